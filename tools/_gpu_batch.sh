mkdir -p gpurun_out; timeout -k 5 60 python -c "
import sys; sys.path.insert(0, '.')
import numpy as np, toypathtracer_b200 as tpt
ctx = tpt.Context(0); w,h=192,108
ctx.set_scene(*tpt.reference_scene(w, h))
a=np.zeros((h,w,4),np.float32); b=np.zeros((h,w,4),np.float32)
ctx.set_option('exact_lanes', 32); ra=ctx.draw(0,2,w,h,a,flags=2,mode=0)
ctx.set_option('exact_lanes', 70); rb=ctx.draw(0,2,w,h,b,flags=2,mode=0)
print('cluster small:', ra, rb, (a.view(np.uint32)!=b.view(np.uint32)).sum())
" > gpurun_out/cluster_small.log 2>&1; echo "cluster small rc $?"; tail -3 gpurun_out/cluster_small.log; timeout -k 5 300 python tools/exact_probe.py 3 > gpurun_out/exact_probe7.log 2>&1; echo "probe rc $?"; head -8 gpurun_out/exact_probe7.log; timeout -k 5 200 python tools/fast_probe.py v8 > gpurun_out/fast_probe_v8c.log 2>&1; grep -E 'variant": 4|e2e|"720p x4spp", "w": 1280, "h": 720, "frames": 1, "variant": 3' gpurun_out/fast_probe_v8c.log; timeout -k 5 400 python -m pytest tests/test_gpu_exact.py -m gpu -q -x > gpurun_out/pytest9.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest9.log
