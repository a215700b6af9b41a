mkdir -p gpurun_out
L=$PWD/toypathtracer_b200
(timeout 200 python tools/fast_probe.py wave 2>&1 | grep "stress" | grep '"variant": 3'
TPT_LIB_PATH=$L/libtpt_ab_big1024.so timeout 200 python tools/fast_probe.py wave 2>&1 | grep "stress" | grep '"variant": 3' | sed 's/"case": "/"case": "1024thr /') > gpurun_out/ab10.jsonl
cat gpurun_out/ab10.jsonl
timeout 600 python -m pytest tests/test_gpu_fast.py -q 2>&1 | tail -3
