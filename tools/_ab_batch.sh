mkdir -p gpurun_out
L=$PWD/toypathtracer_b200
TPT_LIB_PATH=$L/libtpt_ab_trace.so TPT_TRACE_FILE=gpurun_out/trace_compact.bin timeout 60 python tools/warp_trace.py 2>&1 | tail -2 > gpurun_out/trace_compact.json
cat gpurun_out/trace_compact.json
(timeout 100 python tools/ab_fast.py compact 3 200 1,2 nobig
TPT_LIB_PATH=$L/libtpt_ab_nocompact.so timeout 100 python tools/ab_fast.py nocompact 3 200 1,2 nobig) 2>&1 | grep -v Warn > gpurun_out/ab4.jsonl
cat gpurun_out/ab4.jsonl
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fast.py tests/test_gpu_api.py -q -x -k "fast or sweep or api" 2>&1 | tail -5 > gpurun_out/ab4_tests.log
cat gpurun_out/ab4_tests.log
