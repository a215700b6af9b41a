mkdir -p gpurun_out
(TPT_LIB_PATH=$PWD/toypathtracer_b200/libtpt_b200.so timeout 120 python tools/ab_fast.py h16p2 3,9,8 200 2>&1 | tail -5
TPT_LIB_PATH=$PWD/toypathtracer_b200/libtpt_ab_h32p2.so timeout 120 python tools/ab_fast.py h32p2 3,9 200 2>&1 | tail -4) > gpurun_out/ab2.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fast.py tests/test_gpu_modes.py -q -x -k "fast or refgpu or sweep" 2>&1 | tail -5 > gpurun_out/ab2_tests.log
cat gpurun_out/ab2.jsonl gpurun_out/ab2_tests.log
