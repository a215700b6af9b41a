mkdir -p gpurun_out
timeout 100 python tools/ab_fast.py k2c 3,8 200 1 2>&1 | grep -v Warn > gpurun_out/ab6.jsonl
timeout 200 python tools/fast_probe.py wave 2>&1 | grep stress >> gpurun_out/ab6.jsonl
cat gpurun_out/ab6.jsonl
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fast.py tests/test_gpu_api.py tests/test_gpu_modes.py -q -k "fast or sweep or api or c5 or refgpu" 2>&1 | tail -40 > gpurun_out/ab6_tests.log
tail -15 gpurun_out/ab6_tests.log
