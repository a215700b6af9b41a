mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fast.py -m gpu -x -q -rf > gpurun_out/pytest14.log 2>&1; tail -8 gpurun_out/pytest14.log
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:k_fast_queue -c 1 -f -o gpurun_out/stress_k2c python tools/prof_run.py fast 3 1920 1080 2 1 4096 > gpurun_out/ncu_stress.log 2>&1; tail -3 gpurun_out/ncu_stress.log
