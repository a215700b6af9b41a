mkdir -p gpurun_out
timeout -k 5 200 python tools/exact_probe.py 1 > gpurun_out/exact_probe10.log 2>&1; echo "probe rc $?"; head -12 gpurun_out/exact_probe10.log
timeout -k 5 300 python -m pytest tests/test_gpu_modes.py -m gpu -q -x > gpurun_out/pytest11.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest11.log
