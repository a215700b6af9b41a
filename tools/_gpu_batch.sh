mkdir -p gpurun_out
timeout -k 5 200 python tools/exact_probe.py 1 > gpurun_out/exact_probe9.log 2>&1; echo "probe rc $?"; head -5 gpurun_out/exact_probe9.log
