mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_full3.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_full3.log
timeout -k 5 600 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_r02c.log 2> gpurun_out/bench_r02c.err; echo "bench rc $?"; cut -c1-160 gpurun_out/bench_r02c.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/smoke_r02.log
timeout -k 5 300 python tools/exact_probe.py > gpurun_out/exact_probe_final.log 2>&1; echo "probe rc $?"; cat gpurun_out/exact_probe_final.log
