mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_full.log 2>&1; tail -6 gpurun_out/pytest_full.log
timeout -k 5 600 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -k 5 300 python tools/fast_probe.py wave > gpurun_out/fast_probe_wave.log 2>&1
