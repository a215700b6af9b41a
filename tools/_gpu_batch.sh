# round-end evidence run (one GPU): full GPU test suite, bench line, ncu captures of the shipped fast kernels, launch list
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_full.log 2>&1; tail -4 gpurun_out/pytest_full.log
timeout -k 5 600 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:k_fast_queue -s 1 -c 1 -o gpurun_out/prof_fastq python tools/prof_run.py fast 3 1280 720 1 3 > gpurun_out/prof_fastq.log 2>&1
timeout 300 $NCU -k regex:k_fast_group -s 1 -c 1 -o gpurun_out/prof_group4k python tools/prof_run.py fast 8 3840 2160 1 3 > gpurun_out/prof_group4k.log 2>&1
timeout 300 $NCU -k regex:k_fast_queue -s 1 -c 1 -o gpurun_out/prof_stress python tools/prof_run.py fast 3 1920 1080 2 2 4096 > gpurun_out/prof_stress.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12
