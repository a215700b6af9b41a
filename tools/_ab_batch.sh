mkdir -p gpurun_out
L=$PWD/toypathtracer_b200
TPT_LIB_PATH=$L/libtpt_ab_trace.so TPT_TRACE_FILE=gpurun_out/trace_split.bin timeout 60 python tools/warp_trace.py 2>&1 | tail -1 > gpurun_out/trace_split.json
cat gpurun_out/trace_split.json
(for i in 1 2; do timeout 100 python tools/ab_fast.py split2 3 300 1 nobig
TPT_LIB_PATH=$L/libtpt_ab_nosplit.so timeout 100 python tools/ab_fast.py nosplit 3 300 1 nobig; done
timeout 100 python tools/ab_fast.py split2 3 100 2,16 nobig) 2>&1 | grep -v Warn > gpurun_out/ab9.jsonl
cat gpurun_out/ab9.jsonl
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_api.py tests/test_gpu_configs.py -q -k "fast or api or c5" 2>&1 | tail -5 > gpurun_out/ab9_tests.log
tail -3 gpurun_out/ab9_tests.log
