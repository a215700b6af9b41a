mkdir -p gpurun_out
L=$PWD/toypathtracer_b200
(timeout 100 python tools/ab_fast.py match 8 300 1,16 nobig
for n in cas go12 go16; do TPT_LIB_PATH=$L/libtpt_ab_$n.so timeout 100 python tools/ab_fast.py $n 8 300 1,16 nobig; done) 2>&1 | grep -v Warn > gpurun_out/ab7.jsonl
cat gpurun_out/ab7.jsonl
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_api.py tests/test_gpu_configs.py -q -k "fast or api" 2>&1 | tail -5 > gpurun_out/ab7_tests.log
tail -5 gpurun_out/ab7_tests.log
