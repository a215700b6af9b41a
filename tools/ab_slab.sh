#!/bin/bash
# A/B harness: build variants with e.g.
#   nvcc ... -DTPT_SLAB_PIX=32 -c tpt_fast.cu -o build/s32/tpt_fast.o; nvcc -shared -o ../libtpt_b200_slab32.so build/tpt_exact.o build/s32/tpt_fast.o build/tpt_api.o
# and list them here; TPT_LIB_PATH selects the library bench.py loads.
for lib in libtpt_b200.so libtpt_b200_slab32.so libtpt_b200_t64.so libtpt_b200_t256.so; do
  TPT_LIB_PATH=$PWD/toypathtracer_b200/$lib timeout 200 python bench.py --no-cpu-baseline --steps 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['e2e']['ms_per_step'])"
done
