#!/bin/bash
# A/B: slab size of the fast queue kernel (experiment builds libtpt_b200_slab{64,256}.so)
for lib in libtpt_b200.so libtpt_b200_slab32.so; do
  TPT_LIB_PATH=$PWD/toypathtracer_b200/$lib timeout 200 python bench.py --no-cpu-baseline --steps 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['e2e']['ms_per_step'])"
done
