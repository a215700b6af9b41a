"""Small draws of every new kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import toypathtracer_b200 as tpt

ctx = tpt.Context(0)
w, h = 96, 54
ctx.set_scene(*tpt.reference_scene(w, h))
buf = ctx.mem_alloc(w * h * 16)
for lanes in (65, 66, 32):
    ctx.set_option("exact_lanes", lanes)
    print("exact", lanes, ctx.draw(0, 1, w, h, buf, flags=2, mode=tpt.MODE_EXACT), flush=True)
ctx.set_option("exact_lanes", 0)
for variant, nf in ((8, 1), (8, 3), (3, 1), (5, 2)):
    ctx.set_option("fast_variant", variant)
    print("fast", variant, nf, ctx.draw(0, nf, w, h, buf, flags=2, mode=tpt.MODE_FAST), flush=True)
for mode in (tpt.MODE_REFGPU, tpt.MODE_REFGPU_FAST):
    print("refgpu", mode, ctx.draw(1, 2, w, h, buf, flags=2, mode=mode), flush=True)
ctx.set_option("exact_lookahead", 3)
host = np.zeros((h, w, 4), np.float32)
for f in range(4):
    print("lookahead", f, ctx.draw(f, 1, w, h, host, flags=2, mode=tpt.MODE_EXACT), flush=True)
