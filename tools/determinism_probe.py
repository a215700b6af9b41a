"""Are ray counts reproducible per sweep form? (same per-path RNG streams: they must be, run to run)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import toypathtracer_b200 as tpt
ctx = tpt.Context(0)
ctx.set_option('fast_variant', 3)
for name, scene, (w, h, nf) in (("stress", None, (1920, 1080, 2)), ("ref", None, (1280, 720, 4))):
    if name == "stress":
        s = tpt.stress_scene(w, h, count=4096); ctx.set_scene(s[0], s[1], s[2], None)
    else:
        ctx.set_scene(*tpt.reference_scene(w, h))
    buf = ctx.mem_alloc(w * h * 16)
    for kform in (0, 1, 2):
        ctx.set_option("fast_kform", kform)
        print(name, "kform", kform, [ctx.draw(0, nf, w, h, buf, flags=2, mode=1) for _ in range(3)], flush=True)
    ctx.mem_free(buf)
