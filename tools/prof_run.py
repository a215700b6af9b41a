"""Tiny driver for ncu captures: python tools/prof_run.py <mode fast|exact> <variant/lanes> <w> <h> <nframes> <reps> [scene]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import toypathtracer_b200 as tpt

mode, var, w, h, nf, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
scene = sys.argv[7] if len(sys.argv) > 7 else "ref"
ctx = tpt.Context(0)
if scene == "ref":
    sph, mats, cam, em = tpt.reference_scene(w, h)
else:
    sph, mats, cam, em = tpt.stress_scene(w, h, int(scene))
ctx.set_scene(sph, mats, cam, em)
buf = ctx.mem_alloc(w * h * 16)          # device-resident image: one kernel launch per draw (no host-band split)
if mode == "fast":
    ctx.set_option("fast_variant", var)
    m = tpt.MODE_FAST
    if len(sys.argv) > 8:
        ctx.set_option("fast_kform", int(sys.argv[8]))
else:
    ctx.set_option("exact_lanes", var)
    m = tpt.MODE_EXACT
for r in range(reps):
    rays = ctx.draw(r * nf, nf, w, h, buf, flags=0 if nf == 1 else 2, mode=m)
    print(r, rays, ctx.last_kernel_ms(), "ms", rays / ctx.last_kernel_ms() / 1e3, "Mray/s", flush=True)
