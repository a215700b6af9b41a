"""Developer check run under gpurun: exact-mode parity at full size + first timings of every kernel variant."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import toypathtracer_b200 as tpt
from oracle import pyoracle

def log(*a):
    print(*a, flush=True)

ctx = tpt.Context(0)
w, h = 1280, 720
sph, mats, cam, em = tpt.reference_scene(w, h)
ctx.set_scene(sph, mats, cam, em)

# device libm vs glibc on the path's whole sin/cos domain
k = np.arange(1 << 24, dtype=np.uint32)
a = (k.astype(np.float32) / np.float32(16777216.0)) * np.float32(2.0) * np.float32(3.1415926)
for fn, name in ((0, "sinf"), (1, "cosf")):
    d = ctx.debug_libm(fn, a); g = pyoracle.libm_eval(name, a)
    log(f"libm {name}: mismatches {(d.view(np.uint32) != g.view(np.uint32)).sum()} of {a.size}")
x = np.random.default_rng(1).integers(0, 1 << 32, 1 << 24, dtype=np.uint64).astype(np.uint32).view(np.float32)
d = ctx.debug_libm(2, x); g = pyoracle.libm_eval("powf", x, 5.0)
bad = (d.view(np.uint32) != g.view(np.uint32)) & ~(np.isnan(d) & np.isnan(g))
log(f"libm powf(x,5): mismatches {bad.sum()} of {x.size}")

# exact: frame 0..1 at 1280x720 vs reference, all lane configs
rbuf, rrays = pyoracle.ref_render(w, h, 0, 2, flags=2)
for lanes in (32, 8, 1):
    ctx.set_option("exact_lanes", lanes)
    buf = np.zeros((h, w, 4), np.float32)
    t = time.time()
    r0 = ctx.draw(0, 1, w, h, buf, flags=2, mode=tpt.MODE_EXACT)
    ms0 = ctx.last_kernel_ms()
    r1 = ctx.draw(1, 1, w, h, buf, flags=2, mode=tpt.MODE_EXACT)
    ms1 = ctx.last_kernel_ms()
    nd = int((buf.view(np.uint32) != rbuf.view(np.uint32)).any(axis=2).sum())
    log(f"exact lanes={lanes}: rays {[r0, r1]} ref {rrays} equal={[r0, r1] == rrays} pixels differing={nd} kernel ms {ms0:.2f} {ms1:.2f} -> {r1/ms1/1e3:.1f} Mray/s")

# exact batched: 16 frames in one call vs reference progressive
rbuf16, rrays16 = pyoracle.ref_render(w, h, 0, 16, flags=2)
for lanes in (0, 1, 2, 8, 9, 32):
    ctx.set_option("exact_lanes", lanes)
    buf = np.zeros((h, w, 4), np.float32)
    tot, pf = ctx.draw(0, 16, w, h, buf, flags=2, mode=tpt.MODE_EXACT, per_frame=True)
    ms = ctx.last_kernel_ms()
    nd = int((buf.view(np.uint32) != rbuf16.view(np.uint32)).any(axis=2).sum())
    log(f"exact batch16 lanes={lanes}: rays equal={pf == rrays16} pixels differing={nd} kernel ms {ms:.2f} -> {tot/ms/1e3:.1f} Mray/s")
ctx.set_option("exact_lanes", 0)

# fast variants
import torch
dbuf = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
for var in (0, 2, 3, 5, 7):
    ctx.set_option("fast_variant", var)
    for rep in range(3):
        rays = ctx.draw(rep, 1, w, h, dbuf, flags=0, mode=tpt.MODE_FAST)
        ms = ctx.last_kernel_ms()
    img = dbuf.cpu().numpy()
    log(f"fast variant={var}: rays {rays} rays/sample {rays/(w*h*4):.4f} kernel ms {ms:.3f} -> {rays/ms/1e3:.1f} Mray/s  mean {img[...,:3].mean(axis=(0,1))} ref mean {rbuf[...,:3].mean(axis=(0,1))}")
# fast 64 spp accumulate vs ref 64 spp
for var in (3, 5, 6, 7):
    ctx.set_option("fast_variant", var)
    dbuf.zero_()
    rays = ctx.draw(0, 16, w, h, dbuf, flags=2, mode=tpt.MODE_FAST)
    ms = ctx.last_kernel_ms()
    img = dbuf.cpu().numpy()[..., :3].astype(np.float64); ref = rbuf16[..., :3].astype(np.float64)
    rel = np.sqrt(((img - ref) ** 2).sum() / (ref ** 2).sum())
    log(f"fast variant={var} 64spp: rays {rays} ({rays/(w*h*64):.4f}/sample, ref {sum(rrays16)/(w*h*64):.4f}) ms {ms:.2f} -> {rays/ms/1e3:.1f} Mray/s relL2 vs ref64 {rel:.4e} (independent-render floor ~ {0.194/8:.4e})")
log("done")
