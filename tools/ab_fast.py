"""A/B timing of one build of the fast kernel (TPT_LIB_PATH selects the library): the bench configuration (46 spheres,
1280x720, 4 spp, device buffer, L2 flushed between draws, CUDA events) for the variants given, plus 4K x 16 spp.
    TPT_LIB_PATH=... python tools/ab_fast.py <label> [variants=3] [reps=200]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toypathtracer_b200 as tpt

label = sys.argv[1]
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "3").split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
ctx = tpt.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sh = stream.cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(case, w, h, nf, variant, reps, flags=0):
    ctx.set_scene(*tpt.reference_scene(w, h))
    ctx.set_option("fast_variant", variant); ctx.set_option("fast_kform", 2)
    img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    for r in range(3):
        ctx.draw(r * nf, nf, w, h, img, flags=flags, mode=1, stream=sh, want_rays=False)
    ctx.read_ray_count(sh)
    ms = []
    for r in range(reps):
        flush.fill_(r & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.draw((3 + r) * nf, nf, w, h, img, flags=flags, mode=1, stream=sh, want_rays=False)
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    rays = ctx.read_ray_count(sh)
    tot = sum(ms)
    print(json.dumps({"build": label, "case": case, "variant": variant, "ms": tot / reps, "ms_min": min(ms),
                      "mray_s": rays / tot / 1e3}), flush=True)


frames = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "1").split(",")]
for v in variants:
    for nf in frames:
        run("720p x%dspp" % (4 * nf), 1280, 720, nf, v, max(reps // nf, 3), flags=0 if nf == 1 else 2)
if len(sys.argv) <= 5:
    run("4K x16spp", 3840, 2160, 4, variants[0], max(reps // 20, 3), flags=2)
