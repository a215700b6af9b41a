"""Times the exact mode's kernels (exact_lanes settings) on the shapes that matter: one 720p frame (the drop-in's
default call), one 4K frame, 16 frames of 720p. Device buffers, CUDA events. Prints one JSON line per case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toypathtracer_b200 as tpt

ctx = tpt.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sh = stream.cuda_stream
cases = [(1280, 720, 1, (32, 64, 65, 66, 69, 70, 71)), (3840, 2160, 1, (32, 65)), (1280, 720, 16, (32, 8, 65)), (1280, 720, 256, (1,))]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for (w, h, nf, lanes_list) in cases:
    ctx.set_scene(*tpt.reference_scene(w, h))
    img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    for lanes in lanes_list:
        ctx.set_option("exact_lanes", lanes)
        ctx.draw(0, nf, w, h, img, flags=0, mode=0, stream=sh, want_rays=False)
        ctx.read_ray_count(sh)
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for r in range(reps):
            ctx.draw(1 + r * nf, nf, w, h, img, flags=0, mode=0, stream=sh, want_rays=False)
        e1.record(stream)
        torch.cuda.synchronize()
        rays = ctx.read_ray_count(sh)
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"w": w, "h": h, "frames": nf, "exact_lanes": lanes, "ms": ms, "mray_s": rays / reps / ms / 1e3}), flush=True)

# the reference-GPU-compatible mode (per-pixel streams): strict and native arithmetic, 720p x 4 spp, flags = 0
w, h = 1280, 720
ctx.set_option("exact_lanes", 0)
ctx.set_scene(*tpt.reference_scene(w, h))
img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
for mode, name in ((tpt.MODE_REFGPU, "refgpu"), (tpt.MODE_REFGPU_FAST, "refgpu_fast"), (tpt.MODE_FAST, "fast")):
    ctx.draw(0, 1, w, h, img, flags=0, mode=mode, stream=sh, want_rays=False)
    ctx.read_ray_count(sh)
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r in range(reps):
        ctx.draw(1 + r, 1, w, h, img, flags=0, mode=mode, stream=sh, want_rays=False)
    e1.record(stream)
    torch.cuda.synchronize()
    rays = ctx.read_ray_count(sh)
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"w": w, "h": h, "mode": name, "ms": ms, "mray_s": rays / reps / ms / 1e3}), flush=True)
