"""Times the fast-mode kernels: sweep forms (fast_kform 0/1/2) x variants on the BASELINE shapes. Device buffers, CUDA
events per draw, L2 flushed between draws. One JSON line per case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toypathtracer_b200 as tpt

ctx = tpt.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sh = stream.cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(label, w, h, nf, variant, kform, scene, reps, flags=0):
    ctx.set_scene(*scene)
    ctx.set_option("fast_variant", variant); ctx.set_option("fast_kform", kform)
    img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    for r in range(2):
        ctx.draw(r * nf, nf, w, h, img, flags=flags, mode=1, stream=sh, want_rays=False)
    ctx.read_ray_count(sh)
    tot = 0.0
    for r in range(reps):
        flush.fill_(r & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.draw((2 + r) * nf, nf, w, h, img, flags=flags, mode=1, stream=sh, want_rays=False)
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    rays = ctx.read_ray_count(sh)
    print(json.dumps({"case": label, "w": w, "h": h, "frames": nf, "variant": variant, "kform": kform, "ms": tot / reps,
                      "mray_s": rays / tot / 1e3}), flush=True)


ref720 = tpt.reference_scene(1280, 720)
ref4k = tpt.reference_scene(3840, 2160)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "kform"):
    for kform in (0, 1, 2):
        run("720p x4spp", 1280, 720, 1, 3, kform, ref720, 50)
    for kform in (1, 2):
        run("4K x64spp", 3840, 2160, 16, 3, kform, ref4k, 3, flags=2)
        run("4K x64spp", 3840, 2160, 16, 5, kform, ref4k, 3, flags=2)
        run("720p x4spp", 1280, 720, 1, 5, kform, ref720, 20)
if which in ("all", "v8"):
    for variant in (3, 4, 8):
        run("720p x4spp", 1280, 720, 1, variant, 2, ref720, 50)
        run("720p x4spp progressive", 1280, 720, 1, variant, 2, ref720, 50, flags=2)
    for variant in (3, 8):
        run("4K x64spp", 3840, 2160, 16, variant, 2, ref4k, 3, flags=2)
        run("720p x64spp", 1280, 720, 16, variant, 2, ref720, 5, flags=2)
    # end to end with a pinned host buffer (what bench.py's e2e does): set_scene + draw, wall clock
    import time
    host = torch.zeros((720, 1280, 4), dtype=torch.float32).pin_memory().numpy()
    ctx.set_scene(*ref720); ctx.set_option("fast_kform", 2)
    for variant in (3, 8):
        ctx.set_option("fast_variant", variant)
        for i in range(5):
            ctx.set_scene(*ref720); ctx.draw(i, 1, 1280, 720, host, flags=0, mode=1)
        t0 = time.perf_counter(); rays = 0
        for i in range(100):
            ctx.set_scene(*ref720); rays += ctx.draw(5 + i, 1, 1280, 720, host, flags=0, mode=1)
        dt = time.perf_counter() - t0
        print(json.dumps({"case": "e2e host pinned 720p x4spp", "variant": variant, "ms": dt * 10, "mray_s": rays / dt / 1e6}), flush=True)
if which in ("all", "wave"):
    # the material-sorted block wavefront (variants 6/7) against the queue kernel where sorting should pay most
    for variant in (3, 6, 7):
        run("720p x64spp", 1280, 720, 16, variant, 2, ref720, 3, flags=2)
    stress = tpt.stress_scene(1920, 1080, count=4096)
    sc = (stress[0], stress[1], stress[2], None)
    for variant in (3, 6, 7):
        run("stress4096 1080p x8spp", 1920, 1080, 2, variant, 2, sc, 2, flags=2)
