"""Where does the end-to-end step time go? (host-buffer draws, 1280x720x4spp fast)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toypathtracer_b200 as tpt
W, H = 1280, 720
ctx = tpt.Context(0)
sph, mats, cam, em = tpt.reference_scene(W, H)
ctx.set_scene(sph, mats, cam, em)
host = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory().numpy()
dev = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
def timeit(fn, n=60):
    for i in range(5): fn(i)
    t0 = time.perf_counter()
    for i in range(n): fn(5 + i)
    return (time.perf_counter() - t0) / n * 1e3
for bands in (1, 2, 3, 4, 6, 8):
    ctx.set_option("host_bands", bands)
    a = timeit(lambda f: ctx.draw(f, 1, W, H, host, flags=0, mode=1))
    b = timeit(lambda f: (ctx.set_scene(sph, mats, cam, em), ctx.draw(f, 1, W, H, host, flags=0, mode=1)))
    c = timeit(lambda f: ctx.draw(f, 1, W, H, host, flags=0, mode=1, want_rays=False))
    print(f"bands {bands}: draw {a:.3f} ms, set_scene+draw {b:.3f} ms, draw w/o ray readback {c:.3f} ms", flush=True)
d = timeit(lambda f: ctx.draw(f, 1, W, H, dev, flags=0, mode=1))
print(f"device buffer draw + ray readback (sync): {d:.3f} ms; kernel ms {ctx.last_kernel_ms():.3f}")
e = timeit(lambda f: ctx.set_scene(sph, mats, cam, em))
print(f"set_scene alone: {e:.3f} ms")
import ctypes
L = tpt._load_lib()
L.tpt_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
ctx.set_option("host_bands", 3); ctx.set_option("host_progress", 1)
for pb in (8, 4, 2):
    ctx.set_option("progress_bands", pb)
    for i in range(3):
        t0 = time.perf_counter(); ctx.draw(100 + i, 1, W, H, host, flags=0, mode=1); t1 = time.perf_counter()
    arr = (ctypes.c_float * 20)()
    n = -L.tpt_debug_timeline(ctx._h, arr, 20)
    print(f"progress_bands {pb}: wall {1e3*(t1-t0):.3f} ms; kernel end {arr[0]:.3f}; band copies end {[round(arr[k],3) for k in range(1,n-1)]}; draw end {arr[n-1]:.3f}")
