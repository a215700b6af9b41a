"""TEST INFRASTRUCTURE — ctypes access to the CPU oracles. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product (toypathtracer_b200/) never does.

  ref_*   oracle/_ref/libtoyref.so  — the UNMODIFIED reference C++ path (Test.cpp/Maths.cpp/enkiTS) behind a
                                      C-ABI harness (oracle/ref_harness.cpp), built by oracle/Makefile from
                                      /root/reference where it lies; prebuilt file travels to the GPU box.
  orc_*   oracle/_ref/liboracle.so  — the CPU restatement with a runtime scene (oracle/restate.cpp), pinned
                                      bitwise to the reference on its 46-sphere scene (tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_DIR, "_ref", "libtoyref.so")
ORC_SO = os.path.join(_DIR, "_ref", "liboracle.so")
RGO_SO = os.path.join(_DIR, "_ref", "librefgpu.so")
_ref = {}
_orc = None
_rgo = None


def build(quiet: bool = True):
    """Compiles the oracles (restatement always; the reference only where /root/reference exists)."""
    subprocess.run(["make", "-C", _DIR, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.STDOUT if quiet else None)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ref_so(variant: str = "") -> str:
    """variant: "" (stock), "mitsuba" (DO_MITSUBA_COMPARE 1), "small" (DO_BIG_SCENE 0), "small_mitsuba" — the
    reference compiled from an edited temporary copy of its sources by oracle/Makefile."""
    return os.path.join(_DIR, "_ref", f"libtoyref{'_' + variant if variant else ''}.so")


def have_ref_variant(variant: str) -> bool:
    return os.path.exists(ref_so(variant))


def ref_lib(variant: str = ""):
    if variant not in _ref:
        path = ref_so(variant)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run oracle/Makefile where /root/reference is mounted")
        L = ctypes.CDLL(path)             # RTLD_LOCAL: every variant keeps its own statics (scene, scheduler)
        L.ref_render.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ref_render.restype = ctypes.c_int
        _ref[variant] = L
    return _ref[variant]


def orc_lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORC_SO):
            build()
        L = ctypes.CDLL(ORC_SO)
        L.orc_render.restype = ctypes.c_int
        L.orc_render_rows.restype = ctypes.c_int
        _orc = L
    return _orc


def ref_render(w, h, frame0, nframes, flags=0, time=0.0, buf=None, want_seconds=False, variant=""):
    """Reference shells' loop: UpdateTest + DrawTest per frame into `buf` (zeros if None).
    Returns (buf[h,w,4] float32, rays per frame[, seconds per frame])."""
    L = ref_lib(variant)
    if buf is None:
        buf = np.zeros((h, w, 4), np.float32)
    rays = (ctypes.c_longlong * nframes)()
    secs = (ctypes.c_double * nframes)()
    L.ref_render(w, h, frame0, nframes, time, flags, _vp(buf), rays, secs)
    out = (buf, [int(r) for r in rays])
    return out + ([float(s) for s in secs],) if want_seconds else out


def ref_scene(w, h, time=0.0, flags=0, variant=""):
    """Raw scene through the reference's own GetObjectCount/GetSceneDesc: (spheres[n,5], mats[n,9] raw f32 view,
    cam[22], emissive ids)."""
    L = ref_lib(variant)
    cnt = (ctypes.c_int * 4)()
    L.ref_object_count(ctypes.byref(cnt, 0), ctypes.byref(cnt, 4), ctypes.byref(cnt, 8), ctypes.byref(cnt, 12))
    n = cnt[0]
    assert (cnt[1], cnt[2], cnt[3]) == (20, 36, 88)
    sph = np.zeros((n, 5), np.float32); mats = np.zeros((n, 9), np.float32); cam = np.zeros(22, np.float32)
    em = np.zeros(n, np.int32); ec = ctypes.c_int()
    L.ref_scene_desc(ctypes.c_float(time), 0, w, h, flags, _vp(sph), _vp(mats), _vp(cam), _vp(em), ctypes.byref(ec))
    return sph, mats, cam, em[: ec.value].copy()


def orc_render(spheres, mats, cam, w, h, frame0, nframes, flags=0, spp=4, simd_tie=1, buf=None, nthreads=0,
               want_seconds=False, rows=None, mitsuba=False):
    """CPU restatement on an arbitrary scene. Returns (buf, rays per frame, pad pixel list [(x,y,frame)...]
    [, seconds per frame]). rows = (row0, numRows, rowStep): only those rows are traced (into their place in the
    full-size buf) and counted."""
    L = orc_lib()
    spheres = np.ascontiguousarray(spheres); mats = np.ascontiguousarray(mats); cam = np.ascontiguousarray(cam)
    n = spheres.nbytes // 20
    assert spheres.nbytes == n * 20 and mats.nbytes == n * 36 and cam.nbytes == 88
    if buf is None:
        buf = np.zeros((h, w, 4), np.float32)
    rays = (ctypes.c_longlong * nframes)()
    secs = (ctypes.c_double * nframes)()
    pad = ctypes.c_longlong(0)
    cap = 4096
    padxy = np.zeros((cap, 3), np.int32)
    L.orc_set_mitsuba(1 if mitsuba else 0)
    row0, nrows, rstep = rows if rows is not None else (0, h, 1)
    assert nrows >= 0 and rstep >= 1 and row0 >= 0 and (nrows == 0 or row0 + (nrows - 1) * rstep < h)
    L.orc_render_rows(_vp(spheres), _vp(mats), n, _vp(cam), w, h, frame0, nframes, ctypes.c_uint(flags), spp, simd_tie,
                      _vp(buf), rays, ctypes.byref(pad), secs, nthreads, _vp(padxy), cap, row0, nrows, rstep)
    pads = [tuple(int(v) for v in p) for p in padxy[: min(pad.value, cap)]]
    out = (buf, [int(r) for r in rays], pads)
    return out + ([float(s) for s in secs],) if want_seconds else out


def rgo_render(spheres, mats, cam, w, h, frame0, nframes, flags=0, spp=4, mitsuba=False, buf=None, nthreads=0):
    """CPU restatement of the reference's GPU compute shader (oracle/refgpu_restate.cpp): (buf, rays per frame)."""
    global _rgo
    if _rgo is None:
        if not os.path.exists(RGO_SO):
            build()
        _rgo = ctypes.CDLL(RGO_SO)
        _rgo.rgo_render.restype = ctypes.c_int
    spheres = np.ascontiguousarray(spheres); mats = np.ascontiguousarray(mats); cam = np.ascontiguousarray(cam)
    n = spheres.nbytes // 20
    assert spheres.nbytes == n * 20 and mats.nbytes == n * 36 and cam.nbytes == 88
    if buf is None:
        buf = np.zeros((h, w, 4), np.float32)
    rays = (ctypes.c_longlong * nframes)()
    _rgo.rgo_render(_vp(spheres), _vp(mats), n, _vp(cam), w, h, frame0, nframes, ctypes.c_uint(flags), spp,
                    1 if mitsuba else 0, _vp(buf), rays, nthreads)
    return buf, [int(r) for r in rays]


def libm_eval(fn, x, y=None):
    """The platform libm the oracles link (glibc): fn 'sinf' | 'cosf' | 'powf'."""
    L = orc_lib()
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    if fn == "powf":
        y = np.ascontiguousarray(np.broadcast_to(np.asarray(y, np.float32), x.shape))
        L.orc_powf(_vp(x), _vp(y), _vp(out), ctypes.c_longlong(x.size))
    else:
        getattr(L, "orc_" + fn)(_vp(x), _vp(out), ctypes.c_longlong(x.size))
    return out


# The reference keeps animated sphere positions in its static scene for the rest of the process
# (Test.cpp:304-308), so anything that passes kFlagAnimate runs in a fresh interpreter.
def _isolated_worker(q, fn, args, kwargs):
    q.put(globals()[fn](*args, **kwargs))


def isolated(fn: str, *args, **kwargs):
    """Runs ref_render / ref_scene in a spawned process and returns its result."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_isolated_worker, args=(q, fn, args, kwargs))
    p.start()
    out = q.get(timeout=600)
    p.join()
    return out
