"""toypathtracer_b200 — host-side mirror of the reference's renderer API (Cpp/Source/Test.h:10-17) on top of
the B200 CUDA library (include/tpt_b200.h).

Two layers, both thin ctypes bindings over in-tree shared libraries built by `csrc/Makefile`:

* the six functions of the reference's ``Test.h`` — ``InitializeTest, ShutdownTest, UpdateTest, DrawTest,
  GetObjectCount, GetSceneDesc`` — bound to the *C++-mangled* symbols of ``libtoytest_b200.so``, i.e. exactly
  what a reference shell (Cpp/Windows/TestWin.cpp:76,258,265,315-316) would link against;
* :class:`Context`, the C-ABI itself (``tpt_create/tpt_set_scene/tpt_draw/...``) for runtime scenes, frame
  batching, row sharding across GPUs and device-resident buffers.

There is no CPU rendering path: importing works anywhere (so that CPU-only tests can check symbols), but any
draw without the compiled library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TPT_LIB_PATH") or os.path.join(_HERE, "libtpt_b200.so")   # override: A/B builds in experiments
SHIM_PATH = os.path.join(_HERE, "libtoytest_b200.so")

MODE_EXACT = 0
MODE_FAST = 1
MODE_REFGPU = 2        # the reference's GPU-shader estimator (ComputeShader.hlsl), strict arithmetic
MODE_REFGPU_FAST = 3   # the same with GPU-native arithmetic
kFlagAnimate = 1      # Cpp/Source/Test.h:6
kFlagProgressive = 2  # Cpp/Source/Test.h:7

SPHERE_DTYPE = np.dtype([("center", np.float32, 3), ("radius", np.float32), ("invRadius", np.float32)])  # 20 B, Maths.h:354-364
MATERIAL_DTYPE = np.dtype([("type", np.int32), ("albedo", np.float32, 3), ("emissive", np.float32, 3),
                           ("roughness", np.float32), ("ri", np.float32)])                                # 36 B, Test.cpp:36-44
CAMERA_DTYPE = np.dtype([("origin", np.float32, 3), ("lowerLeftCorner", np.float32, 3), ("horizontal", np.float32, 3),
                         ("vertical", np.float32, 3), ("uu", np.float32, 3), ("vv", np.float32, 3), ("ww", np.float32, 3),
                         ("lensRadius", np.float32)])                                                     # 88 B, Maths.h:444-449
assert SPHERE_DTYPE.itemsize == 20 and MATERIAL_DTYPE.itemsize == 36 and CAMERA_DTYPE.itemsize == 88


class TptError(RuntimeError):
    pass


_lib = None
_shim = None


def _load_lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TptError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(or `make -C toypathtracer_b200/csrc`). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    vp, ci, cu, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_longlong
    L.tpt_create.argtypes = [ci, ctypes.POINTER(vp)]; L.tpt_create.restype = ci
    L.tpt_destroy.argtypes = [vp]; L.tpt_destroy.restype = None
    L.tpt_device_count.argtypes = []; L.tpt_device_count.restype = ci
    L.tpt_last_error.argtypes = [vp]; L.tpt_last_error.restype = ctypes.c_char_p
    L.tpt_set_scene.argtypes = [vp, vp, vp, ci, vp, vp, ci]; L.tpt_set_scene.restype = ci
    L.tpt_set_camera.argtypes = [vp, vp]; L.tpt_set_camera.restype = ci
    L.tpt_set_spp.argtypes = [vp, ci]; L.tpt_set_spp.restype = ci
    L.tpt_set_option.argtypes = [vp, ctypes.c_char_p, ci]; L.tpt_set_option.restype = ci
    L.tpt_draw.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, ci, cu, ci, ctypes.POINTER(cll), ctypes.POINTER(cll), vp]
    L.tpt_draw.restype = ci
    L.tpt_read_ray_count.argtypes = [vp, vp, ctypes.POINTER(cll)]; L.tpt_read_ray_count.restype = ci
    L.tpt_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]; L.tpt_last_kernel_ms.restype = ci
    L.tpt_last_launch_count.argtypes = [vp]; L.tpt_last_launch_count.restype = ci
    L.tpt_last_scene_upload_bytes.argtypes = [vp]; L.tpt_last_scene_upload_bytes.restype = ctypes.c_longlong
    L.tpt_tonemap_srgb8.argtypes = [vp, vp, ci, ci, ci, vp, ci, vp]; L.tpt_tonemap_srgb8.restype = ci
    L.tpt_tonemap_rgba8.argtypes = [vp, vp, ci, ci, ci, vp, ci, ci, ci, ci, vp]; L.tpt_tonemap_rgba8.restype = ci
    L.tpt_debug_libm.argtypes = [vp, ci, vp, vp, cll]; L.tpt_debug_libm.restype = ci
    L.tpt_debug_hit.argtypes = [vp, ci, vp, vp, vp, cll]; L.tpt_debug_hit.restype = ci
    cull = ctypes.c_ulonglong
    L.tpt_mem_alloc.argtypes = [vp, cull, ctypes.POINTER(vp)]; L.tpt_mem_alloc.restype = ci
    L.tpt_mem_free.argtypes = [vp, vp]; L.tpt_mem_free.restype = ci
    L.tpt_mem_copy.argtypes = [vp, vp, vp, cull, ci]; L.tpt_mem_copy.restype = ci
    L.tpt_ipc_export.argtypes = [vp, vp, vp]; L.tpt_ipc_export.restype = ci
    L.tpt_ipc_open.argtypes = [vp, vp, ctypes.POINTER(vp)]; L.tpt_ipc_open.restype = ci
    L.tpt_ipc_close.argtypes = [vp, vp]; L.tpt_ipc_close.restype = ci
    _lib = L
    return L


def device_count() -> int:
    return int(_load_lib().tpt_device_count())


class DevicePtr:
    """A raw device address (e.g. an image living in another rank's HBM, opened with Context.ipc_open)."""

    def __init__(self, addr: int):
        self.addr = int(addr)


def _as_ptr(buf) -> Tuple[int, bool, object]:
    """(address, on_device, keepalive) for a numpy array, a torch tensor or a raw device address."""
    if isinstance(buf, DevicePtr):
        return buf.addr, True, buf
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.float32 or not buf.flags["C_CONTIGUOUS"]:
            raise TptError("backbuffer must be a C-contiguous float32 array")
        return buf.ctypes.data, False, buf
    if hasattr(buf, "data_ptr"):  # torch tensor
        if not buf.is_contiguous() or str(buf.dtype) != "torch.float32":
            raise TptError("backbuffer tensor must be contiguous float32")
        return int(buf.data_ptr()), bool(buf.is_cuda), buf
    raise TptError("unsupported backbuffer type")


class Context:
    """One CUDA device's renderer state (scene blob, scratch, counters): ``tpt_context`` of include/tpt_b200.h."""

    def __init__(self, device: int = 0):
        self._L = _load_lib()
        h = ctypes.c_void_p()
        rc = self._L.tpt_create(device, ctypes.byref(h))
        if rc != 0 or not h.value:
            raise TptError(f"tpt_create(device={device}) failed with CUDA error {rc}: no usable CUDA device "
                           "(this package has no CPU path)")
        self._h = h
        self.device = device
        self.count = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.tpt_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise TptError(f"{what} failed ({rc}): {self._L.tpt_last_error(self._h).decode()}")

    def set_scene(self, spheres: np.ndarray, materials: np.ndarray, camera: np.ndarray,
                  emissives: Optional[np.ndarray] = None):
        """Raw scene as exported by GetSceneDesc (Test.cpp:377-384)."""
        spheres = np.ascontiguousarray(spheres); materials = np.ascontiguousarray(materials)
        camera = np.ascontiguousarray(camera)
        n = spheres.nbytes // 20
        if spheres.nbytes != n * 20 or materials.nbytes != n * 36 or camera.nbytes != 88:
            raise TptError("scene arrays must be n*20 B spheres, n*36 B materials, 88 B camera")
        if emissives is not None:
            emissives = np.ascontiguousarray(emissives, dtype=np.int32)
            ep, ec = emissives.ctypes.data, int(emissives.size)
        else:
            ep, ec = None, 0
        self._check(self._L.tpt_set_scene(self._h, spheres.ctypes.data, materials.ctypes.data, n,
                                          camera.ctypes.data, ep, ec), "tpt_set_scene")
        self.count = n

    def set_camera(self, camera: np.ndarray):
        camera = np.ascontiguousarray(camera)
        self._check(self._L.tpt_set_camera(self._h, camera.ctypes.data), "tpt_set_camera")

    def set_spp(self, spp: int):
        self._check(self._L.tpt_set_spp(self._h, spp), "tpt_set_spp")

    def set_option(self, key: str, value: int):
        self._check(self._L.tpt_set_option(self._h, key.encode(), int(value)), f"tpt_set_option({key})")

    def draw(self, frame: int, num_frames: int, width: int, height: int, backbuffer, flags: int = 0,
             mode: int = MODE_EXACT, rows: Optional[Tuple[int, int, int, int]] = None, stream: int = 0,
             want_rays: bool = True, per_frame: bool = False):
        """tpt_draw. rows = (row0, numRows, rowStep, packed); default = the whole image.
        Returns total rays (int) if want_rays, plus the per-frame list if per_frame (exact mode)."""
        row0, nrows, step, packed = rows if rows is not None else (0, height, 1, 0)
        addr, on_dev, _keep = _as_ptr(backbuffer)
        total = ctypes.c_longlong(0)
        pf = (ctypes.c_longlong * num_frames)() if per_frame else None
        rc = self._L.tpt_draw(self._h, frame, num_frames, width, height, row0, nrows, step, packed,
                              ctypes.c_void_p(addr), 1 if on_dev else 0, flags, mode,
                              ctypes.byref(total) if want_rays else None,
                              pf if per_frame else None, ctypes.c_void_p(stream) if stream else None)
        self._check(rc, "tpt_draw")
        if per_frame:
            return int(total.value), [int(v) for v in pf]
        return int(total.value) if want_rays else None

    def read_ray_count(self, stream: int = 0) -> int:
        out = ctypes.c_longlong(0)
        self._check(self._L.tpt_read_ray_count(self._h, ctypes.c_void_p(stream) if stream else None, ctypes.byref(out)),
                    "tpt_read_ray_count")
        return int(out.value)

    def last_kernel_ms(self) -> float:
        out = ctypes.c_float(0)
        self._check(self._L.tpt_last_kernel_ms(self._h, ctypes.byref(out)), "tpt_last_kernel_ms")
        return float(out.value)

    def last_launch_count(self) -> int:
        return int(self._L.tpt_last_launch_count(self._h))

    def last_scene_upload_bytes(self) -> int:
        """Bytes the most recent set_scene copied to the device (0: the scene was already resident)."""
        return int(self._L.tpt_last_scene_upload_bytes(self._h))

    # ---- device memory shared between ranks (CUDA IPC), see include/tpt_b200.h
    def mem_alloc(self, nbytes: int) -> DevicePtr:
        p = ctypes.c_void_p()
        self._check(self._L.tpt_mem_alloc(self._h, nbytes, ctypes.byref(p)), "tpt_mem_alloc")
        return DevicePtr(p.value)

    def mem_free(self, ptr: DevicePtr):
        self._check(self._L.tpt_mem_free(self._h, ctypes.c_void_p(ptr.addr)), "tpt_mem_free")

    def mem_to_host(self, ptr: DevicePtr, out: np.ndarray):
        self._check(self._L.tpt_mem_copy(self._h, out.ctypes.data, ctypes.c_void_p(ptr.addr), out.nbytes, 2), "tpt_mem_copy")
        return out

    def mem_from_host(self, ptr: DevicePtr, src: np.ndarray):
        src = np.ascontiguousarray(src)
        self._check(self._L.tpt_mem_copy(self._h, ctypes.c_void_p(ptr.addr), src.ctypes.data, src.nbytes, 1), "tpt_mem_copy")

    def ipc_export(self, ptr: DevicePtr) -> bytes:
        h = ctypes.create_string_buffer(64)
        self._check(self._L.tpt_ipc_export(self._h, ctypes.c_void_p(ptr.addr), h), "tpt_ipc_export")
        return bytes(h.raw)

    def ipc_open(self, handle: bytes) -> DevicePtr:
        p = ctypes.c_void_p()
        self._check(self._L.tpt_ipc_open(self._h, handle, ctypes.byref(p)), "tpt_ipc_open")
        return DevicePtr(p.value)

    def ipc_close(self, ptr: DevicePtr):
        self._check(self._L.tpt_ipc_close(self._h, ctypes.c_void_p(ptr.addr)), "tpt_ipc_close")

    def debug_libm(self, fn: int, x: np.ndarray) -> np.ndarray:
        """Device-side libm restatement of the exact mode: fn 0 sinf, 1 cosf, 2 powf(x,5)."""
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        self._check(self._L.tpt_debug_libm(self._h, fn, x.ctypes.data, out.ctypes.data, x.size), "tpt_debug_libm")
        return out

    def debug_hit(self, kform: int, rays: np.ndarray):
        """Nearest hit of rays[n, 6] = {o.xyz, d.xyz} with sweep form `kform` of the fast kernels -> (ids int32, t float32)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        ids = np.empty(len(rays), np.int32)
        t = np.empty(len(rays), np.float32)
        self._check(self._L.tpt_debug_hit(self._h, kform, rays.ctypes.data, ids.ctypes.data, t.ctypes.data, len(rays)), "tpt_debug_hit")
        return ids, t

    def tonemap_srgb8(self, image, width: int, height: int) -> np.ndarray:
        addr, on_dev, _keep = _as_ptr(image)
        out = np.empty((height, width, 4), np.uint8)
        self._check(self._L.tpt_tonemap_srgb8(self._h, ctypes.c_void_p(addr), 1 if on_dev else 0, width, height,
                                              out.ctypes.data, 0, None), "tpt_tonemap_srgb8")
        return out

    def tonemap_rgba8(self, image, width: int, height: int, transfer: int = 0, bgr: bool = False, flip_y: bool = True) -> np.ndarray:
        """transfer 0: LinearToSRGB (PixelShader.hlsl), 1: sqrt gamma (Emscripten/main.cpp:67-79), 2: the C# TGA writer's
        (Cs/Program.cs:34-68; use bgr=True, flip_y=False for its byte layout)."""
        addr, on_dev, _keep = _as_ptr(image)
        out = np.empty((height, width, 4), np.uint8)
        self._check(self._L.tpt_tonemap_rgba8(self._h, ctypes.c_void_p(addr), 1 if on_dev else 0, width, height,
                                              out.ctypes.data, 0, transfer, 1 if bgr else 0, 1 if flip_y else 0, None),
                    "tpt_tonemap_rgba8")
        return out


# ---- the reference's Test.h API (C++-mangled symbols of the drop-in shim) -----------------------------------
def _load_shim():
    global _shim
    if _shim is not None:
        return _shim
    _load_lib()
    if not os.path.exists(SHIM_PATH):
        raise TptError(f"{SHIM_PATH} is missing: build with `make -C toypathtracer_b200/csrc`")
    S = ctypes.CDLL(SHIM_PATH)
    ci, cu, cf, vp = ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_void_p
    S._Z14InitializeTestv.argtypes = []; S._Z14InitializeTestv.restype = None
    S._Z12ShutdownTestv.argtypes = []; S._Z12ShutdownTestv.restype = None
    S._Z10UpdateTestfiiij.argtypes = [cf, ci, ci, ci, cu]; S._Z10UpdateTestfiiij.restype = None
    S._Z8DrawTestfiiiPfRij.argtypes = [cf, ci, ci, ci, vp, ctypes.POINTER(ci), cu]; S._Z8DrawTestfiiiPfRij.restype = None
    S._Z14GetObjectCountRiS_S_S_.argtypes = [ctypes.POINTER(ci)] * 4; S._Z14GetObjectCountRiS_S_S_.restype = None
    S._Z12GetSceneDescPvS_S_S_Pi.argtypes = [vp, vp, vp, vp, ctypes.POINTER(ci)]; S._Z12GetSceneDescPvS_S_S_Pi.restype = None
    S.tpt_shim_set_mode.argtypes = [ci]; S.tpt_shim_set_mode.restype = None
    S.tpt_shim_reset_scene.argtypes = []; S.tpt_shim_reset_scene.restype = None
    S.tpt_shim_set_variant.argtypes = [ci, ci]; S.tpt_shim_set_variant.restype = None
    _shim = S
    return S


def InitializeTest():
    """Test.h:10 — creates the CUDA context (the reference creates its CPU task scheduler here)."""
    _load_shim()._Z14InitializeTestv()


def ShutdownTest():
    """Test.h:11"""
    _load_shim()._Z12ShutdownTestv()


def UpdateTest(time: float, frameCount: int, screenWidth: int, screenHeight: int, testFlags: int = 0):
    """Test.h:13 — host-side scene animation, camera, emissive list; uploads the scene to the device."""
    _load_shim()._Z10UpdateTestfiiij(time, frameCount, screenWidth, screenHeight, testFlags)


def DrawTest(time: float, frameCount: int, screenWidth: int, screenHeight: int, backbuffer: np.ndarray,
             testFlags: int = 0) -> int:
    """Test.h:14 — renders one frame into the caller-owned float RGBA backbuffer; returns outRayCount."""
    if not (isinstance(backbuffer, np.ndarray) and backbuffer.dtype == np.float32 and backbuffer.flags["C_CONTIGUOUS"]
            and backbuffer.size == screenWidth * screenHeight * 4):
        raise TptError("backbuffer must be a C-contiguous float32 array of width*height*4")
    rc = ctypes.c_int(0)
    _load_shim()._Z8DrawTestfiiiPfRij(time, frameCount, screenWidth, screenHeight, backbuffer.ctypes.data,
                                      ctypes.byref(rc), testFlags)
    return int(rc.value)


def GetObjectCount() -> Tuple[int, int, int, int]:
    """Test.h:16 -> (count, objectSize, materialSize, camSize)"""
    v = [ctypes.c_int(0) for _ in range(4)]
    _load_shim()._Z14GetObjectCountRiS_S_S_(*[ctypes.byref(x) for x in v])
    return tuple(int(x.value) for x in v)


def GetSceneDesc():
    """Test.h:17 -> (spheres[n] SPHERE_DTYPE, materials[n] MATERIAL_DTYPE, camera CAMERA_DTYPE, emissive ids)."""
    n, so, sm, sc = GetObjectCount()
    spheres = np.zeros(n, SPHERE_DTYPE); mats = np.zeros(n, MATERIAL_DTYPE); cam = np.zeros(1, CAMERA_DTYPE)
    em = np.zeros(max(n, 1), np.int32); ec = ctypes.c_int(0)
    _load_shim()._Z12GetSceneDescPvS_S_S_Pi(spheres.ctypes.data, mats.ctypes.data, cam.ctypes.data, em.ctypes.data,
                                            ctypes.byref(ec))
    return spheres, mats, cam, em[: ec.value].copy()


def reset_scene():
    """Un-animated scene again (UpdateTest with kFlagAnimate moves spheres 1 and 8 permanently, Test.cpp:304-308)."""
    _load_shim().tpt_shim_reset_scene()


def set_variant(big_scene: bool = True, mitsuba_compare: bool = False):
    """The reference's compile-time switches DO_BIG_SCENE (Test.cpp:10-11: 46 vs 9 spheres) and DO_MITSUBA_COMPARE
    (Config.h:25: constant sky, zero Metal roughness, zero aperture) at run time; resets the scene."""
    _load_shim().tpt_shim_set_variant(1 if big_scene else 0, 1 if mitsuba_compare else 0)


def set_mode(mode: int):
    """Mode used by DrawTest(): MODE_EXACT (default, bit-identical to the reference) or MODE_FAST."""
    _load_shim().tpt_shim_set_mode(mode)


from .scenes import reference_scene, stress_scene, make_camera  # noqa: E402,F401
