"""The product's glibc-faithful sinf/cosf/powf (toypathtracer_b200/csrc/tpt_libm.cuh, host build) against the
platform libm the reference links — exhaustively over the hot path's argument domain."""
import ctypes


def test_sincos_whole_path_domain(host_sim):
    L = host_sim["libm_check"]
    L.check_sincos_domain.restype = ctypes.c_longlong
    first = ctypes.c_longlong(-1)
    assert L.check_sincos_domain(ctypes.byref(first)) == 0, f"first mismatch at k={first.value}"


def test_sincos_sampled_full_range(host_sim):
    L = host_sim["libm_check"]
    L.check_sincos_range.restype = ctypes.c_longlong
    assert L.check_sincos_range(ctypes.c_uint32(61)) == 0


def test_powf_x5_sampled_and_random_pairs(host_sim):
    L = host_sim["libm_check"]
    L.check_powf_random.restype = ctypes.c_longlong
    assert L.check_powf_random(ctypes.c_longlong(4_000_000), ctypes.c_uint32(0xC0FFEE)) == 0


def test_powf_x5_all_floats(host_sim):
    """All 2^32 bit patterns of x for the reference's only call shape powf(x, 5) (Maths.h:331)."""
    import os
    L = host_sim["libm_check"]
    L.check_powf_all_x.restype = ctypes.c_longlong
    first = ctypes.c_ulonglong(0)
    nthreads = max(1, os.cpu_count() or 1)
    assert L.check_powf_all_x(ctypes.c_float(5.0), nthreads, ctypes.byref(first)) == 0, hex(first.value)


def test_powf_cuberoot_whole_rand01_domain(host_sim):
    """pow(RandomFloat01(), 1.0/3.0) of the reference's GPU sampler (ComputeShader.hlsl:33), all 2^24 arguments."""
    L = host_sim["libm_check"]
    L.check_powf_rand01_domain.restype = ctypes.c_longlong
    assert L.check_powf_rand01_domain(ctypes.c_float(1.0 / 3.0)) == 0
