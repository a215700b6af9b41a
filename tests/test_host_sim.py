"""HOST build of the product's integrator source (tests/host_sim/exact_sim.cpp includes
toypathtracer_b200/csrc/tpt_integrator.cuh, EXACT instantiation) against the oracle and the golden vectors:
checks the kernel's logic bit-for-bit on a machine without a GPU. (The GPU run of the same source is
tests/test_gpu_exact.py.)"""
import ctypes
import os

import numpy as np

from conftest import bits_differ
from test_oracle import GOLD, golden_scene


def sim_render(host_sim, sph, mats, cam, w, h, f0, nf, flags, spp=4, flat=False, split=False):
    L = host_sim["exact_sim"]
    buf = np.zeros((h, w, 4), np.float32)
    rays = (ctypes.c_longlong * nf)()
    vp = lambda a: np.ascontiguousarray(a).ctypes.data_as(ctypes.c_void_p)
    sph = np.ascontiguousarray(sph); mats = np.ascontiguousarray(mats); cam = np.ascontiguousarray(cam)
    fn = L.sim_render_exact_split if split else (L.sim_render_exact_flat if flat else L.sim_render_exact)
    fn(vp(sph), vp(mats), sph.nbytes // 20, vp(cam), w, h, f0, nf, ctypes.c_uint(flags), spp, vp(buf), rays, 0)
    return buf, [int(r) for r in rays]


def test_product_source_matches_golden(host_sim):
    g = np.load(os.path.join(GOLD, "ref_192x108_f0-3.npz"))
    sph, mats, cam, em = golden_scene()
    buf, rays = sim_render(host_sim, sph, mats, cam, 192, 108, 0, 4, 2)
    assert rays == [int(r) for r in g["rays"]]
    assert not bits_differ(buf, g["image"]).any()


def test_product_source_matches_oracle_on_runtime_scene(host_sim, oracle):
    import toypathtracer_b200 as tpt
    sph, mats, cam, em = tpt.stress_scene(160, 90, count=203)   # 203 % 4 != 0: padded spheres in play
    obuf, orays, pads = oracle.orc_render(sph, mats, cam, 160, 90, 3, 2, flags=2)
    buf, rays = sim_render(host_sim, sph, mats, cam, 160, 90, 3, 2, 2)
    assert rays == orays
    assert not bits_differ(buf, obuf, pads).any()


def test_flat_state_machine_matches_golden_and_oracle(host_sim, oracle):
    """xchain_step (one sweep per step; the batched LANES = 1 kernel's form) is bit-identical to the nested form."""
    g = np.load(os.path.join(GOLD, "ref_192x108_f0-3.npz"))
    sph, mats, cam, em = golden_scene()
    buf, rays = sim_render(host_sim, sph, mats, cam, 192, 108, 0, 4, 2, flat=True)
    assert rays == [int(r) for r in g["rays"]]
    assert not bits_differ(buf, g["image"]).any()
    import toypathtracer_b200 as tpt
    s2, m2, c2, e2 = tpt.stress_scene(160, 90, count=203)
    obuf, orays, pads = oracle.orc_render(s2, m2, c2, 160, 90, 3, 2, flags=2)
    buf, rays = sim_render(host_sim, s2, m2, c2, 160, 90, 3, 2, 2, flat=True)
    assert rays == orays and not bits_differ(buf, obuf, pads).any()


def test_split_path_and_shade_streams_match_golden_and_oracle(host_sim, oracle):
    """xpath_sample / xshade_event (the two-warp kernel's form: the whole row's path stream first, shading afterwards)
    is bit-identical to the nested form, incl. padded-sphere hits and scenes with 0, 1 and 6 lights."""
    g = np.load(os.path.join(GOLD, "ref_192x108_f0-3.npz"))
    sph, mats, cam, em = golden_scene()
    buf, rays = sim_render(host_sim, sph, mats, cam, 192, 108, 0, 4, 2, split=True)
    assert rays == [int(r) for r in g["rays"]]
    assert not bits_differ(buf, g["image"]).any()
    import toypathtracer_b200 as tpt
    s2, m2, c2, e2 = tpt.stress_scene(160, 90, count=203)
    obuf, orays, pads = oracle.orc_render(s2, m2, c2, 160, 90, 3, 2, flags=2)
    buf, rays = sim_render(host_sim, s2, m2, c2, 160, 90, 3, 2, 2, split=True)
    assert rays == orays and not bits_differ(buf, obuf, pads).any()
    cam = tpt.make_camera((0, 1, 4), (0, 0, 0), (0, 1, 0), 45, 2.0, 0.05, 4)
    for n in (1, 2, 5):
        s3 = np.zeros(n, tpt.SPHERE_DTYPE); m3 = np.zeros(n, tpt.MATERIAL_DTYPE)
        for i in range(n):
            s3[i] = ((i - n / 2, 0, 0), 0.45, 0)
            m3[i] = (i % 3, (0.7, 0.6, 0.5), (4, 4, 4) if i == 1 else (0, 0, 0), 0.1, 1.5)
        obuf, orays, pads = oracle.orc_render(s3, m3, cam, 64, 32, 0, 2, flags=2, spp=3)
        buf, rays = sim_render(host_sim, s3, m3, cam, 64, 32, 0, 2, 2, spp=3, split=True)
        assert rays == orays and not bits_differ(buf, obuf, pads).any()


def test_fastdiv_matches_integer_division(host_sim):
    """tpt_fastdiv.h (path index -> pixel/sample in the wavefront kernel): every divisor shape, n < 2^31."""
    L = host_sim["fastdiv_check"]
    L.check_fastdiv.restype = ctypes.c_longlong
    divisors = [1, 2, 3, 4, 5, 7, 16, 64, 100, 255, 256, 257, 1000, 1280, 1920, 3840, 4096, 65535, 65536, 65537, 1 << 20, (1 << 20) + 7,
                (1 << 30) - 1, 1 << 30, 0x7fffffff]
    for d in divisors:
        assert L.check_fastdiv(ctypes.c_uint32(d), ctypes.c_uint32(d * 2654435761 & 0xffffffff), ctypes.c_longlong(200000)) == 0, d
