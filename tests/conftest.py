import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import toypathtracer_b200 as tpt
        return tpt.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`-m gpu` on a machine without a usable device (or without the built CUDA library) must fail loudly rather than
    deselect/skip its way to a green run: there is no CPU fallback to test."""
    expr = (config.getoption("-m") or "").strip()
    if expr == "gpu" and any(it.get_closest_marker("gpu") for it in items) and not _have_gpu():
        pytest.exit("-m gpu was requested but toypathtracer_b200 finds no CUDA device / libtpt_b200.so "
                    "(the product has no CPU path)", returncode=1)


@pytest.fixture(scope="session")
def libs():
    """Product libraries (built in-tree by csrc/Makefile; rebuilt here only when nvcc is available)."""
    import toypathtracer_b200 as tpt
    if not (os.path.exists(tpt.LIB_PATH) and os.path.exists(tpt.SHIM_PATH)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "toypathtracer_b200", "csrc"), "all"], check=True)
    return tpt


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    if not os.path.exists(pyoracle.ORC_SO):
        pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def host_sim(tmp_path_factory):
    """Test-only HOST builds of product sources (tests/host_sim/*.cpp)."""
    import ctypes
    out = tmp_path_factory.mktemp("host_sim")
    libs_ = {}
    for name in ("libm_check", "exact_sim", "fastdiv_check"):
        so = str(out / f"{name}.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", so,
                        os.path.join(ROOT, "tests", "host_sim", f"{name}.cpp"), "-lpthread", "-lm"], check=True)
        libs_[name] = ctypes.CDLL(so)
    return libs_


@pytest.fixture(scope="session")
def gpu_ctx(libs):
    ctx = libs.Context(0)   # raises TptError without a device: -m gpu tests then FAIL (no silent fallback)
    yield ctx
    ctx.close()


def bits_differ(a, b, exclude=()):
    """Boolean [h,w] map of pixels whose RGBA bits differ, minus (x, y[, frame]) pixels in `exclude`."""
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    # NaN payload/sign propagation is not specified by IEEE 754 (x86 keeps the operand's payload, the GPU emits
    # the canonical NaN): a NaN matches a NaN, everything else is compared bit for bit.
    d = ((a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))).any(axis=2)
    for p in exclude:
        d[p[1], p[0]] = False
    return d


def rel_l2(img, ref):
    img = np.asarray(img, np.float64)[..., :3]; ref = np.asarray(ref, np.float64)[..., :3]
    return float(np.sqrt(((img - ref) ** 2).sum() / (ref ** 2).sum()))
