"""The design claims that can be read off the shipped binary, pinned on every build (CPU box: cuobjdump only, no GPU):
TMA bulk staging in every trace kernel, packed f32x2 math in the default sweep, 128-bit reductions / stores, warp
reductions in the exact mode, shared-memory addressing of the scene, sm_100a only, no tensor-core opcodes."""
import os, shutil, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "toypathtracer_b200", "libtpt_b200.so")
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def kernels():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_counts
    return {name: (n, cnt) for name, n, cnt in sass_counts.kernel_counts(LIB)}


def test_only_sm_100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    cubins = [l for l in out.splitlines() if ".cubin" in l]
    assert cubins and all("sm_100a" in l for l in cubins), cubins
    ptx = subprocess.run(["cuobjdump", "-lptx", LIB], capture_output=True, text=True).stdout
    assert ".ptx" not in ptx                      # no JIT fallback for other architectures


def test_trace_kernels_stage_the_scene_with_tma_and_use_no_tensor_cores(kernels):
    trace = [k for k in kernels if k.startswith(("k_fast_", "k_trace_exact", "k_refgpu"))]
    assert len(trace) > 30
    for k in trace:
        n, c = kernels[k]
        assert c["UBLKCP"] >= 1, k                # cp.async.bulk (TMA) staging of the sphere/material blob
    assert all(c["tensor"] == 0 for _, c in kernels.values())


def test_default_fast_kernel(kernels):
    n, c = kernels["k_fast_queue<128, 6, 2>"]     # what bench.py's `value` runs
    assert c["FFMA2"] >= 300 and c["FADD2"] >= 40          # packed-pair sweep (fma.rn.f32x2 / add.rn.f32x2)
    assert c["REDG.F32x4"] >= 1                            # one red.global.add.v4.f32 per finished path
    assert c["LD.E (generic)"] == 0 and c["BRX"] == 0      # scene addressed as shared memory; no jump table
    n8, c8 = kernels["k_fast_queue<128, 8, 2>"]            # the 64-register instance for long draws
    assert c8["FFMA2"] >= 300
    nb, cb = kernels["k_fast_queue<768, 1, 3>"]            # big scenes: conservative packed sweep
    assert cb["FFMA2"] >= 300


def test_group_kernel_writes_with_128_bit_stores(kernels):
    n, c = kernels["k_fast_group<6, 2, 1>"]       # what bench.py's `e2e` runs (pinned host buffer)
    assert c["STG.128"] >= 1 and c["REDG.F32x4"] == 0 and c["LD.E (generic)"] == 0


def test_exact_kernels_reduce_the_nearest_hit_in_hardware(kernels):
    for k in ("k_trace_exact<32>", "k_trace_exact<8>", "k_trace_exact_split<2, 0, 1>"):
        n, c = kernels[k]
        assert c["(C)REDUX"] >= 2, k              # __reduce_min_sync pair (t, then tie key)
        assert c["FFMA2"] == 0, k                 # exact mode: reference-form scalar arithmetic only
