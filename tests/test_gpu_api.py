"""C-ABI behaviour around the kernels: buffer ownership (only the rendered rows are written), alpha, asynchronous scene
upload vs draws in flight, per-context options, argument validation. (ADVICE r01 findings, each with its test.)"""
import numpy as np
import pytest

from conftest import bits_differ, rel_l2
from test_oracle import golden_scene

pytestmark = pytest.mark.gpu

W, H = 320, 180


def test_fast_host_draw_touches_only_its_rows(gpu_ctx):
    """A fast-mode draw into an unpacked HOST buffer that covers only some rows (TraceRowJob(start,end), Test.cpp:266)
    must leave every other row of the caller's buffer alone — also when `prev` has zero weight and is not uploaded."""
    sph, mats, cam, em = golden_scene()
    gpu_ctx.set_scene(sph, mats, cam, em)
    for variant in (0, 1, 3, 5, 8):
        gpu_ctx.set_option("fast_variant", variant)
        for rows in [(0, 32, 1, 0), (40, 17, 1, 0), (1, 20, 7, 0)]:
            for flags in (0, 2):
                buf = np.full((H, W, 4), 7.0, np.float32)
                gpu_ctx.draw(3, 1, W, H, buf, flags=flags, mode=1, rows=rows)
                mine = np.zeros(H, bool)
                mine[rows[0]:rows[0] + rows[1] * rows[2]:rows[2]] = True
                assert (buf[~mine] == 7.0).all(), (variant, rows, flags)
                assert np.isfinite(buf[mine]).all() and (buf[mine][..., :3] != 7.0).any(), (variant, rows, flags)
                if flags == 2:
                    assert (buf[mine][..., 3] == 7.0).all()      # progressive: prev uploaded, alpha preserved
    gpu_ctx.set_option("fast_variant", 3)
    # exact mode: same ownership rule
    buf = np.full((H, W, 4), 7.0, np.float32)
    gpu_ctx.draw(3, 1, W, H, buf, flags=0, mode=0, rows=(1, 20, 7, 0))
    mine = np.zeros(H, bool); mine[1:141:7] = True
    assert (buf[~mine] == 7.0).all() and (buf[mine][..., 3] == 7.0).all() and (buf[mine][..., :3] != 7.0).any()


def test_fast_device_buffer_keeps_alpha(gpu_ctx):
    """The reference never writes alpha (Maths.h:38). Fast mode on a device buffer keeps it even when prev has zero
    weight; "fast_alpha_zero" waives that."""
    import torch
    sph, mats, cam, em = golden_scene()
    gpu_ctx.set_scene(sph, mats, cam, em)
    for variant in (0, 1, 3, 5, 6, 8):
        gpu_ctx.set_option("fast_variant", variant)
        img = torch.full((H, W, 4), float("nan"), dtype=torch.float32, device="cuda")    # prev RGB must not be read as a number
        img[..., 3] = 0.5
        gpu_ctx.draw(0, 1, W, H, img, flags=0, mode=1)
        out = img.cpu().numpy()
        assert np.isfinite(out).all(), variant
        assert (out[..., 3] == 0.5).all(), variant
    gpu_ctx.set_option("fast_variant", 3)
    gpu_ctx.set_option("fast_alpha_zero", 1)
    img = torch.full((H, W, 4), 0.5, dtype=torch.float32, device="cuda")
    gpu_ctx.draw(0, 1, W, H, img, flags=0, mode=1)
    gpu_ctx.set_option("fast_alpha_zero", 0)
    assert (img[..., 3] == 0).all()


def test_scene_update_does_not_race_draws_in_flight(libs):
    """tpt_set_scene right after an asynchronous device-buffer draw on the caller's stream (what an animated shell does
    every frame, TestWin.cpp:261-283): the draw in flight must still see ITS scene. Alternates two scenes 40 times."""
    import torch
    ctx = libs.Context(0)
    sphA, mats, cam, em = golden_scene()
    sphB = sphA.copy()
    sphB.view(np.float32).reshape(-1, 5)[1:9, 1] += 0.75          # lift the hero spheres
    stream = torch.cuda.Stream()
    w, h = 640, 360
    want = {}
    for name, sph in (("A", sphA), ("B", sphB)):
        ctx.set_scene(sph, mats, cam, em)
        ref = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        ctx.draw(5, 1, w, h, ref, flags=0, mode=1)
        torch.cuda.synchronize()
        want[name] = ref.cpu().numpy()
    assert rel_l2(want["A"], want["B"]) > 1e-2
    imgs = [torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(40)]
    with torch.cuda.stream(stream):
        for i in range(40):
            ctx.set_scene(sphA if i % 2 == 0 else sphB, mats, cam, em)
            ctx.draw(5, 1, w, h, imgs[i], flags=0, mode=1, stream=stream.cuda_stream, want_rays=False)
    torch.cuda.synchronize()
    for i in range(40):
        assert rel_l2(imgs[i].cpu().numpy(), want["A" if i % 2 == 0 else "B"]) < 1e-5, i
    # draws of one context on two different streams are ordered too (shared counters)
    s2 = torch.cuda.Stream()
    a = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    b = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.read_ray_count()
    ctx.draw(5, 1, w, h, a, flags=0, mode=1, stream=stream.cuda_stream, want_rays=False)
    ctx.draw(5, 1, w, h, b, flags=0, mode=1, stream=s2.cuda_stream, want_rays=False)
    torch.cuda.synchronize()
    one = ctx.draw(5, 1, w, h, a, flags=0, mode=1)
    assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    assert ctx.read_ray_count() == 3 * one
    ctx.close()


def test_kform_option_is_per_context(libs):
    sph, mats, cam, em = golden_scene()
    a, b = libs.Context(0), libs.Context(0)
    for c in (a, b):
        c.set_scene(sph, mats, cam, em)
    ref = np.zeros((H, W, 4), np.float32)
    r_ref = b.draw(0, 4, W, H, ref, flags=2, mode=1)
    a.set_option("fast_kform", 0)                                   # must not leak into context b
    out = np.zeros((H, W, 4), np.float32)
    r_out = b.draw(0, 4, W, H, out, flags=2, mode=1)
    assert r_out == r_ref and rel_l2(out, ref) < 1e-5
    oa = np.zeros((H, W, 4), np.float32)
    r_a = a.draw(0, 4, W, H, oa, flags=2, mode=1)
    assert r_a != r_ref or rel_l2(oa, ref) > 0                      # the other sweep form rounds differently somewhere
    assert abs(r_a / r_ref - 1) < 1e-3
    a.close(); b.close()


def test_scene_validation_and_empty_shard(libs):
    ctx = libs.Context(0)
    sph, mats, cam, em = golden_scene()
    with pytest.raises(libs.TptError, match="emissive id"):
        ctx.set_scene(sph, mats, cam, np.array([8, 46], np.int32))
    with pytest.raises(libs.TptError, match="emissive id"):
        ctx.set_scene(sph, mats, cam, np.array([-1], np.int32))
    ctx.set_scene(sph, mats, cam, em)
    buf = np.full((8, 8, 4), 3.0, np.float32)
    assert ctx.draw(0, 1, 8, 8, buf, rows=(0, 0, 1, 0)) == 0        # more ranks than rows: an empty shard is a no-op
    assert (buf == 3.0).all()
    from toypathtracer_b200 import multigpu as mg
    assert mg.rows_of_rank(4, 6, 8) == (6, 0, 8)
    ctx.close()


def test_identical_scene_is_not_uploaded_twice_but_changes_are_seen(gpu_ctx):
    sph, mats, cam, em = golden_scene()
    gpu_ctx.set_scene(sph, mats, cam, em)
    a = np.zeros((H, W, 4), np.float32); gpu_ctx.draw(0, 1, W, H, a, flags=0, mode=0)
    gpu_ctx.set_scene(sph, mats, cam, em)                           # same bytes: no upload, same result
    b = np.zeros((H, W, 4), np.float32); gpu_ctx.draw(0, 1, W, H, b, flags=0, mode=0)
    assert not bits_differ(a, b).any()
    m2 = mats.copy(); m2.view(np.float32).reshape(-1, 9)[0, 1:4] = (0.2, 0.9, 0.2)   # green ground
    gpu_ctx.set_scene(sph, m2, cam, em)
    c = np.zeros((H, W, 4), np.float32); gpu_ctx.draw(0, 1, W, H, c, flags=0, mode=0)
    assert bits_differ(a, c).mean() > 0.3
    gpu_ctx.set_scene(sph, mats, cam, em)
    d = np.zeros((H, W, 4), np.float32); gpu_ctx.draw(0, 1, W, H, d, flags=0, mode=0)
    assert not bits_differ(a, d).any()


@pytest.mark.gpu
def test_scene_upload_accounting(libs):
    """tpt_set_scene skips the copy when the bytes are resident (a shell calls UpdateTest every frame) and says so;
    "scene_upload_always" (bench.py's end-to-end leg) copies the packed blob every time without invalidating anything."""
    ctx = libs.Context(0)
    sph, mats, cam, em = golden_scene()
    ctx.set_scene(sph, mats, cam, em)
    first = ctx.last_scene_upload_bytes()
    assert first >= 46 * 20 + 46 * 36                  # the packed blob holds at least the caller's arrays
    ctx.set_scene(sph, mats, cam, em)
    assert ctx.last_scene_upload_bytes() == 0          # same bytes: nothing copied
    ctx.set_option("scene_upload_always", 1)
    img = np.zeros((108, 192, 4), np.float32)
    ctx.set_option("exact_lookahead", 4)
    rays = []
    for f in range(4):
        ctx.set_scene(sph, mats, cam, em)
        assert ctx.last_scene_upload_bytes() == first
        rays.append(ctx.draw(f, 1, 192, 108, img, flags=2, mode=0))
        assert ctx.last_launch_count() == (3 if f == 0 else 2)     # forced re-uploads of identical bytes keep the frame cache
    ctx.set_option("scene_upload_always", 0); ctx.set_option("exact_lookahead", 0)
    ref = np.zeros((108, 192, 4), np.float32)
    assert rays == [ctx.draw(f, 1, 192, 108, ref, flags=2, mode=0) for f in range(4)] and (ref.view(np.uint32) == img.view(np.uint32)).all()
    ctx.close()
