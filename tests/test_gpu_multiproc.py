"""Two processes sharing ONE GPU exercise the multi-GPU path end to end (CUDA IPC works across processes on the same
device): rank 1 opens rank 0's image and both ranks' kernels write their interleaved rows straight into it. In exact
mode the result must be bit-identical to a single-process render."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 192, 108


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    import toypathtracer_b200 as tpt
    from toypathtracer_b200 import multigpu as mg
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = tpt.Context(0)
    ctx.set_scene(*tpt.reference_scene(W, H))
    shared = mg.SharedImage(ctx, W, H, rank)
    row0, nrows, step = mg.rows_of_rank(H, rank, world)
    rays = ctx.draw(0, 3, W, H, shared.ptr, flags=2, mode=tpt.MODE_EXACT, rows=(row0, nrows, step, 0))
    total = mg.sum_ray_counts(rays, "cpu")
    dist.barrier()
    if rank == 0:
        np.savez(out_path, image=shared.to_host(), rays=total)
    dist.barrier()
    shared.close()
    dist.destroy_process_group()


def test_two_ranks_write_one_image_over_ipc(gpu_ctx, tmp_path):
    import torch.multiprocessing as mp
    import toypathtracer_b200 as tpt
    out = str(tmp_path / "shared.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    gpu_ctx.set_scene(*tpt.reference_scene(W, H))
    ref = np.zeros((H, W, 4), np.float32)
    rays = gpu_ctx.draw(0, 3, W, H, ref, flags=2, mode=tpt.MODE_EXACT)
    assert int(got["rays"]) == rays
    assert (got["image"].view(np.uint32) == ref.view(np.uint32)).all()
