"""world_size-2 (and 3) gloo runs of the multi-GPU host logic on CPU: row-interleaved sharding + gather, frame
sharding + combine, ragged heights."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, height, width):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from toypathtracer_b200 import multigpu as mg
    row0, nrows, step = mg.rows_of_rank(height, rank, world)
    assert row0 == rank and step == world
    ys = torch.arange(nrows) * step + row0
    band = torch.zeros((nrows, width, 4))
    band[:, :, 0] = ys[:, None].float()                       # pixel = (y, x, rank, 1)
    band[:, :, 1] = torch.arange(width)[None, :].float()
    band[:, :, 2] = rank
    band[:, :, 3] = 1
    img = mg.gather_rows(band, height, rank, world)
    assert img.shape == (height, width, 4)
    assert (img[:, :, 0] == torch.arange(height)[:, None].float()).all()
    assert (img[:, :, 1] == torch.arange(width)[None, :].float()).all()
    assert (img[:, :, 2] == (torch.arange(height) % world)[:, None].float()).all()
    # every row rendered exactly once across ranks
    owned = torch.zeros(height); owned[ys] = 1
    dist.all_reduce(owned); assert (owned == 1).all()
    # frames
    frames = mg.frames_of_rank(10, 7, rank, world)
    allf = [None] * world
    dist.all_gather_object(allf, frames)
    assert sorted(sum(allf, [])) == list(range(10, 17))
    local = torch.full((4, 4, 4), float(sum(frames)) / max(1, len(frames)))
    mean = mg.combine_frame_means(local, len(frames))
    assert torch.allclose(mean, torch.full((4, 4, 4), sum(range(10, 17)) / 7.0))
    assert mg.sum_ray_counts(100 + rank, "cpu") == sum(100 + r for r in range(world))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 16), (2, 17), (3, 10)])
def test_row_and_frame_sharding_gloo(world, height):
    mp.spawn(_worker, args=(world, _free_port(), height, 8), nprocs=world, join=True)
