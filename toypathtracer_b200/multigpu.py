"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed, NCCL over NVLink), no collective
inside the tracing itself.

Two partitionings, both over units the reference treats as independent:

* rows   — every (frame,row) is an independent RNG chain (Cpp/Source/Test.cpp:278-280). Rank r renders rows
           r, r+W, r+2W, ... (row-interleaved: per-row cost varies smoothly with y, so interleaving balances
           load) into a packed band, then ONE all_gather assembles the image (`gather_rows`). Works for the exact
           mode too (bit-identical to a single-GPU render).
* frames — N spp = N/4 frames with different seeds (Test.cpp:280) accumulated by a progressive mean
           (Test.cpp:272-276). Rank r accumulates frames r, r+W, ... locally; ONE all_reduce(sum) at the end of the
           accumulation combines the per-rank means (`combine_frame_means`). Fast mode only (float summation order
           differs from the serial lerp; SURVEY §9.3 measures that difference at 2.7e-7 relL2).

The functions below are backend-agnostic (gloo on CPU in tests/test_dist_cpu.py, nccl on the GPUs)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def rows_of_rank(height: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(row0, numRows, rowStep) of the row-interleaved shard; heights that are not a multiple of `world` give the
    low ranks one extra row."""
    n = (height - rank + world - 1) // world if rank < height else 0
    return rank, n, world


def gather_rows(band: torch.Tensor, height: int, rank: int, world: int, group=None) -> torch.Tensor:
    """band: [rows_of_rank, W, 4] packed rows of this rank. Returns the assembled [H, W, 4] image on every rank.
    Payload: H*W*16 bytes in total (3840x2160: 132.7 MB -> 16.6 MB per GPU at 8 ranks)."""
    w = band.shape[1]
    nmax = (height + world - 1) // world
    if band.shape[0] < nmax:                                  # ragged tail: pad so all_gather sees equal shapes
        pad = torch.zeros((nmax - band.shape[0], w, 4), dtype=band.dtype, device=band.device)
        band = torch.cat([band, pad], 0)
    gathered = torch.empty((world * nmax, w, 4), dtype=band.dtype, device=band.device)   # concat form (gloo + nccl)
    dist.all_gather_into_tensor(gathered, band.contiguous(), group=group)
    gathered = gathered.view(world, nmax, w, 4)
    # gathered[r, i] is image row r + i*world  ->  [nmax, world, W, 4] is row-major over y
    img = gathered.permute(1, 0, 2, 3).reshape(nmax * world, w, 4)
    return img[:height].contiguous()


class SharedImage:
    """The root rank's image, writable by every rank's kernels over NVLink (CUDA IPC peer mapping): the fused
    alternative to gather_rows — each rank's trace kernel stores its finished pixels straight into the root's HBM.
    `ctx` is this rank's toypathtracer_b200.Context."""

    def __init__(self, ctx, width: int, height: int, rank: int, root: int = 0, group=None):
        self.ctx, self.rank, self.root, self.width, self.height = ctx, rank, root, width, height
        self.nbytes = width * height * 16
        handle = [None]
        if rank == root:
            self.ptr = ctx.mem_alloc(self.nbytes)
            handle[0] = ctx.ipc_export(self.ptr)
        dist.broadcast_object_list(handle, src=root, group=group)
        if rank != root:
            self.ptr = ctx.ipc_open(handle[0])

    def to_host(self):
        """Root only: the assembled image as a numpy array [H, W, 4]."""
        import numpy as np
        assert self.rank == self.root
        out = np.empty((self.height, self.width, 4), np.float32)
        return self.ctx.mem_to_host(self.ptr, out)

    def close(self):
        if self.rank == self.root:
            self.ctx.mem_free(self.ptr)
        else:
            self.ctx.ipc_close(self.ptr)


def frames_of_rank(frame0: int, num_frames: int, rank: int, world: int) -> List[int]:
    return [f for f in range(frame0, frame0 + num_frames) if (f - frame0) % world == rank]


def combine_frame_means(local_mean: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """Each rank holds the mean of `local_count` frames; returns the mean over all ranks' frames (progressive mean
    of the whole accumulation), on every rank."""
    acc = local_mean * float(local_count)
    cnt = torch.tensor([float(local_count)], dtype=torch.float64, device=local_mean.device)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return acc / cnt.item()


def sum_ray_counts(rays: int, device, group=None) -> int:
    t = torch.tensor([rays], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())
