"""Strong scaling of ONE image across GPUs (BASELINE configs[3]: 3840x2160, 64 spp, rows interleaved over ranks).
Run under torchrun. Two assembly methods:
  A  NCCL baseline: every rank renders a packed band (fast variant 3), one all_gather assembles the image;
  B  fused peer write-out: every rank's tile kernel (fast variant 5) stores finished pixels straight into the
     root's image over NVLink (CUDA IPC mapping), no gather — only a ray-count all_reduce as completion barrier.
Prints one JSON line on rank 0."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import toypathtracer_b200 as tpt
from toypathtracer_b200 import multigpu as mg

W = int(os.environ.get("TPT_W", 3840)); H = int(os.environ.get("TPT_H", 2160)); NF = int(os.environ.get("TPT_FRAMES", 16))
REPS = int(os.environ.get("TPT_REPS", 3))
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
ctx = tpt.Context(local)
ctx.set_scene(*tpt.reference_scene(W, H))
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); sh = stream.cuda_stream
row0, nrows, step = mg.rows_of_rank(H, rank, world)

def timed(fn):
    best = None; out = None
    for r in range(REPS):
        torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); out = fn(); e1.record(stream)
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = t.item() if best is None else min(best, t.item())
    return best, out

# ---- A: packed band + all_gather
ctx.set_option("fast_variant", 3)
band = torch.zeros((nrows, W, 4), dtype=torch.float32, device=dev)
def method_a():
    ctx.draw(0, NF, W, H, band, flags=2, mode=tpt.MODE_FAST, rows=(row0, nrows, step, 1), stream=sh, want_rays=False)
    return mg.gather_rows(band, H, rank, world)
ctx.read_ray_count(sh)
ms_a, img_a = timed(method_a)
rays_a = mg.sum_ray_counts(ctx.read_ray_count(sh), dev) // REPS

# ---- B: fused peer write-out into the root's image
ctx.set_option("fast_variant", 5)
shared = mg.SharedImage(ctx, W, H, rank)
def method_b():
    ctx.draw(0, NF, W, H, shared.ptr, flags=2, mode=tpt.MODE_FAST, rows=(row0, nrows, step, 0), stream=sh, want_rays=False)
    done = torch.zeros(1, device=dev); dist.all_reduce(done)          # completion barrier on the stream
    return None
ms_b, _ = timed(method_b)
rays_b = mg.sum_ray_counts(ctx.read_ray_count(sh), dev) // REPS
# diagnostics: the same tile kernel into LOCAL memory (packed band and strided full image), no collective
local_full = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
ms_b_local_packed, _ = timed(lambda: ctx.draw(0, NF, W, H, band, flags=2, mode=tpt.MODE_FAST, rows=(row0, nrows, step, 1), stream=sh, want_rays=False))
ms_b_local_strided, _ = timed(lambda: ctx.draw(0, NF, W, H, local_full, flags=2, mode=tpt.MODE_FAST, rows=(row0, nrows, step, 0), stream=sh, want_rays=False))
ms_b_peer_nobarrier, _ = timed(lambda: ctx.draw(0, NF, W, H, shared.ptr, flags=2, mode=tpt.MODE_FAST, rows=(row0, nrows, step, 0), stream=sh, want_rays=False))
ctx.set_option("fast_variant", 3)
ms_a_kernel_only, _ = timed(lambda: ctx.draw(0, NF, W, H, band, flags=2, mode=tpt.MODE_FAST, rows=(row0, nrows, step, 1), stream=sh, want_rays=False))
ctx.read_ray_count(sh)
torch.cuda.synchronize(dev); dist.barrier()
if rank == 0:
    img_b = shared.to_host()[..., :3].astype(np.float64)
    a = img_a.cpu().numpy()[..., :3].astype(np.float64)
    rel = float(np.sqrt(((a - img_b) ** 2).sum() / (a ** 2).sum()))
    print(json.dumps({"workload": f"{W}x{H} {NF * 4} spp, 46 spheres, rows interleaved over {world} GPUs", "n_gpus": world,
                      "A_nccl_allgather": {"ms": ms_a, "mray_s": rays_a / ms_a / 1e3, "rays": rays_a, "kernel": "k_fast_queue (variant 3)"},
                      "B_fused_peer_writeout": {"ms": ms_b, "mray_s": rays_b / ms_b / 1e3, "rays": rays_b, "kernel": "k_fast_tileq (variant 5)"},
                      "relL2_A_vs_B": rel, "diag_ms": {"v5_local_packed": ms_b_local_packed, "v5_local_strided": ms_b_local_strided,
                                            "v5_peer_no_barrier": ms_b_peer_nobarrier, "v3_kernel_only": ms_a_kernel_only}, "noise_floor_two_independent_renders": 0.194 / np.sqrt(NF * 4) * np.sqrt(2)}), flush=True)
dist.barrier()
shared.close()
dist.destroy_process_group()
