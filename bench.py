#!/usr/bin/env python
"""bench.py — Mray/s of the ToyPathTracer hot path (DrawTest -> Trace/HitWorld/Scatter) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode fast|exact]

Workload (BASELINE.json configs[1]): the reference's 46-sphere scene, 1280x720, 4 spp per frame, flags = 0
(non-progressive: every frame is a fresh image, BASELINE "DO_PROGRESSIVE off"). One STEP = one frame = one DrawTest
call of the reference (UpdateTest + DrawTest): ~16.8 M rays (camera + bounce + shadow, Test.cpp:122,199).

  value   whole-job Mray/s with the image resident in HBM: rays of the K timed steps / device time of those steps
          (CUDA events on the launching stream, one pair per step, summed; L2 flushed between steps, flush not
          timed), max over ranks.
  e2e     the same metric through the drop-in C-ABI with HOST buffers: per step tpt_set_scene (scene H2D, what
          UpdateTest+GetSceneDesc+UpdateSubresource do in the reference's GPU shells, TestWin.cpp:258-283) +
          tpt_draw (kernel, image D2H to pinned host memory, ray count D2H), wall clock around the synchronous call.
  N > 1   frames are independent units (Test.cpp:280 seeds depend on the frame index): step s renders frames
          s*N .. s*N+N-1, one per GPU (weak scaling: one frame per GPU per step); ONE all_reduce(sum) of the
          14.7 MB image at the end of the timed region combines the ranks' frames (the exchange step of a
          frame-sharded accumulation, multigpu.combine_frame_means). No collective inside the tracing.

--impl reference times the UNMODIFIED reference C++ path (oracle/_ref/libtoyref.so, enkiTS on all host threads)
on the same workload; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SPP = 1280, 720, 4
WORKLOAD = "46-sphere reference scene, 1280x720, 4 spp/frame, flags=0 (BASELINE configs[1]); step = one frame (one DrawTest)"
SPHERES = 46
FP32_TESTS_PER_S_PEAK = 148 * 128 * 1.965e9 / 17.0   # ~16 FP32 ops + compare per ray-sphere test, no FMA


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
    (per launch, same workload); None when no capture of this round's kernel is committed."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01", "ncu_traffic.json")))
        return t.get("k_fast_queue_1280x720_4spp_dram_bytes")
    except Exception:
        return None


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        busy = [s for s in sm if s > 300] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(args, rank):
    """The reference's own CPU implementation on the host cores (its enkiTS scheduler uses every hardware thread,
    Cpp/Source/enkiTS/TaskScheduler.cpp:1301). One step = one 1280x720 4-spp frame, like ours."""
    if rank != 0:
        return
    from oracle import pyoracle
    if not pyoracle.have_ref():
        pyoracle.build()
    kind = "reference" if pyoracle.have_ref() else "port"
    cores = os.cpu_count() or 1
    buf = np.zeros((H, W, 4), np.float32)
    if kind == "reference":
        render = lambda f0, n: pyoracle.ref_render(W, H, f0, n, flags=0, buf=buf, want_seconds=True)[1:]
    else:
        import toypathtracer_b200 as tpt
        sph, mats, cam, em = tpt.reference_scene(W, H)
        render = lambda f0, n: (lambda r: (r[1], r[3]))(pyoracle.orc_render(sph, mats, cam, W, H, f0, n, flags=0, buf=buf, want_seconds=True))
    render(0, max(2, args.warmup))            # first frames pay thread spin-up
    rays, secs = render(args.warmup, args.steps)
    total_s = sum(secs)
    value = sum(rays) / total_s / 1e6
    line = {"impl": "reference", "metric": "Mray/s on 46-sphere scene @1280x720", "value": value, "unit": "Mray/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "threads": cores},
            "cpu_baseline": {"value": value, "unit": "Mray/s", "cores": cores, "kind": kind,
                             "sample": f"{args.steps} frames of 1280x720x4spp after {max(2, args.warmup)} warm-up frames, "
                                       "UpdateTest+DrawTest per frame, steady clock"},
            "e2e": {"value": value, "unit": "Mray/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    """Bounded sample of the reference CPU path on this box's host cores (rank 0, N=1 only): ~10-20 s."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.have_ref() else "port"
    cores = os.cpu_count() or 1
    buf = np.zeros((H, W, 4), np.float32)
    if kind == "reference":
        run = lambda f0, n: pyoracle.ref_render(W, H, f0, n, flags=0, buf=buf, want_seconds=True)[1:]
    else:
        import toypathtracer_b200 as tpt
        sph, mats, cam, em = tpt.reference_scene(W, H)
        run = lambda f0, n: (lambda r: (r[1], r[3]))(pyoracle.orc_render(sph, mats, cam, W, H, f0, n, flags=0, buf=buf, want_seconds=True))
    run(0, 3)
    t0 = time.time()
    rays, secs, frames = [], [], 0
    while time.time() - t0 < 12.0 and frames < 240:
        r, s = run(3 + frames, 8)
        rays += r; secs += s; frames += 8
    # median frame (discard warm-up, median of >= 30 frames)
    per_frame = sorted(r / s / 1e6 for r, s in zip(rays, secs))
    return {"value": per_frame[len(per_frame) // 2], "unit": "Mray/s", "cores": cores, "kind": kind,
            "sample": f"median of {frames} frames of 1280x720x4spp (UpdateTest+DrawTest each) after 3 warm-up frames; "
                      f"mean {sum(rays) / sum(secs) / 1e6:.1f} Mray/s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--variant", type=int, default=-1, help="fast kernel variant (default: library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import toypathtracer_b200 as tpt
    from toypathtracer_b200 import multigpu as mg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = tpt.Context(local_rank)
    sph, mats, cam, em = tpt.reference_scene(W, H)
    ctx.set_scene(sph, mats, cam, em)
    if args.variant >= 0:
        ctx.set_option("fast_variant", args.variant)
    mode = tpt.MODE_FAST if args.mode == "fast" else tpt.MODE_EXACT
    # flags = 0 exactly like the reference step; at N > 1 rank r renders global frame s*N + r.
    image = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.Stream(dev)          # a real (non-NULL) stream: the library treats NULL as "my own stream"
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream

    def step(s, timed_events=None):
        frame = s * world + rank
        flush.fill_(s & 0xFF)                                             # L2 flush, outside the timed events
        if timed_events is not None:
            timed_events[0].record(stream)
        ctx.draw(frame, 1, W, H, image, flags=0, mode=mode, stream=sh, want_rays=False)
        if timed_events is not None:
            timed_events[1].record(stream)

    for s in range(args.warmup):
        step(s)
    if world > 1:
        dist.all_reduce(image, op=dist.ReduceOp.SUM)                      # warm-up includes the exchange step (NCCL sets up
        dist.all_reduce(image, op=dist.ReduceOp.SUM)                      # its buffers / NVLS on the first large collective)
    ctx.read_ray_count(sh)                                                # reset the accumulated counter
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    wall0 = time.perf_counter()
    for s in range(args.steps):
        step(args.warmup + s, evs[s])
    extra = None
    if world > 1:
        extra = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        extra[0].record(stream)
        # the exchange step of the frame-sharded accumulation: one all_reduce(sum) of the image
        dist.all_reduce(image, op=dist.ReduceOp.SUM)
        extra[1].record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = sum(step_ms) + (extra[0].elapsed_time(extra[1]) if extra else 0.0)
    rays = ctx.read_ray_count(sh)
    launches = ctx.last_launch_count() * args.steps

    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        rays = mg.sum_ray_counts(rays, dev)
        lt = torch.tensor([launches], dtype=torch.int64, device=dev); dist.all_reduce(lt); launches = int(lt.item())

    # ---- end to end through the C-ABI with host buffers (every rank; rank 0 reports the max time)
    e2e_steps = min(args.steps, 50)
    host = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory().numpy()
    for s in range(3):
        ctx.set_scene(sph, mats, cam, em)
        ctx.draw(s * world + rank, 1, W, H, host, flags=0, mode=mode)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for s in range(e2e_steps):
        ctx.set_scene(sph, mats, cam, em)                                  # scene H2D (UpdateTest + upload)
        e2e_rays += ctx.draw((args.warmup + s) * world + rank, 1, W, H, host, flags=0, mode=mode)  # kernel + image D2H + count D2H
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
        e2e_rays = mg.sum_ray_counts(e2e_rays, dev)
    scene_bytes = 46 * 20 + 46 * 36 + 88 + 2 * 4
    h2d = scene_bytes + (W * H * 16 if mode == tpt.MODE_EXACT else 0)      # exact mode uploads prev (bit parity)
    d2h = W * H * 16 + 8

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        value = rays / (dev_ms * 1e-3) / 1e6
        kms = statistics.mean(step_ms)
        alg_bytes = W * H * 16                                             # one float4 per pixel written, flags=0: no read
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        tests_per_s = (rays / world / args.steps) * (SPHERES + 2) / (kms * 1e-3)   # 48 padded spheres swept per ray
        line = {
            "metric": "Mray/s on 46-sphere scene @1280x720", "value": value, "unit": "Mray/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mode": args.mode, "parallelism": f"frames x{world}" if world > 1 else "1 GPU",
                       "l2": "flushed between steps (256 MiB fill), not timed",
                       "timing": "CUDA events per step on the launching stream, summed; max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mray/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "tpt_set_scene + tpt_draw(host backbuffer) per step, wall clock"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": load_traffic() if args.mode == "fast" else None, "peak_source": peak_src,
                         "kernel": "k_fast_queue" if args.mode == "fast" else "k_trace_exact",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "the path is FP32-issue bound, not HBM bound (SURVEY §8d): 16 B/pixel written per launch; "
                                 "see fp32 for the binding roofline",
                         "fp32": {"sphere_tests_per_s": tests_per_s, "peak_tests_per_s": FP32_TESTS_PER_S_PEAK,
                                  "frac": tests_per_s / FP32_TESTS_PER_S_PEAK,
                                  "peak_def": "148 SM x 128 lanes x 1.965 GHz / 17 FP32 issue slots per test (no FMA)"}},
            "wall_s": wall,
        }
        if world == 1 and args.mode == "fast":
            # the bit-exact mode on the same step, for the record (same API, same buffers; latency-bound: one serial RNG
            # chain per image row, Test.cpp:280)
            ex_steps = 8
            for s_ in range(2):
                ctx.draw(s_, 1, W, H, image, flags=0, mode=tpt.MODE_EXACT, stream=sh, want_rays=False)
            ctx.read_ray_count(sh)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for s_ in range(ex_steps):
                ctx.draw(2 + s_, 1, W, H, image, flags=0, mode=tpt.MODE_EXACT, stream=sh, want_rays=False)
            ev1.record(stream)
            torch.cuda.synchronize(dev)
            ex_rays = ctx.read_ray_count(sh)
            ex_ms = ev0.elapsed_time(ev1)
            line["exact_mode"] = {"value": ex_rays / ex_ms / 1e3, "unit": "Mray/s", "ms_per_step": ex_ms / ex_steps, "steps": ex_steps,
                                  "note": "TPT_MODE_EXACT: pixels and ray counts bit-identical to the reference CPU path"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_sample()
            except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"error": str(ex)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
