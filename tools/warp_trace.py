"""Per-warp timeline of k_fast_queue (diagnostic build, -DTPT_TRACE_WARPS=1):
    TPT_LIB_PATH=.../libtpt_ab_trace.so TPT_TRACE_FILE=gpurun_out/trace.bin python tools/warp_trace.py [analyze-only file]
Each warp records global-timer stamps (entry, first trip, queue exhausted, exit), trip counts and the lane-slots used
after the queue ran dry. Prints where the kernel's end effects go."""
import json, os, sys
import numpy as np


def analyze(path):
    t = np.fromfile(path, dtype=np.uint64).reshape(-1, 8)
    t = t[t[:, 0] > 0]
    e, r, d, x = (t[:, i].astype(np.int64) for i in range(4))
    trips, dtrips, dlanes, sm = (t[:, i].astype(np.int64) for i in range(4, 8))
    t0, t1 = e.min(), x.max()
    has_dry = d > 0
    d = np.where(has_dry, d, x)
    us = lambda v: float(v) / 1e3
    out = {
        "warps": int(len(t)), "kernel_us": us(t1 - t0),
        "entry_spread_us": us(e.max() - t0), "first_trip_after_entry_us_mean": us((r - e).mean()),
        "first_trip_latest_us": us(r.max() - t0),
        "queue_dry_first_us": us(d.min() - t0), "queue_dry_median_us": us(np.median(d) - t0), "queue_dry_last_us": us(d.max() - t0),
        "exit_p10_us": us(np.percentile(x, 10) - t0), "exit_median_us": us(np.median(x) - t0), "exit_p90_us": us(np.percentile(x, 90) - t0),
        "exit_p99_us": us(np.percentile(x, 99) - t0),
        "drain_us_mean": us((x - d).mean()), "drain_us_max": us((x - d).max()),
        "trips_mean": float(trips.mean()), "dry_trips_mean": float(dtrips.mean()), "dry_trips_max": int(dtrips.max()),
        "dry_lane_fill": float(dlanes.sum() / max(dtrips.sum() * 32, 1)),
        "warp_time_lost_to_early_exit_pct": float(100 * (t1 - x).sum() / ((t1 - t0) * len(t))),
        "warp_time_in_drain_pct": float(100 * (x - d).sum() / ((t1 - t0) * len(t))),
    }
    # per-SM: when does the SM's last warp leave
    last = {}
    for s, xx in zip(sm, x):
        last[s] = max(last.get(s, 0), xx)
    lv = np.array(list(last.values()))
    out["sm_last_exit_median_us"] = us(np.median(lv) - t0)
    out["sm_last_exit_min_us"] = us(lv.min() - t0)
    return out


if len(sys.argv) > 1:
    print(json.dumps(analyze(sys.argv[1])))
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toypathtracer_b200 as tpt
ctx = tpt.Context(0)
w, h = 1280, 720
ctx.set_scene(*tpt.reference_scene(w, h))
ctx.set_option("fast_variant", 3); ctx.set_option("fast_kform", 2)
img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
for r in range(4):
    rays = ctx.draw(r, 1, w, h, img, flags=0, mode=1)
print(json.dumps(dict(analyze(os.environ["TPT_TRACE_FILE"]), rays=int(rays), kernel_ms=ctx.last_kernel_ms())))
