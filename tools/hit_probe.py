"""Ray-by-ray comparison of the fast kernels' sweep forms (tpt_debug_hit): ids and distance bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import toypathtracer_b200 as tpt


def make_rays(spheres, n, seed):
    rng = np.random.default_rng(seed)
    c = np.stack([spheres["center"][:, 0], spheres["center"][:, 1], spheres["center"][:, 2]], 1).astype(np.float64) \
        if spheres["center"].ndim == 2 else None
    r = spheres["radius"].astype(np.float64)
    pick = rng.integers(0, len(r), n)
    pick[: n // 3] = 0                                   # a third of the rays leave the ground sphere
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    g = pick == 0
    # ground: points near the populated area (top of the sphere)
    xz = rng.uniform(-40, 40, (n, 2))
    top = np.stack([xz[:, 0], np.sqrt(r[0] ** 2 - xz[:, 0] ** 2 - xz[:, 1] ** 2), xz[:, 1]], 1) / r[0]
    nrm[g] = top[g]
    o = c[pick] + nrm * r[pick, None]
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    flip = (d * nrm).sum(1) < 0
    flip &= rng.random(n) < 0.8                          # most leave the surface, some go inside (refraction)
    d[flip] = -d[flip]
    graze = rng.random(n) < 0.2                          # grazing directions: the interesting ones for rounding
    d[graze] = d[graze] - 0.98 * (d[graze] * nrm[graze]).sum(1, keepdims=True) * nrm[graze]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cam = rng.random(n) < 0.2
    o[cam] = np.array([0, 6, 14.0]) + rng.normal(size=(cam.sum(), 3)) * 0.02
    return np.concatenate([o, d], 1).astype(np.float32)


def main():
    ctx = tpt.Context(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    for name in ("stress", "ref"):
        if name == "stress":
            s = tpt.stress_scene(1920, 1080, count=4096)
        else:
            s = tpt.reference_scene(1280, 720)
        ctx.set_scene(s[0], s[1], s[2], None if name == "stress" else s[3])
        rays = make_rays(s[0], n, 7)
        res = {k: ctx.debug_hit(k, rays) for k in (0, 1, 2, 3)}
        print(name, "hit fraction", float((res[0][0] >= 0).mean()), flush=True)
        for a, b in ((0, 3), (1, 2), (0, 1)):
            ida, ta = res[a]; idb, tb = res[b]
            bad = (ida != idb) | (ta.view(np.uint32) != tb.view(np.uint32))
            print(f"  kform {a} vs {b}: {int(bad.sum())} of {n} rays differ (ids differ: {int((ida != idb).sum())})", flush=True)
            for i in np.nonzero(bad)[0][:6] if (a, b) != (0, 1) else []:
                print("    ray", rays[i].tolist(), "->", int(ida[i]), float(ta[i]), "|", int(idb[i]), float(tb[i]), flush=True)


if __name__ == "__main__":
    main()
