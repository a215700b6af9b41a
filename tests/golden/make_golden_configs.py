"""Golden fixtures for the BASELINE.json configurations themselves (slow: minutes of CPU; run in the build container
where /root/reference is mounted and oracle/_ref/libtoyref.so exists):

    python tests/golden/make_golden_configs.py [c2] [c3] [c4] [c5]

Writes tests/golden/configs.json (+ c5_rows_1920x1080.npz):
  c2  1280x720, frames 0..255 with kFlagProgressive (= 1024 spp, BASELINE configs[1] correctness statement):
      per-frame ray counts from the UNMODIFIED reference, sha256 of its final float image with the padded-sphere
      pixels zeroed, and those pixels (x, y, frame) from the restatement (oracle/restate.cpp reports them; the
      reference's colour there is undefined behaviour).
  c3  3840x2160, frames 0..3 progressive (16 spp, configs[2]): the same three items.
  c4  3840x2160, frames 0..15, flags 0 (64 spp, configs[3]): per-frame ray counts of the reference.
  c5  4096-sphere stress scene, 1920x1080, frames 0..1 progressive (8 spp, configs[4]): full-frame ray counts of the
      restatement (the reference cannot run this scene) and the float pixels of 8 rows (npz).
Pixel hashes are over the little-endian float32 bytes, rows bottom-up as in the backbuffer.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(here, "configs.json")
C5_ROWS = (7, 8, 134)           # rows 7, 141, ..., 945 (row0, numRows, rowStep)


def image_hash(img, pads):
    img = img.copy()
    for p in pads:
        img[p[1], p[0]] = 0
    return hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()


def main():
    which = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5"}
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}

    if "c2" in which:
        t0 = time.time()
        w, h, n = 1280, 720, 256
        sph, mats, cam, em = pyoracle.ref_scene(w, h)
        rbuf, rrays = pyoracle.ref_render(w, h, 0, n, flags=2)
        obuf, orays, pads = pyoracle.orc_render(sph, mats, cam, w, h, 0, n, flags=2)
        assert orays == rrays, "restatement and reference disagree on ray counts"
        d = (rbuf.view(np.uint32) != obuf.view(np.uint32)).any(axis=2)
        for p in pads:
            d[p[1], p[0]] = False
        assert not d.any(), "restatement and reference disagree on pixels"
        out["c2_1280x720_flags2_frames0-255"] = {"rays": rrays, "total": int(sum(rrays)), "pads": [list(p) for p in pads],
                                                 "sha256_pads_zeroed": image_hash(rbuf, pads)}
        print("c2", sum(rrays), len(pads), time.time() - t0, flush=True)
        json.dump(out, open(OUT, "w"), indent=1)

    if "c3" in which:
        t0 = time.time()
        w, h, n = 3840, 2160, 4
        sph, mats, cam, em = pyoracle.ref_scene(w, h)
        rbuf, rrays = pyoracle.ref_render(w, h, 0, n, flags=2)
        obuf, orays, pads = pyoracle.orc_render(sph, mats, cam, w, h, 0, n, flags=2)
        assert orays == rrays
        d = (rbuf.view(np.uint32) != obuf.view(np.uint32)).any(axis=2)
        for p in pads:
            d[p[1], p[0]] = False
        assert not d.any()
        out["c3_3840x2160_flags2_frames0-3"] = {"rays": rrays, "total": int(sum(rrays)), "pads": [list(p) for p in pads],
                                                "sha256_pads_zeroed": image_hash(rbuf, pads)}
        print("c3", rrays, pads, time.time() - t0, flush=True)
        json.dump(out, open(OUT, "w"), indent=1)

    if "c4" in which:
        t0 = time.time()
        w, h, n = 3840, 2160, 16
        _, rrays = pyoracle.ref_render(w, h, 0, n, flags=0)
        out["c4_3840x2160_flags0_frames0-15"] = {"rays": rrays, "total": int(sum(rrays))}
        print("c4", sum(rrays), time.time() - t0, flush=True)
        json.dump(out, open(OUT, "w"), indent=1)

    if "c5" in which:
        t0 = time.time()
        import toypathtracer_b200 as tpt
        w, h, n = 1920, 1080, 2
        sph, mats, cam, em = tpt.stress_scene(w, h, count=4096)
        buf, rays, pads = pyoracle.orc_render(sph, mats, cam, w, h, 0, n, flags=2)
        r0, nr, rs = C5_ROWS
        rows = buf[r0:r0 + nr * rs:rs].copy()
        _, row_rays, row_pads = pyoracle.orc_render(sph, mats, cam, w, h, 0, n, flags=2, rows=C5_ROWS)
        out["c5_stress4096_1920x1080_flags2_frames0-1"] = {"rays": rays, "total": int(sum(rays)), "pads": [list(p) for p in pads],
                                                           "rows": list(C5_ROWS), "row_rays": row_rays,
                                                           "sha256_pads_zeroed": image_hash(buf, pads)}
        np.savez_compressed(os.path.join(here, "c5_rows_1920x1080.npz"), rows=rows)
        print("c5", rays, pads, time.time() - t0, flush=True)
        json.dump(out, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
