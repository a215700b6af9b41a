"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/libtoyref.so, built by
oracle/Makefile from /root/reference). Run in the build container where the reference is mounted:

    python tests/golden/make_golden.py

Outputs (small, committed):
  scene46_1280x720.npz     raw GetSceneDesc export of the reference scene (spheres 46x20 B, materials 46x36 B,
                           camera 88 B, emissive ids) after UpdateTest(0, 0, 1280, 720, 0)
  ref_192x108_f0-3.npz     float RGBA image after frames 0..3 with kFlagProgressive (16 spp) + per-frame ray counts
  ref_counts.json          ray counts of larger runs (no images): 1280x720 frames 0..5, 3840x2160 frame 0
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
sph, mats, cam, em = pyoracle.ref_scene(1280, 720)
np.savez(os.path.join(here, "scene46_1280x720.npz"), spheres=sph, materials=mats.view(np.uint32), camera=cam, emissives=em)

buf, rays = pyoracle.ref_render(192, 108, 0, 4, flags=2)
np.savez_compressed(os.path.join(here, "ref_192x108_f0-3.npz"), image=buf, rays=np.asarray(rays, np.int64))

counts = {}
_, r = pyoracle.ref_render(1280, 720, 0, 6, flags=0)
counts["1280x720_flags0_frames0-5"] = r
_, r = pyoracle.ref_render(3840, 2160, 0, 1, flags=0)
counts["3840x2160_flags0_frame0"] = r
_, r = pyoracle.ref_render(1920, 1080, 0, 2, flags=0)
counts["1920x1080_flags0_frames0-1"] = r
json.dump(counts, open(os.path.join(here, "ref_counts.json"), "w"), indent=1)
print(counts)
