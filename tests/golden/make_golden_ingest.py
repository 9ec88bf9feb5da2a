"""Generates tests/golden/ingest_golden.npz from the COMPILED REFERENCE (oracle/_ref/libref_ingest.so): accelerated-arrays colour -> gray as
src/tracker/image.cpp:360-366 builds it, and Undistorter::buildMono(...)->undistort (src/tracker/undistorter.cpp) for a distorted pinhole
and a Kannala-Brandt fisheye camera, together with the camera mapping tables (reference Camera classes). Small frames: the .npz travels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybvio_b200 import synth            # noqa: E402
from oracle import ingest_oracle as io   # noqa: E402

ref = io.RefIngest()
rng = np.random.RandomState(5)
out = {}
for c in (3, 4):
    img = rng.randint(0, 256, (60, 101, c)).astype(np.uint8)
    out[f"rgb{c}"] = img
    out[f"rgb{c}_gray"] = ref.gray(img)
img, _ = synth.stereo_frame(2, 200, 120)
out["frame"] = img
for name, (fish, dist, f, zoom) in {"pinhole": (0, [-0.28, 0.07, 0.0002], 150.0, 1.0), "fisheye": (1, [-0.01, 0.02, -0.01, 0.003], 120.0, 1.0),
                                     "zoomout": (0, [-0.28, 0.07, 0.0002], 150.0, 0.7)}.items():
    und, table = ref.undistort(img, fish, f, f, 99.5, 59.5, dist, zoom)
    out[name + "_out"] = und
    out[name + "_table"] = table.view(np.uint8).reshape(-1, 12)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ingest_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
