"""Generates tests/golden/gftt_golden.npz from the COMPILED REFERENCE (oracle/_ref/libref_detect.so = the reference's own
src/tracker/feature_detector.cpp on CPU images): corner lists of FeatureDetector::detect and the cv::cornerMinEigenVal response of small
synthetic frames. Run in the build container (needs /root/reference at build time); the .npz travels to the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybvio_b200 import synth          # noqa: E402
from oracle import gftt_oracle        # noqa: E402

ref = gftt_oracle.RefGftt()
out = {}
for name, (w, h, k) in {"A": (752, 480, 3), "B": (512, 512, 5), "C": (203, 77, 1)}.items():
    img, _ = synth.stereo_frame(k, w, h)
    prev = synth.interior_points(40, w, h, seed=5, margin=5.0)
    out[name + "_img"] = img
    out[name + "_prev"] = prev
    out[name + "_corners_r0"] = ref.detect(img, None, 0, 150)
    out[name + "_corners_r50"] = ref.detect(img, prev, 50, 150)
    if name == "C":
        out[name + "_response"] = ref.response(img)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gftt_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
