"""Corner detection on the device: hv_gftt_detect (csrc/gftt.cu) through the C ABI against
  * oracle/hv_oracle_gftt.c on the same frames: key points (x, y) and responses BIT-exact (the kernel uses the oracle's operation order),
  * the reference's golden corner lists (tests/golden/gftt_golden.npz, generated from the compiled reference's FeatureDetector::detect):
    after the host half of the detector (sort, resize quirk, applyMinDistance: here the oracle's restatement; in the adapter
    hybvio_b200/host/cuda_feature_detector.cpp the reference's own code) the corner lists are identical."""
import os

import numpy as np
import pytest

from hybvio_b200 import synth
from oracle import gftt_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gftt_golden.npz")


@pytest.fixture(scope="module")
def orc():
    return gftt_oracle.OracleGftt()


def device_keypoints(hv, img, cell=32, min_response=1e-3):
    p = hv.pyramid(img.shape[1], img.shape[0], 31, 1)
    p.build(np.ascontiguousarray(img))
    kp = p.gftt_detect(3, cell, min_response)
    p.release()
    return kp


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,k", [(752, 480, 3), (512, 512, 5), (203, 77, 1), (64, 64, 2), (97, 130, 7), (751, 479, 4), (1280, 720, 6)])
@pytest.mark.parametrize("cell", [32, 16, 8])
def test_key_points_bit_exact_vs_oracle(hv, orc, w, h, k, cell):
    img, _ = synth.stereo_frame(k, w, h)
    kp = device_keypoints(hv, img, cell)
    ok = orc.collect(orc.response(img), cell, 1e-3)
    assert kp.shape == ok.shape
    assert np.array_equal(kp[:, :2], ok[:, :2]), np.nonzero((kp[:, :2] != ok[:, :2]).any(axis=1))[0][:10]
    assert np.array_equal(kp[:, 2].view(np.uint32), ok[:, 2].view(np.uint32))


@pytest.mark.gpu
def test_flat_cells_report_zero_and_borders_reflect(hv, orc):
    img, _ = synth.stereo_frame(0, 256, 128)
    img[:, :128] = 100
    kp = device_keypoints(hv, img)
    assert np.array_equal(kp, orc.collect(orc.response(img)))
    # strong structure in the very first / last rows and columns: the reflect-101 borders of Sobel and of the box filter decide
    rng = np.random.RandomState(3)
    img = rng.randint(0, 255, (96, 160)).astype(np.uint8)
    assert np.array_equal(device_keypoints(hv, img), orc.collect(orc.response(img)))


@pytest.mark.gpu
def test_corner_lists_match_the_reference_golden(hv, orc):
    g = np.load(GOLD)
    for name in "ABC":
        img, prev = g[name + "_img"], g[name + "_prev"]
        kp = device_keypoints(hv, img)
        assert np.array_equal(orc.corners(kp, None, 0, 150), g[name + "_corners_r0"]), name
        assert np.array_equal(orc.corners(kp, prev, 50, 150), g[name + "_corners_r50"]), name


def test_oracle_matches_the_reference_golden(orc):
    """CPU: the restatement against the committed vectors of the compiled reference (runs without oracle/_ref)."""
    g = np.load(GOLD)
    for name in "ABC":
        img, prev = g[name + "_img"], g[name + "_prev"]
        assert np.array_equal(orc.detect(img, None, 0, 150), g[name + "_corners_r0"])
        assert np.array_equal(orc.detect(img, prev, 50, 150), g[name + "_corners_r50"])
    r = orc.response(g["C_img"])
    assert np.all(np.abs(r - g["C_response"]) <= 1e-6 + 1e-5 * np.abs(g["C_response"]))
