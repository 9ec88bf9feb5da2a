"""Corner detection: the C restatement oracle/hv_oracle_gftt.c against the compiled reference
(oracle/_ref/libref_detect.so = src/tracker/feature_detector.cpp + feature_detector_legacy.cpp, unmodified, on CPU images).
Tolerances: the corner response is fp32 and OpenCV forms its 3 x 3 box sums as running sums (and switches code paths with the CPU:
AVX / FMA dispatch), so |d response| <= 1e-6 + 1e-5 |r| (measured: 1e-9); key point coordinates and the final corner list are
integers / exact floats and must be IDENTICAL on these inputs (a block whose two best responses tie within the response tolerance
could legitimately differ; none does here)."""
import numpy as np
import pytest

from hybvio_b200 import synth
from oracle import gftt_oracle


@pytest.fixture(scope="module")
def orc():
    import os, subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", ROOT, "oracle"], stdout=subprocess.DEVNULL)
    return gftt_oracle.OracleGftt()


@pytest.fixture(scope="module")
def ref():
    if not gftt_oracle.have_ref():
        pytest.skip("oracle/_ref/libref_detect.so not built (needs /root/reference)")
    return gftt_oracle.RefGftt()


CASES = [(752, 480, 3), (512, 512, 5), (203, 77, 1), (64, 64, 2), (97, 130, 7)]


@pytest.mark.parametrize("w,h,k", CASES)
def test_response_matches_cv_corner_min_eigen_val(orc, ref, w, h, k):
    img, _ = synth.stereo_frame(k, w, h)
    a, b = orc.response(img), ref.response(img)
    assert np.all(np.abs(a - b) <= 1e-6 + 1e-5 * np.abs(b)), float(np.abs(a - b).max())
    flat = np.full((h, w), 77, np.uint8)
    assert np.abs(orc.response(flat)).max() == 0.0 and np.abs(ref.response(flat)).max() == 0.0


@pytest.mark.parametrize("w,h,k", CASES)
@pytest.mark.parametrize("mask_radius", [0, 50, 17])
def test_corner_list_matches_the_reference_detector(orc, ref, w, h, k, mask_radius):
    """detect(): cell maxima (GAIN 16, > gfttMinResponse), stable sort, the n leading zero points of `corners.resize(n)` followed by
    push_back (feature_detector.cpp:632-634), applyMinDistance against previous corners and against the accepted ones, maxTracks."""
    img, _ = synth.stereo_frame(k, w, h)
    prev = synth.interior_points(40, w, h, seed=5, margin=5.0)
    for p in (None, prev):
        a = orc.detect(img, p, mask_radius, max_tracks=150)
        b = ref.detect(img, p, mask_radius, max_tracks=150)
        assert a.shape == b.shape and np.array_equal(a, b)
    if mask_radius == 0 and w >= 64:
        assert len(a) == 2 * (w // 32) * (h // 32) and not a[: (w // 32) * (h // 32)].any()       # the resize quirk


def test_cells_without_a_qualifying_pixel_report_zero(orc, ref):
    img, _ = synth.stereo_frame(0, 256, 128)
    img[:, :128] = 100                                   # flat half: response 0 < gfttMinResponse
    kp = orc.collect(orc.response(img))
    assert kp.shape == (8 * 4, 3)
    flat = kp.reshape(4, 8, 3)[:, :3]
    assert (flat[..., 2] == np.float32(-1e10)).all() and not flat[..., :2].any()
    assert np.array_equal(orc.detect(img, None, 50, 100), ref.detect(img, None, 50, 100))
