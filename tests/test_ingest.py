"""Frame ingest: colour -> gray and undistortion / rectification.
CPU: oracle (orc_gray / orc_remap, oracle/hv_oracle_gftt.c) against the compiled reference (accelerated-arrays pixelwiseAffine as image.cpp builds
it; Undistorter::buildMono -> undistort of src/tracker/undistorter.cpp) and against the committed golden vectors generated from it.
GPU: hv_ingest_frame through the C ABI against the oracle and the golden vectors: BIT-exact (integer output, fp32 interpolation in the
reference's operation order), and the pyramid built from the ingested frame equals the pyramid of the reference's output image.
Pixels whose bilinear taps leave the reference's image buffer (it reads them without a bounds check, undistorter.cpp:101) are undefined in the
reference itself (two runs differ) and are excluded."""
import os

import numpy as np
import pytest

from hybvio_b200 import synth
from oracle import ingest_oracle as io

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest_golden.npz")


@pytest.fixture(scope="module")
def orc():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"], stdout=subprocess.DEVNULL)
    return io.OracleIngest()


def defined(table, w, h):
    t = table.reshape(h, w)
    return (t["x0"] == io.INVALID) | ((t["x0"] + 1 < w) & (t["y0"] + 1 < h))


def test_oracle_matches_reference_golden(orc):
    g = np.load(GOLD)
    for c in (3, 4):
        assert np.array_equal(orc.gray(g[f"rgb{c}"]), g[f"rgb{c}_gray"])
    img = g["frame"]
    h, w = img.shape
    for name in ("pinhole", "fisheye", "zoomout"):
        table = g[name + "_table"].reshape(-1).view(io.REMAP_DTYPE)
        ok = defined(table, w, h)
        assert ok.mean() > 0.95
        assert np.array_equal(orc.remap(img, table)[ok], g[name + "_out"][ok]), name
    assert (g["zoomout_table"].reshape(-1).view(io.REMAP_DTYPE)["x0"] == io.INVALID).mean() > 0.1      # the zoomed-out case has pixels without a source


def test_oracle_matches_compiled_reference(orc):
    if not io.have_ref():
        pytest.skip("oracle/_ref/libref_ingest.so not built (needs /root/reference)")
    ref = io.RefIngest()
    rng = np.random.RandomState(2)
    for c in (3, 4):
        img = rng.randint(0, 256, (77, 203, c)).astype(np.uint8)
        assert np.array_equal(orc.gray(img), ref.gray(img))
    img, _ = synth.stereo_frame(2, 320, 240)
    for fish, dist, f, zoom in ((0, [-0.28, 0.07, 0.0002], 230.0, 1.0), (1, [-0.01, 0.02, -0.01, 0.003], 190.0, 1.0), (0, [-0.28, 0.07, 0.0002], 230.0, 0.7)):
        out, table = ref.undistort(img, fish, f, f, 159.5, 119.5, dist, zoom)
        ok = defined(table, 320, 240)
        assert np.array_equal(orc.remap(img, table)[ok], out[ok])


@pytest.mark.gpu
def test_ingest_kernels_bit_exact_and_pyramid_follows(hv, orc, oracle_lk):
    from hybvio_b200 import capi
    g = np.load(GOLD)
    # colour -> gray (3 and 4 channels), straight into level 0 of a pyramid
    for c in (3, 4):
        img = g[f"rgb{c}"]
        h, w = img.shape[:2]
        ing, pyr = capi.Ingest(hv, w, h), hv.pyramid(w, h, 31, 1)
        gray = ing.frame(img, pyr)
        assert np.array_equal(gray, g[f"rgb{c}_gray"]) and np.array_equal(gray, orc.gray(img))
        assert np.array_equal(pyr.download(0)[0], gray)
        ing.close(); pyr.release()
    # undistortion tables of the reference's cameras
    img = g["frame"]
    h, w = img.shape
    ing, pyr = capi.Ingest(hv, w, h), hv.pyramid(w, h, 31, 2)
    for name in ("pinhole", "fisheye", "zoomout"):
        table = g[name + "_table"].reshape(-1).view(io.REMAP_DTYPE)
        ing.set_remap(table)
        out = ing.frame(img, pyr)
        assert np.array_equal(out, orc.remap(img, table)), name                       # every pixel, incl. the clamped out-of-buffer taps
        ok = defined(table, w, h)
        assert np.array_equal(out[ok], g[name + "_out"][ok]), name
        o = oracle_lk.pyramid(out, 31, 2)                                             # the pyramid was built from the ingested frame in place
        for lv in range(pyr.levels):
            gg, dd = pyr.download(lv)
            og, od = o.download(lv, padded=False)
            assert np.array_equal(gg, og) and np.array_equal(dd, od), (name, lv)
    # colour AND remap in one call; random table incl. entries without a source and border taps
    rng = np.random.RandomState(9)
    rgb = rng.randint(0, 256, (h, w, 4)).astype(np.uint8)
    table = np.zeros(w * h, io.REMAP_DTYPE)
    table["x0"] = rng.randint(0, w, w * h); table["y0"] = rng.randint(0, h, w * h)
    table["xfrac"] = rng.rand(w * h).astype(np.float32); table["yfrac"] = rng.rand(w * h).astype(np.float32)
    table["x0"][rng.rand(w * h) < 0.1] = io.INVALID
    ing.set_remap(table)
    out = ing.frame(rgb, pyr)
    assert np.array_equal(out, orc.remap(orc.gray(rgb), table))
    ing.set_remap(None)
    assert np.array_equal(ing.frame(img, pyr), img)                                    # plain gray frame: identity (= hv_pyr_build)
    ing.close(); pyr.release()
