"""Interface-level parity of the tracker adapters (hybvio_b200/host/cuda_tracker_backends.cpp): the reference's OWN abstract
interfaces tracker::ImagePyramid::Factory::compute / tracker::OpticalFlow::compute are driven once with the reference's CPU back
ends (src/tracker/image_pyramid.cpp, optical_flow.cpp, compiled unmodified over the vendored OpenCV) and once with the CUDA back
ends on the same accelerated::Image frames (oracle/ref_build/tracker_iface_test.cpp, built by build_tracker_iface.sh).
Tolerance: Feature::Status identical; end points <= 1e-3 px for >= 99.9 %, < 3e-2 px for all."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "run_tracker_iface_test")


def _run(mode):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/run_tracker_iface_test not built (needs /root/reference at build time)")
    r = subprocess.run([EXE, mode], cwd=os.path.join(ROOT, "oracle", "_ref"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "tracker interface test passed" in r.stdout
    return r.stdout


def test_tracker_interface_harness_with_reference_backends():
    """CPU: the harness itself, with the reference back ends only (plausible TRACKED / FAILED_FLOW / FLOW_OUT_OF_RANGE counts)."""
    out = _run("ref")
    assert out.count("reference back ends") == 2


@pytest.mark.gpu
def test_cuda_tracker_backends_match_reference_backends_through_the_reference_interfaces():
    out = _run("both")
    assert out.count("0 status differences") == 2 and out.count(": ok") == 2
