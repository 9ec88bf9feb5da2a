"""GPU tests of the opt-in tracker kernels. (1) The second-generation pyramid kernel (hv_pyr_fused2_kernel, HV_PYR_V2=1, read once per process): every
pyramid parity test of tests/test_gpu_pyramid_lk.py -- oracle, the reference's golden vectors, batches, strided and device-resident
input, and the LK tests that consume those pyramids -- repeated in a child process with the switch set. The kernel body is bit-exact
on the host emulator (tests/test_emu_kernels.py); this file is what the first GPU session after it was written has to pass before the
switch becomes the default. (2) The 8-warp CTA-per-feature LK kernel (HV_LK_CTA_WARPS=8). Named zzz so that they run after the tests of the
default path."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_second_generation_pyramid_kernel_passes_the_pyramid_and_lk_parity_tests():
    if os.environ.get("HV_PYR_V2") or os.environ.get("HV_LK_CTA_WARPS"):
        pytest.skip("this is a child run")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_pyramid_lk.py"), "-q", "-m", "gpu", "-p", "no:cacheprovider"],
                       env={**os.environ, "HV_PYR_V2": "1"}, cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]


@pytest.mark.gpu
def test_eight_warps_per_feature_lk_kernel_passes_the_lk_parity_tests():
    """HV_LK_CTA_WARPS=8: hv_lk_cta_kernel<31, 8> (4 window rows per warp instead of 8). The sums are exact integers, so the results
    must be bit-identical to the 4-warp kernel: the same LK tests (oracle in the kernel's arithmetic, golden vectors of the compiled
    reference, edge cases, full-size properties), in a child process with the switch set."""
    if os.environ.get("HV_PYR_V2") or os.environ.get("HV_LK_CTA_WARPS"):
        pytest.skip("this is a child run")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_pyramid_lk.py"), "-q", "-m", "gpu", "-k", "lk", "-p", "no:cacheprovider"],
                       env={**os.environ, "HV_LK_CTA_WARPS": "8"}, cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
