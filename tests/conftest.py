import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a kernel that never finishes blocks its caller inside the CUDA runtime, where no Python signal handler
    runs) must not hold the GPU box until the session limit: with pytest-timeout present every gpu test gets a 15-minute limit
    enforced from a watchdog thread (stack dump + exit)."""
    if config.pluginmanager.hasplugin("timeout"):
        for it in items:
            if it.get_closest_marker("gpu") and not it.get_closest_marker("timeout"):
                it.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session")
def oracle_lk():
    """The plain-C restatement (oracle/hv_oracle_lk.c); built on demand."""
    import subprocess
    from oracle import lk_oracle
    if not os.path.exists(lk_oracle.ORACLE_SO):
        subprocess.check_call(["make", "-C", ROOT, "oracle"])
    return lk_oracle.OracleLK()


@pytest.fixture(scope="session")
def ref_lk():
    """The compiled reference (oracle/_ref); skipped where it was not built."""
    from oracle import lk_oracle
    if not lk_oracle.have_ref():
        pytest.skip("oracle/_ref/libref_lk.so not built (needs /root/reference)")
    return lk_oracle.RefLK()


@pytest.fixture(scope="session")
def hv():
    """CUDA context through the C ABI. Fails (does not skip) if the library is missing on a GPU box."""
    from hybvio_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
