import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


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
