"""CPU tests (no GPU): the C EKF oracle (oracle/hv_oracle_ekf.c) against the reference's own unit-test vectors,
the golden trajectory produced by the compiled reference EKF, and -- where oracle/_ref exists -- the compiled
reference itself, op by op."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import ekf_common as C
import ekf_script
from oracle import ekf_oracle


@pytest.fixture(scope="module")
def gold():
    return np.load(C.GOLD)


@pytest.fixture(scope="module")
def orc(oracle_lk):     # oracle_lk builds oracle/libhv_oracle.so on demand
    o = ekf_oracle.OracleEKF()
    o.close()
    return lambda p: ekf_oracle.OracleEKF(p)


def default_params():
    o = ekf_oracle.OracleEKF(); p = o.default_params(); o.close()
    return p


def test_state_dimension_and_defaults(orc):
    e = orc(default_params())
    assert e.N == 160                                   # 20 + 7 * 20 (ekf.cpp:156-158)
    m, P = e.download()
    assert m[6] == 1 and (m[16:19] == 1).all() and np.count_nonzero(m) == 4
    assert np.allclose(np.diag(P)[20:23], 100 ** 2 * 1e4) and np.count_nonzero(P - np.diag(np.diag(P))) == 0
    e.close()
    e = orc(C.params_with(default_params, 6))
    assert e.N == 62
    e.close()


def test_chi2_table_matches_scipy_and_reference_values(orc):
    from scipy.stats import chi2
    e = orc(default_params())
    for n in (1, 2, 8, 20, 40, 84, 160, 200):
        assert abs(e.chi2inv95(n) - chi2.ppf(0.95, n)) < 1e-9 * chi2.ppf(0.95, n)
    assert abs(e.chi2inv95(1) - 3.841458820694124) < 1e-12      # odometry/util.hpp:23, first entries
    assert abs(e.chi2inv95(2) - 5.991464547107981) < 1e-12
    e.close()


def test_reference_unit_test_chi2_kat(orc, gold):
    C.check_reference_chi2_kat(orc, default_params, gold)


def test_reference_unit_test_der_predict(orc, gold):
    C.check_reference_der_predict(orc, default_params, gold)


def test_reference_unit_test_transform_roundtrip(orc, gold):
    C.check_reference_transform_roundtrip(orc, default_params, gold)


def test_oracle_matches_reference_golden_n62(orc, gold):
    C.check_against_golden(orc, default_params, gold, "n62", 6, (8, 20))


def test_oracle_matches_reference_golden_n160(orc, gold):
    C.check_against_golden(orc, default_params, gold, "n160", 20, (8, 20, 40, 84))


@pytest.mark.parametrize("trail,map_size,frames,nlist", [(6, 0, 6, (8, 20)), (20, 0, 6, (8, 40, 84)), (5, 2, 5, (4, 12))])
def test_oracle_vs_compiled_reference_live(orc, trail, map_size, frames, nlist):
    if not ekf_oracle.have_ref():
        pytest.skip("oracle/_ref/libref_ekf.so not built (needs /root/reference)")
    p = C.params_with(default_params, trail, map_size)
    a, b = orc(p), ekf_oracle.RefEKF(p)
    C.check_pair(a, b, frames, nlist)
    a.close(); b.close()


def test_bookkeeping_matches_reference_semantics(orc):
    e = orc(C.params_with(default_params, 3))
    e.predict(10.0, [0, 0, 0], [0, 0, 9.8])           # first sample: no-op, sets times (ekf.cpp:357-370)
    m0, P0 = e.download()
    assert e.platform_time() == 10.0
    e.predict(10.0, [0, 0, 0], [0, 0, 9.8])           # dt <= 0: skipped
    m1, P1 = e.download()
    assert np.array_equal(m0, m1) and np.array_equal(P0, P1)
    e.predict(10.5, [0, 0, 0.1], [0, 0, 9.8])
    assert abs(e.platform_time() - 10.5) < 1e-12 and e.pose_count() == 1
    for k in range(5):
        e.augment(-1)
    assert e.pose_count() == 4                         # capped at trail + 1 (ekf.cpp:877-883)
    assert abs(e.history_time(0) - 10.5) < 1e-12
    e.unaugment()
    assert e.pose_count() == 3
    assert not e.was_stationary()
    e.update_zupt(1e-2)
    assert e.was_stationary()
    e.close()
