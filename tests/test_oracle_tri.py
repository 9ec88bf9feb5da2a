"""CPU tests of the oracle for the per-track measurement model (triangulation + prepareVisualUpdate):
the plain-C restatement (oracle/hv_oracle_tri.c) against golden vectors produced by the reference's own code
(tests/golden/make_golden_tri.py) and, where the compiled reference is present (build container), against it directly on
many more seeded tracks. Tolerances: statuses identical; values 1e-9 relative to the largest entry (fp64, different operation
order inside 3x3 / 3x2 factorisations)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import tri_common  # noqa: E402
import make_golden_tri  # noqa: E402
from oracle import tri_oracle  # noqa: E402

TOL = 1e-9


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if b.size else 0.0


@pytest.fixture(scope="module")
def orc():
    assert os.path.exists(tri_oracle.ORACLE_SO), "run make oracle"
    return tri_oracle.OracleTri()


def check(a, b, where):
    """b: reference, a: oracle"""
    assert (a["tri_status"], a["vu_status"]) == (b["tri_status"], b["vu_status"]), where
    if b["tri_status"] == 0:
        for k in ("pf", "dpf", "H", "f"):
            assert a[k].shape == b[k].shape, (where, k)
            # tracks that pass the rcond gate with derivatives of 1e6 and more (random observations) are ill-conditioned
            tol = TOL if np.abs(b["dpf"]).max() < 1e6 else 1e-6
            assert rel(a[k], b[k]) < tol, (where, k, rel(a[k], b[k]))
        assert abs(a["depth"] - b["depth"]) < TOL * max(1.0, b["depth"]), where
    else:
        assert a["H"].size == 0 and not a["dpf"].any(), where


def test_oracle_matches_reference_golden_vectors(orc):
    g = np.load(os.path.join(HERE, "golden", "tri_golden.npz"))
    cs = make_golden_tri.cases()
    assert int(g["ncases"][0]) == len(cs)
    seen = set()
    for i, (seed, kw, cor) in enumerate(cs):
        t = make_golden_tri.build(seed, kw, cor)
        for ets in (1, 0):
            p = f"c{i}_t{ets}_"
            ref = {"tri_status": int(g[p + "status"][0]), "vu_status": int(g[p + "status"][1]), "pf": g[p + "pf"], "dpf": g[p + "dpf"],
                   "depth": float(g[p + "depth"][0]), "H": g[p + "H"], "f": g[p + "f"]}
            a = orc.track_model(t["m"], t["trail"], t["stereo"], t["idx"], t["T1"], t["T2"], t["ip"], t["vel"], bool(ets))
            check(a, ref, (i, seed, cor, ets))
            seen.add(ref["tri_status"])
    assert seen == {0, 2, 3, 4}        # OK, BEHIND, BAD_COND, NO_CONVERGENCE all exercised


def test_known_answers_of_the_reference_test_suite(orc):
    """The known-answer tests test/triangulation.cpp holds for this path: "visual" (Matlab-generated track, :56-167), "pinv"
    (Matlab pinv, :477-485), "triangulateWithTwoCameras" (right triangle, :487-519); with the reference's own tolerances."""
    import ctypes
    k = tri_common.reference_visual_kat()
    impls = [orc] + ([tri_oracle.RefTri()] if tri_oracle.have_ref() else [])
    for impl in impls:
        o = impl.track_model(k["m"], k["trail"], False, k["idx"], k["T1"], k["T2"], k["ip"], k["vel"], True)
        assert o["tri_status"] == 0 and o["vu_status"] == 0
        assert np.abs(o["pf"] - k["pf_expected"]).sum() < 1e-5
        assert o["H"].shape == (20, 83)                       # 10 observations, truncated after trail slot 8 (CAM + 7 * 8 + 7)
    L = orc.lib
    A = np.array([[1.0, 4.0], [3.0, 2.0], [-1.0, -3.0]])      # = m.transpose() of the reference test
    iA = np.zeros((2, 3))
    L.orc_pinv32(A.ctypes.data_as(ctypes.c_void_p), iA.ctypes.data_as(ctypes.c_void_p))
    expect = np.array([[-0.153333, 0.206667], [0.406667, -0.113333], [0.0666667, -0.133333]]).T
    assert np.abs(iA - expect).sum() < 1e-5
    assert np.abs(iA - np.linalg.pinv(A)).max() < 1e-14
    # test/util.cpp:9-57: quat2rmat / quat2rmat_d against Matlab (q = rotm2quat(rotx(15) roty(3) rotz(-5)))
    q = np.array([0.990310843256666, 0.129225220713441, 0.031619820909086, -0.039817872689419])
    R, dR = np.zeros(9), np.zeros(36)
    L.orc_quat2rmat_d(q.ctypes.data_as(ctypes.c_void_p), R.ctypes.data_as(ctypes.c_void_p), dR.ctypes.data_as(ctypes.c_void_p))
    rmat_e = np.array([0.994829447880333, 0.087036298831283, 0.052335956242944, -0.070691985487699, 0.963430758692103, -0.258464342596353,
                       -0.072917849789463, 0.253428206582672, 0.964602058514480])
    a, b, c, d = 1.980621686513332, 0.258450441426882, 0.063239641818172, 0.079635745378838
    dR_e = np.array([[a, d, c, -d, a, -b, -c, b, a], [b, c, -d, c, -b, -a, -d, a, -b], [-c, b, a, b, c, -d, -a, -d, -c], [d, -a, b, a, d, c, b, c, -d]])
    assert np.abs(R - rmat_e).sum() < 1e-5
    for i in range(4):
        assert np.abs(dR[9 * i:9 * i + 9] - dR_e[i]).sum() < 1e-5
    # test/util.cpp:105-107: the LDLT-based reciprocal condition number of the identity is exactly 1
    L.orc_rcond_ldlt3.restype = ctypes.c_double
    assert L.orc_rcond_ldlt3(np.eye(3).ravel().ctypes.data_as(ctypes.c_void_p)) == 1.0
    pf = np.zeros(3)
    eye = np.eye(3).ravel()
    args = [np.array([1.0, 1.0, 0.0]), eye, np.array([1.0, 2.0, 0.0]), eye, np.array([0.0, 1.0]), np.array([0.0, 0.0]), pf]
    L.orc_two_cameras_point(*[x.ctypes.data_as(ctypes.c_void_p) for x in args])
    assert np.abs(pf - np.array([0.0, 1.0, 1.0])).sum() < 1e-5


def test_measurement_model_is_consistent_with_the_scene(orc):
    """Independent of the reference: on a nearly clean track the triangulated point is the true point, the predicted observations
    are the observations, and H is the total derivative of the predictions w.r.t. the pose states, the re-triangulated point
    included (finite differences through the whole chain).
    Two properties of the reference that parity has to keep (both visible here):
      * exactly noise-free observations never satisfy the relative convergence test |dJ / J| < 1e-2 (triangulation.cpp:337-345)
        and return NO_CONVERGENCE, hence the 1e-5 noise;
      * the conversion of the derivatives back from inverse depth (triangulation.cpp:381-387) leaves out the dependence of the
        first camera's position on the first pose's quaternion through the IMU-camera lever arm (-dR0^T baseline), so the
        quaternion columns of the first pose are exact only for a zero lever arm of the first camera (checked both ways)."""
    t = tri_common.make_track(5, npose=5, stereo=True, noise=0.0)
    rng = np.random.RandomState(1)

    def with_lever_arm(keep):
        T1, T2 = t["T1"].copy(), t["T2"].copy()
        if not keep:
            T2[:3, 3] -= T1[:3, 3]; T1[:3, 3] = 0
        ip = []
        for T in (T1, T2):
            for i in t["idx"]:
                o = 0 if i == 0 else 20 + 7 * (i - 1)
                q = t["m"][6:10] if i == 0 else t["m"][o + 3:o + 7]
                R = T[:3, :3] @ tri_common.quat2rmat(q)
                c = R @ (t["pf_true"] - (t["m"][o:o + 3] - R.T @ T[:3, 3]))
                ip.append(c[:2] / c[2])
        return T1, T2, np.array(ip) + rng.normal(0, 1e-5, (len(ip), 2))

    def run(m, T1, T2, ip):
        return orc.track_model(m, t["trail"], True, t["idx"], T1, T2, ip, t["vel"], True)

    for keep in (False, True):
        T1, T2, ip = with_lever_arm(keep)
        a = run(t["m"], T1, T2, ip)
        assert a["tri_status"] == 0 and a["vu_status"] == 0
        assert np.abs(a["pf"] - t["pf_true"]).max() < 2e-3
        assert np.abs(a["f"] - ip.ravel()).max() < 5e-5
        H = a["H"]
        worst_q0 = 0.0
        for col in (0, 1, 2, 6, 7, 8, 9, 20 + 7 * (int(t["idx"][1]) - 1), 20 + 7 * (int(t["idx"][2]) - 1) + 4):
            eps = 1e-6
            mp, mm_ = t["m"].copy(), t["m"].copy()
            mp[col] += eps; mm_[col] -= eps
            fd = (run(mp, T1, T2, ip)["f"] - run(mm_, T1, T2, ip)["f"]) / (2 * eps)
            err = np.abs(fd - H[:, col]).max()
            if keep and 6 <= col <= 9:
                worst_q0 = max(worst_q0, err)
            else:
                assert err < 2e-8 * max(1.0, np.abs(H[:, col]).max()), (keep, col, err)
        if keep:
            assert 1e-3 < worst_q0 < 5e-2          # the omitted lever-arm term, of the size of |baseline| = 2.3 cm


@pytest.mark.skipif(not tri_oracle.have_ref(), reason="compiled reference (oracle/_ref/libref_tri.so) only exists in the build container")
def test_oracle_matches_compiled_reference_on_many_tracks(orc):
    ref = tri_oracle.RefTri()
    statuses = set()
    for seed in range(240):
        kw = dict(npose=2 + seed % 9, stereo=seed % 2 == 0, noise=[1e-3, 3e-3, 1e-2][seed % 3], depth=[2, 5, 15, 40][seed % 4])
        t = tri_common.make_track(seed, **kw)
        tri_common.corrupt(t, ["none", "none", "outlier", "flip", "static", "garbage"][(seed // 2) % 6], seed)
        for ets in (True, False):
            b = ref.track_model(t["m"], t["trail"], t["stereo"], t["idx"], t["T1"], t["T2"], t["ip"], t["vel"], ets)
            a = orc.track_model(t["m"], t["trail"], t["stereo"], t["idx"], t["T1"], t["T2"], t["ip"], t["vel"], ets)
            check(a, b, (seed, ets))
            statuses.add(b["tri_status"])
    assert {0, 2, 3}.issubset(statuses)
