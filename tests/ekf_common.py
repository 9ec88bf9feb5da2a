"""Back-end independent EKF checks used by both the CPU (oracle) and the GPU (CUDA) test modules. `make(params)`
builds a filter of the back end under test."""
import os

import numpy as np

import ekf_script

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ekf_golden.npz")
# Tolerances (BASELINE.json north_star: pose within 1e-4 m; SURVEY.md 8(d): covariance relative <= 1e-9).
TOL_M = 1e-9          # absolute, on every state entry (positions in metres: 5 orders tighter than the 1e-4 m gate)
TOL_P_REL = 1e-9      # max |dP| / max |P|


def params_with(make_default, trail, map_size=0, noise_scale=None):
    p = make_default()
    p.camera_trail_length = trail
    p.hybrid_map_size = map_size
    if noise_scale is not None:
        p.noise_scale = noise_scale
    return p


def check_reference_chi2_kat(make, default_params, gold):
    """test/ekf.cpp:19-71: v' M^-1 v = 1.7626 +- 0.1, through visualTrackOutlierCheck's statistic (ekf.cpp:815)."""
    M, v = gold["reftest_M"], gold["reftest_v"]
    e = make(params_with(default_params, 5, noise_scale=1.0))
    r = 1e-3
    P = np.zeros((e.N, e.N)); P[:20, :20] = M - (r * r) * np.eye(20)
    P[20:, 20:] = np.eye(e.N - 20)
    e.upload(np.zeros(e.N), P)
    st, chi2 = e.visual_check(np.eye(20), np.zeros(20), v, r)
    assert abs(chi2 - 1.7626) < 1e-1
    assert abs(chi2 - float(v @ np.linalg.solve(M, v))) < 1e-9
    assert st == 0    # 1.76 << chi2inv95[20] = 31.4
    e.close()


def check_reference_der_predict(make, default_params, gold):
    """test/ekf.cpp:73-117: predict()'s analytic Jacobian (getDydx) vs central differences of its mean, < 1e-3."""
    poses, gyro, acc = gold["reftest_poses"], gold["reftest_gyro"], gold["reftest_acc"]
    e0 = make(params_with(default_params, 5))
    m = np.zeros(e0.N)
    m[0:3] = poses[0:3]; m[6:10] = poses[3:7]
    for i in range(5):
        m[20 + 7 * i: 27 + 7 * i] = poses[(i + 1) * 7:(i + 2) * 7]
    t, dt = 0.01, 0.01
    e0.set_first_sample_time(t)

    def run(x):
        c = e0.clone()
        mm = c.download()[0]; mm[:20] = x
        c.upload(mm, None)
        c.predict(t + dt, gyro, acc)
        out = c.download()[0][:20].copy(), c.get_dydx()
        c.close()
        return out

    x0 = m[:20].copy()
    _, analytic = run(x0)
    h = 1e-7
    numeric = np.zeros((20, 20))
    for j in range(20):
        xp, xm = x0.copy(), x0.copy(); xp[j] += h; xm[j] -= h
        numeric[:, j] = (run(xp)[0] - run(xm)[0]) / (2 * h)
    assert np.abs(analytic - numeric).max() < 1e-3
    e0.close()


def check_reference_transform_roundtrip(make, default_params, gold):
    """test/ekf.cpp:119-145 with the reference's fixtures test/data/P.csv, m.csv."""
    P0, m0 = gold["reftest_P0"], gold["reftest_m0"]
    e = make(params_with(default_params, 5))
    e.upload(m0, P0)
    A = 2
    pos0, rot0 = m0[20 + 7 * A:23 + 7 * A].copy(), m0[23 + 7 * A:27 + 7 * A].copy()
    e.transform_to([0, 1, 0], [1, 0, 0, 0], A)
    m1, _ = e.download()
    assert np.linalg.norm(m1[20 + 7 * A:23 + 7 * A] - [0, 1, 0]) < 1e-6
    assert np.linalg.norm(m1[23 + 7 * A:27 + 7 * A] - [1, 0, 0, 0]) < 1e-6
    e.transform_to(pos0, rot0, A)
    m2, P2 = e.download()
    assert np.linalg.norm(m2 - m0) < 1e-3 and np.linalg.norm(P2 - P0) < 1e-3
    e.close()


def check_against_golden(make, default_params, gold, name, trail, nlist, fused=False):
    e = make(params_with(default_params, trail))
    frames = int(gold[f"{name}_frames"])
    snaps, checks = [], []
    t = ekf_script.run_frames(e, frames=frames, n_list=nlist, snapshots=snaps, checks=checks, fused=fused)
    assert np.array_equal(np.array([c[0] for c in checks], np.int32), gold[f"{name}_check_status"])
    for i in range(len(snaps)):
        if f"{name}_m_{i}" in gold:
            assert np.abs(snaps[i][0] - gold[f"{name}_m_{i}"]).max() < TOL_M, f"{name} frame {i}: m"
            assert ekf_script.rel_err(snaps[i][1], gold[f"{name}_P_{i}"]) < TOL_P_REL, f"{name} frame {i}: P"
    misc = []
    ekf_script.run_misc_ops(e, misc, t)
    assert len(misc) == int(gold[f"{name}_misc_count"])
    for i, (m, P) in enumerate(misc):
        assert np.abs(m - gold[f"{name}_misc_m_{i}"]).max() < TOL_M, f"{name} misc op {i}: m"
        if f"{name}_misc_P_{i}" in gold:
            assert ekf_script.rel_err(P, gold[f"{name}_misc_P_{i}"]) < TOL_P_REL, f"{name} misc op {i}: P"
    e.close()


def check_pair(a, b, frames, nlist, fused_a=False, map_size=0, TOL_M=TOL_M, TOL_P_REL=TOL_P_REL):
    """Runs the same script on two live back ends and compares after every frame / op."""
    sa, sb, ca, cb = [], [], [], []
    ta = ekf_script.run_frames(a, frames=frames, n_list=nlist, snapshots=sa, checks=ca, fused=fused_a)
    tb = ekf_script.run_frames(b, frames=frames, n_list=nlist, snapshots=sb, checks=cb)
    assert [c[0] for c in ca] == [c[0] for c in cb]
    for (x, y) in zip(ca, cb):
        if x[1] is not None and y[1] is not None and y[1] > 0:
            assert abs(x[1] - y[1]) <= 1e-8 * max(1.0, abs(y[1]))
    for i, ((ma, Pa), (mb, Pb)) in enumerate(zip(sa, sb)):
        assert np.abs(ma - mb).max() < TOL_M, f"frame {i}: m differs by {np.abs(ma - mb).max()}"
        assert ekf_script.rel_err(Pa, Pb) < TOL_P_REL, f"frame {i}: P rel err {ekf_script.rel_err(Pa, Pb)}"
    ma_, mb_ = [], []
    ekf_script.run_misc_ops(a, ma_, ta); ekf_script.run_misc_ops(b, mb_, tb)
    for i, ((ma, Pa), (mb, Pb)) in enumerate(zip(ma_, mb_)):
        assert np.abs(ma - mb).max() < TOL_M, f"misc {i}: m differs by {np.abs(ma - mb).max()}"
        assert ekf_script.rel_err(Pa, Pb) < TOL_P_REL, f"misc {i}: P rel err {ekf_script.rel_err(Pa, Pb)}"
