"""GPU parity of the per-track measurement model on the device (hv_ekf_track_models: triangulation + prepareVisualUpdate,
SURVEY.md 8(f) N1) against the C oracle (oracle/hv_oracle_tri.c, itself pinned against the compiled reference and its golden
vectors in test_oracle_tri.py), through the C ABI. Statuses identical; pf, d pf, H, f within 1e-9 relative (fp64, different
summation order). The file sorts after the other GPU tests on purpose: this row was added after the round's last GPU session
and has so far only run on the host emulator (tests/test_emu_kernels.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import tri_common  # noqa: E402
import make_golden_tri  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if b.size else 0.0


@pytest.fixture(scope="module")
def env():
    from hybvio_b200 import capi
    from oracle import tri_oracle
    hv = capi.Context(0)
    yield capi, hv, tri_oracle.OracleTri()
    hv.close()


def make_ekf(capi, hv, t):
    import ctypes
    p = capi.EkfParams()
    capi.load().hv_ekf_default_params(ctypes.byref(p))
    p.camera_trail_length = t["trail"]
    e = capi.Ekf(hv, p)
    e.upload(m=t["m"])
    return e


def compare(dev, orc, where):
    assert (dev["tri_status"], dev["vu_status"]) == (orc["tri_status"], orc["vu_status"]), where
    if orc["tri_status"] != 0:
        assert dev["rows"] == 0 and not dev["dpf"].any(), where
        return
    tol = TOL if np.abs(orc["dpf"]).max() < 1e6 else 1e-6
    assert dev["H"].shape == orc["H"].shape, where
    for k in ("pf", "dpf", "H", "f"):
        assert rel(dev[k], orc[k]) < tol, (where, k, rel(dev[k], orc[k]))
    assert abs(dev["depth"] - orc["depth"]) < tol * max(1.0, orc["depth"]), where


@pytest.mark.parametrize("stereo", [True, False])
@pytest.mark.parametrize("time_shift", [True, False])
def test_track_models_match_oracle(env, stereo, time_shift):
    """One launch, many tracks against the same state: clean and spoiled tracks, 2..10 poses each."""
    capi, hv, orc = env
    base = tri_common.make_track(0, npose=4, stereo=stereo)
    e = make_ekf(capi, hv, base)
    e.set_camera_model(base["T1"], base["T2"], use_stereo=stereo, estimate_time_shift=time_shift)
    tracks, expect = [], []
    rng = np.random.RandomState(5)
    for k in range(96):
        # every track of a launch shares the state (and rig) of `base`: rebuild the observations of a new point from it
        t = tri_common.make_track(0, npose=2 + k % 9, stereo=stereo, noise=[1e-3, 3e-3, 1e-2][k % 3], depth=[2, 5, 15, 40][k % 4])
        assert np.array_equal(t["m"], base["m"])
        idx = np.concatenate([[0], np.sort(rng.choice(np.arange(1, 21), len(t["idx"]) - 1, replace=False))]).astype(np.int32)
        ip = tri_common.project(base["m"], idx, base["T1"], base["T2"], stereo, t["pf_true"] + rng.normal(0, 0.3, 3)) + rng.normal(0, 1e-3, (len(idx) * (2 if stereo else 1), 2))
        t = dict(t, idx=idx, ip=ip, vel=rng.normal(0, 0.05, ip.shape))
        tri_common.corrupt_observations(t, ["none", "none", "none", "outlier", "flip", "garbage"][k % 6], 100 + k)
        tracks.append((t["idx"], t["ip"], t["vel"]))
        expect.append(orc.track_model(base["m"], base["trail"], stereo, t["idx"], base["T1"], base["T2"], t["ip"], t["vel"], time_shift))
    got = e.track_models(tracks)
    seen = set()
    for k, (g, o) in enumerate(zip(got, expect)):
        compare(g, o, (stereo, time_shift, k))
        seen.add(o["tri_status"])
    assert {0, 2}.issubset(seen)
    e.close()


def test_track_models_match_reference_golden_vectors(env):
    """The committed outputs of the reference's own triangulation.cpp (tests/golden/tri_golden.npz)."""
    capi, hv, _ = env
    g = np.load(os.path.join(HERE, "golden", "tri_golden.npz"))
    seen = set()
    for i, (seed, kw, cor) in enumerate(make_golden_tri.cases()):
        t = make_golden_tri.build(seed, kw, cor)
        e = make_ekf(capi, hv, t)
        for ets in (1, 0):
            p = f"c{i}_t{ets}_"
            ref = {"tri_status": int(g[p + "status"][0]), "vu_status": int(g[p + "status"][1]), "pf": g[p + "pf"], "dpf": g[p + "dpf"],
                   "depth": float(g[p + "depth"][0]), "H": g[p + "H"], "f": g[p + "f"]}
            e.set_camera_model(t["T1"], t["T2"], use_stereo=t["stereo"], estimate_time_shift=bool(ets))
            dev = e.track_models([(t["idx"], t["ip"], t["vel"])])[0]
            compare(dev, ref, (i, seed, cor, ets))
            seen.add(ref["tri_status"])
        e.close()
    assert seen == {0, 2, 3, 4}


def test_known_answer_of_the_reference_test_suite(env):
    """TEST_CASE "visual" of the reference's test/triangulation.cpp (:56-167): a Matlab-generated track; the reference requires
    TriangulatorStatus::OK and sum |pf - pf_e| < 1e-5."""
    capi, hv, orc = env
    k = tri_common.reference_visual_kat()
    e = make_ekf(capi, hv, k)
    e.set_camera_model(k["T1"], None, use_stereo=False, estimate_time_shift=True)
    d = e.track_models([(k["idx"], k["ip"], k["vel"])])[0]
    assert d["tri_status"] == 0 and d["vu_status"] == 0 and d["H"].shape == (20, 83)
    assert np.abs(d["pf"] - k["pf_expected"]).sum() < 1e-5
    compare(d, orc.track_model(k["m"], k["trail"], False, k["idx"], k["T1"], k["T2"], k["ip"], k["vel"], True), "visual KAT")
    e.close()


def test_device_H_feeds_the_outlier_check_and_update(env):
    """H, f, y never leave the device: check + update from the pointers hv_ekf_track_models returns equals the same calls with
    the oracle's H uploaded from the host."""
    capi, hv, orc = env
    t = tri_common.make_track(3, npose=6, stereo=True)
    a, b = make_ekf(capi, hv, t), make_ekf(capi, hv, t)
    A = np.random.RandomState(3).normal(0, 1, (a.N, a.N))
    P0 = 1e-4 * (A @ A.T) / a.N + np.diag(np.full(a.N, 1e-4))          # a well-conditioned prior: the comparison is about H, not about the filter
    a.upload(P=P0); b.upload(P=P0)
    a.set_camera_model(t["T1"], t["T2"], use_stereo=True)
    d = a.track_models([(t["idx"], t["ip"], t["vel"])])[0]
    o = orc.track_model(t["m"], t["trail"], True, t["idx"], t["T1"], t["T2"], t["ip"], t["vel"], True)
    assert d["tri_status"] == 0 and d["vu_status"] == 0
    st_a, chi2_a = a.visual_track(d, 0.02, 5.0, mode=2)          # check, then update if inlier, on the device-resident H
    st_b, chi2_b, _ = b.visual_check_update(o["H"], o["f"], t["ip"].ravel(), 0.02, 5.0)
    ma, Pa = a.download(); mb, Pb = b.download()
    assert st_a == st_b and abs(chi2_a - chi2_b) < 1e-9 * max(1.0, abs(chi2_b))
    assert np.abs(ma - mb).max() < 1e-9 and np.abs(Pa - Pb).max() / np.abs(Pb).max() < 1e-9
    # the asynchronous update-only mode on a second pair of filters
    a2, b2 = make_ekf(capi, hv, t), make_ekf(capi, hv, t)
    a2.upload(P=P0); b2.upload(P=P0)
    a2.set_camera_model(t["T1"], t["T2"], use_stereo=True)
    d2 = a2.track_models([(t["idx"], t["ip"], t["vel"])], download=False)[0]
    assert a2.visual_track(d2, 0.02, mode=1) is None
    b2.visual_update(o["H"], o["f"], t["ip"].ravel(), 0.02)
    m2a, P2a = a2.download(); m2b, P2b = b2.download()
    assert np.abs(m2a - m2b).max() < 1e-9 and np.abs(P2a - P2b).max() / np.abs(P2b).max() < 1e-9
    a2.close(); b2.close()
    a.close(); b.close()


def sequential_reference_flow(orc, okf, m0_tracks, base, chi_r, vis_r, max_succ):
    """The per-track loop of Session::trackerVisualUpdate (backend.cpp:1012-1252, per-track mode) with the oracles: every track
    is modelled against the state the previous track's update left behind."""
    out, succ = [], 0
    for idx, ip, vel in m0_tracks:
        if succ >= max_succ:
            out.append({"tri_status": -1, "vu_status": -1, "outlier_status": 1, "updated": False})
            continue
        m, _ = okf.download()
        o = orc.track_model(m, base["trail"], base["stereo"], idx, base["T1"], base["T2"], ip, vel, True)
        r = {"tri_status": o["tri_status"], "vu_status": o["vu_status"], "outlier_status": 1, "updated": False, "pf": o["pf"]}
        if o["tri_status"] == 0 and o["vu_status"] == 0:
            st, chi2 = okf.visual_check(o["H"], o["f"], np.asarray(ip).ravel(), chi_r, -1.0)
            r["outlier_status"], r["chi2"] = st, chi2
            if st == 0:
                okf.visual_update(o["H"], o["f"], np.asarray(ip).ravel(), vis_r)
                r["updated"] = True
                succ += 1
        out.append(r)
    return out, succ


def chain_tracks(base, count, seed):
    rng = np.random.RandomState(seed)
    tracks = []
    for k in range(count):
        npose = 4 + (k * 5) % 9
        idx = np.concatenate([[0], np.sort(rng.choice(np.arange(1, 21), npose - 1, replace=False))]).astype(np.int32)
        pf = base["pf_true"] + rng.normal(0, 0.4, 3)
        ip = tri_common.project(base["m"], idx, base["T1"], base["T2"], base["stereo"], pf)
        ip = ip + rng.normal(0, 2e-3, ip.shape)
        if k % 4 == 1:
            ip[rng.randint(len(ip))] += [0.08, -0.06]          # a gross outlier: rejected by the chi2 test (or by the triangulation)
        if k % 7 == 3:
            ip = -ip                                           # behind the cameras: no measurement model
        tracks.append((idx, ip, rng.normal(0, 0.05, ip.shape)))
    return tracks


@pytest.mark.parametrize("lookahead", [0, 3])
def test_device_gated_chain_equals_the_per_track_loop(env, lookahead):
    """hv_ekf_visual_tracks (model -> check -> update if inlier, control flow on the device, one synchronisation per `lookahead`
    tracks) against the same loop driven track by track with the oracles."""
    capi, hv, orc = env
    from oracle import ekf_oracle
    base = tri_common.make_track(7, npose=4, stereo=True)
    tracks = chain_tracks(base, 14, 21)
    e = make_ekf(capi, hv, base)
    okf = ekf_oracle.OracleEKF(e.params)
    rngP = np.random.RandomState(3)
    A = rngP.normal(0, 1, (e.N, e.N))
    P0 = 1e-4 * (A @ A.T) / e.N + np.diag(np.full(e.N, 1e-4))
    e.upload(m=base["m"], P=P0); okf.upload(m=base["m"], P=P0)
    e.set_camera_model(base["T1"], base["T2"], use_stereo=True)
    chi_r, vis_r, max_succ = 0.01, 0.004, 5
    exp, exp_succ = sequential_reference_flow(orc, okf, tracks, base, chi_r, vis_r, max_succ)
    got, succ = e.visual_tracks(tracks, chi_r, vis_r, max_successful_updates=max_succ, lookahead=lookahead)
    assert succ == exp_succ
    assert exp_succ == max_succ and any(x["outlier_status"] == 3 for x in exp) and any(x["tri_status"] == 2 for x in exp) and exp[-1]["tri_status"] == -1
    for k, (g, x) in enumerate(zip(got, exp)):
        assert (g["tri_status"], g["vu_status"], g["outlier_status"], g["updated"]) == (x["tri_status"], x["vu_status"], x["outlier_status"], x["updated"]), (k, g, x)
        if x["tri_status"] == 0:
            assert np.abs(g["pf"] - x["pf"]).max() < 1e-9 * max(1.0, np.abs(x["pf"]).max())
        if "chi2" in x:
            assert abs(g["chi2"] - x["chi2"]) < 1e-8 * max(1.0, abs(x["chi2"]))
    ma, Pa = e.download(); mb, Pb = okf.download()
    assert np.abs(ma - mb).max() < 1e-9 and np.abs(Pa - Pb).max() / np.abs(Pb).max() < 1e-9
    e.close(); okf.close()


def test_full_frame_properties(env):
    """BASELINE config 2 size (150 stereo tracks, 2..21 poses, trail 20) through properties that need no oracle: a track's result does
    not depend on its position in the batch (bitwise), H is zero in the columns of poses the track does not touch and its column
    count is the truncation point of prepareVisualUpdate, f reproduces the observations to the noise level, repeated calls agree bitwise."""
    capi, hv, _ = env
    base = tri_common.make_track(0, npose=4, stereo=True)
    rng = np.random.RandomState(11)
    tracks = []
    for k in range(150):
        npose = 2 + (k * 7) % 20
        idx = np.concatenate([[0], np.sort(rng.choice(np.arange(1, 21), npose - 1, replace=False))]).astype(np.int32)
        pf = base["pf_true"] * [2.0, 4.0, 8.0, 16.0][k % 4] / 5.0 + rng.normal(0, 0.2, 3)
        ip = tri_common.project(base["m"], idx, base["T1"], base["T2"], True, pf) + rng.normal(0, 1e-3, (2 * npose, 2))
        tracks.append((idx, ip, rng.normal(0, 0.05, ip.shape)))
    e = make_ekf(capi, hv, base)
    e.set_camera_model(base["T1"], base["T2"], use_stereo=True)
    a = e.track_models(tracks)
    perm = rng.permutation(150)
    b = e.track_models([tracks[i] for i in perm])
    c = e.track_models(tracks)
    ok = 0
    for k in range(150):
        x, y, z = a[k], b[int(np.where(perm == k)[0][0])], c[k]
        for other in (y, z):
            assert (x["tri_status"], x["vu_status"]) == (other["tri_status"], other["vu_status"])
            assert np.array_equal(x["H"], other["H"]) and np.array_equal(x["f"], other["f"]) and np.array_equal(x["pf"], other["pf"])
        if x["tri_status"] != 0:
            continue
        ok += 1
        idx = tracks[k][0]
        used = np.zeros(x["cols"], bool)
        used[0:3] = used[6:10] = True
        used[19] = True                                          # the time-shift column
        for i in idx[1:]:
            used[20 + 7 * (i - 1):20 + 7 * i] = True
        assert x["cols"] == 20 + 7 * int(idx.max()) and x["rows"] == 4 * len(idx)
        assert not x["H"][:, ~used].any() and np.abs(x["H"][:, used]).max() > 0
        assert np.abs(x["f"] - tracks[k][1].ravel()).max() < 2e-2
    assert ok >= 140
    e.close()


def test_track_models_reject_bad_input(env):
    capi, hv, _ = env
    t = tri_common.make_track(1, npose=4, stereo=False)
    e = make_ekf(capi, hv, t)
    with pytest.raises(RuntimeError):
        e.track_models([(t["idx"], t["ip"], t["vel"])])                # camera model not set
    e.set_camera_model(t["T1"], None, use_stereo=False)
    with pytest.raises(RuntimeError):
        e.track_models([(t["idx"][:1], t["ip"][:1], t["vel"][:1])])    # a single pose
    with pytest.raises(RuntimeError):
        e.track_models([(np.array([0, 25], np.int32), t["ip"][:2], t["vel"][:2])])   # index beyond the trail
    e.close()


def test_track_model_through_the_reference_interfaces():
    """oracle/ref_build/track_model_iface_test.cpp: the reference's unmodified extractCameraPoseTrail / Triangulator::triangulate /
    prepareVisualUpdate / EKF::visualTrackOutlierCheck / updateVisualTrack (H built on the host) against cudaTrackModels /
    cudaVisualTrackOutlierCheck / cudaUpdateVisualTrack (hybvio_b200/host/cuda_track_model.hpp), both on the CUDA EKF."""
    import subprocess
    root = os.path.dirname(HERE)
    exe = os.path.join(root, "oracle", "_ref", "run_track_model_iface_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/run_track_model_iface_test not built (needs /root/reference at build time)")
    r = subprocess.run([exe], cwd=os.path.join(root, "oracle", "_ref"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "track model interface: all ok" in r.stdout, r.stdout[-500:]
