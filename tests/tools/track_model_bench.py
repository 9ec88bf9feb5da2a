"""Measures the per-track measurement-model kernel (hv_track_model_kernel: triangulation + prepareVisualUpdate on the device,
SURVEY.md 8(f) N1) on cuda:0 and checks it against the C oracle in the same run. Prints ONE JSON object. bench.py runs this as a
separate process after its own measurement (a problem here cannot disturb the headline line); it can also be run by hand:
    python tests/tools/track_model_bench.py [--tracks 150] [--reps 50]
Workload: the tracks of one EuRoC-shaped stereo frame (BASELINE config 2: 150 tracks, trail 20, stereo), 2..21 poses per track,
all evaluated against one resident state in one launch (one CTA per track)."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cpu_reference_loop(tracks, base, p, P0, chi_r, vis_r, reps=3):
    """The same loop with the reference's own code on the host: triangulation.cpp + prepareVisualUpdate (oracle/_ref/libref_tri.so)
    and ekf.cpp (libref_ekf.so), one thread, as Session::trackerVisualUpdate runs it. None where oracle/_ref was not built."""
    from oracle import ekf_oracle, tri_oracle
    if not (tri_oracle.have_ref() and os.path.exists(ekf_oracle.REF_SO)):
        return None
    tri, best, dec = tri_oracle.RefTri(), [], []
    for _ in range(reps):
        kf = ekf_oracle.RefEKF(p)
        kf.upload(m=base["m"], P=P0)
        dec, succ = [], 0
        t0 = time.perf_counter()
        for idx, ip, vel in tracks:
            if succ >= 5:
                break
            m, _P = kf.download()
            o = tri.track_model(m, base["trail"], True, idx, base["T1"], base["T2"], ip, vel, True)
            st = 1
            if o["tri_status"] == 0 and o["vu_status"] == 0:
                st, _ = kf.visual_check(o["H"], o["f"], np.asarray(ip).ravel(), chi_r, -1.0)
                if st == 0:
                    kf.visual_update(o["H"], o["f"], np.asarray(ip).ravel(), vis_r)
                    succ += 1
            dec.append((o["tri_status"], st))
        best.append(time.perf_counter() - t0)
        kf.close()
    return {"us": round(float(np.median(best)) * 1e6, 1), "successful_updates": succ, "decisions": dec,
            "note": "reference triangulation.cpp + ekf.cpp through ctypes, one host thread; includes a state download per track"}


def visual_update_loop(capi, hv, base, p, ntracks=20, reps=20):
    """Session::trackerVisualUpdate's per-track loop (backend.cpp:1012-1252; 20 candidate tracks, at most 5 successful updates)
    three ways through the C ABI, same tracks and start state: (a) device-gated chain hv_ekf_visual_tracks, one synchronisation;
    (b) the same chain synchronising every 4 tracks; (c) track by track: hv_ekf_track_models -> hv_ekf_visual_track(check) ->
    hv_ekf_visual_track(update), two synchronisations per track. Wall-clock per loop incl. ctypes; results must agree."""
    import tri_common
    rng = np.random.RandomState(5)
    tracks = []
    for k in range(ntracks):
        npose = 4 + (k * 5) % 9
        idx = np.concatenate([[0], np.sort(rng.choice(np.arange(1, 21), npose - 1, replace=False))]).astype(np.int32)
        ip = tri_common.project(base["m"], idx, base["T1"], base["T2"], True, base["pf_true"] + rng.normal(0, 0.4, 3))
        ip = ip + rng.normal(0, 2e-3, ip.shape)
        if k % 3 == 1:
            ip[rng.randint(len(ip))] += [0.08, -0.06]
        tracks.append((idx, ip, rng.normal(0, 0.05, ip.shape)))
    ekf = capi.Ekf(hv, p)
    ekf.set_camera_model(base["T1"], base["T2"], use_stereo=True, estimate_time_shift=True)
    A = np.random.RandomState(3).normal(0, 1, (ekf.N, ekf.N))
    P0 = 1e-4 * (A @ A.T) / ekf.N + np.diag(np.full(ekf.N, 1e-4))
    chi_r, vis_r = 0.01, 0.004

    def reset():
        ekf.upload(m=base["m"], P=P0)
        ekf.flush()
        hv.sync()

    def chain(lookahead):
        return ekf.visual_tracks(tracks, chi_r, vis_r, max_successful_updates=5, lookahead=lookahead)

    def per_track():
        succ, res = 0, []
        for t in tracks:
            if succ >= 5:
                break
            d = ekf.track_models([t], download=False)[0]
            st = 1
            if d["tri_status"] == 0 and d["vu_status"] == 0:
                st, _ = ekf.visual_track(d, chi_r, mode=0)
                if st == 0:
                    ekf.visual_track(d, vis_r, mode=1)
                    succ += 1
            res.append((d["tri_status"], st))
        return res, succ

    def timed(fn):
        best, state, ret = [], None, None
        for i in range(reps + 2):
            reset()
            t0 = time.perf_counter()
            ret = fn()
            hv.sync()
            dt = time.perf_counter() - t0
            if i >= 2:
                best.append(dt)
            state = ekf.download()
        return float(np.median(best)) * 1e6, state, ret

    errors = {}

    def attempt(name, fn):
        try:
            return timed(fn)
        except Exception as ex:       # noqa: BLE001 -- keep whatever else could be measured
            errors[name] = repr(ex)[:300]
            return None, None, None

    try:
        cpu = cpu_reference_loop(tracks, base, p, P0, chi_r, vis_r)
    except Exception as ex:           # noqa: BLE001
        cpu, errors["cpu_reference_loop"] = None, repr(ex)[:300]
    us_a, st_a, ret_a = attempt("chain_one_sync", lambda: chain(0))
    us_b, st_b, _ = attempt("chain_sync_every_4", lambda: chain(4))
    us_c, st_c, ret_c = attempt("per_track_calls", per_track)
    try:
        ekf.close()
    except Exception:             # noqa: BLE001
        pass
    out = {"tracks": ntracks, "max_successful_updates": 5,
           "chain_one_sync_us": None if us_a is None else round(us_a, 1), "chain_sync_every_4_us": None if us_b is None else round(us_b, 1),
           "per_track_calls_us": None if us_c is None else round(us_c, 1), "cpu_reference_loop": cpu,
           "variant": "model + fused check/update per track",
           "note": "median wall-clock of the whole loop through ctypes, state re-uploaded before every repetition (outside the timed region)"}
    if ret_a is not None:
        out["successful_updates"] = ret_a[1]
        seq = [(r["tri_status"], r["outlier_status"]) for r in ret_a[0]]
        if ret_c is not None:
            out["same_decisions"] = bool(seq[:len(ret_c[0])] == ret_c[0] and ret_a[1] == ret_c[1])
            out["max_state_difference_chain_vs_per_track"] = float(max(np.abs(st_a[0] - st_c[0]).max(), np.abs(st_b[0] - st_c[0]).max() if st_b is not None else 0.0))
        if cpu is not None:
            out["same_decisions_as_cpu_reference"] = bool([tuple(x) for x in cpu["decisions"]] == seq[:len(cpu["decisions"])])
    if errors:
        out["errors"] = errors
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=150)
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    import tri_common
    from hybvio_b200 import capi
    from oracle import tri_oracle

    base = tri_common.make_track(0, npose=4, stereo=True)
    rng = np.random.RandomState(11)
    tracks = []
    for k in range(args.tracks):
        npose = 2 + (k * 7) % 20                       # 2..21 poses: the whole range of a trail-20 filter
        idx = np.concatenate([[0], np.sort(rng.choice(np.arange(1, 21), npose - 1, replace=False))]).astype(np.int32)
        depth = [2.0, 4.0, 8.0, 16.0][k % 4]
        pf = base["pf_true"] * depth / 5.0 + rng.normal(0, 0.2, 3)
        ip = tri_common.project(base["m"], idx, base["T1"], base["T2"], True, pf) + rng.normal(0, 1e-3, (2 * npose, 2))
        tracks.append((idx, ip, rng.normal(0, 0.05, ip.shape)))
    hv = capi.Context(0)
    p = capi.EkfParams()
    capi.load().hv_ekf_default_params(ctypes.byref(p))
    p.camera_trail_length = base["trail"]
    ekf = capi.Ekf(hv, p)
    ekf.upload(m=base["m"])
    ekf.set_camera_model(base["T1"], base["T2"], use_stereo=True, estimate_time_shift=True)
    got = ekf.track_models(tracks)                     # warm-up + results for the parity check
    for _ in range(3):
        ekf.track_models(tracks, download=False)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        ekf.track_models(tracks, download=False)
    call_us = (time.perf_counter() - t0) / args.reps * 1e6
    kernel_us = ekf.track_models_time(args.reps)
    # parity against the oracle (checker only)
    orc = tri_oracle.OracleTri()
    worst, mism, ok_tracks, h_bytes, obs = 0.0, 0, 0, 0, 0
    t0 = time.perf_counter()
    exp = [orc.track_model(base["m"], base["trail"], True, i, base["T1"], base["T2"], ip, v, True) for i, ip, v in tracks]
    port_us = (time.perf_counter() - t0) / len(tracks) * 1e6
    for g, o in zip(got, exp):
        if (g["tri_status"], g["vu_status"]) != (o["tri_status"], o["vu_status"]):
            mism += 1
            continue
        if o["tri_status"] == 0:
            ok_tracks += 1
            h_bytes += 8 * o["H"].size
            obs += o["H"].shape[0] // 2
            for key in ("H", "f", "pf", "dpf"):
                worst = max(worst, float(np.abs(g[key] - o[key]).max() / max(np.abs(o[key]).max(), 1e-300)))
    out = {"kernel": "hv_track_model_kernel", "tracks": len(tracks), "stereo": True, "trail": base["trail"],
           "kernel_us_per_launch": round(kernel_us, 2), "kernel_us_per_track": round(kernel_us / len(tracks), 3),
           "tracks_per_s_kernel": round(len(tracks) / kernel_us * 1e6, 1),
           "call_us": round(call_us, 2), "call_note": "hv_ekf_track_models through ctypes: pack + H2D of the observations, launch, D2H of statuses, synchronise",
           "algo_bytes_per_launch": int(h_bytes + 8 * 160 * len(tracks) + 8 * 4 * obs),
           "gbs": round((h_bytes + 8 * 160 * len(tracks) + 8 * 4 * obs) / kernel_us * 1e-3, 2),
           "bytes_note": "H written (2 n_obs x l x 8 B per triangulated track) + state mean read per CTA + observations; the kernel is bound by the Gauss-Newton dependency chain, not by HBM",
           "parity": {"checker": "oracle/hv_oracle_tri.c", "status_mismatches": mism, "triangulated": ok_tracks, "worst_relative_difference": worst},
           "cpu_port_us_per_track": round(port_us, 2)}
    if tri_oracle.have_ref():
        ref = tri_oracle.RefTri()
        t0 = time.perf_counter()
        for i, ip, v in tracks:
            ref.track_model(base["m"], base["trail"], True, i, base["T1"], base["T2"], ip, v, True)
        out["cpu_reference_us_per_track"] = round((time.perf_counter() - t0) / len(tracks) * 1e6, 2)
        out["cpu_reference_note"] = "the reference's own triangulation.cpp + prepareVisualUpdate (oracle/_ref/libref_tri.so), one host thread (-O2, as oracle/ref_build/build_tri.sh compiles it)"
    try:
        out["visual_update_loop"] = visual_update_loop(capi, hv, base, p)
    except Exception as ex:       # noqa: BLE001 -- keep the kernel numbers above even if the loop comparison fails
        out["visual_update_loop"] = {"error": repr(ex)[:300]}
    print(json.dumps(out), flush=True)
    try:
        ekf.close(); hv.close()
    except Exception:             # noqa: BLE001 -- a sticky CUDA error from a failed variant must not eat the line above
        pass


if __name__ == "__main__":
    main()
