"""Launches the track-model kernel and one visual-update chain with BASELINE-config shapes, for ncu captures (profiles/):
    ncu --set full --clock-control none --import-source on -k regex:hv_track_model\\|ekf_update_cluster2 -o gpurun_out/r02_track_model \\
        python tools/prof_track_model.py
Launch order: hv_track_model_kernel x reps (150 tracks, one CTA each), then one chain of 8 tracks = 8 x (hv_track_model_kernel (1 CTA),
ekf_update_cluster2_kernel check, ekf_update_cluster2_kernel update)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tri_common  # noqa: E402
from hybvio_b200 import capi  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
base = tri_common.make_track(0, npose=4, stereo=True)
rng = np.random.RandomState(11)
tracks = []
for k in range(150):
    npose = 2 + (k * 7) % 20
    idx = np.concatenate([[0], np.sort(rng.choice(np.arange(1, 21), npose - 1, replace=False))]).astype(np.int32)
    pf = base["pf_true"] * [2.0, 4.0, 8.0, 16.0][k % 4] / 5.0 + rng.normal(0, 0.2, 3)
    ip = tri_common.project(base["m"], idx, base["T1"], base["T2"], True, pf) + rng.normal(0, 1e-3, (2 * npose, 2))
    tracks.append((idx, ip, rng.normal(0, 0.05, ip.shape)))
hv = capi.Context(0)
p = capi.EkfParams()
capi.load().hv_ekf_default_params(ctypes.byref(p))
p.camera_trail_length = 20
ekf = capi.Ekf(hv, p)
A = np.random.RandomState(3).normal(0, 1, (ekf.N, ekf.N))
ekf.upload(m=base["m"], P=1e-4 * (A @ A.T) / ekf.N + np.diag(np.full(ekf.N, 1e-4)))
ekf.set_camera_model(base["T1"], base["T2"], use_stereo=True, estimate_time_shift=True)
for _ in range(reps):
    ekf.track_models(tracks, download=False)
res, succ = ekf.visual_tracks([t for t in tracks if len(t[0]) >= 6][:8], 0.01, 0.004, max_successful_updates=5, lookahead=0)
hv.sync()
print("ok", hv.launches, "launches;", succ, "updates;", [(r["tri_status"], r["outlier_status"]) for r in res])
