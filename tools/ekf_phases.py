"""Per-phase timing of the cluster update kernel (needs a library built with -DHV_EKF_TIMING: make TIMING=1)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hybvio_b200 import capi
hv = capi.Context(0)
p = capi.EkfParams(); capi.load().hv_ekf_default_params(ctypes.byref(p))
ekf = capi.Ekf(hv, p)
ekf.initialize_orientation([0.1, 0.2, 9.8])
lib = capi.load(); lib.hv_ekf_debug_result_words.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
rng = np.random.RandomState(0)
if os.environ.get("HV_EKF_CLUSTER_V1"):
    names = {0: "start", 1: "shift(aug)", 2: "H+P staged, residual", 3: "HP (phase A)", 4: "partial S stored", 5: "cluster.sync", 6: "S reduced+gathered",
             7: "elimination", 8: "chi2/decision", 9: "Z exchange + P/m update"}
else:   # ekf_cluster2.cuh
    names = {0: "start", 1: "P block + H staged, residual", 2: "HP", 3: "partial S", 4: "S reduced (DSMEM)", 5: "elimination", 6: "chi2/decision",
             7: "Z gathered (DSMEM)", 8: "P block downdate + m", 9: "stores",
             10: "Joseph: K", 11: "T1 columns", 12: "cluster.sync", 13: "special columns of G gathered (DSMEM)", 14: "Joseph product", 15: "cluster.sync",
             16: "mirrored entries fetched (DSMEM)"}
for n in (8, 20, 40, 84):
    l = min(160, 20 + 7 * max(1, n // 4))
    Hm = torch.from_numpy(np.asfortranarray(rng.normal(0, 0.1, (n, l))).ravel(order="F").copy()).cuda()
    f = torch.from_numpy(rng.normal(0, 0.5, n)).cuda(); y = f + 0.02 * torch.from_numpy(rng.normal(0, 1, n)).cuda()
    for mode in (0, 2):
        for rep in range(3):
            ekf.visual_device(Hm, n, l, f, y, 0.05, -1.0, mode)
        w = np.zeros(32); lib.hv_ekf_debug_result_words(ekf.h, w.ctypes.data)
        ts = w[8:18]
        keys = [k for k in range(10) if ts[k] > 0]
        line = f"n={n:3d} mode={mode}: total {(max(ts[keys]) - ts[0]) / 1e3:6.1f} us | "
        prev = ts[0]
        for k in keys[1:]:
            if ts[k] >= prev:
                line += f"{names[k]} {(ts[k] - prev) / 1e3:.1f} | "; prev = ts[k]
        print(line)
    ekf.symmetrize(); ekf.augment(-1)

for rep in range(3):
    ekf.symmetrize(); ekf.augment(-1)
ekf.flush()
w = np.zeros(32); lib.hv_ekf_debug_result_words(ekf.h, w.ctypes.data)
ts = w[8:25]
keys = sorted([k for k in range(17) if ts[k] > 0], key=lambda k: ts[k])
line = f"symmetrise+augment: total {(max(ts[keys]) - ts[0]) / 1e3:6.1f} us | "
prev = ts[0]
for k in keys[1:]:
    line += f"{names[k]} {(ts[k] - prev) / 1e3:.1f} | "; prev = ts[k]
print(line)

t = 1.0
for rep in range(3):
    for k in range(10):
        t += 0.005
        ekf.predict(t, [0.01, 0.02, 0.2], [0.1, 0.2, 9.8]); ekf.normalize_quaternions(True)
    ekf.flush()
w = np.zeros(32); lib.hv_ekf_debug_result_words(ekf.h, w.ctypes.data)
ts = w[8:14]
print("predict x10 (one launch): loads %.1f us | A_k + quaternion chain %.1f | Jacobians + W_k %.1f | covariance recursion %.1f | write-back + strips %.1f | total %.1f"
      % tuple([(ts[i + 1] - ts[i]) / 1e3 for i in range(5)] + [(ts[5] - ts[0]) / 1e3]))
