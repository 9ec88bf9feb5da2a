"""Launches each hot kernel a few times with BASELINE-config shapes (for ncu captures; see profiles/)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hybvio_b200 import capi, synth

hv = capi.Context(0)
W, H, N = 752, 480, 150
fr = synth.stereo_frames_torch(0, 3, W, H, device="cuda")
pyr = [hv.pyramid(W, H) for _ in range(4)]
pts = torch.from_numpy(synth.interior_points(N)).cuda()
fx, fy = synth.true_flow(0, 1)
init = (pts + torch.tensor([fx, fy], device="cuda", dtype=torch.float32)).contiguous()
nxt = torch.zeros_like(pts); nxt2 = torch.zeros_like(pts)
st = torch.zeros(N, dtype=torch.uint8, device="cuda"); ts = torch.zeros(N, dtype=torch.int32, device="cuda")
p = capi.EkfParams(); capi.load().hv_ekf_default_params(ctypes.byref(p))
ekf = capi.Ekf(hv, p)
ekf.initialize_orientation([0.1, 0.2, 9.8])
rng = np.random.RandomState(0)
t = 0.0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for it in range(reps):
    hv.build_pyramids(pyr[0:2], [fr[0, 0], fr[0, 1]], device=True)
    hv.build_pyramids(pyr[2:4], [fr[1, 0], fr[1, 1]], device=True)
    nxt.copy_(init)
    hv.lk_track_device(pyr[0], pyr[2], pts, nxt, st, ts, N, True)
    hv.lk_track_device(pyr[2], pyr[3], nxt, nxt2, st, ts, N, False)
    for s in range(10):                      # one IMU burst: 10 x (predict + normalizeQuaternions(true)) -> one launch
        t += 0.005
        ekf.predict(t, [0.01, 0.02, 0.2], [0.1, 0.2, 9.8]); ekf.normalize_quaternions(True)
    ekf.flush()
    for n in (8, 20, 40, 84):
        l = min(160, 20 + 7 * max(1, n // 4))
        Hm = torch.from_numpy(np.asfortranarray(rng.normal(0, 0.1, (n, l))).ravel(order="F").copy()).cuda()
        f = torch.from_numpy(rng.normal(0, 0.5, n)).cuda(); y = f + 0.02 * torch.from_numpy(rng.normal(0, 1, n)).cuda()
        ekf.visual_device(Hm, n, l, f, y, 0.05, -1.0, 0)
        ekf.visual_device(Hm, n, l, f, y, 0.05, -1.0, 2)
    ekf.symmetrize()
    ekf.augment(-1)
hv.sync()
print("ok", hv.launches)
