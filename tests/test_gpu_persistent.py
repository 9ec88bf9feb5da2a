"""GPU test of the persistent sequence of measurements (HV_EKF_PERSIST=1: consecutive check+update / update ops of hv_ekf_run_device
become ONE launch of ekf_update_multi_cluster2_kernel, the P blocks stay in shared memory in between) against the same
measurements applied one by one through the C oracle. Written after the last GPU session of round 1 (sorted last on purpose); the
kernel body has run on the host emulator only (tests/emu/emu_multi.cpp). The switch is read once per process, hence the child run."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def _child():
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from hybvio_b200 import capi
    from oracle import ekf_oracle
    hv = capi.Context(0)
    p = capi.EkfParams()
    capi.load().hv_ekf_default_params(ctypes.byref(p))
    a, b = capi.Ekf(hv, p), ekf_oracle.OracleEKF(p)
    rng = np.random.RandomState(9)
    N = a.N
    A = rng.normal(0, 1, (N, N))
    P0 = 0.05 * (A @ A.T) / N + np.diag(np.full(N, 0.5))
    m0 = 0.3 * (rng.rand(N) - 0.5)
    for k in range(21):
        o = 6 if k == 0 else 20 + 7 * (k - 1) + 3
        q = rng.normal(0, 0.3, 4) + [1, 0, 0, 0]
        m0[o:o + 4] = q / np.linalg.norm(q)
    a.upload(m=m0, P=P0); b.upload(m=m0, P=P0)
    launches0 = hv.launches
    seq = [(8, 34, 2, 0.02), (20, 55, 2, 0.02), (40, 90, 2, 40.0), (84, 160, 2, 0.02), (13, 41, 1, 0.02), (8, 34, 2, 0.02)]
    ops = (capi.EkfOp * len(seq))()
    keep, r = [], 0.05
    for i, (n, l, mode, ys) in enumerate(seq):
        H = np.asfortranarray(rng.normal(0, 0.1, (n, l)))
        f = rng.normal(0, 0.5, n)
        y = f + ys * rng.normal(0, 1, n)
        dH = torch.from_numpy(H.ravel(order="F").copy()).cuda(); df = torch.from_numpy(f).cuda(); dy = torch.from_numpy(y).cuda()
        keep.append((dH, df, dy))
        ops[i].kind, ops[i].n, ops[i].l, ops[i].mode, ops[i].r, ops[i].rmse_thr = capi.OP_VISUAL, n, l, mode, r, -1.0
        ops[i].H, ops[i].f, ops[i].y = dH.data_ptr(), df.data_ptr(), dy.data_ptr()
        st = 0
        if mode != 1:
            st, _ = b.visual_check(H, f, y, r)
        if mode == 1 or st == 0:
            b.visual_update(H, f, y, r)
    torch.cuda.synchronize()
    a.run_device(ops, len(seq))
    ma, Pa = a.download(); mb, Pb = b.download()
    used = hv.launches - launches0
    assert used == 1, f"{used} launches for the sequence: the persistent path was not taken"
    assert np.abs(ma - mb).max() < 1e-9 and np.abs(Pa - Pb).max() / np.abs(Pb).max() < 1e-9, (np.abs(ma - mb).max(), np.abs(Pa - Pb).max() / np.abs(Pb).max())
    a.close(); b.close(); hv.close()
    print("persistent sequence ok")


def test_persistent_sequence_matches_oracle():
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env={**os.environ, "HV_EKF_PERSIST": "1"}, capture_output=True, text=True,
                       timeout=600, cwd=os.path.dirname(HERE))
    assert r.returncode == 0 and "persistent sequence ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "child":
    _child()
