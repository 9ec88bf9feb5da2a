"""TEST INFRASTRUCTURE: ctypes wrappers of the per-track measurement model (triangulation + prepareVisualUpdate,
src/odometry/triangulation.cpp; the next hot-path row):
  RefTri     the reference's own code compiled unmodified (oracle/_ref/libref_tri.so, oracle/ref_build/build_tri.sh)
  OracleTri  the plain-C restatement (oracle/hv_oracle_tri.c in oracle/libhv_oracle.so)
Both expose track_model(m, trail, use_stereo, pose_index, imu_to_cam, imu_to_cam2, ip, vel, estimate_time_shift)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libref_tri.so")
ORACLE_SO = os.path.join(_HERE, "libhv_oracle.so")


def have_ref():
    return os.path.exists(REF_SO)


class _Tri:
    def __init__(self, path, symbol):
        self.lib = ctypes.CDLL(path)
        self.fn = getattr(self.lib, symbol)
        self.fn.restype = ctypes.c_int

    def track_model(self, m, trail, use_stereo, pose_index, imu_to_cam, imu_to_cam2, ip, vel, estimate_time_shift=True):
        m = np.ascontiguousarray(m, np.float64)
        idx = np.ascontiguousarray(pose_index, np.int32)
        npose = len(idx)
        nobs = npose * (2 if use_stereo else 1)
        ip = np.ascontiguousarray(ip, np.float64).reshape(nobs, 2)
        vel = np.ascontiguousarray(vel, np.float64).reshape(nobs, 2)
        a = np.asfortranarray(imu_to_cam, np.float64)
        b = np.asfortranarray(imu_to_cam2, np.float64)
        N = len(m)
        tri, vu, rows, cols = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        depth = ctypes.c_double()
        pf = np.zeros(3)
        dpf = np.zeros(3 * (7 * npose + 1))
        H = np.zeros(2 * nobs * N)
        f = np.zeros(2 * nobs)
        p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
        rc = self.fn(p(m), ctypes.c_int(trail), ctypes.c_int(1 if use_stereo else 0), p(idx), ctypes.c_int(npose), p(a), p(b), p(ip), p(vel),
                     ctypes.c_int(1 if estimate_time_shift else 0), ctypes.byref(tri), p(pf), p(dpf), ctypes.byref(depth), ctypes.byref(vu),
                     ctypes.byref(rows), ctypes.byref(cols), p(H), p(f))
        assert rc == 0
        r, c = rows.value, cols.value
        return {"tri_status": tri.value, "pf": pf, "dpf": dpf.reshape((3, 7 * npose + 1), order="F"), "depth": depth.value,
                "vu_status": vu.value, "H": H[:r * c].reshape((r, c), order="F").copy(), "f": f[:r].copy()}


class RefTri(_Tri):
    def __init__(self):
        super().__init__(REF_SO, "ref_track_model")


class OracleTri(_Tri):
    def __init__(self):
        super().__init__(ORACLE_SO, "orc_track_model")
