"""ctypes wrappers for the EKF checkers. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  OracleEKF  oracle/libhv_oracle.so    plain-C restatement (oracle/hv_oracle_ekf.c)
  RefEKF     oracle/_ref/libref_ekf.so the reference's own src/odometry/ekf.cpp + vendored Eigen, compiled unmodified
                                        (oracle/ref_build/build_ekf.sh)
Both expose the method set of hybvio_b200.capi.Ekf so parity tests can drive the three back ends with one script.
"""
import ctypes
import os

import numpy as np

from hybvio_b200.capi import EkfParams

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libhv_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_ekf.so")
vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double


def have_ref():
    return os.path.exists(REF_SO)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _CpuEkf:
    """Common driver; `pre` is the symbol prefix."""

    def __init__(self, lib, pre, params=None, handle=None):
        self.L, self.pre = lib, pre
        f = lambda name: getattr(lib, pre + name)
        self._f = f
        if not getattr(lib, "_hv_bound_" + pre, False):
            f("create").restype = vp; f("create").argtypes = [ctypes.POINTER(EkfParams)]
            f("clone").restype = vp; f("clone").argtypes = [vp]
            f("default_params").argtypes = [ctypes.POINTER(EkfParams)]
            for n in ("destroy", "state_dim", "pose_count", "was_stationary", "update_zupt_initialization", "unaugment", "symmetrize",
                      "condition_on_last_pose", "lock_biases"):
                f(n).argtypes = [vp]
            f("platform_time").restype = cd; f("platform_time").argtypes = [vp]
            f("history_time").restype = cd; f("history_time").argtypes = [vp, ci]
            f("set_first_sample_time").argtypes = [vp, cd]
            for n in ("upload", "download", "download_inertial", "set_inertial_state"):
                f(n).argtypes = [vp, vp, vp]
            for n in ("set_process_noise", "get_dydx", "initialize_orientation", "update_zrupt", "translate_to"):
                f(n).argtypes = [vp, vp]
            f("predict").argtypes = [vp, cd, vp, vp]
            f("update_zupt").argtypes = [vp, cd]
            f("update_zero_height").argtypes = [vp, cd]
            f("update_pseudo_velocity").argtypes = [vp, cd, cd]
            f("update_position").argtypes = [vp, vp, cd]
            f("update_orientation").argtypes = [vp, vp, cd]
            f("visual_update").argtypes = [vp, vp, ci, ci, vp, vp, cd]
            f("augment").argtypes = [vp, ci]
            f("normalize_quaternions").argtypes = [vp, ci]
            f("transform_to").argtypes = [vp, vp, vp, ci]
            f("insert_map_point").argtypes = [vp, ci, vp]
            setattr(lib, "_hv_bound_" + pre, True)
        if handle is not None:
            self.h = handle
        else:
            if params is None:
                params = self.default_params()
            self.params = params
            self.h = f("create")(ctypes.byref(params))
        self.N = f("state_dim")(self.h)

    def default_params(self):
        p = EkfParams()
        self._f("default_params")(ctypes.byref(p))
        return p

    def clone(self):
        c = type(self).__new__(type(self))
        _CpuEkf.__init__(c, self.L, self.pre, handle=self._f("clone")(self.h))
        c.params = self.params
        return c

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    # state access
    def upload(self, m=None, P=None):
        m = None if m is None else _d(m)
        P = None if P is None else np.asfortranarray(P, dtype=np.float64)
        self._f("upload")(self.h, None if m is None else m.ctypes.data, None if P is None else P.ctypes.data)

    def download(self):
        m = np.zeros(self.N); P = np.zeros((self.N, self.N), order="F")
        self._f("download")(self.h, m.ctypes.data, P.ctypes.data)
        return m, P

    def download_inertial(self):
        m = np.zeros(20); P = np.zeros((20, 20), order="F")
        self._f("download_inertial")(self.h, m.ctypes.data, P.ctypes.data)
        return m, P

    def set_inertial_state(self, m20, P20):
        m20 = _d(m20); P20 = np.asfortranarray(P20, dtype=np.float64)
        self._f("set_inertial_state")(self.h, m20.ctypes.data, P20.ctypes.data)

    def set_process_noise(self, Q):
        Q = np.asfortranarray(Q, dtype=np.float64)
        self._f("set_process_noise")(self.h, Q.ctypes.data)

    def get_dydx(self):
        d = np.zeros((20, 20), order="F")
        self._f("get_dydx")(self.h, d.ctypes.data)
        return d

    def pose_count(self): return self._f("pose_count")(self.h)
    def platform_time(self): return self._f("platform_time")(self.h)
    def history_time(self, i): return self._f("history_time")(self.h, i)
    def was_stationary(self): return bool(self._f("was_stationary")(self.h))
    def set_first_sample_time(self, t): self._f("set_first_sample_time")(self.h, t)

    # operations
    def initialize_orientation(self, acc): a = _d(acc); self._f("initialize_orientation")(self.h, a.ctypes.data)
    def predict(self, t, gyro, acc): g, a = _d(gyro), _d(acc); self._f("predict")(self.h, t, g.ctypes.data, a.ctypes.data)
    def update_zupt(self, r): self._f("update_zupt")(self.h, r)
    def update_zupt_initialization(self): self._f("update_zupt_initialization")(self.h)
    def update_zrupt(self, gyro): g = _d(gyro); self._f("update_zrupt")(self.h, g.ctypes.data)
    def update_pseudo_velocity(self, speed, r): self._f("update_pseudo_velocity")(self.h, speed, r)
    def update_position(self, pos, r): p = _d(pos); self._f("update_position")(self.h, p.ctypes.data, r)
    def update_zero_height(self, r): self._f("update_zero_height")(self.h, r)
    def update_orientation(self, q, r): q = _d(q); self._f("update_orientation")(self.h, q.ctypes.data, r)

    def visual_update(self, H, f, y, r):
        H = np.asfortranarray(H, dtype=np.float64); f, y = _d(f), _d(y)
        self._f("visual_update")(self.h, H.ctypes.data, H.shape[0], H.shape[1], f.ctypes.data, y.ctypes.data, r)

    def visual_check_update(self, H, f, y, r, rmse_thr=-1.0):
        st, chi2 = self.visual_check(H, f, y, r, rmse_thr)
        if st == 0:
            self.visual_update(H, f, y, r)
        return st, chi2, self.download()[0]

    def augment(self, drop=-1): self._f("augment")(self.h, drop)
    def unaugment(self): self._f("unaugment")(self.h)
    def symmetrize(self): self._f("symmetrize")(self.h)
    def normalize_quaternions(self, only_current=False): self._f("normalize_quaternions")(self.h, 1 if only_current else 0)
    def translate_to(self, pos): p = _d(pos); self._f("translate_to")(self.h, p.ctypes.data)
    def transform_to(self, pos, q, i=-1): p, q = _d(pos), _d(q); self._f("transform_to")(self.h, p.ctypes.data, q.ctypes.data, i)
    def insert_map_point(self, idx, pf): p = _d(pf); self._f("insert_map_point")(self.h, idx, p.ctypes.data)
    def condition_on_last_pose(self): self._f("condition_on_last_pose")(self.h)
    def lock_biases(self): self._f("lock_biases")(self.h)


class OracleEKF(_CpuEkf):
    name = "port"

    def __init__(self, params=None):
        L = ctypes.CDLL(ORACLE_SO)
        L.orc_ekf_visual_check.argtypes = [vp, vp, ci, ci, vp, vp, cd, cd, ctypes.POINTER(cd)]
        L.orc_chi2inv95.restype = cd; L.orc_chi2inv95.argtypes = [ci]
        super().__init__(L, "orc_ekf_", params)

    def visual_check(self, H, f, y, r, rmse_thr=-1.0):
        H = np.asfortranarray(H, dtype=np.float64); f, y = _d(f), _d(y)
        chi2 = cd(0.0)
        st = self.L.orc_ekf_visual_check(self.h, H.ctypes.data, H.shape[0], H.shape[1], f.ctypes.data, y.ctypes.data, r, rmse_thr, ctypes.byref(chi2))
        return st, chi2.value

    def chi2inv95(self, n):
        return self.L.orc_chi2inv95(n)


class RefEKF(_CpuEkf):
    """The compiled reference. visual_check returns (status, None): the reference does not expose the statistic."""
    name = "reference"

    def __init__(self, params=None):
        L = ctypes.CDLL(REF_SO)
        L.ref_ekf_visual_check.argtypes = [vp, vp, ci, ci, vp, vp, cd, cd]
        super().__init__(L, "ref_ekf_", params)

    def visual_check(self, H, f, y, r, rmse_thr=-1.0):
        H = np.asfortranarray(H, dtype=np.float64); f, y = _d(f), _d(y)
        return self.L.ref_ekf_visual_check(self.h, H.ctypes.data, H.shape[0], H.shape[1], f.ctypes.data, y.ctypes.data, r, rmse_thr), None
