"""ctypes wrappers for the pyramid + LK checkers. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  OracleLK  oracle/libhv_oracle.so   plain-C restatement (oracle/hv_oracle_lk.c), single-threaded
  RefLK     oracle/_ref/libref_lk.so the reference's own vendored OpenCV 4.3 pyramid + LK compiled unmodified
                                      (oracle/ref_build/Makefile.lk), OpenCV pthreads parallel_for_ on all cores
Both expose the same interface.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libhv_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_lk.so")
vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double


def have_ref():
    return os.path.exists(REF_SO)


class _Pyr:
    def __init__(self, owner, h, win):
        self.owner, self.h, self.win = owner, h, win
        self.levels = owner._levels(h)

    def level_size(self, level):
        w, h = ci(), ci()
        self.owner._level_size(self.h, level, ctypes.byref(w), ctypes.byref(h))
        return w.value, h.value

    def download(self, level, padded=True):
        w, h = self.level_size(level)
        W, H = w + 2 * self.win, h + 2 * self.win
        g = np.zeros((H, W), np.uint8)
        d = np.zeros((H, W, 2), np.int16)
        self.owner._get_padded(self.h, level, g.ctypes.data, d.ctypes.data)
        if padded:
            return g, d
        k = self.win
        return np.ascontiguousarray(g[k:-k, k:-k]), np.ascontiguousarray(d[k:-k, k:-k])

    def free(self):
        if self.h:
            self.owner._free(self.h)
            self.h = None


class OracleLK:
    """C restatement. accum_mode 0 = reference fp32 lane order (bit-exact with RefLK), 1 = exact-integer sums
    (the CUDA kernel's arithmetic)."""
    name = "port"

    def __init__(self):
        L = ctypes.CDLL(ORACLE_SO)
        L.orc_pyr_create.restype = vp
        L.orc_pyr_create.argtypes = [vp, ci, ci, ci, ci, ci]
        L.orc_pyr_levels.argtypes = [vp]
        L.orc_pyr_level_size.argtypes = [vp, ci, vp, vp]
        L.orc_pyr_get_level_padded.argtypes = [vp, ci, vp, vp]
        L.orc_pyr_free.argtypes = [vp]
        L.orc_lk.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, cd, ci, cd, ci]
        L.orc_lk_track_status.argtypes = [vp, vp, ci, ci, ci, vp]
        self.L = L
        self._levels, self._level_size, self._get_padded, self._free = L.orc_pyr_levels, L.orc_pyr_level_size, L.orc_pyr_get_level_padded, L.orc_pyr_free
        self.threads = 1

    def pyramid(self, gray, win=31, max_level=3):
        gray = np.ascontiguousarray(gray)
        h, w = gray.shape
        return _Pyr(self, self.L.orc_pyr_create(gray.ctypes.data, w, h, gray.strides[0], win, max_level), win)

    def lk(self, prev, nxt, prev_xy, next_xy=None, max_level=3, max_iter=20, eps=0.03, min_eig=1e-3, accum_mode=0):
        prev_xy = np.ascontiguousarray(prev_xy, np.float32)
        n = prev_xy.shape[0]
        use_initial = next_xy is not None
        out = np.ascontiguousarray(next_xy, np.float32).copy() if use_initial else np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.uint8)
        rc = self.L.orc_lk(prev.h, nxt.h, prev_xy.ctypes.data, out.ctypes.data, st.ctypes.data, n, max_level, max_iter, eps,
                           1 if use_initial else 0, min_eig, accum_mode)
        assert rc == 0
        return out, st, self.track_status(out, st, *nxt.level_size(0))

    def track_status(self, pts, st, width, height):
        ts = np.zeros(len(st), np.int32)
        pts = np.ascontiguousarray(pts, np.float32)
        self.L.orc_lk_track_status(pts.ctypes.data, st.ctypes.data, len(st), width, height, ts.ctypes.data)
        return ts


class RefLK:
    """The compiled reference (vendored OpenCV 4.3). Exists only where oracle/_ref was built."""
    name = "reference"

    def __init__(self):
        L = ctypes.CDLL(REF_SO)
        L.ref_pyr_build.restype = vp
        L.ref_pyr_build.argtypes = [vp, ci, ci, ci, ci, ci]
        L.ref_pyr_rebuild.argtypes = [vp, vp, ci, ci, ci, ci]
        L.ref_pyr_levels.argtypes = [vp]
        L.ref_pyr_level_size.argtypes = [vp, ci, vp, vp]
        L.ref_pyr_get_level_padded.argtypes = [vp, ci, vp, vp]
        L.ref_pyr_free.argtypes = [vp]
        L.ref_lk.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, cd, ci, cd]
        L.ref_set_num_threads.argtypes = [ci]
        self.L = L
        self._levels, self._level_size, self._get_padded, self._free = L.ref_pyr_levels, L.ref_pyr_level_size, L.ref_pyr_get_level_padded, L.ref_pyr_free
        self._orc = None

    @property
    def threads(self):
        return self.L.ref_get_num_threads()

    def set_threads(self, n):
        self.L.ref_set_num_threads(n)

    def pyramid(self, gray, win=31, max_level=3):
        gray = np.ascontiguousarray(gray)
        h, w = gray.shape
        p = _Pyr(self, self.L.ref_pyr_build(gray.ctypes.data, w, h, gray.strides[0], win, max_level), win)
        p.max_level = max_level
        return p

    def rebuild(self, pyr, gray):
        h, w = gray.shape
        self.L.ref_pyr_rebuild(pyr.h, gray.ctypes.data, w, h, gray.strides[0], pyr.max_level)

    def lk(self, prev, nxt, prev_xy, next_xy=None, max_level=3, max_iter=20, eps=0.03, min_eig=1e-3, accum_mode=0):
        prev_xy = np.ascontiguousarray(prev_xy, np.float32)
        n = prev_xy.shape[0]
        use_initial = next_xy is not None
        out = np.ascontiguousarray(next_xy, np.float32).copy() if use_initial else np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.uint8)
        self.L.ref_lk(prev.h, nxt.h, prev_xy.ctypes.data, out.ctypes.data, st.ctypes.data, n, prev.win, max_level, max_iter, eps,
                      1 if use_initial else 0, min_eig)
        w, h = nxt.level_size(0)
        # status mapping of src/tracker/optical_flow.cpp:52-58 (reference's own adapter needs accelerated-arrays; restated)
        ts = np.where(st == 0, 2, 0).astype(np.int32)
        oob = (out[:, 0] < 0) | (out[:, 0] >= np.float32(w)) | (out[:, 1] < 0) | (out[:, 1] >= np.float32(h))
        ts[oob] = 4
        return out, st, ts
