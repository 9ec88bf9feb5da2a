"""TEST INFRASTRUCTURE: ctypes wrappers of the corner-detector oracle (oracle/hv_oracle_gftt.c) and of the compiled reference
(oracle/_ref/libref_detect.so = the reference's feature_detector.cpp behind ref_build/ref_detect_shim.cpp)."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libhv_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_detect.so")
GAIN = 16.0


def have_ref():
    return os.path.exists(REF_SO)


class OracleGftt:
    def __init__(self):
        self.lib = ctypes.CDLL(ORACLE_SO)
        self.lib.orc_gftt_collect.restype = ctypes.c_int
        self.lib.orc_gftt_corners.restype = ctypes.c_int

    def response(self, img, block_size=3):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.zeros((h, w), np.float32)
        self.lib.orc_gftt_response(img.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(w), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(block_size),
                                   out.ctypes.data_as(ctypes.c_void_p))
        return out

    def collect(self, response, bs=32, min_response=1e-3):
        h, w = response.shape
        kp = np.zeros(((w // bs) * (h // bs), 3), np.float32)
        n = self.lib.orc_gftt_collect(np.ascontiguousarray(response).ctypes.data_as(ctypes.c_void_p), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(bs),
                                      ctypes.c_float(min_response), kp.ctypes.data_as(ctypes.c_void_p))
        assert n == len(kp)
        return kp

    def corners(self, kp, prev=None, mask_radius=0, max_tracks=200):
        prev = np.zeros((0, 2), np.float32) if prev is None else np.ascontiguousarray(prev, np.float32)
        kp = np.ascontiguousarray(kp, np.float32)
        out = np.zeros((2 * len(kp) + 1, 2), np.float32)
        n = self.lib.orc_gftt_corners(kp.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(kp)), prev.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(prev)),
                                      ctypes.c_int(mask_radius), ctypes.c_int(max_tracks), out.ctypes.data_as(ctypes.c_void_p))
        return out[:n].copy()

    def detect(self, img, prev=None, mask_radius=0, max_tracks=200, bs=32, min_response=1e-3):
        return self.corners(self.collect(self.response(img), bs, min_response), prev, mask_radius, max_tracks)


class RefGftt:
    def __init__(self):
        self.lib = ctypes.CDLL(REF_SO)
        self.lib.hv_ref_detect.restype = ctypes.c_int

    def response(self, img, block_size=3):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.zeros((h, w), np.float32)
        self.lib.hv_ref_corner_min_eigen_val(img.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(block_size), out.ctypes.data_as(ctypes.c_void_p))
        return out

    def detect(self, img, prev=None, mask_radius=0, max_tracks=200, min_distance=50.0, min_response=1e-3):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        prev = np.zeros((0, 2), np.float32) if prev is None else np.ascontiguousarray(prev, np.float32)
        cap = 4096
        out = np.zeros((cap, 2), np.float32)
        n = self.lib.hv_ref_detect(img.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(max_tracks), ctypes.c_double(min_distance),
                                   ctypes.c_float(min_response), prev.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(prev)), ctypes.c_int(mask_radius),
                                   out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(cap))
        return out[:min(n, cap)].copy()
