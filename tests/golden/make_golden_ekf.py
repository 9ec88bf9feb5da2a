"""Generates tests/golden/ekf_golden.npz from the COMPILED REFERENCE EKF (oracle/_ref/libref_ekf.so = the
reference's src/odometry/ekf.cpp + vendored Eigen, built by oracle/ref_build/build_ekf.sh) and embeds the
reference's own unit-test vectors so that they travel to boxes without /root/reference:
  * test/ekf.cpp:19-71   "chi-squared innovation test": 20x20 M, v, expected v' M^-1 v = 1.7626 +- 0.1
  * test/ekf.cpp:73-117  "der_predict": state, gyro, acc of the predict-Jacobian test
  * test/ekf.cpp:119-145 "tranformTo": fixtures test/data/P.csv (55x55), test/data/m.csv (55)
Run in the build container:  python tests/golden/make_golden_ekf.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ekf_script
from oracle.ekf_oracle import RefEKF

REF = "/root/reference"


def parse_reference_tests():
    src = open(os.path.join(REF, "test", "ekf.cpp")).read()
    nums = lambda s: np.array([float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", s)])
    blk = src[src.index("chi-squared innovation test"):src.index("der_predict")]
    parts = re.findall(r"<<(.*?);", blk, flags=re.S)
    left, right, v = nums(parts[0]).reshape(20, 10), nums(parts[1]).reshape(20, 10), nums(parts[2])
    M = np.hstack([left, right]) * 1e3
    assert v.shape == (20,)
    blk2 = src[src.index("der_predict"):src.index("tranformTo")]
    poses = nums(re.search(r"poses; poses <<(.*?);", blk2, flags=re.S).group(1))
    gyro = nums(re.search(r"gyro; gyro <<(.*?);", blk2).group(1))
    acc = nums(re.search(r"acc; acc <<(.*?);", blk2).group(1))
    assert poses.shape == (70,)
    P0 = np.loadtxt(os.path.join(REF, "test", "data", "P.csv"), delimiter=",")
    m0 = np.loadtxt(os.path.join(REF, "test", "data", "m.csv"), delimiter=",").ravel()
    assert P0.shape == (55, 55) and m0.shape == (55,)
    return dict(reftest_M=M, reftest_v=v, reftest_poses=poses, reftest_gyro=gyro, reftest_acc=acc, reftest_P0=P0, reftest_m0=m0)


def main():
    out = parse_reference_tests()
    for name, trail, frames, nlist in (("n62", 6, 8, (8, 20)), ("n160", 20, 24, (8, 20, 40, 84))):
        e = RefEKF()
        p = e.default_params(); p.camera_trail_length = trail
        e.close()
        e = RefEKF(p)
        snaps, checks = [], []
        t = ekf_script.run_frames(e, frames=frames, n_list=nlist, snapshots=snaps, checks=checks)
        out[f"{name}_check_status"] = np.array([c[0] for c in checks], np.int32)
        keep = range(len(snaps)) if name == "n62" else (len(snaps) - 1,)
        for i in keep:
            out[f"{name}_m_{i}"] = snaps[i][0]; out[f"{name}_P_{i}"] = snaps[i][1]
        out[f"{name}_frames"] = np.int32(frames)
        misc = []
        ekf_script.run_misc_ops(e, misc, t)
        for i, (m, P) in enumerate(misc):
            out[f"{name}_misc_m_{i}"] = m
            if name == "n62" or i == len(misc) - 3:
                out[f"{name}_misc_P_{i}"] = P
        out[f"{name}_misc_count"] = np.int32(len(misc))
        print(name, "N =", e.N, "checks:", np.bincount(out[f"{name}_check_status"], minlength=4), "poses", e.pose_count())
        e.close()
    path = os.path.join(ROOT, "tests", "golden", "ekf_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
