"""Generates tests/golden/tri_golden.npz: inputs and outputs of the REFERENCE's own per-track measurement model (triangulation
+ prepareVisualUpdate, src/odometry/triangulation.cpp compiled unmodified into oracle/_ref/libref_tri.so by
oracle/ref_build/build_tri.sh) on synthetic tracks (tests/tri_common.py). Run in the build container (needs /root/reference):
    python tests/golden/make_golden_tri.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import tri_common  # noqa: E402
from oracle import tri_oracle  # noqa: E402


def cases():
    """(seed, kwargs, corruption) covering OK / BEHIND / BAD_COND / NO_CONVERGENCE, mono and stereo, 2..10 poses."""
    out = []
    for seed in range(16):
        out.append((seed, dict(npose=2 + seed % 9, stereo=seed % 2 == 0, noise=[1e-3, 3e-3][seed % 2], depth=[2, 5, 15, 40][seed % 4]), "none"))
    # seeds found by scanning so that every TriangulatorStatus the path can return is present (BEHIND 2, BAD_COND 3, NO_CONVERGENCE 4)
    for seed, cor in ((102, "outlier"), (103, "outlier"), (100, "flip"), (101, "flip"), (101, "static"), (103, "static"), (100, "static"),
                      (107, "garbage"), (113, "garbage"), (114, "garbage"), (114, "flip"), (129, "flip"), (161, "outlier"), (200, "outlier")):
        out.append((seed, dict(npose=3 + seed % 5, stereo=seed % 2 == 0, depth=[3, 8, 60][seed % 3]), cor))
    return out


def build(seed, kw, corruption):
    t = tri_common.make_track(seed, **kw)
    tri_common.corrupt(t, corruption, seed)
    return t


def main():
    ref = tri_oracle.RefTri()
    blob = {}
    cs = cases()
    for i, (seed, kw, cor) in enumerate(cs):
        t = build(seed, kw, cor)
        for ets in (1, 0):
            o = ref.track_model(t["m"], t["trail"], t["stereo"], t["idx"], t["T1"], t["T2"], t["ip"], t["vel"], bool(ets))
            p = f"c{i}_t{ets}_"
            blob[p + "status"] = np.array([o["tri_status"], o["vu_status"]], np.int32)
            blob[p + "pf"] = o["pf"]; blob[p + "dpf"] = o["dpf"]; blob[p + "depth"] = np.array([o["depth"]])
            blob[p + "H"] = o["H"]; blob[p + "f"] = o["f"]
    blob["ncases"] = np.array([len(cs)], np.int32)
    np.savez_compressed(os.path.join(HERE, "tri_golden.npz"), **blob)
    st = [tuple(blob[f"c{i}_t1_status"]) for i in range(len(cs))]
    print("wrote tri_golden.npz:", len(cs), "cases; statuses", sorted(set(st)))


if __name__ == "__main__":
    main()
