"""Synthetic, physically consistent tracks for the per-track measurement model (triangulation + prepareVisualUpdate): a pose
trail in the EKF state layout, a stereo rig given by imuToCamera / secondImuToCamera, one 3-D point projected into every
observing camera pose (+ noise)."""
import numpy as np


def quat2rmat(q):
    """src/odometry/util.cpp:10-47 (row-major as written there)."""
    q0, q1, q2, q3 = q
    return np.array([[q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * q1 * q2 - 2 * q0 * q3, 2 * q1 * q3 + 2 * q0 * q2],
                     [2 * q1 * q2 + 2 * q0 * q3, q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * q2 * q3 - 2 * q0 * q1],
                     [2 * q1 * q3 - 2 * q0 * q2, 2 * q2 * q3 + 2 * q0 * q1, q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3]])


def project(m, idx, T1, T2, stereo, pf):
    """Normalised image coordinates of the world point pf in the cameras of the poses idx (camera 0 poses, then camera 1)."""
    ip = []
    for T in ([T1, T2] if stereo else [T1]):
        for i in idx:
            o = 0 if i == 0 else 20 + 7 * (i - 1)
            q = m[6:10] if i == 0 else m[o + 3:o + 7]
            R = T[:3, :3] @ quat2rmat(q)
            c = R @ (pf - (m[o:o + 3] - R.T @ T[:3, 3]))
            ip.append(c[:2] / c[2])
    return np.array(ip)


def corrupt_observations(t, kind, seed):
    """corrupt() restricted to the kinds that leave the state alone (tracks of one launch share it)."""
    assert kind in ("none", "outlier", "flip", "garbage")
    return corrupt(t, kind, seed)


def make_track(seed, trail=20, npose=6, stereo=True, noise=1e-3, depth=5.0, baseline=0.11):
    rng = np.random.RandomState(seed)
    N = 20 + 7 * trail
    m = np.zeros(N)
    # a smooth path: the device moves sideways ~8 cm per pose and turns slowly
    for k in range(trail + 1):
        pos = np.array([0.08 * k, 0.02 * np.sin(0.7 * k), 0.01 * k]) + rng.normal(0, 0.005, 3)
        ang = np.array([0.02 * k, -0.015 * k, 0.01 * np.sin(k)]) + rng.normal(0, 0.003, 3)
        q = np.array([1.0, *(0.5 * ang)]); q /= np.linalg.norm(q)
        o = 0 if k == 0 else 20 + 7 * (k - 1)
        m[o:o + 3] = pos
        if k == 0:
            m[6:10] = q
        else:
            m[o + 3:o + 7] = q
    m[3:6] = rng.normal(0, 0.1, 3); m[16:19] = 1.0
    # camera looks along the IMU z axis, slightly rotated; second camera displaced to the right (negative x translation,
    # src/tracker/util.cpp:95-104)
    a = rng.normal(0, 0.02, 3)
    qc = np.array([1.0, *(0.5 * a)]); qc /= np.linalg.norm(qc)
    T1 = np.eye(4); T1[:3, :3] = quat2rmat(qc); T1[:3, 3] = [0.01, -0.02, 0.005]
    T2 = T1.copy(); T2[:3, 3] = T1[:3, 3] + [-baseline, 0.0, 0.0]
    idx = np.sort(rng.choice(np.arange(1, trail + 1), npose - 1, replace=False))
    idx = np.concatenate([[0], idx]).astype(np.int32)
    # the point, in front of the first camera of pose idx[0]
    def cam(i, T):
        o = 0 if i == 0 else 20 + 7 * (i - 1)
        p = m[o:o + 3]
        q = m[6:10] if i == 0 else m[o + 3:o + 7]
        R = T[:3, :3] @ quat2rmat(q)
        return p - R.T @ T[:3, 3], R
    p0, R0 = cam(0, T1)
    pf = p0 + R0.T @ np.array([rng.uniform(-0.3, 0.3) * depth, rng.uniform(-0.2, 0.2) * depth, depth])
    ip = []
    for T in ([T1, T2] if stereo else [T1]):
        for i in idx:
            pc, R = cam(i, T)
            c = R @ (pf - pc)
            ip.append(c[:2] / c[2] + rng.normal(0, noise, 2))
    ip = np.array(ip)
    vel = rng.normal(0, 0.05, ip.shape)
    return dict(m=m, trail=trail, stereo=stereo, idx=idx, T1=T1, T2=T2, ip=ip, vel=vel, pf_true=pf)


def corrupt(t, kind, seed):
    """Spoils a track so that the reference's failure branches are taken: a gross outlier, mirrored observations (point behind
    the cameras), a static camera (no parallax -> BAD_COND), random observations."""
    rng = np.random.RandomState(10000 + seed)
    if kind == "outlier":
        t["ip"][rng.randint(len(t["ip"]))] += rng.normal(0, 0.5, 2)
    elif kind == "flip":
        t["ip"] = -t["ip"] + rng.normal(0, 0.05, t["ip"].shape)
    elif kind == "static":
        m = t["m"]
        for k in range(1, t["trail"] + 1):
            o = 20 + 7 * (k - 1)
            m[o:o + 3] = m[0:3] + rng.normal(0, 1e-7, 3)
            m[o + 3:o + 7] = m[6:10]
        t["ip"][:] = t["ip"][0] + rng.normal(0, 1e-6, t["ip"].shape)
    elif kind == "garbage":
        t["ip"] = rng.normal(0, 0.5, t["ip"].shape)
    else:
        assert kind == "none"
    return t
