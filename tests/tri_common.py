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


def reference_visual_kat():
    """The known-answer test of the reference's own test suite (test/triangulation.cpp:56-167, TEST_CASE "visual", values generated
    by the authors' Matlab code): 10 poses (the current one + 9 trail slots) of a trail-20 filter, one mono track of 10 observations,
    imuToCamera = diag(1, -1, -1) (the default imuToCameraMatrix through tracker::util::automaticCameraParametersWhereUnset),
    feature velocities (0.1, 0.1), time shift estimated. Expected: TriangulatorStatus::OK and sum |pf - pf_e| < 1e-5."""
    poses = np.array([
        -1.115954259678003, -2.830379937574711, 0.360953864756080, 0.228275363465427, -0.064194730744503, -0.594104812214096, -0.772824444840030,
        -1.080393253042482, -2.763692958718615, 0.332645073392916, 0.196322489942363, -0.083909476935720, -0.628312037667580, -0.752388564841313,
        -1.053635192163148, -2.698599740902574, 0.304049959330811, 0.171347617609120, -0.090804163156838, -0.627022749727822, -0.749919482080305,
        -1.031838101194812, -2.623526076445418, 0.281408008477340, 0.155625729177218, -0.090380891656242, -0.639892913358913, -0.737146980096418,
        -1.009828260492951, -2.544268915819571, 0.273217018299048, 0.153209864083974, -0.090234014840705, -0.636707261073876, -0.737354342707954,
        -0.986215006493242, -2.468647298253558, 0.272275808868746, 0.157856184323099, -0.083435652262512, -0.606327170014471, -0.761376924834563,
        -0.961600705821358, -2.396757542411821, 0.267737813520921, 0.163130732364498, -0.079219306292358, -0.594278868691105, -0.765754228906657,
        -0.933757923541281, -2.325217937044675, 0.255438002606821, 0.172957779390792, -0.084991869290214, -0.593937386185525, -0.762521999377893,
        -0.898272888273739, -2.253889975199411, 0.239108878766994, 0.189256086747472, -0.090322497349436, -0.593833321653932, -0.758101862911017,
        -0.858474881652736, -2.184122374378553, 0.228789583088852, 0.204536006494471, -0.092660683000154, -0.580153035798419, -0.761692686677209])
    m = np.zeros(20 + 7 * 20)
    m[0:3] = poses[0:3]; m[6:10] = poses[3:7]
    for i in range(9):
        m[20 + 7 * i:20 + 7 * i + 7] = poses[7 * (i + 1):7 * (i + 2)]
    uv = np.array([[-0.182574266004879, -0.078574171780591], [-0.158898685463446, -0.007691759819452], [-0.131230597106084, -0.013212139610991],
                   [-0.110637420135181, 0.020800938142075], [-0.107508132406555, 0.002175057216783], [-0.108465120810051, -0.080045047328712],
                   [-0.111911566078740, -0.103534929832195], [-0.135452929226407, -0.099277664417604], [-0.165840298753357, -0.093731544303972],
                   [-0.188661852179662, -0.133908509900881]])
    T = np.diag([1.0, -1.0, -1.0, 1.0])
    return dict(m=m, trail=20, stereo=False, idx=np.arange(10, dtype=np.int32), T1=T, T2=T, ip=uv, vel=np.full((10, 2), 0.1),
                pf_expected=np.array([-2.32842, -8.02612, -0.619833]))
