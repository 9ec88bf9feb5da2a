"""Scripted EKF operation sequences shared by the golden generator and the parity tests. Every back end
(oracle port, compiled reference, CUDA) exposes the same methods, so one script drives all three.

Synthetic inputs follow SURVEY.md 8(d): H (n x l) entries N(0, 0.1^2) with l = 20 + 7*(n/4) capped at N, seed 3;
IMU 200 Hz, gyro ~ N(0, 0.05^2) around a 0.2 rad/s yaw, acc = R'g + N(0, 0.2^2), seed 11; per frame 10 x predict,
visual checks / updates, symmetrise, augment.
"""
import numpy as np

VISUAL_R = 0.05     # odometry.visualR (codegen/parameter_definitions.c:91)


def make_params(params_cls_instance, trail=20, map_size=0):
    p = params_cls_instance
    p.camera_trail_length = trail
    p.hybrid_map_size = map_size
    return p


def imu_sample(rng, k):
    gyro = np.array([0.0, 0.0, 0.2]) + rng.normal(0, 0.05, 3)
    acc = np.array([0.3 * np.sin(0.01 * k), 0.2 * np.cos(0.013 * k), 9.819]) + rng.normal(0, 0.2, 3)
    return gyro, acc


def visual_measurement(rng, n, N, scale_v):
    l = min(N, 20 + 7 * max(1, n // 4))
    H = rng.normal(0, 0.1, (n, l))
    f = rng.normal(0, 0.5, n)
    y = f + rng.normal(0, scale_v, n)
    return H, f, y


def run_frames(ekf, frames=6, n_list=(8, 20, 40), seed=3, fused=False, snapshots=None, checks=None,
               updates_per_frame=3, checks_per_frame=6, t0=0.0, frame0=0, norm_each=False):
    """The per-frame pattern of Session::process (src/odometry/backend.cpp:716-867): IMU predicts, optional
    un-augmentation of a non-keyframe, outlier checks + visual updates, symmetrise, augmentation."""
    rng = np.random.RandomState(seed)
    irng = np.random.RandomState(11)
    N = ekf.N
    t = t0
    if frame0 == 0:
        g, a = imu_sample(irng, 0)
        ekf.initialize_orientation(a)
    k = 0
    for fr in range(frame0, frame0 + frames):
        for _ in range(10):
            k += 1
            t += 0.005
            g, a = imu_sample(irng, k)
            ekf.predict(t, g, a)
            if norm_each:                        # the reference's sample loop (backend.cpp:734-735)
                ekf.normalize_quaternions(True)
        if not norm_each:
            ekf.normalize_quaternions(True)
        if fr % 4 == 3 and ekf.pose_count() > 1:
            ekf.unaugment()                      # non-keyframe (backend.cpp:793-796)
        done = 0
        for c in range(checks_per_frame):
            n = n_list[(fr + c) % len(n_list)]
            scale_v = 0.02 if (c % 3) else 40.0  # every third candidate is a gross outlier
            H, f, y = visual_measurement(rng, n, N, scale_v)
            if fused:
                st, chi2, _ = ekf.visual_check_update(H, f, y, VISUAL_R)
            else:
                st, chi2 = ekf.visual_check(H, f, y, VISUAL_R)
                if st == 0 and done < updates_per_frame:
                    ekf.visual_update(H, f, y, VISUAL_R)
            if checks is not None:
                checks.append((st, chi2))
            if st == 0:
                done += 1
            if done >= updates_per_frame:
                break
        ekf.symmetrize()
        ekf.augment(-1 if fr % 3 else (fr % max(1, min(ekf.pose_count(), ekf.params.camera_trail_length)) ))
        if snapshots is not None:
            snapshots.append(ekf.download())
    return t


def run_misc_ops(ekf, snapshots, t):
    """Everything else on the interface: the fixed-H updates, rigid transforms, structural edits."""
    irng = np.random.RandomState(5)
    snap = lambda: snapshots.append(ekf.download())
    g, a = imu_sample(irng, 1)
    for i in range(60):                    # 0.3 s so that the 4 Hz rate limits (ekf.cpp:574, 615) open
        t += 0.005
        ekf.predict(t, g, a)
    ekf.update_zupt_initialization(); snap()
    ekf.update_zupt(1e-2); snap()
    ekf.update_zupt(1e-2)                  # rate limited: no-op
    ekf.update_zrupt(g); snap()
    ekf.update_pseudo_velocity(0.7, 1.0); snap()
    ekf.update_position([0.1, -0.2, 0.05], 1e-3); snap()
    ekf.update_zero_height(1e-3); snap()
    q = np.array([0.9, 0.1, -0.2, 0.3]); q /= np.linalg.norm(q)
    ekf.update_orientation(q, 1e-2); snap()
    ekf.translate_to([1.0, 2.0, 3.0]); snap()
    q2 = np.array([0.7, -0.1, 0.2, 0.6]); q2 /= np.linalg.norm(q2)
    ekf.transform_to([0.5, -0.5, 0.25], q2, -1); snap()
    ekf.transform_to([0.0, 1.0, 0.0], [1.0, 0.0, 0.0, 0.0], 2); snap()
    m20, P20 = ekf.download_inertial()
    snapshots.append((m20, P20))
    ekf.lock_biases(); snap()
    if ekf.params.hybrid_map_size == 0:
        ekf.condition_on_last_pose(); snap()
    else:
        ekf.insert_map_point(1, [3.0, -2.0, 8.0]); snap()
    c = ekf.clone()
    c.predict(t + 0.005, g, a)
    snapshots.append(c.download())
    c.close()
    ekf.set_inertial_state(m20 * 1.01, P20 * 1.1); snap()
    assert ekf.pose_count() == 1
    return t


def rel_err(a, b):
    """max |a - b| / max |b| : the covariance tolerance is stated relative to the largest entry of P."""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max()))
