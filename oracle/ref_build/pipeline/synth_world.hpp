// TEST INFRASTRUCTURE: a physically consistent synthetic stereo + IMU stream (SURVEY.md 8(d) "pipeline-level stream").
//
// One 3-D scene (the inside of a textured box, walls 3-6 m away) is rendered through pinhole intrinsics into the left and
// the right camera from the SAME analytic IMU trajectory the gyroscope / accelerometer samples are derived from, with the
// SAME imuToCamera / secondImuToCamera transforms the odometry::Parameters declare -- otherwise the stereo epipolar gate
// (src/tracker/tracker.cpp:348-376) or the EKF chi^2 gate rejects everything and the run never leaves INIT.
// Conventions follow the reference: quat2rmat(q) = world-to-IMU rotation (src/odometry/util.cpp:10-26), gyro in the body
// frame with q <- exp(-dt/2 Omega(w)) q (src/odometry/ekf.cpp:414-425), v' = R^T (a) + g with g = (0,0,-9.81)
// (ekf.cpp:172, 436), x_cam = imuToCamera * x_imu (src/odometry/triangulation.cpp:65-103).
#ifndef HV_PIPELINE_SYNTH_WORLD_HPP_
#define HV_PIPELINE_SYNTH_WORLD_HPP_
#include <Eigen/Dense>
#include <cmath>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>

namespace synth {

inline uint32_t hash2(int x, int y, uint32_t seed) {
    uint32_t h = (uint32_t)x * 0x8da6b343u ^ (uint32_t)y * 0xd8163841u ^ seed * 0xcb1ab31fu;
    h ^= h >> 13; h *= 0x85ebca6bu; h ^= h >> 16; h *= 0xc2b2ae35u; h ^= h >> 15;
    return h;
}
inline double lattice(int x, int y, uint32_t seed) { return (hash2(x, y, seed) & 0xffff) / 32767.5 - 1.0; }   // [-1, 1]
inline double valueNoise(double u, double v, uint32_t seed) {
    const double fu = std::floor(u), fv = std::floor(v);
    const int iu = (int)fu, iv = (int)fv;
    double a = u - fu, b = v - fv;
    a = a * a * (3 - 2 * a); b = b * b * (3 - 2 * b);
    return (1 - b) * ((1 - a) * lattice(iu, iv, seed) + a * lattice(iu + 1, iv, seed)) +
           b * ((1 - a) * lattice(iu, iv + 1, seed) + a * lattice(iu + 1, iv + 1, seed));
}
// texture of a plane in metres: three octaves of value noise + random blocks (sharp corners for the GFTT detector)
inline double texture(double u, double v, uint32_t seed) {
    double t = 128.0 + 38.0 * valueNoise(u / 0.06, v / 0.06, seed) + 30.0 * valueNoise(u / 0.19, v / 0.19, seed + 1) +
               24.0 * valueNoise(u / 0.55, v / 0.55, seed + 2);
    t += 34.0 * lattice((int)std::floor(u / 0.11), (int)std::floor(v / 0.11), seed + 3);
    return t;
}

struct Pose {                  // IMU pose in the world
    Eigen::Vector3d p;         // position
    Eigen::Matrix3d Rwb;       // body-to-world rotation (world-to-IMU = Rwb^T)
};

struct Trajectory {
    double still = 1.0;        // seconds at rest before the motion fades in
    double scale = 1.0;        // amplitude scale
    static double fade(double t, double t0) {       // C2 smooth step over 2 s
        double s = (t - t0) / 2.0;
        if (s <= 0) return 0; if (s >= 1) return 1;
        return s * s * s * (10 - 15 * s + 6 * s * s);
    }
    Pose at(double t) const {
        const double k = fade(t, still) * scale;
        Pose P;
        P.p = Eigen::Vector3d(0.55 * k * std::sin(0.45 * t), 0.45 * k * std::sin(0.31 * t + 0.7), 0.25 * k * std::sin(0.53 * t + 0.3));
        const double yaw = 0.35 * k * std::sin(0.23 * t), pitch = 0.12 * k * std::sin(0.37 * t + 1.0), roll = 0.10 * k * std::sin(0.29 * t + 2.0);
        P.Rwb = (Eigen::AngleAxisd(yaw, Eigen::Vector3d::UnitZ()) * Eigen::AngleAxisd(pitch, Eigen::Vector3d::UnitY()) *
                 Eigen::AngleAxisd(roll, Eigen::Vector3d::UnitX())).toRotationMatrix();
        return P;
    }
    // body-frame angular rate and specific force by central differences of the analytic pose (h = 1e-4 s: error ~1e-8)
    void imu(double t, Eigen::Vector3d& gyro, Eigen::Vector3d& acc, double gravity = 9.81) const {
        const double h = 1e-4;
        const Pose a = at(t - h), b = at(t), c = at(t + h);
        const Eigen::Matrix3d W = b.Rwb.transpose() * (c.Rwb - a.Rwb) / (2 * h);      // [w]x = Rwb^T dRwb/dt
        gyro = Eigen::Vector3d(W(2, 1) - W(1, 2), W(0, 2) - W(2, 0), W(1, 0) - W(0, 1)) * 0.5;
        const Eigen::Vector3d pdd = (c.p - 2 * b.p + a.p) / (h * h);
        acc = b.Rwb.transpose() * (pdd + Eigen::Vector3d(0, 0, gravity));
    }
};

struct Camera {                // pinhole; x_cam = Rc * x_imu + tc
    double fx, fy, cx, cy;
    Eigen::Matrix3d Rc;
    Eigen::Vector3d tc;
};

// Axis-aligned box room centred on the origin, half sizes (hx, hy, hz); the IMU looks along +x at rest.
struct Room {
    double hx = 4.5, hy = 3.5, hz = 2.6;
    uint32_t seed = 42;
    // first wall hit by the ray o + s d (o inside the box): gray value
    double shade(const Eigen::Vector3d& o, const Eigen::Vector3d& d) const {
        double best = 1e30; int face = -1;
        const double half[3] = {hx, hy, hz};
        for (int ax = 0; ax < 3; ax++) {
            if (d[ax] == 0) continue;
            const double wall = d[ax] > 0 ? half[ax] : -half[ax];
            const double s = (wall - o[ax]) / d[ax];
            if (s > 0 && s < best) { best = s; face = 2 * ax + (d[ax] > 0 ? 0 : 1); }
        }
        const Eigen::Vector3d x = o + best * d;
        const int ax = face / 2;
        const double u = x[(ax + 1) % 3], v = x[(ax + 2) % 3];
        return texture(u + 10.0, v + 10.0, seed + 17u * (uint32_t)face);
    }
};

// Renders the gray image of one camera at IMU pose P (2x2 supersampling; rows split over threads; deterministic).
inline void render(const Room& room, const Camera& cam, const Pose& P, int w, int h, std::vector<uint8_t>& out, int threads = 8) {
    out.resize((size_t)w * h);
    const Eigen::Matrix3d Rcw = cam.Rc * P.Rwb.transpose();                 // world-to-camera
    const Eigen::Vector3d centre = P.p - Rcw.transpose() * cam.tc;          // triangulation.cpp:88
    const Eigen::Matrix3d Rwc = Rcw.transpose();
    auto rows = [&](int y0, int y1) {
        for (int y = y0; y < y1; y++)
            for (int x = 0; x < w; x++) {
                double acc = 0;
                for (int sy = 0; sy < 2; sy++)
                    for (int sx = 0; sx < 2; sx++) {
                        const double px = x + (sx - 0.5) * 0.5, py = y + (sy - 0.5) * 0.5;      // pixel centre at integer coordinates
                        const Eigen::Vector3d dc((px - cam.cx) / cam.fx, (py - cam.cy) / cam.fy, 1.0);
                        acc += room.shade(centre, Rwc * dc);
                    }
                const double g = acc * 0.25;
                out[(size_t)y * w + x] = (uint8_t)(g < 0 ? 0 : g > 255 ? 255 : std::lround(g));
            }
    };
    std::vector<std::thread> pool;
    const int per = (h + threads - 1) / threads;
    for (int t = 0; t < threads; t++) pool.emplace_back(rows, std::min(h, t * per), std::min(h, (t + 1) * per));
    for (auto& t : pool) t.join();
}

} // namespace synth
#endif
