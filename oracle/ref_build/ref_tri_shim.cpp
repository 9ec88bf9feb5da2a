// TEST INFRASTRUCTURE: C shim around the reference's own per-track measurement model, compiled UNMODIFIED from
// /root/reference (src/odometry/triangulation.cpp, src/odometry/ekf.cpp, src/tracker/camera.cpp): the sequence
// Session::trackerVisualUpdate runs for one track (src/odometry/backend.cpp:1050-1160):
//   extractCameraPoseTrail -> Triangulator::triangulate (Gauss-Newton, with derivatives) -> stereo: sum the per-pose
//   derivatives of both cameras -> depth gate -> prepareVisualUpdate(truncated) -> H (2 n_obs x l), f.
// Oracle for the next hot-path row; checked against oracle/hv_oracle_tri.c in tests/test_oracle_tri.py.
#include "triangulation.hpp"
#include "ekf.hpp"
#include "parameters.hpp"
#include <cstring>
#include <memory>

extern "C" {

// m: state mean (N = 20 + 7 trail); poseTrailIndex: npose indices (0 = current pose, k = trail slot k - 1), as
// EkfStateIndex::createTrackIndex produces them; imuToCam / imuToCam2: 4x4 column-major; ip / vel: normalized image points and
// their velocities, [camera 0 poses..., camera 1 poses...] x 2; outputs: triStatus (TriangulatorStatus), pf[3],
// dpf[3 x (7 npose + 1)] column-major AFTER the stereo sum, vuStatus (PrepareVuStatus), H column-major rows x cols, f[rows].
int ref_track_model(const double* m, int trail, int useStereo, const int* poseTrailIndex, int npose, const double* imuToCam,
                    const double* imuToCam2, const double* ip, const double* vel, int estimateTimeShift, int* triStatus, double* pf,
                    double* dpf, double* depth, int* vuStatus, int* rows, int* cols, double* H, double* f)
{
    // the filter is only the state store behind extractCameraPoseTrail: build it once per trail length (timing runs call this
    // function per track; EKF::build allocates N x N matrices)
    static odometry::Parameters params;
    static std::unique_ptr<odometry::EKF> ekf;
    static int builtTrail = -1;
    params.odometry.estimateImuCameraTimeShift = estimateTimeShift != 0;
    params.tracker.useStereo = useStereo != 0;
    params.imuToCamera = Eigen::Map<const Eigen::Matrix4d>(imuToCam);
    params.secondImuToCamera = Eigen::Map<const Eigen::Matrix4d>(imuToCam2);
    if (builtTrail != trail) {
        params.odometry.cameraTrailLength = trail;
        ekf = odometry::EKF::build(params);
        builtTrail = trail;
    }
    const int N = ekf->getStateDim();
    ekf->setState(Eigen::Map<const Eigen::VectorXd>(m, N));
    std::vector<int> idx(poseTrailIndex, poseTrailIndex + npose);
    odometry::CameraPoseTrail tr;
    odometry::extractCameraPoseTrail(*ekf, idx, params, useStereo != 0, tr);
    const int nobs = npose * (useStereo ? 2 : 1);
    vecVector2d feats(nobs), vels(nobs);          // global alias of src/odometry/util.hpp
    for (int i = 0; i < nobs; i++) { feats[i] = Eigen::Vector2d(ip[2 * i], ip[2 * i + 1]); vels[i] = Eigen::Vector2d(vel[2 * i], vel[2 * i + 1]); }
    odometry::Triangulator triangulator(params.odometry);
    odometry::TriangulationArgsOut out;
    const odometry::TriangulationArgsIn args {
        .imageFeatures = feats, .featureVelocities = vels, .trail = tr, .stereo = useStereo != 0, .calculateDerivatives = true,
        .estimateImuCameraTimeShift = estimateTimeShift != 0,
    };
    odometry::TriangulatorStatus st = triangulator.triangulate(args, out);
    if (st != odometry::TriangulatorStatus::OK) { out.dpfdp.clear(); out.dpfdq.clear(); }          // backend.cpp:1100-1104
    if (useStereo && st == odometry::TriangulatorStatus::OK) {                                     // backend.cpp:1105-1116
        const size_t n = idx.size();
        for (size_t i = 0; i < n; ++i) { out.dpfdp[i] += out.dpfdp[i + n]; out.dpfdq[i] += out.dpfdq[i + n]; }
        out.dpfdp.resize(n); out.dpfdq.resize(n);
    }
    *depth = (out.pf - tr.at(0).p).norm();
    *triStatus = static_cast<int>(st);
    for (int i = 0; i < 3; i++) pf[i] = out.pf(i);
    std::memset(dpf, 0, sizeof(double) * 3 * (7 * npose + 1));
    if (st == odometry::TriangulatorStatus::OK) {
        for (int j = 0; j < npose; j++) {
            Eigen::Map<Eigen::Matrix<double, 3, 3>>(dpf + 3 * (7 * j)) = out.dpfdp[j];
            Eigen::Map<Eigen::Matrix<double, 3, 4>>(dpf + 3 * (7 * j + 3)) = out.dpfdq[j];
        }
        for (int i = 0; i < 3; i++) dpf[3 * 7 * npose + i] = out.dpfdt(i);
    }
    *vuStatus = -1; *rows = 0; *cols = 0;
    if (st != odometry::TriangulatorStatus::OK) return 0;
    Eigen::MatrixXd Hm; Eigen::VectorXd fv;
    const odometry::PrepareVisualUpdateArgsIn pargs {
        .triangulationOut = out, .featureVelocities = vels, .trail = tr, .poseTrailIndex = idx, .stateDim = N,
        .useStereo = useStereo != 0, .truncated = true, .mapPointOffset = -1, .estimateImuCameraTimeShift = estimateTimeShift != 0,
    };
    *vuStatus = static_cast<int>(odometry::prepareVisualUpdate(pargs, Hm, fv));
    *rows = (int)Hm.rows(); *cols = (int)Hm.cols();
    std::memcpy(H, Hm.data(), sizeof(double) * Hm.size());
    std::memcpy(f, fv.data(), sizeof(double) * fv.size());
    return 0;
}

}
