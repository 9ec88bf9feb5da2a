// hybvio_b200/host/adapter_e2e_driver.cpp -- bench.py's `e2e_adapter`: one VIO frame after the other through the REFERENCE'S OWN
// virtual interfaces, in the order Session::process issues the calls (src/odometry/backend.cpp:716-867, 1158-1185):
//     10 x { EKF::predict, EKF::normalizeQuaternions(true) }                                   backend.cpp:734-735
//     ImagePyramid::Factory::compute (left, right), OpticalFlow::compute (temporal, with predicted corners; then stereo)
//                                                                                                image.cpp:87-106, tracker.cpp:395-438
//     per candidate track: EKF::visualTrackOutlierCheck(H, f, y, chiOutlierR, rmse)  -- ONE synchronous virtual call per track --
//                          and, for an INLIER while fewer than maxSuccessfulVisualUpdates: EKF::updateVisualTrack(H, f, y, visualR)
//     EKF::maintainPositiveSemiDefinite, EKF::updateVisualPoseAugmentation                      backend.cpp:1267, 805
//     EKF::position / orientation (output.setFromEKF)                                           backend.cpp:841
// No batching, no fused check+update, H handed over as a caller-owned Eigen matrix per call: what an unmodified backend.cpp costs.
//
// The file only uses the abstract interfaces and the three factory symbols, so the SAME source is linked twice:
//   hybvio_b200/libhv_adapter_e2e.so       + cuda_ekf.cpp, cuda_tracker_backends.cpp (-DHV_REPLACE_OPENCV_BACKENDS), libhybvio_b200.so
//   oracle/_ref/libref_adapter_e2e.so      + the reference's ekf.cpp, image_pyramid.cpp, optical_flow.cpp, vendored OpenCV (bench.py --impl reference)
// Bench harness, not part of the product library. Needs the reference headers: built where /root/reference exists (build()).
#include "ekf.hpp"
#include "image_pyramid.hpp"
#include "optical_flow.hpp"
#include "parameters.hpp"

#include <accelerated-arrays/cpu/image.hpp>
#include <chrono>
#include <cstdint>
#include <vector>

extern "C" {
typedef struct hv_adapter_frame {
    const uint8_t* left; const uint8_t* right;     // host gray images, right == NULL: mono
    const float* init_xy;                          // n x 2 predicted corners of the temporal flow
    const double* imu;                             // nimu x 7: t, gyro xyz, acc xyz
    int nimu;
    const double* tracks;                          // ntracks x { H (n x l column-major), f (n), y (n) }, packed
    const int* track_n; const int* track_l;
    int ntracks;
} hv_adapter_frame;

// pose_out: 7 doubles (position, orientation) after the last frame. counts: {outlier checks, inliers, updates}.
int hv_adapter_e2e_run(int width, int height, int maxTracks, int maxLevel, int trail, const float* points, int n, const hv_adapter_frame* frames,
                       int nframes, int warmup, int maxUpdates, double chiOutlierR, double visualR, double* pose_out, double* wall_ms, long long* counts)
{
    using clk = std::chrono::steady_clock;
    odometry::Parameters params;
    params.tracker.maxTracks = maxTracks; params.tracker.pyrLKMaxLevel = maxLevel;
    params.odometry.cameraTrailLength = trail;
    auto factory = tracker::ImagePyramid::Factory::buildOpenCv(params.tracker);                   // image.cpp:55
    auto flow = tracker::OpticalFlow::buildOpenCv(params.tracker);                                // image.cpp:56
    auto ekf = odometry::EKF::build(params);                                                      // backend.cpp:187
    std::vector<tracker::Feature::Point> prev(n), cur, right;
    for (int i = 0; i < n; i++) prev[i] = { points[2 * i], points[2 * i + 1] };
    std::vector<tracker::Feature::Status> status;
    std::shared_ptr<tracker::ImagePyramid> prevL;
    Eigen::MatrixXd H; Eigen::VectorXd f, y;
    long long nCheck = 0, nInlier = 0, nUpdate = 0;
    if (nframes > 0) { const double* u = frames[0].imu; ekf->initializeOrientation(Eigen::Vector3d(u[4], u[5], u[6])); }
    auto image = [&](const uint8_t* p) {
        return std::shared_ptr<accelerated::Image>(accelerated::cpu::Image::createReference(width, height, 1, accelerated::ImageTypeSpec::DataType::UFIXED8,
                                                                                           const_cast<uint8_t*>(p)));
    };
    clk::time_point t0 = clk::now();
    double predictorSink = 0.0;
    for (int k = 0; k < nframes; k++) {
        if (k == warmup) { t0 = clk::now(); nCheck = nInlier = nUpdate = 0; }
        const hv_adapter_frame& fr = frames[k];
        for (int s = 0; s < fr.nimu; s++) {
            const double* u = fr.imu + 7 * s;
            ekf->predict(u[0], Eigen::Vector3d(u[1], u[2], u[3]), Eigen::Vector3d(u[4], u[5], u[6]));
            ekf->normalizeQuaternions(true);
        }
        // the flow predictor reads the propagated pose and the newest pose of the trail before the optical flow (backend.cpp:547-600)
        {
            const Eigen::Vector3d pp = ekf->position(), hp = ekf->historyPosition(0);
            const Eigen::Vector4d po = ekf->orientation(), ho = ekf->historyOrientation(0);
            predictorSink += pp[0] + po[0] + hp[0] + ho[0];
        }
        auto pyrL = factory->compute(image(fr.left));
        std::shared_ptr<tracker::ImagePyramid> pyrR;
        if (fr.right) pyrR = factory->compute(image(fr.right));
        if (prevL) {
            cur.resize(n);
            for (int i = 0; i < n; i++) cur[i] = { fr.init_xy[2 * i], fr.init_xy[2 * i + 1] };
            flow->compute(*prevL, *pyrL, prev, cur, status, true);                                 // tracker.cpp:395-408
            if (pyrR) { right = cur; flow->compute(*pyrL, *pyrR, cur, right, status, false); }     // tracker.cpp:426-438
        }
        prevL = pyrL;
        int ok = 0;
        const double* p = fr.tracks;
        for (int c = 0; c < fr.ntracks; c++) {
            const int rows = fr.track_n[c], cols = fr.track_l[c];
            H = Eigen::Map<const Eigen::MatrixXd>(p, rows, cols); p += (size_t)rows * cols;       // prepareVisualUpdate fills caller-owned H, f (backend.cpp:1149)
            f = Eigen::Map<const Eigen::VectorXd>(p, rows); p += rows;
            y = Eigen::Map<const Eigen::VectorXd>(p, rows); p += rows;
            const auto st = ekf->visualTrackOutlierCheck(H, f, y, chiOutlierR, -1.0);              // backend.cpp:1158
            nCheck++;
            if (st == odometry::VuOutlierStatus::INLIER) {
                nInlier++;
                if (ok < maxUpdates) { ekf->updateVisualTrack(H, f, y, visualR); ok++; nUpdate++; }   // backend.cpp:1185
            }
        }
        ekf->maintainPositiveSemiDefinite();                                                       // backend.cpp:1267
        ekf->updateVisualPoseAugmentation(-1);                                                     // backend.cpp:805
        const Eigen::Vector3d pos = ekf->position();                                               // output: reads the state back
        const Eigen::Vector4d ori = ekf->orientation();
        if (pose_out) { for (int i = 0; i < 3; i++) pose_out[i] = pos[i]; for (int i = 0; i < 4; i++) pose_out[3 + i] = ori[i]; }
    }
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
    if (counts) { counts[0] = nCheck; counts[1] = nInlier; counts[2] = nUpdate; }
    if (pose_out && !(predictorSink == predictorSink)) pose_out[0] = predictorSink;      // (keeps the reads; NaN propagates)
    return 0;
}
}
