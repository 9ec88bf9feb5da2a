// TEST INFRASTRUCTURE: run_pipeline -- the UNMODIFIED reference core (tracker.cpp, image.cpp, backend.cpp, control.cpp,
// triangulation.cpp, RANSAC, detector ...; SLAM off) driven at odometry::Control exactly as the API drives it
// (src/api/api.cpp:415-430 gyro -> processSyncedSamples, :570-628 stereo frames, :879-897 image factories) on a physically
// consistent synthetic stereo + IMU stream, with the back ends chosen per pipeline (backends.hpp):
//   --mode ref       stock reference back ends only (runs without a GPU: stream sanity, CPU frames/s)
//   --mode cuda      CUDA back ends only
//   --mode lockstep  ONE pipeline whose back ends are lock-step pairs: every pyramid / LK / EKF call goes to the reference
//                    class and to the CUDA class with identical inputs and is compared; a second TrackerImplementation runs
//                    on CUDA-flavoured images beside the reference-driven one (Tracker::Output IDs / statuses / points)
//   --mode free      TWO complete pipelines (reference, CUDA) run independently on the same stream: first-divergence frame of
//                    Tracker::Output, pose difference over time
// Writes one JSON document (--out) that tests/test_gpu_pipeline.py asserts on.
#include "backends.hpp"
#include "synth_world.hpp"

#include "control.hpp"
#include "ekf.hpp"
#include "parameters.hpp"
#include "../tracker/camera.hpp"
#include "../tracker/image.hpp"
#include "../tracker/util.hpp"
#include "../odometry/tagged_frame.hpp"

#include <accelerated-arrays/cpu/image.hpp>
#include <accelerated-arrays/cpu/operations.hpp>
#include <accelerated-arrays/future.hpp>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

using harness::Flavour;

namespace {
struct Config {
    int id = 2, w = 752, h = 480, maxTracks = 150, maxLevel = 3, trail = 20;
    bool stereo = true;
    double f = 458.0, baseline = 0.11;
    const char* name = "config2: stereo 752x480, 150 features, 4-level pyramid, N=160";
};
Config makeConfig(int id) {
    Config c;
    if (id == 4) { c.id = 4; c.w = 512; c.h = 512; c.maxTracks = 200; c.trail = 6; c.f = 260.0; c.baseline = 0.10;
                   c.name = "config4: stereo 512x512, 200 features, 4-level pyramid, N=62"; }
    else if (id == 1) { c.id = 1; c.stereo = false; c.maxTracks = 100; c.maxLevel = 2;
                        c.name = "config1: mono 752x480, 100 features, 3-level pyramid, N=160"; }
    return c;
}

void fillParameters(odometry::Parameters& p, const Config& c) {
    auto& t = p.tracker; auto& o = p.odometry;
    t.focalLength = c.f; t.principalPointX = 0.5 * (c.w - 1); t.principalPointY = 0.5 * (c.h - 1);
    t.maxTracks = c.maxTracks; t.pyrLKMaxLevel = c.maxLevel;
    t.useStereo = c.stereo;
    // featureDetector stays at its default "GPU-GFTT": with CPU images that is CpuCornerResponse + CollectMax::cpuImplementation
    // (src/tracker/feature_detector.cpp:666-669), the path the CUDA detector replaces
    t.targetFps = 20;
    o.cameraTrailLength = c.trail;
    if (c.trail < 10) o.cameraTrailHanoiLength = 2;
    // IMU: x forward, y left, z up.  Camera: z forward, x right, y down.  x_cam = Rc x_imu + t
    Eigen::Matrix4d T1 = Eigen::Matrix4d::Identity(), T2;
    T1.topLeftCorner<3, 3>() << 0, -1, 0,   0, 0, -1,   1, 0, 0;
    T2 = T1;
    // right camera displaced by +baseline along the camera x axis: t2 = -Rc c2 = (-baseline, 0, 0)  (src/tracker/util.cpp:95-104)
    T2(0, 3) = -c.baseline;
    o.imuToCameraMatrix.assign(T1.data(), T1.data() + 16);              // column-major (src/odometry/util.hpp:92-109)
    o.secondImuToCameraMatrix.assign(T2.data(), T2.data() + 16);
    p.slam.useSlam = false;
    tracker::util::automaticCameraParametersWhereUnset(p);
}

struct Pipeline {
    Flavour flavour;
    const char* name;
    std::unique_ptr<accelerated::Queue> queue;
    std::unique_ptr<accelerated::Image::Factory> imageFactory;
    std::unique_ptr<accelerated::operations::StandardFactory> opsFactory;
    std::unique_ptr<tracker::Image::Factory> trackerImages;
    std::unique_ptr<odometry::Control> control;
    // extra CUDA-flavoured image factory for the shadow tracker of the lock-step mode
    std::unique_ptr<tracker::Image::Factory> shadowImages;
    int framesOut = 0;
    double seconds = 0;
    std::vector<double> frameMs;
    struct Out { double t; Eigen::Vector3d p; Eigen::Vector4d q; int status; Eigen::VectorXd m; double Pmax; Eigen::MatrixXd P; };
    std::vector<Out> outs;

    Pipeline(Flavour f, const char* n, const odometry::Parameters& params) : flavour(f), name(n) {
        harness::setFlavour(f);
        queue = accelerated::Processor::createQueue();
        imageFactory = accelerated::cpu::Image::createFactory();
        opsFactory = accelerated::cpu::operations::createFactory(*queue);
        trackerImages = tracker::Image::buildFactory(*queue, *imageFactory, *opsFactory, params);      // api.cpp:893
        if (f == Flavour::DUAL) shadowImages = tracker::Image::buildFactory(*queue, *imageFactory, *opsFactory, params);
        control = odometry::Control::build(params);                                                    // api.cpp:80-81
    }
};

std::string jsonVec(const std::vector<double>& v) {
    std::ostringstream s; s.precision(6); s << "[";
    for (size_t i = 0; i < v.size(); i++) { if (i) s << ","; s << v[i]; }
    s << "]"; return s.str();
}
} // namespace

int main(int argc, char** argv) {
    std::string mode = "ref", outPath;
    int configId = 2, frames = 120, threads = 0, keepP = 0, rectify = 0;
    double still = 1.0, motion = 1.0;
    for (int i = 1; i < argc; i++) {
        auto arg = [&](const char* k) { return !std::strcmp(argv[i], k) && i + 1 < argc; };
        if (arg("--mode")) mode = argv[++i];
        else if (arg("--config")) configId = std::atoi(argv[++i]);
        else if (arg("--frames")) frames = std::atoi(argv[++i]);
        else if (arg("--out")) outPath = argv[++i];
        else if (arg("--threads")) threads = std::atoi(argv[++i]);
        else if (arg("--still")) still = std::atof(argv[++i]);
        else if (arg("--motion")) motion = std::atof(argv[++i]);
        else if (arg("--keep-cov")) keepP = std::atoi(argv[++i]);
        else if (arg("--cuda-detector")) harness::setUseCudaDetector(std::atoi(argv[++i]) != 0);
        else if (arg("--rectify")) rectify = std::atoi(argv[++i]);
        else { std::fprintf(stderr, "usage: run_pipeline --mode ref|cuda|lockstep|free [--config 2|4|1] [--frames N] [--out file.json] [--threads T]\n"); return 2; }
    }
    const Config cfg = makeConfig(configId);
    odometry::Parameters params;
    fillParameters(params, cfg);
    // --rectify 1: stereo rectification on (tracker.useRectification): every frame goes through StereoRectifier + Undistorter
    // (src/tracker/image.cpp:316-332), i.e. through the frame-ingest row N4
    if (rectify) { params.tracker.useRectification = true; }
    if (threads > 0) harness::setRefThreads(threads);

    synth::Room room;
    synth::Trajectory traj; traj.still = still; traj.scale = motion;
    synth::Camera camL { cfg.f, cfg.f, 0.5 * (cfg.w - 1), 0.5 * (cfg.h - 1), params.imuToCamera.topLeftCorner<3, 3>(), params.imuToCamera.block<3, 1>(0, 3) };
    synth::Camera camR = camL;
    camR.Rc = params.secondImuToCamera.topLeftCorner<3, 3>(); camR.tc = params.secondImuToCamera.block<3, 1>(0, 3);

    std::vector<std::unique_ptr<Pipeline>> pipes;
    if (mode == "ref") pipes.emplace_back(new Pipeline(Flavour::REF, "reference", params));
    else if (mode == "cuda") pipes.emplace_back(new Pipeline(Flavour::CUDA, "cuda", params));
    else if (mode == "lockstep") pipes.emplace_back(new Pipeline(Flavour::DUAL, "lockstep", params));
    else if (mode == "free") { pipes.emplace_back(new Pipeline(Flavour::REF, "reference", params)); pipes.emplace_back(new Pipeline(Flavour::CUDA, "cuda", params)); }
    else { std::fprintf(stderr, "unknown mode %s\n", mode.c_str()); return 2; }

    const double imuDt = 0.005, frameDt = 0.05;
    const int imuPerFrame = 10;
    std::mt19937 rng(11);
    std::normal_distribution<double> gyroNoise(0.0, 0.002), accNoise(0.0, 0.02);
    std::vector<uint8_t> imgL, imgR;
    const api::CameraParameters intrinsic = [&] { api::CameraParameters k; k.focalLengthX = cfg.f; k.focalLengthY = cfg.f;
        k.principalPointX = camL.cx; k.principalPointY = camL.cy; return k; }();
    const auto kind = tracker::Camera::Kind::PINHOLE;
    using clk = std::chrono::steady_clock;

    long imuIndex = 0;
    for (int k = 0; k < frames; k++) {
        // ---- IMU samples of this frame interval (200 Hz), shared by all pipelines
        struct Imu { double t; Eigen::Vector3d g, a; };
        std::vector<Imu> imu;
        for (int j = 0; j < imuPerFrame; j++, imuIndex++) {
            Imu s; s.t = imuIndex * imuDt;
            traj.imu(s.t, s.g, s.a, params.odometry.gravity);
            for (int d = 0; d < 3; d++) { s.g[d] += gyroNoise(rng); s.a[d] += accNoise(rng); }
            imu.push_back(s);
        }
        const double tFrame = k * frameDt;
        const synth::Pose pose = traj.at(tFrame);
        synth::render(room, camL, pose, cfg.w, cfg.h, imgL);
        if (cfg.stereo) synth::render(room, camR, pose, cfg.w, cfg.h, imgR);

        for (auto& pp : pipes) {
            Pipeline& P = *pp;
            harness::setFlavour(P.flavour);
            const auto t0 = clk::now();
            // frame first (time stamp = first IMU sample of the interval), then the IMU samples up to the next frame
            auto makeImages = [&](tracker::Image::Factory& fac) {
                auto accL = accelerated::cpu::Image::createReference(cfg.w, cfg.h, 1, accelerated::ImageTypeSpec::DataType::UFIXED8, imgL.data());
                std::pair<std::unique_ptr<tracker::Image>, std::unique_ptr<tracker::Image>> r;
                std::shared_ptr<const tracker::Camera> c0 = tracker::buildCamera(intrinsic, kind, params.tracker, cfg.w, cfg.h, params.tracker.distortionCoeffs);
                if (cfg.stereo) {
                    auto accR = accelerated::cpu::Image::createReference(cfg.w, cfg.h, 1, accelerated::ImageTypeSpec::DataType::UFIXED8, imgR.data());
                    std::shared_ptr<const tracker::Camera> c1 = tracker::buildCamera(intrinsic, kind, params.tracker, cfg.w, cfg.h, params.tracker.secondDistortionCoeffs);
                    r = fac.buildStereo(*accL, *accR, c0, c1);                                          // api.cpp:603-606
                } else {
                    r.first = fac.build(*accL, c0);                                                    // api.cpp:553
                }
                P.queue->processAll();                                                                 // api.cpp:622
                return r;
            };
            auto images = makeImages(*P.trackerImages);
            if (P.shadowImages) {
                harness::setFlavour(Flavour::CUDA);
                auto shadow = makeImages(*P.shadowImages);
                harness::setFlavour(P.flavour);
                harness::registerShadowImage(images.first.get(), std::shared_ptr<tracker::Image>(std::move(shadow.first)));
                if (images.second) harness::registerShadowImage(images.second.get(), std::shared_ptr<tracker::Image>(std::move(shadow.second)));
            }
            if (cfg.stereo) P.control->processStereoFrames(tFrame, std::move(images.first), std::move(images.second), {});
            else P.control->processFrame(tFrame, std::move(images.first), {});
            for (const Imu& s : imu) {
                P.control->processAccelerometerSample(s.t, api::Vector3d { s.a[0], s.a[1], s.a[2] });
                P.control->processGyroSample(s.t, api::Vector3d { s.g[0], s.g[1], s.g[2] });
                harness::setFrameIndex(P.framesOut);
                const auto tf0 = clk::now();
                const auto res = P.control->processSyncedSamples(2);                                   // api.cpp:424-428
                if (res == odometry::Control::SampleProcessResult::FRAMES) {
                    P.frameMs.push_back(std::chrono::duration<double, std::milli>(clk::now() - tf0).count());
                    const odometry::Output o = P.control->getOutput();
                    const odometry::EKF& ekf = P.control->getEKF();
                    Pipeline::Out rec;
                    rec.t = o.t; rec.p = ekf.position(); rec.q = ekf.orientation(); rec.status = (int)o.trackingStatus;
                    rec.m = ekf.getState();
                    const Eigen::MatrixXd& cov = ekf.getStateCovarianceRef();
                    rec.Pmax = cov.cwiseAbs().maxCoeff();
                    if (keepP || mode == "free") rec.P = cov;
                    P.outs.push_back(std::move(rec));
                    P.framesOut++;
                }
            }
            P.seconds += std::chrono::duration<double>(clk::now() - t0).count();
        }
    }

    // ------------------------------------------------------------------------------------------------ report
    std::ostringstream js; js.precision(9);
    js << "{\n \"mode\": \"" << mode << "\", \"config\": \"" << cfg.name << "\", \"frames_fed\": " << frames
       << ", \"state_dim\": " << (20 + 7 * cfg.trail) << ", \"reference_opencv_threads\": " << harness::refThreads() << ",\n \"pipelines\": [";
    for (size_t i = 0; i < pipes.size(); i++) {
        Pipeline& P = *pipes[i];
        int tracking = 0, firstTracking = -1;
        for (size_t k = 0; k < P.outs.size(); k++) if (P.outs[k].status == (int)api::TrackingStatus::TRACKING) { tracking++; if (firstTracking < 0) firstTracking = (int)k; }
        std::vector<double> sorted = P.frameMs; std::sort(sorted.begin(), sorted.end());
        // error against the ground-truth trajectory (the filter starts at the origin with the initial orientation: compare path length scale only)
        double gtErr = 0;
        if (!P.outs.empty()) {
            const auto& last = P.outs.back();
            const synth::Pose gt = traj.at(last.t), gt0 = traj.at(0);
            gtErr = ((last.p - P.outs.front().p) - (gt.p - gt0.p)).norm();
        }
        js << (i ? "," : "") << "\n  {\"name\": \"" << P.name << "\", \"frames_processed\": " << P.framesOut << ", \"frames_tracking\": " << tracking
           << ", \"first_tracking_frame\": " << firstTracking << ", \"wall_seconds\": " << P.seconds
           << ", \"median_frame_ms\": " << (sorted.empty() ? 0.0 : sorted[sorted.size() / 2])
           << ", \"final_position\": [" << (P.outs.empty() ? 0 : P.outs.back().p[0]) << "," << (P.outs.empty() ? 0 : P.outs.back().p[1]) << "," << (P.outs.empty() ? 0 : P.outs.back().p[2]) << "]"
           << ", \"position_error_vs_ground_truth_m\": " << gtErr << "}";
    }
    js << "\n ]";
    harness::Stats& S = harness::stats();
    auto trackSummary = [&](const std::vector<harness::FrameTracks>& a, const std::vector<harness::FrameTracks>& b) {
        // frame-by-frame equality of Tracker::Output: IDs and statuses bit-equal, points within tolerance
        std::ostringstream o; o.precision(6);
        const size_t n = std::min(a.size(), b.size());
        long tracks = 0; int firstDiff = -1; double maxPt = 0; long idDiff = 0, stDiff = 0;
        for (size_t k = 0; k < n; k++) {
            bool same = a[k].ids.size() == b[k].ids.size() && a[k].keyframe == b[k].keyframe;
            if (same) for (size_t j = 0; j < a[k].ids.size(); j++) {
                tracks++;
                if (a[k].ids[j] != b[k].ids[j]) { idDiff++; same = false; }
                else if (a[k].status[j] != b[k].status[j]) { stDiff++; same = false; }
                else if (firstDiff < 0) for (int c = 0; c < 4; c++) maxPt = std::max(maxPt, (double)std::fabs(a[k].pts[4 * j + c] - b[k].pts[4 * j + c]));
            }
            if (!same && firstDiff < 0) firstDiff = (int)k;
        }
        o << "{\"frames_compared\": " << n << ", \"tracks_compared\": " << tracks << ", \"first_divergence_frame\": " << firstDiff
          << ", \"id_differences\": " << idDiff << ", \"status_differences\": " << stDiff << ", \"max_point_diff_before_divergence_px\": " << maxPt << "}";
        return o.str();
    };
    if (mode == "lockstep") {
        js << ",\n \"lockstep\": {\n  \"pyramid\": {\"pyramids\": " << S.pyramidsCompared << ", \"levels\": " << S.pyramidLevelsCompared << ", \"mismatching_bytes\": " << S.pyramidMismatchBytes << "},\n"
           << "  \"lk\": {\"calls\": " << S.lkCalls << ", \"points\": " << S.lkPoints << ", \"tracked\": " << S.lkTracked << ", \"status_mismatch\": " << S.lkStatusMismatch
           << ", \"over_1e-3_px\": " << S.lkOver1e3 << ", \"max_diff_px\": " << S.lkMaxDiff << ", \"outliers\": [";
        for (size_t i = 0; i < S.lkOutliers.size() && i < 50; i++) { const auto& o = S.lkOutliers[i];
            js << (i ? "," : "") << "{\"frame\": " << o.frame << ", \"call\": " << o.call << ", \"index\": " << o.index << ", \"dx\": " << (std::isnan(o.dx) ? -1e9 : o.dx) << ", \"dy\": " << (std::isnan(o.dy) ? -1e9 : o.dy) << "}"; }
        js << "]},\n  \"ekf\": {\"compares\": " << S.ekfCompares << ", \"max_position_diff_m\": " << S.ekfMaxPos << ", \"max_state_diff\": " << S.ekfMaxM << ", \"max_cov_rel_diff\": " << S.ekfMaxPrel
           << ", \"outlier_checks\": " << S.ekfChecks << ", \"decision_mismatch\": " << S.ekfCheckMismatch << ", \"ops\": {";
        bool first = true;
        for (const auto& kv : S.ekfOps) { js << (first ? "" : ", ") << "\"" << kv.first << "\": {\"calls\": " << kv.second.calls << ", \"pos\": " << kv.second.maxPos << ", \"m\": " << kv.second.maxM << ", \"P_rel\": " << kv.second.maxPrel << "}"; first = false; }
        js << "},\n   \"position_diff_by_frame\": " << jsonVec(S.framePos) << ",\n   \"cov_rel_diff_by_frame\": " << jsonVec(S.framePrel) << "},\n"
           << "  \"detector\": {\"calls\": " << S.detCalls << ", \"corners\": " << S.detCorners << ", \"mismatch\": " << S.detMismatch << ", \"first_mismatch_frame\": " << S.detFirstMismatchFrame << "},\n"
           << "  \"undistorter\": {\"calls\": " << S.undCalls << ", \"pixels\": " << S.undPixels << ", \"mismatch\": " << S.undMismatch << ", \"undefined_in_reference\": " << S.undUndefined << "},\n"
           << "  \"tracker\": {\"frames\": " << S.trkFrames << ", \"tracks\": " << S.trkTracks << ", \"id_mismatch\": " << S.trkIdMismatch << ", \"status_mismatch\": " << S.trkStatusMismatch
           << ", \"size_mismatch\": " << S.trkSizeMismatch << ", \"keyframe_mismatch\": " << S.trkKeyframeMismatch << ", \"first_mismatch_frame\": " << S.trkFirstMismatchFrame
           << ", \"max_point_diff_px\": " << S.trkMaxPointDiff << "}\n }";
    }
    if (mode == "free") {
        const Pipeline& A = *pipes[0]; const Pipeline& B = *pipes[1];
        const size_t n = std::min(A.outs.size(), B.outs.size());
        std::vector<double> dpos, dang, dP; int firstOver = -1;
        for (size_t k = 0; k < n; k++) {
            const double d = (A.outs[k].p - B.outs[k].p).norm();
            dpos.push_back(d);
            const double dot = std::min(1.0, std::fabs(A.outs[k].q.dot(B.outs[k].q) / (A.outs[k].q.norm() * B.outs[k].q.norm())));
            dang.push_back(2 * std::acos(dot));
            dP.push_back((A.outs[k].P - B.outs[k].P).cwiseAbs().maxCoeff() / std::max(A.outs[k].Pmax, 1e-300));
            if (d > 1e-4 && firstOver < 0) firstOver = (int)k;
        }
        js << ",\n \"free_running\": {\n  \"tracker_output\": " << trackSummary(S.trackLog[0], S.trackLog[1]) << ",\n  \"first_frame_position_diff_over_1e-4_m\": " << firstOver
           << ",\n  \"position_diff_m_by_frame\": " << jsonVec(dpos) << ",\n  \"orientation_diff_rad_by_frame\": " << jsonVec(dang) << ",\n  \"cov_rel_diff_by_frame\": " << jsonVec(dP) << "\n }";
    }
    js << "\n}\n";
    if (!outPath.empty()) { std::ofstream f(outPath); f << js.str(); }
    std::fputs(js.str().c_str(), stdout);
    pipes.clear();
    return 0;
}
