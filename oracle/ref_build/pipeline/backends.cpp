// TEST INFRASTRUCTURE: defines the three factory symbols through which the UNMODIFIED reference core chooses its back ends
//   tracker::ImagePyramid::Factory::buildOpenCv / tracker::OpticalFlow::buildOpenCv   (src/tracker/image.cpp:55-56)
//   odometry::EKF::build                                                              (src/odometry/backend.cpp:187)
// and hands out the reference's classes, the CUDA classes or a lock-step pair of both (backends.hpp). tracker::Tracker::build
// (src/tracker/tracker.cpp:161-163) is intercepted at LINK time (-Wl,--wrap) so that tracker.cpp stays as it is: the wrapper
// records every Tracker::Output and, in lock-step mode, runs a second TrackerImplementation on CUDA-flavoured images next to
// the reference-driven one and compares IDs / statuses / points frame by frame.
#include "backends.hpp"

#include "ekf.hpp"
#include "feature_detector.hpp"
#include "image.hpp"
#include "image_pyramid.hpp"
#include "optical_flow.hpp"
#include "parameters.hpp"
#include "tracker.hpp"
#include "undistorter.hpp"
#include "camera.hpp"
#include "hybvio_b200.h"
#include "undistort_table.hpp"
#include <accelerated-arrays/cpu/image.hpp>

#include <opencv2/core.hpp>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_map>

// the reference's own back ends (oracle/_ref/libref_backends.so, ref_backends_shim.cpp)
extern "C" {
tracker::ImagePyramid::Factory* hv_ref_build_pyramid_factory(const odometry::ParametersTracker*);
tracker::OpticalFlow* hv_ref_build_optical_flow(const odometry::ParametersTracker*);
odometry::EKF* hv_ref_build_ekf(const odometry::Parameters*);
void hv_ref_set_num_threads(int);
int hv_ref_get_num_threads(void);
}
// the CUDA back ends (hybvio_b200/host/)
namespace tracker {
std::unique_ptr<ImagePyramid::Factory> buildCudaImagePyramidFactory(const odometry::ParametersTracker&);
std::unique_ptr<OpticalFlow> buildCudaOpticalFlow(const odometry::ParametersTracker&);
hv_pyr* cudaPyramidHandle(ImagePyramid&);
std::unique_ptr<FeatureDetector> buildCudaFeatureDetector(int w, int h, const odometry::ParametersTracker&);
std::unique_ptr<Undistorter> buildCudaUndistorter(int w, int h, std::shared_ptr<const Camera> rectifiedCamera, accelerated::Image::Factory& ifac,
                                                  const odometry::ParametersTracker& p);
}
namespace odometry { std::unique_ptr<EKF> buildCudaEKF(const Parameters&); }

namespace harness {
namespace {
Flavour g_flavour = Flavour::REF;
int g_frame = -1;
Stats g_stats;
std::unordered_map<tracker::Image*, std::shared_ptr<tracker::Image>> g_shadow;
}
bool g_cudaDetector = true;
void setUseCudaDetector(bool on) { g_cudaDetector = on; }
bool useCudaDetector() { return g_cudaDetector; }
void setFlavour(Flavour f) { g_flavour = f; }
Flavour flavour() { return g_flavour; }
void setFrameIndex(int f) { g_frame = f; }
void setRefThreads(int n) { hv_ref_set_num_threads(n); }
int refThreads() { return hv_ref_get_num_threads(); }
Stats& stats() { return g_stats; }
void registerShadowImage(tracker::Image* ref, std::shared_ptr<tracker::Image> cuda) {
    g_shadow[ref] = std::move(cuda);                 // consumed (erased) by DualTracker::add
}

namespace {
using namespace tracker;
using odometry::EKF;

// ------------------------------------------------------------------------------------------------ pyramid + LK, lock-step
struct DualPyramid : ImagePyramid {
    std::shared_ptr<ImagePyramid> ref, cuda;
    accelerated::Image& getGrayLevel(std::size_t i) final { return ref->getGrayLevel(i); }
    accelerated::Image& getGradientLevel(std::size_t i) final { return ref->getGradientLevel(i); }
    const std::vector<cv::Mat>& getOpenCv() final { return ref->getOpenCv(); }
};

struct DualPyramidFactory : ImagePyramid::Factory {
    std::unique_ptr<ImagePyramid::Factory> ref, cuda;
    std::vector<uint8_t> g; std::vector<int16_t> d;
    std::shared_ptr<ImagePyramid> compute(std::shared_ptr<accelerated::Image> img) final {
        auto p = std::make_shared<DualPyramid>();
        p->ref = ref->compute(img);
        p->cuda = cuda->compute(img);
        // every level, gray and gradients, bit for bit against the reference's cv::Mat views (ROI inside the padded buffers)
        const std::vector<cv::Mat>& mats = p->ref->getOpenCv();      // [g0, d0, g1, d1, ...] (lkpyramid.cpp:733-735)
        hv_pyr* h = cudaPyramidHandle(*p->cuda);
        const int levels = hv_pyr_levels(h);
        Stats& s = g_stats;
        s.pyramidsCompared++;
        if ((int)mats.size() != 2 * levels) { s.pyramidMismatchBytes += 1000000; return p; }
        for (int l = 0; l < levels; l++) {
            int w = 0, hh = 0;
            hv_pyr_level_size(h, l, &w, &hh);
            g.resize((size_t)w * hh); d.resize((size_t)w * hh * 2);
            if (hv_pyr_download_level(h, l, g.data(), d.data()) != HV_OK) { s.pyramidMismatchBytes += 1000000; continue; }
            const cv::Mat& G = mats[2 * l]; const cv::Mat& D = mats[2 * l + 1];
            if (G.cols != w || G.rows != hh || D.cols != w || D.rows != hh) { s.pyramidMismatchBytes += 1000000; continue; }
            for (int y = 0; y < hh; y++) {
                const uint8_t* gr = G.ptr<uint8_t>(y); const int16_t* dr = D.ptr<int16_t>(y);
                for (int x = 0; x < w; x++) if (gr[x] != g[(size_t)y * w + x]) s.pyramidMismatchBytes++;
                for (int x = 0; x < 2 * w; x++) if (dr[x] != d[(size_t)y * w * 2 + x]) s.pyramidMismatchBytes += 2;
            }
            s.pyramidLevelsCompared++;
        }
        return p;
    }
};

struct DualOpticalFlow : OpticalFlow {
    std::unique_ptr<OpticalFlow> ref, cuda;
    std::vector<Feature::Point> cornersC;
    std::vector<Feature::Status> statusC;
    void compute(ImagePyramid& prevP, ImagePyramid& curP, const std::vector<Feature::Point>& prevCorners,
                 std::vector<Feature::Point>& corners, std::vector<Feature::Status>& status, bool useInitial, int overrideMaxIter) final {
        auto& a = static_cast<DualPyramid&>(prevP);
        auto& b = static_cast<DualPyramid&>(curP);
        cornersC = corners; statusC = status;
        ref->compute(*a.ref, *b.ref, prevCorners, corners, status, useInitial, overrideMaxIter);
        cuda->compute(*a.cuda, *b.cuda, prevCorners, cornersC, statusC, useInitial, overrideMaxIter);
        Stats& s = g_stats;
        const int call = (int)s.lkCalls++;
        if (corners.size() != cornersC.size() || status.size() != statusC.size()) { s.lkStatusMismatch += 1000000; return; }
        for (size_t i = 0; i < status.size(); i++) {
            s.lkPoints++;
            if (status[i] != statusC[i]) { s.lkStatusMismatch++; s.lkOutliers.push_back({g_frame, call, (int)i, NAN, NAN}); continue; }
            if (status[i] != Feature::Status::TRACKED) continue;
            s.lkTracked++;
            const float dx = cornersC[i].x - corners[i].x, dy = cornersC[i].y - corners[i].y;
            const double e = std::max(std::fabs(dx), std::fabs(dy));
            s.lkMaxDiff = std::max(s.lkMaxDiff, e);
            if (!(e <= 1e-3)) { s.lkOver1e3++; s.lkOutliers.push_back({g_frame, call, (int)i, dx, dy}); }
        }
    }
};

// ------------------------------------------------------------------------------------------------ corner detector, lock-step
struct DualDetector : FeatureDetector {
    std::unique_ptr<FeatureDetector> ref, cuda;
    std::vector<Feature::Point> cornersC;
    DualDetector(const odometry::ParametersTracker& p, std::unique_ptr<FeatureDetector> r, std::unique_ptr<FeatureDetector> c)
        : FeatureDetector(p), ref(std::move(r)), cuda(std::move(c)) {}
    void compare(const std::vector<Feature::Point>& a, const std::vector<Feature::Point>& b) {
        Stats& s = g_stats;
        s.detCalls++; s.detCorners += (long)a.size();
        if (a.size() != b.size()) { s.detMismatch += 1 + (long)std::max(a.size(), b.size()) - (long)std::min(a.size(), b.size()); if (s.detFirstMismatchFrame < 0) s.detFirstMismatchFrame = g_frame; return; }
        for (size_t i = 0; i < a.size(); i++)
            if (a[i].x != b[i].x || a[i].y != b[i].y) { s.detMismatch++; if (s.detFirstMismatchFrame < 0) s.detFirstMismatchFrame = g_frame; }
    }
    void detect(Image& image, std::vector<Feature::Point>& corners, const std::vector<Feature::Point>& prev, int maskRadius) final {
        ref->detect(image, corners, prev, maskRadius);
        cuda->detect(image, cornersC, prev, maskRadius);
        compare(corners, cornersC);
    }
    accelerated::Future detect(accelerated::Image& image, std::vector<Feature::Point>& corners, const std::vector<Feature::Point>& prev, int maskRadius) final {
        ref->detect(image, corners, prev, maskRadius).wait();
        cuda->detect(image, cornersC, prev, maskRadius).wait();
        compare(corners, cornersC);
        return accelerated::Future::instantlyResolved();
    }
    bool supportsAsync() const final { return false; }
    void debugVisualize(cv::Mat& m) final { ref->debugVisualize(m); }
};

// ------------------------------------------------------------------------------------------------ undistortion / rectification, lock-step
// Both run on the same input; the images must be bit-identical wherever the reference's own bilinear taps stay inside its image buffer
// (it reads a tap right of the last column / below the last row without a bounds check, undistorter.cpp:101: those pixels are undefined in
// the reference itself and are counted separately).
struct DualUndistorter : Undistorter {
    std::unique_ptr<Undistorter> ref, cuda;
    std::shared_ptr<const Camera> rectified;
    int w, h;
    std::vector<hv_remap_entry> table;
    std::string tableFor;
    Result undistort(accelerated::Image& image, std::shared_ptr<const Camera> camera) final {
        Result a = ref->undistort(image, camera);
        Result b = cuda->undistort(image, camera);
        a.future.wait(); b.future.wait();
        if (camera->serialize() != tableFor) { hybvio_b200::buildUndistortTable(*a.camera, *camera, w, h, table); tableFor = camera->serialize(); }
        const uint8_t* pa = accelerated::cpu::Image::castFrom(*a.image).getDataRaw();
        const uint8_t* pb = accelerated::cpu::Image::castFrom(*b.image).getDataRaw();
        Stats& s = g_stats;
        s.undCalls++;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const hv_remap_entry& e = table[(size_t)y * w + x];
                const bool defined = e.x0 == HV_REMAP_INVALID_X0 || (e.x0 + 1 < w && e.y0 + 1 < h);
                if (!defined) { s.undUndefined++; continue; }
                s.undPixels++;
                if (pa[(size_t)y * w + x] != pb[(size_t)y * w + x]) s.undMismatch++;
            }
        return a;
    }
};

// ------------------------------------------------------------------------------------------------ EKF, lock-step
// Every call goes to both filters with identical arguments; getters and decisions come from the reference filter (so the
// pipeline follows the reference trajectory exactly); after every mutating call other than predict / normalizeQuaternions(true)
// -- whose batching (one launch per IMU burst) is part of what is being tested -- the two states are compared.
struct DualEKF : EKF {
    std::unique_ptr<EKF> r, c;
    DualEKF(std::unique_ptr<EKF> r_, std::unique_ptr<EKF> c_) : r(std::move(r_)), c(std::move(c_)) {}
    std::unique_ptr<EKF> clone() const final { return std::unique_ptr<EKF>(new DualEKF(r->clone(), c->clone())); }

    void compare(const char* op) {
        const Eigen::VectorXd& mr = r->getState(); const Eigen::VectorXd& mc = c->getState();
        const Eigen::MatrixXd& Pr = r->getStateCovarianceRef(); const Eigen::MatrixXd& Pc = c->getStateCovarianceRef();
        const double dm = (mr - mc).cwiseAbs().maxCoeff();
        double dpos = (mr.segment<3>(odometry::POS) - mc.segment<3>(odometry::POS)).cwiseAbs().maxCoeff();
        for (int i = 0; i < r->camTrailSize(); i++)       // every position of the pose trail counts as "pose"
            dpos = std::max(dpos, (mr.segment<3>(odometry::CAM + i * odometry::POSE_DIM) - mc.segment<3>(odometry::CAM + i * odometry::POSE_DIM)).cwiseAbs().maxCoeff());
        const double dP = (Pr - Pc).cwiseAbs().maxCoeff() / std::max(Pr.cwiseAbs().maxCoeff(), 1e-300);
        Stats& s = g_stats;
        OpStat& o = s.ekfOps[op];
        o.calls++; o.maxPos = std::max(o.maxPos, dpos); o.maxM = std::max(o.maxM, dm); o.maxPrel = std::max(o.maxPrel, dP);
        s.ekfCompares++;
        s.ekfMaxPos = std::max(s.ekfMaxPos, dpos); s.ekfMaxM = std::max(s.ekfMaxM, dm); s.ekfMaxPrel = std::max(s.ekfMaxPrel, dP);
        if (g_frame >= 0) {
            if ((int)s.framePos.size() <= g_frame) { s.framePos.resize(g_frame + 1, 0.0); s.framePrel.resize(g_frame + 1, 0.0); }
            s.framePos[g_frame] = std::max(s.framePos[g_frame], dpos); s.framePrel[g_frame] = std::max(s.framePrel[g_frame], dP);
        }
        if (!std::isfinite(dm) || !std::isfinite(dP)) std::fprintf(stderr, "harness: non-finite EKF difference after %s\n", op);
    }
#define BOTH(call, name) do { r->call; c->call; compare(name); } while (0)
    void initializeOrientation(const Eigen::Vector3d& xa) final { BOTH(initializeOrientation(xa), "initializeOrientation"); }
    void predict(double t, const Eigen::Vector3d& xg, const Eigen::Vector3d& xa) final { r->predict(t, xg, xa); c->predict(t, xg, xa); }
    Eigen::Vector3d position() const final { return r->position(); }
    Eigen::Vector3d velocity() const final { return r->velocity(); }
    Eigen::Vector4d orientation() const final { return r->orientation(); }
    Eigen::Vector3d biasGyroscopeAdditive() const final { return r->biasGyroscopeAdditive(); }
    Eigen::Vector3d biasAccelerometerAdditive() const final { return r->biasAccelerometerAdditive(); }
    Eigen::Vector3d biasAccelerometerTransform() const final { return r->biasAccelerometerTransform(); }
    int camTrailSize() const final { return r->camTrailSize(); }
    Eigen::Vector3d historyPosition(int i) const final { return r->historyPosition(i); }
    Eigen::Vector4d historyOrientation(int i) const final { return r->historyOrientation(i); }
    double historyTime(int i) const final {
        const double a = r->historyTime(i), b = c->historyTime(i);
        if (a != b) { g_stats.ekfCheckMismatch++; std::fprintf(stderr, "harness: historyTime(%d) %.9f != %.9f\n", i, a, b); }
        return a;
    }
    double speed() const final { return r->speed(); }
    double horizontalSpeed() const final { return r->horizontalSpeed(); }
    void updateZupt(double rr) final { BOTH(updateZupt(rr), "updateZupt"); }
    void updateZuptInitialization() final { BOTH(updateZuptInitialization(), "updateZuptInitialization"); }
    void updateZrupt(const Eigen::Vector3d& xg) final { BOTH(updateZrupt(xg), "updateZrupt"); }
    void updatePseudoVelocity(double d, double rr) final { BOTH(updatePseudoVelocity(d, rr), "updatePseudoVelocity"); }
    void updatePosition(const Eigen::Vector3d& p, double rr) final { BOTH(updatePosition(p, rr), "updatePosition"); }
    void updateZeroHeight(double rr) final { BOTH(updateZeroHeight(rr), "updateZeroHeight"); }
    void updateOrientation(const Eigen::Vector4d& q, double rr) final { BOTH(updateOrientation(q, rr), "updateOrientation"); }
    void getInertialState(VectorInertialMean& mean, MatrixInertialCov& cov) const final { r->getInertialState(mean, cov); }
    void setInertialState(const VectorInertialMean& mean, const MatrixInertialCov& cov) final { BOTH(setInertialState(mean, cov), "setInertialState"); }
    double getImuToCameraTimeShift() const final { return r->getImuToCameraTimeShift(); }
    void translateTo(const Eigen::Vector3d& pos) final { BOTH(translateTo(pos), "translateTo"); }
    void transformTo(const Eigen::Vector3d& pos, const Eigen::Vector4d& q, int i = -1) final { BOTH(transformTo(pos, q, i), "transformTo"); }
    odometry::VuOutlierStatus visualTrackOutlierCheck(const Eigen::MatrixXd& H, const Eigen::VectorXd& f, const Eigen::VectorXd& y, double rr,
                                                      double rmse) final {
        const auto a = r->visualTrackOutlierCheck(H, f, y, rr, rmse);
        const auto b = c->visualTrackOutlierCheck(H, f, y, rr, rmse);
        g_stats.ekfChecks++;
        if (a != b) { g_stats.ekfCheckMismatch++; std::fprintf(stderr, "harness: frame %d outlier check %d != %d (n=%d)\n", g_frame, (int)a, (int)b, (int)H.rows()); }
        compare("visualTrackOutlierCheck");      // must not have changed anything
        return a;
    }
    void updateVisualTrack(const Eigen::MatrixXd& H, const Eigen::VectorXd& f, const Eigen::VectorXd& y, double rr) final {
        BOTH(updateVisualTrack(H, f, y, rr), "updateVisualTrack");
    }
    void updateVisualPoseAugmentation(int k = -1) final { BOTH(updateVisualPoseAugmentation(k), "updateVisualPoseAugmentation"); }
    void updateUndoAugmentation() final { BOTH(updateUndoAugmentation(), "updateUndoAugmentation"); }
    Eigen::Vector3d getMapPoint(int idx) const final { return r->getMapPoint(idx); }
    void insertMapPoint(int idx, const Eigen::Vector3d& pf) final { BOTH(insertMapPoint(idx, pf), "insertMapPoint"); }
    int getMapPointStateIndex(int idx) const final { return r->getMapPointStateIndex(idx); }
    void conditionOnLastPose() final { BOTH(conditionOnLastPose(), "conditionOnLastPose"); }
    void lockBiases() final { BOTH(lockBiases(), "lockBiases"); }
    void normalizeQuaternions(bool onlyCurrent) final {
        r->normalizeQuaternions(onlyCurrent); c->normalizeQuaternions(onlyCurrent);
        if (!onlyCurrent) compare("normalizeQuaternions");
    }
    void setFirstSampleTime(double t) final { r->setFirstSampleTime(t); c->setFirstSampleTime(t); }
    bool isPositiveSemiDefinite() final { return r->isPositiveSemiDefinite(); }
    void maintainPositiveSemiDefinite() final { BOTH(maintainPositiveSemiDefinite(), "maintainPositiveSemiDefinite"); }
    void setState(const Eigen::VectorXd& m) final { BOTH(setState(m), "setState"); }
    void setStateCovariance(const Eigen::MatrixXd& P) final { BOTH(setStateCovariance(P), "setStateCovariance"); }
    void setProcessNoise(const Eigen::MatrixXd& Q) final { r->setProcessNoise(Q); c->setProcessNoise(Q); }
    double getPlatformTime() const final { return r->getPlatformTime(); }
    int getPoseCount() const final {
        const int a = r->getPoseCount(), b = c->getPoseCount();
        if (a != b) g_stats.ekfCheckMismatch++;
        return a;
    }
    const Eigen::VectorXd& getState() const final { return r->getState(); }
    Eigen::MatrixXd getStateCovariance() const final { return r->getStateCovariance(); }
    const Eigen::MatrixXd& getStateCovarianceRef() const final { return r->getStateCovarianceRef(); }
    Eigen::MatrixXd getVisAugH() const final { return r->getVisAugH(); }
    Eigen::MatrixXd getVisAugA() const final { return r->getVisAugA(); }
    Eigen::MatrixXd getVisAugQ() const final { return r->getVisAugQ(); }
    Eigen::MatrixXd getDydx() const final { return r->getDydx(); }
    std::string stateAsString() const final { return r->stateAsString(); }
    int getStateDim() const final { return r->getStateDim(); }
    bool getWasStationary() const final {
        const bool a = r->getWasStationary(), b = c->getWasStationary();
        if (a != b) g_stats.ekfCheckMismatch++;
        return a;
    }
#undef BOTH
};

// ------------------------------------------------------------------------------------------------ tracker wrapper
void logTracks(std::vector<FrameTracks>& log, const Tracker::Output& out) {
    FrameTracks f; f.frame = g_frame; f.keyframe = out.keyframe;
    for (const Feature& t : out.tracks) {
        f.ids.push_back(t.id); f.status.push_back((int)t.status);
        f.pts.push_back(t.points[0].x); f.pts.push_back(t.points[0].y); f.pts.push_back(t.points[1].x); f.pts.push_back(t.points[1].y);
    }
    log.push_back(std::move(f));
}

struct RecordingTracker : Tracker {
    std::unique_ptr<Tracker> t; int slot;
    RecordingTracker(std::unique_ptr<Tracker> t_, int slot_) : t(std::move(t_)), slot(slot_) {}
    void add(const TrackerArgsIn& args, Output& out) final { t->add(args, out); logTracks(g_stats.trackLog[slot], out); }
    void deleteTrack(int id) final { t->deleteTrack(id); }
};

// Lock-step: `r` sees the pipeline's own images (DUAL back ends, returning the reference's LK results), `c` sees the CUDA
// images of the same frames; same predictor callback, same poses, same deleteTrack calls. tracker.cpp is identical code
// in both, so any difference in IDs / statuses comes from the back ends.
struct DualTracker : Tracker {
    std::unique_ptr<Tracker> r, c;
    Output outC;
    DualTracker(std::unique_ptr<Tracker> r_, std::unique_ptr<Tracker> c_) : r(std::move(r_)), c(std::move(c_)) {}
    void deleteTrack(int id) final { r->deleteTrack(id); c->deleteTrack(id); }
    void add(const TrackerArgsIn& args, Output& out) final {
        r->add(args, out);
        logTracks(g_stats.trackLog[0], out);
        auto f = g_shadow.find(args.firstImage.get());
        if (f == g_shadow.end()) { std::fprintf(stderr, "harness: no CUDA shadow image registered for frame %d\n", g_frame); std::abort(); }
        std::shared_ptr<Image> first = f->second, second;
        g_shadow.erase(f);
        if (args.secondImage) { auto g = g_shadow.find(args.secondImage.get()); second = g->second; g_shadow.erase(g); }
        TrackerArgsIn argsC { first, second, args.t, args.opticalFlowPredictor, args.poses };
        const Flavour keep = g_flavour;
        g_flavour = Flavour::CUDA;
        c->add(argsC, outC);
        g_flavour = keep;
        logTracks(g_stats.trackLog[1], outC);
        Stats& s = g_stats;
        s.trkFrames++;
        bool bad = false;
        if (out.keyframe != outC.keyframe) { s.trkKeyframeMismatch++; bad = true; }
        if (out.tracks.size() != outC.tracks.size()) { s.trkSizeMismatch++; bad = true; }
        const size_t n = std::min(out.tracks.size(), outC.tracks.size());
        for (size_t i = 0; i < n; i++) {
            const Feature& a = out.tracks[i]; const Feature& b = outC.tracks[i];
            s.trkTracks++;
            if (a.id != b.id) { s.trkIdMismatch++; bad = true; continue; }
            if (a.status != b.status) { s.trkStatusMismatch++; bad = true; continue; }
            for (int k = 0; k < 2; k++) {
                const double e = std::max(std::fabs(a.points[k].x - b.points[k].x), std::fabs(a.points[k].y - b.points[k].y));
                s.trkMaxPointDiff = std::max(s.trkMaxPointDiff, e);
            }
        }
        if (bad && s.trkFirstMismatchFrame < 0) s.trkFirstMismatchFrame = g_frame;
    }
};
} // namespace
} // namespace harness

// ---------------------------------------------------------------------------------------------------- the factory symbols
namespace tracker {
std::unique_ptr<ImagePyramid::Factory> ImagePyramid::Factory::buildOpenCv(const odometry::ParametersTracker& p) {
    using namespace harness;
    switch (flavour()) {
    case Flavour::REF: return std::unique_ptr<Factory>(hv_ref_build_pyramid_factory(&p));
    case Flavour::CUDA: return buildCudaImagePyramidFactory(p);
    default: {
        auto d = std::make_unique<DualPyramidFactory>();
        d->ref.reset(hv_ref_build_pyramid_factory(&p)); d->cuda = buildCudaImagePyramidFactory(p);
        return d;
    }
    }
}
std::unique_ptr<OpticalFlow> OpticalFlow::buildOpenCv(const odometry::ParametersTracker& p) {
    using namespace harness;
    switch (flavour()) {
    case Flavour::REF: return std::unique_ptr<OpticalFlow>(hv_ref_build_optical_flow(&p));
    case Flavour::CUDA: return buildCudaOpticalFlow(p);
    default: {
        auto d = std::make_unique<DualOpticalFlow>();
        d->ref.reset(hv_ref_build_optical_flow(&p)); d->cuda = buildCudaOpticalFlow(p);
        return d;
    }
    }
}
ImagePyramid::~ImagePyramid() = default;
ImagePyramid::Factory::~Factory() = default;
OpticalFlow::~OpticalFlow() = default;
} // namespace tracker

namespace odometry {
std::unique_ptr<EKF> EKF::build(const Parameters& parameters) {
    using namespace harness;
    switch (flavour()) {
    case Flavour::REF: return std::unique_ptr<EKF>(hv_ref_build_ekf(&parameters));
    case Flavour::CUDA: return buildCudaEKF(parameters);
    default: return std::unique_ptr<EKF>(new DualEKF(std::unique_ptr<EKF>(hv_ref_build_ekf(&parameters)), buildCudaEKF(parameters)));
    }
}
} // namespace odometry

// tracker::FeatureDetector::build (src/tracker/image.cpp:52), intercepted with -Wl,--wrap as well: reference detector, CUDA detector
// (hybvio_b200/host/cuda_feature_detector.cpp) or both with the corner lists compared
extern "C" {
#define FD_BUILD _ZN7tracker15FeatureDetector5buildEiiRN11accelerated9ProcessorERNS1_5Image7FactoryERNS1_10operations15StandardFactoryERKN8odometry17ParametersTrackerE
#define FD_CAT2(a, b) a##b
#define FD_CAT(a, b) FD_CAT2(a, b)
std::unique_ptr<tracker::FeatureDetector> FD_CAT(__real_, FD_BUILD)(int, int, accelerated::Processor&, accelerated::Image::Factory&,
                                                                    accelerated::operations::StandardFactory&, const odometry::ParametersTracker&);
std::unique_ptr<tracker::FeatureDetector> FD_CAT(__wrap_, FD_BUILD)(int w, int h, accelerated::Processor& proc, accelerated::Image::Factory& ifac,
                                                                    accelerated::operations::StandardFactory& ofac, const odometry::ParametersTracker& p) {
    using namespace harness;
    if (flavour() == Flavour::REF || !useCudaDetector() || p.featureDetector != "GPU-GFTT") return FD_CAT(__real_, FD_BUILD)(w, h, proc, ifac, ofac, p);
    if (flavour() == Flavour::CUDA) return tracker::buildCudaFeatureDetector(w, h, p);
    return std::unique_ptr<tracker::FeatureDetector>(new DualDetector(p, FD_CAT(__real_, FD_BUILD)(w, h, proc, ifac, ofac, p), tracker::buildCudaFeatureDetector(w, h, p)));
}
}

// tracker::Undistorter::buildRectified / buildMono (src/tracker/image.cpp:323-336), intercepted with -Wl,--wrap
extern "C" {
#define UR_BUILD _ZN7tracker11Undistorter14buildRectifiedEiiSt10shared_ptrIKNS_6CameraEERN11accelerated5Image7FactoryERNS5_10operations15StandardFactoryERKN8odometry17ParametersTrackerE
std::unique_ptr<tracker::Undistorter> FD_CAT(__real_, UR_BUILD)(int, int, std::shared_ptr<const tracker::Camera>, accelerated::Image::Factory&,
                                                                accelerated::operations::StandardFactory&, const odometry::ParametersTracker&);
std::unique_ptr<tracker::Undistorter> FD_CAT(__wrap_, UR_BUILD)(int w, int h, std::shared_ptr<const tracker::Camera> cam, accelerated::Image::Factory& ifac,
                                                                accelerated::operations::StandardFactory& ofac, const odometry::ParametersTracker& p) {
    using namespace harness;
    auto real = FD_CAT(__real_, UR_BUILD)(w, h, cam, ifac, ofac, p);
    if (!real || flavour() == Flavour::REF) return real;
    if (flavour() == Flavour::CUDA) return tracker::buildCudaUndistorter(w, h, cam, ifac, p);
    auto d = std::make_unique<DualUndistorter>();
    d->ref = std::move(real); d->cuda = tracker::buildCudaUndistorter(w, h, cam, ifac, p); d->rectified = cam; d->w = w; d->h = h;
    return d;
}
}

// tracker::Tracker::build, intercepted with  -Wl,--wrap=_ZN7tracker7Tracker5buildERKN8odometry10ParametersE
extern "C" {
std::unique_ptr<tracker::Tracker> __real__ZN7tracker7Tracker5buildERKN8odometry10ParametersE(const odometry::Parameters&);
std::unique_ptr<tracker::Tracker> __wrap__ZN7tracker7Tracker5buildERKN8odometry10ParametersE(const odometry::Parameters& p) {
    using namespace harness;
    auto real = __real__ZN7tracker7Tracker5buildERKN8odometry10ParametersE(p);
    switch (flavour()) {
    case Flavour::REF: return std::unique_ptr<tracker::Tracker>(new RecordingTracker(std::move(real), 0));
    case Flavour::CUDA: return std::unique_ptr<tracker::Tracker>(new RecordingTracker(std::move(real), 1));
    default: return std::unique_ptr<tracker::Tracker>(new DualTracker(std::move(real), __real__ZN7tracker7Tracker5buildERKN8odometry10ParametersE(p)));
    }
}
}
