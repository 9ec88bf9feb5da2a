// hybvio_b200/host/cuda_ekf.cpp -- odometry::EKF implemented on top of the hv_ekf_* C ABI (include/hybvio_b200.h).
//
// Drop-in replacement for the reference's src/odometry/ekf.cpp: it defines odometry::EKF::build / ~EKF and a class
// with the same 56 virtuals (src/odometry/ekf.hpp:62-174), so backend.cpp, triangulation.cpp, output.cpp, api.cpp, the
// viewers and the reference's own unit tests (test/ekf.cpp, test/triangulation.cpp) compile and link UNCHANGED when this
// file is compiled in place of ekf.cpp. State m and covariance P live on the GPU; the getters that
// return Eigen references (getState, getStateCovarianceRef) are served from a lazily synchronised host mirror.
// There is no CPU fallback: build() aborts with the library's error text if no B200 context can be created.
#include "ekf.hpp"
#include "parameters.hpp"
#include "../../include/hybvio_b200.h"
#include "cuda_context.hpp"
#include "cuda_track_model.hpp"

#include <Eigen/Eigenvalues>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <sstream>

namespace {
using namespace odometry;

using hybvio_b200::sharedContext;   // ONE context / stream per process, shared with the tracker back ends (cuda_context.hpp)

[[noreturn]] void fail(const char* what) { hybvio_b200::hvFail(what); }
// Programming errors and CUDA errors abort (the reference asserts in the same places); NUMERICAL conditions do not:
// see visualTrackOutlierCheck / updateVisualTrack below.
#define HV(call) do { if ((call) != HV_OK) fail(#call); } while (0)

struct CudaEKF : public EKF {
    const Parameters& parameters;
    hv_ekf* h = nullptr;
    const int camPoseCount, hybridMapDim, stateDim;
    const double noiseScale;
    mutable Eigen::VectorXd m;          // host mirrors
    mutable Eigen::MatrixXd P;
    mutable bool mStale = true, PStale = true, inertialOnly = false;   // inertialOnly: the mirror is stale in its 20 inertial states only

    explicit CudaEKF(const Parameters& p)
        : parameters(p), camPoseCount(p.odometry.cameraTrailLength), hybridMapDim(p.odometry.hybridMapSize * MAP_POINT_DIM),
          stateDim(INER_DIM + camPoseCount * POSE_DIM + hybridMapDim), noiseScale(p.odometry.noiseScale * p.odometry.noiseScale) {
        const ParametersOdometry& po = p.odometry;
        hv_ekf_params q;
        q.camera_trail_length = po.cameraTrailLength; q.hybrid_map_size = po.hybridMapSize;
        q.noise_scale = po.noiseScale; q.gravity = po.gravity;
        q.noise_initial_pos = po.noiseInitialPos; q.noise_initial_vel = po.noiseInitialVel; q.noise_initial_ori = po.noiseInitialOri;
        q.noise_initial_bga = po.noiseInitialBGA; q.noise_initial_baa = po.noiseInitialBAA; q.noise_initial_bat = po.noiseInitialBAT;
        q.noise_initial_sft = po.noiseInitialSFT; q.noise_initial_pos_trail = po.noiseInitialPosTrail; q.noise_initial_ori_trail = po.noiseInitialOriTrail;
        q.noise_process_acc = po.noiseProcessAcc; q.noise_process_gyro = po.noiseProcessGyro;
        q.noise_process_baa = po.noiseProcessBAA; q.noise_process_baa_rev = po.noiseProcessBAARev;
        q.noise_process_bga = po.noiseProcessBGA; q.noise_process_bga_rev = po.noiseProcessBGARev;
        q.augment_r = po.augmentR; q.init_zupt_r = po.initZuptR; q.rotation_zupt_r = po.rotationZuptR;
        HV(hv_ekf_create(sharedContext(), &q, &h));
        m.resize(stateDim); P.resize(stateDim, stateDim);
    }
    CudaEKF(const CudaEKF& o)
        : EKF(o), parameters(o.parameters), camPoseCount(o.camPoseCount), hybridMapDim(o.hybridMapDim), stateDim(o.stateDim),
          noiseScale(o.noiseScale), m(o.m), P(o.P), mStale(o.mStale), PStale(o.PStale), inertialOnly(o.inertialOnly) {
        HV(hv_ekf_clone(o.h, &h));
    }
    ~CudaEKF() override { hv_ekf_destroy(h); }
    std::unique_ptr<EKF> clone() const final { return std::unique_ptr<EKF>(new CudaEKF(*this)); }

    void touched() { mStale = true; inertialOnly = false; PStale = true; }
    // predict() and normalizeQuaternions(true) change the 20 inertial states only: behind a burst of them the mirror is refreshed from the
    // mean launch of the queued burst (hv_ekf_predicted_mean: ~5 us + 160 bytes) instead of the full launch and the whole mean -- what
    // the flow predictor's position() / orientation() reads cost (backend.cpp:547-600)
    void predicted() { if (!mStale) inertialOnly = true; mStale = true; PStale = true; }
    const Eigen::VectorXd& mean() const {
        if (mStale) {
            if (inertialOnly) HV(hv_ekf_predicted_mean(h, m.data()));
            else HV(hv_ekf_download(h, m.data(), nullptr));
            mStale = false; inertialOnly = false;
        }
        return m;
    }
    const Eigen::MatrixXd& cov() const { if (PStale) { HV(hv_ekf_download(h, nullptr, P.data())); PStale = false; } return P; }

    void initializeOrientation(const Eigen::Vector3d& xa) final { HV(hv_ekf_initialize_orientation(h, xa.data())); touched(); }
    void predict(double t, const Eigen::Vector3d& xg, const Eigen::Vector3d& xa) final { HV(hv_ekf_predict(h, t, xg.data(), xa.data())); predicted(); }
    Eigen::Vector3d position() const final { return mean().segment(POS, 3); }
    Eigen::Vector3d velocity() const final { return mean().segment(VEL, 3); }
    Eigen::Vector4d orientation() const final { return mean().segment(ORI, 4); }
    Eigen::Vector3d biasGyroscopeAdditive() const final { return mean().segment(BGA, 3); }
    Eigen::Vector3d biasAccelerometerAdditive() const final { return mean().segment(BAA, 3); }
    Eigen::Vector3d biasAccelerometerTransform() const final { return mean().segment(BAT, 3); }
    int camTrailSize() const final { return camPoseCount; }
    Eigen::Vector3d historyPosition(int i) const final { return i == -1 ? position() : Eigen::Vector3d(mean().segment(CAM + POSE_DIM * i, 3)); }
    Eigen::Vector4d historyOrientation(int i) const final { return i == -1 ? orientation() : Eigen::Vector4d(mean().segment(CAM + POSE_DIM * i + 3, 4)); }
    double historyTime(int i) const final { return hv_ekf_history_time(h, i); }
    double speed() const final { return mean().segment(VEL, 3).norm(); }
    double horizontalSpeed() const final { return mean().segment(VEL, 2).norm(); }
    void updateZupt(double r) final { HV(hv_ekf_update_zupt(h, r)); touched(); }
    void updateZuptInitialization() final { HV(hv_ekf_update_zupt_initialization(h)); touched(); }
    void updateZrupt(const Eigen::Vector3d& xg) final { HV(hv_ekf_update_zrupt(h, xg.data())); touched(); }
    void updatePseudoVelocity(double defaultSpeed, double r) final { HV(hv_ekf_update_pseudo_velocity(h, defaultSpeed, r)); touched(); }
    void updatePosition(const Eigen::Vector3d& pos, double r) final { HV(hv_ekf_update_position(h, pos.data(), r)); touched(); }
    void updateZeroHeight(double r) final { HV(hv_ekf_update_zero_height(h, r)); touched(); }
    void updateOrientation(const Eigen::Vector4d& q, double r) final { HV(hv_ekf_update_orientation(h, q.data(), r)); touched(); }

    void getInertialState(VectorInertialMean& mean_, MatrixInertialCov& cov_) const final { HV(hv_ekf_download_inertial(h, mean_.data(), cov_.data())); }
    void setInertialState(const VectorInertialMean& mean_, const MatrixInertialCov& cov_) final { HV(hv_ekf_set_inertial_state(h, mean_.data(), cov_.data())); touched(); }
    double getImuToCameraTimeShift() const final { return mean()(SFT); }
    void translateTo(const Eigen::Vector3d& pos) final { HV(hv_ekf_translate_to(h, pos.data())); touched(); }
    void transformTo(const Eigen::Vector3d& pos, const Eigen::Vector4d& q, int i = -1) final { HV(hv_ekf_transform_to(h, pos.data(), q.data(), i)); touched(); }

    VuOutlierStatus visualTrackOutlierCheck(const Eigen::MatrixXd& visH, const Eigen::VectorXd& f, const Eigen::VectorXd& y, double r,
                                            double trackRmseThreshold) final {
        int st = 0;
        const int rc = hv_ekf_visual_check(h, visH.data(), (int)visH.rows(), (int)visH.cols(), f.data(), y.data(), r, trackRmseThreshold, &st, nullptr);
        if (rc == HV_ERR_STATE) {
            // innovation covariance not positive definite: the reference's pivoted LDLT carries on with whatever it gets
            // (ekf.cpp:787-819); here the track is simply not used this frame instead of killing the host process
            std::fprintf(stderr, "hybvio_b200: visualTrackOutlierCheck: %s -> NOT_COMPUTED\n", hv_last_error());
            return VuOutlierStatus::NOT_COMPUTED;
        }
        if (rc != HV_OK) fail("hv_ekf_visual_check");
        return static_cast<VuOutlierStatus>(st);
    }
    void updateVisualTrack(const Eigen::MatrixXd& visH, const Eigen::VectorXd& f, const Eigen::VectorXd& y, double r) final {
        const int rc = hv_ekf_visual_update(h, visH.data(), (int)visH.rows(), (int)visH.cols(), f.data(), y.data(), r);
        if (rc == HV_ERR_STATE) { std::fprintf(stderr, "hybvio_b200: updateVisualTrack skipped: %s\n", hv_last_error()); return; }
        if (rc != HV_OK) fail("hv_ekf_visual_update");
        touched();
    }
    void updateVisualPoseAugmentation(int discardedPoseIndex = -1) final { HV(hv_ekf_augment(h, discardedPoseIndex)); touched(); }
    void updateUndoAugmentation() final { HV(hv_ekf_unaugment(h)); touched(); }

    Eigen::Vector3d getMapPoint(int idx) const final { return mean().segment<3>(getMapPointStateIndex(idx)); }
    void insertMapPoint(int idx, const Eigen::Vector3d& pf) final { HV(hv_ekf_insert_map_point(h, idx, pf.data())); touched(); }
    int getMapPointStateIndex(int idx) const final { return idx == -1 ? -1 : stateDim - hybridMapDim + idx * MAP_POINT_DIM; }

    void conditionOnLastPose() final { HV(hv_ekf_condition_on_last_pose(h)); touched(); }
    void lockBiases() final { HV(hv_ekf_lock_biases(h)); touched(); }
    void normalizeQuaternions(bool onlyCurrent) final {
        HV(hv_ekf_normalize_quaternions(h, onlyCurrent ? 1 : 0));
        if (onlyCurrent) { if (!mStale) inertialOnly = true; mStale = true; } else { mStale = true; inertialOnly = false; }
    }
    void setFirstSampleTime(double t) final { HV(hv_ekf_set_first_sample_time(h, t)); }
    bool isPositiveSemiDefinite() final {   // "Expensive, use only for debugging" (ekf.cpp:1043-1057)
        Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> es(cov());
        return es.info() == Eigen::Success && es.eigenvalues().minCoeff() >= 0.0;
    }
    void maintainPositiveSemiDefinite() final { HV(hv_ekf_symmetrize(h)); PStale = true; }
    void setState(const Eigen::VectorXd& m_) final { HV(hv_ekf_upload(h, m_.data(), nullptr)); mStale = true; inertialOnly = false; }
    void setStateCovariance(const Eigen::MatrixXd& P_) final { HV(hv_ekf_upload(h, nullptr, P_.data())); PStale = true; }
    void setProcessNoise(const Eigen::MatrixXd& Q_) final { Eigen::Matrix<double, Q_DIM, Q_DIM> q = Q_; HV(hv_ekf_set_process_noise(h, q.data())); }
    double getPlatformTime() const final { return hv_ekf_platform_time(h); }
    int getPoseCount() const final { return hv_ekf_pose_count(h); }
    const Eigen::VectorXd& getState() const final { return mean(); }
    Eigen::MatrixXd getStateCovariance() const final { return cov(); }
    const Eigen::MatrixXd& getStateCovarianceRef() const final { return cov(); }
    int getStateDim() const final { return stateDim; }
    bool getWasStationary() const final { return hv_ekf_was_stationary(h) != 0; }

    // constant structure matrices, only used by debug viewers / tests (ekf.cpp:229-291)
    Eigen::MatrixXd getVisAugH() const final {
        Eigen::MatrixXd H = Eigen::MatrixXd::Zero(POSE_DIM, stateDim);
        for (int i = 0; i < 3; i++) { H(i, POS + i) = 1; H(i, CAM + i) = -1; }
        for (int i = 0; i < 4; i++) { H(3 + i, ORI + i) = 1; H(3 + i, CAM + 3 + i) = -1; }
        return H;
    }
    Eigen::MatrixXd getVisAugA() const final {   // drops the last pose of the trail
        Eigen::MatrixXd A = Eigen::MatrixXd::Zero(stateDim, stateDim);
        const int drop = camPoseCount - 1;
        for (int i = 0; i < CAM; i++) A(i, i) = 1;
        for (int i = CAM; i < CAM + drop * POSE_DIM; ++i) A(i + POSE_DIM, i) = 1;
        for (int i = CAM + (drop + 1) * POSE_DIM; i < stateDim; i++) A(i, i) = 1;
        return A;
    }
    Eigen::MatrixXd getVisAugQ() const final {
        Eigen::MatrixXd Q = Eigen::MatrixXd::Zero(stateDim, stateDim);
        const ParametersOdometry& po = parameters.odometry;
        for (int i = CAM; i < CAM + 3; i++) Q(i, i) = po.noiseInitialPosTrail * po.noiseInitialPosTrail * noiseScale;
        for (int i = CAM + 3; i < CAM + POSE_DIM; i++) Q(i, i) = po.noiseInitialOriTrail * po.noiseInitialOriTrail * noiseScale;
        return Q;
    }
    Eigen::MatrixXd getDydx() const final {
        Eigen::MatrixXd full = Eigen::MatrixXd::Identity(stateDim, stateDim);
        Eigen::Matrix<double, INER_DIM, INER_DIM> d;
        HV(hv_ekf_get_dydx(h, d.data()));
        full.topLeftCorner<INER_DIM, INER_DIM>() = d;
        return full;
    }
    std::string stateAsString() const final {   // same layout as ekf.cpp:997-1022
        std::stringstream ss;
        Eigen::Matrix<double, INER_DIM, 1> var = cov().block(0, 0, INER_DIM, INER_DIM).diagonal();
        for (size_t i = 0; i < STATE_PARTS.size(); i++) {
            const int part = STATE_PARTS[i], size = STATE_PART_SIZES[i];
            ss << STATE_PART_NAMES[i] << " ";
            for (int j = 0; j < size; j++) ss << std::setprecision(3) << mean()(part + j) << " ";
            ss << std::setprecision(2) << " [" << std::sqrt(var.segment(part, size).maxCoeff()) << "], ";
            if (i == 2) ss << std::endl << " ";
        }
        ss << std::fixed << std::setprecision(3) << "t " << (hv_ekf_platform_time(h));
        return ss.str();
    }
};
} // namespace

namespace odometry {
EKF::~EKF() = default;
EKF::EKF(const EKF& other) = default;
EKF::EKF() {}
std::unique_ptr<EKF> buildCudaEKF(const Parameters& parameters) { return std::unique_ptr<EKF>(new CudaEKF(parameters)); }
#ifndef HV_PIPELINE_HARNESS
// compiled INSTEAD of src/odometry/ekf.cpp: the reference's factory symbol. (The parity harness oracle/ref_build/pipeline defines
// HV_PIPELINE_HARNESS and provides its own EKF::build that chooses between the reference, this class and a lock-step pair.)
std::unique_ptr<EKF> EKF::build(const Parameters& parameters) { return buildCudaEKF(parameters); }
#endif

// ---- cuda_track_model.hpp: the per-track measurement model on the device (backend.cpp:1050-1160)
static CudaEKF& cudaEkf(EKF& ekf) {
    CudaEKF* c = dynamic_cast<CudaEKF*>(&ekf);
    if (!c) { std::fprintf(stderr, "hybvio_b200: the EKF was not built by the CUDA EKF::build\n"); std::abort(); }
    return *c;
}

static void setCameraModel(CudaEKF& e, const Parameters& parameters) {
    const ParametersOdometry& po = parameters.odometry;
    hv_camera_model cam;
    hv_camera_model_defaults(&cam);
    Eigen::Map<Eigen::Matrix4d>(cam.imu_to_camera) = parameters.imuToCamera;
    Eigen::Map<Eigen::Matrix4d>(cam.second_imu_to_camera) = parameters.secondImuToCamera;
    cam.use_stereo = parameters.tracker.useStereo ? 1 : 0;
    cam.estimate_imu_camera_time_shift = po.estimateImuCameraTimeShift ? 1 : 0;
    cam.gauss_newton_iterations = po.triangulationGaussNewtonIterations;
    cam.convergence_threshold = po.triangulationConvergenceThreshold; cam.convergence_r = po.triangulationConvergenceR;
    cam.rcond_threshold = po.triangulationRcondThreshold; cam.min_dist = po.triangulationMinDist; cam.max_dist = po.triangulationMaxDist;
    HV(hv_ekf_set_camera_model(e.h, &cam));
}

static std::vector<hv_track_obs> toObs(const std::vector<CudaTrackIn>& in) {
    std::vector<hv_track_obs> obs(in.size());
    for (size_t k = 0; k < in.size(); k++) {
        obs[k].npose = (int)in[k].poseTrailIndex->size();
        obs[k].pose_trail_index = in[k].poseTrailIndex->data();
        obs[k].ip = in[k].imageFeatures->front().data();             // contiguous Vector2d storage: x0 y0 x1 y1 ...
        obs[k].velocities = in[k].featureVelocities->front().data();
    }
    return obs;
}

void cudaTrackModels(EKF& ekf, const Parameters& parameters, const std::vector<CudaTrackIn>& in, std::vector<CudaTrackOut>& out) {
    CudaEKF& e = cudaEkf(ekf);
    setCameraModel(e, parameters);
    const std::vector<hv_track_obs> obs = toObs(in);
    std::vector<hv_track_model> res(in.size());
    HV(hv_ekf_track_models(e.h, obs.data(), (int)obs.size(), res.data()));
    out.resize(in.size());
    for (size_t k = 0; k < in.size(); k++) {
        CudaTrackOut& o = out[k];
        o.triangulateStatus = static_cast<TriangulatorStatus>(res[k].triangulator_status);
        o.prepareVuStatus = static_cast<PrepareVuStatus>(res[k].prepare_vu_status < 0 ? 0 : res[k].prepare_vu_status);
        o.pf = Eigen::Vector3d(res[k].pf[0], res[k].pf[1], res[k].pf[2]);
        o.depth = res[k].depth; o.rows = res[k].rows; o.cols = res[k].cols;
        o.dH = res[k].d_H; o.df = res[k].d_f; o.dy = res[k].d_y; o.index = (int)k;
    }
}

int cudaVisualTracks(EKF& ekf, const Parameters& parameters, const std::vector<CudaTrackIn>& in, double chiOutlierR, double rmseThreshold,
                     double visualR, int maxSuccessfulUpdates, int lookahead, std::vector<CudaTrackResult>& out) {
    CudaEKF& e = cudaEkf(ekf);
    setCameraModel(e, parameters);
    const std::vector<hv_track_obs> obs = toObs(in);
    hv_visual_update_params p;
    p.chi_outlier_r = chiOutlierR; p.track_rmse_threshold = rmseThreshold; p.visual_r = visualR;
    p.max_successful_updates = maxSuccessfulUpdates; p.lookahead = lookahead;
    std::vector<hv_track_result> res(in.size());
    int succ = 0;
    HV(hv_ekf_visual_tracks(e.h, obs.data(), (int)obs.size(), &p, res.data(), &succ));
    e.touched();
    out.resize(in.size());
    for (size_t k = 0; k < in.size(); k++) {
        CudaTrackResult& o = out[k];
        o.attempted = res[k].triangulator_status >= 0;
        o.triangulateStatus = static_cast<TriangulatorStatus>(o.attempted ? res[k].triangulator_status : 0);
        o.prepareVuStatus = static_cast<PrepareVuStatus>(res[k].prepare_vu_status < 0 ? 0 : res[k].prepare_vu_status);
        o.outlierStatus = static_cast<VuOutlierStatus>(res[k].outlier_status);
        o.updated = res[k].updated != 0; o.chi2 = res[k].chi2; o.depth = res[k].depth;
        o.pf = Eigen::Vector3d(res[k].pf[0], res[k].pf[1], res[k].pf[2]);
    }
    return succ;
}

static hv_track_model toAbi(const CudaTrackOut& t) {
    hv_track_model m;
    m.triangulator_status = static_cast<int>(t.triangulateStatus); m.prepare_vu_status = static_cast<int>(t.prepareVuStatus);
    m.rows = t.rows; m.cols = t.cols; m.pf[0] = t.pf(0); m.pf[1] = t.pf(1); m.pf[2] = t.pf(2); m.depth = t.depth;
    m.d_H = t.dH; m.d_f = t.df; m.d_y = t.dy;
    return m;
}

VuOutlierStatus cudaVisualTrackOutlierCheck(EKF& ekf, const CudaTrackOut& track, double r, double trackRmseThreshold) {
    const hv_track_model m = toAbi(track);
    int st = 0;
    HV(hv_ekf_visual_track(cudaEkf(ekf).h, &m, r, trackRmseThreshold, 0, &st, nullptr));
    return static_cast<VuOutlierStatus>(st);
}

void cudaUpdateVisualTrack(EKF& ekf, const CudaTrackOut& track, double r) {
    CudaEKF& e = cudaEkf(ekf);
    const hv_track_model m = toAbi(track);
    HV(hv_ekf_visual_track(e.h, &m, r, -1.0, 1, nullptr, nullptr));
    e.touched();
}

void cudaTrackModelDownload(EKF& ekf, const CudaTrackOut& track, Eigen::MatrixXd& H, Eigen::VectorXd& f) {
    H.resize(track.rows, track.cols); f.resize(track.rows);
    HV(hv_ekf_track_model_download(cudaEkf(ekf).h, track.index, H.data(), f.data(), nullptr));
}
} // namespace odometry
