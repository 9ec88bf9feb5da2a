// TEST INFRASTRUCTURE ONLY (oracle/_ref): C shim around the *reference's own* odometry::EKF
// (src/odometry/ekf.hpp:62-174, implemented by src/odometry/ekf.cpp, compiled unmodified) so that tests/ and
// bench.py's cpu_baseline / --impl reference legs can drive it through ctypes with the same call shapes as the
// hv_ekf_* C ABI (include/hybvio_b200.h).
#include "ekf.hpp"
#include "parameters.hpp"
#include <cstring>
#include <memory>

namespace {
struct Params {   // same layout as hv_ekf_params
    int camera_trail_length, hybrid_map_size;
    double noise_scale, gravity;
    double noise_initial_pos, noise_initial_vel, noise_initial_ori, noise_initial_bga, noise_initial_baa, noise_initial_bat, noise_initial_sft;
    double noise_initial_pos_trail, noise_initial_ori_trail;
    double noise_process_acc, noise_process_gyro, noise_process_baa, noise_process_baa_rev, noise_process_bga, noise_process_bga_rev;
    double augment_r, init_zupt_r, rotation_zupt_r;
};
struct Ref {
    odometry::Parameters params;   // EKFImplementation keeps a reference to it
    std::unique_ptr<odometry::EKF> ekf;
};
using Eigen::Map; using Eigen::MatrixXd; using Eigen::VectorXd;
}

extern "C" {

void ref_ekf_default_params(Params* p) {
    odometry::Parameters d; const auto& o = d.odometry;
    p->camera_trail_length = o.cameraTrailLength; p->hybrid_map_size = o.hybridMapSize;
    p->noise_scale = o.noiseScale; p->gravity = o.gravity;
    p->noise_initial_pos = o.noiseInitialPos; p->noise_initial_vel = o.noiseInitialVel; p->noise_initial_ori = o.noiseInitialOri;
    p->noise_initial_bga = o.noiseInitialBGA; p->noise_initial_baa = o.noiseInitialBAA; p->noise_initial_bat = o.noiseInitialBAT;
    p->noise_initial_sft = o.noiseInitialSFT; p->noise_initial_pos_trail = o.noiseInitialPosTrail; p->noise_initial_ori_trail = o.noiseInitialOriTrail;
    p->noise_process_acc = o.noiseProcessAcc; p->noise_process_gyro = o.noiseProcessGyro;
    p->noise_process_baa = o.noiseProcessBAA; p->noise_process_baa_rev = o.noiseProcessBAARev;
    p->noise_process_bga = o.noiseProcessBGA; p->noise_process_bga_rev = o.noiseProcessBGARev;
    p->augment_r = o.augmentR; p->init_zupt_r = o.initZuptR; p->rotation_zupt_r = o.rotationZuptR;
}

void* ref_ekf_create(const Params* p) {
    Ref* r = new Ref; auto& o = r->params.odometry;
    o.cameraTrailLength = p->camera_trail_length; o.hybridMapSize = p->hybrid_map_size;
    o.noiseScale = p->noise_scale; o.gravity = p->gravity;
    o.noiseInitialPos = p->noise_initial_pos; o.noiseInitialVel = p->noise_initial_vel; o.noiseInitialOri = p->noise_initial_ori;
    o.noiseInitialBGA = p->noise_initial_bga; o.noiseInitialBAA = p->noise_initial_baa; o.noiseInitialBAT = p->noise_initial_bat;
    o.noiseInitialSFT = p->noise_initial_sft; o.noiseInitialPosTrail = p->noise_initial_pos_trail; o.noiseInitialOriTrail = p->noise_initial_ori_trail;
    o.noiseProcessAcc = p->noise_process_acc; o.noiseProcessGyro = p->noise_process_gyro;
    o.noiseProcessBAA = p->noise_process_baa; o.noiseProcessBAARev = p->noise_process_baa_rev;
    o.noiseProcessBGA = p->noise_process_bga; o.noiseProcessBGARev = p->noise_process_bga_rev;
    o.augmentR = p->augment_r; o.initZuptR = p->init_zupt_r; o.rotationZuptR = p->rotation_zupt_r;
    r->ekf = odometry::EKF::build(r->params);
    return r;
}
void* ref_ekf_clone(void* h) { Ref* s = (Ref*)h; Ref* r = new Ref; r->params = s->params; r->ekf = s->ekf->clone(); return r; }
// NB: the clone keeps referring to the source's Parameters object (EKFImplementation copies the reference), as in the reference's tests.
void ref_ekf_destroy(void* h) { delete (Ref*)h; }
#define E (((Ref*)h)->ekf)
int ref_ekf_state_dim(void* h) { return E->getStateDim(); }
int ref_ekf_pose_count(void* h) { return E->getPoseCount(); }
double ref_ekf_platform_time(void* h) { return E->getPlatformTime(); }
double ref_ekf_history_time(void* h, int i) { return E->historyTime(i); }
int ref_ekf_was_stationary(void* h) { return E->getWasStationary() ? 1 : 0; }
void ref_ekf_set_first_sample_time(void* h, double t) { E->setFirstSampleTime(t); }
void ref_ekf_upload(void* h, const double* m, const double* P) {
    int N = E->getStateDim();
    if (m) E->setState(Map<const VectorXd>(m, N));
    if (P) E->setStateCovariance(Map<const MatrixXd>(P, N, N));
}
void ref_ekf_download(void* h, double* m, double* P) {
    int N = E->getStateDim();
    if (m) Map<VectorXd>(m, N) = E->getState();
    if (P) Map<MatrixXd>(P, N, N) = E->getStateCovarianceRef();
}
void ref_ekf_download_inertial(void* h, double* m20, double* P20) {
    odometry::EKF::VectorInertialMean m; odometry::EKF::MatrixInertialCov c; E->getInertialState(m, c);
    std::memcpy(m20, m.data(), sizeof(double) * 20); std::memcpy(P20, c.data(), sizeof(double) * 400);
}
void ref_ekf_set_inertial_state(void* h, const double* m20, const double* P20) {
    odometry::EKF::VectorInertialMean m = Map<const odometry::EKF::VectorInertialMean>(m20);
    odometry::EKF::MatrixInertialCov c = Map<const odometry::EKF::MatrixInertialCov>(P20);
    E->setInertialState(m, c);
}
void ref_ekf_set_process_noise(void* h, const double* Q) { E->setProcessNoise(Map<const Eigen::Matrix<double, 12, 12>>(Q)); }
void ref_ekf_get_dydx(void* h, double* out) {
    MatrixXd f = E->getDydx();
    Map<Eigen::Matrix<double, 20, 20>> o(out);
    o = f.topLeftCorner<20, 20>();
}
void ref_ekf_initialize_orientation(void* h, const double* a) { E->initializeOrientation(Eigen::Vector3d(a[0], a[1], a[2])); }
void ref_ekf_predict(void* h, double t, const double* g, const double* a) {
    E->predict(t, Eigen::Vector3d(g[0], g[1], g[2]), Eigen::Vector3d(a[0], a[1], a[2]));
}
void ref_ekf_update_zupt(void* h, double r) { E->updateZupt(r); }
void ref_ekf_update_zupt_initialization(void* h) { E->updateZuptInitialization(); }
void ref_ekf_update_zrupt(void* h, const double* g) { E->updateZrupt(Eigen::Vector3d(g[0], g[1], g[2])); }
void ref_ekf_update_pseudo_velocity(void* h, double s, double r) { E->updatePseudoVelocity(s, r); }
void ref_ekf_update_position(void* h, const double* p, double r) { E->updatePosition(Eigen::Vector3d(p[0], p[1], p[2]), r); }
void ref_ekf_update_zero_height(void* h, double r) { E->updateZeroHeight(r); }
void ref_ekf_update_orientation(void* h, const double* q, double r) { E->updateOrientation(Eigen::Vector4d(q[0], q[1], q[2], q[3]), r); }
int ref_ekf_visual_check(void* h, const double* H, int n, int l, const double* f, const double* y, double r, double thr) {
    MatrixXd Hm = Map<const MatrixXd>(H, n, l); VectorXd fv = Map<const VectorXd>(f, n), yv = Map<const VectorXd>(y, n);
    return (int)E->visualTrackOutlierCheck(Hm, fv, yv, r, thr);
}
void ref_ekf_visual_update(void* h, const double* H, int n, int l, const double* f, const double* y, double r) {
    MatrixXd Hm = Map<const MatrixXd>(H, n, l); VectorXd fv = Map<const VectorXd>(f, n), yv = Map<const VectorXd>(y, n);
    E->updateVisualTrack(Hm, fv, yv, r);
}
void ref_ekf_augment(void* h, int d) { E->updateVisualPoseAugmentation(d); }
void ref_ekf_unaugment(void* h) { E->updateUndoAugmentation(); }
void ref_ekf_symmetrize(void* h) { E->maintainPositiveSemiDefinite(); }
void ref_ekf_normalize_quaternions(void* h, int only) { E->normalizeQuaternions(only != 0); }
void ref_ekf_translate_to(void* h, const double* p) { E->translateTo(Eigen::Vector3d(p[0], p[1], p[2])); }
void ref_ekf_transform_to(void* h, const double* p, const double* q, int i) {
    E->transformTo(Eigen::Vector3d(p[0], p[1], p[2]), Eigen::Vector4d(q[0], q[1], q[2], q[3]), i);
}
void ref_ekf_insert_map_point(void* h, int idx, const double* p) { E->insertMapPoint(idx, Eigen::Vector3d(p[0], p[1], p[2])); }
void ref_ekf_condition_on_last_pose(void* h) { E->conditionOnLastPose(); }
void ref_ekf_lock_biases(void* h) { E->lockBiases(); }
}
