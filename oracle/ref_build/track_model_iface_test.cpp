// TEST INFRASTRUCTURE: drives the per-track measurement model through the reference's own interfaces, twice, on the same
// odometry::EKF (built by hybvio_b200/host/cuda_ekf.cpp):
//   reference path  extractCameraPoseTrail -> Triangulator::triangulate -> stereo sum -> prepareVisualUpdate -> EKF::
//                   visualTrackOutlierCheck / updateVisualTrack   (src/odometry/triangulation.cpp + backend.cpp:1050-1190, UNMODIFIED,
//                   H built on the host and uploaded)
//   device path     cudaTrackModels -> cudaVisualTrackOutlierCheck / cudaUpdateVisualTrack (hybvio_b200/host/cuda_track_model.hpp)
// and requires identical statuses and H, f, state within 1e-9. Built by build_ref_tests.sh, run on the GPU by
// tests/test_zz_gpu_track_model.py.
#include "triangulation.hpp"
#include "ekf.hpp"
#include "parameters.hpp"
#include "cuda_track_model.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace odometry;

static double urand() { return rand() / (double)RAND_MAX; }
static double nrand() { double s = 0; for (int i = 0; i < 12; i++) s += urand(); return s - 6.0; }

int main()
{
    int fails = 0;
    for (int stereo = 0; stereo < 2; stereo++) {
        srand(42 + stereo);
        Parameters params;
        params.odometry.cameraTrailLength = 20;
        params.tracker.useStereo = stereo != 0;
        Eigen::Matrix4d T1 = Eigen::Matrix4d::Identity();
        T1.topLeftCorner<3, 3>() = util::quat2rmat(Eigen::Vector4d(1, 0.01, -0.02, 0.015).normalized());
        T1.block<3, 1>(0, 3) = Eigen::Vector3d(0.01, -0.02, 0.005);
        Eigen::Matrix4d T2 = T1; T2(0, 3) -= 0.11;
        params.imuToCamera = T1; params.secondImuToCamera = T2;
        auto ekf = EKF::build(params);
        const int N = ekf->getStateDim(), trail = 20;
        Eigen::VectorXd m = Eigen::VectorXd::Zero(N);
        for (int k = 0; k <= trail; k++) {
            const Eigen::Vector3d pos(0.08 * k + 0.005 * nrand(), 0.02 * std::sin(0.7 * k) + 0.005 * nrand(), 0.01 * k + 0.005 * nrand());
            Eigen::Vector4d q(1.0, 0.01 * k + 0.0015 * nrand(), -0.0075 * k + 0.0015 * nrand(), 0.005 * std::sin((double)k)); q.normalize();
            const int o = k == 0 ? 0 : CAM + 7 * (k - 1);
            m.segment<3>(o) = pos;
            m.segment<4>(k == 0 ? ORI : o + 3) = q;
        }
        m.segment<3>(BAT) = Eigen::Vector3d::Ones();
        ekf->setState(m);
        // tracks: a point in front of the current camera, observed from npose poses
        const int ntracks = 24;
        std::vector<std::vector<int>> idx(ntracks);
        std::vector<vecVector2d> feats(ntracks), vels(ntracks);
        for (int t = 0; t < ntracks; t++) {
            const int npose = 2 + t % 9;
            std::vector<int> pool; for (int k = 1; k <= trail; k++) pool.push_back(k);
            for (int k = 0; k < npose - 1; k++) std::swap(pool[k], pool[k + rand() % (int)(pool.size() - k)]);
            idx[t].assign(pool.begin(), pool.begin() + npose - 1); std::sort(idx[t].begin(), idx[t].end()); idx[t].insert(idx[t].begin(), 0);
            CameraPoseTrail tr;
            extractCameraPoseTrail(*ekf, idx[t], params, stereo != 0, tr);
            const double depth = 2.0 + 3.0 * (t % 5);
            const Eigen::Vector3d pf = tr[0].p + tr[0].R.transpose() * Eigen::Vector3d((urand() * 0.6 - 0.3) * depth, (urand() * 0.4 - 0.2) * depth, depth);
            for (const auto& pose : tr) {
                const Eigen::Vector3d c = pose.R * (pf - pose.p);
                Eigen::Vector2d ip(c(0) / c(2) + 1e-3 * nrand(), c(1) / c(2) + 1e-3 * nrand());
                if (t % 6 == 4) ip = -ip;                                   // behind the cameras
                if (t % 6 == 5 && feats[t].size() == 1) ip += Eigen::Vector2d(0.4, -0.3);   // gross outlier
                feats[t].push_back(ip); vels[t].push_back(Eigen::Vector2d(0.05 * nrand(), 0.05 * nrand()));
            }
        }
        // device path: all tracks, one launch
        std::vector<CudaTrackIn> in(ntracks);
        for (int t = 0; t < ntracks; t++) in[t] = CudaTrackIn{&idx[t], &feats[t], &vels[t]};
        std::vector<CudaTrackOut> dev;
        cudaTrackModels(*ekf, params, in, dev);
        Triangulator triangulator(params.odometry);
        double worstH = 0, worstF = 0, worstPf = 0; int nOk = 0, nCheck = 0;
        for (int t = 0; t < ntracks; t++) {
            CameraPoseTrail tr;
            extractCameraPoseTrail(*ekf, idx[t], params, stereo != 0, tr);
            TriangulationArgsOut out;
            const TriangulationArgsIn args { .imageFeatures = feats[t], .featureVelocities = vels[t], .trail = tr, .stereo = stereo != 0,
                .calculateDerivatives = true, .estimateImuCameraTimeShift = params.odometry.estimateImuCameraTimeShift };
            TriangulatorStatus st = triangulator.triangulate(args, out);
            if (st != TriangulatorStatus::OK) { out.dpfdp.clear(); out.dpfdq.clear(); }
            if (stereo && st == TriangulatorStatus::OK) {
                const size_t n = idx[t].size();
                for (size_t i = 0; i < n; ++i) { out.dpfdp[i] += out.dpfdp[i + n]; out.dpfdq[i] += out.dpfdq[i + n]; }
                out.dpfdp.resize(n); out.dpfdq.resize(n);
            }
            bool ok = st == dev[t].triangulateStatus;
            if (ok && st == TriangulatorStatus::OK) {
                nOk++;
                Eigen::MatrixXd H, Hd; Eigen::VectorXd f, fd;
                const PrepareVisualUpdateArgsIn pargs { .triangulationOut = out, .featureVelocities = vels[t], .trail = tr, .poseTrailIndex = idx[t],
                    .stateDim = N, .useStereo = stereo != 0, .truncated = true, .mapPointOffset = -1,
                    .estimateImuCameraTimeShift = params.odometry.estimateImuCameraTimeShift };
                const PrepareVuStatus vu = prepareVisualUpdate(pargs, H, f);
                ok = vu == dev[t].prepareVuStatus && vu == PREPARE_VU_OK && H.rows() == dev[t].rows && H.cols() == dev[t].cols;
                if (ok) {
                    cudaTrackModelDownload(*ekf, dev[t], Hd, fd);
                    const double eh = (H - Hd).cwiseAbs().maxCoeff() / H.cwiseAbs().maxCoeff(), ef = (f - fd).cwiseAbs().maxCoeff();
                    const double ep = (out.pf - dev[t].pf).cwiseAbs().maxCoeff();
                    worstH = std::max(worstH, eh); worstF = std::max(worstF, ef); worstPf = std::max(worstPf, ep);
                    ok = eh < 1e-9 && ef < 1e-9 && ep < 1e-9;
                    // outlier check on the uploaded H vs on the device-resident H
                    Eigen::VectorXd y(2 * feats[t].size());
                    for (size_t i = 0; i < feats[t].size(); i++) y.segment<2>(2 * i) = feats[t][i];
                    const VuOutlierStatus a = ekf->visualTrackOutlierCheck(H, f, y, 0.02, -1.0);
                    const VuOutlierStatus b = cudaVisualTrackOutlierCheck(*ekf, dev[t], 0.02, -1.0);
                    ok = ok && a == b; nCheck++;
                    if (t == 0) {                                           // one update each way on two copies of the filter
                        auto e1 = ekf->clone(), e2 = ekf->clone();
                        e1->updateVisualTrack(H, f, y, 0.02);
                        std::vector<CudaTrackOut> dev2; cudaTrackModels(*e2, params, {in[t]}, dev2);
                        cudaUpdateVisualTrack(*e2, dev2[0], 0.02);
                        const double em = (e1->getState() - e2->getState()).cwiseAbs().maxCoeff();
                        const double eP = (e1->getStateCovariance() - e2->getStateCovariance()).cwiseAbs().maxCoeff() / e1->getStateCovariance().cwiseAbs().maxCoeff();
                        std::printf("  update through both paths: |dm| %.2e |dP|/|P| %.2e\n", em, eP);
                        ok = ok && em < 1e-9 && eP < 1e-9;
                    }
                }
            }
            if (!ok) { fails++; std::printf("track %d (stereo %d): reference status %d, device %d  FAIL\n", t, stereo, (int)st, (int)dev[t].triangulateStatus); }
        }
        std::printf("stereo %d: %d tracks, %d triangulated, %d checks; worst |dH|/|H| %.2e |df| %.2e |dpf| %.2e\n", stereo, ntracks, nOk, nCheck, worstH, worstF, worstPf);
        // ---- the whole per-track loop (backend.cpp:1012-1252, per-track mode): reference calls track by track on one copy of the
        // filter, cudaVisualTracks (one device-gated chain) on another
        {
            const double chiR = 0.05, visR = 0.05; const int maxSucc = 5;
            auto e1 = ekf->clone(), e2 = ekf->clone();
            std::vector<int> rTri(ntracks, -1), rOut(ntracks, (int)VuOutlierStatus::NOT_COMPUTED), rUpd(ntracks, 0);
            int succ = 0;
            for (int t = 0; t < ntracks && succ < maxSucc; t++) {
                CameraPoseTrail tr;
                extractCameraPoseTrail(*e1, idx[t], params, stereo != 0, tr);
                TriangulationArgsOut out;
                const TriangulationArgsIn args { .imageFeatures = feats[t], .featureVelocities = vels[t], .trail = tr, .stereo = stereo != 0,
                    .calculateDerivatives = true, .estimateImuCameraTimeShift = params.odometry.estimateImuCameraTimeShift };
                TriangulatorStatus st = triangulator.triangulate(args, out);
                rTri[t] = (int)st;
                if (st != TriangulatorStatus::OK) continue;
                if (stereo) {
                    const size_t n = idx[t].size();
                    for (size_t i = 0; i < n; ++i) { out.dpfdp[i] += out.dpfdp[i + n]; out.dpfdq[i] += out.dpfdq[i + n]; }
                    out.dpfdp.resize(n); out.dpfdq.resize(n);
                }
                Eigen::MatrixXd H; Eigen::VectorXd f, y(2 * feats[t].size());
                const PrepareVisualUpdateArgsIn pargs { .triangulationOut = out, .featureVelocities = vels[t], .trail = tr, .poseTrailIndex = idx[t],
                    .stateDim = N, .useStereo = stereo != 0, .truncated = true, .mapPointOffset = -1,
                    .estimateImuCameraTimeShift = params.odometry.estimateImuCameraTimeShift };
                if (prepareVisualUpdate(pargs, H, f) != PREPARE_VU_OK) continue;
                for (size_t i = 0; i < feats[t].size(); i++) y.segment<2>(2 * i) = feats[t][i];
                const VuOutlierStatus os = e1->visualTrackOutlierCheck(H, f, y, chiR, -1.0);
                rOut[t] = (int)os;
                if (os == VuOutlierStatus::INLIER) { e1->updateVisualTrack(H, f, y, visR); rUpd[t] = 1; succ++; }
            }
            std::vector<CudaTrackResult> res;
            const int dsucc = cudaVisualTracks(*e2, params, in, chiR, -1.0, visR, maxSucc, 0, res);
            bool ok = dsucc == succ;
            for (int t = 0; t < ntracks; t++) {
                const int dTri = res[t].attempted ? (int)res[t].triangulateStatus : -1;
                if (dTri != rTri[t] || (int)res[t].outlierStatus != rOut[t] || (int)res[t].updated != rUpd[t]) {
                    ok = false;
                    std::printf("  loop track %d: reference %d/%d/%d, chain %d/%d/%d  FAIL\n", t, rTri[t], rOut[t], rUpd[t], dTri, (int)res[t].outlierStatus, (int)res[t].updated);
                }
            }
            const double em = (e1->getState() - e2->getState()).cwiseAbs().maxCoeff();
            const double eP = (e1->getStateCovariance() - e2->getStateCovariance()).cwiseAbs().maxCoeff() / e1->getStateCovariance().cwiseAbs().maxCoeff();
            // Five chained updates on the freshly built filter (prior variances noiseInitial^2 x noiseScale 1e4 against r = 0.05) are
            // badly conditioned: rounding-level differences in H (1e-15) grow to 1e-8 in m -- also between two runs of the
            // reference's own EKF fed with the two H (measured here on the CPU). Hence 1e-5 (the north-star gate is 1e-4 m); the
            // 1e-9 bar is held by the single update above and, with a well-conditioned prior, by tests/test_zz_gpu_track_model.py.
            ok = ok && em < 1e-5 && eP < 1e-6;
            std::printf("  per-track loop vs device-gated chain: %d / %d updates, |dm| %.2e |dP|/|P| %.2e  %s\n", succ, dsucc, em, eP, ok ? "ok" : "FAIL");
            fails += !ok;
        }
    }
    std::printf(fails ? "FAILED (%d)\n" : "track model interface: all ok\n", fails);
    return fails;
}
