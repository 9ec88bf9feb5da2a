// hybvio_b200/host/cuda_track_model.hpp -- the per-track measurement model of Session::trackerVisualUpdate
// (src/odometry/backend.cpp:1050-1160: extractCameraPoseTrail -> Triangulator::triangulate -> per-pose stereo sum ->
// prepareVisualUpdate(truncated)) evaluated on the device for an odometry::EKF built by hybvio_b200/host/cuda_ekf.cpp.
// H, f and the measurement vector stay in HBM; the outlier check / update read them there.
// Uses the reference's own types (TriangulatorStatus, PrepareVuStatus, VuOutlierStatus, vecVector2d).
#pragma once
#include "ekf.hpp"
#include "output.hpp"
#include "parameters.hpp"
#include "util.hpp"

#include <vector>

namespace odometry {

struct CudaTrackIn {
    const std::vector<int>* poseTrailIndex;       // as EkfStateIndex::createTrackIndex fills it (backend.cpp:1040-1048)
    const vecVector2d* imageFeatures;             // TriangulationArgsIn::imageFeatures (first camera, then second)
    const vecVector2d* featureVelocities;         // TriangulationArgsIn::featureVelocities
};

struct CudaTrackOut {
    TriangulatorStatus triangulateStatus;         // incl. the BAD_DEPTH gate of backend.cpp:1095-1098
    PrepareVuStatus prepareVuStatus;              // meaningful if triangulateStatus == OK
    Eigen::Vector3d pf;                           // TriangulationArgsOut::pf
    double depth;
    int rows, cols;                               // of H
    const double *dH, *df, *dy;                   // device pointers, valid until the next cudaTrackModels call on this EKF
    int index;                                    // position in the batch
};

/// All tracks against the CURRENT state of `ekf`, one kernel launch. `ekf` must come from the CUDA EKF::build.
void cudaTrackModels(EKF& ekf, const Parameters& parameters, const std::vector<CudaTrackIn>& in, std::vector<CudaTrackOut>& out);
/// EKF::visualTrackOutlierCheck on the device-resident H of a track (ekf.cpp:787-819)
VuOutlierStatus cudaVisualTrackOutlierCheck(EKF& ekf, const CudaTrackOut& track, double r, double trackRmseThreshold);
/// EKF::updateVisualTrack on the device-resident H of a track (ekf.cpp:829-844)
void cudaUpdateVisualTrack(EKF& ekf, const CudaTrackOut& track, double r);
/// The whole per-track loop (model -> visualTrackOutlierCheck -> updateVisualTrack if INLIER, at most maxSuccessfulUpdates
/// updates; backend.cpp:1012-1252 in per-track mode) as one chain with the control flow on the device: one host
/// synchronisation per `lookahead` tracks (0: one in total) instead of two per track.
struct CudaTrackResult {
    TriangulatorStatus triangulateStatus; bool attempted;   // attempted == false: the chain already had its successful updates
    PrepareVuStatus prepareVuStatus;
    VuOutlierStatus outlierStatus;
    bool updated;
    double chi2, depth;
    Eigen::Vector3d pf;
};
int cudaVisualTracks(EKF& ekf, const Parameters& parameters, const std::vector<CudaTrackIn>& in, double chiOutlierR, double rmseThreshold,
                     double visualR, int maxSuccessfulUpdates, int lookahead, std::vector<CudaTrackResult>& out);
/// Host copies of H and f (viewers, tests)
void cudaTrackModelDownload(EKF& ekf, const CudaTrackOut& track, Eigen::MatrixXd& H, Eigen::VectorXd& f);

}  // namespace odometry
