// TEST INFRASTRUCTURE: back-end selection and lock-step comparison for the whole-core parity harness (run_pipeline).
#ifndef HV_PIPELINE_BACKENDS_HPP_
#define HV_PIPELINE_BACKENDS_HPP_
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace tracker { struct Image; }

namespace harness {
// Which implementation the reference's three factory symbols hand out (read at the moment of the call):
//   REF  = the reference's own classes (oracle/_ref/libref_backends.so: image_pyramid.cpp, optical_flow.cpp, ekf.cpp)
//   CUDA = hybvio_b200/host/cuda_tracker_backends.cpp, cuda_ekf.cpp over libhybvio_b200.so
//   DUAL = lock-step pair: every call goes to both with identical inputs, the results are compared, the REFERENCE result
//          is returned to the (unmodified) pipeline
enum class Flavour { REF, CUDA, DUAL };
void setFlavour(Flavour f);
Flavour flavour();
void setFrameIndex(int frame);
void setUseCudaDetector(bool on);      // CUDA / DUAL flavours: corner detector on the device too; default on
bool useCudaDetector();
void setRefThreads(int n);
int refThreads();

// lock-step tracker shadow: image of the CUDA-flavoured Image::Factory that belongs to a reference-flavoured image
void registerShadowImage(tracker::Image* ref, std::shared_ptr<tracker::Image> cuda);

struct Outlier { int frame, call, index; float dx, dy; };
struct OpStat { long calls = 0; double maxPos = 0, maxM = 0, maxPrel = 0; };
struct FrameTracks { int frame; std::vector<int> ids, status; std::vector<float> pts; bool keyframe; };

struct Stats {
    // pyramid (DUAL)
    long pyramidsCompared = 0, pyramidLevelsCompared = 0, pyramidMismatchBytes = 0;
    // LK (DUAL)
    long lkCalls = 0, lkPoints = 0, lkTracked = 0, lkStatusMismatch = 0, lkOver1e3 = 0;
    double lkMaxDiff = 0;
    std::vector<Outlier> lkOutliers;
    // EKF (DUAL)
    std::map<std::string, OpStat> ekfOps;
    long ekfChecks = 0, ekfCheckMismatch = 0, ekfCompares = 0;
    double ekfMaxPos = 0, ekfMaxM = 0, ekfMaxPrel = 0;
    std::vector<double> framePos, framePrel;          // running maximum within each frame
    // corner detector (DUAL): corner lists of FeatureDetector::detect, reference against CUDA
    long detCalls = 0, detCorners = 0, detMismatch = 0;
    int detFirstMismatchFrame = -1;
    // undistortion / rectification (DUAL): output images of Undistorter::undistort, reference against CUDA
    long undCalls = 0, undPixels = 0, undMismatch = 0, undUndefined = 0;
    // tracker (DUAL): reference-driven tracker against the CUDA-driven shadow tracker
    long trkFrames = 0, trkTracks = 0, trkIdMismatch = 0, trkStatusMismatch = 0, trkSizeMismatch = 0, trkKeyframeMismatch = 0;
    double trkMaxPointDiff = 0;
    int trkFirstMismatchFrame = -1;
    // per-flavour log of Tracker::Output (REF / CUDA free-running)
    std::vector<FrameTracks> trackLog[2];
};
Stats& stats();
} // namespace harness
#endif
