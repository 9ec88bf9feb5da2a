/*
 * hybvio_b200.h -- C ABI of libhybvio_b200.so: the B200 (sm_100a) implementation of HybVIO's per-frame hot path.
 *
 * Boundary. Each entry point names the reference interface it replaces (paths relative to the
 * reference root; OCV = 3rdparty/mobile-cv-suite/opencv/modules). The reference-side C++ adapter classes
 * (CudaImagePyramid / CudaOpticalFlow / CudaEKF, hybvio_b200/host/) sit on top of exactly these calls.
 *
 * Conventions: every function returns HV_OK (0) or a negative hv_status; nothing throws across the boundary;
 * handles are opaque; the caller owns host buffers, the library owns device buffers; matrices are fp64
 * COLUMN-MAJOR with leading dimension = rows (Eigen's default layout). All work of a context is issued on one
 * CUDA stream; calls taking host output buffers synchronise that stream before returning, `_device`/`_async`
 * variants do not. A context must be used from one thread at a time (the reference drives Tracker and EKF from
 * a single thread, src/api/api.cpp:425-428).
 *
 * There is NO CPU fallback: hv_ctx_create fails with HV_ERR_NO_DEVICE when no sm_100 GPU is present.
 */
#ifndef HYBVIO_B200_H_
#define HYBVIO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hv_status {
    HV_OK = 0,
    HV_ERR_INVALID = -1,      /* bad argument (NULL handle, n < 0, unsupported window size ...) */
    HV_ERR_NO_DEVICE = -2,    /* no CUDA device / not sm_100 */
    HV_ERR_CUDA = -3,         /* a CUDA runtime call failed; see hv_last_error */
    HV_ERR_OOM = -4,
    HV_ERR_UNSUPPORTED = -5,  /* e.g. pyrLKWindowSize not in {11,15,21,31}, maxLevel > 5 */
    HV_ERR_STATE = -6         /* call order violated (e.g. unaugment with no augmented pose) */
} hv_status;

typedef struct hv_ctx hv_ctx;
typedef struct hv_pyr hv_pyr;
typedef struct hv_ekf hv_ekf;

/* ---------------------------------------------------------------- context ------------------------------- */
const char* hv_version(void);
/* Last error text of this thread (valid until the next failing call). */
const char* hv_last_error(void);
int hv_device_count(void);
/* Creates a context on `device` with its own non-blocking stream. */
int hv_ctx_create(int device, hv_ctx** out);
/* Same, but issues all work on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream). */
int hv_ctx_create_on_stream(int device, void* cuda_stream, hv_ctx** out);
int hv_ctx_destroy(hv_ctx* ctx);
/* Waits for everything issued through the context: its stream and the library's side stream (hv_ekf_run_device_results). */
int hv_ctx_sync(hv_ctx* ctx);
/* The stream as a cudaStream_t (for event timing by the caller). */
void* hv_ctx_stream(hv_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long hv_ctx_launch_count(hv_ctx* ctx);

/* ---------------------------------------------------------------- image pyramid ------------------------- */
/* Replaces tracker::ImagePyramid + CpuImagePyramidFactory (src/tracker/image_pyramid.hpp:18-42,
 * image_pyramid.cpp:28-48): one object per camera frame, recycled through a pool like util::Allocator.
 * win = tracker.pyrLKWindowSize, max_level = tracker.pyrLKMaxLevel. Level sizes follow
 * cv::buildOpticalFlowPyramid (OCV/video/src/lkpyramid.cpp:726-822): ((w+1)/2, (h+1)/2), stopping early when
 * the next level would be <= win. */
int hv_pyr_create(hv_ctx* ctx, int width, int height, int win, int max_level, hv_pyr** out);
int hv_pyr_release(hv_pyr* pyr);
int hv_pyr_levels(const hv_pyr* pyr);
int hv_pyr_level_size(const hv_pyr* pyr, int level, int* width, int* height);
/* Replaces ImagePyramid::Factory::compute -> cv::buildOpticalFlowPyramid. `gray` is a HOST 8-bit image
 * (accelerated::Image CPU storage, row stride in bytes). Asynchronous: H2D copy + one fused kernel on the
 * context stream; the host buffer must stay valid until the next synchronising call (use pinned memory for
 * a truly asynchronous copy). */
int hv_pyr_build(hv_pyr* pyr, const uint8_t* gray, size_t stride_bytes);
/* Same for `n` images in ONE kernel launch (stereo pair: n = 2). src_is_device != 0: `gray[i]` are device
 * pointers (frame already in HBM). All pyramids must belong to one context and have equal level-0 size. */
int hv_pyr_build_batch(hv_pyr* const* pyrs, const uint8_t* const* gray, const size_t* stride_bytes, int n,
                       int src_is_device);
/* Test/debug accessors, replacing ImagePyramid::getGrayLevel/getGradientLevel/getOpenCv: copy one level to the
 * host. gray: w*h u8; deriv: w*h*2 int16 interleaved (Ix,Iy), GRADIENT_SCALE_0_255 = 1/32
 * (image_pyramid.hpp:19-26). Either pointer may be NULL. Synchronises. */
int hv_pyr_download_level(hv_pyr* pyr, int level, uint8_t* gray, int16_t* deriv);
/* Same, laid out exactly like the reference's cv::Mat level *with* its win-pixel padding (gray REFLECT_101,
 * deriv CONSTANT 0; lkpyramid.cpp:761-808): (w+2win) x (h+2win). The device stores levels unpadded; the
 * border is materialised on the host by this accessor only. */
int hv_pyr_download_level_padded(hv_pyr* pyr, int level, uint8_t* gray, int16_t* deriv);

/* ---------------------------------------------------------------- Lucas-Kanade -------------------------- */
/* Replaces tracker::OpticalFlow::compute -> cv::calcOpticalFlowPyrLK (src/tracker/optical_flow.cpp:10-59,
 * 78-102; OCV/video/src/lkpyramid.cpp:1236-1401, 183-724) with TermCriteria(COUNT|EPS, max_iter, eps),
 * flags = use_initial ? OPTFLOW_USE_INITIAL_FLOW : 0, minEigThreshold = min_eig, `err` requested.
 *   prev_xy   n x (x,y) float32, host
 *   next_xy   n x (x,y) float32, host; in: initial guesses when use_initial, out: tracked end points
 *   status    n x uint8, OpenCV status (1 tracked, 0 failed); may be NULL
 *   track_status  n x int32 tracker::Feature::Status (src/tracker/track.hpp:9-20): TRACKED=0, FAILED_FLOW=2,
 *             FLOW_OUT_OF_RANGE=4, exactly as optical_flow.cpp:52-58 derives it; may be NULL
 * Synchronises. */
int hv_lk_track(hv_ctx* ctx, hv_pyr* prev, hv_pyr* next, const float* prev_xy, float* next_xy, uint8_t* status,
                int32_t* track_status, int n, int use_initial, int max_iter, double eps, double min_eig);
/* Device-resident variant: all pointers are device pointers, nothing is copied, no synchronisation. */
int hv_lk_track_device(hv_ctx* ctx, hv_pyr* prev, hv_pyr* next, const float* d_prev_xy, float* d_next_xy,
                       uint8_t* d_status, int32_t* d_track_status, int n, int use_initial, int max_iter,
                       double eps, double min_eig);
/* The same launch on a stream of the CALLER instead of the context's own (cuda_stream: a cudaStream_t of the same device). For pipelines
 * that keep the tracker and the filter on ONE stream, so that the optical flow of a frame, its visual updates and the next propagation
 * follow each other without cross-stream events, while the pyramid builds keep the context's stream. The caller orders the call
 * behind the builds of both pyramids (an event recorded on hv_ctx_stream(ctx)) and the next build of either pyramid behind this call.
 * d_init_xy (n x 2, may be NULL = start at d_prev_xy): the predicted end points (OPTFLOW_USE_INITIAL_FLOW) are READ from there and the
 * results written to d_next_xy -- a predictor's output buffer is used as it is, no copy into the result buffer first. */
int hv_lk_track_device_on_stream(hv_ctx* ctx, void* cuda_stream, hv_pyr* prev, hv_pyr* next, const float* d_prev_xy, const float* d_init_xy,
                                 float* d_next_xy, uint8_t* d_status, int32_t* d_track_status, int n, int max_iter, double eps, double min_eig);
/* Several independent LK calls (e.g. one per stream) in one launch; device pointers. */
typedef struct hv_lk_job {
    hv_pyr* prev; hv_pyr* next;
    const float* d_prev_xy; float* d_next_xy; uint8_t* d_status; int32_t* d_track_status;
    int n; int use_initial;
} hv_lk_job;
int hv_lk_track_batch_device(hv_ctx* ctx, const hv_lk_job* jobs, int njobs, int max_iter, double eps, double min_eig);

/* ---------------------------------------------------------------- corner detection ---- */
/* Device part of tracker::FeatureDetector::detect on CPU images (src/tracker/feature_detector.cpp:566-682, the default "GPU-GFTT"
 * detector without OpenGL images): CpuCornerResponse = cv::cornerMinEigenVal(gray, block_size, 3) (feature_detector.cpp:281-310,
 * OCV/imgproc/src/corner.cpp:238-320) and CollectMax::cpuImplementation (feature_detector.cpp:393-417: the best response of every
 * cell x cell block, x GAIN 16, kept if > min_response), on the level-0 gray image of `pyr`, which hv_pyr_build has already put into
 * HBM. cell = the detector's block size (32 for tracker.gfttMinDistance >= 32, feature_detector.cpp:425-433), block_size =
 * tracker.gfttBlockSize (3), min_response = tracker.gfttMinResponse.
 *   kp    (w / cell) * (h / cell) x (x, y, response) float32 in row-major cell order; a cell without a qualifying pixel reports
 *         (0, 0, -1e10) exactly as the reference does. The sort by response, the reference's resize quirk and applyMinDistance
 *         (feature_detector.cpp:625-638) stay on the host: hybvio_b200/host/cuda_feature_detector.cpp.
 * hv_gftt_detect: host output, synchronises. hv_gftt_detect_device: device output, no synchronisation. */
int hv_gftt_cells(const hv_pyr* pyr, int cell, int* cells_x, int* cells_y);
int hv_gftt_detect(hv_ctx* ctx, hv_pyr* pyr, int block_size, int cell, float min_response, float* kp);
int hv_gftt_detect_device(hv_ctx* ctx, hv_pyr* pyr, int block_size, int cell, float min_response, float* d_kp);

/* ---------------------------------------------------------------- frame ingest -------- */
/* Device part of tracker::Image::Factory::build / buildStereo (src/tracker/image.cpp:243-308): colour -> gray
 * (accelerated-arrays pixelwiseAffine, image.cpp:360-366) and undistortion / rectification (UndistorterImplementation::undistort,
 * src/tracker/undistorter.cpp:77-118), with the result written straight into level 0 of the frame's pyramid, followed by the fused
 * pyramid kernel -- the frame crosses PCIe once, as it arrives.
 *   hv_remap_entry: per OUTPUT pixel the source position the reference's cameras map it to, as undistorter.cpp:93-94 forms it:
 *   x0 = floor(px), y0 = floor(py), xfrac = (float)(px - x0), yfrac = (float)(py - y0); x0 = HV_REMAP_INVALID_X0 where pixelToRay /
 *   rayToPixel fail or (px, py) lies outside [0, w) x [0, h). The table belongs to a (rectified camera, original camera) pair and is
 *   computed once by the adapter with the reference's own Camera classes (hybvio_b200/host/cuda_undistorter.cpp).
 * hv_ingest_frame: src = HOST frame (w x h, `channels` interleaved 8-bit channels, row stride in bytes). channels > 1: gray =
 * sum_j coeff[j] * channel_j in the reference's fixed-point arithmetic (coeff NULL: 0.299, 0.587, 0.114, 0). table (device handle from
 * hv_ingest_set_remap) optional. dst: the pyramid to build from the ingested frame. gray_out (host, optional, w x h, tightly packed)
 * receives the ingested gray image (the tracker::Image keeps it on the host for cornerSubPix / SLAM); asynchronous like hv_pyr_build:
 * synchronise (hv_ctx_sync or any synchronising call) before reading it. */
#define HV_REMAP_INVALID_X0 (-32768)
typedef struct hv_remap_entry { int16_t x0, y0; float xfrac, yfrac; } hv_remap_entry;
typedef struct hv_ingest hv_ingest;
int hv_ingest_create(hv_ctx* ctx, int width, int height, hv_ingest** out);
int hv_ingest_destroy(hv_ingest* ing);
int hv_ingest_set_remap(hv_ingest* ing, const hv_remap_entry* table);     /* width * height entries, host; NULL removes the table */
int hv_ingest_frame(hv_ingest* ing, const uint8_t* src, size_t stride_bytes, int channels, const double* coeff, hv_pyr* dst, uint8_t* gray_out);

/* ---------------------------------------------------------------- EKF ----------------------------------- */
/* Replaces odometry::EKF / EKFImplementation (src/odometry/ekf.hpp:62-174, ekf.cpp). State m (N) and
 * covariance P (N x N) are fp64 and live in HBM; N = 20 + 7*trail + 3*map (ekf.cpp:156-158). */
typedef struct hv_ekf_params {          /* the odometry::Parameters fields EKFImplementation reads */
    int camera_trail_length;            /* odometry.cameraTrailLength */
    int hybrid_map_size;                /* odometry.hybridMapSize */
    double noise_scale;                 /* odometry.noiseScale (the EKF uses its square, ekf.cpp:154) */
    double gravity;                     /* odometry.gravity */
    double noise_initial_pos, noise_initial_vel, noise_initial_ori;
    double noise_initial_bga, noise_initial_baa, noise_initial_bat, noise_initial_sft;
    double noise_initial_pos_trail, noise_initial_ori_trail;
    double noise_process_acc, noise_process_gyro;
    double noise_process_baa, noise_process_baa_rev, noise_process_bga, noise_process_bga_rev;
    double augment_r, init_zupt_r, rotation_zupt_r;
} hv_ekf_params;
/* Fills `p` with the defaults of codegen/parameter_definitions.c. */
void hv_ekf_default_params(hv_ekf_params* p);

/* EKF::build (ekf.cpp:153-296, 1087-1092) */
int hv_ekf_create(hv_ctx* ctx, const hv_ekf_params* params, hv_ekf** out);
int hv_ekf_destroy(hv_ekf* ekf);
/* EKF::clone (ekf.cpp:1074-1079): device-to-device copy of m, P, Q and the host-side time bookkeeping */
int hv_ekf_clone(const hv_ekf* src, hv_ekf** out);
int hv_ekf_state_dim(const hv_ekf* ekf);                       /* getStateDim */
int hv_ekf_pose_count(const hv_ekf* ekf);                      /* getPoseCount = augmentCount + 1 */
double hv_ekf_platform_time(const hv_ekf* ekf);                /* getPlatformTime */
double hv_ekf_history_time(const hv_ekf* ekf, int i);          /* historyTime (ekf.cpp:556-562) */
int hv_ekf_was_stationary(const hv_ekf* ekf);                  /* getWasStationary */
int hv_ekf_set_first_sample_time(hv_ekf* ekf, double t);       /* setFirstSampleTime (ekf.cpp:1035-1041) */

/* setState / setStateCovariance / getState / getStateCovariance (ekf.cpp:950-981). NULL = skip. download
 * synchronises. */
int hv_ekf_upload(hv_ekf* ekf, const double* m, const double* P);
int hv_ekf_download(hv_ekf* ekf, double* m, double* P);
/* getInertialState / setInertialState (ekf.cpp:679-690): first 20 entries / top-left 20x20 block. */
int hv_ekf_download_inertial(hv_ekf* ekf, double* m20, double* P20x20);
int hv_ekf_set_inertial_state(hv_ekf* ekf, const double* m20, const double* P20x20);
int hv_ekf_set_process_noise(hv_ekf* ekf, const double* Q12x12);   /* setProcessNoise */
int hv_ekf_get_dydx(hv_ekf* ekf, double* dydx20x20);               /* getDydx's non-identity block (tests) */

int hv_ekf_initialize_orientation(hv_ekf* ekf, const double acc[3]);          /* ekf.cpp:299-317 */
/* predict (ekf.cpp:320-514): mean propagation, Jacobians and the structured covariance update all run on the
 * device; the host only keeps the sample-time bookkeeping (first sample / dt <= 0 are no-ops). Asynchronous. */
int hv_ekf_predict(hv_ekf* ekf, double t, const double gyro[3], const double acc[3]);
/* IMU samples are DEFERRED: consecutive predict() calls -- each optionally followed by normalize_quaternions(ekf, 1), as in
 * the reference's sample loop (src/odometry/backend.cpp:734-735) -- are queued on the host (up to max_samples, default and
 * maximum 16) and issued as ONE launch by the next call that needs the state, or by hv_ekf_flush. Results are identical
 * to issuing every sample on its own (max_samples = 1). Likewise hv_ekf_symmetrize directly followed by hv_ekf_augment
 * (backend.cpp:1267 -> 805) is one launch. */
int hv_ekf_flush(hv_ekf* ekf);
/* The 20 inertial states (position, velocity, orientation, biases: ekf.hpp:26-50) that the QUEUED IMU samples lead to, written to
 * d_mean20 (device, 20 doubles) by a small launch of its own on the context's stream -- the mean part of predict() alone, a quarter of
 * the full launch, which stays queued. For consumers that need the propagated pose but not the covariance: the optical-flow predictor
 * (src/odometry/backend.cpp:547-600 reads ekf->position() / orientation() and the pose trail, which predict() does not change), so that
 * the tracker can start while the covariance is still being propagated. Bit-identical to what the full launch leaves in the state. */
int hv_ekf_predicted_mean_device(hv_ekf* ekf, double* d_mean20);
/* The same into host memory (launch, 160-byte read-back, one synchronisation): what odometry::EKF::position() / orientation() / velocity()
 * cost behind a burst of predict() calls -- the flow predictor's reads -- instead of the full launch plus a read-back of the whole mean. */
int hv_ekf_predicted_mean(hv_ekf* ekf, double* mean20);
/* After this call the queued FULL launch (hv_ekf_flush, or whatever issues the queue next) goes to a stream of the library instead of the
 * context's stream, and the next call that touches the filter waits for it there: work the caller puts on the context's stream in between
 * without touching the filter -- the optical flow of a pipeline that keeps tracker and filter on one stream -- does not queue behind the
 * covariance propagation. (Not in throughput mode, HV_EKF_NO_PDL=1.) */
int hv_ekf_set_imu_batching(hv_ekf* ekf, int max_samples);

/* The fixed-H updates (ekf.cpp:573-677); rate limits and early-outs as in the reference. Asynchronous. */
int hv_ekf_update_zupt(hv_ekf* ekf, double r);
int hv_ekf_update_zupt_initialization(hv_ekf* ekf);
int hv_ekf_update_zrupt(hv_ekf* ekf, const double gyro[3]);
int hv_ekf_update_pseudo_velocity(hv_ekf* ekf, double default_speed, double r);
int hv_ekf_update_position(hv_ekf* ekf, const double pos[3], double r);
int hv_ekf_update_zero_height(hv_ekf* ekf, double r);
int hv_ekf_update_orientation(hv_ekf* ekf, const double q[4], double r);

/* visualTrackOutlierCheck (ekf.cpp:787-819). H is n x l column-major (host), f and y length n (host).
 * *vu_status: odometry::VuOutlierStatus (ekf.hpp:54-59) INLIER=0, NOT_COMPUTED=1, RMSE=2, CHI2=3.
 * *chi2 (optional) receives noiseScale * v' S^-1 v. Synchronises (the caller branches on the result,
 * src/odometry/backend.cpp:1158-1161). */
int hv_ekf_visual_check(hv_ekf* ekf, const double* H, int n, int l, const double* f, const double* y, double r,
                        double track_rmse_threshold, int* vu_status, double* chi2);
/* updateVisualTrack (ekf.cpp:829-844). Asynchronous. */
int hv_ekf_visual_update(hv_ekf* ekf, const double* H, int n, int l, const double* f, const double* y, double r);
/* Fused check + conditional update in one kernel and ONE host round trip: applies updateVisualTrack iff the
 * check returns INLIER, and (if m_out != NULL) returns the updated state mean, which the caller needs to
 * build the next track's H (backend.cpp:1054-1056). Equivalent to check followed by update. Synchronises. */
int hv_ekf_visual_check_update(hv_ekf* ekf, const double* H, int n, int l, const double* f, const double* y,
                               double r, double track_rmse_threshold, int* vu_status, double* chi2, double* m_out);
/* Device-resident variant of the two calls above (H, f, y already in HBM; result words written to
 * d_result[0] = status, d_result[1] = chi2 as doubles); mode: 0 = check only, 1 = update only,
 * 2 = check then update-if-inlier. Asynchronous. */
int hv_ekf_visual_device(hv_ekf* ekf, const double* d_H, int n, int l, const double* d_f, const double* d_y,
                         double r, double track_rmse_threshold, int mode, double* d_result);

/* Batch submission: executes `nops` EKF calls in order with ONE crossing of the language boundary (what
 * Session::process issues per frame: the IMU predicts, the per-track checks/updates, symmetrise, augment;
 * src/odometry/backend.cpp:716-867). Each op is exactly the single call of the same name. */
typedef enum hv_ekf_op_kind {
    HV_EKF_OP_PREDICT = 0,        /* t, gyro, acc */
    HV_EKF_OP_VISUAL = 1,         /* H, n, l, f, y, r, rmse_thr, mode (0 check, 1 update, 2 check+update-if-inlier) */
    HV_EKF_OP_SYMMETRIZE = 2,
    HV_EKF_OP_AUGMENT = 3,        /* index = discarded pose */
    HV_EKF_OP_UNAUGMENT = 4,
    HV_EKF_OP_NORMALIZE = 5       /* index = only_current */
} hv_ekf_op_kind;
typedef struct hv_ekf_op {
    int kind;
    int n, l, mode, index;
    double t, r, rmse_thr;
    double gyro[3], acc[3];
    const double* H; const double* f; const double* y;
} hv_ekf_op;
/* H/f/y are DEVICE pointers; fully asynchronous. Consecutive independent outlier checks (mode 0) are issued as one launch, one
 * cluster per measurement. The measurement inputs of a list are PREPARED inputs: the kernels read them while their predecessor on the
 * stream may still be running (programmatic dependent launch), so they must not be produced by work queued on this stream after that
 * predecessor. For H produced by the caller's own kernel right before the call use hv_ekf_visual_device, which reads its inputs only
 * after the dependency on all earlier work of the stream has been resolved. The inputs must stay valid until hv_ctx_sync or
 * hv_ekf_run_device_results has returned (synchronising the context's STREAM alone is not enough, see below). */
int hv_ekf_run_device(hv_ekf* ekf, const hv_ekf_op* ops, int nops);
/* What the VISUAL ops of the most recent hv_ekf_run_device list decided: VuOutlierStatus / chi2 into vu_status[i] / chi2[i] for the ops
 * [0, nops) of that list (entries of other ops untouched; lists of up to 256 ops report). The list itself returns nothing and does not
 * wait: a run of outlier checks that is followed by the pose augmentation (the end of a frame, backend.cpp:1012-1270) is issued on a side
 * stream of the library -- the checks only read the state and the augmentation writes second buffers that are swapped in -- so that
 * the augmentation, the next IMU burst and the next visual updates do not queue behind them. This call waits for all of it. */
int hv_ekf_run_device_results(hv_ekf* ekf, int nops, int* vu_status, double* chi2);
/* H/f/y are HOST pointers; every VISUAL op with mode 0 or 2 returns its VuOutlierStatus / chi2 into
 * vu_status[i] / chi2[i] (arrays of length nops, entries of other ops untouched) -- i.e. each such op is a host
 * round trip, as in the reference interface. m_out (optional, N doubles) receives the final state mean. */
int hv_ekf_run_host(hv_ekf* ekf, const hv_ekf_op* ops, int nops, int* vu_status, double* chi2, double* m_out);

int hv_ekf_augment(hv_ekf* ekf, int discarded_pose_index);     /* updateVisualPoseAugmentation (ekf.cpp:848-885) */
int hv_ekf_unaugment(hv_ekf* ekf);                             /* updateUndoAugmentation (ekf.cpp:888-903) */
int hv_ekf_symmetrize(hv_ekf* ekf);                            /* maintainPositiveSemiDefinite (ekf.cpp:1059-1067) */
int hv_ekf_normalize_quaternions(hv_ekf* ekf, int only_current);   /* ekf.cpp:1024-1032 */
int hv_ekf_translate_to(hv_ekf* ekf, const double pos[3]);     /* ekf.cpp:696-702 */
int hv_ekf_transform_to(hv_ekf* ekf, const double pos[3], const double q[4], int pose_index); /* ekf.cpp:704-758 */
int hv_ekf_insert_map_point(hv_ekf* ekf, int idx, const double pf[3]);   /* ekf.cpp:911-921 */
int hv_ekf_condition_on_last_pose(hv_ekf* ekf);                /* ekf.cpp:928-942 */
int hv_ekf_lock_biases(hv_ekf* ekf);                           /* ekf.cpp:944-947 */
/* ---- Per-track measurement model on the device (next hot-path row) --------------------------------
 * What Session::trackerVisualUpdate computes on the host for every track before it can call the EKF
 * (src/odometry/backend.cpp:1050-1160): extractCameraPoseTrail (src/odometry/triangulation.cpp:65-103) ->
 * Triangulator::triangulate with derivatives (:120-407, two-view start :610-710) -> per-pose sum of the two cameras
 * (backend.cpp:1105-1116) -> prepareVisualUpdate(truncated) (:897-987). Here it runs on the device from the resident state
 * mean; H, f (and the measurement vector y = the observations) stay in HBM, where hv_ekf_visual_device / hv_ekf_run_device
 * read them, so H is neither built on nor uploaded from the host. */
typedef struct hv_camera_model {
    double imu_to_camera[16];            /* odometry::Parameters::imuToCamera, 4x4 column-major (Eigen) */
    double second_imu_to_camera[16];     /* ...::secondImuToCamera (ignored unless use_stereo) */
    int use_stereo;                      /* tracker.useStereo */
    int estimate_imu_camera_time_shift;  /* odometry.estimateImuCameraTimeShift */
    unsigned gauss_newton_iterations;    /* odometry.triangulationGaussNewtonIterations (10) */
    double convergence_threshold;        /* odometry.triangulationConvergenceThreshold (1e-2) */
    double convergence_r;                /* odometry.triangulationConvergenceR (11) */
    double rcond_threshold;              /* odometry.triangulationRcondThreshold (1e-8) */
    double min_dist, max_dist;           /* odometry.triangulationMinDist / MaxDist (0, 1e300): BAD_DEPTH gate, backend.cpp:1095-1098 */
} hv_camera_model;
void hv_camera_model_defaults(hv_camera_model* c);     /* the reference defaults above; the two matrices are zeroed */
int hv_ekf_set_camera_model(hv_ekf* ekf, const hv_camera_model* c);
#define HV_TRACK_MAX_POSES 21            /* pose-trail indices per track (cameraTrailLength 20 + the current pose) */
typedef struct hv_track_obs {
    int npose;                           /* 2 .. HV_TRACK_MAX_POSES */
    const int* pose_trail_index;         /* npose indices as EkfStateIndex::createTrackIndex gives them: 0 = current pose, k = trail slot k - 1 */
    const double* ip;                    /* normalised image points x,y: npose of the first camera [, npose of the second] */
    const double* velocities;            /* their velocities, same layout (TriangulationArgsIn::featureVelocities) */
} hv_track_obs;
typedef struct hv_track_model {
    int triangulator_status;             /* odometry::TriangulatorStatus (output.hpp:21-29): OK 0, BEHIND 2, BAD_COND 3, NO_CONVERGENCE 4, BAD_DEPTH 5, UNKNOWN_PROBLEM 6 */
    int prepare_vu_status;               /* odometry::PrepareVuStatus (output.hpp:15-19), -1 if triangulation failed */
    int rows, cols;                      /* of H: 2 n_obs x l (truncated at the last pose the track touches) */
    double pf[3], depth;                 /* triangulated point, |pf - first camera| */
    const double* d_H;                   /* DEVICE: rows x cols, column-major, ld = rows */
    const double* d_f;                   /* DEVICE: predicted observations, rows */
    const double* d_y;                   /* DEVICE: the observations, rows (the y of visualTrackOutlierCheck / updateVisualTrack) */
} hv_track_model;
/* All tracks are evaluated against the CURRENT state mean in one launch (one CTA per track); the device pointers in out[]
 * stay valid until the next call on this ekf. Synchronises (the caller branches on the statuses). */
int hv_ekf_track_models(hv_ekf* ekf, const hv_track_obs* tracks, int ntracks, hv_track_model* out);
/* visualTrackOutlierCheck (mode 0) / updateVisualTrack (mode 1) / check then update-if-inlier (mode 2) (ekf.cpp:787-844) on
 * the device-resident H, f, y of one track of the last hv_ekf_track_models call. Modes 0 and 2 return the VuOutlierStatus and
 * chi2 (one host round trip, like hv_ekf_visual_check); mode 1 is asynchronous. Note that an update changes the state: the
 * models of the other tracks of that call were evaluated against the state before it (as in the reference's batch mode,
 * backend.cpp:1170-1183; per-track mode re-evaluates the next track with a new hv_ekf_track_models call). */
int hv_ekf_visual_track(hv_ekf* ekf, const hv_track_model* t, double r, double track_rmse_threshold, int mode, int* vu_status,
                        double* chi2);
/* The per-track loop of Session::trackerVisualUpdate in per-track mode (src/odometry/backend.cpp:1012-1252) as ONE chain on the
 * stream, with the control flow on the device: for every track, in order,
 *     measurement model against the CURRENT state (the previous track's update included)
 *  -> visualTrackOutlierCheck(chi_outlier_r, track_rmse_threshold)       if the model is valid
 *  -> updateVisualTrack(visual_r)                                         if the check says INLIER,
 * and nothing more once max_successful_updates updates have been applied (backend.cpp:1240-1247). Check and update of a track
 * run as ONE kernel (H P and H P H' formed once, factorised with each of the two noise levels; HV_CHAIN_SEPARATE=1 in the
 * environment issues two gated launches instead), so a track costs two launches. Each kernel is gated by
 * words its predecessors wrote, so the host does not synchronise per track but once per `lookahead` tracks (0: once). The
 * caller applies its own pre-filters (track score, trackMinFrames, blacklist, maxVisualUpdates: backend.cpp:1020-1047, 1241)
 * by choosing which tracks to submit. Results are those of the per-track calls hv_ekf_track_models ->
 * hv_ekf_visual_track(mode 0) -> hv_ekf_visual_track(mode 1) issued track by track.
 * Not covered: trackOutlierThresholdGrowthFactor != 1 (the thresholds are fixed for the chain), hybrid map points. */
typedef struct hv_visual_update_params {
    double chi_outlier_r;            /* r of the check: odometry.trackChiTestOutlierR / focal length (backend.cpp:996) */
    double track_rmse_threshold;     /* odometry.trackRmseThreshold / focal length (backend.cpp:995); < 0: off */
    double visual_r;                 /* r of the update: odometry.visualR / focal length (backend.cpp:997) */
    int max_successful_updates;      /* odometry.maxSuccessfulVisualUpdates; <= 0: unlimited */
    int lookahead;                   /* tracks issued per host synchronisation; 0 = all */
} hv_visual_update_params;
typedef struct hv_track_result {
    int triangulator_status;         /* odometry::TriangulatorStatus; -1: not attempted (enough successful updates before it) */
    int prepare_vu_status;           /* odometry::PrepareVuStatus, -1 if triangulation failed / not attempted */
    int outlier_status;              /* odometry::VuOutlierStatus (INLIER 0, NOT_COMPUTED 1, RMSE 2, CHI2 3) */
    int updated;                     /* 1: updateVisualTrack was applied with this track */
    double chi2, pf[3], depth;
} hv_track_result;
int hv_ekf_visual_tracks(hv_ekf* ekf, const hv_track_obs* tracks, int ntracks, const hv_visual_update_params* params,
                         hv_track_result* out, int* successful_updates);
/* Test / debug: copies H (rows x cols), f (rows) and d pf / d (poses, t) (3 x (7 npose + 1), column-major, after the stereo
 * sum) of track `track` of the last hv_ekf_track_models call to the host; any pointer may be NULL. */
int hv_ekf_track_model_download(hv_ekf* ekf, int track, double* H, double* f, double* dpf);

/* Measurement aid (bench): repeats the kernel of the last hv_ekf_track_models call `reps` times between two CUDA events on the
 * context's stream and returns the average device time per launch in milliseconds. */
int hv_ekf_track_models_time(hv_ekf* ekf, int reps, float* ms_per_launch);
/* Debug: the 32 result words of the last update kernel ([0] status, [1] chi2, [2] flag, [8..] phase timestamps when
 * the library is built with -DHV_EKF_TIMING). */
int hv_ekf_debug_result_words(hv_ekf* ekf, double* out32);
/* Host wall time of the most recent hv_ekf_run_host list that had something to hand back: out4 = {issuing the list, waiting in its one
 * synchronisation, total} in microseconds and the number of ops (bench.py reports them next to `e2e`). */
int hv_ekf_debug_host_times(hv_ekf* ekf, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* HYBVIO_B200_H_ */
