// hybvio_b200/csrc/track_model.cu -- hv_track_model_kernel: triangulation + prepareVisualUpdate of a batch of tracks on the
// device (one CTA per track; body and algorithm notes in track_model.cuh). Replaces, for the EKF's visual updates, the host
// sequence extractCameraPoseTrail -> Triangulator::triangulate -> prepareVisualUpdate of src/odometry/backend.cpp:1050-1160.
#include "track_model.cuh"

__global__ void __launch_bounds__(TM_NT, 2) hv_track_model_kernel(TmArgs a)
{
    extern __shared__ __align__(16) double tm_dyn[];
    tm_body(a, tm_dyn);
}

cudaError_t tm_launch(const TmArgs& a, cudaStream_t s)
{
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(hv_track_model_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tm_smem_bytes());
        if (e != cudaSuccess) return e;
        attr = true;
    }
    hv_track_model_kernel<<<a.ntracks, TM_NT, tm_smem_bytes(), s>>>(a);
    return cudaGetLastError();
}
