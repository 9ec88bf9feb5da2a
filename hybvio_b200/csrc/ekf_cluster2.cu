// hybvio_b200/csrc/ekf_cluster2.cu -- kernels and launchers of the second-generation cluster update (ekf_cluster2.cuh):
// P column blocks resident in shared memory, all inter-CTA exchanges through distributed shared memory.
#include <cooperative_groups.h>
#include "hv_device_once.cuh"
#include <math.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#ifdef HV_EKF_TIMING
// phase timestamps (globaltimer, ns) into res[8 + i] (tools/ekf_phases.py)
#define EK2_PHASE(i) do { if (c == 0 && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.b.res[8 + (i)] = (double)t_; } } while (0)
#endif
#include "ekf_cluster2.cuh"

__global__ void __launch_bounds__(EK2_NT) ekf_update_cluster2_kernel(EkfUpdateArgs a)
{
    extern __shared__ __align__(16) double ek2_sm[];
    ek2_body(a, ek2_sm, cg::this_cluster());
}

// Batched outlier checks: cluster i works on measurement i against the same state (read-only), with its own result words.
// b.withAugment: one more cluster runs the pose augmentation that FOLLOWS the checks in the caller's sequence (aug: its argument block).
// The checks only read (m, P) and the augmentation writes its result to the second buffers (aug.specP / aug.specM, adopted by the
// host with a pointer swap), so the two are independent and share the launch instead of queueing behind each other.
__global__ void __launch_bounds__(EK2_NT) ekf_check_batch_cluster2_kernel(EkfUpdateArgs a, EkfCheckBatch b, EkfUpdateArgs aug)
{
    extern __shared__ __align__(16) double ek2_sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int inst = blockIdx.x / (int)cluster.num_blocks();
    if (inst >= b.count) a = aug;                               // (one call of the body: its code is 340 KB)
    else {
        const EkfCheckItem& it = b.it[inst];
        a.H = it.H; a.f = it.f; a.y = it.y; a.n = it.n; a.l = it.l;
        a.Rdiag = it.Rdiag; a.chi2Thr = it.chi2Thr; a.rmseThr = it.rmseThr; a.skipChi2 = it.skipChi2;
        if (a.sig) a.sig += 4 * inst;
        if (a.slot) a.slot += 4 * inst;
    }
    a.b.res += (size_t)EKF_RES_STRIDE * inst;
    a.b.cwork += (size_t)inst * 10 * a.b.N * a.b.N;           // own exchange area (Z | reduced S | partial S)
    ek2_body(a, ek2_sm, cluster);
}

#define EK2_STATIC_SMEM (sizeof(double) * (2 + EK2_LINV_DOUBLES + 2 + EK2_MAXN) + 256)
#define EK2_SMEM_LIMIT (227 * 1024)

// Cluster size 8 (the portable maximum). Measured on B200 (round 2, profiles/r02_ab_settled.md): a 16-CTA cluster (non-portable size)
// was slower for the frame as a whole (3755 against 4152 frames/s): the dense products halve but every exchange gets slower.
static int ek2_cluster_size() { return 8; }
static int ek2_cluster_size_for(int, bool) { return 8; }

bool ekf_cluster2_fits(int n, int l, int N, bool joseph)
{
    return N <= EK2_MAXN && ek2_smem_bytes(n, l, N, joseph, ek2_cluster_size()) + EK2_STATIC_SMEM <= EK2_SMEM_LIMIT;
}

template <class K, class... Args>
static cudaError_t ek2_launch(K kernel, int C, int nclusters, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C * nclusters); cfg.blockDim = dim3(EK2_NT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    // programmatic dependent launch: this kernel may start while the previous KERNEL of the stream is still running (it waits
    // in griddepcontrol.wait before it reads the filter state); HV_EKF_NO_PDL=1 switches it off (A/B)
    static const bool pdl = getenv("HV_EKF_NO_PDL") == nullptr;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <class K>
static cudaError_t ek2_prepare(K kernel, int C, size_t staticSmem = EK2_STATIC_SMEM)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(EK2_SMEM_LIMIT - staticSmem));
    if (e != cudaSuccess) return e;
    (void)C;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}

cudaError_t ekf_launch_update_cluster2(const EkfUpdateArgs& a, cudaStream_t s)
{
    const int C = ek2_cluster_size_for(a.n, false);
    static bool seen[64];                             // per device (hv_common.cuh)
    if (hv_first_use_on_device(seen)) { cudaError_t e = ek2_prepare(ekf_update_cluster2_kernel, C); if (e != cudaSuccess) return e; }
    const size_t smem = ek2_smem_bytes(a.n, a.l, a.b.N, a.op == EKF_OP_AUGMENT, C);
    return ek2_launch(ekf_update_cluster2_kernel, C, 1, smem, s, a);
}

cudaError_t ekf_launch_check_batch2(const EkfUpdateArgs& a, const EkfCheckBatch& b, cudaStream_t s, const EkfUpdateArgs* aug)
{
    const int C = ek2_cluster_size();
    static bool seen[64];
    if (hv_first_use_on_device(seen)) { cudaError_t e = ek2_prepare(ekf_check_batch_cluster2_kernel, C); if (e != cudaSuccess) return e; }
    size_t smem = 0;
    for (int i = 0; i < b.count; i++) { const size_t v = ek2_smem_bytes(b.it[i].n, b.it[i].l, a.b.N, false, C); if (v > smem) smem = v; }
    if (aug) { const size_t v = ek2_smem_bytes(aug->n, aug->l, a.b.N, true, C); if (v > smem) smem = v; }
    return ek2_launch(ekf_check_batch_cluster2_kernel, C, b.count + (aug ? 1 : 0), smem, s, a, b, aug ? *aug : a);
}
