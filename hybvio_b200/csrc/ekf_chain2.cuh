// hybvio_b200/csrc/ekf_chain2.cuh -- the per-track loop of Session::trackerVisualUpdate (src/odometry/backend.cpp:1012-1252, per-track
// mode) as ONE persistent launch of an 8-CTA cluster: for every track of the list, in order,
//   CTA 0:    measurement model against the current state mean (tm_body of track_model.cuh; skipped once the success counter is full)
//   cluster:  outlier check with the check's noise level and, if INLIER, the update with the update's noise level
//             (ek2_body of ekf_cluster2.cuh in its two-R mode), the P column blocks staying in shared memory from track to track.
// Control flow is the same set of words the multi-launch chain uses (model status, success counter, result slots); what
// disappears are two launches and one staging of P per track. The model's scratch lives in the region H / Z occupy during the
// update (dead between two updates), sized by xCap.
#pragma once
#include "ekf_cluster2.cuh"
#include "track_model.cuh"

__host__ __device__ inline size_t ek2_chain_smem_bytes(const EkfChainList& c, int N, int C, int* xCap, int* tCap)
{
    int X = TM_S_TOTAL, T = 0, rest = 0, PB = 0;
    for (int i = 0; i < c.count; i++) {
        const Ek2Geom g = ek2_geom(c.it[i].n, c.it[i].l, N, false, C);
        if (g.X > X) X = g.X;
        if (g.T > T) T = g.T;
        if (g.RS > rest) rest = g.RS;
        PB = g.PB;
    }
    X = (X + 1) & ~1;                                  // keeps the tableau 16-byte aligned whatever TM_S_TOTAL is
    if (xCap) *xCap = X;
    if (tCap) *tCap = T;
    return ((size_t)X + T + PB + rest) * sizeof(double);
}

template <class Cluster>
__device__ __forceinline__ void ek2_chain_body(const EkfUpdateArgs& a, const TmArgs& tm, const EkfChainList& list, double* sm, Cluster cluster)
{
    const int c = (int)cluster.block_rank(), C = (int)cluster.num_blocks(), tid = threadIdx.x;
    const int N = a.b.N;
    // stage the own column block of P once (what ek2_body does for a first measurement); every ek2_body call below keeps it
    {
        const Ek2Geom g = ek2_geom(list.it[0].n, list.it[0].l, N, false, C);
        double* PB = sm + a.xCap + a.tCap;
        const int B = g.B, LD = g.LD, J0 = c * B, Bc = max(0, min(B, N - J0));
        ek2_pdl_launch_dependents();
        ek2_pdl_wait();
        const double* src = a.b.P + (size_t)J0 * N;
        for (int i = tid; i < N * Bc; i += EK2_NT) PB[(i % N) + (size_t)(i / N) * LD] = src[i];
        __syncthreads();
    }
    for (int k = 0; k < list.count; k++) {
        const int trk = list.first + k;
        // model: CTA 0, in the region of H / Z (its own early-outs -- skipped, triangulation failed -- are uniform over the CTA)
        if (c == 0) {
            TmArgs t = tm;
            t.trackOffset = trk; t.pdl = 0;
            tm_body(t, sm);
        }
        cluster.sync();                                   // H, f and the model's status words are visible to the whole cluster
        const EkfChainItem& it = list.it[k];
        EkfUpdateArgs b = a;
        b.H = tm.H + (size_t)trk * tm.Hstride; b.f = tm.f + (size_t)trk * 2 * TM_MAXOBS; b.y = tm.ip + (size_t)trk * 2 * TM_MAXOBS;
        b.n = it.n; b.l = it.l; b.mode = EKF_MODE_CHECK_UPDATE; b.skipChi2 = 0;
        b.Rdiag = list.RdiagCheck; b.Rdiag2 = list.RdiagUpdate; b.chi2Thr = it.chi2Thr; b.rmseThr = list.rmseThr; b.slot = it.slot;
        b.gateI = tm.status + 4 * (size_t)trk + 1; b.gateIExpect = 0; b.counter = tm.counter; b.counterMax = tm.counterMax; b.bump = (int*)tm.counter;
        b.keepBlock = 1; b.lateH = 1;
        ek2_body(b, sm, cluster);
        cluster.sync();                                   // the counter bump and the new state mean are visible before the next model reads them
    }
}
