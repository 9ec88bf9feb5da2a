// hybvio_b200/csrc/ekf_cluster.cu -- Kalman update / outlier check / pose augmentation as ONE launch of a
// thread-block CLUSTER of 8 CTAs (8 SMs), fp64, working set in shared memory.
//
// Same algebra as ekf.cu (elimination tableau, Z = D^-1/2 L^-1 HP, P -= Z'Z, Joseph form for the augmentation) and
// the same reference functions (src/odometry/ekf.cpp:57-82, 573-677, 760-844, 848-885); what changes is the mapping:
//
//   * the state dimension is split into 8 column blocks J_c (20 columns each for N = 160); CTA c owns P[:, J_c];
//   * phase A  every CTA stages H (n x l) and P[0:l, J_c] in shared memory with coalesced loads and forms its slice
//              HP[:, J_c] -- the n x l x N product is spread over 8 SMs;
//   * phase B  S = HP[:, 0:l] H' + R is a sum over column blocks: every CTA forms its partial n x n product, the
//              partials are reduce-scattered / all-gathered through L2 in a FIXED order (bitwise deterministic);
//   * phase C  every CTA eliminates its own tableau [S | HP_Jc | v] -- the S part redundantly (identical operations
//              in identical order, hence identical bits and identical accept/reject decisions in all CTAs);
//   * phase E  the slices Z[:, J_c] are exchanged through L2 once, then CTA c applies P[:, J_c] -= Z' Z[:, J_c]
//              from shared memory with 4x2 register tiles, and CTA 0 applies m += Z' z_v and the normalisation.
//
// Three cluster barriers per update (two per check). Nothing is latency-bound on single global loads any more:
// every global access is a coalesced block copy with many loads in flight, all arithmetic runs out of shared memory.
#include "ekf.cuh"
#include <cooperative_groups.h>
#include <math.h>
namespace cg = cooperative_groups;

#define EKC 8            // cluster size (portable maximum)
#define EKC_NT 512
// Optional phase timestamps (globaltimer, ns) into res[8 + i]; compiled in with -DHV_EKF_TIMING (tools/ekf_phases.py)
#ifdef HV_EKF_TIMING
#define PHASE_MARK(i) do { if (c == 0 && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.b.res[8 + (i)] = (double)t_; } } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

// Block copy global -> shared with 8 independent loads in flight per thread (one L2 round trip moves 8 x 512 doubles)
__device__ __forceinline__ void cta_copy8(double* __restrict__ dst, const double* __restrict__ src, int count, int tid)
{
    for (int base = 0; base < count; base += 8 * EKC_NT) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const int i = base + u * EKC_NT + tid; r[u] = i < count ? src[i] : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; u++) { const int i = base + u * EKC_NT + tid; if (i < count) dst[i] = r[u]; }
    }
}

__device__ __forceinline__ void ck_normalize_quat(double* q)
{
    const double z = (q[0] * q[0] + q[2] * q[2]) + (q[1] * q[1] + q[3] * q[3]);
    if (z > 0.0) { const double nrm = sqrt(z); q[0] /= nrm; q[1] /= nrm; q[2] /= nrm; q[3] /= nrm; }
}
__device__ __forceinline__ int ck_aug_src(int i, int drop)
{
    if (i < EKF_CAM) return i;
    if (i < EKF_CAM + EKF_POSE) return -1;
    if (i < EKF_CAM + (drop + 1) * EKF_POSE) return i - EKF_POSE;
    return i;
}
__device__ __forceinline__ int ck_special_col(int c) { return c < 3 ? EKF_POS + c : c < 7 ? EKF_ORI + c - 3 : EKF_CAM + c - 7; }


#include "ekf_elim.cuh"

struct EkcGeom { int B, X, W, T, PC; };   // column block, sizes (doubles) of the shared-memory regions
__host__ __device__ inline EkcGeom ekc_geom(int n, int l, int N, bool joseph)
{
    EkcGeom g;
    g.B = (N + EKC - 1) / EKC;
    g.X = n * (l > N ? l : N);                         // H (n x l), later the gathered Z (n x N)
    g.W = (n + g.B + 1 + (joseph ? n : 0)) | 1;        // tableau row: [S | HP_J | v | (I)]
    g.T = n * g.W;
    const int pc = l * g.B, jz = joseph ? N * 21 : 0;  // P[0:l, J] staging, later K (N x 7) and T1's 14 columns
    g.PC = pc > jz ? pc : jz;
    return g;
}

__device__ __forceinline__ void ekf_cluster_body(EkfUpdateArgs& a)
{
    extern __shared__ double sm[];
    __shared__ double s_scalar[2];
    __shared__ double s_elim[ELIM_SMEM_DOUBLES];
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5, nwarps = EKC_NT / 32;
    const int N = a.b.N, n = a.n, l = a.l;
    const bool joseph = a.op == EKF_OP_AUGMENT;
    const EkcGeom g = ekc_geom(n, l, N, joseph);
    double* X = sm;                 // H, later Z
    double* T = X + g.X;            // tableau
    double* PC = T + g.T;           // P[0:l, J], later K | T1c
    const int W = g.W, B = g.B;
    const int J0 = c * B, Bc = max(0, min(B, N - J0));
    const int vcol = n + B, cend = joseph ? vcol + n : vcol;
    double* m = a.b.m;
    double* P = a.b.P;
    double* Spart = a.b.cwork;                       // EKC x n x n partial innovation covariances
    double* Sg = Spart + (size_t)EKC * N * N;        // reduced S
    double* Zg = Sg + (size_t)N * N;                 // gathered Z, row-major n x N

    PHASE_MARK(0);
    // ---- phase 0 (augmentation): m = A m, P2 = A P A' + visAugQ (ekf.cpp:853-857); CTA c writes columns J_c
    if (joseph) {
        const int drop = a.dropIdx;
        double* P2 = a.b.P2;
        for (int idx = tid; idx < N * Bc; idx += EKC_NT) {
            const int i = idx % N, j = J0 + idx / N;
            const int si = ck_aug_src(i, drop), sj = ck_aug_src(j, drop);
            double v = (si < 0 || sj < 0) ? 0.0 : P[si + (size_t)sj * N];
            // deferred maintainPositiveSemiDefinite (ekf.cpp:1059-1067): 0.5 (P + P') evaluated while the shift reads P
            if (a.symFirst && si >= 0 && sj >= 0 && si != sj) v = 0.5 * (v + P[sj + (size_t)si * N]);
            if (i == j && i >= EKF_CAM && i < EKF_CAM + EKF_POSE) v += (i - EKF_CAM) < 3 ? a.augNoisePos : a.augNoiseOri;
            P2[i + (size_t)j * N] = v;
        }
        if (c == 0) {
            double tmp[2];
#pragma unroll
            for (int r = 0; r < 2; r++) { const int i = tid + r * EKC_NT; if (i < N) { const int s = ck_aug_src(i, drop); tmp[r] = s < 0 ? 0.0 : m[s]; } }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 2; r++) { const int i = tid + r * EKC_NT; if (i < N) m[i] = tmp[r]; }
        }
        cluster.sync();
        P = P2;
    }

    PHASE_MARK(1);
    // ---- phase 1: measurement model into shared memory (ld = n)
    double hspeed = 0.0;
    if (a.op == EKF_OP_DENSE) {
        cta_copy8(X, a.H, n * l, tid);
    } else {
        for (int i = tid; i < n * l; i += EKC_NT) X[i] = 0.0;
        if (a.op == EKF_OP_PSEUDO_VELOCITY) {
            hspeed = sqrt(m[EKF_VEL] * m[EKF_VEL] + m[EKF_VEL + 1] * m[EKF_VEL + 1]);
            if (hspeed <= 1e-7) return;                               // ekf.cpp:635-637 (uniform over the cluster)
        }
        __syncthreads();
        if (tid == 0) {
            switch (a.op) {
                case EKF_OP_ZUPT: for (int i = 0; i < 3; i++) X[i + (EKF_VEL + i) * n] = 1.0; break;
                case EKF_OP_ZRUPT: for (int i = 0; i < 3; i++) X[i + (EKF_BGA + i) * n] = 1.0; break;
                case EKF_OP_PSEUDO_VELOCITY: for (int i = 0; i < 2; i++) X[(EKF_VEL + i) * n] = m[EKF_VEL + i] / hspeed; break;
                case EKF_OP_POSITION: for (int i = 0; i < 3; i++) X[i + (EKF_POS + i) * n] = 1.0; break;
                case EKF_OP_ZERO_HEIGHT: X[(EKF_POS + 2) * n] = 1.0; break;
                case EKF_OP_ORIENTATION: for (int i = 0; i < 4; i++) X[i + (EKF_ORI + i) * n] = 1.0; break;
                case EKF_OP_AUGMENT:
                    for (int i = 0; i < 3; i++) { X[i + (EKF_POS + i) * n] = 1.0; X[i + (EKF_CAM + i) * n] = -1.0; }
                    for (int i = 0; i < 4; i++) { X[3 + i + (EKF_ORI + i) * n] = 1.0; X[3 + i + (EKF_CAM + 3 + i) * n] = -1.0; }
                    break;
            }
        }
    }
    // stage P[0:l, J_c] (columns are contiguous in memory)
    if (l == N) cta_copy8(PC, P + (size_t)J0 * N, l * Bc, tid);      // whole columns: one contiguous block
    else {
#pragma unroll 4
        for (int idx = tid; idx < l * Bc; idx += EKC_NT) {
            const int k = idx % l, jj = idx / l;
            PC[idx] = P[k + (size_t)(J0 + jj) * N];
        }
    }
    __syncthreads();
    const double* Hs = X;

    // residual (identical in every CTA)
    for (int i = tid; i < n; i += EKC_NT) {
        double v;
        if (a.op == EKF_OP_PSEUDO_VELOCITY) v = a.defaultSpeed - hspeed;
        else {
            const double yi = a.y ? a.y[i] : a.ysmall[i];
            double fi = 0.0;
            if (a.f) fi = a.f[i];
            else for (int k = 0; k < l; k++) fi += Hs[i + (size_t)k * n] * m[k];
            v = yi - fi;
        }
        T[(size_t)i * W + vcol] = v;
    }
    if (joseph) for (int t = tid; t < n * n; t += EKC_NT) T[(size_t)(t / n) * W + vcol + 1 + (t % n)] = (t / n == t % n) ? 1.0 : 0.0;
    __syncthreads();

    const bool checking = a.mode != EKF_MODE_UPDATE;
    if (checking && a.rmseThr >= 0.0) {               // ekf.cpp:797-801
        if (tid == 0) { double ss = 0.0; for (int i = 0; i < n; i++) { const double v = T[(size_t)i * W + vcol]; ss += v * v; } s_scalar[0] = sqrt(ss / n); }
        __syncthreads();
        if (s_scalar[0] > a.rmseThr) { if (c == 0 && tid == 0) { a.b.res[0] = 2.0; a.b.res[1] = 0.0; a.b.res[2] = 0.0; } return; }
    }
    if (checking && a.skipChi2 && a.mode == EKF_MODE_CHECK) {
        if (c == 0 && tid == 0) { a.b.res[0] = 0.0; a.b.res[1] = 0.0; a.b.res[2] = 0.0; }
        return;
    }

    PHASE_MARK(2);
    // ---- phase A: HP[:, J_c] = H P[0:l, J_c]   (2 x 2 register tiles out of shared memory)
    {
        // 4 x 4 register tiles (rows i0 + x*tm: consecutive lanes -> consecutive H entries, columns j0 + y*tn): 8 shared
        // loads feed 16 DFMAs per k. The k range is split over KS thread groups whose partial tiles are summed in
        // shared memory, so that all 512 threads work even when there are few tiles.
        const int tm = (n + 3) >> 2, tn = (Bc + 3) >> 2, ntile = tm * tn;
        int KS = min(EKC_NT / max(ntile, 1), l / 24); KS = KS < 1 ? 1 : (KS > 8 ? 8 : KS);   // a slice is worth >= 24 k's
        const int klen = (l + KS - 1) / KS;
        // zero the HP slice (partials are accumulated into it)
        for (int t = tid; t < n * Bc; t += EKC_NT) T[(size_t)(t / Bc) * W + n + (t % Bc)] = 0.0;
        __syncthreads();
        // every group computes the partial tile of its own k-slice in registers (all groups in parallel) ...
        const int g = tid / ntile, t = tid - g * ntile;
        const int ti = t % tm, tj = t / tm;
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) acc[x][y] = 0.0;
        if (g < KS) {
            const int k0 = g * klen, k1 = min(l, k0 + klen);
            int iv[4], jv[4];
#pragma unroll
            for (int x = 0; x < 4; x++) { iv[x] = min(ti + x * tm, n - 1); jv[x] = min(tj + x * tn, Bc - 1); }
            for (int k = k0; k < k1; k++) {
                double hv[4], bv[4];
#pragma unroll
                for (int x = 0; x < 4; x++) { hv[x] = Hs[iv[x] + (size_t)k * n]; bv[x] = PC[k + (size_t)jv[x] * l]; }
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y] += hv[x] * bv[y];
            }
        }
        // ... and the groups add them into the tile one after the other (fixed order: deterministic, race-free)
        for (int ks = 0; ks < KS; ks++) {
            if (g == ks) {
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) {
                        const int i = ti + x * tm, j = tj + y * tn;
                        if (i < n && j < Bc) T[(size_t)i * W + n + j] += acc[x][y];
                    }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    PHASE_MARK(3);
    // ---- phase B: partial S over the own columns that lie inside [0, l), to global, fixed-order reduction
    {
        const int kc = max(0, min(Bc, l - J0));
        double* mine = Spart + (size_t)c * n * n;
        // 2 x 4 register tiles over S(i, ip): per k two broadcast loads of HP and four consecutive loads of H per warp
        // instead of two loads per FMA
        const int ti_n = (n + 1) >> 1, tp_n = (n + 3) >> 2;
        for (int t = tid; t < ti_n * tp_n; t += EKC_NT) {
            const int tp = t % tp_n, ti = t / tp_n;
            const int i0 = ti, i1 = min(ti + ti_n, n - 1);
            int pv[4];
#pragma unroll
            for (int x = 0; x < 4; x++) pv[x] = min(tp + x * tp_n, n - 1);
            const double* hp0 = T + (size_t)i0 * W + n;
            const double* hp1 = T + (size_t)i1 * W + n;
            const double* hh = Hs + (size_t)J0 * n;
            double acc[2][4];
#pragma unroll
            for (int x = 0; x < 4; x++) { acc[0][x] = 0.0; acc[1][x] = 0.0; }
            for (int k = 0; k < kc; k++) {
                const double a0 = hp0[k], a1 = hp1[k];
#pragma unroll
                for (int x = 0; x < 4; x++) { const double h = hh[pv[x] + (size_t)k * n]; acc[0][x] += a0 * h; acc[1][x] += a1 * h; }
            }
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const int ip = tp + x * tp_n;
                if (ip < n) { mine[(size_t)i0 * n + ip] = acc[0][x]; if (ti + ti_n < n) mine[(size_t)(ti + ti_n) * n + ip] = acc[1][x]; }
            }
        }
    }
    PHASE_MARK(4);
    cluster.sync();
    PHASE_MARK(5);
    {
        const int E = (n * n + EKC - 1) / EKC, e0 = c * E, e1 = min(n * n, e0 + E);
        for (int e = e0 + tid; e < e1; e += EKC_NT) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < EKC; r++) s += Spart[(size_t)r * n * n + e];
            if (e / n == e % n) s += a.Rdiag;
            Sg[e] = s;
        }
    }
    cluster.sync();
#pragma unroll 4
    for (int t = tid; t < n * n; t += EKC_NT) T[(size_t)(t / n) * W + (t % n)] = Sg[t];
    __syncthreads();

    PHASE_MARK(6);
    // ---- phase C + D: unpivoted forward elimination of [S | HP_Jc | v | (I)], then Z = D^-1/2 (.)
    // Register-resident, cyclically distributed: warp w owns rows {w, w+16, ...}, lane q owns columns {q, q+32, ...}
    // (<= 6 x 4 elements per thread, n <= 96, row length <= 128). Per step only the pivot row and the multiplier
    // column travel through shared memory (double-buffered: one barrier per step); the rank-1 update itself is
    // 24 register FMAs per thread with warp-uniform skipping of retired rows / column blocks.
    bool bad = false;
    if (n <= ELIM_RA * 32 && cend < ELIM_CJ * 32) {
        double t[ELIM_RA][2][ELIM_CJ];
#pragma unroll
        for (int aa = 0; aa < ELIM_RA; aa++)
#pragma unroll
            for (int sr = 0; sr < 2; sr++)
#pragma unroll
                for (int bb = 0; bb < ELIM_CJ; bb++) {
                    const int i = elim_row(wrp, aa, sr), j = lane + 32 * bb;
                    t[aa][sr][bb] = (i < n && j <= cend) ? T[(size_t)i * W + j] : 0.0;
                }
        bad = !elim_dispatch(t, n, cend + 1, lane, wrp, s_elim);
        if (!bad) {
            __syncthreads();
            const double* pivots = s_elim + 2 * 2 * ELIM_ROWBUF + 8;
#pragma unroll
            for (int aa = 0; aa < ELIM_RA; aa++)
#pragma unroll
                for (int sr = 0; sr < 2; sr++) {
                    const int i = elim_row(wrp, aa, sr);
                    if (i < n) {
                        const double sc = 1.0 / sqrt(pivots[i]);
#pragma unroll
                        for (int bb = 0; bb < ELIM_CJ; bb++) {
                            const int j = lane + 32 * bb;
                            if (j >= n && j <= cend) T[(size_t)i * W + j] = t[aa][sr][bb] * sc;
                        }
                    }
                }
        }
    } else {
        for (int k = 0; k < n; k++) {
            const double piv = T[(size_t)k * W + k];
            if (!(piv > 0.0)) { bad = true; break; }
            const double rinv = 1.0 / piv;
            const double* rk = T + (size_t)k * W;
            for (int i = k + 1 + wrp; i < n; i += nwarps) {
                double* ri = T + (size_t)i * W;
                const double f = ri[k] * rinv;
                for (int j = k + 1 + lane; j <= cend; j += 32) ri[j] -= f * rk[j];
            }
            __syncthreads();
        }
        if (!bad) {
            for (int k = wrp; k < n; k += nwarps) {
                const double sc = 1.0 / sqrt(T[(size_t)k * W + k]);
                double* rk = T + (size_t)k * W;
                for (int j = n + lane; j <= cend; j += 32) rk[j] *= sc;
            }
        }
    }
    if (bad) { if (c == 0 && tid == 0) { a.b.res[0] = 1.0; a.b.res[1] = 0.0; a.b.res[2] = 1.0; } return; }   // uniform over the cluster
    __syncthreads();
    PHASE_MARK(7);
    if (tid == 0) { double t = 0.0; for (int k = 0; k < n; k++) { const double z = T[(size_t)k * W + vcol]; t += z * z; } s_scalar[1] = a.noiseScale * t; }
    __syncthreads();
    const double chi2 = s_scalar[1];
    if (checking) {
        const bool outlier = !a.skipChi2 && chi2 > a.chi2Thr;
        if (c == 0 && tid == 0) { a.b.res[0] = outlier ? 3.0 : 0.0; a.b.res[1] = chi2; a.b.res[2] = 0.0; }
        if (outlier || a.mode == EKF_MODE_CHECK) return;
    } else if (c == 0 && tid == 0) { a.b.res[0] = 0.0; a.b.res[1] = chi2; a.b.res[2] = 0.0; }

    PHASE_MARK(8);
    // ---- phase E: exchange Z slices through L2, then P[:, J_c] -= Z' Z[:, J_c]; CTA 0: m += Z' z_v
    for (int t = tid; t < n * Bc; t += EKC_NT) {
        const int jj = t % Bc, k = t / Bc;
        Zg[(size_t)k * N + J0 + jj] = T[(size_t)k * W + n + jj];
    }
    cluster.sync();
    double* Z = X;                                   // n x N row-major (H is dead)
    cta_copy8(Z, Zg, n * N, tid);
    __syncthreads();
    {
        // 4 x 2 register tiles; a thread's four rows are ti, ti + R4, ti + 2 R4, ti + 3 R4 so that the lanes of a warp read
        // CONSECUTIVE doubles of a Z row (conflict-free) and update consecutive rows of a P column (coalesced)
        const int R4 = (N + 3) >> 2, tj_n = (Bc + 1) >> 1;
        for (int t = tid; t < R4 * tj_n; t += EKC_NT) {
            const int ti = t % R4, tj = t / R4;
            const int j0 = J0 + tj * 2;
            const bool j1ok = tj * 2 + 1 < Bc;
            int iv[4];
#pragma unroll
            for (int x = 0; x < 4; x++) iv[x] = min(ti + x * R4, N - 1);
            double acc[4][2];
#pragma unroll
            for (int x = 0; x < 4; x++) { acc[x][0] = 0.0; acc[x][1] = 0.0; }
            const double* zr = Z;
#pragma unroll 2
            for (int k = 0; k < n; k++) {
                double av[4];
#pragma unroll
                for (int x = 0; x < 4; x++) av[x] = zr[iv[x]];
                const double b0 = zr[j0], b1 = zr[j1ok ? j0 + 1 : j0];
#pragma unroll
                for (int x = 0; x < 4; x++) { acc[x][0] += av[x] * b0; acc[x][1] += av[x] * b1; }
                zr += N;
            }
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const int i = ti + x * R4;
                if (i < N) { P[i + (size_t)j0 * N] -= acc[x][0]; if (j1ok) P[i + (size_t)(j0 + 1) * N] -= acc[x][1]; }
            }
        }
    }
    if (c == 0) {
        for (int i = tid; i < N; i += EKC_NT) {
            double s = 0.0;
            for (int k = 0; k < n; k++) s += Z[(size_t)k * N + i] * T[(size_t)k * W + vcol];
            m[i] += s;
        }
        __syncthreads();
        // quaternion normalisation: updateCommon normalises the current orientation only, the visual update and
        // the augmentation all of them (ekf.cpp:31, 843, 874)
        for (int q = tid; q < (a.normalizeAll ? a.b.trail + 1 : 1); q += EKC_NT)
            ck_normalize_quat(q == 0 ? m + EKF_ORI : m + EKF_CAM + EKF_POSE * (q - 1) + 3);
    }

    PHASE_MARK(9);
    double* Pfinal = P;
    if (joseph) {
        // ---- Joseph form (ekf.cpp:35-50): P'' = G T1' + K R K', G = T1 P' (in P2 now), T1 = I - K visAugH; see ekf.cu
        double* Ks = PC;                             // N x 7
        double* T1c = PC + (size_t)N * EKF_POSE;     // N x 14
        for (int t = tid; t < N * EKF_POSE; t += EKC_NT) {
            const int i = t % N, r = t / N;
            double s = 0.0;
            for (int k = 0; k < n; k++) s += Z[(size_t)k * N + i] * T[(size_t)k * W + vcol + 1 + r];
            Ks[t] = s;
        }
        __syncthreads();
        for (int t = tid; t < N * 14; t += EKC_NT) {
            const int j = t % N, cc = t / N;
            const double kv = cc < 7 ? -Ks[j + cc * N] : Ks[j + (cc - 7) * N];
            T1c[t] = (j == ck_special_col(cc) ? 1.0 : 0.0) + kv;
        }
        cluster.sync();                              // all of G is final (every CTA finished its columns)
        const double* G = P;
        double* Pout = a.b.P;
        for (int idx = tid; idx < N * Bc; idx += EKC_NT) {
            const int i = idx % N, j = J0 + idx / N;
            const bool jsp = j < 3 || (j >= EKF_ORI && j < EKF_ORI + 4) || (j >= EKF_CAM && j < EKF_CAM + EKF_POSE);
            double s = jsp ? 0.0 : G[i + (size_t)j * N];
#pragma unroll
            for (int cc = 0; cc < 14; cc++) s += G[i + (size_t)ck_special_col(cc) * N] * T1c[j + cc * N];
            double kr = 0.0;
#pragma unroll
            for (int r = 0; r < EKF_POSE; r++) kr += Ks[i + r * N] * (a.Rdiag * Ks[j + r * N]);
            Pout[i + (size_t)j * N] = s + kr;
        }
        Pfinal = Pout;
    }
    if (a.symmetrize) {
        cluster.sync();                              // all columns of the final P are written
        for (int idx = tid; idx < N * Bc; idx += EKC_NT) {
            const int i = idx % N, j = J0 + idx / N;
            if (i > j) {
                const double s = 0.5 * (Pfinal[i + (size_t)j * N] + Pfinal[j + (size_t)i * N]);
                Pfinal[i + (size_t)j * N] = s; Pfinal[j + (size_t)i * N] = s;
            }
        }
    }
}

__global__ void __cluster_dims__(EKC, 1, 1) __launch_bounds__(EKC_NT) ekf_update_cluster_kernel(EkfUpdateArgs a)
{
    ekf_cluster_body(a);
}

// Batched outlier checks: cluster i works on measurement i against the same state (read-only), with its own
// exchange buffers and result words. This is what makes the 15-20 candidate tracks of a frame one launch.
__global__ void __cluster_dims__(EKC, 1, 1) __launch_bounds__(EKC_NT) ekf_check_batch_cluster_kernel(EkfUpdateArgs a, EkfCheckBatch b)
{
    const int inst = blockIdx.x / EKC;
    const EkfCheckItem& it = b.it[inst];
    a.H = it.H; a.f = it.f; a.y = it.y; a.n = it.n; a.l = it.l;
    a.Rdiag = it.Rdiag; a.chi2Thr = it.chi2Thr; a.rmseThr = it.rmseThr; a.skipChi2 = it.skipChi2;
    a.b.res += (size_t)EKF_RES_STRIDE * inst;
    a.b.cwork += (size_t)inst * 10 * a.b.N * a.b.N;
    ekf_cluster_body(a);
}

size_t ekf_cluster_smem_bytes(int n, int l, int N, bool joseph)
{
    const EkcGeom g = ekc_geom(n, l, N, joseph);
    return ((size_t)g.X + g.T + g.PC) * sizeof(double);
}

cudaError_t ekf_launch_update_cluster(const EkfUpdateArgs& a, cudaStream_t s)
{
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ekf_update_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    const size_t smem = ekf_cluster_smem_bytes(a.n, a.l, a.b.N, a.op == EKF_OP_AUGMENT);
    ekf_update_cluster_kernel<<<EKC, EKC_NT, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t ekf_launch_check_batch(const EkfUpdateArgs& a, const EkfCheckBatch& b, cudaStream_t s)
{
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(ekf_check_batch_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    size_t smem = 0;
    for (int i = 0; i < b.count; i++) { const size_t v = ekf_cluster_smem_bytes(b.it[i].n, b.it[i].l, a.b.N, false); if (v > smem) smem = v; }
    ekf_check_batch_cluster_kernel<<<EKC * b.count, EKC_NT, smem, s>>>(a, b);
    return cudaGetLastError();
}
