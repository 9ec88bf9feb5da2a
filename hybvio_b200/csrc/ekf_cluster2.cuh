// hybvio_b200/csrc/ekf_cluster2.cuh -- body of the Kalman update / outlier check / pose augmentation kernel (cluster of C
// CTAs, fp64): second generation of ekf_cluster.cu. Same reference functions (src/odometry/ekf.cpp:57-82, 573-677,
// 760-844, 848-885, 35-50) and the same algebra (elimination tableau [S | HP | v], Z = D^-1/2 L^-1 HP, P -= Z'Z, Joseph
// form with the explicit 14-column T1); what changed is where the data lives and how the CTAs talk:
//
//   * CTA c owns the column block J_c of P and keeps ALL of it (N x B) in shared memory from the first load to the
//     final store: P is read once and written once per update (the first generation re-read it for the downdate);
//   * every exchange between CTAs goes through DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank) instead of a
//     write -> cluster barrier -> read round trip through L2:
//       - innovation covariance: every CTA leaves its partial S in its own tableau; after one cluster barrier the
//         partials are summed in a fixed order (bitwise identical in every CTA) -- directly by everybody when S is small
//         (n*n <= 1024), otherwise reduce-scatter + all-gather;
//       - the Z slices are gathered straight out of the neighbours' tableaus;
//       - symmetrisation reads the mirrored entry from the owner's block; the Joseph form reads the 14 special columns
//         of G from the CTAs that own them;
//   * the augmentation builds its shifted block A P A' + Q while loading (no P2 pass, no barrier), and a deferred
//     maintainPositiveSemiDefinite() is applied in the same pass (EkfUpdateArgs::symFirst);
//   * the state mean is staged in shared memory; CTA 0 writes it back once.
//
// Decisions (chi2 / RMSE / pivot sign) are computed redundantly from identical data in identical order, so that every
// CTA takes the same branch; every path that leaves the kernel after the first exposure of shared memory to the
// neighbours passes a final cluster barrier (a CTA must not exit while its shared memory may still be read).
//
// Written against the primitives tools/emu can run on the host (tools/emu/emu_update.cpp runs this body against the C
// oracle without a GPU).
#pragma once
#include "ekf.cuh"
#include "ekf_elim.cuh"

#define EK2_NT 512
#define EK2_MAXN 768
#ifndef EK2_PHASE
#define EK2_PHASE(i) do { } while (0)
#endif

struct Ek2Geom { int C, B, X, W, T, PB, RS, EXTRA, oneStage; };
__host__ __device__ inline Ek2Geom ek2_geom(int n, int l, int N, bool joseph, int C)
{
    Ek2Geom g;
    g.C = C;
    g.B = (N + C - 1) / C;
    g.X = n * (l > N ? l : N);                              // H (n x l, ld n), later the gathered Z (n x N, ld N)
    g.W = (n + g.B + 1 + (joseph ? n : 0)) | 1;             // tableau row: [S | HP_J | v | (I)]
    g.T = n * g.W;
    g.PB = N * g.B;                                         // own column block of P, ld N
    const int E = (n * n + C - 1) / C;
    const int cend = n + g.B + (joseph ? n : 0);
    // small S is summed directly by every CTA into RS (and eliminated from there: register path only)
    g.oneStage = (n * n <= 1024 && n <= ELIM_RA * 32 && cend < ELIM_CJ * 32) ? 1 : 0;
    g.RS = g.oneStage ? n * n : E;                          // reduced S (small n: all of it; else the own slice)
    g.EXTRA = joseph ? N * (EKF_POSE + 14 + 14) + N * g.B : 0;   // K | T1c | special columns of G | P'' block
    return g;
}
__host__ __device__ inline size_t ek2_smem_bytes(int n, int l, int N, bool joseph, int C)
{
    const Ek2Geom g = ek2_geom(n, l, N, joseph, C);
    return ((size_t)g.X + g.T + g.PB + g.RS + g.EXTRA) * sizeof(double);
}

__device__ __forceinline__ void ek2_copy8(double* __restrict__ dst, const double* __restrict__ src, int count, int tid)
{
    for (int base = 0; base < count; base += 8 * EK2_NT) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const int i = base + u * EK2_NT + tid; r[u] = i < count ? src[i] : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; u++) { const int i = base + u * EK2_NT + tid; if (i < count) dst[i] = r[u]; }
    }
}
__device__ __forceinline__ void ek2_normalize_quat(double* q)
{
    const double z = (q[0] * q[0] + q[2] * q[2]) + (q[1] * q[1] + q[3] * q[3]);
    if (z > 0.0) { const double nrm = sqrt(z); q[0] /= nrm; q[1] /= nrm; q[2] /= nrm; q[3] /= nrm; }
}
__device__ __forceinline__ int ek2_aug_src(int i, int drop)
{
    if (i < EKF_CAM) return i;
    if (i < EKF_CAM + EKF_POSE) return -1;
    if (i < EKF_CAM + (drop + 1) * EKF_POSE) return i - EKF_POSE;
    return i;
}
__device__ __forceinline__ int ek2_special_col(int c) { return c < 3 ? EKF_POS + c : c < 7 ? EKF_ORI + c - 3 : EKF_CAM + c - 7; }

// `Cluster` is cooperative_groups::cluster_group (or the emulator's stand-in).
template <class Cluster>
__device__ __forceinline__ void ek2_body(EkfUpdateArgs& a, double* sm, Cluster cluster)
{
    __shared__ double s_scalar[2];
    __shared__ double s_elim[ELIM_SMEM_DOUBLES];
    __shared__ double s_m[EK2_MAXN];
    const int c = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5, nwarps = EK2_NT / 32;
    const int N = a.b.N, n = a.n, l = a.l;
    const bool joseph = a.op == EKF_OP_AUGMENT;
    const Ek2Geom g = ek2_geom(n, l, N, joseph, C);
    double* X = sm;                 // H, later Z
    double* T = X + g.X;            // tableau
    double* PB = T + g.T;           // P[:, J_c]
    double* RS = PB + g.PB;         // reduced S
    double* EX = RS + g.RS;         // Joseph-form extras
    const int W = g.W, B = g.B;
    const int J0 = c * B, Bc = max(0, min(B, N - J0));
    const int vcol = n + B, cend = joseph ? vcol + n : vcol;
    const bool oneStage = g.oneStage != 0;
    double* const P = a.b.P;

    EK2_PHASE(0);
    // ---- stage the state mean and the own column block of P (augmentation: of A P A' + visAugQ, ekf.cpp:853-857)
    if (joseph) {
        const int drop = a.dropIdx;
        for (int i = tid; i < N; i += EK2_NT) { const int s = ek2_aug_src(i, drop); s_m[i] = s < 0 ? 0.0 : a.b.m[s]; }
        for (int idx = tid; idx < N * Bc; idx += EK2_NT) {
            const int i = idx % N, j = J0 + idx / N;
            const int si = ek2_aug_src(i, drop), sj = ek2_aug_src(j, drop);
            double v = (si < 0 || sj < 0) ? 0.0 : P[si + (size_t)sj * N];
            // deferred maintainPositiveSemiDefinite (ekf.cpp:1059-1067): 0.5 (P + P') evaluated while the shift reads P
            if (a.symFirst && si >= 0 && sj >= 0 && si != sj) v = 0.5 * (v + P[sj + (size_t)si * N]);
            if (i == j && i >= EKF_CAM && i < EKF_CAM + EKF_POSE) v += (i - EKF_CAM) < 3 ? a.augNoisePos : a.augNoiseOri;
            PB[idx] = v;
        }
    } else {
        for (int i = tid; i < N; i += EK2_NT) s_m[i] = a.b.m[i];
        ek2_copy8(PB, P + (size_t)J0 * N, N * Bc, tid);          // whole columns: one contiguous block
    }

    // ---- measurement model into shared memory (ld = n)
    double hspeed = 0.0;
    if (a.op == EKF_OP_DENSE) {
        ek2_copy8(X, a.H, n * l, tid);
        __syncthreads();
    } else {
        for (int i = tid; i < n * l; i += EK2_NT) X[i] = 0.0;
        __syncthreads();                                                  // s_m staged, X zeroed
        if (a.op == EKF_OP_PSEUDO_VELOCITY) {
            hspeed = sqrt(s_m[EKF_VEL] * s_m[EKF_VEL] + s_m[EKF_VEL + 1] * s_m[EKF_VEL + 1]);
            if (hspeed <= 1e-7) return;                                   // ekf.cpp:635-637 (uniform; nothing exposed yet)
        }
        if (tid == 0) {
            switch (a.op) {
                case EKF_OP_ZUPT: for (int i = 0; i < 3; i++) X[i + (EKF_VEL + i) * n] = 1.0; break;
                case EKF_OP_ZRUPT: for (int i = 0; i < 3; i++) X[i + (EKF_BGA + i) * n] = 1.0; break;
                case EKF_OP_PSEUDO_VELOCITY: for (int i = 0; i < 2; i++) X[(EKF_VEL + i) * n] = s_m[EKF_VEL + i] / hspeed; break;
                case EKF_OP_POSITION: for (int i = 0; i < 3; i++) X[i + (EKF_POS + i) * n] = 1.0; break;
                case EKF_OP_ZERO_HEIGHT: X[(EKF_POS + 2) * n] = 1.0; break;
                case EKF_OP_ORIENTATION: for (int i = 0; i < 4; i++) X[i + (EKF_ORI + i) * n] = 1.0; break;
                case EKF_OP_AUGMENT:
                    for (int i = 0; i < 3; i++) { X[i + (EKF_POS + i) * n] = 1.0; X[i + (EKF_CAM + i) * n] = -1.0; }
                    for (int i = 0; i < 4; i++) { X[3 + i + (EKF_ORI + i) * n] = 1.0; X[3 + i + (EKF_CAM + 3 + i) * n] = -1.0; }
                    break;
            }
        }
        __syncthreads();
    }
    const double* Hs = X;

    // residual (identical in every CTA)
    for (int i = tid; i < n; i += EK2_NT) {
        double v;
        if (a.op == EKF_OP_PSEUDO_VELOCITY) v = a.defaultSpeed - hspeed;
        else {
            const double yi = a.y ? a.y[i] : a.ysmall[i];
            double fi = 0.0;
            if (a.f) fi = a.f[i];
            else for (int k = 0; k < l; k++) fi += Hs[i + (size_t)k * n] * s_m[k];
            v = yi - fi;
        }
        T[(size_t)i * W + vcol] = v;
    }
    if (joseph) for (int t = tid; t < n * n; t += EK2_NT) T[(size_t)(t / n) * W + vcol + 1 + (t % n)] = (t / n == t % n) ? 1.0 : 0.0;
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

    EK2_PHASE(1);
    // ---- phase A: HP[:, J_c] = H P[0:l, J_c]  (4 x 4 register tiles out of shared memory; the k range is split over KS
    // thread groups whose partial tiles are summed in shared memory in a fixed order)
    {
        const int tm = (n + 3) >> 2, tn = (Bc + 3) >> 2, ntile = tm * tn;
        int KS = min(EK2_NT / max(ntile, 1), l / 24); KS = KS < 1 ? 1 : (KS > 8 ? 8 : KS);   // a slice is worth >= 24 k's
        const int klen = (l + KS - 1) / KS;
        for (int t = tid; t < n * Bc; t += EK2_NT) T[(size_t)(t / Bc) * W + n + (t % Bc)] = 0.0;
        __syncthreads();
        const int grp = tid / max(ntile, 1), t = tid - grp * ntile;
        const int ti = t % max(tm, 1), tj = t / max(tm, 1);
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) acc[x][y] = 0.0;
        if (grp < KS && ntile > 0) {
            const int k0 = grp * klen, k1 = min(l, k0 + klen);
            int iv[4], jv[4];
#pragma unroll
            for (int x = 0; x < 4; x++) { iv[x] = min(ti + x * tm, n - 1); jv[x] = min(tj + x * tn, Bc - 1); }
            for (int k = k0; k < k1; k++) {
                double hv[4], bv[4];
#pragma unroll
                for (int x = 0; x < 4; x++) { hv[x] = Hs[iv[x] + (size_t)k * n]; bv[x] = PB[k + (size_t)jv[x] * N]; }
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y] += hv[x] * bv[y];
            }
        }
        for (int ks = 0; ks < KS; ks++) {
            if (grp == ks && ntile > 0) {
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
    EK2_PHASE(2);
    // ---- phase B: partial S over the own columns inside [0, l) into the S part of the own tableau
    {
        const int kc = max(0, min(Bc, l - J0));
        const int ti_n = (n + 1) >> 1, tp_n = (n + 3) >> 2;
        for (int t = tid; t < ti_n * tp_n; t += EK2_NT) {
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
                if (ip < n) { T[(size_t)i0 * W + ip] = acc[0][x]; if (ti + ti_n < n) T[(size_t)(ti + ti_n) * W + ip] = acc[1][x]; }
            }
        }
    }
    EK2_PHASE(3);
    cluster.sync();                                   // #1: every partial S is in place (and from here on shared memory is exposed)
    // ---- reduce S through distributed shared memory, fixed order r = 0 .. C-1 (+ R on the diagonal)
    if (oneStage) {
        for (int e = tid; e < n * n; e += EK2_NT) {
            const int i = e / n, ip = e - i * n;
            double s = 0.0;
            for (int r = 0; r < C; r++) s += cluster.map_shared_rank(T, r)[(size_t)i * W + ip];
            if (i == ip) s += a.Rdiag;
            RS[e] = s;
        }
        __syncthreads();
    } else {
        const int E = (n * n + C - 1) / C, e0 = c * E, e1 = min(n * n, e0 + E);
        for (int e = e0 + tid; e < e1; e += EK2_NT) {
            const int i = e / n, ip = e - i * n;
            double s = 0.0;
            for (int r = 0; r < C; r++) s += cluster.map_shared_rank(T, r)[(size_t)i * W + ip];
            if (i == ip) s += a.Rdiag;
            RS[e - e0] = s;
        }
        cluster.sync();                               // #2: all slices reduced; nobody reads the partials any more
        for (int e = tid; e < n * n; e += EK2_NT) {
            const int r = e / E;
            T[(size_t)(e / n) * W + (e % n)] = cluster.map_shared_rank(RS, r)[e - r * E];
        }
        __syncthreads();
    }

    EK2_PHASE(4);
    // ---- unpivoted forward elimination of [S | HP_Jc | v | (I)], then Z = D^-1/2 (.)   (ekf_elim.cuh)
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
                    double v = 0.0;
                    if (i < n && j <= cend) v = (oneStage && j < n) ? RS[i * n + j] : T[(size_t)i * W + j];
                    t[aa][sr][bb] = v;
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
        // oversized rows (n > 96 or a very wide block): in place in shared memory, one pivot per barrier (two-stage S only)
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
    if (bad) {                                        // uniform over the cluster
        if (c == 0 && tid == 0) { a.b.res[0] = 1.0; a.b.res[1] = 0.0; a.b.res[2] = 1.0; }
        cluster.sync();
        return;
    }
    __syncthreads();
    EK2_PHASE(5);
    if (tid == 0) { double t = 0.0; for (int k = 0; k < n; k++) { const double z = T[(size_t)k * W + vcol]; t += z * z; } s_scalar[1] = a.noiseScale * t; }
    __syncthreads();
    const double chi2 = s_scalar[1];
    if (checking) {
        const bool outlier = !a.skipChi2 && chi2 > a.chi2Thr;
        if (c == 0 && tid == 0) { a.b.res[0] = outlier ? 3.0 : 0.0; a.b.res[1] = chi2; a.b.res[2] = 0.0; }
        if (outlier || a.mode == EKF_MODE_CHECK) { cluster.sync(); return; }
    } else if (c == 0 && tid == 0) { a.b.res[0] = 0.0; a.b.res[1] = chi2; a.b.res[2] = 0.0; }

    EK2_PHASE(6);
    // ---- gather Z (n x N, row-major) out of the neighbours' tableaus, then P[:, J_c] -= Z' Z[:, J_c] in shared memory
    cluster.sync();                                   // #3: every Z slice is final
    double* Z = X;                                    // H is dead
    for (int r = 0; r < C; r++) {
        const int j0r = r * B, bcr = max(0, min(B, N - j0r));
        const double* Tr = cluster.map_shared_rank(T, r);
        for (int t = tid; t < n * bcr; t += EK2_NT) {
            const int jj = t % bcr, k = t / bcr;
            Z[(size_t)k * N + j0r + jj] = Tr[(size_t)k * W + n + jj];
        }
    }
    __syncthreads();
    EK2_PHASE(7);
    {
        // 4 x 2 register tiles; a thread's four rows are ti, ti + R4, ti + 2 R4, ti + 3 R4 so that the lanes of a warp read
        // CONSECUTIVE doubles of a Z row (conflict-free) and update consecutive rows of a P column
        const int R4 = (N + 3) >> 2, tj_n = (Bc + 1) >> 1;
        for (int t = tid; t < R4 * tj_n; t += EK2_NT) {
            const int ti = t % R4, tj = t / R4;
            const int jj0 = tj * 2, j0 = J0 + jj0;
            const bool j1ok = jj0 + 1 < Bc;
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
                if (i < N) { PB[i + (size_t)jj0 * N] -= acc[x][0]; if (j1ok) PB[i + (size_t)(jj0 + 1) * N] -= acc[x][1]; }
            }
        }
    }
    // state mean: m += Z' z_v (CTA 0 owns the write-back; quaternion normalisation: updateCommon normalises the current
    // orientation only, the visual update and the augmentation all of them, ekf.cpp:31, 843, 874)
    if (c == 0) {
        for (int i = tid; i < N; i += EK2_NT) {
            double s = 0.0;
            for (int k = 0; k < n; k++) s += Z[(size_t)k * N + i] * T[(size_t)k * W + vcol];
            s_m[i] += s;
        }
        __syncthreads();
        for (int q = tid; q < (a.normalizeAll ? a.b.trail + 1 : 1); q += EK2_NT)
            ek2_normalize_quat(q == 0 ? s_m + EKF_ORI : s_m + EKF_CAM + EKF_POSE * (q - 1) + 3);
        __syncthreads();
        for (int i = tid; i < N; i += EK2_NT) a.b.m[i] = s_m[i];
    } else __syncthreads();
    EK2_PHASE(8);

    double* Pblk = PB;                                // block holding this CTA's final columns
    if (joseph) {
        // ---- Joseph form (ekf.cpp:35-50): P'' = G T1' + K R K', G = T1 P' = P' - Z'Z (in PB now), T1 = I - K visAugH
        double* Ks = EX;                              // N x 7
        double* T1c = Ks + (size_t)N * EKF_POSE;      // N x 14
        double* GS = T1c + (size_t)N * 14;            // N x 14: the special columns of G
        double* P2 = GS + (size_t)N * 14;             // N x B: P'' block
        for (int t = tid; t < N * EKF_POSE; t += EK2_NT) {
            const int i = t % N, r = t / N;
            double s = 0.0;
            for (int k = 0; k < n; k++) s += Z[(size_t)k * N + i] * T[(size_t)k * W + vcol + 1 + r];
            Ks[t] = s;
        }
        __syncthreads();
        for (int t = tid; t < N * 14; t += EK2_NT) {
            const int j = t % N, cc = t / N;
            const double kv = cc < 7 ? -Ks[j + cc * N] : Ks[j + (cc - 7) * N];
            T1c[t] = (j == ek2_special_col(cc) ? 1.0 : 0.0) + kv;
        }
        cluster.sync();                               // #4: all of G is final
        for (int t = tid; t < N * 14; t += EK2_NT) {
            const int i = t % N, cc = t / N;
            const int col = ek2_special_col(cc), r = col / B;
            GS[t] = cluster.map_shared_rank(PB, r)[i + (size_t)(col - r * B) * N];
        }
        __syncthreads();
        for (int idx = tid; idx < N * Bc; idx += EK2_NT) {
            const int i = idx % N, j = J0 + idx / N;
            const bool jsp = j < 3 || (j >= EKF_ORI && j < EKF_ORI + 4) || (j >= EKF_CAM && j < EKF_CAM + EKF_POSE);
            double s = jsp ? 0.0 : PB[idx];
#pragma unroll
            for (int cc = 0; cc < 14; cc++) s += GS[i + cc * N] * T1c[j + cc * N];
            double kr = 0.0;
#pragma unroll
            for (int r = 0; r < EKF_POSE; r++) kr += Ks[i + r * N] * (a.Rdiag * Ks[j + r * N]);
            P2[idx] = s + kr;
        }
        Pblk = P2;
    }
    if (a.symmetrize) {
        cluster.sync();                               // #5: every final block is in shared memory
        for (int idx = tid; idx < N * Bc; idx += EK2_NT) {
            const int i = idx % N, j = J0 + idx / N;
            double v = Pblk[idx];
            if (i != j) {
                const int r = i / B;                  // owner of column i, which holds P(j, i)
                const double w = cluster.map_shared_rank(Pblk, r)[j + (size_t)(i - r * B) * N];
                v = i > j ? 0.5 * (v + w) : 0.5 * (w + v);      // same operand order as P(i>j) + P(j<i) on both sides
            }
            P[i + (size_t)j * N] = v;
        }
    } else {
        for (int idx = tid; idx < N * Bc; idx += EK2_NT) P[(size_t)J0 * N + idx] = Pblk[idx];
    }
    EK2_PHASE(9);
    cluster.sync();                                   // nobody may leave while its shared memory can still be read
}
