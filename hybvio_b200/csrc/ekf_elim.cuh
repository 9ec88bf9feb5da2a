// hybvio_b200/csrc/ekf_elim.cuh -- register-resident forward elimination of the EKF tableau [S | HP | v | (I)] on one
// CTA of 512 threads (16 warps), two pivots per barrier.
//
// Layout: rows are dealt to warps in PAIRS (rows 2p, 2p+1 -> warp p mod 16, slot p div 16), columns cyclically to
// lanes (column j -> lane j mod 32, block j div 32); a thread keeps its <= 3 x 2 x 4 elements in registers for the
// whole factorisation. The warp owning the pivot pair (k0, k1) eliminates k0 from row k1 locally (shuffles only),
// publishes both rows and 1/pivots through shared memory (double-buffered), ONE barrier, then every thread applies
// the rank-2 update to its registers. S is symmetric, so the multipliers of row i are read from the published rows
// (S[k0,i] / p0 and S'[k1,i] / p1): the strictly lower triangle of S is never needed and never updated.
//
// Look-ahead: see elim_group.
//
// Why this shape: on B200 a dependent DFMA costs 24 cycles, an fp64 division ~70-90, a publish-barrier-read handshake
// ~100, and at most ~45 DFMA/clk/SM issue (tools/ubench.cu); profiling the one-pivot-per-barrier version showed 60 % of
// all warp samples waiting at the barrier for the pivot owner's serial chain (profiles/), i.e. the factorisation is
// bound by that chain, not by arithmetic. Pairing halves the number of chains and barriers.
//
// Compile-time specialisation: NRA = ceil(n / 32) row-pair slots, NCJ column blocks, and per group of 16 pairs the
// slot KA (which also fixes the column block of the pivots, KA) -- every register index is a constant.
#pragma once
#define ELIM_RA 3          // max row-pair slots per thread (16 warps x 2 rows x 3 -> n <= 96)
#define ELIM_CJ 4          // max column blocks per thread (32 lanes -> row length <= 128)
#define ELIM_ROWBUF (ELIM_CJ * 32)

// shared scratch (doubles): rows[2 buffers][2 rows][ELIM_ROWBUF] | piv[2 buffers][4] | pivots[96]
#define ELIM_SMEM_DOUBLES (2 * 2 * ELIM_ROWBUF + 2 * 4 + ELIM_RA * 32)

// Owner-side work for pivot pair (k0, k0+1) held in t[KA][0..1][*] of the calling warp: eliminate k0 from row k1
// locally, publish both rows, the pivots and their reciprocals into buffer (pp & 1).
template <int NCJ, int KA>
__device__ __forceinline__ void elim_publish(double (&t)[ELIM_RA][2][ELIM_CJ], int n, int lane, int pp, double* sh)
{
    constexpr int KB = KA;
    const int k0 = 32 * KA + 2 * pp, k1 = k0 + 1;
    const bool has1 = k1 < n;
    double* row0 = sh + (pp & 1) * (2 * ELIM_ROWBUF);
    double* row1 = row0 + ELIM_ROWBUF;
    double* pv = sh + 2 * 2 * ELIM_ROWBUF + (pp & 1) * 4;
    double* s_pivots = sh + 2 * 2 * ELIM_ROWBUF + 8;
    const double p0 = __shfl_sync(0xffffffffu, t[KA][0][KB], k0 & 31);
    const double rinv0 = 1.0 / p0;
    const double f = __shfl_sync(0xffffffffu, t[KA][0][KB], k1 & 31) * rinv0;   // S[k0,k1] / p0
#pragma unroll
    for (int bb = KB; bb < NCJ; bb++) {
        row0[lane + 32 * bb] = t[KA][0][bb];
        if (bb > KB || lane + 32 * KB > k0) t[KA][1][bb] -= f * t[KA][0][bb];
        row1[lane + 32 * bb] = t[KA][1][bb];
    }
    const double p1 = has1 ? __shfl_sync(0xffffffffu, t[KA][1][KB], k1 & 31) : 1.0;
    const double rinv1 = 1.0 / p1;                    // all lanes: no divergent section on the critical chain
    if (lane == 0) { pv[0] = p0; pv[1] = rinv0; pv[2] = p1; pv[3] = rinv1; s_pivots[k0] = p0; if (has1) s_pivots[k1] = p1; }
}

// One group of 16 pivot pairs. The barrier of pair pp separates "rows of pair pp are published" from "everyone
// applies them"; the warp that owns pair pp+1 applies them to ITS pivot rows first and immediately runs the
// owner-side chain of pair pp+1 (look-ahead), so that chain overlaps with the other warps' bulk update instead of
// preceding it.
template <int NRA, int NCJ, int KA>
__device__ __forceinline__ bool elim_group(double (&t)[ELIM_RA][2][ELIM_CJ], int n, int lane, int wrp, double* sh)
{
    constexpr int KB = KA;                            // pivots 32 KA .. 32 KA + 31 live in column block KA
    double* s_piv = sh + 2 * 2 * ELIM_ROWBUF;
    if (32 * KA >= n) return true;
    if (wrp == 0) elim_publish<NCJ, KA>(t, n, lane, 0, sh);          // first pair of the group: no look-ahead possible
#pragma unroll 1
    for (int pp = 0; pp < 16; pp++) {
        const int k0 = 32 * KA + 2 * pp, k1 = k0 + 1;
        if (k0 >= n) return true;
        const double* row0 = sh + (pp & 1) * (2 * ELIM_ROWBUF);
        const double* row1 = row0 + ELIM_ROWBUF;
        const double* pv = s_piv + (pp & 1) * 4;
        __syncthreads();
        const double p0 = pv[0], rinv0 = pv[1], p1 = pv[2], rinv1 = pv[3];
        if (!(p0 > 0.0) || !(p1 > 0.0)) return false;
        double rb0[NCJ], rb1[NCJ];
#pragma unroll
        for (int bb = KB; bb < NCJ; bb++) { rb0[bb] = row0[lane + 32 * bb]; rb1[bb] = row1[lane + 32 * bb]; }
        const bool diagCol = lane + 32 * KB > k1;     // per lane: column of block KB right of both pivots
        const bool nextOwner = wrp == pp + 1 && pp + 1 < 16 && k0 + 2 < n;   // warp-uniform
#pragma unroll
        for (int aa = KA; aa < NRA; aa++) {
            const bool later = aa > KA || wrp > pp;   // warp-uniform: this thread's row pair comes after the pivot pair
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int i = 2 * (wrp + 16 * aa) + s;
                const bool act = later && i < n;
                const double f = act ? row0[i] * rinv0 : 0.0;     // symmetric S: S[k0,i] / p0
                const double g = act ? row1[i] * rinv1 : 0.0;     // S'[k1,i] / p1 (row k1 is already reduced by k0)
                if (diagCol) t[aa][s][KB] = fma(-g, rb1[KB], fma(-f, rb0[KB], t[aa][s][KB]));
#pragma unroll
                for (int bb = KB + 1; bb < NCJ; bb++) t[aa][s][bb] = fma(-g, rb1[bb], fma(-f, rb0[bb], t[aa][s][bb]));
            }
            // look-ahead: my slot-KA rows are the next pivot pair and are now up to date -> publish them right away
            if (aa == KA && nextOwner) elim_publish<NCJ, KA>(t, n, lane, pp + 1, sh);
        }
    }
    return true;
}

template <int NRA, int NCJ>
__device__ __forceinline__ bool elim_all(double (&t)[ELIM_RA][2][ELIM_CJ], int n, int lane, int wrp, double* sh)
{
    bool ok = elim_group<NRA, NCJ, 0>(t, n, lane, wrp, sh);
    if (NRA > 1) ok = ok && elim_group<NRA, (NCJ > 1 ? NCJ : 2), (NRA > 1 ? 1 : 0)>(t, n, lane, wrp, sh);
    if (NRA > 2) ok = ok && elim_group<NRA, (NCJ > 2 ? NCJ : 3), (NRA > 2 ? 2 : 0)>(t, n, lane, wrp, sh);
    return ok;
}

// Runtime dispatch on n rows and ncols columns. Returns false on a non-positive pivot (uniform over the CTA).
// Requires n <= 96, ncols <= 128 and ncols > n (so that NCJ >= NRA).
__device__ __forceinline__ bool elim_dispatch(double (&t)[ELIM_RA][2][ELIM_CJ], int n, int ncols, int lane, int wrp, double* sh)
{
    const int nra = (n + 31) >> 5, ncj = (ncols + 31) >> 5;
    if (nra == 1) return ncj <= 1 ? elim_all<1, 1>(t, n, lane, wrp, sh) : ncj == 2 ? elim_all<1, 2>(t, n, lane, wrp, sh) : elim_all<1, 4>(t, n, lane, wrp, sh);
    if (nra == 2) return ncj <= 2 ? elim_all<2, 2>(t, n, lane, wrp, sh) : ncj == 3 ? elim_all<2, 3>(t, n, lane, wrp, sh) : elim_all<2, 4>(t, n, lane, wrp, sh);
    return ncj <= 3 ? elim_all<3, 3>(t, n, lane, wrp, sh) : elim_all<3, 4>(t, n, lane, wrp, sh);
}

// tableau row of (warp, slot, sub)
__device__ __forceinline__ int elim_row(int wrp, int aa, int s) { return 2 * (wrp + 16 * aa) + s; }
