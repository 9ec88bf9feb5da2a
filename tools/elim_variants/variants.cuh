// tools/elim_variants/variants.cuh -- experimental versions of the blocked elimination of ekf_cluster2.cuh (EK2_VARIANT = 1..4),
// substituted through EK2_ELIM_OVERRIDE by tools/ubench_elim2.cu. Measurements: profiles/r02_elimination_variants.md.
#if EK2_VARIANT == 1
#define EK2_EB 8
#define EK2_LINV_DOUBLES 128
// 2^-e for x = f 2^e, f in [1, 2): exact scaling factor out of the exponent field (x > 0, normal)
__device__ __forceinline__ double ek2_pow2_inv(double x)
{
#ifdef HV_EMU
    int e; frexp(x, &e); return ldexp(1.0, -(e - 1));
#else
    return __hiloint2double(0x7fe00000 - (__double2hiint(x) & 0x7ff00000), 0);
#endif
}
// 1 / sqrt(x) to double precision, straight-line: hardware approximation (2^-22) + 2 Newton steps
__device__ __forceinline__ double ek2_rsqrt(double x)
{
#ifdef HV_EMU
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
#endif
}


// 8 x 8, division-free row operations, shuffles (lanes 0..7: D columns, 8..15: I columns)
__device__ __forceinline__ void ek2_diag_factor(const double* T, int W, int r0, int nb, int lane, double* linv, volatile int* s_bad)
{
    double v[8];
    const int cidx = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double t = T[(size_t)(r0 + min(i, nb - 1)) * W + r0 + min(cidx, nb - 1)];
        double x = (i == cidx) ? 1.0 : 0.0;
        if (lane < 8 && i < nb && cidx < nb) x = t;
        if (lane >= 16) x = 0.0;
        v[i] = x;
    }
    bool ok = true;
    double G = 1.0, mine = 1.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const double akk = __shfl_sync(0xffffffffu, v[k], k);
        double raw[8];
#pragma unroll
        for (int i = k + 1; i < 8; i++) raw[i] = __shfl_sync(0xffffffffu, v[k], i);
        if (!(akk > 0.0)) ok = false;
        const double s = ek2_pow2_inv(akk);
        const double ps = akk * s, vks = v[k] * s;
        if (cidx == k) mine = G * akk;
        G *= ps;
#pragma unroll
        for (int i = k + 1; i < 8; i++) v[i] = fma(ps, v[i], -(raw[i] * vks));
    }
    const double sc = ok ? ek2_rsqrt(mine) : 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] *= __shfl_sync(0xffffffffu, sc, k);
    if (lane >= 8 && lane < 16) {
#pragma unroll
        for (int i = 0; i < 8; i++) linv[i * 8 + (lane - 8)] = v[i];
    }
    if (!ok && lane == 0) *s_bad = 1;
}
// Blocked forward elimination of the tableau T = [ S | Y ] (n rows, columns 0 .. ncols-1, row-major, ld W) in shared
// memory: S = L L' (never pivoted: R > 0 makes S positive definite), Y <- L^-1 Y, by 8-row blocks j:
//   a. all warps: rows of block j <- L_jj^-1 * rows (8 x 8 x 8 DMMA per column tile);
//   b. trailing update T[i, c] -= U_j[:, i]' U_j[:, c] for the rows below, upper triangle of S and all of Y (8 x 8 x 8 DMMA
//      per tile, several tiles of a warp in flight) by warps 1..15, WHILE warp 0 updates the next diagonal tile first and
//      factors it (ek2_diag_factor: look-ahead), so that the serial pivot chain overlaps the bulk work.
// Two barriers per 8 pivots; the first generation (ekf_elim.cuh) needed one barrier per two pivots and kept the tableau
// in registers, which bounded n <= 96. Returns false (uniformly) on a non-positive pivot. s_linv: 2 x 64 doubles.
__device__ __forceinline__ bool ek2_block_eliminate(double* T, int W, int n, int ncols, int wrp, int lane, double* s_linv, volatile int* s_bad)
{
    const int g8 = lane >> 2, t4 = lane & 3;
    const int nwarps = EK2_NT / 32;
    const int MT = (n + 7) >> 3, CT = (ncols + 7) >> 3;
    EK2_ELIM_DECL
    if (wrp == 0) {
        if (lane == 0) *s_bad = 0;
        __syncwarp();
        ek2_diag_factor(T, W, 0, min(8, n), lane, s_linv, s_bad);
    }
    __syncthreads();
    for (int j = 0; j < MT; j++) {
        if (*s_bad) return false;
        const int r0 = 8 * j, nb = min(8, n - r0);
        const double* linv = s_linv + (j & 1) * 64;
        EK2_ELIM_MARK(0);
        // ---- a. rows of the block <- L_jj^-1 * rows, column tiles j .. CT-1 (loads clamped into the tableau: no branches)
        for (int ct = j + wrp; ct < CT; ct += nwarps) {
            double c0 = 0.0, c1 = 0.0;
            const int colbc = min(8 * ct + g8, ncols - 1);
            double bf[2];
#pragma unroll
            for (int kt = 0; kt < 2; kt++) {
                const int k = kt * 4 + t4;
                const double x = T[(size_t)(r0 + min(k, nb - 1)) * W + colbc];
                bf[kt] = k < nb ? x : 0.0;
            }
#pragma unroll
            for (int kt = 0; kt < 2; kt++) hv_dmma(c0, c1, linv[g8 * 8 + kt * 4 + t4], bf[kt]);
            const int col = 8 * ct + 2 * t4;
            if (g8 < nb) { if (col < ncols) T[(size_t)(r0 + g8) * W + col] = c0; if (col + 1 < ncols) T[(size_t)(r0 + g8) * W + col + 1] = c1; }
        }
        EK2_ELIM_MARK(1);
        __syncthreads();
        EK2_ELIM_MARK(2);
        // ---- b. trailing update: row tiles mt > j, column tiles nt >= mt (row-major list; entry 0 is the next diagonal tile).
        // Warp 0 takes entry 0 and then factors it; warps 1..15 walk contiguous ranges of the rest, reloading the A
        // fragment (-U_j[:, row tile]') only when the row tile changes.
        if (j + 1 < MT) {
            const int first = j + 1;
            int total = 0;
            for (int mt = first; mt < MT; mt++) total += CT - mt;
            const bool ahead = wrp == 0;
            // Workers: the warps that do NOT share warp 0's scheduler / FP64 pipe (warp id mod 4 != 0). The pivot chain of the
            // look-ahead factorisation is a sequence of dependent fp64 operations; every DMMA a sibling warp queues on the
            // same pipe (16 cycles each) would sit in front of them.
            const int nwork = nwarps - nwarps / 4, widx = wrp - wrp / 4 - 1;           // 12 workers, index 0..11
            const bool worker = (wrp & 3) != 0;
            const int rest = total - 1;
            int lo = ahead ? 0 : worker ? 1 + (int)(((long long)rest * widx) / nwork) : 0;
            const int hi = ahead ? 1 : worker ? 1 + (int)(((long long)rest * (widx + 1)) / nwork) : 0;
            const double* rowk0 = T + (size_t)(r0 + min(t4, nb - 1)) * W;        // k = t4
            const double* rowk1 = T + (size_t)(r0 + min(4 + t4, nb - 1)) * W;    // k = 4 + t4
            const bool k0v = t4 < nb, k1v = 4 + t4 < nb;
            if (ahead) {
                // the next diagonal tile (first, first): A and B fragments are the same column block of U_j
                const int cb = 8 * first, am = min(cb + g8, ncols - 1);
                const double x0 = rowk0[am], x1 = rowk1[am];
                const int rowi = cb + g8, coli = cb + 2 * t4;
                double* crow = T + (size_t)min(rowi, n - 1) * W;
                double c0 = crow[min(coli, ncols - 1)], c1 = crow[min(coli + 1, ncols - 1)];
                hv_dmma(c0, c1, k0v ? -x0 : 0.0, x0);
                hv_dmma(c0, c1, k1v ? -x1 : 0.0, x1);
                if (rowi < n) { if (coli < ncols) crow[coli] = c0; if (coli + 1 < ncols) crow[coli + 1] = c1; }
            } else if (worker) {
                int mt = first, nt, idx = min(lo, total - 1);
                while (idx >= CT - mt) { idx -= CT - mt; mt++; }
                nt = mt + idx;
                int curMt = -1;
                double a0 = 0.0, a1 = 0.0;
                for (; lo < hi; lo++) {
                    if (mt != curMt) {
                        const int am = min(8 * mt + g8, ncols - 1);
                        const double x0 = rowk0[am], x1 = rowk1[am];
                        a0 = k0v ? -x0 : 0.0; a1 = k1v ? -x1 : 0.0;                   // A[m][k] = -U_j[k][8 mt + m]
                        curMt = mt;
                    }
                    const int rowi = 8 * mt + g8, coli = 8 * nt + 2 * t4, bn = min(8 * nt + g8, ncols - 1);
                    double* crow = T + (size_t)min(rowi, n - 1) * W;
                    double c0 = crow[min(coli, ncols - 1)], c1 = crow[min(coli + 1, ncols - 1)];
                    const double b0 = rowk0[bn], b1 = rowk1[bn];                       // B[k][nn] = U_j[k][8 nt + nn]
                    hv_dmma(c0, c1, a0, b0);
                    hv_dmma(c0, c1, a1, b1);
                    if (rowi < n) { if (coli < ncols) crow[coli] = c0; if (coli + 1 < ncols) crow[coli + 1] = c1; }
                    if (++nt == CT) { mt++; nt = mt; }
                }
            }
            EK2_ELIM_MARK(3);
            if (ahead) {
                __syncwarp();
                ek2_diag_factor(T, W, 8 * first, min(8, n - 8 * first), lane, s_linv + (first & 1) * 64, s_bad);
            }
            EK2_ELIM_MARK(4);
            __syncthreads();
            EK2_ELIM_MARK(5);
        }
    }
    return !*s_bad;
}

#endif
#if EK2_VARIANT == 2
#define EK2_EB 8
#define EK2_LINV_DOUBLES (128 + 16)
// 2^-e for x = f 2^e, f in [1, 2): exact scaling factor out of the exponent field (x > 0, normal)
__device__ __forceinline__ double ek2_pow2_inv(double x)
{
#ifdef HV_EMU
    int e; frexp(x, &e); return ldexp(1.0, -(e - 1));
#else
    return __hiloint2double(0x7fe00000 - (__double2hiint(x) & 0x7ff00000), 0);
#endif
}
// 1 / sqrt(x) to double precision, straight-line: hardware approximation (2^-22) + 2 Newton steps
__device__ __forceinline__ double ek2_rsqrt(double x)
{
#ifdef HV_EMU
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
#endif
}


// 8 x 8, division-free, row k through shared memory (scr: 2 x 8 doubles)
__device__ __forceinline__ void ek2_diag_factor(const double* T, int W, int r0, int nb, int lane, double* linv, volatile int* s_bad, double* scr)
{
    double v[8];
    const int cidx = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double t = T[(size_t)(r0 + min(i, nb - 1)) * W + r0 + min(cidx, nb - 1)];
        double x = (i == cidx) ? 1.0 : 0.0;
        if (lane < 8 && i < nb && cidx < nb) x = t;
        if (lane >= 16) x = 0.0;
        v[i] = x;
    }
    bool ok = true;
    double G = 1.0, mine = 1.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double* buf = scr + (k & 1) * 8;
        if (lane < 8) buf[cidx] = v[k];
        __syncwarp();
        double raw[8];
#pragma unroll
        for (int p = (k >> 1); p < 4; p++) { const double2 q = *reinterpret_cast<const double2*>(buf + 2 * p); raw[2 * p] = q.x; raw[2 * p + 1] = q.y; }
        const double akk = raw[k];
        if (!(akk > 0.0)) ok = false;
        const double s = ek2_pow2_inv(akk);
        const double ps = akk * s, vks = v[k] * s;
        if (cidx == k) mine = G * akk;
        G *= ps;
#pragma unroll
        for (int i = k + 1; i < 8; i++) v[i] = fma(ps, v[i], -(raw[i] * vks));
    }
    const double sc = ok ? ek2_rsqrt(mine) : 0.0;
    __syncwarp();
    if (lane < 8) scr[cidx] = sc;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] *= scr[k];
    if (lane >= 8 && lane < 16) {
#pragma unroll
        for (int i = 0; i < 8; i++) linv[i * 8 + (lane - 8)] = v[i];
    }
    if (!ok && lane == 0) *s_bad = 1;
    __syncwarp();
}
// Blocked forward elimination of the tableau T = [ S | Y ] (n rows, columns 0 .. ncols-1, row-major, ld W) in shared
// memory: S = L L' (never pivoted: R > 0 makes S positive definite), Y <- L^-1 Y, by 8-row blocks j:
//   a. all warps: rows of block j <- L_jj^-1 * rows (8 x 8 x 8 DMMA per column tile);
//   b. trailing update T[i, c] -= U_j[:, i]' U_j[:, c] for the rows below, upper triangle of S and all of Y (8 x 8 x 8 DMMA
//      per tile, several tiles of a warp in flight) by warps 1..15, WHILE warp 0 updates the next diagonal tile first and
//      factors it (ek2_diag_factor: look-ahead), so that the serial pivot chain overlaps the bulk work.
// Two barriers per 8 pivots; the first generation (ekf_elim.cuh) needed one barrier per two pivots and kept the tableau
// in registers, which bounded n <= 96. Returns false (uniformly) on a non-positive pivot. s_linv: 2 x 64 doubles.
__device__ __forceinline__ bool ek2_block_eliminate(double* T, int W, int n, int ncols, int wrp, int lane, double* s_linv, volatile int* s_bad)
{
    const int g8 = lane >> 2, t4 = lane & 3;
    const int nwarps = EK2_NT / 32;
    const int MT = (n + 7) >> 3, CT = (ncols + 7) >> 3;
    EK2_ELIM_DECL
    if (wrp == 0) {
        if (lane == 0) *s_bad = 0;
        __syncwarp();
        ek2_diag_factor(T, W, 0, min(8, n), lane, s_linv, s_bad, s_linv + 128);
    }
    __syncthreads();
    for (int j = 0; j < MT; j++) {
        if (*s_bad) return false;
        const int r0 = 8 * j, nb = min(8, n - r0);
        const double* linv = s_linv + (j & 1) * 64;
        EK2_ELIM_MARK(0);
        // ---- a. rows of the block <- L_jj^-1 * rows, column tiles j .. CT-1 (loads clamped into the tableau: no branches)
        for (int ct = j + wrp; ct < CT; ct += nwarps) {
            double c0 = 0.0, c1 = 0.0;
            const int colbc = min(8 * ct + g8, ncols - 1);
            double bf[2];
#pragma unroll
            for (int kt = 0; kt < 2; kt++) {
                const int k = kt * 4 + t4;
                const double x = T[(size_t)(r0 + min(k, nb - 1)) * W + colbc];
                bf[kt] = k < nb ? x : 0.0;
            }
#pragma unroll
            for (int kt = 0; kt < 2; kt++) hv_dmma(c0, c1, linv[g8 * 8 + kt * 4 + t4], bf[kt]);
            const int col = 8 * ct + 2 * t4;
            if (g8 < nb) { if (col < ncols) T[(size_t)(r0 + g8) * W + col] = c0; if (col + 1 < ncols) T[(size_t)(r0 + g8) * W + col + 1] = c1; }
        }
        EK2_ELIM_MARK(1);
        __syncthreads();
        EK2_ELIM_MARK(2);
        // ---- b. trailing update: row tiles mt > j, column tiles nt >= mt (row-major list; entry 0 is the next diagonal tile).
        // Warp 0 takes entry 0 and then factors it; warps 1..15 walk contiguous ranges of the rest, reloading the A
        // fragment (-U_j[:, row tile]') only when the row tile changes.
        if (j + 1 < MT) {
            const int first = j + 1;
            int total = 0;
            for (int mt = first; mt < MT; mt++) total += CT - mt;
            const bool ahead = wrp == 0;
            // Workers: the warps that do NOT share warp 0's scheduler / FP64 pipe (warp id mod 4 != 0). The pivot chain of the
            // look-ahead factorisation is a sequence of dependent fp64 operations; every DMMA a sibling warp queues on the
            // same pipe (16 cycles each) would sit in front of them.
            const int nwork = nwarps - nwarps / 4, widx = wrp - wrp / 4 - 1;           // 12 workers, index 0..11
            const bool worker = (wrp & 3) != 0;
            const int rest = total - 1;
            int lo = ahead ? 0 : worker ? 1 + (int)(((long long)rest * widx) / nwork) : 0;
            const int hi = ahead ? 1 : worker ? 1 + (int)(((long long)rest * (widx + 1)) / nwork) : 0;
            const double* rowk0 = T + (size_t)(r0 + min(t4, nb - 1)) * W;        // k = t4
            const double* rowk1 = T + (size_t)(r0 + min(4 + t4, nb - 1)) * W;    // k = 4 + t4
            const bool k0v = t4 < nb, k1v = 4 + t4 < nb;
            if (ahead) {
                // the next diagonal tile (first, first): A and B fragments are the same column block of U_j
                const int cb = 8 * first, am = min(cb + g8, ncols - 1);
                const double x0 = rowk0[am], x1 = rowk1[am];
                const int rowi = cb + g8, coli = cb + 2 * t4;
                double* crow = T + (size_t)min(rowi, n - 1) * W;
                double c0 = crow[min(coli, ncols - 1)], c1 = crow[min(coli + 1, ncols - 1)];
                hv_dmma(c0, c1, k0v ? -x0 : 0.0, x0);
                hv_dmma(c0, c1, k1v ? -x1 : 0.0, x1);
                if (rowi < n) { if (coli < ncols) crow[coli] = c0; if (coli + 1 < ncols) crow[coli + 1] = c1; }
            } else if (worker) {
                int mt = first, nt, idx = min(lo, total - 1);
                while (idx >= CT - mt) { idx -= CT - mt; mt++; }
                nt = mt + idx;
                int curMt = -1;
                double a0 = 0.0, a1 = 0.0;
                for (; lo < hi; lo++) {
                    if (mt != curMt) {
                        const int am = min(8 * mt + g8, ncols - 1);
                        const double x0 = rowk0[am], x1 = rowk1[am];
                        a0 = k0v ? -x0 : 0.0; a1 = k1v ? -x1 : 0.0;                   // A[m][k] = -U_j[k][8 mt + m]
                        curMt = mt;
                    }
                    const int rowi = 8 * mt + g8, coli = 8 * nt + 2 * t4, bn = min(8 * nt + g8, ncols - 1);
                    double* crow = T + (size_t)min(rowi, n - 1) * W;
                    double c0 = crow[min(coli, ncols - 1)], c1 = crow[min(coli + 1, ncols - 1)];
                    const double b0 = rowk0[bn], b1 = rowk1[bn];                       // B[k][nn] = U_j[k][8 nt + nn]
                    hv_dmma(c0, c1, a0, b0);
                    hv_dmma(c0, c1, a1, b1);
                    if (rowi < n) { if (coli < ncols) crow[coli] = c0; if (coli + 1 < ncols) crow[coli + 1] = c1; }
                    if (++nt == CT) { mt++; nt = mt; }
                }
            }
            EK2_ELIM_MARK(3);
            if (ahead) {
                __syncwarp();
                ek2_diag_factor(T, W, 8 * first, min(8, n - 8 * first), lane, s_linv + (first & 1) * 64, s_bad, s_linv + 128);
            }
            EK2_ELIM_MARK(4);
            __syncthreads();
            EK2_ELIM_MARK(5);
        }
    }
    return !*s_bad;
}

#endif
#if EK2_VARIANT == 3
// ---- blocked forward elimination, 16 pivots per block ---------------------------------------------------------------------------
// The serial part of the update is the pivot chain of S = L L' (n dependent pivots). Measured on B200 (tools/ubench_elim2.cu,
// profiles/r01_ubench_elim2.txt) the first version of this routine -- 8 x 8 blocks, Cholesky row operations with one rsqrt per
// pivot -- spent ~200 cycles per pivot in the chain shuffle -> rsqrt -> multiply -> FMA, plus two barriers, a rows-solve and a
// look-ahead tile per 8 pivots: ~3200 cycles per block, 17 us for n = 84. This version
//   * takes the reciprocal square root OFF the chain: the row operations are division-free, row_i <- s (a_kk row_i - a_ik row_k),
//     with s = 2^-exponent(a_kk) (exact, two integer instructions) keeping the magnitudes in range; every active row carries
//     the same accumulated factor G_k, so the rows of D^-1/2 L^-1 are recovered at the end with ONE rsqrt per row, all 16 of a
//     block in parallel (lane k does row k). Chain per pivot: shuffle -> exponent -> multiply -> multiply -> FMA;
//   * works on 16 x 16 blocks: half the barriers, rows-solves and look-ahead steps per pivot, and every trailing tile gets 4
//     DMMAs per load / store of its accumulators instead of 2.
#define EK2_EB 16
#define EK2_LINV_DOUBLES 512

// 2^-e for x = f 2^e, f in [1, 2): exact scaling factor out of the exponent field (x > 0, normal)
__device__ __forceinline__ double ek2_pow2_inv(double x)
{
#ifdef HV_EMU
    int e; frexp(x, &e); return ldexp(1.0, -(e - 1));
#else
    return __hiloint2double(0x7fe00000 - (__double2hiint(x) & 0x7ff00000), 0);
#endif
}
// 1 / sqrt(x) to double precision, straight-line: hardware approximation (2^-22) + 2 Newton steps
__device__ __forceinline__ double ek2_rsqrt(double x)
{
#ifdef HV_EMU
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
#endif
}

// Factorisation of the 16 x 16 diagonal block D = T[r0 .. r0+nb, r0 .. r0+nb] (upper triangle read) of the current Schur complement
// by ONE warp with shuffles only: lanes 0..15 hold the columns of D (padded with the identity), lanes 16..31 those of I; the row
// operations applied to both leave L_jj' in the first and L_jj^-1 (lower triangular) in the second group, which is written
// to linv (16 x 16, row-major).
__device__ __forceinline__ void ek2_diag_factor(const double* T, int W, int r0, int nb, int lane, double* linv, volatile int* s_bad)
{
    double v[EK2_EB];
    const int cidx = lane & 15;
    const bool left = lane < 16;
#pragma unroll
    for (int i = 0; i < EK2_EB; i++) {
        const int lo = min(i, cidx), hi = max(i, cidx);                  // S is kept on and above the diagonal tiles only
        const double t = T[(size_t)(r0 + min(lo, nb - 1)) * W + r0 + min(hi, nb - 1)];
        v[i] = (left && hi < nb) ? t : (i == cidx ? 1.0 : 0.0);
    }
    bool ok = true;
    double G = 1.0, mine = 1.0;                       // G_k: common factor of the active rows; mine = G_k p_k of row k = cidx
#pragma unroll
    for (int k = 0; k < EK2_EB; k++) {
        const double akk = __shfl_sync(0xffffffffu, v[k], k);
        double raw[EK2_EB];
#pragma unroll
        for (int i = k + 1; i < EK2_EB; i++) raw[i] = __shfl_sync(0xffffffffu, v[k], i);    // a_ik = a_ki: entry i of row k
        if (!(akk > 0.0)) ok = false;
        const double s = ek2_pow2_inv(akk);
        const double ps = akk * s, vks = v[k] * s;    // exact scalings
        if (cidx == k) mine = G * akk;
        G *= ps;
#pragma unroll
        for (int i = k + 1; i < EK2_EB; i++) v[i] = fma(ps, v[i], -(raw[i] * vks));
    }
    // row k of D^-1/2 L^-1 [D | I] = stored row k / sqrt(G_k p_k)
    const double sc = ok ? ek2_rsqrt(mine) : 0.0;
#pragma unroll
    for (int k = 0; k < EK2_EB; k++) v[k] *= __shfl_sync(0xffffffffu, sc, k);
    if (!left) {
#pragma unroll
        for (int i = 0; i < EK2_EB; i++) linv[i * EK2_EB + cidx] = v[i];
    }
    if (!ok && lane == 0) *s_bad = 1;
}

// One trailing tile: C(8 mt .., 8 nt ..) -= U_j[:, 8 mt ..]' U_j[:, 8 nt ..], K = kb rows of the block (4 k-steps of 4)
__device__ __forceinline__ void ek2_trailing_tile(double* T, int W, int n, int ncols, int mt, int nt, const double* const* rowk, const double* a,
                                                  int g8, int t4)
{
    const int rowi = 8 * mt + g8, coli = 8 * nt + 2 * t4, bn = min(8 * nt + g8, ncols - 1);
    double* crow = T + (size_t)min(rowi, n - 1) * W;
    double c0 = crow[min(coli, ncols - 1)], c1 = crow[min(coli + 1, ncols - 1)];
    double b[4];
#pragma unroll
    for (int kt = 0; kt < 4; kt++) b[kt] = rowk[kt][bn];                 // B[k][nn] = U_j[k][8 nt + nn] (rows past the block: times a = 0)
#pragma unroll
    for (int kt = 0; kt < 4; kt++) hv_dmma(c0, c1, a[kt], b[kt]);
    if (rowi < n) { if (coli < ncols) crow[coli] = c0; if (coli + 1 < ncols) crow[coli + 1] = c1; }
}

// Blocked forward elimination of the tableau T = [ S | Y ] (n rows, columns 0 .. ncols-1, row-major, ld W) in shared
// memory: S = L L' (never pivoted: R > 0 makes S positive definite), Y <- L^-1 Y, by 16-row blocks j:
//   a. all warps: rows of block j <- L_jj^-1 * rows (one 16 x 8 column tile per warp and turn, 6 DMMAs);
//   b. trailing update T[i, c] -= U_j[:, i]' U_j[:, c] for the rows below, upper triangle of S and all of Y (8 x 8 x 16 per tile) by
//      the worker warps, WHILE warp 0 updates the three tiles of the next diagonal block first and factors it
//      (ek2_diag_factor: look-ahead), so that the serial pivot chain overlaps the bulk work.
// Two barriers per 16 pivots. Returns false (uniformly) on a non-positive pivot. s_linv: 2 x 256 doubles.
__device__ __forceinline__ bool ek2_block_eliminate(double* T, int W, int n, int ncols, int wrp, int lane, double* s_linv, volatile int* s_bad)
{
    const int g8 = lane >> 2, t4 = lane & 3;
    const int nwarps = EK2_NT / 32;
    const int MB = (n + EK2_EB - 1) / EK2_EB, MT = (n + 7) >> 3, CT = (ncols + 7) >> 3;
    EK2_ELIM_DECL
    if (wrp == 0) {
        if (lane == 0) *s_bad = 0;
        __syncwarp();
        ek2_diag_factor(T, W, 0, min(EK2_EB, n), lane, s_linv, s_bad);
    }
    __syncthreads();
    for (int j = 0; j < MB; j++) {
        if (*s_bad) return false;
        const int r0 = EK2_EB * j, nb = min(EK2_EB, n - r0);
        const double* linv = s_linv + (j & 1) * (EK2_EB * EK2_EB);
        EK2_ELIM_MARK(0);
        // ---- a. rows of the block <- L_jj^-1 * rows, column tiles 2 j .. CT-1 (loads clamped into the tableau: no branches).
        // L_jj^-1 is lower triangular: the upper row tile needs k < 8 only.
        {
            double la[6];                                                 // A fragments: (row tile 0: k-steps 0, 1) (row tile 1: k-steps 0 .. 3)
#pragma unroll
            for (int kt = 0; kt < 2; kt++) la[kt] = linv[g8 * EK2_EB + kt * 4 + t4];
#pragma unroll
            for (int kt = 0; kt < 4; kt++) la[2 + kt] = linv[(8 + g8) * EK2_EB + kt * 4 + t4];
            for (int ct = 2 * j + wrp; ct < CT; ct += nwarps) {
                const int colbc = min(8 * ct + g8, ncols - 1);
                double bf[4];
#pragma unroll
                for (int kt = 0; kt < 4; kt++) {
                    const int k = kt * 4 + t4;
                    const double x = T[(size_t)(r0 + min(k, nb - 1)) * W + colbc];
                    bf[kt] = k < nb ? x : 0.0;
                }
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                hv_dmma(c0, c1, la[0], bf[0]); hv_dmma(d0, d1, la[2], bf[0]);
                hv_dmma(c0, c1, la[1], bf[1]); hv_dmma(d0, d1, la[3], bf[1]);
                if (nb > 8) { hv_dmma(d0, d1, la[4], bf[2]); hv_dmma(d0, d1, la[5], bf[3]); }
                const int col = 8 * ct + 2 * t4;
                if (g8 < nb) { if (col < ncols) T[(size_t)(r0 + g8) * W + col] = c0; if (col + 1 < ncols) T[(size_t)(r0 + g8) * W + col + 1] = c1; }
                if (8 + g8 < nb) { if (col < ncols) T[(size_t)(r0 + 8 + g8) * W + col] = d0; if (col + 1 < ncols) T[(size_t)(r0 + 8 + g8) * W + col + 1] = d1; }
            }
        }
        EK2_ELIM_MARK(1);
        __syncthreads();
        EK2_ELIM_MARK(2);
        // ---- b. trailing update: row tiles mt >= m0 = 2 (j + 1), column tiles nt >= mt. Warp 0 takes the tiles of the next diagonal
        // block, (m0, m0), (m0, m0 + 1), (m0 + 1, m0 + 1), and then factors it; the workers walk contiguous ranges of the rest
        // (row m0: nt >= m0 + 2; row m0 + 1: nt >= m0 + 2; row mt: nt >= mt), reloading the A fragments (-U_j[:, row tile]') only
        // when the row tile changes.
        if (j + 1 < MB) {                                                 // a following block exists: this one is full (nb = 16)
            const int m0 = 2 * (j + 1);
            const double* rowk[4];
#pragma unroll
            for (int kt = 0; kt < 4; kt++) rowk[kt] = T + (size_t)(r0 + kt * 4 + t4) * W;
            const bool ahead = wrp == 0;
            // Workers: the warps that do NOT share warp 0's scheduler / FP64 pipe (warp id mod 4 != 0). The pivot chain of the
            // look-ahead factorisation is a sequence of dependent fp64 operations; every DMMA a sibling warp queues on the
            // same pipe (16 cycles each) would sit in front of them.
            const int nwork = nwarps - nwarps / 4, widx = wrp - wrp / 4 - 1;           // 12 workers, index 0..11
            const bool worker = (wrp & 3) != 0;
            if (ahead) {
                for (int mt = m0; mt <= min(m0 + 1, MT - 1); mt++) {
                    double a[4];
                    const int am = min(8 * mt + g8, ncols - 1);
#pragma unroll
                    for (int kt = 0; kt < 4; kt++) a[kt] = -rowk[kt][am];
                    for (int nt = mt; nt <= min(m0 + 1, CT - 1); nt++) ek2_trailing_tile(T, W, n, ncols, mt, nt, rowk, a, g8, t4);
                }
            } else if (worker) {
                const int skip = m0 + 2;                                  // first column tile of the workers in rows m0, m0 + 1
                auto cnt = [&](int mt) { return max(0, CT - max(mt, skip)); };
                int total = 0;
                for (int mt = m0; mt < MT; mt++) total += cnt(mt);
                int lo = (int)(((long long)total * widx) / nwork);
                const int hi = (int)(((long long)total * (widx + 1)) / nwork);
                int mt = m0, idx = lo;
                while (mt < MT - 1 && idx >= cnt(mt)) { idx -= cnt(mt); mt++; }
                int nt = max(mt, skip) + idx;
                int curMt = -1;
                double a[4] = {0.0, 0.0, 0.0, 0.0};
                for (; lo < hi; lo++) {
                    if (mt != curMt) {
                        const int am = min(8 * mt + g8, ncols - 1);
#pragma unroll
                        for (int kt = 0; kt < 4; kt++) a[kt] = -rowk[kt][am];           // A[m][k] = -U_j[k][8 mt + m]
                        curMt = mt;
                    }
                    ek2_trailing_tile(T, W, n, ncols, mt, nt, rowk, a, g8, t4);
                    if (++nt >= CT) { do mt++; while (mt < MT - 1 && cnt(mt) == 0); nt = max(mt, skip); }
                }
            }
            EK2_ELIM_MARK(3);
            if (ahead) {
                __syncwarp();
                ek2_diag_factor(T, W, EK2_EB * (j + 1), min(EK2_EB, n - EK2_EB * (j + 1)), lane, s_linv + ((j + 1) & 1) * (EK2_EB * EK2_EB), s_bad);
            }
            EK2_ELIM_MARK(4);
            __syncthreads();
            EK2_ELIM_MARK(5);
        }
    }
    return !*s_bad;
}

#endif
#if EK2_VARIANT == 4
#define EK2_EB 16
#define EK2_LINV_DOUBLES (512 + 32)
// 2^-e for x = f 2^e, f in [1, 2): exact scaling factor out of the exponent field (x > 0, normal)
__device__ __forceinline__ double ek2_pow2_inv(double x)
{
#ifdef HV_EMU
    int e; frexp(x, &e); return ldexp(1.0, -(e - 1));
#else
    return __hiloint2double(0x7fe00000 - (__double2hiint(x) & 0x7ff00000), 0);
#endif
}
// 1 / sqrt(x) to double precision, straight-line: hardware approximation (2^-22) + 2 Newton steps
__device__ __forceinline__ double ek2_rsqrt(double x)
{
#ifdef HV_EMU
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
#endif
}


__device__ __forceinline__ void ek2_diag_factor(const double* T, int W, int r0, int nb, int lane, double* linv, volatile int* s_bad, double* scr)
{
    double v[EK2_EB];
    const int cidx = lane & 15;
    const bool left = lane < 16;
#pragma unroll
    for (int i = 0; i < EK2_EB; i++) {
        const int lo = min(i, cidx), hi = max(i, cidx);
        const double t = T[(size_t)(r0 + min(lo, nb - 1)) * W + r0 + min(hi, nb - 1)];
        v[i] = (left && hi < nb) ? t : (i == cidx ? 1.0 : 0.0);
    }
    bool ok = true;
    double G = 1.0, mine = 1.0;
#pragma unroll
    for (int k = 0; k < EK2_EB; k++) {
        double* buf = scr + (k & 1) * EK2_EB;
        if (left) buf[cidx] = v[k];                                      // row k of the D part: entry i from lane i
        __syncwarp();
        double raw[EK2_EB];
#pragma unroll
        for (int p = (k >> 1); p < EK2_EB / 2; p++) { const double2 q = *reinterpret_cast<const double2*>(buf + 2 * p); raw[2 * p] = q.x; raw[2 * p + 1] = q.y; }
        const double akk = raw[k];
        if (!(akk > 0.0)) ok = false;
        const double s = ek2_pow2_inv(akk);
        const double ps = akk * s, vks = v[k] * s;
        if (cidx == k) mine = G * akk;
        G *= ps;
#pragma unroll
        for (int i = k + 1; i < EK2_EB; i++) v[i] = fma(ps, v[i], -(raw[i] * vks));
    }
    const double sc = ok ? ek2_rsqrt(mine) : 0.0;
    __syncwarp();
    if (left) scr[cidx] = sc;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < EK2_EB; k++) v[k] *= scr[k];
    if (!left) {
#pragma unroll
        for (int i = 0; i < EK2_EB; i++) linv[i * EK2_EB + cidx] = v[i];
    }
    if (!ok && lane == 0) *s_bad = 1;
    __syncwarp();
}
// One trailing tile: C(8 mt .., 8 nt ..) -= U_j[:, 8 mt ..]' U_j[:, 8 nt ..], K = kb rows of the block (4 k-steps of 4)
__device__ __forceinline__ void ek2_trailing_tile(double* T, int W, int n, int ncols, int mt, int nt, const double* const* rowk, const double* a,
                                                  int g8, int t4)
{
    const int rowi = 8 * mt + g8, coli = 8 * nt + 2 * t4, bn = min(8 * nt + g8, ncols - 1);
    double* crow = T + (size_t)min(rowi, n - 1) * W;
    double c0 = crow[min(coli, ncols - 1)], c1 = crow[min(coli + 1, ncols - 1)];
    double b[4];
#pragma unroll
    for (int kt = 0; kt < 4; kt++) b[kt] = rowk[kt][bn];                 // B[k][nn] = U_j[k][8 nt + nn] (rows past the block: times a = 0)
#pragma unroll
    for (int kt = 0; kt < 4; kt++) hv_dmma(c0, c1, a[kt], b[kt]);
    if (rowi < n) { if (coli < ncols) crow[coli] = c0; if (coli + 1 < ncols) crow[coli + 1] = c1; }
}

// Blocked forward elimination of the tableau T = [ S | Y ] (n rows, columns 0 .. ncols-1, row-major, ld W) in shared
// memory: S = L L' (never pivoted: R > 0 makes S positive definite), Y <- L^-1 Y, by 16-row blocks j:
//   a. all warps: rows of block j <- L_jj^-1 * rows (one 16 x 8 column tile per warp and turn, 6 DMMAs);
//   b. trailing update T[i, c] -= U_j[:, i]' U_j[:, c] for the rows below, upper triangle of S and all of Y (8 x 8 x 16 per tile) by
//      the worker warps, WHILE warp 0 updates the three tiles of the next diagonal block first and factors it
//      (ek2_diag_factor: look-ahead), so that the serial pivot chain overlaps the bulk work.
// Two barriers per 16 pivots. Returns false (uniformly) on a non-positive pivot. s_linv: 2 x 256 doubles.
__device__ __forceinline__ bool ek2_block_eliminate(double* T, int W, int n, int ncols, int wrp, int lane, double* s_linv, volatile int* s_bad)
{
    const int g8 = lane >> 2, t4 = lane & 3;
    const int nwarps = EK2_NT / 32;
    const int MB = (n + EK2_EB - 1) / EK2_EB, MT = (n + 7) >> 3, CT = (ncols + 7) >> 3;
    EK2_ELIM_DECL
    if (wrp == 0) {
        if (lane == 0) *s_bad = 0;
        __syncwarp();
        ek2_diag_factor(T, W, 0, min(EK2_EB, n), lane, s_linv, s_bad, s_linv + 512);
    }
    __syncthreads();
    for (int j = 0; j < MB; j++) {
        if (*s_bad) return false;
        const int r0 = EK2_EB * j, nb = min(EK2_EB, n - r0);
        const double* linv = s_linv + (j & 1) * (EK2_EB * EK2_EB);
        EK2_ELIM_MARK(0);
        // ---- a. rows of the block <- L_jj^-1 * rows, column tiles 2 j .. CT-1 (loads clamped into the tableau: no branches).
        // L_jj^-1 is lower triangular: the upper row tile needs k < 8 only.
        {
            double la[6];                                                 // A fragments: (row tile 0: k-steps 0, 1) (row tile 1: k-steps 0 .. 3)
#pragma unroll
            for (int kt = 0; kt < 2; kt++) la[kt] = linv[g8 * EK2_EB + kt * 4 + t4];
#pragma unroll
            for (int kt = 0; kt < 4; kt++) la[2 + kt] = linv[(8 + g8) * EK2_EB + kt * 4 + t4];
            for (int ct = 2 * j + wrp; ct < CT; ct += nwarps) {
                const int colbc = min(8 * ct + g8, ncols - 1);
                double bf[4];
#pragma unroll
                for (int kt = 0; kt < 4; kt++) {
                    const int k = kt * 4 + t4;
                    const double x = T[(size_t)(r0 + min(k, nb - 1)) * W + colbc];
                    bf[kt] = k < nb ? x : 0.0;
                }
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                hv_dmma(c0, c1, la[0], bf[0]); hv_dmma(d0, d1, la[2], bf[0]);
                hv_dmma(c0, c1, la[1], bf[1]); hv_dmma(d0, d1, la[3], bf[1]);
                if (nb > 8) { hv_dmma(d0, d1, la[4], bf[2]); hv_dmma(d0, d1, la[5], bf[3]); }
                const int col = 8 * ct + 2 * t4;
                if (g8 < nb) { if (col < ncols) T[(size_t)(r0 + g8) * W + col] = c0; if (col + 1 < ncols) T[(size_t)(r0 + g8) * W + col + 1] = c1; }
                if (8 + g8 < nb) { if (col < ncols) T[(size_t)(r0 + 8 + g8) * W + col] = d0; if (col + 1 < ncols) T[(size_t)(r0 + 8 + g8) * W + col + 1] = d1; }
            }
        }
        EK2_ELIM_MARK(1);
        __syncthreads();
        EK2_ELIM_MARK(2);
        // ---- b. trailing update: row tiles mt >= m0 = 2 (j + 1), column tiles nt >= mt. Warp 0 takes the tiles of the next diagonal
        // block, (m0, m0), (m0, m0 + 1), (m0 + 1, m0 + 1), and then factors it; the workers walk contiguous ranges of the rest
        // (row m0: nt >= m0 + 2; row m0 + 1: nt >= m0 + 2; row mt: nt >= mt), reloading the A fragments (-U_j[:, row tile]') only
        // when the row tile changes.
        if (j + 1 < MB) {                                                 // a following block exists: this one is full (nb = 16)
            const int m0 = 2 * (j + 1);
            const double* rowk[4];
#pragma unroll
            for (int kt = 0; kt < 4; kt++) rowk[kt] = T + (size_t)(r0 + kt * 4 + t4) * W;
            const bool ahead = wrp == 0;
            // Workers: the warps that do NOT share warp 0's scheduler / FP64 pipe (warp id mod 4 != 0). The pivot chain of the
            // look-ahead factorisation is a sequence of dependent fp64 operations; every DMMA a sibling warp queues on the
            // same pipe (16 cycles each) would sit in front of them.
            const int nwork = nwarps - nwarps / 4, widx = wrp - wrp / 4 - 1;           // 12 workers, index 0..11
            const bool worker = (wrp & 3) != 0;
            if (ahead) {
                for (int mt = m0; mt <= min(m0 + 1, MT - 1); mt++) {
                    double a[4];
                    const int am = min(8 * mt + g8, ncols - 1);
#pragma unroll
                    for (int kt = 0; kt < 4; kt++) a[kt] = -rowk[kt][am];
                    for (int nt = mt; nt <= min(m0 + 1, CT - 1); nt++) ek2_trailing_tile(T, W, n, ncols, mt, nt, rowk, a, g8, t4);
                }
            } else if (worker) {
                const int skip = m0 + 2;                                  // first column tile of the workers in rows m0, m0 + 1
                auto cnt = [&](int mt) { return max(0, CT - max(mt, skip)); };
                int total = 0;
                for (int mt = m0; mt < MT; mt++) total += cnt(mt);
                int lo = (int)(((long long)total * widx) / nwork);
                const int hi = (int)(((long long)total * (widx + 1)) / nwork);
                int mt = m0, idx = lo;
                while (mt < MT - 1 && idx >= cnt(mt)) { idx -= cnt(mt); mt++; }
                int nt = max(mt, skip) + idx;
                int curMt = -1;
                double a[4] = {0.0, 0.0, 0.0, 0.0};
                for (; lo < hi; lo++) {
                    if (mt != curMt) {
                        const int am = min(8 * mt + g8, ncols - 1);
#pragma unroll
                        for (int kt = 0; kt < 4; kt++) a[kt] = -rowk[kt][am];           // A[m][k] = -U_j[k][8 mt + m]
                        curMt = mt;
                    }
                    ek2_trailing_tile(T, W, n, ncols, mt, nt, rowk, a, g8, t4);
                    if (++nt >= CT) { do mt++; while (mt < MT - 1 && cnt(mt) == 0); nt = max(mt, skip); }
                }
            }
            EK2_ELIM_MARK(3);
            if (ahead) {
                __syncwarp();
                ek2_diag_factor(T, W, EK2_EB * (j + 1), min(EK2_EB, n - EK2_EB * (j + 1)), lane, s_linv + ((j + 1) & 1) * (EK2_EB * EK2_EB), s_bad, s_linv + 512);
            }
            EK2_ELIM_MARK(4);
            __syncthreads();
            EK2_ELIM_MARK(5);
        }
    }
    return !*s_bad;
}

#endif
