// hybvio_b200/csrc/ekf_predict.cuh -- body of the fused IMU-burst predict kernel (ekf.cu: ekf_predict_kernel).
//
// predict() of odometry::EKF (src/odometry/ekf.cpp:320-514) applied to `count` (<= 16) consecutive IMU samples in ONE
// launch of one CTA. What is sequential in the reference is only
//   * the quaternion chain            q_{k+1} = A_k q_k [/ |.|]               (4x4 mat-vec per sample; the
//                                     normalizeQuaternions(true) the reference calls after every predict is folded in),
//   * velocity / position sums        v_{k+1} = v_k + dv_k, p_{k+1} = p_k + v_k dt_k,
//   * the 20x20 covariance recursion  P00 <- D_k P00 D_k' + W_k,  Dacc <- D_k Dacc   (two barriers per sample);
// everything else of a sample depends on the state only through quantities that are known up front:
//   * the gyro / accelerometer biases only decay by a fixed factor per sample (ekf.cpp:443-448), so the angular rates
//     w_k, the rotation exp(-dt/2 Omega(w_k)) = A_k (cos / sin / sqrt in fp64: the longest scalar chain of a sample)
//     and T o a - b_a are computed for ALL samples at once, one warp per sample;
//   * once the quaternions are known, the Jacobians D_k (20x20), G_k (20x12) and the process-noise term
//     W_k = G_k Q_k G_k' of all samples are again computed in parallel, one warp per sample.
// The version before this one walked through ~9 barrier-separated stages per sample (4.1 us per sample, 45 us per
// 10-sample burst on B200); here a burst costs 3 parallel stages + 2 short chains + 2 barriers per sample.
//
// The two off-diagonal strips of P are transformed once at the end with Dacc = D_{c-1} ... D_0 (algebraically what
// the reference does sample by sample; fp64 differences are association-order rounding, ~1e-16 relative).
//
// Written against the small set of CUDA primitives that tests/emu/ can run on the host (threads, __syncthreads,
// __syncwarp, full-warp shuffles), so that the indexing logic is testable without a GPU (tests/emu/emu_predict.cpp).
#pragma once
#include "ekf.cuh"
#include "hv_dmma.cuh"

// per-sample scratch in dynamic shared memory (doubles)
#define PS_D 0          // dydx 20 x 20, column-major
#define PS_G 400        // dydq 20 x 12
#define PS_W 640        // G Q G' 20 x 20
#define PS_G1 1040      // G Q 20 x 12
#define PS_A 1280       // exp(-dt/2 Omega) 4 x 4, row-major
#define PS_B 1296       // 3 x 4
#define PS_TX 1308      // T o a - b_a
#define PS_DV 1311      // velocity increment
#define PS_STRIDE 1320
#define PDX(i, j) D[(i) + (j) * 20]
#define PDQ(i, j) G[(i) + (j) * 20]

__host__ __device__ inline size_t ekf_predict_smem_bytes(int count) { return (size_t)count * PS_STRIDE * sizeof(double); }

// Q of sample k: the stored Q with the two drift blocks replaced by the value in force at sample k (ekf.cpp:397-412)
__device__ __forceinline__ double ps_qval(const double* s_Q, int a, int b, double qBaa, double qBga)
{
    const bool baaA = a >= EKF_Q_BAA_DRIFT && a < EKF_Q_BAA_DRIFT + 3, baaB = b >= EKF_Q_BAA_DRIFT && b < EKF_Q_BAA_DRIFT + 3;
    const bool bgaA = a >= EKF_Q_BGA_DRIFT && a < EKF_Q_BGA_DRIFT + 3, bgaB = b >= EKF_Q_BGA_DRIFT && b < EKF_Q_BGA_DRIFT + 3;
    if (qBaa >= 0.0 && baaA && baaB) return a == b ? qBaa : 0.0;
    if (qBga >= 0.0 && bgaA && bgaB) return a == b ? qBga : 0.0;
    return s_Q[a + b * 12];
}

// Programmatic dependent launch (sm_90+); no-ops on the host emulator
__device__ __forceinline__ void ekf_pdl_launch_dependents()
{
#ifndef HV_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void ekf_pdl_wait()
{
#ifndef HV_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__device__ __forceinline__ void ekf_predict_body(const EkfPredictArgs& a, double* dyn)
{
    __shared__ double s_Q[144], s_P00[400], s_T1[400];
    __shared__ __align__(16) double s_acc[400];
    __shared__ double s_m[EKF_INER], s_q[(EKF_MAX_PREDICT + 1) * 4], s_qn[EKF_MAX_PREDICT * 4], s_mfinal[EKF_INER];
    __shared__ double s_u[(EKF_MAX_PREDICT + 1) * 4], s_nrm[EKF_MAX_PREDICT + 1];
    const int tid = threadIdx.x, N = a.b.N, cnt = a.count;
    const int lane = tid & 31, wrp = tid >> 5;
    double* P = a.b.P;
    const bool meanOnly = a.meanOut != nullptr;
    EKF_PMARK(0);
    ekf_pdl_launch_dependents();                      // programmatic dependent launch, as in ekf_cluster2.cuh: nothing of the filter
    ekf_pdl_wait();                                   // state is read before the previous kernel of the stream has completed
    if (!meanOnly) {
        for (int i = tid; i < 400; i += EKF_NT) s_P00[i] = P[(i % 20) + (size_t)(i / 20) * N];
        for (int i = tid; i < 144; i += EKF_NT) s_Q[i] = a.b.Q[i];
    }
    if (tid < EKF_INER) s_m[tid] = a.b.m[tid];
    __syncthreads();
    EKF_PMARK(1);

    // ---- stage 1 (warp k = sample k): decayed biases, A_k = exp(S) = cos(th) I + sin(th)/th S with S = -dt/2 Omega(w)
    // (closed form of the reference's Pade S.exp(), ekf.cpp:414-425: Omega^2 = -|w|^2 I), T o a - b_a, identity/zero fill
    for (int k = wrp; k < cnt; k += EKF_NT / 32) {
        const EkfPredictSample& S = a.s[k];
        double* smp = dyn + (size_t)k * PS_STRIDE;
        double* D = smp + PS_D; double* G = smp + PS_G;
        if (!meanOnly) {
            for (int i = lane; i < 400; i += 32) D[i] = (i % 21 == 0) ? 1.0 : 0.0;
            for (int i = lane; i < 240; i += 32) G[i] = 0.0;
        }
        double bg0 = s_m[EKF_BGA], bg1 = s_m[EKF_BGA + 1], bg2 = s_m[EKF_BGA + 2];
        double ba = s_m[EKF_BAA + (lane % 3)];
        for (int j = 0; j < k; j++) { const double dg = a.s[j].bgaDecay, da = a.s[j].baaDecay; bg0 *= dg; bg1 *= dg; bg2 *= dg; ba *= da; }
        const double dt = S.dt;
        const double w0 = S.xg[0] - bg0, w1 = S.xg[1] - bg1, w2 = S.xg[2] - bg2;
        const double c = -dt / 2;
        // cos(th) and sin(th)/th are even in th = |w| dt / 2: for the small angles of an IMU step (th^2 < 0.01) their Taylor
        // series in x = th^2 reach double precision with 6 terms -- no sqrt, no fp64 sin/cos on the critical path
        const double x = (w0 * w0 + w1 * w1 + w2 * w2) * (c * c);
        double ct, sc;
        if (x < 0.01) {
            ct = 1.0 + x * (-1.0 / 2 + x * (1.0 / 24 + x * (-1.0 / 720 + x * (1.0 / 40320 + x * (-1.0 / 3628800 + x * (1.0 / 479001600))))));
            sc = 1.0 + x * (-1.0 / 6 + x * (1.0 / 120 + x * (-1.0 / 5040 + x * (1.0 / 362880 + x * (-1.0 / 39916800 + x * (1.0 / 6227020800.0))))));
        } else {
            const double th = sqrt(x);
            ct = cos(th); sc = sin(th) / th;
        }
        if (lane < 16) {
            // Omega row-major {0,-w0,-w1,-w2, w0,0,-w2,w1, w1,w2,0,-w0, w2,-w1,w0,0}: component (3 = zero) / sign per entry
            const int comp = (0xC6396C93u >> (2 * lane)) & 3;
            const bool neg = (0x284Eu >> lane) & 1;
            double wv = comp == 0 ? w0 : comp == 1 ? w1 : comp == 2 ? w2 : 0.0;
            if (neg) wv = -wv;
            smp[PS_A + lane] = sc * (wv * c) + ((lane % 5 == 0) ? ct : 0.0);
        } else if (lane < 19) {
            // ba is the decayed accelerometer bias of component lane % 3; lanes 16, 17, 18 -> components 1, 2, 0
            smp[PS_TX + (lane % 3)] = s_m[EKF_BAT + (lane % 3)] * S.xa[lane % 3] - ba;
        }
    }
    __syncthreads();

    // ---- quaternion chain (warp 0). The rotations are applied to the UNNORMALISED chain u_{k+1} = A_k u_k (4 FMAs per
    // sample on the critical path); the normalizeQuaternions(true) calls that follow samples in the reference loop only
    // rescale: state before sample k = u_k / n_j, state left by predict k = u_{k+1} / n_j, with n_j = |u_j| of the latest
    // normalisation j <= k (ekf.cpp:1024-1032) -- all divisions happen afterwards, in parallel over the samples.
    if (wrp == 0) {
        const int r = lane & 3;
        double u = s_m[EKF_ORI + r];
        for (int k = 0; k < cnt; k++) {
            if (lane < 4) s_u[k * 4 + lane] = u;
            const double* A = dyn + (size_t)k * PS_STRIDE + PS_A;
            const double u0 = __shfl_sync(0xffffffffu, u, 0), u1 = __shfl_sync(0xffffffffu, u, 1);
            const double u2 = __shfl_sync(0xffffffffu, u, 2), u3 = __shfl_sync(0xffffffffu, u, 3);
            double v = 0;
            v += A[r * 4] * u0; v += A[r * 4 + 1] * u1; v += A[r * 4 + 2] * u2; v += A[r * 4 + 3] * u3;
            u = v;
        }
        if (lane < 4) s_u[cnt * 4 + lane] = u;
        __syncwarp();
        if (lane <= cnt) {                       // lane j: n_j = |u_j| (only used where a normalisation ends at j)
            const double* uj = s_u + lane * 4;
            const double z = (uj[0] * uj[0] + uj[2] * uj[2]) + (uj[1] * uj[1] + uj[3] * uj[3]);
            s_nrm[lane] = z > 0.0 ? sqrt(z) : 1.0;
        }
        __syncwarp();
        if (lane <= cnt) {                       // lane k: scale in force while sample k runs
            int j = lane;
            while (j > 0 && !a.s[j - 1].normAfter) j--;
            const double nj = j > 0 ? s_nrm[j] : 1.0;
            const bool scaled = j > 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                s_q[lane * 4 + i] = scaled ? s_u[lane * 4 + i] / nj : s_u[lane * 4 + i];
                if (lane < cnt) s_qn[lane * 4 + i] = scaled ? s_u[(lane + 1) * 4 + i] / nj : s_u[(lane + 1) * 4 + i];
            }
        }
    }
    __syncthreads();
    EKF_PMARK(2);

    // ---- stage 2 (warp k = sample k): Jacobians (ekf.cpp:450-498), velocity increment, W_k = G_k Q_k G_k'
    for (int k = wrp; k < cnt; k += EKF_NT / 32) {
        const EkfPredictSample& S = a.s[k];
        double* smp = dyn + (size_t)k * PS_STRIDE;
        double* D = smp + PS_D; double* G = smp + PS_G; double* A = smp + PS_A; double* B = smp + PS_B;
        double* W = smp + PS_W; double* G1 = smp + PS_G1;
        const double* Tx = smp + PS_TX;
        const double dt = S.dt;
        const double* qo = s_q + k * 4;           // orientation before the sample
        const double* qn = s_qn + k * 4;          // and after (before a normalizeQuaternions that may follow)
        if (meanOnly) {                           // velocity increment only: the same expression as below, in the same order
            if (lane >= 8 && lane < 11) {
                const double* q = qn;
                const int i = lane - 8;
                double R[9];
                R[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3]; R[1] = 2 * q[1] * q[2] - 2 * q[0] * q[3]; R[2] = 2 * q[1] * q[3] + 2 * q[0] * q[2];
                R[3] = 2 * q[1] * q[2] + 2 * q[0] * q[3]; R[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3]; R[5] = 2 * q[2] * q[3] - 2 * q[0] * q[1];
                R[6] = 2 * q[1] * q[3] - 2 * q[0] * q[2]; R[7] = 2 * q[2] * q[3] + 2 * q[0] * q[1]; R[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
                const double gi = i == 2 ? -a.gravity : 0.0;
                smp[PS_DV + i] = (R[i] * Tx[0] + R[3 + i] * Tx[1] + R[6 + i] * Tx[2] + gi) * dt;
            }
            continue;
        }
        if (lane < 3) {
            // d(orientation)/d(gyro noise) columns A dS_j q (ekf.cpp:470-476) and their negatives (ekf.cpp:492)
            const int j = lane;
            const double h = dt / 2;
            const double q0 = qo[0], q1 = qo[1], q2 = qo[2], q3 = qo[3];
            double t0, t1, t2, t3;                // dS_j * q
            if (j == 0) { t0 = h * q1; t1 = -h * q0; t2 = h * q3; t3 = -h * q2; }
            else if (j == 1) { t0 = h * q2; t1 = -h * q3; t2 = -h * q0; t3 = h * q1; }
            else { t0 = h * q3; t1 = h * q2; t2 = -h * q1; t3 = -h * q0; }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double v = A[i * 4] * t0 + A[i * 4 + 1] * t1 + A[i * 4 + 2] * t2 + A[i * 4 + 3] * t3;
                PDQ(EKF_ORI + i, EKF_Q_GYRO + j) = v;
                PDX(EKF_ORI + i, EKF_BGA + j) = -v;
            }
            PDX(EKF_POS + lane, EKF_VEL + lane) = dt;
            PDQ(EKF_BGA + lane, EKF_Q_BGA_DRIFT + lane) = 1.0; PDQ(EKF_BAA + lane, EKF_Q_BAA_DRIFT + lane) = 1.0;
        } else if (lane >= 4 && lane < 8) {
            // B[:, qi] = dR[qi]' Txab dt with the rotation of the NEW quaternion (src/odometry/util.cpp:10-47)
            const int qi = lane - 4;
            const double a2 = 2 * qn[0], b2 = 2 * qn[1], c2 = 2 * qn[2], d2 = 2 * qn[3];
            double Dm[9];
            if (qi == 0) { Dm[0] = a2; Dm[1] = -d2; Dm[2] = c2; Dm[3] = d2; Dm[4] = a2; Dm[5] = -b2; Dm[6] = -c2; Dm[7] = b2; Dm[8] = a2; }
            else if (qi == 1) { Dm[0] = b2; Dm[1] = c2; Dm[2] = d2; Dm[3] = c2; Dm[4] = -b2; Dm[5] = -a2; Dm[6] = d2; Dm[7] = a2; Dm[8] = -b2; }
            else if (qi == 2) { Dm[0] = -c2; Dm[1] = b2; Dm[2] = a2; Dm[3] = b2; Dm[4] = c2; Dm[5] = d2; Dm[6] = -a2; Dm[7] = d2; Dm[8] = -c2; }
            else { Dm[0] = -d2; Dm[1] = -a2; Dm[2] = b2; Dm[3] = a2; Dm[4] = -d2; Dm[5] = c2; Dm[6] = b2; Dm[7] = c2; Dm[8] = d2; }
#pragma unroll
            for (int i = 0; i < 3; i++) B[i * 4 + qi] = (Dm[i] * Tx[0] + Dm[3 + i] * Tx[1] + Dm[6 + i] * Tx[2]) * dt;
        } else if (lane >= 8 && lane < 11) {
            const double* q = qn;
            const int i = lane - 8;
            double R[9];
            R[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3]; R[1] = 2 * q[1] * q[2] - 2 * q[0] * q[3]; R[2] = 2 * q[1] * q[3] + 2 * q[0] * q[2];
            R[3] = 2 * q[1] * q[2] + 2 * q[0] * q[3]; R[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3]; R[5] = 2 * q[2] * q[3] - 2 * q[0] * q[1];
            R[6] = 2 * q[1] * q[3] - 2 * q[0] * q[2]; R[7] = 2 * q[2] * q[3] + 2 * q[0] * q[1]; R[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
            const double gi = i == 2 ? -a.gravity : 0.0;
            // velocity += (R' Txab + g) dt with the biases before this sample's decay (ekf.cpp:435-436)
            smp[PS_DV + i] = (R[i] * Tx[0] + R[3 + i] * Tx[1] + R[6 + i] * Tx[2] + gi) * dt;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                PDQ(EKF_VEL + i, EKF_Q_ACC + j) = R[j * 3 + i] * dt;
                PDX(EKF_VEL + i, EKF_BAA + j) = -R[j * 3 + i] * dt;
                PDX(EKF_VEL + i, EKF_BAT + j) = R[j * 3 + i] * S.xa[j] * dt;
            }
        } else if (lane >= 16) {
            PDX(EKF_ORI + (lane - 16) / 4, EKF_ORI + (lane - 16) % 4) = A[lane - 16];
        }
        __syncwarp();
        // d vel / d quat = B A  (ekf.cpp:458-461)
        if (lane < 12) {
            const int i = lane / 4, j = lane % 4;
            double v = 0;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) v += B[i * 4 + kk] * A[kk * 4 + j];
            PDX(EKF_VEL + i, EKF_ORI + j) = v;
        }
        __syncwarp();
        // d vel / d gyro noise and d vel / d gyro bias (ekf.cpp:486-489)
        if (lane < 9) {
            const int i = lane / 3, j = lane % 3;
            double v = 0;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) v += PDX(EKF_VEL + i, EKF_ORI + kk) * PDQ(EKF_ORI + kk, EKF_Q_GYRO + j);
            PDQ(EKF_VEL + i, EKF_Q_GYRO + j) = v;
            PDX(EKF_VEL + i, EKF_BGA + j) = -v;
        }
        __syncwarp();
        // drift-block values in force at this sample (the last one set at or before k; ekf.cpp:397-412)
        double qBaa = -1.0, qBga = -1.0;
        for (int j = 0; j <= k; j++) { if (a.s[j].qBaa >= 0.0) qBaa = a.s[j].qBaa; if (a.s[j].qBga >= 0.0) qBga = a.s[j].qBga; }
        // G1 = G Q_k (20 x 12 x 12) and W = G1 G' (20 x 20 x 12) on the fp64 tensor cores (hv_dmma.cuh): 8 x 8 tiles, k in steps of 4
        const int g8 = lane >> 2, t4 = lane & 3;
        for (int mt = 0; mt < 3; mt++)
            for (int nt = 0; nt < 2; nt++) {
                double c0 = 0.0, c1 = 0.0;
                const int row = mt * 8 + g8, colb = nt * 8 + g8;
#pragma unroll
                for (int kt = 0; kt < 3; kt++) {
                    const int kk = kt * 4 + t4;
                    const double av = row < 20 ? PDQ(row, kk) : 0.0;
                    const double bv = colb < 12 ? ps_qval(s_Q, kk, colb, qBaa, qBga) : 0.0;
                    hv_dmma(c0, c1, av, bv);
                }
                const int col = nt * 8 + 2 * t4;
                if (row < 20 && col < 12) { G1[row + col * 20] = c0; G1[row + (col + 1) * 20] = c1; }
            }
        __syncwarp();
        for (int mt = 0; mt < 3; mt++)
            for (int nt = 0; nt < 3; nt++) {
                double c0 = 0.0, c1 = 0.0;
                const int row = mt * 8 + g8, colb = nt * 8 + g8;
#pragma unroll
                for (int kt = 0; kt < 3; kt++) {
                    const int kk = kt * 4 + t4;
                    const double av = row < 20 ? G1[row + kk * 20] : 0.0;
                    const double bv = colb < 20 ? PDQ(colb, kk) : 0.0;       // B = G': B[k][n] = G(n, k)
                    hv_dmma(c0, c1, av, bv);
                }
                const int col = nt * 8 + 2 * t4;
                if (row < 20 && col < 20) { W[row + col * 20] = c0; W[row + (col + 1) * 20] = c1; }
            }
    }
    __syncthreads();
    EKF_PMARK(3);

    // ---- mean chains (last warp, idle during the covariance recursion): position with the OLD velocity, velocity,
    // mean-reverting biases (ekf.cpp:432-448); orientation = end of the quaternion chain
    if (wrp == EKF_NT / 32 - 1) {
        if (lane < 3) {
            double p = s_m[EKF_POS + lane], v = s_m[EKF_VEL + lane];
            for (int k = 0; k < cnt; k++) { p = p + v * a.s[k].dt; v = v + dyn[(size_t)k * PS_STRIDE + PS_DV + lane]; }
            s_mfinal[EKF_POS + lane] = p; s_mfinal[EKF_VEL + lane] = v;
        } else if (lane < 7) s_mfinal[EKF_ORI + lane - 3] = s_q[cnt * 4 + lane - 3];
        else if (lane < 10) { double b = s_m[EKF_BAA + lane - 7]; for (int k = 0; k < cnt; k++) b *= a.s[k].baaDecay; s_mfinal[EKF_BAA + lane - 7] = b; }
        else if (lane < 13) { double b = s_m[EKF_BGA + lane - 10]; for (int k = 0; k < cnt; k++) b *= a.s[k].bgaDecay; s_mfinal[EKF_BGA + lane - 10] = b; }
        else if (lane < 17) s_mfinal[EKF_BAT + lane - 13] = s_m[EKF_BAT + lane - 13];     // BAT (3) and SFT (1) are constant
    }

    if (meanOnly) {
        __syncthreads();
        if (tid < EKF_INER) a.meanOut[tid] = s_mfinal[tid];
        return;
    }

    // ---- covariance recursion on the fp64 tensor cores: T1 = D P00 and Dacc' = D Dacc (one 8 x 8 tile of each per warp,
    // warps 0..8), barrier, P00 = T1 D' + W, barrier. Two barriers and two 5-deep DMMA chains per sample.
    {
        const int g8 = lane >> 2, t4 = lane & 3;
        const int mt = wrp / 3, nt = wrp % 3;                 // tile of warps 0..8
        const int row = mt * 8 + g8, colb = nt * 8 + g8, col = nt * 8 + 2 * t4;
        const bool tileWarp = wrp < 9;
        for (int k = 0; k < cnt; k++) {
            const double* D = dyn + (size_t)k * PS_STRIDE + PS_D;
            const double* W = dyn + (size_t)k * PS_STRIDE + PS_W;
            double a0 = 0.0, a1 = 0.0;                        // Dacc' tile
            if (tileWarp) {
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int kt = 0; kt < 5; kt++) {
                    const int kk = kt * 4 + t4;
                    const double dv = row < 20 ? PDX(row, kk) : 0.0;
                    const double pv = colb < 20 ? s_P00[kk + colb * 20] : 0.0;
                    hv_dmma(t0, t1, dv, pv);
                    if (k > 0) { const double cv = colb < 20 ? s_acc[kk + colb * 20] : 0.0; hv_dmma(a0, a1, dv, cv); }
                }
                if (k == 0 && row < 20 && col < 20) { a0 = PDX(row, col); a1 = PDX(row, col + 1); }      // Dacc = D_0
                if (row < 20 && col < 20) { s_T1[row + col * 20] = t0; s_T1[row + (col + 1) * 20] = t1; }   // T1 was last read before the previous barrier
            }
            __syncthreads();
            if (tileWarp) {
                if (row < 20 && col < 20) { s_acc[row + col * 20] = a0; s_acc[row + (col + 1) * 20] = a1; }   // all reads of Dacc happened before the barrier
                double p0 = 0.0, p1 = 0.0;
                if (row < 20 && col < 20) { p0 = W[row + col * 20]; p1 = W[row + (col + 1) * 20]; }
#pragma unroll
                for (int kt = 0; kt < 5; kt++) {
                    const int kk = kt * 4 + t4;
                    const double tv = row < 20 ? s_T1[row + kk * 20] : 0.0;
                    const double dv = colb < 20 ? PDX(colb, kk) : 0.0;      // B = D': B[k][n] = D(n, k)
                    hv_dmma(p0, p1, tv, dv);
                }
                if (row < 20 && col < 20) { s_P00[row + col * 20] = p0; s_P00[row + (col + 1) * 20] = p1; }
            }
            __syncthreads();
        }
    }
    EKF_PMARK(4);

    // ---- write back the inertial block, then transform the two strips with the accumulated Jacobian
    if (cnt > 0) {
        const double* Dlast = dyn + (size_t)(cnt - 1) * PS_STRIDE + PS_D;
        if (tid < EKF_INER) a.b.m[tid] = s_mfinal[tid];
        for (int i = tid; i < 400; i += EKF_NT) { a.b.dydx[i] = Dlast[i]; P[(i % 20) + (size_t)(i / 20) * N] = s_P00[i]; }
        double qBaa = -1.0, qBga = -1.0;
        for (int j = 0; j < cnt; j++) { if (a.s[j].qBaa >= 0.0) qBaa = a.s[j].qBaa; if (a.s[j].qBga >= 0.0) qBga = a.s[j].qBga; }
        for (int i = tid; i < 144; i += EKF_NT) a.b.Q[i] = ps_qval(s_Q, i % 12, i / 12, qBaa, qBga);
    }
    // ---- the two strips as ONE product on the tensor cores: X = [ P[20:, 0:20] ; P[0:20, 20:]' ] (2 (N-20) x 20),
    // Y = X Dacc' (ekf.cpp:506-508: P[20:,0:20] Dacc'  and  Dacc P[0:20,20:]); 8-row tiles dealt to the warps, the A
    // fragments of all tiles of a warp are loaded first (one L2 round trip), Dacc' fragments live in registers.
    if (cnt > 0) {
        const int g8 = lane >> 2, t4 = lane & 3;
        const int rest = N - EKF_INER, rows = 2 * rest, ntile = (rows + 7) >> 3;
        double bf[3][5];
#pragma unroll
        for (int nt = 0; nt < 3; nt++)
#pragma unroll
            for (int kt = 0; kt < 5; kt++) { const int nn = nt * 8 + g8, kk = kt * 4 + t4; bf[nt][kt] = nn < 20 ? s_acc[nn + kk * 20] : 0.0; }   // B[k][n] = Dacc(n, k)
        for (int tb = wrp; tb < ntile; tb += 3 * (EKF_NT / 32)) {
            double af[3][5];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int r = (tb + q * (EKF_NT / 32)) * 8 + g8;
#pragma unroll
                for (int kt = 0; kt < 5; kt++) {
                    const int kk = kt * 4 + t4;
                    double v = 0.0;
                    if (r < rest) v = P[EKF_INER + r + (size_t)kk * N];
                    else if (r < rows) v = P[kk + (size_t)(EKF_INER + r - rest) * N];
                    af[q][kt] = v;
                }
            }
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int tile = tb + q * (EKF_NT / 32);
                if (tile >= ntile) break;                                   // warp-uniform
                const int r = tile * 8 + g8;
#pragma unroll
                for (int nt = 0; nt < 3; nt++) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int kt = 0; kt < 5; kt++) hv_dmma(c0, c1, af[q][kt], bf[nt][kt]);
                    const int col = nt * 8 + 2 * t4;
                    if (col < 20) {
                        if (r < rest) { P[EKF_INER + r + (size_t)col * N] = c0; P[EKF_INER + r + (size_t)(col + 1) * N] = c1; }
                        else if (r < rows) { double* cp = P + (size_t)(EKF_INER + r - rest) * N; cp[col] = c0; cp[col + 1] = c1; }
                    }
                }
            }
        }
    }
    EKF_PMARK(5);
}
