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
// Written against the small set of CUDA primitives that tools/emu/ can run on the host (threads, __syncthreads,
// __syncwarp, full-warp shuffles), so that the indexing logic is testable without a GPU (tools/emu/emu_predict.cpp).
#pragma once
#include "ekf.cuh"

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

__device__ __forceinline__ void ekf_predict_body(const EkfPredictArgs& a, double* dyn)
{
    __shared__ double s_Q[144], s_P00[400], s_T1[400];
    __shared__ __align__(16) double s_acc[400];
    __shared__ double s_m[EKF_INER], s_q[(EKF_MAX_PREDICT + 1) * 4], s_qn[EKF_MAX_PREDICT * 4], s_mfinal[EKF_INER];
    const int tid = threadIdx.x, N = a.b.N, cnt = a.count;
    const int lane = tid & 31, wrp = tid >> 5;
    double* P = a.b.P;
    EKF_PMARK(0);
    for (int i = tid; i < 400; i += EKF_NT) s_P00[i] = P[(i % 20) + (size_t)(i / 20) * N];
    for (int i = tid; i < 144; i += EKF_NT) s_Q[i] = a.b.Q[i];
    if (tid < EKF_INER) s_m[tid] = a.b.m[tid];
    __syncthreads();
    EKF_PMARK(1);

    // ---- stage 1 (warp k = sample k): decayed biases, A_k = exp(S) = cos(th) I + sin(th)/th S with S = -dt/2 Omega(w)
    // (closed form of the reference's Pade S.exp(), ekf.cpp:414-425: Omega^2 = -|w|^2 I), T o a - b_a, identity/zero fill
    for (int k = wrp; k < cnt; k += EKF_NT / 32) {
        const EkfPredictSample& S = a.s[k];
        double* smp = dyn + (size_t)k * PS_STRIDE;
        double* D = smp + PS_D; double* G = smp + PS_G;
        for (int i = lane; i < 400; i += 32) D[i] = (i % 21 == 0) ? 1.0 : 0.0;
        for (int i = lane; i < 240; i += 32) G[i] = 0.0;
        double bg0 = s_m[EKF_BGA], bg1 = s_m[EKF_BGA + 1], bg2 = s_m[EKF_BGA + 2];
        double ba = s_m[EKF_BAA + (lane % 3)];
        for (int j = 0; j < k; j++) { const double dg = a.s[j].bgaDecay, da = a.s[j].baaDecay; bg0 *= dg; bg1 *= dg; bg2 *= dg; ba *= da; }
        const double dt = S.dt;
        const double w0 = S.xg[0] - bg0, w1 = S.xg[1] - bg1, w2 = S.xg[2] - bg2;
        const double c = -dt / 2;
        const double th = sqrt(w0 * w0 + w1 * w1 + w2 * w2) * fabs(c);
        const double ct = cos(th), sc = th < 1e-8 ? 1.0 - th * th / 6.0 : sin(th) / th;
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

    // ---- quaternion chain q_{k+1} = A_k q_k (warp 0; every lane takes part in the shuffles, lanes 0..3 hold q)
    if (wrp == 0) {
        const int r = lane & 3;
        double q = s_m[EKF_ORI + r];
        for (int k = 0; k < cnt; k++) {
            if (lane < 4) s_q[k * 4 + lane] = q;
            const double* A = dyn + (size_t)k * PS_STRIDE + PS_A;
            double v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) v += A[r * 4 + j] * __shfl_sync(0xffffffffu, q, j);
            if (lane < 4) s_qn[k * 4 + lane] = v;   // what predict() itself leaves in the state: the Jacobians use this one
            if (a.s[k].normAfter) {               // normalizeQuaternions(true) right after this sample (ekf.cpp:1024-1032)
                const double sq = v * v;
                const double z0 = __shfl_sync(0xffffffffu, sq, 0), z1 = __shfl_sync(0xffffffffu, sq, 1);
                const double z2 = __shfl_sync(0xffffffffu, sq, 2), z3 = __shfl_sync(0xffffffffu, sq, 3);
                const double z = (z0 + z2) + (z1 + z3);
                if (z > 0.0) v /= sqrt(z);
            }
            q = v;
        }
        if (lane < 4) s_q[cnt * 4 + lane] = q;
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
        for (int t = lane; t < 240; t += 32) {
            const int i = t % 20, j = t / 20;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
            for (int kk = 0; kk < 12; kk += 4) {
                s0 += PDQ(i, kk) * ps_qval(s_Q, kk, j, qBaa, qBga); s1 += PDQ(i, kk + 1) * ps_qval(s_Q, kk + 1, j, qBaa, qBga);
                s2 += PDQ(i, kk + 2) * ps_qval(s_Q, kk + 2, j, qBaa, qBga); s3 += PDQ(i, kk + 3) * ps_qval(s_Q, kk + 3, j, qBaa, qBga);
            }
            G1[t] = (s0 + s1) + (s2 + s3);
        }
        __syncwarp();
        for (int t = lane; t < 400; t += 32) {
            const int i = t % 20, j = t / 20;
            double g0 = 0, g1 = 0, g2 = 0, g3 = 0;
#pragma unroll
            for (int kk = 0; kk < 12; kk += 4) {
                g0 += G1[i + kk * 20] * PDQ(j, kk); g1 += G1[i + (kk + 1) * 20] * PDQ(j, kk + 1);
                g2 += G1[i + (kk + 2) * 20] * PDQ(j, kk + 2); g3 += G1[i + (kk + 3) * 20] * PDQ(j, kk + 3);
            }
            W[t] = (g0 + g1) + (g2 + g3);
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

    // ---- covariance recursion: P00 = D P00 D' + W and Dacc = D Dacc (4 interleaved partial sums per dot product)
    for (int k = 0; k < cnt; k++) {
        const double* D = dyn + (size_t)k * PS_STRIDE + PS_D;
        const double* W = dyn + (size_t)k * PS_STRIDE + PS_W;
        double accNew = 0.0;
        if (tid < 400) {
            const int i = tid % 20, j = tid / 20;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0, u0 = 0, u1 = 0, u2 = 0, u3 = 0;
            if (k == 0) {
#pragma unroll
                for (int kk = 0; kk < 20; kk += 4) {
                    s0 += PDX(i, kk) * s_P00[kk + j * 20]; s1 += PDX(i, kk + 1) * s_P00[kk + 1 + j * 20]; s2 += PDX(i, kk + 2) * s_P00[kk + 2 + j * 20]; s3 += PDX(i, kk + 3) * s_P00[kk + 3 + j * 20];
                }
                accNew = PDX(i, j);
            } else {
#pragma unroll
                for (int kk = 0; kk < 20; kk += 4) {
                    s0 += PDX(i, kk) * s_P00[kk + j * 20]; s1 += PDX(i, kk + 1) * s_P00[kk + 1 + j * 20]; s2 += PDX(i, kk + 2) * s_P00[kk + 2 + j * 20]; s3 += PDX(i, kk + 3) * s_P00[kk + 3 + j * 20];
                    u0 += PDX(i, kk) * s_acc[kk + j * 20]; u1 += PDX(i, kk + 1) * s_acc[kk + 1 + j * 20]; u2 += PDX(i, kk + 2) * s_acc[kk + 2 + j * 20]; u3 += PDX(i, kk + 3) * s_acc[kk + 3 + j * 20];
                }
                accNew = (u0 + u1) + (u2 + u3);
            }
            s_T1[tid] = (s0 + s1) + (s2 + s3);      // T1 was last read before the previous barrier
        }
        __syncthreads();
        if (tid < 400) {
            const int i = tid % 20, j = tid / 20;
            s_acc[tid] = accNew;                    // all reads of Dacc happened before the barrier above
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
            for (int kk = 0; kk < 20; kk += 4) { s0 += s_T1[i + kk * 20] * PDX(j, kk); s1 += s_T1[i + (kk + 1) * 20] * PDX(j, kk + 1); s2 += s_T1[i + (kk + 2) * 20] * PDX(j, kk + 2); s3 += s_T1[i + (kk + 3) * 20] * PDX(j, kk + 3); }
            s_P00[tid] = ((s0 + s1) + (s2 + s3)) + W[tid];
        }
        __syncthreads();
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
#define PAC(i, j) s_acc[(i) + (j) * 20]
    const int rest = N - EKF_INER;
    for (int r = tid; r < 2 * rest && cnt > 0; r += EKF_NT) {
        if (r < rest) {                                     // P[20+r, 0:20] = P[20+r, 0:20] * Dacc'
            const int i = EKF_INER + r;
            double row[20], out[20];
#pragma unroll
            for (int k = 0; k < 20; k++) row[k] = P[i + (size_t)k * N];
#pragma unroll
            for (int j = 0; j < 20; j++) out[j] = 0.0;
#pragma unroll
            for (int k = 0; k < 20; k++) {
                const double rk = row[k];
                const double2* ac = reinterpret_cast<const double2*>(&PAC(0, k));     // column k of Dacc: 10 x 16-byte loads
#pragma unroll
                for (int j = 0; j < 10; j++) { const double2 v = ac[j]; out[2 * j] += rk * v.x; out[2 * j + 1] += rk * v.y; }
            }
#pragma unroll
            for (int j = 0; j < 20; j++) P[i + (size_t)j * N] = out[j];
        } else {                                            // P[0:20, 20+c] = Dacc * P[0:20, 20+c]
            double* colp = P + (size_t)(EKF_INER + r - rest) * N;
            double col[20], out[20];
#pragma unroll
            for (int k = 0; k < 20; k++) col[k] = colp[k];
#pragma unroll
            for (int j = 0; j < 20; j++) out[j] = 0.0;
#pragma unroll
            for (int k = 0; k < 20; k++) {
                const double ck = col[k];
                const double2* ac = reinterpret_cast<const double2*>(&PAC(0, k));
#pragma unroll
                for (int j = 0; j < 10; j++) { const double2 v = ac[j]; out[2 * j] += v.x * ck; out[2 * j + 1] += v.y * ck; }
            }
#pragma unroll
            for (int j = 0; j < 20; j++) colp[j] = out[j];
        }
    }
    EKF_PMARK(5);
}
