/*
 * oracle/hv_oracle_ekf.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (fp64) CPU restatement of the reference's EKF, src/odometry/ekf.cpp (EKFImplementation), written in
 * the reference's own algebra: explicit HP, S, Kalman gain K = (S^-1 HP)', P -= K HP, and the Joseph-form
 * augmentation with its two dense N x N x N products (ekf.cpp:35-50, 848-885) -- deliberately NOT the
 * elimination-tableau formulation the CUDA kernels use, so that the two are independent derivations.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED. tests/test_oracle_ekf.py checks every function below against oracle/_ref/libref_ekf.so
 * (the reference's ekf.cpp compiled unmodified with the vendored Eigen, oracle/ref_build/build_ekf.sh), against
 * tests/golden/ekf_golden.npz generated from that library, and against the reference's own unit-test vectors
 * (test/ekf.cpp:19-71 chi-square KAT, test/data/P.csv + m.csv transformTo round trip).
 * Differences to Eigen are summation order only (Eigen: blocked GEMM + pivoted LDLT; here: plain loops +
 * Cholesky), i.e. ~1e-13 relative.
 *
 * Function <-> reference map (all in src/odometry/ekf.cpp):
 *   orc_ekf_create 153-296 | initialize_orientation 299-317 | predict 320-514 | update() helper 57-82
 *   update_zupt.. 573-677 | get/set_inertial 679-690 | translate_to 696-702 | transform_to 704-758
 *   visual_check 760-819 | visual_update 829-844 | augment 848-885 (+35-50) | unaugment 888-903
 *   insert_map_point 911-921 | condition_on_last_pose 928-942 | lock_biases 944-947
 *   normalize_quaternions 1024-1032 | symmetrize 1059-1067;  quat2rmat_d: src/odometry/util.cpp:10-47
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { POS = 0, VEL = 3, ORI = 6, BGA = 10, BAA = 13, BAT = 16, SFT = 19, CAM = 20, INER = 20, POSE = 7, MAPPT = 3 };
enum { Q_ACC = 0, Q_GYRO = 3, Q_BGA_DRIFT = 6, Q_BAA_DRIFT = 9, Q_DIM = 12 };

typedef struct {   /* same layout as hv_ekf_params (include/hybvio_b200.h) */
    int camera_trail_length, hybrid_map_size;
    double noise_scale, gravity;
    double noise_initial_pos, noise_initial_vel, noise_initial_ori, noise_initial_bga, noise_initial_baa, noise_initial_bat, noise_initial_sft;
    double noise_initial_pos_trail, noise_initial_ori_trail;
    double noise_process_acc, noise_process_gyro, noise_process_baa, noise_process_baa_rev, noise_process_bga, noise_process_bga_rev;
    double augment_r, init_zupt_r, rotation_zupt_r;
} orc_params;

typedef struct {
    orc_params prm;
    int N, trail, mapDim;
    double noiseScale;
    double *m, *P, *Q, dydx[400];
    int augmentCount; double* augmentTimes; int nAugTimes;
    double time, ZUPTtime, ZRUPTtime, initZUPTtime; int wasStationary;
    double prevSampleT, firstSampleT; int firstSample;
} orc_ekf;

#define Pm(i, j) e->P[(i) + (size_t)(j) * e->N]
static double sq(double x) { return x * x; }

/* ---- chi2inv95 (odometry/util.hpp:23): recomputed by inverting the regularised incomplete gamma function */
static double gamma_p(double a, double x)
{
    if (x <= 0) return 0.0;
    double gln = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, sum = 1.0 / a, del = sum;
        for (int i = 0; i < 2000; i++) { ap += 1.0; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-17) break; }
        return sum * exp(-x + a * log(x) - gln);
    }
    double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; i++) {
        double an = -i * (i - a); b += 2.0;
        d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
        c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
        d = 1.0 / d; double del = d * c; h *= del;
        if (fabs(del - 1.0) < 1e-17) break;
    }
    return 1.0 - exp(-x + a * log(x) - gln) * h;
}
double orc_chi2inv95(int k)
{
    if (k <= 0) return 0.0;
    double lo = 0.0, hi = k + 10.0 * sqrt(2.0 * k) + 20.0;
    for (int it = 0; it < 200; it++) {
        double mid = 0.5 * (lo + hi);
        if (gamma_p(0.5 * k, 0.5 * mid) < 0.95) lo = mid; else hi = mid;
        if (hi - lo < 1e-15 * hi) break;
    }
    return 0.5 * (lo + hi);
}

void orc_ekf_default_params(orc_params* p)
{   /* codegen/parameter_definitions.c:68-160 */
    p->camera_trail_length = 20; p->hybrid_map_size = 0; p->noise_scale = 100; p->gravity = 9.819;
    p->noise_initial_pos = 1e-5; p->noise_initial_vel = 0.1; p->noise_initial_ori = 0.0316227766;
    p->noise_initial_bga = 1e-3; p->noise_initial_baa = 1e-6; p->noise_initial_bat = 1e-5; p->noise_initial_sft = 1e-5;
    p->noise_initial_pos_trail = 100; p->noise_initial_ori_trail = 3.16227766;
    p->noise_process_acc = 0.003; p->noise_process_gyro = 0.00017;
    p->noise_process_baa = 1e-4; p->noise_process_baa_rev = 0.1; p->noise_process_bga = 0; p->noise_process_bga_rev = 0.1;
    p->augment_r = 1e-9; p->init_zupt_r = 1e-4; p->rotation_zupt_r = 1e-6;
}

orc_ekf* orc_ekf_create(const orc_params* prm)
{
    orc_ekf* e = (orc_ekf*)calloc(1, sizeof(orc_ekf));
    e->prm = *prm; e->trail = prm->camera_trail_length; e->mapDim = prm->hybrid_map_size * MAPPT;
    e->N = INER + e->trail * POSE + e->mapDim;
    e->noiseScale = prm->noise_scale * prm->noise_scale;
    int N = e->N;
    e->m = (double*)calloc(N, sizeof(double)); e->P = (double*)calloc((size_t)N * N, sizeof(double)); e->Q = (double*)calloc(144, sizeof(double));
    e->augmentTimes = (double*)calloc(e->trail + 2, sizeof(double));
    e->ZUPTtime = e->ZRUPTtime = e->initZUPTtime = -1.0; e->prevSampleT = e->firstSampleT = -1.0; e->firstSample = 1;
    e->m[ORI] = 1.0;
    for (int i = 0; i < 3; i++) e->m[BAT + i] = 1.0;
    for (int i = 0; i < 3; i++) { Pm(POS + i, POS + i) = sq(prm->noise_initial_pos); Pm(VEL + i, VEL + i) = sq(prm->noise_initial_vel); }
    for (int i = 0; i < 4; i++) Pm(ORI + i, ORI + i) = 1.0;
    for (int i = 0; i < 3; i++) { Pm(BGA + i, BGA + i) = sq(prm->noise_initial_bga); Pm(BAA + i, BAA + i) = sq(prm->noise_initial_baa); Pm(BAT + i, BAT + i) = sq(prm->noise_initial_bat); }
    Pm(SFT, SFT) = sq(prm->noise_initial_sft);
    for (int p = 0; p < e->trail; p++) {
        int o = CAM + p * POSE;
        for (int i = 0; i < 3; i++) Pm(o + i, o + i) = sq(prm->noise_initial_pos_trail);
        for (int i = 0; i < 4; i++) Pm(o + 3 + i, o + 3 + i) = sq(prm->noise_initial_ori_trail);
    }
    for (int i = 0; i < 3; i++) { e->Q[(Q_ACC + i) * 13] = sq(prm->noise_process_acc); e->Q[(Q_GYRO + i) * 13] = sq(prm->noise_process_gyro); }
    for (size_t i = 0; i < (size_t)N * N; i++) e->P[i] *= e->noiseScale;
    for (int i = 0; i < 144; i++) e->Q[i] *= e->noiseScale;
    for (int i = 0; i < 400; i++) e->dydx[i] = 0.0;
    return e;
}
void orc_ekf_destroy(orc_ekf* e) { if (!e) return; free(e->m); free(e->P); free(e->Q); free(e->augmentTimes); free(e); }
orc_ekf* orc_ekf_clone(const orc_ekf* s)
{
    orc_ekf* e = orc_ekf_create(&s->prm);
    memcpy(e->m, s->m, sizeof(double) * s->N); memcpy(e->P, s->P, sizeof(double) * s->N * s->N); memcpy(e->Q, s->Q, sizeof(double) * 144);
    memcpy(e->dydx, s->dydx, sizeof(e->dydx)); memcpy(e->augmentTimes, s->augmentTimes, sizeof(double) * (s->trail + 2));
    e->augmentCount = s->augmentCount; e->nAugTimes = s->nAugTimes; e->time = s->time; e->ZUPTtime = s->ZUPTtime; e->ZRUPTtime = s->ZRUPTtime;
    e->initZUPTtime = s->initZUPTtime; e->wasStationary = s->wasStationary; e->prevSampleT = s->prevSampleT; e->firstSampleT = s->firstSampleT;
    e->firstSample = s->firstSample;
    return e;
}
int orc_ekf_state_dim(const orc_ekf* e) { return e->N; }
int orc_ekf_pose_count(const orc_ekf* e) { return e->augmentCount + 1; }
double orc_ekf_platform_time(const orc_ekf* e) { return e->firstSampleT + e->time; }
double orc_ekf_history_time(const orc_ekf* e, int i) { return i == -1 ? orc_ekf_platform_time(e) : e->augmentTimes[e->nAugTimes - i - 1]; }
int orc_ekf_was_stationary(const orc_ekf* e) { return e->wasStationary; }
void orc_ekf_set_first_sample_time(orc_ekf* e, double t) { e->firstSample = 0; e->firstSampleT = t; e->prevSampleT = t; e->time = t; }
void orc_ekf_upload(orc_ekf* e, const double* m, const double* P)
{
    if (m) memcpy(e->m, m, sizeof(double) * e->N);
    if (P) memcpy(e->P, P, sizeof(double) * e->N * e->N);
}
void orc_ekf_download(const orc_ekf* e, double* m, double* P)
{
    if (m) memcpy(m, e->m, sizeof(double) * e->N);
    if (P) memcpy(P, e->P, sizeof(double) * e->N * e->N);
}
void orc_ekf_download_inertial(const orc_ekf* e, double* m20, double* P20)
{
    memcpy(m20, e->m, sizeof(double) * 20);
    for (int j = 0; j < 20; j++) for (int i = 0; i < 20; i++) P20[i + j * 20] = Pm(i, j);
}
void orc_ekf_set_inertial_state(orc_ekf* e, const double* m20, const double* P20)
{
    memcpy(e->m, m20, sizeof(double) * 20);
    for (int j = 0; j < 20; j++) for (int i = 0; i < 20; i++) Pm(i, j) = P20[i + j * 20];
    e->augmentCount = 0; e->nAugTimes = 0;
}
void orc_ekf_set_process_noise(orc_ekf* e, const double* Q) { memcpy(e->Q, Q, sizeof(double) * 144); }
void orc_ekf_get_dydx(const orc_ekf* e, double* d) { memcpy(d, e->dydx, sizeof(double) * 400); }

static void normalize4(double* q)
{
    double z = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (z > 0) { double n = sqrt(z); for (int i = 0; i < 4; i++) q[i] /= n; }
}
void orc_ekf_normalize_quaternions(orc_ekf* e, int onlyCurrent)
{
    normalize4(e->m + ORI);
    if (onlyCurrent) return;
    for (int i = 0; i < e->trail; i++) normalize4(e->m + CAM + POSE * i + 3);
}
void orc_ekf_symmetrize(orc_ekf* e)
{
    for (int j = 0; j < e->N; j++) for (int i = j + 1; i < e->N; i++) { double s = 0.5 * (Pm(i, j) + Pm(j, i)); Pm(i, j) = s; Pm(j, i) = s; }
}

void orc_ekf_initialize_orientation(orc_ekf* e, const double* xa)
{   /* Eigen::Quaterniond::FromTwoVectors(-gravity, xa) */
    double nb = sqrt(xa[0] * xa[0] + xa[1] * xa[1] + xa[2] * xa[2]);
    double v0[3] = {0, 0, e->prm.gravity >= 0 ? 1.0 : -1.0}, v1[3] = {xa[0] / nb, xa[1] / nb, xa[2] / nb};
    double c = v0[2] * v1[2], q[4];
    if (c < -1.0 + 1e-12) { if (c < -1.0) c = -1.0; double w2 = (1.0 + c) * 0.5; q[0] = sqrt(w2); q[1] = sqrt(1.0 - w2); q[2] = 0; q[3] = 0; }
    else {
        double ax[3] = {v0[1] * v1[2] - v0[2] * v1[1], v0[2] * v1[0] - v0[0] * v1[2], v0[0] * v1[1] - v0[1] * v1[0]};
        double s = sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
        q[0] = s * 0.5; q[1] = ax[0] * invs; q[2] = ax[1] * invs; q[3] = ax[2] * invs;
    }
    for (int i = 0; i < 4; i++) e->m[ORI + i] = q[i];
    double var = sq(e->prm.noise_initial_ori) * e->noiseScale;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Pm(ORI + i, ORI + j) = (i == j && i < 3) ? var : 0.0;
}

static void quat2rmat_d(const double* q, double R[9], double dR[4][9])
{   /* row-major 3x3; src/odometry/util.cpp:10-47 */
    R[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3]; R[1] = 2 * q[1] * q[2] - 2 * q[0] * q[3]; R[2] = 2 * q[1] * q[3] + 2 * q[0] * q[2];
    R[3] = 2 * q[1] * q[2] + 2 * q[0] * q[3]; R[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3]; R[5] = 2 * q[2] * q[3] - 2 * q[0] * q[1];
    R[6] = 2 * q[1] * q[3] - 2 * q[0] * q[2]; R[7] = 2 * q[2] * q[3] + 2 * q[0] * q[1]; R[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
    double a = 2 * q[0], b = 2 * q[1], c = 2 * q[2], d = 2 * q[3];
    double t0[9] = {a, -d, c, d, a, -b, -c, b, a}, t1[9] = {b, c, d, c, -b, -a, d, a, -b};
    double t2[9] = {-c, b, a, b, c, d, -a, d, -c}, t3[9] = {-d, -a, b, a, -d, c, b, c, d};
    memcpy(dR[0], t0, sizeof(t0)); memcpy(dR[1], t1, sizeof(t1)); memcpy(dR[2], t2, sizeof(t2)); memcpy(dR[3], t3, sizeof(t3));
}

/* 4x4 matrix exponential of S = c * Omega(w) by scaling-and-squaring of a Taylor series (independent of the
 * closed form the CUDA kernel uses; the reference calls Eigen's Pade-based MatrixFunctions exp, ekf.cpp:425) */
static void expm4(const double* S, double* A)
{
    double nrm = 0; for (int i = 0; i < 16; i++) nrm = fmax(nrm, fabs(S[i]));
    int sqn = 0; double sc = 1.0; while (nrm * sc * 4 > 0.25) { sc *= 0.5; sqn++; }
    double X[16], term[16], tmp[16];
    for (int i = 0; i < 16; i++) { X[i] = S[i] * sc; A[i] = (i % 5 == 0) ? 1.0 : 0.0; term[i] = A[i]; }
    for (int k = 1; k <= 18; k++) {
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int r = 0; r < 4; r++) s += term[i * 4 + r] * X[r * 4 + j]; tmp[i * 4 + j] = s / k; }
        memcpy(term, tmp, sizeof(tmp));
        for (int i = 0; i < 16; i++) A[i] += term[i];
    }
    for (int s = 0; s < sqn; s++) {
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double v = 0; for (int r = 0; r < 4; r++) v += A[i * 4 + r] * A[r * 4 + j]; tmp[i * 4 + j] = v; }
        memcpy(A, tmp, sizeof(tmp));
    }
}

#define DX(i, j) e->dydx[(i) + (j) * 20]
#define DQ(i, j) dydq[(i) + (j) * 20]
void orc_ekf_predict(orc_ekf* e, double t, const double* xg, const double* xa)
{
    const orc_params* po = &e->prm;
    double dt = 0.0;
    if (!e->firstSample) { dt = t - e->prevSampleT; e->time = t - e->firstSampleT; } else { e->firstSampleT = t; e->firstSample = 0; }
    e->prevSampleT = t;
    if (dt <= 0.0) return;
    int N = e->N; double* m = e->m;
    double dydq[240];
    for (int i = 0; i < 400; i++) e->dydx[i] = (i % 21 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 240; i++) dydq[i] = 0.0;
    if (po->noise_process_baa > 0.0) {
        double v = e->noiseScale * sq(po->noise_process_baa), th = po->noise_process_baa_rev;
        if (th > 0.0) v *= (1 - exp(-2 * dt * th)) / (2 * th);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) e->Q[(Q_BAA_DRIFT + i) + (Q_BAA_DRIFT + j) * 12] = i == j ? v : 0.0;
    }
    if (po->noise_process_bga > 0.0) {
        double v = e->noiseScale * sq(po->noise_process_bga), th = po->noise_process_bga_rev;
        if (th > 0.0) v *= (1 - exp(-2 * dt * th)) / (2 * th);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) e->Q[(Q_BGA_DRIFT + i) + (Q_BGA_DRIFT + j) * 12] = i == j ? v : 0.0;
    }
    double w[3] = {xg[0] - m[BGA], xg[1] - m[BGA + 1], xg[2] - m[BGA + 2]};
    double S[16] = {0, -w[0], -w[1], -w[2], w[0], 0, -w[2], w[1], w[1], w[2], 0, -w[0], w[2], -w[1], w[0], 0};
    for (int i = 0; i < 16; i++) S[i] *= -dt / 2;
    double A[16]; expm4(S, A);
    double q[4] = {m[ORI], m[ORI + 1], m[ORI + 2], m[ORI + 3]}, qn[4];
    for (int i = 0; i < 4; i++) qn[i] = A[i * 4] * q[0] + A[i * 4 + 1] * q[1] + A[i * 4 + 2] * q[2] + A[i * 4 + 3] * q[3];
    double R[9], dR[4][9]; quat2rmat_d(qn, R, dR);
    for (int i = 0; i < 3; i++) m[POS + i] += m[VEL + i] * dt;
    double Txab[3]; for (int i = 0; i < 3; i++) Txab[i] = m[BAT + i] * xa[i] - m[BAA + i];
    double g[3] = {0, 0, -po->gravity};
    for (int i = 0; i < 3; i++) m[VEL + i] += (R[i] * Txab[0] + R[3 + i] * Txab[1] + R[6 + i] * Txab[2] + g[i]) * dt;
    for (int i = 0; i < 4; i++) m[ORI + i] = qn[i];
    if (po->noise_process_baa > 0.0) for (int i = 0; i < 3; i++) m[BAA + i] *= exp(-dt * po->noise_process_baa_rev);
    if (po->noise_process_bga > 0.0) for (int i = 0; i < 3; i++) m[BGA + i] *= exp(-dt * po->noise_process_bga_rev);
    for (int i = 0; i < 3; i++) DX(POS + i, VEL + i) = dt;
    double B[12];
    for (int qi = 0; qi < 4; qi++) for (int i = 0; i < 3; i++) B[i * 4 + qi] = (dR[qi][i] * Txab[0] + dR[qi][3 + i] * Txab[1] + dR[qi][6 + i] * Txab[2]) * dt;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += B[i * 4 + k] * A[k * 4 + j]; DX(VEL + i, ORI + j) = s; }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) DX(ORI + i, ORI + j) = A[i * 4 + j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) DQ(VEL + i, Q_ACC + j) = R[j * 3 + i] * dt;
    double h = dt / 2;
    double dS[3][16] = {{0, h, 0, 0, -h, 0, 0, 0, 0, 0, 0, h, 0, 0, -h, 0}, {0, 0, h, 0, 0, 0, 0, -h, -h, 0, 0, 0, 0, h, 0, 0}, {0, 0, 0, h, 0, 0, h, 0, 0, -h, 0, 0, -h, 0, 0, 0}};
    for (int j = 0; j < 3; j++) {
        double tq[4];
        for (int i = 0; i < 4; i++) tq[i] = dS[j][i * 4] * q[0] + dS[j][i * 4 + 1] * q[1] + dS[j][i * 4 + 2] * q[2] + dS[j][i * 4 + 3] * q[3];
        for (int i = 0; i < 4; i++) DQ(ORI + i, Q_GYRO + j) = A[i * 4] * tq[0] + A[i * 4 + 1] * tq[1] + A[i * 4 + 2] * tq[2] + A[i * 4 + 3] * tq[3];
    }
    for (int i = 0; i < 3; i++) { DQ(BGA + i, Q_BGA_DRIFT + i) = 1.0; DQ(BAA + i, Q_BAA_DRIFT + i) = 1.0; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 4; k++) s += DX(VEL + i, ORI + k) * DQ(ORI + k, Q_GYRO + j);
        DQ(VEL + i, Q_GYRO + j) = s; DX(VEL + i, BGA + j) = -s;
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 3; j++) DX(ORI + i, BGA + j) = -DQ(ORI + i, Q_GYRO + j);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { DX(VEL + i, BAA + j) = -R[j * 3 + i] * dt; DX(VEL + i, BAT + j) = R[j * 3 + i] * xa[j] * dt; }
    /* covariance, ekf.cpp:504-508 */
    double P00[400], T1[400], G1[240], newP00[400];
    for (int j = 0; j < 20; j++) for (int i = 0; i < 20; i++) P00[i + j * 20] = Pm(i, j);
    for (int j = 0; j < 20; j++) for (int i = 0; i < 20; i++) { double s = 0; for (int k = 0; k < 20; k++) s += DX(i, k) * P00[k + j * 20]; T1[i + j * 20] = s; }
    for (int j = 0; j < 12; j++) for (int i = 0; i < 20; i++) { double s = 0; for (int k = 0; k < 12; k++) s += DQ(i, k) * e->Q[k + j * 12]; G1[i + j * 20] = s; }
    for (int j = 0; j < 20; j++) for (int i = 0; i < 20; i++) {
        double s = 0, gq = 0;
        for (int k = 0; k < 20; k++) s += T1[i + k * 20] * DX(j, k);
        for (int k = 0; k < 12; k++) gq += G1[i + k * 20] * DQ(j, k);
        newP00[i + j * 20] = s + gq;
    }
    double row[20], col[20];
    for (int i = INER; i < N; i++) {
        for (int k = 0; k < 20; k++) row[k] = Pm(i, k);
        for (int j = 0; j < 20; j++) { double s = 0; for (int k = 0; k < 20; k++) s += row[k] * DX(j, k); Pm(i, j) = s; }
    }
    for (int c = INER; c < N; c++) {
        for (int k = 0; k < 20; k++) col[k] = Pm(k, c);
        for (int j = 0; j < 20; j++) { double s = 0; for (int k = 0; k < 20; k++) s += DX(j, k) * col[k]; Pm(j, c) = s; }
    }
    for (int j = 0; j < 20; j++) for (int i = 0; i < 20; i++) Pm(i, j) = newP00[i + j * 20];
}

/* Cholesky solve S X = B (S n x n SPD, B n x nrhs column-major, in place). Returns 0 on success. */
static int chol_solve(double* S, int n, double* B, int nrhs)
{
    for (int j = 0; j < n; j++) {
        for (int k = 0; k < j; k++) { double ljk = S[j + k * n]; for (int i = j; i < n; i++) S[i + j * n] -= S[i + k * n] * ljk; }
        if (!(S[j + j * n] > 0)) return 1;
        double d = sqrt(S[j + j * n]);
        for (int i = j; i < n; i++) S[i + j * n] /= d;
    }
    for (int c = 0; c < nrhs; c++) {
        double* b = B + (size_t)c * n;
        for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= S[i + k * n] * b[k]; b[i] = s / S[i + i * n]; }
        for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= S[k + i * n] * b[k]; b[i] = s / S[i + i * n]; }
    }
    return 0;
}

/* HP = H P[0:l,:] (n x N), S = HP[:,0:l] H' + Rdiag I.  H is n x l column-major. */
static void hp_and_s(const orc_ekf* e, const double* H, int n, int l, double Rdiag, double* HP, double* S)
{
    int N = e->N;
    for (int j = 0; j < N; j++) for (int i = 0; i < n; i++) { double s = 0; for (int k = 0; k < l; k++) s += H[i + k * n] * Pm(k, j); HP[i + (size_t)j * n] = s; }
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) { double s = 0; for (int k = 0; k < l; k++) s += HP[i + (size_t)k * n] * H[j + k * n]; S[i + j * n] = s + (i == j ? Rdiag : 0.0); }
}

/* generic update: K = (S^-1 HP)', m += K v, P -= K HP  (ekf.cpp:57-82 / 829-844). v given. */
static int kalman_update(orc_ekf* e, const double* H, int n, int l, double Rdiag, const double* v)
{
    int N = e->N;
    double* HP = (double*)malloc(sizeof(double) * n * N), *S = (double*)malloc(sizeof(double) * n * n), *W = (double*)malloc(sizeof(double) * n * N);
    hp_and_s(e, H, n, l, Rdiag, HP, S);
    memcpy(W, HP, sizeof(double) * n * N);
    int rc = chol_solve(S, n, W, N);              /* W = S^-1 HP = K' */
    if (!rc) {
        for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < n; k++) s += W[k + (size_t)i * n] * v[k]; e->m[i] += s; }
        for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < n; k++) s += W[k + (size_t)i * n] * HP[k + (size_t)j * n]; Pm(i, j) -= s; }
    }
    free(HP); free(S); free(W);
    return rc;
}

static void small_update(orc_ekf* e, int n, int l, const int* rowsel, const double* y, double Rdiag)
{   /* selector H: row i has a 1 at column rowsel[i]; update() then updateCommon normalises the current quaternion */
    double H[4 * 20]; memset(H, 0, sizeof(H));
    for (int i = 0; i < n; i++) H[i + rowsel[i] * n] = 1.0;
    double v[4];
    for (int i = 0; i < n; i++) { double s = 0; for (int k = 0; k < l; k++) s += H[i + k * n] * e->m[k]; v[i] = y[i] - s; }
    kalman_update(e, H, n, l, Rdiag, v);
    normalize4(e->m + ORI);
}

void orc_ekf_update_zupt(orc_ekf* e, double r)
{
    if (e->time - e->ZUPTtime < 0.25) return;
    e->ZUPTtime = e->time; e->wasStationary = 1;
    int sel[3] = {VEL, VEL + 1, VEL + 2}; double y[3] = {0, 0, 0};
    small_update(e, 3, VEL + 3, sel, y, r * e->noiseScale);
}
void orc_ekf_update_zupt_initialization(orc_ekf* e)
{
    if (e->wasStationary || e->time > 60 || e->time - e->initZUPTtime < 0.1) return;
    e->initZUPTtime = e->time;
    int sel[3] = {VEL, VEL + 1, VEL + 2}; double y[3] = {0, 0, 0};
    small_update(e, 3, VEL + 3, sel, y, e->prm.init_zupt_r * e->noiseScale * exp(0.5 * e->time));
}
void orc_ekf_update_zrupt(orc_ekf* e, const double* xg)
{
    if (e->time - e->ZRUPTtime < 0.25) return;
    e->ZRUPTtime = e->time;
    int sel[3] = {BGA, BGA + 1, BGA + 2};
    small_update(e, 3, BGA + 3, sel, xg, e->prm.rotation_zupt_r * e->noiseScale);
}
void orc_ekf_update_pseudo_velocity(orc_ekf* e, double defaultSpeed, double r)
{
    double h = sqrt(e->m[VEL] * e->m[VEL] + e->m[VEL + 1] * e->m[VEL + 1]);
    if (h <= 1e-7) return;
    double H[VEL + 2]; memset(H, 0, sizeof(H));
    H[VEL] = e->m[VEL] / h; H[VEL + 1] = e->m[VEL + 1] / h;
    double v = defaultSpeed - h;
    kalman_update(e, H, 1, VEL + 2, r * e->noiseScale, &v);
    normalize4(e->m + ORI);
}
void orc_ekf_update_position(orc_ekf* e, const double* y, double r)
{
    int sel[3] = {POS, POS + 1, POS + 2};
    small_update(e, 3, POS + 3, sel, y, r * e->noiseScale);
    orc_ekf_symmetrize(e);
}
void orc_ekf_update_zero_height(orc_ekf* e, double r)
{
    int sel[1] = {POS + 2}; double y[1] = {0};
    small_update(e, 1, POS + 3, sel, y, r * e->noiseScale);
    orc_ekf_symmetrize(e);
}
void orc_ekf_update_orientation(orc_ekf* e, const double* q, double r)
{
    int sel[4] = {ORI, ORI + 1, ORI + 2, ORI + 3};
    small_update(e, 4, ORI + 4, sel, q, r * e->noiseScale);
    orc_ekf_normalize_quaternions(e, 0);
    orc_ekf_symmetrize(e);
}

/* returns VuOutlierStatus (ekf.hpp:54-59): 0 INLIER, 2 RMSE, 3 CHI2; *chi2 = noiseScale v' S^-1 v */
int orc_ekf_visual_check(const orc_ekf* e, const double* H, int n, int l, const double* f, const double* y, double r, double rmseThr, double* chi2)
{
    double* v = (double*)malloc(sizeof(double) * n * 2), *sv = v + n;
    for (int i = 0; i < n; i++) v[i] = y[i] - f[i];
    if (chi2) *chi2 = 0.0;
    if (rmseThr >= 0.0) { double ss = 0; for (int i = 0; i < n; i++) ss += v[i] * v[i]; if (sqrt(ss / n) > rmseThr) { free(v); return 2; } }
    if (r < 0.0) { free(v); return 0; }
    double* HP = (double*)malloc(sizeof(double) * n * e->N), *S = (double*)malloc(sizeof(double) * n * n);
    hp_and_s(e, H, n, l, (r * r) * e->noiseScale, HP, S);
    memcpy(sv, v, sizeof(double) * n);
    chol_solve(S, n, sv, 1);
    double t = 0; for (int i = 0; i < n; i++) t += sv[i] * v[i];
    t *= e->noiseScale;
    if (chi2) *chi2 = t;
    int st = t > orc_chi2inv95(n) ? 3 : 0;
    free(HP); free(S); free(v);
    return st;
}
void orc_ekf_visual_update(orc_ekf* e, const double* H, int n, int l, const double* f, const double* y, double r)
{
    double* v = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) v[i] = y[i] - f[i];
    kalman_update(e, H, n, l, (r * r) * e->noiseScale, v);
    orc_ekf_normalize_quaternions(e, 0);
    free(v);
}

static int aug_src(int i, int drop) { if (i < CAM) return i; if (i < CAM + POSE) return -1; if (i < CAM + (drop + 1) * POSE) return i - POSE; return i; }

void orc_ekf_augment(orc_ekf* e, int drop)
{
    int N = e->N;
    if (drop == -1) drop = e->trail - 1;
    size_t NN = (size_t)N * N;
    double* P2 = (double*)malloc(sizeof(double) * NN), *m2 = (double*)malloc(sizeof(double) * N);
    for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) { int si = aug_src(i, drop), sj = aug_src(j, drop); P2[i + (size_t)j * N] = (si < 0 || sj < 0) ? 0.0 : Pm(si, sj); }
    for (int i = 0; i < N; i++) { int s = aug_src(i, drop); m2[i] = s < 0 ? 0.0 : e->m[s]; }
    memcpy(e->P, P2, sizeof(double) * NN); memcpy(e->m, m2, sizeof(double) * N);
    for (int i = 0; i < 3; i++) Pm(CAM + i, CAM + i) += sq(e->prm.noise_initial_pos_trail) * e->noiseScale;
    for (int i = 3; i < 7; i++) Pm(CAM + i, CAM + i) += sq(e->prm.noise_initial_ori_trail) * e->noiseScale;
    /* dense visAugH (7 x N), as the reference multiplies it */
    double* H = (double*)calloc((size_t)7 * N, sizeof(double));
    for (int i = 0; i < 3; i++) { H[i + (POS + i) * 7] = 1; H[i + (CAM + i) * 7] = -1; }
    for (int i = 0; i < 4; i++) { H[3 + i + (ORI + i) * 7] = 1; H[3 + i + (CAM + 3 + i) * 7] = -1; }
    double Rd = e->prm.augment_r * e->noiseScale;
    double* HP = (double*)malloc(sizeof(double) * 7 * N), S[49], *W = (double*)malloc(sizeof(double) * 7 * N);
    hp_and_s(e, H, 7, N, Rd, HP, S);
    memcpy(W, HP, sizeof(double) * 7 * N);
    chol_solve(S, 7, W, N);                                   /* K' */
    double v[7];
    for (int i = 0; i < 7; i++) { double s = 0; for (int k = 0; k < N; k++) s += H[i + k * 7] * e->m[k]; v[i] = -s; }
    for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < 7; k++) s += W[k + (size_t)i * 7] * v[k]; e->m[i] += s; }
    /* Joseph form exactly as updateCommonJosephForm: T1 = I - K H; T0 = T1 P; P = T0 T1' + K (R K') */
    double* T1 = (double*)malloc(sizeof(double) * NN), *T0 = (double*)malloc(sizeof(double) * NN);
    for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < 7; k++) s += W[k + (size_t)i * 7] * H[k + j * 7]; T1[i + (size_t)j * N] = (i == j ? 1.0 : 0.0) - s; }
    for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < N; k++) s += T1[i + (size_t)k * N] * Pm(k, j); T0[i + (size_t)j * N] = s; }
    for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) {
        double s = 0; for (int k = 0; k < N; k++) s += T0[i + (size_t)k * N] * T1[j + (size_t)k * N];
        double kr = 0; for (int k = 0; k < 7; k++) kr += W[k + (size_t)i * 7] * Rd * W[k + (size_t)j * 7];
        P2[i + (size_t)j * N] = s + kr;
    }
    memcpy(e->P, P2, sizeof(double) * NN);
    orc_ekf_symmetrize(e);
    orc_ekf_normalize_quaternions(e, 0);
    e->augmentTimes[e->nAugTimes++] = orc_ekf_platform_time(e);
    if (e->augmentCount < e->trail) e->augmentCount++;
    else { memmove(e->augmentTimes, e->augmentTimes + 1, sizeof(double) * (e->nAugTimes - 1)); e->nAugTimes--; }
    free(P2); free(m2); free(H); free(HP); free(W); free(T1); free(T0);
}

void orc_ekf_unaugment(orc_ekf* e)
{
    int N = e->N, ptd = N - e->mapDim;
    size_t NN = (size_t)N * N;
    double* P2 = (double*)malloc(sizeof(double) * NN), *m2 = (double*)malloc(sizeof(double) * N);
#define USRC(i) ((i) < CAM ? (i) : (i) >= ptd ? (i) : ((i) + POSE < ptd ? (i) + POSE : -1))
    for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) { int si = USRC(i), sj = USRC(j); P2[i + (size_t)j * N] = (si < 0 || sj < 0) ? 0.0 : Pm(si, sj); }
    for (int i = 0; i < N; i++) { int s = USRC(i); m2[i] = s < 0 ? 0.0 : e->m[s]; }
    memcpy(e->P, P2, sizeof(double) * NN); memcpy(e->m, m2, sizeof(double) * N);
    e->nAugTimes--; e->augmentCount--;
    free(P2); free(m2);
}

void orc_ekf_translate_to(orc_ekf* e, const double* pos)
{
    double d[3]; for (int k = 0; k < 3; k++) d[k] = pos[k] - e->m[POS + k];
    for (int k = 0; k < 3; k++) e->m[POS + k] += d[k];
    for (int i = 0; i < e->trail; i++) for (int k = 0; k < 3; k++) e->m[CAM + POSE * i + k] += d[k];
}

void orc_ekf_transform_to(orc_ekf* e, const double* pos, const double* q1, int pi)
{
    int N = e->N; size_t NN = (size_t)N * N;
    const double* q0 = pi < 0 ? e->m + ORI : e->m + CAM + POSE * pi + 3;
    const double* rp = pi < 0 ? e->m + POS : e->m + CAM + POSE * pi;
    double aw = q0[0], ax = -q0[1], ay = -q0[2], az = -q0[3], bw = q1[0], bx = q1[1], by = q1[2], bz = q1[3];
    double qc[4] = {aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx};
    double p1 = qc[0], p2 = qc[1], p3 = qc[2], p4 = qc[3];
    double Qm[16] = {p1, -p2, -p3, -p4, p2, p1, p4, -p3, p3, -p4, p1, p2, p4, p3, -p2, p1};
    double tx = 2 * qc[1], ty = 2 * qc[2], tz = 2 * qc[3], twx = tx * qc[0], twy = ty * qc[0], twz = tz * qc[0];
    double txx = tx * qc[1], txy = ty * qc[1], txz = tz * qc[1], tyy = ty * qc[2], tyz = tz * qc[2], tzz = tz * qc[3];
    double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    double Pc[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Pc[i * 3 + j] = R[j * 3 + i];
    double tr[3]; for (int i = 0; i < 3; i++) tr[i] = pos[i] - (Pc[i * 3] * rp[0] + Pc[i * 3 + 1] * rp[1] + Pc[i * 3 + 2] * rp[2]);
    double* A = (double*)calloc(NN, sizeof(double)), *T0 = (double*)malloc(sizeof(double) * NN), *m2 = (double*)malloc(sizeof(double) * N);
    for (int i = 0; i < N; i++) A[i + (size_t)i * N] = 1.0;
#define SETP(o) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[(o) + i + (size_t)((o) + j) * N] = Pc[i * 3 + j]
#define SETQ(o) for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) A[(o) + i + (size_t)((o) + j) * N] = Qm[i * 4 + j]
    SETP(POS); SETP(VEL); SETQ(ORI);
    for (int p = 0; p < e->trail; p++) { int o = CAM + p * POSE; SETP(o); SETQ(o + 3); }
    for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < N; k++) s += A[i + (size_t)k * N] * e->m[k]; m2[i] = s; }
    memcpy(e->m, m2, sizeof(double) * N);
    for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < N; k++) s += Pm(i, k) * A[j + (size_t)k * N]; T0[i + (size_t)j * N] = s; }
    for (int j = 0; j < N; j++) for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < N; k++) s += A[i + (size_t)k * N] * T0[k + (size_t)j * N]; Pm(i, j) = s; }
    double np[3]; for (int k = 0; k < 3; k++) np[k] = e->m[POS + k] + tr[k];
    orc_ekf_translate_to(e, np);
    free(A); free(T0); free(m2);
}

void orc_ekf_insert_map_point(orc_ekf* e, int idx, const double* pf)
{
    int N = e->N, off = N - e->mapDim + idx * MAPPT;
    for (int k = 0; k < 3; k++) for (int j = 0; j < N; j++) { Pm(off + k, j) = 0; Pm(j, off + k) = 0; }
    for (int k = 0; k < 3; k++) { Pm(off + k, off + k) = 1e3 * 1e3; e->m[off + k] = pf[k]; }
}

void orc_ekf_condition_on_last_pose(orc_ekf* e)
{
    int N = e->N, mm = N - POSE;
    double M[7][14];
    for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++) { M[i][j] = Pm(mm + i, mm + j); M[i][7 + j] = i == j; }
    for (int c = 0; c < 7; c++) {
        int p = c; for (int r = c + 1; r < 7; r++) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
        if (p != c) for (int j = 0; j < 14; j++) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
        double inv = 1.0 / M[c][c]; for (int j = 0; j < 14; j++) M[c][j] *= inv;
        for (int r = 0; r < 7; r++) if (r != c) { double f = M[r][c]; for (int j = 0; j < 14; j++) M[r][j] -= f * M[c][j]; }
    }
    double* T = (double*)malloc(sizeof(double) * mm * 7);
    for (int k = 0; k < 7; k++) for (int i = 0; i < mm; i++) { double s = 0; for (int r = 0; r < 7; r++) s += Pm(i, mm + r) * M[r][7 + k]; T[i + k * mm] = s; }
    for (int j = 0; j < mm; j++) for (int i = 0; i < mm; i++) { double s = 0; for (int k = 0; k < 7; k++) s += T[i + k * mm] * Pm(mm + k, j); Pm(i, j) -= s; }
    for (int k = 0; k < 7; k++) for (int i = 0; i < mm; i++) { Pm(i, mm + k) = 0; Pm(mm + k, i) = 0; }
    for (int i = 0; i < 7; i++) for (int j = 0; j < 7; j++) Pm(mm + i, mm + j) = i == j ? 1e3 * 1e3 : 0.0;
    free(T);
}

void orc_ekf_lock_biases(orc_ekf* e)
{
    for (int k = 0; k < 9; k++) for (int j = 0; j < e->N; j++) { Pm(BGA + k, j) = 0; Pm(j, BGA + k) = 0; }
}
