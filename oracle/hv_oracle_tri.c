/* TEST INFRASTRUCTURE (never linked into, imported by or executed from the product path).
 *
 * Plain-C restatement of the per-track measurement model of HybVIO, the next hot-path row: what
 * Session::trackerVisualUpdate does for ONE track between "the EKF holds a pose trail" and "H, f go into the outlier
 * check" (src/odometry/backend.cpp:1050-1160):
 *
 *   pose trail of the observing cameras      extractCameraPoseTrail          src/odometry/triangulation.cpp:65-103
 *   two-view mid-point start + derivatives   triangulateWithTwoCameras       src/odometry/triangulation.cpp:610-710
 *     pseudo-inverse of the 3x2 ray matrix   pinv / dpinv                    src/odometry/triangulation.cpp:1000-1004, :32-51
 *   Gauss-Newton in inverse depth, with the derivative of every iterate w.r.t. every pose and the
 *   IMU-camera time shift                    Triangulator::triangulate       src/odometry/triangulation.cpp:120-407
 *   stereo: per-pose sum of both cameras     backend.cpp:1105-1116
 *   measurement Jacobian and prediction      prepareVisualUpdate (truncated) src/odometry/triangulation.cpp:897-987
 *   (x, y, z) <-> (x/z, y/z, 1/z)            inverseDepth                    src/odometry/triangulation.cpp:1006-1030
 *
 * Parameters are the reference defaults (oracle/_ref/gen/output/parameters.cpp:55-60): 10 Gauss-Newton iterations,
 * convergence threshold 1e-2, convergence R 11, rcond threshold 1e-8; useLinearTriangulation and
 * useIndependentStereoTriangulation off (their defaults).
 *
 * The 3x3 solves follow the published algorithms the reference gets from Eigen 3.3 (third party, vendored by the reference under
 * 3rdparty/mobile-cv-suite/eigen): LDL^T with diagonal pivoting (Eigen/src/Cholesky/LDLT.h) and the Hager / Higham 1-norm
 * condition estimate (Eigen/src/Core/ConditionEstimator.h), so that the BAD_COND decision is taken on the same number; the
 * pseudo-inverse is computed from a column-pivoted QR (rank threshold 2 eps like Eigen's completeOrthogonalDecomposition).
 *
 * Pinned against the reference's own code, compiled unmodified (oracle/_ref/libref_tri.so, oracle/ref_build/build_tri.sh), by
 * tests/test_oracle_tri.py and against tests/golden/tri_golden.npz generated from it (tests/golden/make_golden_tri.py).
 * All 3x3 matrices here are row-major; the 4x4 imuToCamera inputs and the outputs dpf and H are column-major like Eigen's.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { POS = 0, ORI = 6, SFT = 19, CAM = 20, POSE_DIM = 7 };
enum { TRI_OK = 0, TRI_HYBRID, TRI_BEHIND, TRI_BAD_COND, TRI_NO_CONVERGENCE, TRI_BAD_DEPTH, TRI_UNKNOWN_PROBLEM };   /* output.hpp:21-29 */
enum { PREPARE_VU_OK = 0, PREPARE_VU_ZERO_DEPTH = 1, PREPARE_VU_BEHIND = 2 };                                        /* output.hpp:15-19 */

static const unsigned GN_ITERATIONS = 10;
static const double CONVERGENCE_THRESHOLD = 1e-2, CONVERGENCE_R = 11.0, RCOND_THRESHOLD = 1e-8;

typedef struct { double p[3], R[9], base[3], dR[4][9]; } tri_pose;          /* CameraPose, triangulation.hpp */

/* ---- small dense helpers: C(ar x bc) = A(ar x ac) B(ac x bc), row-major ---- */
static void mm(const double* A, int ar, int ac, const double* B, int bc, double* C)
{
    for (int i = 0; i < ar; i++) for (int j = 0; j < bc; j++) {
        double s = 0; for (int k = 0; k < ac; k++) s += A[i * ac + k] * B[k * bc + j];
        C[i * bc + j] = s;
    }
}
static void tr(const double* A, int r, int c, double* At) { for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) At[j * r + i] = A[i * c + j]; }
static void mv3(const double* A, const double* x, double* y) { for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2]; }
static void mtv3(const double* A, const double* x, double* y) { for (int i = 0; i < 3; i++) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2]; }
static double nrm3(const double* x) { return sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); }

static void quat2rmat_d(const double* q, double R[9], double dR[4][9])
{   /* src/odometry/util.cpp:10-47 */
    R[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3]; R[1] = 2 * q[1] * q[2] - 2 * q[0] * q[3]; R[2] = 2 * q[1] * q[3] + 2 * q[0] * q[2];
    R[3] = 2 * q[1] * q[2] + 2 * q[0] * q[3]; R[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3]; R[5] = 2 * q[2] * q[3] - 2 * q[0] * q[1];
    R[6] = 2 * q[1] * q[3] - 2 * q[0] * q[2]; R[7] = 2 * q[2] * q[3] + 2 * q[0] * q[1]; R[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
    double a = 2 * q[0], b = 2 * q[1], c = 2 * q[2], d = 2 * q[3];
    double t0[9] = {a, -d, c, d, a, -b, -c, b, a}, t1[9] = {b, c, d, c, -b, -a, d, a, -b};
    double t2[9] = {-c, b, a, b, c, d, -a, d, -c}, t3[9] = {-d, -a, b, a, -d, c, b, c, d};
    memcpy(dR[0], t0, sizeof(t0)); memcpy(dR[1], t1, sizeof(t1)); memcpy(dR[2], t2, sizeof(t2)); memcpy(dR[3], t3, sizeof(t3));
}

/* triangulation.cpp:1006-1030 (first derivative only): ip = (x, y, 1) / z */
static void inverse_depth(const double* p, double* ip, double* dip)
{
    ip[0] = p[0] / p[2]; ip[1] = p[1] / p[2]; ip[2] = 1.0 / p[2];
    memset(dip, 0, 9 * sizeof(double));
    dip[0] = 1.0 / p[2]; dip[4] = 1.0 / p[2];
    for (int i = 0; i < 3; i++) dip[3 * i + 2] = -ip[i] / p[2];
}

/* triangulation.cpp:65-103. poses of camera 0 for every index, then (stereo) of camera 1. */
static void pose_trail(const double* m, const int* idx, int npose, int ncam, const double* T0, const double* T1, tri_pose* out)
{
    for (int cam = 0; cam < ncam; cam++) {
        const double* T = cam == 0 ? T0 : T1;                  /* column-major 4x4 */
        double Rc[9], R[9], dR[4][9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rc[3 * r + c] = T[4 * c + r];
        for (int k = 0; k < npose; k++) {
            tri_pose* o = &out[cam * npose + k];
            const int i = idx[k];
            const double* p = i == 0 ? m + POS : m + CAM + POSE_DIM * (i - 1);        /* historyPosition(i - 1), ekf.cpp */
            const double* q = i == 0 ? m + ORI : m + CAM + POSE_DIM * (i - 1) + 3;    /* historyOrientation(i - 1) */
            quat2rmat_d(q, R, dR);
            mm(Rc, 3, 3, R, 3, o->R);
            for (int j = 0; j < 4; j++) mm(Rc, 3, 3, dR[j], 3, o->dR[j]);
            for (int r = 0; r < 3; r++) o->base[r] = T[12 + r];
            double rb[3]; mtv3(o->R, o->base, rb);
            for (int r = 0; r < 3; r++) o->p[r] = p[r] - rb[r];
        }
    }
}

/* Moore-Penrose inverse of a 3x2 matrix (row-major A[3][2] -> iA[2][3]); triangulation.cpp:1000-1004 */
static void pinv32(const double* A, double* iA)
{
    double c[2][3] = {{A[0], A[2], A[4]}, {A[1], A[3], A[5]}};
    const int a = nrm3(c[1]) > nrm3(c[0]) ? 1 : 0, b = 1 - a;       /* column pivoting: the longer column first */
    const double r11 = nrm3(c[a]);
    double q1[3], q2[3], u[3];
    for (int i = 0; i < 3; i++) q1[i] = c[a][i] / r11;
    double r12 = q1[0] * c[b][0] + q1[1] * c[b][1] + q1[2] * c[b][2];
    for (int i = 0; i < 3; i++) u[i] = c[b][i] - r12 * q1[i];
    const double r12b = q1[0] * u[0] + q1[1] * u[1] + q1[2] * u[2];           /* one re-orthogonalisation step */
    for (int i = 0; i < 3; i++) u[i] -= r12b * q1[i];
    r12 += r12b;
    const double r22 = nrm3(u);
    if (r22 <= 2 * DBL_EPSILON * r11) {                                         /* rank 1: A = q1 [r11 r12] P^T */
        const double s = r11 * r11 + r12 * r12;
        for (int i = 0; i < 3; i++) { iA[3 * a + i] = r11 * q1[i] / s; iA[3 * b + i] = r12 * q1[i] / s; }
        return;
    }
    for (int i = 0; i < 3; i++) q2[i] = u[i] / r22;
    /* A P = Q R  ->  A^+ = P R^-1 Q^T */
    for (int i = 0; i < 3; i++) {
        iA[3 * b + i] = q2[i] / r22;
        iA[3 * a + i] = (q1[i] - r12 * q2[i] / r22) / r11;
    }
}

/* derivative of the pseudo-inverse (Golub & Pereyra 1973, eq. 4.12); triangulation.cpp:32-51 */
static void dpinv32(const double* A, const double* iA, const double* dA, double* diA)
{
    double iAT[6], dAT[6], t1[6], t2[6], AiA[9], iAA[4], G2[4], G3[9], P3[9], P2[4], u[6], w[6];
    tr(iA, 2, 3, iAT); tr(dA, 3, 2, dAT);
    mm(iA, 2, 3, dA, 2, iAA /* 2x2 tmp */); mm(iAA, 2, 2, iA, 3, t1);                 /* iA dA iA */
    mm(A, 3, 2, iA, 3, AiA); for (int i = 0; i < 9; i++) P3[i] = (i % 4 == 0 ? 1.0 : 0.0) - AiA[i];
    mm(iA, 2, 3, iAT, 2, G2); mm(G2, 2, 2, dAT, 3, u); mm(u, 2, 3, P3, 3, t2);        /* (iA iA') dA' (I - A iA) */
    mm(iA, 2, 3, A, 2, iAA); for (int i = 0; i < 4; i++) P2[i] = (i % 3 == 0 ? 1.0 : 0.0) - iAA[i];
    mm(iAT, 3, 2, iA, 3, G3); mm(P2, 2, 2, dAT, 3, u); mm(u, 2, 3, G3, 3, w);         /* (I - iA A) dA' (iA' iA) */
    for (int i = 0; i < 6; i++) diA[i] = -t1[i] + t2[i] + w[i];
}

/* triangulation.cpp:610-710. pf in the coordinates of pose0; dpf[15][3]: columns p0 (3), q0 (4), p1 (3), q1 (4), t. */
static void two_cameras(const tri_pose* P0, const tri_pose* P1, const double* ip0, const double* ip1, const double* vel0,
                        const double* vel1, int timeShift, double* pf, double (*dpf)[3])
{
    double R1T[9], C[9], d[3], b[3];
    tr(P1->R, 3, 3, R1T); mm(P0->R, 3, 3, R1T, 3, C);
    for (int i = 0; i < 3; i++) d[i] = P1->p[i] - P0->p[i];
    mv3(P0->R, d, b);
    const double v0[3] = {ip0[0], ip0[1], 1.0}, v1[3] = {ip1[0], ip1[1], 1.0};
    const double n0 = nrm3(v0), n1 = nrm3(v1);
    double vn0[3], vn1[3], Cv[3], A[6], iA[6];
    for (int i = 0; i < 3; i++) { vn0[i] = v0[i] / n0; vn1[i] = v1[i] / n1; }
    mv3(C, vn1, Cv);
    for (int i = 0; i < 3; i++) { A[2 * i] = vn0[i]; A[2 * i + 1] = -Cv[i]; }
    pinv32(A, iA);
    double s[2]; mm(iA, 2, 3, b, 1, s);
    for (int i = 0; i < 3; i++) pf[i] = s[0] * vn0[i];

    double dA[15][6], db[15][3];
    memset(dA, 0, sizeof(dA)); memset(db, 0, sizeof(db));
    for (int i = 0; i < 4; i++) {
        double dC0[9], dC1[9], dR1T[9], x[3];
        mm(P0->dR[i], 3, 3, R1T, 3, dC0);
        tr(P1->dR[i], 3, 3, dR1T); mm(P0->R, 3, 3, dR1T, 3, dC1);
        mv3(dC0, vn1, x); for (int r = 0; r < 3; r++) dA[3 + i][2 * r + 1] = -x[r];
        mv3(dC1, vn1, x); for (int r = 0; r < 3; r++) dA[10 + i][2 * r + 1] = -x[r];
        mv3(P0->dR[i], d, db[3 + i]);
        if (i < 3) for (int r = 0; r < 3; r++) { db[i][r] = -P0->R[3 * r + i]; db[7 + i][r] = P0->R[3 * r + i]; }
    }
    for (int i = 0; i < 14; i++) {
        double diA[6], x[2], y[2];
        dpinv32(A, iA, dA[i], diA);
        mm(iA, 2, 3, db[i], 1, x); mm(diA, 2, 3, b, 1, y);
        for (int r = 0; r < 3; r++) dpf[i][r] = (x[0] + y[0]) * vn0[r];
    }
    if (timeShift) {
        double B0[9], B1[9], w0[3] = {vel0[0], vel0[1], 0.0}, w1[3] = {vel1[0], vel1[1], 0.0}, x[3], y[3], diA[6], ds[2];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
            B0[3 * r + c] = ((r == c ? 1.0 : 0.0) - vn0[r] * vn0[c]) / n0;
            B1[3 * r + c] = ((r == c ? 1.0 : 0.0) - vn1[r] * vn1[c]) / n1;
        }
        mv3(B0, w0, x); mv3(B1, w1, y);
        double Cy[3]; mv3(C, y, Cy);
        for (int r = 0; r < 3; r++) { dA[14][2 * r] = x[r]; dA[14][2 * r + 1] = -Cy[r]; }
        dpinv32(A, iA, dA[14], diA);
        mm(diA, 2, 3, b, 1, ds);
        for (int r = 0; r < 3; r++) dpf[14][r] = s[0] * x[r] + vn0[r] * ds[0];
    } else {
        for (int r = 0; r < 3; r++) dpf[14][r] = 0.0;
    }
}

/* ---- 3x3 LDL^T with diagonal pivoting (Eigen/src/Cholesky/LDLT.h, lower variant, unblocked) ---- */
typedef struct { double M[9]; int tp[3]; double l1; } ldlt3;

static void ldlt3_compute(const double* Ain, ldlt3* X)
{
    double* M = X->M;
    memcpy(M, Ain, 9 * sizeof(double));
    X->l1 = 0;
    for (int c = 0; c < 3; c++) {                      /* max abs column sum of the self-adjoint matrix stored in the lower triangle */
        double s = 0;
        for (int r = c; r < 3; r++) s += fabs(M[3 * r + c]);
        for (int k = 0; k < c; k++) s += fabs(M[3 * c + k]);
        if (s > X->l1) X->l1 = s;
    }
    for (int k = 0; k < 3; k++) {
        int big = k;
        for (int i = k + 1; i < 3; i++) if (fabs(M[4 * i]) > fabs(M[4 * big])) big = i;
        X->tp[k] = big;
        if (big != k) {                                /* symmetric row / column exchange on the lower triangle */
            for (int c = 0; c < k; c++) { double t = M[3 * k + c]; M[3 * k + c] = M[3 * big + c]; M[3 * big + c] = t; }
            for (int r = big + 1; r < 3; r++) { double t = M[3 * r + k]; M[3 * r + k] = M[3 * r + big]; M[3 * r + big] = t; }
            { double t = M[4 * k]; M[4 * k] = M[4 * big]; M[4 * big] = t; }
            for (int i = k + 1; i < big; i++) { double t = M[3 * i + k]; M[3 * i + k] = M[3 * big + i]; M[3 * big + i] = t; }
        }
        double temp[3];
        for (int c = 0; c < k; c++) temp[c] = M[4 * c] * M[3 * k + c];
        for (int c = 0; c < k; c++) M[4 * k] -= M[3 * k + c] * temp[c];
        for (int r = k + 1; r < 3; r++) for (int c = 0; c < k; c++) M[3 * r + k] -= M[3 * r + c] * temp[c];
        const double akk = M[4 * k];
        if (k == 0 && !(fabs(akk) > 0)) { for (int j = 0; j < 3; j++) X->tp[j] = j; return; }
        if (fabs(akk) > 0) for (int r = k + 1; r < 3; r++) M[3 * r + k] /= akk;
    }
}

static void ldlt3_solve(const ldlt3* X, const double* rhs, double* x)
{
    const double* M = X->M;
    double v[3] = {rhs[0], rhs[1], rhs[2]};
    for (int k = 0; k < 3; k++) if (X->tp[k] != k) { double t = v[k]; v[k] = v[X->tp[k]]; v[X->tp[k]] = t; }
    for (int r = 1; r < 3; r++) for (int c = 0; c < r; c++) v[r] -= M[3 * r + c] * v[c];
    for (int i = 0; i < 3; i++) v[i] = fabs(M[4 * i]) > DBL_MIN ? v[i] / M[4 * i] : 0.0;        /* pseudo-inverse of D */
    for (int r = 1; r >= 0; r--) for (int c = r + 1; c < 3; c++) v[r] -= M[3 * c + r] * v[c];
    for (int k = 2; k >= 0; k--) if (X->tp[k] != k) { double t = v[k]; v[k] = v[X->tp[k]]; v[X->tp[k]] = t; }
    memcpy(x, v, sizeof(v));
}

/* reciprocal condition number in the 1-norm: Hager's estimator with Higham's alternating-sign safeguard
 * (Eigen/src/Core/ConditionEstimator.h); the decomposition is self-adjoint, so adjoint().solve == solve */
static double ldlt3_rcond(const ldlt3* X)
{
    if (X->l1 == 0) return 0;
    double v[3] = {1.0 / 3, 1.0 / 3, 1.0 / 3}, sgn[3], old_sgn[3] = {0, 0, 0};
    ldlt3_solve(X, v, v);
    double lower = fabs(v[0]) + fabs(v[1]) + fabs(v[2]), old_lower = lower;
    int jmax = -1, old_jmax = -1;
    for (int k = 0; k < 4; k++) {
        for (int i = 0; i < 3; i++) sgn[i] = v[i] < 0 ? -1.0 : 1.0;
        if (k > 0 && sgn[0] == old_sgn[0] && sgn[1] == old_sgn[1] && sgn[2] == old_sgn[2]) break;
        ldlt3_solve(X, sgn, v);
        jmax = 0; for (int i = 1; i < 3; i++) if (fabs(v[i]) > fabs(v[jmax])) jmax = i;
        if (jmax == old_jmax) break;
        double e[3] = {0, 0, 0}; e[jmax] = 1.0;
        ldlt3_solve(X, e, v);
        lower = fabs(v[0]) + fabs(v[1]) + fabs(v[2]);
        if (lower <= old_lower) break;
        memcpy(old_sgn, sgn, sizeof(sgn));
        old_jmax = jmax; old_lower = lower;
    }
    double a[3] = {1.0, -1.5, 2.0};
    ldlt3_solve(X, a, a);
    const double alt = 2 * (fabs(a[0]) + fabs(a[1]) + fabs(a[2])) / 9.0;
    const double inv = lower > alt ? lower : alt;
    return inv == 0 ? 0 : (1.0 / inv) / X->l1;
}

/* One term of the product rule used throughout triangulation.cpp:216-318: derivative of the residual block (2) and of its
 * Jacobian w.r.t. the inverse-depth point (2x3) given the derivative of C, t, the point and an additive term of the residual. */
static void d_error_block(const double* C, const double* t, const double* h, const double* pfiab, double rho, const double* dC,
                          const double* dt, const double* dq /* derivative of pfi (3) */, const double* extra, double* dErr, double* dE)
{
    const double dpfiab[3] = {dq[0], dq[1], 0.0};
    double a[3], b[3], dh[3];
    mv3(dC, pfiab, a); mv3(C, dpfiab, b);
    for (int r = 0; r < 3; r++) dh[r] = a[r] + b[r] + dq[2] * t[r] + rho * dt[r];
    const double ih2sq = 1.0 / (h[2] * h[2]);
    const double dih2 = -dh[2] / (h[2] * h[2]);
    const double dih2sq = -2 * dh[2] * ih2sq / h[2];
    for (int r = 0; r < 2; r++) {
        dErr[r] = extra[r] - dh[r] / h[2] - dih2 * h[r];
        for (int c = 0; c < 2; c++)
            dE[3 * r + c] = -dih2 * C[3 * r + c] + (-1 / h[2]) * dC[3 * r + c] + (dh[r] * ih2sq + dih2sq * h[r]) * C[6 + c] + h[r] * ih2sq * dC[6 + c];
        dE[3 * r + 2] = -dt[r] / h[2] - t[r] * dih2 + dh[r] * ih2sq * t[2] + h[r] * dih2sq * t[2] + h[r] * ih2sq * dt[2];
    }
}

/* Gauss-Newton triangulation with derivatives; triangulation.cpp:120-407. dpf: 3 x (7 n + 1) column-major (dpfi of the reference). */
static int triangulate(const tri_pose* trail, int n, int stereo, const double* ip, const double* vel, int timeShift, double* pf, double* dpfi)
{
    const int ind1 = stereo ? n / 2 - 1 : n - 1;
    const int dDim = n * POSE_DIM;
    double d2[15][3], dpfi_dpf[9], pfi[3];
    two_cameras(&trail[0], &trail[ind1], ip, ip + 2 * ind1, vel, vel + 2 * ind1, timeShift, pf, d2);
    inverse_depth(pf, pfi, dpfi_dpf);
    memset(dpfi, 0, sizeof(double) * 3 * (dDim + 1));
    for (int j = 0; j < POSE_DIM; j++) {
        mv3(dpfi_dpf, d2[j], dpfi + 3 * j);
        mv3(dpfi_dpf, d2[POSE_DIM + j], dpfi + 3 * (POSE_DIM * ind1 + j));
    }
    mv3(dpfi_dpf, d2[14], dpfi + 3 * dDim);

    double R0T[9]; tr(trail[0].R, 3, 3, R0T);
    const double* p0 = trail[0].p;
    double* dEerror = (double*)calloc((size_t)3 * (dDim + 1), sizeof(double));
    double* dETE = (double*)calloc((size_t)9 * (dDim + 1), sizeof(double));
    double rcond = 0, Jprev = 1e10;
    int converged = 0;
    static const double zero3[3] = {0, 0, 0}, zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (unsigned it = 0; it < GN_ITERATIONS; it++) {
        double ETE[9] = {0}, Eerror[3] = {0}, error2 = 0;
        memset(dEerror, 0, sizeof(double) * 3 * (dDim + 1));
        memset(dETE, 0, sizeof(double) * 9 * (dDim + 1));
        for (int i = 0; i < n; i++) {
            const tri_pose* cur = &trail[i];
            double C[9], t[3], dp[3], h[3], Ch[3], err[2], E[6];
            mm(cur->R, 3, 3, R0T, 3, C);
            for (int r = 0; r < 3; r++) dp[r] = p0[r] - cur->p[r];
            mv3(cur->R, dp, t);
            const double pfiab[3] = {pfi[0], pfi[1], 1.0};
            mv3(C, pfiab, Ch);
            for (int r = 0; r < 3; r++) h[r] = Ch[r] + pfi[2] * t[r];
            const double ih2sq = 1.0 / (h[2] * h[2]);
            for (int r = 0; r < 2; r++) {
                err[r] = ip[2 * i + r] - h[r] / h[2];
                for (int c = 0; c < 2; c++) E[3 * r + c] = (-1 / h[2]) * C[3 * r + c] + h[r] * ih2sq * C[6 + c];
                E[3 * r + 2] = -t[r] / h[2] + h[r] * ih2sq * t[2];
            }
            error2 += err[0] * err[0] + err[1] * err[1];
            for (int a = 0; a < 3; a++) {
                for (int b = 0; b < 3; b++) ETE[3 * a + b] += E[a] * E[b] + E[3 + a] * E[3 + b];
                Eerror[a] += E[a] * err[0] + E[3 + a] * err[1];
            }
            for (int j = 0; j <= dDim; j++) {
                double dC[9], dt[3], dErr[2], dE[6];
                const double* extra = zero3;
                if (j == dDim) {                                   /* the time-shift column: only the point moves, residual += velocity */
                    if (!timeShift) continue;
                    memcpy(dC, zero9, sizeof(dC)); memcpy(dt, zero3, sizeof(dt));
                    extra = vel + 2 * i;
                } else {
                    const int pose = j / POSE_DIM, comp = j % POSE_DIM;
                    const double *dRi = zero9, *dR0 = zero9;
                    double dp0[3] = {0, 0, 0}, dpi[3] = {0, 0, 0};
                    if (comp < 3) {
                        if (pose == i) dpi[comp] = 1;
                        if (pose == 0) dp0[comp] = 1;
                    } else {
                        if (pose == i) { dRi = cur->dR[comp - 3]; mtv3(dRi, cur->base, dpi); for (int r = 0; r < 3; r++) dpi[r] = -dpi[r]; }
                        if (pose == 0) { dR0 = trail[0].dR[comp - 3]; mtv3(dR0, trail[0].base, dp0); for (int r = 0; r < 3; r++) dp0[r] = -dp0[r]; }
                    }
                    double a[9], b[9], dR0T[9], x[3], y[3], dd[3];
                    mm(dRi, 3, 3, R0T, 3, a); tr(dR0, 3, 3, dR0T); mm(cur->R, 3, 3, dR0T, 3, b);
                    for (int r = 0; r < 9; r++) dC[r] = a[r] + b[r];
                    for (int r = 0; r < 3; r++) dd[r] = dp0[r] - dpi[r];
                    mv3(dRi, dp, x); mv3(cur->R, dd, y);
                    for (int r = 0; r < 3; r++) dt[r] = x[r] + y[r];
                }
                d_error_block(C, t, h, pfiab, pfi[2], dC, dt, dpfi + 3 * j, extra, dErr, dE);
                for (int a = 0; a < 3; a++) {
                    dEerror[3 * j + a] += (dE[a] * err[0] + dE[3 + a] * err[1]) + (E[a] * dErr[0] + E[3 + a] * dErr[1]);
                    for (int b = 0; b < 3; b++)
                        dETE[9 * j + 3 * a + b] += (dE[a] * E[b] + dE[3 + a] * E[3 + b]) + (E[a] * dE[b] + E[3 + a] * dE[3 + b]);
                }
            }
        }
        ldlt3 X; ldlt3_compute(ETE, &X);
        double step[3]; ldlt3_solve(&X, Eerror, step);
        for (int r = 0; r < 3; r++) pfi[r] += -step[r];
        for (int j = 0; j <= dDim; j++) {                          /* d(A^-1 b) = A^-1 db - A^-1 dA A^-1 b */
            double w[3], u[3], g[3];
            mv3(dETE + 9 * j, step, w); ldlt3_solve(&X, w, u);
            ldlt3_solve(&X, dEerror + 3 * j, g);
            for (int r = 0; r < 3; r++) dpfi[3 * j + r] += -g[r] - (-u[r]);
        }
        rcond = ldlt3_rcond(&X);
        const double J = 0.5 * error2 / (CONVERGENCE_R * CONVERGENCE_R);
        const double Jd = fabs((J - Jprev) / J);
        Jprev = J;
        if (Jd < CONVERGENCE_THRESHOLD) { converged = 1; break; }
    }
    free(dEerror); free(dETE);
    if (!converged) return TRI_NO_CONVERGENCE;
    if (rcond < RCOND_THRESHOLD) return TRI_BAD_COND;

    double dpf0_dpfi[9], pf0[3], rp[3], M[9];
    inverse_depth(pfi, pf0, dpf0_dpfi);
    mv3(R0T, pf0, rp);
    for (int r = 0; r < 3; r++) pf[r] = rp[r] + p0[r];
    if (pf[0] == p0[0] && pf[1] == p0[1] && pf[2] == p0[2]) return TRI_UNKNOWN_PROBLEM;
    mm(R0T, 3, 3, dpf0_dpfi, 3, M);
    for (int j = 0; j <= dDim; j++) {
        double x[3], y[3] = {0, 0, 0};
        mv3(M, dpfi + 3 * j, x);
        if (j >= 3 && j < POSE_DIM) mtv3(trail[0].dR[j - 3], pf0, y);       /* d(R0^T)/dq_k pf0 */
        for (int r = 0; r < 3; r++) dpfi[3 * j + r] = y[r] + x[r] + (j == r ? 1.0 : 0.0);
    }
    for (int i = 0; i < n; i++) {                                           /* isBehind, triangulation.cpp:53-59 */
        double d[3], a[3];
        for (int r = 0; r < 3; r++) d[r] = pf[r] - trail[i].p[r];
        mv3(trail[i].R, d, a);
        if (a[2] < 0) return TRI_BEHIND;
    }
    return TRI_OK;
}

/* triangulation.cpp:897-987 with truncated = true, mapPointOffset <= 0. dpf: 3 x (7 npose + 1) after the stereo sum. */
static int prepare_visual_update(const tri_pose* trail, int nobs, const int* idx, int npose, const double* pf, const double* dpf,
                                 int haveDerivatives, const double* vel, int timeShift, int* rows, int* cols, double* H, double* f)
{
    int end = 0;
    for (int k = 0; k < npose; k++) {
        const int ori = idx[k] == 0 ? ORI : CAM + POSE_DIM * (idx[k] - 1) + 3;
        const int pos = idx[k] == 0 ? POS : CAM + POSE_DIM * (idx[k] - 1);
        if (pos + 3 > end) end = pos + 3;
        if (ori + 4 > end) end = ori + 4;
    }
    const int R = 2 * nobs;
    *rows = R; *cols = end;
    memset(H, 0, sizeof(double) * R * end);
    memset(f, 0, sizeof(double) * R);
#define Hm(r, c) H[(size_t)(c) * R + (r)]
    for (int i = 0; i < nobs; i++) {
        const tri_pose* P = &trail[i];
        double pt[3], pfc[3], ipH[3], dipH[9], dipR[6];
        for (int r = 0; r < 3; r++) pt[r] = pf[r] - P->p[r];
        mv3(P->R, pt, pfc);
        if (pfc[2] == 0) return PREPARE_VU_ZERO_DEPTH;
        if (pfc[2] < 0) return PREPARE_VU_BEHIND;
        inverse_depth(pfc, ipH, dipH);
        f[2 * i] = ipH[0]; f[2 * i + 1] = ipH[1];
        mm(dipH, 2, 3, P->R, 3, dipR);                                 /* dip * R (2x3); dip = top 2 rows of dipHomog */
        const int k = i % npose;
        const int iPos = idx[k] == 0 ? POS : CAM + POSE_DIM * (idx[k] - 1), iOri = idx[k] == 0 ? ORI : iPos + 3;
        for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) Hm(2 * i + r, iPos + c) = -dipR[3 * r + c];
        for (int j = 0; j < 4; j++) {
            double a[3], b[3], c3[3], col[3];
            mv3(P->dR[j], pt, a); mtv3(P->dR[j], P->base, b); mv3(P->R, b, c3);
            for (int r = 0; r < 3; r++) col[r] = a[r] + c3[r];
            for (int r = 0; r < 2; r++) Hm(2 * i + r, iOri + j) = dipH[3 * r] * col[0] + dipH[3 * r + 1] * col[1] + dipH[3 * r + 2] * col[2];
        }
        if (haveDerivatives) {
            for (int j = 0; j < npose; j++) {
                const int jPos = idx[j] == 0 ? POS : CAM + POSE_DIM * (idx[j] - 1), jOri = idx[j] == 0 ? ORI : jPos + 3;
                for (int c = 0; c < POSE_DIM; c++) {
                    const double* d = dpf + 3 * (POSE_DIM * j + c);
                    const int col = c < 3 ? jPos + c : jOri + (c - 3);
                    for (int r = 0; r < 2; r++) Hm(2 * i + r, col) += dipR[3 * r] * d[0] + dipR[3 * r + 1] * d[1] + dipR[3 * r + 2] * d[2];
                }
            }
            if (timeShift) {
                const double* d = dpf + 3 * POSE_DIM * npose;
                for (int r = 0; r < 2; r++) Hm(2 * i + r, SFT) = (dipR[3 * r] * d[0] + dipR[3 * r + 1] * d[1] + dipR[3 * r + 2] * d[2]) - vel[2 * i + r];
            }
        }
    }
#undef Hm
    return PREPARE_VU_OK;
}

/* Same signature and meaning as ref_track_model (oracle/ref_build/ref_tri_shim.cpp). */
int orc_track_model(const double* m, int trail, int useStereo, const int* poseTrailIndex, int npose, const double* imuToCam,
                    const double* imuToCam2, const double* ip, const double* vel, int estimateTimeShift, int* triStatus, double* pf,
                    double* dpf, double* depth, int* vuStatus, int* rows, int* cols, double* H, double* f)
{
    (void)trail;
    const int ncam = useStereo ? 2 : 1, nobs = npose * ncam;
    tri_pose* tr_ = (tri_pose*)malloc(sizeof(tri_pose) * nobs);
    double* d = (double*)calloc((size_t)3 * (POSE_DIM * nobs + 1), sizeof(double));
    pose_trail(m, poseTrailIndex, npose, ncam, imuToCam, imuToCam2, tr_);
    const int st = triangulate(tr_, nobs, useStereo, ip, vel, estimateTimeShift, pf, d);
    memset(dpf, 0, sizeof(double) * 3 * (POSE_DIM * npose + 1));
    if (st == TRI_OK) {
        for (int j = 0; j < npose; j++) for (int c = 0; c < 3 * POSE_DIM; c++)                      /* backend.cpp:1105-1116 */
            dpf[3 * POSE_DIM * j + c] = d[3 * POSE_DIM * j + c] + (useStereo ? d[3 * POSE_DIM * (j + npose) + c] : 0.0);
        for (int r = 0; r < 3; r++) dpf[3 * POSE_DIM * npose + r] = d[3 * POSE_DIM * nobs + r];
    }
    double dd[3]; for (int r = 0; r < 3; r++) dd[r] = pf[r] - tr_[0].p[r];
    *depth = nrm3(dd);                                                                              /* backend.cpp:1095 */
    *triStatus = st; *vuStatus = -1; *rows = 0; *cols = 0;
    if (st == TRI_OK)
        *vuStatus = prepare_visual_update(tr_, nobs, poseTrailIndex, npose, pf, dpf, 1, vel, estimateTimeShift, rows, cols, H, f);
    free(tr_); free(d);
    return 0;
}

/* Entry points for the reference's own known-answer tests of the helpers (test/triangulation.cpp:477-485 "pinv",
 * :487-519 "triangulateWithTwoCameras"): row-major 3x2 -> 2x3; poses as p[3], R[9] row-major. */
void orc_pinv32(const double* A, double* iA) { pinv32(A, iA); }
void orc_two_cameras_point(const double* p0, const double* R0, const double* p1, const double* R1, const double* ip0, const double* ip1, double* pf)
{
    tri_pose a, b; double d[15][3];
    memset(&a, 0, sizeof(a)); memset(&b, 0, sizeof(b));
    memcpy(a.p, p0, sizeof(a.p)); memcpy(a.R, R0, sizeof(a.R)); memcpy(b.p, p1, sizeof(b.p)); memcpy(b.R, R1, sizeof(b.R));
    const double zero[2] = {0, 0};
    two_cameras(&a, &b, ip0, ip1, zero, zero, 0, pf, d);
}
/* test/util.cpp:9-57 "quat2rmat", "quat2rmat_d"; :97-109 "cond" (rcond_ldlt of the identity is 1) */
void orc_quat2rmat_d(const double* q, double* R, double* dR) { double d[4][9]; quat2rmat_d(q, R, d); memcpy(dR, d, sizeof(d)); }
double orc_rcond_ldlt3(const double* A) { ldlt3 X; ldlt3_compute(A, &X); return ldlt3_rcond(&X); }
