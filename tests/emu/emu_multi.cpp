// tests/emu/emu_multi.cpp -- the persistent sequence of measurements (ek2_multi_body of hybvio_b200/csrc/ekf_cluster2.cuh: one
// launch, the P block of every CTA stays in shared memory between the measurements) on the host emulator against the same
// measurements applied one by one through the C oracle. Test infrastructure.
#include "emu_cluster.h"
#include "ekf_cluster2.cuh"
namespace cg = cooperative_groups;

extern "C" {
struct orc_params { int camera_trail_length, hybrid_map_size; double v[20]; };
struct orc_ekf;
void orc_ekf_default_params(orc_params*);
orc_ekf* orc_ekf_create(const orc_params*);
void orc_ekf_destroy(orc_ekf*);
void orc_ekf_upload(orc_ekf*, const double*, const double*);
void orc_ekf_download(const orc_ekf*, double*, double*);
int orc_ekf_state_dim(const orc_ekf*);
double orc_chi2inv95(int);
int orc_ekf_visual_check(const orc_ekf*, const double*, int, int, const double*, const double*, double, double, double*);
void orc_ekf_visual_update(orc_ekf*, const double*, int, int, const double*, const double*, double);
}

struct MultiCtx { EkfUpdateArgs a; EkfMultiList list; };
#if !defined(EMU_CLUSTER_THREADS) || defined(EMU_AS_LIB)
EMU_CLUSTER_BODY(emu_multi_body) { const MultiCtx& c = *(const MultiCtx*)ctx; ek2_multi_body(c.a, c.list, dyn, cg::this_cluster()); }
#endif
#ifndef EMU_AS_LIB

static double rnd() { return rand() / (double)RAND_MAX - 0.5; }
static double gauss() { double s = 0; for (int i = 0; i < 12; i++) s += rand() / (double)RAND_MAX; return s - 6.0; }

struct Meas { int n, l, mode; double yscale, r, r2; };

int main()
{
    // the bench's per-frame sequence (check+update n = 8, 20, 40, 84, 8) with an outlier, a check-only, an update-only and a two-R item mixed in
    const Meas seq[] = {{8, 34, EKF_MODE_CHECK_UPDATE, 0.02, 0.05, 0}, {20, 55, EKF_MODE_CHECK_UPDATE, 0.02, 0.05, 0}, {40, 90, EKF_MODE_CHECK_UPDATE, 40.0, 0.05, 0},
                        {84, 160, EKF_MODE_CHECK_UPDATE, 0.02, 0.05, 0}, {8, 34, EKF_MODE_CHECK, 0.02, 0.05, 0}, {24, 97, EKF_MODE_CHECK_UPDATE, 0.02, 0.05, 0.004},
                        {13, 41, EKF_MODE_UPDATE, 0.02, 0.05, 0}, {8, 34, EKF_MODE_CHECK_UPDATE, 0.02, 0.05, 0}};
    const int cnt = (int)(sizeof(seq) / sizeof(seq[0]));
    srand(321);
    orc_params prm; orc_ekf_default_params(&prm); prm.camera_trail_length = 20;
    orc_ekf* o = orc_ekf_create(&prm);
    const int N = orc_ekf_state_dim(o);
    const double noiseScale = prm.v[0] * prm.v[0];
    emu::Arena arena((size_t)96 << 20);
    double* m = arena.alloc<double>(N); double* P = arena.alloc<double>((size_t)N * N);
    double* res = arena.alloc<double>(64); double* cwork = arena.alloc<double>((size_t)10 * N * N);
    double* slots = arena.alloc<double>(4 * cnt);
    {
        std::vector<double> Bm((size_t)N * N);
        for (auto& x : Bm) x = rnd();
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { double s = 0; for (int k = 0; k < N; k++) s += Bm[i + (size_t)k * N] * Bm[j + (size_t)k * N]; P[i + (size_t)j * N] = 0.05 * s + (i == j ? 0.5 : 0.0); }
        for (int i = 0; i < N; i++) m[i] = 0.3 * rnd();
        for (int p = 0; p <= 20; p++) {
            double* q = p == 0 ? m + EKF_ORI : m + EKF_CAM + EKF_POSE * (p - 1) + 3;
            double nn = 0; for (int i = 0; i < 4; i++) { q[i] = rnd() + (i == 0); nn += q[i] * q[i]; }
            for (int i = 0; i < 4; i++) q[i] /= std::sqrt(nn);
        }
    }
    orc_ekf_upload(o, m, P);
    EkfMultiList list; memset(&list, 0, sizeof(list));
    list.count = cnt;
    std::vector<int> est(cnt, 0); std::vector<double> echi(cnt, 0.0);
    for (int k = 0; k < cnt; k++) {
        const Meas& q = seq[k];
        double* H = arena.alloc<double>((size_t)q.n * q.l); double* f = arena.alloc<double>(q.n); double* y = arena.alloc<double>(q.n);
        for (size_t i = 0; i < (size_t)q.n * q.l; i++) H[i] = 0.1 * gauss();
        for (int i = 0; i < q.n; i++) { f[i] = 0.5 * gauss(); y[i] = f[i] + q.yscale * gauss(); }
        EkfMultiItem& it = list.it[k];
        it.H = H; it.f = f; it.y = y; it.n = q.n; it.l = q.l; it.mode = q.mode; it.skipChi2 = 0;
        it.Rdiag = q.r * q.r * noiseScale; it.Rdiag2 = q.r2 > 0 ? q.r2 * q.r2 * noiseScale : 0.0;
        it.chi2Thr = q.mode == EKF_MODE_UPDATE ? 0.0 : orc_chi2inv95(q.n); it.rmseThr = -1.0; it.slot = slots + 4 * k;
        // oracle, in order
        if (q.mode != EKF_MODE_UPDATE) est[k] = orc_ekf_visual_check(o, H, q.n, q.l, f, y, q.r, -1.0, &echi[k]);
        if (q.mode == EKF_MODE_UPDATE || (q.mode == EKF_MODE_CHECK_UPDATE && est[k] == 0)) orc_ekf_visual_update(o, H, q.n, q.l, f, y, q.r2 > 0 ? q.r2 : q.r);
    }
    EkfUpdateArgs a; memset(&a, 0, sizeof(a));
    a.b.m = m; a.b.P = P; a.b.res = res; a.b.cwork = cwork; a.b.N = N; a.b.trail = 20;
    a.op = EKF_OP_DENSE; a.noiseScale = noiseScale; a.normalizeAll = 1;
    const size_t smem = ek2_multi_smem_bytes(list, N, 8, &a.xCap, &a.tCap);
    MultiCtx mc{a, list};
    const int bad = EMU_LAUNCH_CLUSTER(arena, 8, EK2_NT, smem, emu_multi_body, &mc);
    int fails = bad;
    for (int k = 0; k < cnt; k++) {
        const bool checked = seq[k].mode != EKF_MODE_UPDATE;
        const bool ok = !checked || ((int)slots[4 * k] == est[k] && std::fabs(slots[4 * k + 1] - echi[k]) <= 1e-9 * std::fmax(1.0, std::fabs(echi[k])));
        printf("measurement %d (n=%2d l=%3d mode %d%s): status %d/%d chi2 %.6g/%.6g  %s\n", k, seq[k].n, seq[k].l, seq[k].mode, seq[k].r2 > 0 ? ", two-R" : "", (int)slots[4 * k], est[k],
               slots[4 * k + 1], echi[k], ok ? "ok" : "FAIL");
        fails += !ok;
    }
    std::vector<double> om(N), oP((size_t)N * N);
    orc_ekf_download(o, om.data(), oP.data());
    double em = 0, eP = 0, pmax = 0;
    for (int i = 0; i < N; i++) em = std::fmax(em, std::fabs(om[i] - m[i]));
    for (size_t i = 0; i < oP.size(); i++) { eP = std::fmax(eP, std::fabs(oP[i] - P[i])); pmax = std::fmax(pmax, std::fabs(oP[i])); }
    const bool ok = em < 1e-9 && eP / pmax < 1e-9;
    printf("persistent sequence of %d measurements, smem %.1f KB: max|dm| %.2e  max|dP|/max|P| %.2e  %s\n", cnt, smem / 1024.0, em, eP / pmax, ok ? "ok" : "FAIL");
    fails += !ok;
    orc_ekf_destroy(o);
    return fails;
}
#endif  // EMU_AS_LIB
