// tests/emu/emu_predict.cpp -- runs the REAL body of ekf_predict_kernel (hybvio_b200/csrc/ekf_predict.cuh) on the host
// emulator and compares it with the C oracle (oracle/hv_oracle_ekf.c: orc_ekf_predict, sample by sample).
//   g++ -std=c++20 -O1 -pthread -Itests/emu/stubs -Itests/emu -Ihybvio_b200/csrc tests/emu/emu_predict.cpp oracle/hv_oracle_ekf.c -o build/emu_predict
#include "cuda_emu.h"
#define EKF_PMARK(i) do { } while (0)
#include "ekf_predict.cuh"

extern "C" {
struct orc_params { int camera_trail_length, hybrid_map_size; double v[20]; };
struct orc_ekf;
void orc_ekf_default_params(orc_params*);
orc_ekf* orc_ekf_create(const orc_params*);
void orc_ekf_upload(orc_ekf*, const double*, const double*);
void orc_ekf_download(const orc_ekf*, double*, double*);
void orc_ekf_set_first_sample_time(orc_ekf*, double);
void orc_ekf_predict(orc_ekf*, double, const double*, const double*);
void orc_ekf_get_dydx(const orc_ekf*, double*);
void orc_ekf_normalize_quaternions(orc_ekf*, int);
int orc_ekf_state_dim(const orc_ekf*);
}

int main()
{
    int fails = 0;
    const int ntrials = 12;
    for (int trial = 0; trial < ntrials; trial++) {
        orc_params prm; orc_ekf_default_params(&prm);
        prm.camera_trail_length = (trial % 3) == 2 ? 6 : 20;
        // v[] = hv_ekf_params after the two ints: noise_scale 0, gravity 1, noise_initial_* 2..10, noise_process_acc 11, gyro 12, baa 13, baa_rev 14, bga 15, bga_rev 16
        if ((trial % 3) == 1) prm.v[15] = 2e-5;       // gyro-bias random walk on: exercises bgaDecay / qBga
        orc_ekf* o = orc_ekf_create(&prm);
        const int N = orc_ekf_state_dim(o);
        std::vector<double> m(N), P((size_t)N * N), Q(144, 0.0);
        srand(7 + trial);
        auto rnd = [] { return rand() / (double)RAND_MAX - 0.5; };
        // random SPD covariance and a plausible mean
        std::vector<double> Bm((size_t)N * N);
        for (auto& x : Bm) x = rnd();
        for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { double s = 0; for (int k = 0; k < N; k++) s += Bm[i + (size_t)k * N] * Bm[j + (size_t)k * N]; P[i + (size_t)j * N] = 0.01 * s; }
        for (int i = 0; i < N; i++) m[i] = 0.3 * rnd();
        double qn = 0; for (int i = 0; i < 4; i++) { m[6 + i] = rnd() + (i == 0); qn += m[6 + i] * m[6 + i]; }
        for (int i = 0; i < 4; i++) m[6 + i] /= std::sqrt(qn);
        for (int i = 0; i < 3; i++) m[16 + i] = 1.0 + 0.01 * rnd();
        orc_ekf_upload(o, m.data(), P.data());
        orc_ekf_set_first_sample_time(o, 0.0);
        // device-side copies
        std::vector<double> dm = m, dP = P, dQ(144), ddydx(400), res(64);
        // Q as the oracle holds it after construction: read it through one predict on a clone is overkill -- rebuild (ekf.cpp:224-236)
        const double ns = prm.v[0] * prm.v[0];
        for (int i = 0; i < 3; i++) { dQ[(0 + i) * 13] = ns * prm.v[11] * prm.v[11]; dQ[(3 + i) * 13] = ns * prm.v[12] * prm.v[12]; }
        EkfPredictArgs a; memset(&a, 0, sizeof(a));
        a.b.m = dm.data(); a.b.P = dP.data(); a.b.Q = dQ.data(); a.b.dydx = ddydx.data(); a.b.res = res.data(); a.b.N = N; a.b.trail = prm.camera_trail_length;
        a.gravity = prm.v[1];
        const int cnt = trial == 2 ? 3 : trial < 3 ? 10 : 1 + (trial * 5) % 16;
        a.count = cnt;
        double t = 0.0;
        for (int k = 0; k < cnt; k++) {
            const double dt = 0.005 + 0.0002 * k;
            t += dt;
            double xg[3] = {0.05 * rnd(), 0.05 * rnd(), 0.2 + 0.05 * rnd()}, xa[3] = {0.3 * rnd(), 0.2 * rnd(), 9.8 + 0.2 * rnd()};
            EkfPredictSample& s = a.s[k];
            s.dt = dt; for (int i = 0; i < 3; i++) { s.xg[i] = xg[i]; s.xa[i] = xa[i]; }
            s.qBaa = s.qBga = -1.0; s.baaDecay = s.bgaDecay = 1.0; s.normAfter = ((trial % 3) >= 1 && trial < 6 && k != 1) ? 1 : 0; s.pad = 0;
            if (prm.v[13] > 0) { const double th = prm.v[14]; s.qBaa = ns * prm.v[13] * prm.v[13]; if (th > 0) s.qBaa *= (1 - std::exp(-2 * dt * th)) / (2 * th); s.baaDecay = std::exp(-dt * th); }
            if (prm.v[15] > 0) { const double th = prm.v[16]; s.qBga = ns * prm.v[15] * prm.v[15]; if (th > 0) s.qBga *= (1 - std::exp(-2 * dt * th)) / (2 * th); s.bgaDecay = std::exp(-dt * th); }
            orc_ekf_predict(o, t, xg, xa);
            if (s.normAfter) orc_ekf_normalize_quaternions(o, 1);
        }
        std::vector<double> dyn(ekf_predict_smem_bytes(cnt) / 8);
        // mean-only launch first (hv_ekf_predicted_mean_device): must leave the state alone and produce the bits the full launch will
        std::vector<double> mean(20, -7.0), m0 = dm, P0 = dP;
        { EkfPredictArgs am = a; am.meanOut = mean.data(); emu::launch_cta(EKF_NT, 0, [&] { ekf_predict_body(am, dyn.data()); }); }
        bool meanOk = dm == m0 && dP == P0;
        emu::launch_cta(EKF_NT, 0, [&] { ekf_predict_body(a, dyn.data()); });
        for (int i = 0; i < 20; i++) meanOk = meanOk && mean[i] == dm[i];
        std::vector<double> om(N), oP((size_t)N * N), od(400);
        orc_ekf_download(o, om.data(), oP.data()); orc_ekf_get_dydx(o, od.data());
        double em = 0, eP = 0, pmax = 0, ed = 0;
        for (int i = 0; i < N; i++) em = std::fmax(em, std::fabs(om[i] - dm[i]));
        for (size_t i = 0; i < oP.size(); i++) { eP = std::fmax(eP, std::fabs(oP[i] - dP[i])); pmax = std::fmax(pmax, std::fabs(oP[i])); }
        for (int i = 0; i < 400; i++) ed = std::fmax(ed, std::fabs(od[i] - ddydx[i]));
        const bool ok = em < 1e-12 && eP / pmax < 1e-12 && ed < 1e-12 && meanOk;
        printf("trial %d N=%d cnt=%d: max|dm| %.3e  max|dP|/max|P| %.3e  max|d dydx| %.3e  %s\n", trial, N, cnt, em, eP / pmax, ed, ok ? "ok" : "FAIL");
        fails += !ok;
    }
    return fails;
}
