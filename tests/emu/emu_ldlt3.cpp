// tests/emu/emu_ldlt3.cpp -- the scalar 3x3 pivoted LDL^T / solve / rcond of track_model.cuh against the array version of the
// oracle (oracle/hv_oracle_tri.c, included here for its static functions) on 200k random symmetric matrices: SPD, badly scaled,
// tied diagonals, indefinite; all six pivot patterns occur. Test infrastructure.
#include "cuda_emu.h"
#include "track_model.cuh"
namespace orc {
#include "../../oracle/hv_oracle_tri.c"
}
int main() {
    srand(5); int bad = 0; double worst = 0, worstRc = 0; int pat[3][2] = {{0,0},{0,0},{0,0}};
    for (int t = 0; t < 200000; t++) {
        double B[9], A[9];
        for (double& x : B) x = rand() / (double)RAND_MAX - 0.5;
        const int kind = t % 4;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            double s = 0; for (int k = 0; k < 3; k++) s += B[3*i+k] * B[3*j+k];
            A[3*i+j] = kind == 3 ? 0.5 * (B[3*i+j] + B[3*j+i]) : s;        // SPD or symmetric indefinite
        }
        if (kind == 1) { const double sc[3] = {1e-3, 1.0, 1e3}; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[3*i+j] *= sc[(i + t) % 3] * sc[(j + t) % 3]; }
        if (kind == 2 && t % 8 == 2) { A[4] = A[0]; }                       // tie on the diagonal
        double rhs[3] = {rand() / (double)RAND_MAX, -rand() / (double)RAND_MAX, 0.3}, x1[3], x2[3];
        orc::ldlt3 Xo; orc::ldlt3_compute(A, &Xo); orc::ldlt3_solve(&Xo, rhs, x1);
        TmLdlt X; tm_ldlt(A, X); tm_solve(X, rhs, x2);
        pat[X.p0][X.p1]++;
        if (X.p0 != Xo.tp[0] || (X.p1 ? 2 : 1) != Xo.tp[1]) { bad++; if (bad < 5) printf("pivot mismatch t=%d: %d %d vs %d %d\n", t, X.p0, X.p1, Xo.tp[0], Xo.tp[1]); }
        double nx = 0, e = 0; for (int i = 0; i < 3; i++) { nx = fmax(nx, fabs(x1[i])); e = fmax(e, fabs(x1[i] - x2[i])); }
        // compare relative to cond: use residual instead
        double r = 0, nr = 0; for (int i = 0; i < 3; i++) { double s = -rhs[i]; for (int j = 0; j < 3; j++) s += A[3*i+j] * x2[j]; r = fmax(r, fabs(s)); double an = 0; for (int j = 0; j < 3; j++) an += fabs(A[3*i+j]) * fabs(x2[j]); nr = fmax(nr, an); }
        worst = fmax(worst, r / nr);
        const double rc1 = orc::ldlt3_rcond(&Xo), rc2 = tm_rcond(X);
        worstRc = fmax(worstRc, fabs(rc1 - rc2) / fmax(rc1, 1e-300));
    }
    printf("pivot patterns p0/p1: %d %d | %d %d | %d %d; pivot mismatches %d; worst backward error %.2e; worst relative rcond difference %.2e\n",
           pat[0][0], pat[0][1], pat[1][0], pat[1][1], pat[2][0], pat[2][1], bad, worst, worstRc);
    return bad != 0;
}
