/* Test infrastructure (not product code): the explicit inverse of the box QP's Cholesky factor (cholesky.c:51-74 in the reference),
 * in the two forms csrc/ilqg_kernels.cuh uses -- `seq`: one column after the other in one thread (box_qp<M>), `par`: column c by
 * lane c with the terms the reference's loops start behind skipped by predicate (box_qp<M, CL>, k_backpass_warp) -- on random
 * factors and clamp patterns.  Every entry of the free block must be bit-identical.  Build: g++ -O2 -ffp-contract=off.
 * Prints the number of differing entries per size; exit status 1 if any. */
#include <cstdio>
#include <cstring>
#include <cmath>
#include <random>
constexpr int utri(int r, int c) { return (c * (c + 1)) / 2 + r; }
template <int M> void seq(const double *U, const int *clamped, double *invH) {
    double w[M];
    for (int i = 0; i < M; i++) w[i] = 123.0;
    for (int col = 0; col < M; col++) {
        w[col] = 1.0;
        for (int k = col + 1; k < M; k++) w[k] = 0.0;
        for (int k = col; k < M; k++) {
            double wk = w[k];
            for (int i = col; i < k; i++) { const double t = wk - w[i] * U[utri(i, k)]; wk = clamped[i] ? wk : t; }
            w[k] = wk / U[utri(k, k)];
        }
        for (int k = M - 1; k >= col; k--) {
            double wk = w[k];
            for (int i = k + 1; i < M; i++) { const double t = wk - w[i] * U[utri(k, i)]; wk = clamped[i] ? wk : t; }
            wk = wk / U[utri(k, k)];
            w[k] = wk;
            invH[utri(col, k)] = wk;
        }
    }
}
template <int M> void par(const double *U, const int *clamped, double *invH) {
    double res[M][M];
    for (int lane = 0; lane < M; lane++) {
        const int col = lane;
        double wv[M];
        for (int k = 0; k < M; k++) {
            double wk = (k == col) ? 1.0 : 0.0;
            for (int i = 0; i < k; i++) { const double t = wk - wv[i] * U[utri(i, k)]; wk = (clamped[i] || i < col) ? wk : t; }
            wv[k] = wk / U[utri(k, k)];
        }
        for (int k = M - 1; k >= 0; k--) {
            double wk = wv[k];
            for (int i = k + 1; i < M; i++) { const double t = wk - wv[i] * U[utri(k, i)]; wk = clamped[i] ? wk : t; }
            wk = wk / U[utri(k, k)];
            wv[k] = wk;
        }
        for (int k = 0; k < M; k++) res[lane][k] = wv[k];
    }
    for (int c = 0; c < M; c++) for (int k = c; k < M; k++) invH[utri(c, k)] = res[c][k];
}
template <int M> long run(int n) {
    std::mt19937_64 g(7 + M); std::uniform_real_distribution<double> u(-2, 2);
    long bad = 0;
    for (int t = 0; t < n; t++) {
        double U[M * (M + 1) / 2]; int cl[M];
        for (int i = 0; i < M * (M + 1) / 2; i++) U[i] = u(g);
        for (int i = 0; i < M; i++) { U[utri(i, i)] = 0.1 + fabs(u(g)); cl[i] = (g() % 4 == 0) ? 1 + (int)(g() % 2) : 0; }
        if (t % 50 == 0) U[utri(g() % M, g() % M % M)] = NAN;   // don't-care entries may be NaN: only clamped rows/cols in the kernel, but test robustness of equality on unclamped outputs
        double a[M * (M + 1) / 2], b[M * (M + 1) / 2];
        seq<M>(U, cl, a); par<M>(U, cl, b);
        for (int c = 0; c < M; c++) for (int k = c; k < M; k++)
            if (!cl[c] && !cl[k] && memcmp(&a[utri(c, k)], &b[utri(c, k)], 8)) bad++;
    }
    return bad;
}
int main()
{
    const long b2 = run<2>(200000), b3 = run<3>(200000), b4 = run<4>(200000), b6 = run<6>(100000);
    printf("M=2 bad %ld\nM=3 bad %ld\nM=4 bad %ld\nM=6 bad %ld\n", b2, b3, b4, b6);
    return (b2 | b3 | b4 | b6) ? 1 : 0;
}
