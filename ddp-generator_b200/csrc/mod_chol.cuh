/* mod_chol.cuh -- the reference's modified-Cholesky family as a __device__ unit.
 *
 * Restates mod_chol / mod_chol_solve / mod_chol_inv / perm_tri_square of cholesky.c:129-356 (a pivoted Cholesky factorisation
 * that adds a diagonal E when the matrix is not sufficiently positive definite, after Schnabel & Eskow) on packed upper
 * triangles, operation for operation, so that the results are bit-identical to the reference's functions on the same inputs
 * (tests/test_mod_chol.py: known answers recorded from the reference).  It is NOT wired into the solver: the reference only
 * reaches it with -DMOD_CHOL=1, and its call site (boxQP.c:69-72) overwrites H with the regularised matrix while handing the
 * factor to scratch memory that the next statement reuses, so there are no defined solver semantics to match.
 *
 * Faithfully kept quirks of the reference: in the final 2 x 2 block it addresses the off-diagonal element as
 * UTRI_MAT_IDX(n-1, n-2), i.e. with row > column, which in the packed layout is element (0, n-1) (cholesky.c:281-297); tau and
 * taubar are pow(eps, 1/3) and pow(eps, 2/3) as glibc evaluates them (bit patterns below), not recomputed on the device.
 */
#pragma once
#include "dm_math.h"

namespace ilqg {

/* raw packed-upper-triangle index of the reference's macro (matMult.h:8): NOT symmetrised, rows may exceed columns */
__host__ __device__ constexpr int mc_idx(int r, int c) { return (c * (c + 1)) / 2 + r; }

__host__ __device__ inline double mc_max(double a, double b) { return (a > b) ? a : b; }   /* the reference's own fmax, cholesky.c:76-81 */

__host__ __device__ inline void mc_swap_rc(double *A, int n, int i, int j)   /* i > j: exchange rows and columns i and j */
{
    for (int k = 0; k < j; k++) {
        const double t = A[mc_idx(k, j)];
        A[mc_idx(k, j)] = A[mc_idx(k, i)];
        A[mc_idx(k, i)] = t;
    }
    for (int k = j + 1; k < i; k++) {
        const double t = A[mc_idx(j, k)];
        A[mc_idx(j, k)] = A[mc_idx(k, i)];
        A[mc_idx(k, i)] = t;
    }
    for (int k = i + 1; k < n; k++) {
        const double t = A[mc_idx(j, k)];
        A[mc_idx(j, k)] = A[mc_idx(i, k)];
        A[mc_idx(i, k)] = t;
    }
    const double t = A[mc_idx(j, j)];
    A[mc_idx(j, j)] = A[mc_idx(i, i)];
    A[mc_idx(i, i)] = t;
}

__host__ __device__ inline void mc_eliminate(double *A, int n, int j)   /* j-th step of the outer-product factorisation */
{
    A[mc_idx(j, j)] = sqrt(A[mc_idx(j, j)]);
    for (int i = j + 1; i < n; i++) {
        A[mc_idx(j, i)] /= A[mc_idx(j, j)];
        for (int k = j + 1; k <= i; k++) A[mc_idx(k, i)] -= A[mc_idx(j, i)] * A[mc_idx(j, k)];
    }
}

/* A (packed upper triangle, n x n) is replaced by the factor of P'(A + diag(E))P; returns the last diagonal shift */
__host__ __device__ inline double mod_chol(double *A, int n, double *E, int *P, double *g)
{
    const double tau = dm_from_bits(0x3ed965fea53d6e3eull), taubar = dm_from_bits(0x3dc428a2f98d728dull), mu = 0.1;
    bool phase_one = true;
    double gamma = 0.0, delta = 0.0, deltaprev = 0.0;
    int j;
    if (n == 1) {
        delta = (taubar * fabs(A[0])) - A[0];
        E[0] = (delta > 0.0) ? delta : 0.0;
        if (A[0] == 0.0) E[0] = taubar;
        A[0] = sqrt(A[0] + E[0]);
        P[0] = 0;
        return E[0];
    }
    for (int i = 0; i < n; i++) {
        P[i] = i;
        E[i] = 0.0;
    }
    for (int i = 0; i < n; i++) {
        const double d = fabs(A[mc_idx(i, i)]);
        if (d > gamma) gamma = d;
        if (A[mc_idx(i, i)] < 0.0) phase_one = false;
    }
    /* phase one: plain pivoted Cholesky while the remaining diagonal stays sufficiently positive */
    j = 0;
    while (j < n && phase_one) {
        double dmax_ = A[mc_idx(j, j)], dmin_ = A[mc_idx(j, j)];
        int id = j;
        for (int i = j + 1; i < n; i++) {
            if (dmax_ < A[mc_idx(i, i)]) {
                dmax_ = A[mc_idx(i, i)];
                id = i;
            }
            if (dmin_ > A[mc_idx(i, i)]) dmin_ = A[mc_idx(i, i)];
        }
        if (dmax_ < taubar * gamma || dmin_ < -mu * dmax_) {
            phase_one = false;
            break;
        }
        if (id != j) {
            mc_swap_rc(A, n, id, j);
            const int t = P[id];
            P[id] = P[j];
            P[j] = t;
        }
        double low = 0.0;
        for (int i = j + 1; i < n; i++) {
            const double t = A[mc_idx(i, i)] - A[mc_idx(j, i)] * A[mc_idx(j, i)] / A[mc_idx(j, j)];
            if (low > t) low = t;
        }
        if (low < -mu * gamma) {
            phase_one = false;
            break;
        }
        mc_eliminate(A, n, j);
        j++;
    }
    /* phase two: not positive definite */
    if (!phase_one && j == n - 1) {
        delta = -A[mc_idx(n - 1, n - 1)] + mc_max(tau * A[mc_idx(n - 1, n - 1)] / (tau - 1.), taubar * gamma);
        A[mc_idx(n - 1, n - 1)] += delta;
        A[mc_idx(n - 1, n - 1)] = sqrt(A[mc_idx(n - 1, n - 1)]);
        E[n - 1] = delta;
        deltaprev = delta;
    }
    if (!phase_one && j < n - 1) {
        const int k = j - 1;   /* steps done in phase one, minus one */
        for (int i = k + 1; i < n; i++) {   /* lower Gerschgorin bounds */
            g[i] = A[mc_idx(i, i)];
            for (int c = k + 1; c <= i - 1; c++) g[i] -= fabs(A[mc_idx(c, i)]);
            for (int c = i + 1; c < n; c++) g[i] -= fabs(A[mc_idx(i, c)]);
        }
        for (j = k + 1; j < n - 2; j++) {
            int id = j;
            double best = g[id];
            for (int i = j + 1; i < n; i++)
                if (best < g[i]) {
                    best = g[i];
                    id = i;
                }
            if (id != j) {
                mc_swap_rc(A, n, id, j);
                const int t = P[id];
                P[id] = P[j];
                P[j] = t;
                const double tg = g[id];
                g[id] = g[j];
                g[j] = tg;
            }
            double normj = 0.;
            for (int i = j + 1; i < n; i++) normj += fabs(A[mc_idx(j, i)]);
            delta = mc_max(0.0, mc_max(mc_max(normj, taubar * gamma) - A[mc_idx(j, j)], deltaprev));
            if (delta > 0) {
                A[mc_idx(j, j)] += delta;
                deltaprev = delta;
                E[j] = delta;
            }
            if (A[mc_idx(j, j)] != normj) {
                const double t = 1.0 - normj / A[mc_idx(j, j)];
                for (int i = j + 1; i < n; i++) g[i] += fabs(A[mc_idx(j, i)]) * t;
            }
            mc_eliminate(A, n, j);
        }
        /* final 2 x 2 block; (n-1, n-2) is the reference's (row > column) address, see the header comment */
        const int a11 = mc_idx(n - 2, n - 2), a22 = mc_idx(n - 1, n - 1), a12 = mc_idx(n - 1, n - 2);
        const double root = sqrt((A[a11] - A[a22]) * (A[a11] - A[a22]) + 4.0 * A[a12] * A[a12]);
        const double lambda_hi = ((A[a11] + A[a22]) + root) * 0.5;
        const double lambda_lo = ((A[a11] + A[a22]) - root) * 0.5;
        delta = mc_max(mc_max(0.0, -lambda_lo + mc_max(tau * (lambda_hi - lambda_lo) / (1.0 - tau), taubar * gamma)), deltaprev);
        if (delta > 0) {
            A[a11] += delta;
            A[a22] += delta;
            deltaprev = delta;
            E[n - 2] = delta;
            E[n - 1] = delta;
        }
        A[a11] = sqrt(A[a11]);
        A[a12] /= A[a11];
        A[a22] = sqrt(A[a22] - A[a12] * A[a12]);
    }
    return deltaprev;
}

/* x = (A + E)^-1 b from the factor and its permutation (y is scratch) */
__host__ __device__ inline void mod_chol_solve(const double *L, const int *P, const double *b, double *x, int n, double *y)
{
    for (int k = 0; k < n; k++) x[k] = b[P[k]];
    for (int k = 0; k < n; k++) {
        for (int i = 0; i < k; i++) x[k] -= x[i] * L[mc_idx(i, k)];
        x[k] /= L[mc_idx(k, k)];
    }
    for (int k = n - 1; k >= 0; k--) {
        for (int i = k + 1; i < n; i++) x[k] -= x[i] * L[mc_idx(k, i)];
        x[k] /= L[mc_idx(k, k)];
    }
    for (int k = 0; k < n; k++) y[P[k]] = x[k];
    for (int k = 0; k < n; k++) x[k] = y[k];
}

/* explicit inverse, one unit right-hand side per column, written back through the permutation */
__host__ __device__ inline void mod_chol_inv(const double *L, const int *P, double *invA, int n, double *x)
{
    for (int l = 0; l < n; l++) {
        x[l] = 1.0;
        for (int k = l + 1; k < n; k++) x[k] = 0.0;
        for (int k = l; k < n; k++) {
            for (int i = l; i < k; i++) x[k] -= x[i] * L[mc_idx(i, k)];
            x[k] /= L[mc_idx(k, k)];
        }
        for (int k = n - 1; k >= l; k--) {
            for (int i = k + 1; i < n; i++) x[k] -= x[i] * L[mc_idx(k, i)];
            x[k] /= L[mc_idx(k, k)];
            int rp = P[k], cp = P[l];
            if (rp > cp) {
                const int t = rp;
                rp = cp;
                cp = t;
            }
            invA[mc_idx(rp, cp)] = x[k];
        }
    }
}

/* H = P L'L P': the matrix the factor stands for, in the original ordering */
__host__ __device__ inline void perm_tri_square(const double *L, double *H, const int *P, int n)
{
    for (int c = 0; c < n; c++)
        for (int r = 0; r <= c; r++) {
            int c_ = P[c], r_ = P[r];
            if (r_ > c_) {
                const int t = r_;
                r_ = c_;
                c_ = t;
            }
            H[mc_idx(r_, c_)] = 0.0;
            for (int i = 0; i <= r; i++) H[mc_idx(r_, c_)] += L[mc_idx(i, c)] * L[mc_idx(i, r)];
        }
}

constexpr int MC_MAXN = 16;

#ifdef __CUDACC__
/* one thread per matrix: factor, shift, permutation, return value, inverse, reconstructed matrix and one solve */
__global__ void k_mod_chol(int n, int count, const double *A_in, const double *b_in, double *fac, double *E, int *P, double *ret,
                           double *inv, double *H, double *x_out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int np = (n * (n + 1)) / 2;
    double A[(MC_MAXN * (MC_MAXN + 1)) / 2], e[MC_MAXN], g[MC_MAXN], xs[MC_MAXN], ys[MC_MAXN], iv[(MC_MAXN * (MC_MAXN + 1)) / 2],
        hh[(MC_MAXN * (MC_MAXN + 1)) / 2];
    int p[MC_MAXN];
    for (int i = 0; i < np; i++) {
        A[i] = A_in[(size_t)t * np + i];
        iv[i] = 0.0;
        hh[i] = 0.0;
    }
    for (int i = 0; i < n; i++) g[i] = 0.0;
    ret[t] = mod_chol(A, n, e, p, g);
    mod_chol_inv(A, p, iv, n, xs);
    perm_tri_square(A, hh, p, n);
    mod_chol_solve(A, p, b_in + (size_t)t * n, xs, n, ys);
    for (int i = 0; i < np; i++) {
        fac[(size_t)t * np + i] = A[i];
        inv[(size_t)t * np + i] = iv[i];
        H[(size_t)t * np + i] = hh[i];
    }
    for (int i = 0; i < n; i++) {
        E[(size_t)t * n + i] = e[i];
        P[(size_t)t * n + i] = p[i];
        x_out[(size_t)t * n + i] = xs[i];
    }
}

#endif /* __CUDACC__ */

} /* namespace ilqg */
