/* oracle/port_qp.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's small dense algebra and box-QP.
 *
 * Independent re-write (not a copy) of the arithmetic in /root/reference:
 *   addMulVec / addSquareTri / addMul2Tri   matMult.c:3-12, 14-46, 48-72
 *   cholesky_tri / cholesky_tri_inv         cholesky.c:6-27, 51-74
 *   boxQP                                   boxQP.c:39-238
 * The ORDER of every floating-point operation follows the reference (single accumulator, ascending index,
 * explicit symmetrisation, reciprocal-then-multiply in the factorisation) because parity is bit-exact.
 * Pinned against the compiled reference by tests/test_oracle_port.py (unit KATs + full solves).
 */
#include <math.h>
#include <string.h>
#include "ilqg_compat.h"

#define PK(r, c) SYMTRI_MAT_IDX(r, c)

/* base[c] += sum_r a[r] * b[r, c]   (b column-major n_r x n_c), accumulating directly into base */
void addMulVec(double base[], const double a[], const double b[], const int n_r, const int n_c)
{
    for (int c = 0; c < n_c; c++) {
        const double *col = b + (size_t)c * n_r;
        for (int r = 0; r < n_r; r++)
            base[c] += a[r] * col[r];
    }
}

/* base(packed upper) += a' * B * a,  B packed symmetric n_r x n_r, a dense n_r x n_c, ba = scratch n_r x n_c */
void addSquareTri(double base[], const double b[], const double a[], const int n_r, const int n_c, double ba[])
{
    for (int c = 0; c < n_c; c++)
        for (int r = 0; r < n_r; r++) {
            double acc = 0.0;
            for (int s = 0; s < n_r; s++)
                acc += b[PK(r, s)] * a[s + c * n_r];
            ba[r + c * n_r] = acc;
        }
    for (int c = 0; c < n_c; c++)
        for (int r = 0; r <= c; r++) {
            double acc = 0.0;
            for (int s = 0; s < n_r; s++)
                acc += a[s + r * n_r] * ba[s + c * n_r];
            if (r != c) { /* add the transposed element, then halve: explicit symmetrisation */
                for (int s = 0; s < n_r; s++)
                    acc += a[s + c * n_r] * ba[s + r * n_r];
                acc *= 0.5;
            }
            base[UTRI_MAT_IDX(r, c)] += acc;
        }
}

/* base(dense n_ca x n_cc) += a' * B * c,  B packed symmetric n_ra x n_rc */
void addMul2Tri(double base[], const double b[], const double a[], const int n_ra, const int n_ca,
                const double c[], const int n_rc, const int n_cc, double bc[])
{
    for (int j = 0; j < n_cc; j++)
        for (int r = 0; r < n_ra; r++) {
            double acc = 0.0;
            for (int s = 0; s < n_rc; s++)
                acc += b[PK(r, s)] * c[s + j * n_rc];
            bc[r + j * n_ra] = acc;
        }
    for (int i = 0; i < n_ca; i++)
        for (int j = 0; j < n_cc; j++) {
            double acc = 0.0;
            for (int s = 0; s < n_ra; s++)
                acc += a[s + i * n_ra] * bc[s + j * n_ra];
            base[i + j * n_ca] += acc;
        }
}

/* A = U'U with U packed upper; 0 when a pivot is <= 0 */
int cholesky_tri(const double *A, int n, double *U)
{
    for (int col = 0; col < n; col++)
        for (int row = 0; row <= col; row++) {
            double dot = 0;
            for (int k = 0; k < row; k++)
                dot += U[UTRI_MAT_IDX(k, col)] * U[UTRI_MAT_IDX(k, row)];
            double rem = A[UTRI_MAT_IDX(row, col)] - dot;
            if (row == col) {
                if (rem <= 0.0)
                    return 0;
                U[UTRI_MAT_IDX(row, col)] = sqrt(rem);
            } else {
                U[UTRI_MAT_IDX(row, col)] = 1.0 / U[UTRI_MAT_IDX(row, row)] * rem;
            }
        }
    return 1;
}

/* explicit inverse of A from its factor, one unit right-hand side at a time; w = scratch[n] */
void cholesky_tri_inv(const double *U, double *invA, const int n, double *w)
{
    for (int col = 0; col < n; col++) {
        w[col] = 1.0;
        for (int k = col + 1; k < n; k++)
            w[k] = 0.0;
        for (int k = col; k < n; k++) { /* forward: U' y = e_col */
            for (int i = col; i < k; i++)
                w[k] -= w[i] * U[UTRI_MAT_IDX(i, k)];
            w[k] /= U[UTRI_MAT_IDX(k, k)];
        }
        for (int k = n - 1; k >= col; k--) { /* backward: U x = y */
            for (int i = k + 1; i < n; i++)
                w[k] -= w[i] * U[UTRI_MAT_IDX(k, i)];
            w[k] /= U[UTRI_MAT_IDX(k, k)];
            invA[UTRI_MAT_IDX(col, k)] = w[k];
        }
    }
}

static double qp_objective(const double *H, const double *g, const double *x, int n)
{
    double val = 0.0;
    for (int i = 0; i < n; i++) {
        double hx = 0.0;
        for (int j = 0; j < n; j++)
            hx += H[PK(i, j)] * x[j];
        val += x[i] * (g[i] + 0.5 * hx);
    }
    return val;
}

/* projected-Newton box QP.  Return codes as the reference: -2 no descent, -1 not PD, 1 iteration limit,
 * 2 step underflow, 4 small improvement, 5 small gradient, 6 all clamped. */
int boxQP(double *H, const double *g, const double *lower, const double *upper, double *x, double *Hfree,
          double *U, double *grad, double *grad_clamped, double *search, int *is_clamped, int *n_free_,
          double *invHfree, const int n)
{
    const int max_iter = 100;
    const double min_grad = 1e-8, min_rel_improve = 1e-8, step_dec = 0.6, min_step = 1e-22, armijo = 0.1;
    double value, oldvalue = 0.0;

    memset(Hfree, 0, sizeof(double) * (n * (n + 1)) / 2);
    for (int i = 0; i < n; i++) {
        if (x[i] > upper[i]) x[i] = upper[i];
        if (x[i] < lower[i]) x[i] = lower[i];
        is_clamped[i] = 0;
    }
    value = qp_objective(H, g, x, n);

    for (int iter = 0; iter < max_iter; iter++) {
        if (iter > 0 && (oldvalue - value) < min_rel_improve * fabs(oldvalue))
            return 4;
        oldvalue = value;

        int n_free = 0, changed = 0, all_clamped = 1;
        double gsq = 0.0;
        for (int i = 0; i < n; i++) {
            double hx = 0.0;
            for (int j = 0; j < n; j++)
                hx += H[PK(i, j)] * x[j];
            grad[i] = g[i] + hx;
            const int was = is_clamped[i];
            if (x[i] <= lower[i] && grad[i] > 0)
                is_clamped[i] = 1;
            else if (x[i] >= upper[i] && grad[i] < 0)
                is_clamped[i] = 2;
            else {
                is_clamped[i] = 0;
                all_clamped = 0;
                gsq += grad[i] * grad[i];
                n_free++;
            }
            if ((!was) != (!is_clamped[i]))
                changed = 1;
        }
        n_free_[0] = n_free;
        if (all_clamped)
            return 6;

        if (iter == 0 || changed) { /* re-factorise the free block */
            int jf = 0;
            for (int j = 0; j < n; j++) {
                if (is_clamped[j]) continue;
                int fi = 0;
                for (int i = 0; i <= j; i++) {
                    if (is_clamped[i]) continue;
                    Hfree[UTRI_MAT_IDX(fi, jf)] = H[UTRI_MAT_IDX(i, j)];
                    fi++;
                }
                jf++;
            }
            if (!cholesky_tri(Hfree, n_free, U))
                return -1;
            cholesky_tri_inv(U, invHfree, n_free, search);
        }
        if (gsq < min_grad * min_grad)
            return 5;

        /* Newton direction on the free set: search = -Hfree^-1 (g + H x_clamped) - x */
        {
            int fi = 0;
            for (int i = 0; i < n; i++) {
                if (is_clamped[i]) continue;
                double hc = 0.0;
                for (int j = 0; j < n; j++)
                    if (is_clamped[j])
                        hc += H[PK(i, j)] * x[j];
                grad_clamped[fi++] = g[i] + hc;
            }
            fi = 0;
            for (int i = 0; i < n; i++) {
                if (is_clamped[i]) {
                    search[i] = 0.0;
                    continue;
                }
                search[i] = -x[i];
                for (int jf = 0; jf < n_free; jf++)
                    search[i] -= invHfree[PK(fi, jf)] * grad_clamped[jf];
                fi++;
            }
        }
        double sdotg = 0.0;
        for (int i = 0; i < n; i++)
            sdotg += search[i] * grad[i];
        if (sdotg >= 0.0)
            return -2;

        /* Armijo backtracking on the projected step; candidate lives in grad[] (free after sdotg) */
        double step = 1.0, vc;
        double *xc = grad;
        for (;;) {
            for (int i = 0; i < n; i++) {
                xc[i] = x[i] + step * search[i];
                if (xc[i] > upper[i]) xc[i] = upper[i];
                if (xc[i] < lower[i]) xc[i] = lower[i];
            }
            vc = qp_objective(H, g, xc, n);
            if (((vc - oldvalue) / (step * sdotg)) >= armijo)
                break;
            step = step * step_dec;
            if (step < min_step)
                return 2;
        }
        for (int i = 0; i < n; i++)
            x[i] = xc[i];
        value = vc;
    }
    return 1;
}
