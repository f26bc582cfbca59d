/* mod_chol_host.cpp -- TEST INFRASTRUCTURE: the product's device unit csrc/mod_chol.cuh compiled for the host (its functions are
 * __host__ __device__), exported with C linkage for tests/test_mod_chol.py. */
#include <cmath>
#include <cstdint>
#define __host__
#define __device__
#include "mod_chol.cuh"

extern "C" {
double mch_mod_chol(double *A, int n, double *E, int *P, double *g) { return ilqg::mod_chol(A, n, E, P, g); }
void mch_solve(const double *L, const int *P, const double *b, double *x, int n, double *y) { ilqg::mod_chol_solve(L, P, b, x, n, y); }
void mch_inv(const double *L, const int *P, double *invA, int n, double *x) { ilqg::mod_chol_inv(L, P, invA, n, x); }
void mch_perm_tri_square(const double *L, double *H, const int *P, int n) { ilqg::perm_tri_square(L, H, P, n); }
}
