/* oracle/dm_check.c -- TEST INFRASTRUCTURE: exports the deterministic math layer (csrc/dm_math.h) so tests can compare
 * it with libm and pin its bit patterns (tests/test_dm_math.py). */
#include "dm_math.h"
double dmc_sin(double x) { return dm_sin(x); }
double dmc_cos(double x) { return dm_cos(x); }
double dmc_asin(double x) { return dm_asin(x); }
double dmc_acos(double x) { return dm_acos(x); }
void dmc_vec(int which, const double *x, double *y, long n)
{
    for (long i = 0; i < n; i++)
        y[i] = which == 0 ? dm_sin(x[i]) : which == 1 ? dm_cos(x[i]) : which == 2 ? dm_asin(x[i]) : dm_acos(x[i]);
}
