/* oracle/port_driver.c -- TEST INFRASTRUCTURE: CPU restatement of the iLQG outer loop and its options.
 *
 * Independent re-write of /root/reference iLQG.c:57-78 (defaults), 91-216 (setOptParam), 224-379 (iLQG),
 * 381-386 (makeCandidateNominal), single-threaded path.  Quirks kept on purpose: lambda
 * collapses to exactly 0 at lambdaMin (Q2); the gradient exit also needs lambda < 1e-5 (Q3); `iterations` is
 * the index of the pass that broke out (Q4); a failing derivative pass leaves the previous backPassDone (so
 * the return value can be 1).
 */
#include <math.h>
#include <string.h>
#include <stdio.h>
#include "ilqg_compat.h"

static const double k_default_alpha[8] = {1.0, 0.3727594, 0.1389495, 0.0517947, 0.0193070, 0.0071969, 0.0026827, 0.0010000};

void printParams(double **p, int k) { (void)p; (void)k; }

void standard_parameters(tOptSet *o)
{
    o->alpha = k_default_alpha;
    o->n_alpha = 8;
    o->tolFun = 1e-7;
    o->tolConstraint = 1e-7;
    o->tolGrad = 1e-5;
    o->max_iter = 20;
    o->lambdaInit = 1;
    o->dlambdaInit = 1;
    o->lambdaFactor = 1.6;
    o->lambdaMax = 1e10;
    o->lambdaMin = 1e-6;
    o->regType = 1;
    o->zMin = 0.0;
    o->debug_level = 2;
    o->w_pen_init_l = 1.0;
    o->w_pen_init_f = 1.0;
    o->w_pen_max_l = INF;
    o->w_pen_max_f = INF;
    o->w_pen_fact1 = 4.0;
    o->w_pen_fact2 = 1.0;
}

static char e_scalar[] = "parameter must be scalar";
static char e_alpha_range[] = "all alpha must be in the range [1.0..0.0)";
static char e_alpha_mono[] = "all alpha must be monotonically decreasing";
static char e_pos[] = "parameter must be positive";
static char e_gt1[] = "parameter must be > 1";
static char e_12[] = "parameter must be in range [1..2]";
static char e_01[] = "parameter must be in range [0..1)";
static char e_06[] = "parameter must be in range [0..6]";
static char e_unknown[] = "no such parameter";

enum rule { R_POS_STRICT, R_NONNEG, R_GE1, R_12, R_01, R_06 };

static char *check(enum rule r, double v)
{
    switch (r) {
    case R_POS_STRICT: return (v <= 0.0) ? e_pos : NULL;
    case R_NONNEG: return (v < 0.0) ? e_pos : NULL;
    case R_GE1: return (v < 1.0) ? e_gt1 : NULL;
    case R_12: return (v < 1.0 || v > 2.0) ? e_12 : NULL;
    case R_01: return (v < 0.0 || v >= 1.0) ? e_01 : NULL;
    case R_06: return (v < 0.0 || v > 6.0) ? e_06 : NULL;
    }
    return NULL;
}

char *setOptParam(tOptSet *o, const char *name, const double *value, const int n)
{
    if (strcmp(name, "alpha") == 0) {
        for (int i = 0; i < n; i++) {
            if (value[i] < 0.0 || value[i] > 1.0) return e_alpha_range;
            if (i > 0 && value[i] >= value[i - 1]) return e_alpha_mono;
        }
        o->alpha = value;
        o->n_alpha = n;
        return NULL;
    }
#define DBL_OPT(NAME, RULE) \
    if (strcmp(name, #NAME) == 0) { char *e; if (n != 1) return e_scalar; if ((e = check(RULE, value[0]))) return e; o->NAME = value[0]; return NULL; }
#define INT_OPT(NAME, RULE) \
    if (strcmp(name, #NAME) == 0) { char *e; if (n != 1) return e_scalar; if ((e = check(RULE, value[0]))) return e; o->NAME = (int)value[0]; return NULL; }
    DBL_OPT(tolFun, R_POS_STRICT)
    DBL_OPT(tolConstraint, R_POS_STRICT)
    DBL_OPT(tolGrad, R_POS_STRICT)
    INT_OPT(max_iter, R_NONNEG)
    DBL_OPT(lambdaInit, R_NONNEG)
    DBL_OPT(dlambdaInit, R_NONNEG)
    DBL_OPT(lambdaFactor, R_GE1)
    DBL_OPT(lambdaMax, R_NONNEG)
    DBL_OPT(lambdaMin, R_NONNEG)
    INT_OPT(regType, R_12)
    DBL_OPT(zMin, R_01)
    INT_OPT(debug_level, R_06)
    DBL_OPT(w_pen_init_l, R_NONNEG)
    DBL_OPT(w_pen_init_f, R_NONNEG)
    DBL_OPT(w_pen_max_l, R_NONNEG)
    DBL_OPT(w_pen_max_f, R_NONNEG)
    DBL_OPT(w_pen_fact1, R_GE1)
    DBL_OPT(w_pen_fact2, R_GE1)
    return e_unknown;
}

void makeCandidateNominal(tOptSet *o, int idx)
{
    traj_t *was_nominal = o->nominal;
    o->nominal = o->candidates[idx];
    o->candidates[idx] = was_nominal;
}

static void raise_lambda(tOptSet *o, double *dlambda)
{
    *dlambda = max(*dlambda * o->lambdaFactor, o->lambdaFactor);
    o->lambda = max(o->lambda * *dlambda, o->lambdaMin);
}

static void lower_lambda(tOptSet *o, double *dlambda)
{
    *dlambda = min(*dlambda / o->lambdaFactor, 1.0 / o->lambdaFactor);
    o->lambda = o->lambda * *dlambda * (o->lambda > o->lambdaMin);
}

int iLQG(tOptSet *o)
{
    int iter, back_ok = 0, step_ok = 0, need_derivs = 1;
    double dlambda = o->dlambdaInit;

    o->lambda = o->lambdaInit;
    o->w_pen_l = o->w_pen_init_l;
    o->w_pen_f = o->w_pen_init_f;
    update_multipliers(o, 1);

    for (iter = 0; iter < o->max_iter; iter++) {
        if (need_derivs) {
            if (!calc_derivs(o))
                break;
            need_derivs = 0;
        }
        back_ok = 0;
        while (!back_ok) {
            if (back_pass(o)) {
                raise_lambda(o, &dlambda);
                if (o->lambda > o->lambdaMax)
                    break;
            } else {
                back_ok = 1;
            }
        }
        if (o->g_norm < o->tolGrad && o->lambda < 1e-5) {
            lower_lambda(o, &dlambda);
            break;
        }
        if (!back_ok)
            break;
        step_ok = line_search(o, iter);
        if (step_ok) {
            lower_lambda(o, &dlambda);
            makeCandidateNominal(o, 0);
            o->cost = o->new_cost;
            need_derivs = 1;
            if (o->dcost < o->tolFun)
                break;
            update_multipliers(o, 0);
            forward_pass(o->nominal, o, 0.0, &o->cost, 1);
        } else {
            raise_lambda(o, &dlambda);
            if (o->w_pen_fact2 > 1.0) {
                o->w_pen_l = min(o->w_pen_max_l, o->w_pen_l * o->w_pen_fact2);
                o->w_pen_f = min(o->w_pen_max_f, o->w_pen_f * o->w_pen_fact2);
                forward_pass(o->nominal, o, 0.0, &o->cost, 1);
            }
            if (o->lambda > o->lambdaMax)
                break;
        }
    }
    o->iterations = iter;
    (void)step_ok;
    if (!back_ok)
        return 0;
    if (iter >= o->max_iter)
        return 0;
    return 1;
}
