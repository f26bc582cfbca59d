/* ilqg_dropin.c -- the reference's single-problem entry points, executed on the B200.
 *
 * A caller written against the reference's API (the mex gateway iLQG_mex.c:59-137 is the only shipped one) calls
 *     standard_parameters, setOptParam, init_opt, forward_pass(candidates[0], o, 0.0, &cost, 0),
 *     makeCandidateNominal, iLQG
 * on a caller-allocated tOptSet.  This file exports those symbols and the solver's public phases (calc_derivs, back_pass,
 * line_search, update_multipliers, clampU: iLQG.h:83-86, back_pass.h:7, line_search.h:6) (plus the paramdesc globals the binding
 * reads, iLQG.h:97-99) on top of the batched C ABI with a batch of one: state is marshalled between the caller's
 * array-of-structs trajectories and the device layout, every rollout / derivative / backward pass / line search
 * runs in the CUDA kernels.  Compiled per problem against the generated iLQG_problem.h, like the reference.
 *
 * Postconditions of iLQG(o) as the reference leaves them: nominal->t[k].x/u, nominal->f.x,
 * o->cost/new_cost/dcost/expected/lambda/g_norm/iterations/dV/w_pen_l/w_pen_f, multipliers, optional log_* arrays;
 * additionally l and L of both trajectory buffers.  `nominal`/`candidates[0]` end up swapped exactly when the
 * reference would have swapped them an odd number of times.  NOT maintained on the host: the derivative members of
 * trajEl_t (fx, fu, cx ... live on the device), per-step c and the aux members.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include <math.h>

#include "ilqg_compat.h"
#include "ilqg_b200.h"

int n_params = 0;
int n_vars = 0;
static tParamDesc g_desc[64];
tParamDesc *paramdesc[64];

__attribute__((constructor)) static void dropin_load(void)
{
    int i;
    n_params = ilqgb_n_params();
    for (i = 0; i < n_params && i < 64; i++) {
        g_desc[i].name = (char *)ilqgb_param_name(i);
        g_desc[i].size = ilqgb_param_size(i);
        g_desc[i].is_var = 0;
        paramdesc[i] = &g_desc[i];
    }
}

/* ---- options ---------------------------------------------------------------------------------------------------------- */
static const double k_alpha[8] = {1.0, 0.3727594, 0.1389495, 0.0517947, 0.0193070, 0.0071969, 0.0026827, 0.0010000};

void standard_parameters(tOptSet *o)
{
    o->alpha = k_alpha;
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

typedef struct {
    const char *name;
    size_t off;
    int is_int;
} ofield;
#define OF_D(f) {#f, offsetof(tOptSet, f), 0}
#define OF_I(f) {#f, offsetof(tOptSet, f), 1}
static const ofield k_fields[] = {
    OF_D(tolFun), OF_D(tolConstraint), OF_D(tolGrad), OF_I(max_iter), OF_D(lambdaInit), OF_D(dlambdaInit), OF_D(lambdaFactor),
    OF_D(lambdaMax), OF_D(lambdaMin), OF_I(regType), OF_D(zMin), OF_I(debug_level), OF_D(w_pen_init_l), OF_D(w_pen_init_f),
    OF_D(w_pen_max_l), OF_D(w_pen_max_f), OF_D(w_pen_fact1), OF_D(w_pen_fact2),
};
#define N_FIELDS (sizeof k_fields / sizeof k_fields[0])

char *setOptParam(tOptSet *o, const char *name, const double *value, const int n)
{
    size_t i;
    const char *msg = ilqgb_validate_opt(name, value, n);
    if (msg && strcmp(msg, "at most 16 alpha values are supported") != 0) return (char *)msg;
    if (strcmp(name, "alpha") == 0) { /* the caller's pointer is kept, as in the reference (iLQG.c:101) */
        o->alpha = value;
        o->n_alpha = n;
        return NULL;
    }
    for (i = 0; i < N_FIELDS; i++)
        if (strcmp(name, k_fields[i].name) == 0) {
            if (k_fields[i].is_int)
                *(int *)((char *)o + k_fields[i].off) = (int)value[0];
            else
                *(double *)((char *)o + k_fields[i].off) = value[0];
            return NULL;
        }
    return (char *)"no such parameter";
}

void makeCandidateNominal(tOptSet *o, int idx)
{
    traj_t *t = o->nominal;
    o->nominal = o->candidates[idx];
    o->candidates[idx] = t;
}

void printParams(double **p, int k)
{
    int i, j;
    for (i = 0; i < n_params; i++) {
        const int n = paramdesc[i]->size == -1 ? 1 : paramdesc[i]->size;
        PRNT("%s= ", paramdesc[i]->name);
        for (j = 0; j < n; j++) PRNT("%g ", paramdesc[i]->size == -1 ? p[i][k] : p[i][j]);
        PRNT("\n");
    }
}
/* user outputs of the problem file (iLQG_func.tem:511-521), evaluated on the device like everything else; calcG has no tOptSet:
   like clampU it takes the parameters from p (a [k]-indexed parameter is read at index k only, so any horizon >= k serves) */
int get_g_size() { return ilqgb_eval_size(17); }
static ilqgb_handle *cached_handle_covering(int k);
int calcG(double g[], trajEl_t *t, int k, double *p[])
{
    ilqgb_handle *h;
    int i;
    if (ilqgb_eval_size(17) <= 0) return 1;
    for (i = 0; i < n_params; i++)
        if (paramdesc[i]->size == -1) return 0;   /* the length of a [k]-indexed vector is not known here */
    h = cached_handle_covering(k);
    if (!h) return 0;
    for (i = 0; i < n_params; i++)
        if (ilqgb_set_param(h, i, p[i], paramdesc[i]->size)) return 0;
    return ilqgb_eval(h, 17, k, t->x, t->u, g) ? 0 : 1;
}

/* ---- multipliers: struct members <-> device order [equalities..., inequalities...] ----------------------------------------- */
#define N_EL ((int)(sizeof(multipliersEl_t) / sizeof(double)))
#define N_FIN ((int)(sizeof(multipliersFin_t) / sizeof(double)))

static void mult_layout(int n_tot, int n_eq, int i, int *mu_at, int *last_at)
{
    /* struct = mu_eq[n_eq] last_eq[n_eq] mu_in[n_in] last_in[n_in] (iLQG_problem.tem:70-89) */
    const int n_in = n_tot - n_eq;
    if (i < n_eq) { *mu_at = i; *last_at = n_eq + i; }
    else { *mu_at = 2 * n_eq + (i - n_eq); *last_at = 2 * n_eq + n_in + (i - n_eq); }
}

int init_opt(tOptSet *o)
{
    int i, k;
    o->nominal = &o->trajectories[0];
    for (i = 1; i < NUMBER_OF_THREADS + 1; i++) o->candidates[i - 1] = &o->trajectories[i];
    /* init_multipliers: equalities mu = 0, inequalities mu = 1, last_h = 0 (iLQG_func.tem:364-400) */
    {
        const int n_r = N_EL / 2, n_f = N_FIN / 2;
        int n_le = 0, n_fe = 0;
        ilqgb_mult_counts(&n_le, &n_fe);
        for (k = 0; k < o->n_hor && N_EL > 0; k++) {
            double *m = (double *)&o->multipliers.t[k];
            for (i = 0; i < n_r; i++) {
                int a, b;
                mult_layout(n_r, n_le, i, &a, &b);
                m[a] = i < n_le ? 0.0 : 1.0;
                m[b] = 0.0;
            }
        }
        if (N_FIN > 0) {
            double *m = (double *)&o->multipliers.f;
            for (i = 0; i < n_f; i++) {
                int a, b;
                mult_layout(n_f, n_fe, i, &a, &b);
                m[a] = i < n_fe ? 0.0 : 1.0;
                m[b] = 0.0;
            }
        }
    }
    return 1;
}

/* ---- handle cache and marshalling ---------------------------------------------------------------------------------------- */
static ilqgb_handle *g_h = NULL;
static int g_T = -1, g_trace_len = 1;

static ilqgb_handle *handle_for(int n_hor)
{
    if (g_h && g_T == n_hor) return g_h;
    if (g_h) ilqgb_destroy(g_h);
    g_h = ilqgb_create(0, 1, n_hor, ILQGB_TRACE | ILQGB_CHUNKS(1), NULL);
    g_T = g_h ? n_hor : -1;
    g_trace_len = 1;
    if (!g_h) PRNT("ilqg_b200: %s\n", ilqgb_last_error(NULL));
    return g_h;
}

/* any handle whose horizon reaches step k will do for a single evaluation: keep the cached one when it does */
static ilqgb_handle *cached_handle_covering(int k) { return (g_h && g_T >= k) ? g_h : handle_for(k + 1); }

static int push_params(ilqgb_handle *h, tOptSet *o)
{
    int i;
    for (i = 0; i < n_params; i++)
        if (ilqgb_set_param(h, i, o->p[i], paramdesc[i]->size == -1 ? o->n_hor + 1 : paramdesc[i]->size)) return -1;
    return 0;
}

static int push_options(ilqgb_handle *h, tOptSet *o)
{
    size_t i;
    if (ilqgb_set_opt(h, "alpha", o->alpha, o->n_alpha)) return -1;
    for (i = 0; i < N_FIELDS; i++) {
        double v = k_fields[i].is_int ? (double)*(int *)((char *)o + k_fields[i].off) : *(double *)((char *)o + k_fields[i].off);
        if (!strcmp(k_fields[i].name, "debug_level")) continue;
        if (ilqgb_set_opt(h, k_fields[i].name, &v, 1)) return -1;
    }
    return 0;
}

/* nominal trajectory, control law and multipliers of `o` -> device buffer 0 */
static int push_state(ilqgb_handle *h, tOptSet *o, int with_law)
{
    const int T = o->n_hor, n_r = N_EL / 2, n_f = N_FIN / 2;
    int k, i, zero = 0, n_le = 0, n_fe = 0;
    double *x = (double *)malloc(sizeof(double) * ((size_t)(T + 1) * N_X + (size_t)T * (N_U + N_U + N_U * N_X) + 2 * (size_t)T * (n_r + 1) + 2 * (n_f + 1)));
    double *u = x + (size_t)(T + 1) * N_X, *l = u + (size_t)T * N_U, *L = l + (size_t)T * N_U;
    double *mr = L + (size_t)T * N_U * N_X, *lr = mr + (size_t)T * (n_r + 1), *mf = lr + (size_t)T * (n_r + 1), *lf = mf + n_f + 1;
    int rc = 0;
    ilqgb_mult_counts(&n_le, &n_fe);
    for (k = 0; k < T; k++) {
        const trajEl_t *t = &o->nominal->t[k];
        memcpy(x + (size_t)k * N_X, t->x, sizeof t->x);
        memcpy(u + (size_t)k * N_U, t->u, sizeof t->u);
        memcpy(l + (size_t)k * N_U, t->l, sizeof t->l);
        memcpy(L + (size_t)k * N_U * N_X, t->L, sizeof t->L);
        for (i = 0; i < n_r; i++) {
            int a, b;
            mult_layout(n_r, n_le, i, &a, &b);
            mr[(size_t)k * n_r + i] = ((const double *)&o->multipliers.t[k])[a];
            lr[(size_t)k * n_r + i] = ((const double *)&o->multipliers.t[k])[b];
        }
    }
    memcpy(x + (size_t)T * N_X, o->nominal->f.x, sizeof o->nominal->f.x);
    for (i = 0; i < n_f; i++) {
        int a, b;
        mult_layout(n_f, n_fe, i, &a, &b);
        mf[i] = ((const double *)&o->multipliers.f)[a];
        lf[i] = ((const double *)&o->multipliers.f)[b];
    }
    rc |= ilqgb_put_int(h, "cur", &zero) < 0;
    rc |= ilqgb_put_int(h, "status", &zero) < 0;
    rc |= ilqgb_put(h, "x0", o->x0) < 0;
    rc |= ilqgb_put(h, "x", x) < 0;
    rc |= ilqgb_put(h, "u", u) < 0;
    if (with_law) {
        rc |= ilqgb_put(h, "l", l) < 0;
        rc |= ilqgb_put(h, "L", L) < 0;
    }
    if (n_r) { rc |= ilqgb_put(h, "mu_r", mr) < 0; rc |= ilqgb_put(h, "last_r", lr) < 0; }
    if (n_f) { rc |= ilqgb_put(h, "mu_f", mf) < 0; rc |= ilqgb_put(h, "last_f", lf) < 0; }
    rc |= ilqgb_put(h, "cost", &o->cost) < 0;
    rc |= ilqgb_put(h, "w_pen_l", &o->w_pen_l) < 0;
    rc |= ilqgb_put(h, "w_pen_f", &o->w_pen_f) < 0;
    free(x);
    return rc ? -1 : 0;
}

static int pull_traj(ilqgb_handle *h, traj_t *dst, const char *fx, const char *fu, int T)
{
    double *x = (double *)malloc(sizeof(double) * ((size_t)(T + 1) * N_X + (size_t)T * N_U));
    double *u = x + (size_t)(T + 1) * N_X;
    int k, rc = 0;
    rc |= ilqgb_get(h, fx, x) < 0;
    rc |= ilqgb_get(h, fu, u) < 0;
    for (k = 0; k < T && !rc; k++) {
        memcpy(dst->t[k].x, x + (size_t)k * N_X, sizeof dst->t[k].x);
        memcpy(dst->t[k].u, u + (size_t)k * N_U, sizeof dst->t[k].u);
    }
    if (!rc) memcpy(dst->f.x, x + (size_t)T * N_X, sizeof dst->f.x);
    free(x);
    return rc ? -1 : 0;
}


/* ---- the solver's phases on their own (iLQG.h:83,86; back_pass.h:7; line_search.h:6; iLQG_func.tem:68) --------------------------
 * Each call is self-contained: the caller's tOptSet is marshalled to the device, the phase runs in its kernel, and exactly the
 * members the reference's function writes are copied back.  The derivative members of trajEl_t are filled by calc_derivs for
 * inspection (fx fu cx cxx cu cuu cxu, limits and their gradients; the FULL_DDP tensors fxx/fuu/fxu stay on the device), but
 * back_pass does not read them from the host struct: it re-evaluates them on the device from the nominal x, u -- the same bits. */
static int push_all(ilqgb_handle *h, tOptSet *o, int with_law)
{
    int one = 1, zero = 0, rc = 0;
    if (push_options(h, o) || push_params(h, o) || push_state(h, o, with_law)) return -1;
    rc |= ilqgb_put(h, "lambda", &o->lambda) < 0;
    rc |= ilqgb_put(h, "dV0", &o->dV[0]) < 0;
    rc |= ilqgb_put(h, "dV1", &o->dV[1]) < 0;
    rc |= ilqgb_put(h, "g_norm", &o->g_norm) < 0;
    rc |= ilqgb_put(h, "new_cost", &o->new_cost) < 0;
    rc |= ilqgb_put(h, "dcost", &o->dcost) < 0;
    rc |= ilqgb_put(h, "expected", &o->expected) < 0;
    rc |= ilqgb_put_int(h, "new_deriv", &one) < 0;
    rc |= ilqgb_put_int(h, "deriv_fail", &zero) < 0;
    rc |= ilqgb_put_int(h, "bp_done", &zero) < 0;
    rc |= ilqgb_set_tuning(h, "pass_index", 0) < 0;
    if (o->max_iter > g_trace_len) g_trace_len = o->max_iter;
    return rc ? -1 : 0;
}

int calc_derivs(tOptSet *o)
{
    ilqgb_handle *h = handle_for(o->n_hor);
    const int T = o->n_hor, DS = ilqgb_dense_size();
    int fail = 0, k;
    double *buf, *fd;
    if (!h || push_all(h, o, 0) || ilqgb_phase_derivs(h)) return 0;
    if (ilqgb_get_int(h, "deriv_fail", &fail) < 0 || fail) return 0;
    buf = (double *)malloc(sizeof(double) * ((size_t)T * DS + N_X + sizeofQxx));
    if (!buf) return 0;
    fd = buf + (size_t)T * DS;
    if (ilqgb_get(h, "dense", buf) < 0 || ilqgb_get(h, "fd", fd) < 0) { free(buf); return 0; }
    for (k = 0; k < T; k++) {
        trajEl_t *t = &o->nominal->t[k];
        const double *d = buf + (size_t)k * DS;
#define TAKE(member) do { memcpy(t->member, d, sizeof t->member); d += sizeof t->member / sizeof(double); } while (0)
        TAKE(fx); TAKE(fu); TAKE(cx); TAKE(cxx); TAKE(cu); TAKE(cuu); TAKE(cxu);
        TAKE(lower); TAKE(upper); TAKE(lower_sign); TAKE(upper_sign); TAKE(lower_hx); TAKE(upper_hx);
#undef TAKE
    }
    memcpy(o->nominal->f.cx, fd, sizeof o->nominal->f.cx);
    memcpy(o->nominal->f.cxx, fd + N_X, sizeof o->nominal->f.cxx);
    free(buf);
    return 1;
}

static void pull_law(ilqgb_handle *h, tOptSet *o)
{
    const int T = o->n_hor;
    int k;
    double *l = (double *)malloc(sizeof(double) * (size_t)T * (N_U + N_U * N_X)), *L;
    if (!l) return;
    L = l + (size_t)T * N_U;
    ilqgb_get(h, "l", l);
    ilqgb_get(h, "L", L);
    for (k = 0; k < T; k++) {
        memcpy(o->nominal->t[k].l, l + (size_t)k * N_U, sizeof o->nominal->t[k].l);
        memcpy(o->nominal->t[k].L, L + (size_t)k * N_U * N_X, sizeof o->nominal->t[k].L);
    }
    free(l);
}

/* returns 0 = ok, 1 = the box QP of some step failed (back_pass.c:163-171); o->lambda is the caller's business */
int back_pass(tOptSet *o)
{
    ilqgb_handle *h = handle_for(o->n_hor);
    int done = 0;
    if (!h || push_all(h, o, 1) || ilqgb_phase_derivs(h) || ilqgb_phase_backpass_once(h)) return 1;
    if (ilqgb_get_int(h, "bp_done", &done) < 0) return 1;
    pull_law(h, o);                       /* a failed pass leaves the steps it reached rewritten, like the reference */
    ilqgb_get(h, "dV0", &o->dV[0]);
    ilqgb_get(h, "dV1", &o->dV[1]);
    if (done) ilqgb_get(h, "g_norm", &o->g_norm);
    return done ? 0 : 1;
}

int line_search(tOptSet *o, int iter)
{
    ilqgb_handle *h = handle_for(o->n_hor);
    int cur = 0, rc;
    if (!h || push_all(h, o, 1)) return 0;
    /* every alpha as its own stored rollout, so that candidates[0] holds the last rollout tried, accepted or not */
    if (ilqgb_set_tuning(h, "ls_tail_from", o->n_alpha)) return 0;
    rc = ilqgb_phase_linesearch(h);
    ilqgb_set_tuning(h, "ls_tail_from", -1);
    if (rc || ilqgb_get_int(h, "cur", &cur) < 0) return 0;
    ilqgb_get(h, "new_cost", &o->new_cost);
    ilqgb_get(h, "dcost", &o->dcost);
    ilqgb_get(h, "expected", &o->expected);
    {   /* the phase ran as pass 0: row 0 of the traces is this call's log entry */
        int *a = (int *)malloc(sizeof(int) * g_trace_len);
        double *z = (double *)malloc(sizeof(double) * 2 * g_trace_len), *c = z ? z + g_trace_len : NULL;
        if (a && z && ilqgb_get_int(h, "tr_alpha", a) > 0 && ilqgb_get(h, "tr_z", z) > 0 && ilqgb_get(h, "tr_newcost", c) > 0) {
            if (o->log_linesearch) o->log_linesearch[iter] = a[0];
            if (o->log_z) o->log_z[iter] = z[0];
            if (o->log_cost) o->log_cost[iter] = c[0];
        }
        free(a);
        free(z);
    }
    /* the device has already flipped its buffers for an accepted step; here that stays with the caller (iLQG.c:322) */
    if (pull_traj(h, o->candidates[0], cur ? "x" : "x_cand", cur ? "u" : "u_cand", o->n_hor)) return 0;
    return cur ? 1 : 0;
}

int update_multipliers(tOptSet *o, int init)
{
    ilqgb_handle *h = handle_for(o->n_hor);
    const int T = o->n_hor, n_r = N_EL / 2, n_f = N_FIN / 2;
    int n_le = 0, n_fe = 0, i, k;
    double *mr, *lr, *mf, *lf;
    if (n_r + n_f == 0) return 1;
    if (!h || push_all(h, o, 0) || ilqgb_phase_multipliers(h, init)) return 0;
    ilqgb_mult_counts(&n_le, &n_fe);
    mr = (double *)malloc(sizeof(double) * (2 * (size_t)T * (n_r + 1) + 2 * (n_f + 1)));
    if (!mr) return 0;
    lr = mr + (size_t)T * (n_r + 1); mf = lr + (size_t)T * (n_r + 1); lf = mf + n_f + 1;
    if (n_r) { ilqgb_get(h, "mu_r", mr); ilqgb_get(h, "last_r", lr); }
    if (n_f) { ilqgb_get(h, "mu_f", mf); ilqgb_get(h, "last_f", lf); }
    for (k = 0; k < T; k++)
        for (i = 0; i < n_r; i++) {
            int a, b;
            mult_layout(n_r, n_le, i, &a, &b);
            ((double *)&o->multipliers.t[k])[a] = mr[(size_t)k * n_r + i];
            ((double *)&o->multipliers.t[k])[b] = lr[(size_t)k * n_r + i];
        }
    for (i = 0; i < n_f; i++) {
        int a, b;
        mult_layout(n_f, n_fe, i, &a, &b);
        ((double *)&o->multipliers.f)[a] = mf[i];
        ((double *)&o->multipliers.f)[b] = lf[i];
    }
    free(mr);
    ilqgb_get(h, "w_pen_l", &o->w_pen_l);
    ilqgb_get(h, "w_pen_f", &o->w_pen_f);
    return 1;
}

/* clampU has no tOptSet: the horizon comes from N, the parameters from p */
void clampU(double *u, trajEl_t *t, int k, double **p, int N)
{
    ilqgb_handle *h = handle_for(N);
    int i;
    if (!h) return;
    for (i = 0; i < n_params; i++)
        if (ilqgb_set_param(h, i, p[i], paramdesc[i]->size == -1 ? N + 1 : paramdesc[i]->size)) return;
    ilqgb_clamp_u(h, k, t->x, u);
}

/* ---- forward_pass (iLQG.h:82; iLQG_func.tem:121-185) --------------------------------------------------------------------------- */
int forward_pass(traj_t *c, tOptSet *o, double alpha, double *csum, int cost_only)
{
    ilqgb_handle *h = handle_for(o->n_hor);
    int ok = 0;
    csum[0] = 0.0;
    if (!h) return 0;
    if (push_params(h, o) || push_state(h, o, alpha != 0.0)) return 0;
    if (ilqgb_rollout(h, alpha, cost_only)) return 0;
    if (ilqgb_get(h, "new_cost", csum) < 0 || ilqgb_get_int(h, "result", &ok) < 0) return 0;
    if (!cost_only && pull_traj(h, c, "x_cand", "u_cand", o->n_hor)) return 0;
    return ok;
}

/* ---- iLQG (iLQG.h:80; iLQG.c:224-379) --------------------------------------------------------------------------------------------- */
int iLQG(tOptSet *o)
{
    ilqgb_handle *h = handle_for(o->n_hor);
    const int T = o->n_hor, n_r = N_EL / 2, n_f = N_FIN / 2;
    int result = 0, cur = 0, n_ls = 0, i, k, n_le = 0, n_fe = 0;
    if (!h) return 0;
    ilqgb_mult_counts(&n_le, &n_fe);
    if (push_options(h, o) || push_params(h, o) || push_state(h, o, 0)) return 0;
    if (o->max_iter > g_trace_len) g_trace_len = o->max_iter;
    if (ilqgb_begin(h)) return 0;
    if (o->max_iter > 0 && ilqgb_iterate(h, o->max_iter) < 0) return 0;
    if (ilqgb_finish(h) || ilqgb_sync(h)) return 0;

    ilqgb_get_int(h, "result", &result);
    ilqgb_get_int(h, "iterations", &o->iterations);
    ilqgb_get_int(h, "cur", &cur);
    ilqgb_get_int(h, "n_linesearch", &n_ls);
    ilqgb_get(h, "cost", &o->cost);
    ilqgb_get(h, "new_cost", &o->new_cost);
    ilqgb_get(h, "dcost", &o->dcost);
    ilqgb_get(h, "expected", &o->expected);
    ilqgb_get(h, "lambda", &o->lambda);
    ilqgb_get(h, "g_norm", &o->g_norm);
    ilqgb_get(h, "dV0", &o->dV[0]);
    ilqgb_get(h, "dV1", &o->dV[1]);
    ilqgb_get(h, "w_pen_l", &o->w_pen_l);
    ilqgb_get(h, "w_pen_f", &o->w_pen_f);

    /* the device started with the caller's nominal in buffer 0: an odd number of accepted steps leaves the nominal
       in buffer 1, which is the reference's pointer swap (iLQG.c:322, 381-386) */
    if (cur) makeCandidateNominal(o, 0);
    if (pull_traj(h, o->nominal, "x", "u", T) || pull_traj(h, o->candidates[0], "x_cand", "u_cand", T)) return 0;
    {
        double *l = (double *)malloc(sizeof(double) * (size_t)T * (N_U + N_U * N_X + 2 * (n_r + 1)) + sizeof(double) * 2 * (n_f + 1));
        double *L = l + (size_t)T * N_U, *mr = L + (size_t)T * N_U * N_X, *lr = mr + (size_t)T * (n_r + 1), *mf = lr + (size_t)T * (n_r + 1), *lf = mf + n_f + 1;
        ilqgb_get(h, "l", l);
        ilqgb_get(h, "L", L);
        if (n_r) { ilqgb_get(h, "mu_r", mr); ilqgb_get(h, "last_r", lr); }
        if (n_f) { ilqgb_get(h, "mu_f", mf); ilqgb_get(h, "last_f", lf); }
        for (k = 0; k < T; k++) {
            memcpy(o->nominal->t[k].l, l + (size_t)k * N_U, sizeof o->nominal->t[k].l);
            memcpy(o->nominal->t[k].L, L + (size_t)k * N_U * N_X, sizeof o->nominal->t[k].L);
            for (i = 0; i < n_r; i++) {
                int a, b;
                mult_layout(n_r, n_le, i, &a, &b);
                ((double *)&o->multipliers.t[k])[a] = mr[(size_t)k * n_r + i];
                ((double *)&o->multipliers.t[k])[b] = lr[(size_t)k * n_r + i];
            }
        }
        for (i = 0; i < n_f; i++) {
            int a, b;
            mult_layout(n_f, n_fe, i, &a, &b);
            ((double *)&o->multipliers.f)[a] = mf[i];
            ((double *)&o->multipliers.f)[b] = lf[i];
        }
        free(l);
    }
    if ((o->log_linesearch || o->log_z || o->log_cost) && o->max_iter > 0) {
        int *a = (int *)malloc(sizeof(int) * g_trace_len);
        double *z = (double *)malloc(sizeof(double) * 2 * g_trace_len), *c = z + g_trace_len;
        ilqgb_get_int(h, "tr_alpha", a);
        ilqgb_get(h, "tr_z", z);
        ilqgb_get(h, "tr_newcost", c);
        for (k = 0; k < n_ls && k < o->max_iter; k++) {
            if (o->log_linesearch) o->log_linesearch[k] = a[k];
            if (o->log_z) o->log_z[k] = z[k];
            if (o->log_cost) o->log_cost[k] = c[k];
        }
        free(a);
        free(z);
    }
    return result;
}
