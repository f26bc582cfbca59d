/* oracle/harness.c -- TEST INFRASTRUCTURE.  See harness.h.
 *
 * Spies: the reference core is compiled with  -Dline_search=h_spy_line_search -Dback_pass=h_spy_back_pass
 * on iLQG.c and -DboxQP=h_spy_boxQP on back_pass.c (no source patch; SURVEY.md 8c "zero-patch hooks").
 * The spies record what the parity tests compare (lambda / alpha / active-set sequences) and forward to the
 * real functions, which this file sees under their true names. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <pthread.h>

#include "iLQG.h"
#ifndef H_NO_PHASES /* the solver's phases are public functions of the reference (back_pass.h:7, line_search.h:6) */
#include "line_search.h"
#include "back_pass.h"
#include "boxQP.h"
#endif
#include "harness.h"

#ifndef H_KIND
#define H_KIND "reference"
#endif
#ifndef H_SPIES
#define H_SPIES 1
#endif

/* diagnostics hook: the reference prints through PRNT (iLQG.h:12-14); swallow everything */
int h_prnt(const char *fmt, ...) { (void)fmt; return 0; }

#define LS_FIELDS 11
struct HSolver {
    tOptSet o;
    int n_hor;
    double *x0;
    double **p;
    int *p_len;
    double *alpha_own;
    int max_iter_cap;
    /* traces */
    int n_ls, cap_ls;
    double *ls;     /* [cap][LS_FIELDS] */
    int n_bp, cap_bp;
    double *bp;     /* [cap][2] */
    int qp_on, n_qp, cap_qp;
    int *qp_ret, *qp_nfree, *qp_clamped;
};

static __thread HSolver *g_cur = NULL;

int h_nx(void) { return N_X; }
int h_nu(void) { return N_U; }
int h_full_ddp(void) { return FULL_DDP; }
int h_n_params(void) { return n_params; }
const char *h_param_name(int i) { return paramdesc[i]->name; }
int h_param_size(int i) { return paramdesc[i]->size; }
int h_has_spies(void) { return H_SPIES; }
const char *h_kind(void) { return H_KIND; }

HSolver *h_create(int n_hor)
{
    HSolver *s = (HSolver *)calloc(1, sizeof(HSolver));
    int i;
    s->n_hor = n_hor;
    s->o.n_hor = n_hor;
    s->x0 = (double *)calloc(N_X, sizeof(double));
    s->o.x0 = s->x0;
    s->p = (double **)calloc(n_params > 0 ? n_params : 1, sizeof(double *));
    s->p_len = (int *)calloc(n_params > 0 ? n_params : 1, sizeof(int));
    for (i = 0; i < n_params; i++) {
        int len = paramdesc[i]->size == -1 ? n_hor + 1 : paramdesc[i]->size;
        s->p[i] = (double *)calloc(len, sizeof(double));
        s->p_len[i] = len;
    }
    s->o.p = s->p;
    for (i = 0; i < NUMBER_OF_THREADS + 1; i++)
        s->o.trajectories[i].t = (trajEl_t *)calloc(n_hor > 0 ? n_hor : 1, sizeof(trajEl_t));
    s->o.multipliers.t = (multipliersEl_t *)calloc(n_hor + 1, sizeof(multipliersEl_t) + 1);
    standard_parameters(&s->o);
    s->o.debug_level = 0;
    s->max_iter_cap = 0;
    s->o.w_pen_l = s->o.w_pen_f = 0.0;
    return s;
}

static void ensure_logs(HSolver *s)
{
    int need = s->o.max_iter + 1;
    if (need > s->max_iter_cap) {
        s->o.log_linesearch = (int *)realloc(s->o.log_linesearch, need * sizeof(int));
        s->o.log_z = (double *)realloc(s->o.log_z, need * sizeof(double));
        s->o.log_cost = (double *)realloc(s->o.log_cost, need * sizeof(double));
        s->max_iter_cap = need;
    }
    memset(s->o.log_linesearch, 0, s->max_iter_cap * sizeof(int));
}

void h_destroy(HSolver *s)
{
    int i;
    if (!s) return;
    for (i = 0; i < n_params; i++) free(s->p[i]);
    free(s->p); free(s->p_len); free(s->x0); free(s->alpha_own);
    for (i = 0; i < NUMBER_OF_THREADS + 1; i++) free(s->o.trajectories[i].t);
    free(s->o.multipliers.t);
    free(s->o.log_linesearch); free(s->o.log_z); free(s->o.log_cost);
    free(s->ls); free(s->bp); free(s->qp_ret); free(s->qp_nfree); free(s->qp_clamped);
    free(s);
}

const char *h_set_opt(HSolver *s, const char *name, const double *v, int n)
{
    if (strcmp(name, "alpha") == 0) {   /* setOptParam keeps the caller's pointer (iLQG.c:101): own a copy */
        double *cp = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
        const char *err;
        memcpy(cp, v, sizeof(double) * n);
        err = setOptParam(&s->o, name, cp, n);
        if (err) { free(cp); return err; }
        free(s->alpha_own);
        s->alpha_own = cp;
        return NULL;
    }
    return setOptParam(&s->o, name, v, n);
}

int h_set_param(HSolver *s, int i, const double *v, int n)
{
    if (i < 0 || i >= n_params || n != s->p_len[i]) return 0;
    memcpy(s->p[i], v, sizeof(double) * n);
    return 1;
}

int h_init(HSolver *s, const double *x0, const double *u0)
{
    int k, i;
    g_cur = s;
    memcpy(s->x0, x0, sizeof(double) * N_X);
    s->n_ls = s->n_bp = s->n_qp = 0;
    s->o.w_pen_l = s->o.w_pen_f = 0.0; /* a fresh INIT_OPTSET per call, as in the mex gateway (iLQG_mex.c:24) */
    ensure_logs(s);
    if (!init_opt(&s->o)) return 0;
    for (k = 0; k < s->n_hor; k++)
        for (i = 0; i < N_U; i++)
            s->o.nominal->t[k].u[i] = u0[i + k * N_U];
    if (!forward_pass(s->o.candidates[0], &s->o, 0.0, &s->o.cost, 0)) return 0;
    makeCandidateNominal(&s->o, 0);
    return 1;
}

int h_solve(HSolver *s)
{
    g_cur = s;
    ensure_logs(s);
    return iLQG(&s->o);
}

#ifndef H_NO_PHASES
int h_calc_derivs(HSolver *s) { g_cur = s; return calc_derivs(&s->o); }
int h_back_pass(HSolver *s) { g_cur = s; return back_pass(&s->o); }
int h_line_search(HSolver *s, int iter) { g_cur = s; ensure_logs(s); return line_search(&s->o, iter); }
int h_update_multipliers(HSolver *s, int init) { return update_multipliers(&s->o, init); }
#else /* a library that only exports the entry points the reference's mex gateway binds */
int h_calc_derivs(HSolver *s) { (void)s; return -1; }
int h_back_pass(HSolver *s) { (void)s; return -1; }
int h_line_search(HSolver *s, int iter) { (void)s; (void)iter; return -1; }
int h_update_multipliers(HSolver *s, int init) { (void)s; (void)init; return -1; }
#endif
/* user outputs g of step k of the nominal trajectory (get_g_size / calcG, iLQG.h:87-88) */
int h_calc_g(HSolver *s, int k, double *g) { return calcG(g, &s->o.nominal->t[k], k, s->o.p) ? get_g_size() : -1; }
#ifndef H_NO_PHASES
/* clampU on step k of the nominal trajectory (its state-dependent auxiliaries are those of the last forward pass) */
void h_clamp_u(HSolver *s, int k, double *u) { clampU(u, &s->o.nominal->t[k], k, s->o.p, s->o.n_hor); }
#endif
int h_forward_pass(HSolver *s, double alpha, double *csum, int cost_only)
{
    g_cur = s;
    return forward_pass(cost_only ? s->o.nominal : s->o.candidates[0], &s->o, alpha, csum, cost_only);
}
void h_make_candidate_nominal(HSolver *s) { makeCandidateNominal(&s->o, 0); }
void h_set_scalar(HSolver *s, const char *name, double v)
{
    tOptSet *o = &s->o;
    if (!strcmp(name, "lambda")) o->lambda = v;
    else if (!strcmp(name, "cost")) o->cost = v;
    else if (!strcmp(name, "w_pen_l")) o->w_pen_l = v;
    else if (!strcmp(name, "w_pen_f")) o->w_pen_f = v;
}

double h_scalar(HSolver *s, const char *name)
{
    tOptSet *o = &s->o;
    if (!strcmp(name, "cost")) return o->cost;
    if (!strcmp(name, "new_cost")) return o->new_cost;
    if (!strcmp(name, "dcost")) return o->dcost;
    if (!strcmp(name, "expected")) return o->expected;
    if (!strcmp(name, "lambda")) return o->lambda;
    if (!strcmp(name, "g_norm")) return o->g_norm;
    if (!strcmp(name, "iterations")) return (double)o->iterations;
    if (!strcmp(name, "dV0")) return o->dV[0];
    if (!strcmp(name, "dV1")) return o->dV[1];
    if (!strcmp(name, "w_pen_l")) return o->w_pen_l;
    if (!strcmp(name, "w_pen_f")) return o->w_pen_f;
    if (!strcmp(name, "sizeof_trajEl")) return (double)sizeof(trajEl_t);
    if (!strcmp(name, "n_linesearch")) {   /* loop passes that reached line_search (SURVEY 8d metric) */
        int i, c = 0;
        for (i = 0; i < o->max_iter && i < s->max_iter_cap; i++) c += (o->log_linesearch[i] != 0);
        return (double)c;
    }
    return NAN;
}

#define COPY_STEPS(FIELD, LEN) do { for (k = 0; k < T; k++) memcpy(out + (size_t)k * (LEN), nom->t[k].FIELD, sizeof(double) * (LEN)); return T * (LEN); } while (0)

int h_get(HSolver *s, const char *field, double *out)
{
    traj_t *nom = s->o.nominal;
    int T = s->n_hor, k;
    if (!strcmp(field, "x")) {
        for (k = 0; k < T; k++) memcpy(out + (size_t)k * N_X, nom->t[k].x, sizeof(double) * N_X);
        memcpy(out + (size_t)T * N_X, nom->f.x, sizeof(double) * N_X);
        return (T + 1) * N_X;
    }
    if (!strcmp(field, "u")) COPY_STEPS(u, N_U);
    if (!strcmp(field, "l")) COPY_STEPS(l, N_U);
    if (!strcmp(field, "L")) COPY_STEPS(L, N_U * N_X);
    if (!strcmp(field, "lower")) COPY_STEPS(lower, N_U);
    if (!strcmp(field, "upper")) COPY_STEPS(upper, N_U);
    if (!strcmp(field, "lower_sign")) COPY_STEPS(lower_sign, N_U);
    if (!strcmp(field, "upper_sign")) COPY_STEPS(upper_sign, N_U);
    if (!strcmp(field, "lower_hx")) COPY_STEPS(lower_hx, N_U * N_X);
    if (!strcmp(field, "upper_hx")) COPY_STEPS(upper_hx, N_U * N_X);
    if (!strcmp(field, "fx")) COPY_STEPS(fx, N_X * N_X);
    if (!strcmp(field, "fu")) COPY_STEPS(fu, N_X * N_U);
    if (!strcmp(field, "cu")) COPY_STEPS(cu, N_U);
    if (!strcmp(field, "cuu")) COPY_STEPS(cuu, sizeofQuu);
    if (!strcmp(field, "cxu")) COPY_STEPS(cxu, sizeofQxu);
#if FULL_DDP
    if (!strcmp(field, "fxx")) COPY_STEPS(fxx, N_X * sizeofQxx);
    if (!strcmp(field, "fuu")) COPY_STEPS(fuu, N_X * sizeofQuu);
    if (!strcmp(field, "fxu")) COPY_STEPS(fxu, N_X * sizeofQxu);
#endif
    if (!strcmp(field, "c")) {
        for (k = 0; k < T; k++) out[k] = nom->t[k].c;
        out[T] = nom->f.c;
        return T + 1;
    }
    if (!strcmp(field, "cx")) {
        for (k = 0; k < T; k++) memcpy(out + (size_t)k * N_X, nom->t[k].cx, sizeof(double) * N_X);
        memcpy(out + (size_t)T * N_X, nom->f.cx, sizeof(double) * N_X);
        return (T + 1) * N_X;
    }
    if (!strcmp(field, "cxx")) {
        for (k = 0; k < T; k++) memcpy(out + (size_t)k * sizeofQxx, nom->t[k].cxx, sizeof(double) * sizeofQxx);
        memcpy(out + (size_t)T * sizeofQxx, nom->f.cxx, sizeof(double) * sizeofQxx);
        return (T + 1) * sizeofQxx;
    }
    if (!strcmp(field, "mult_f")) {   /* final multipliers struct as raw doubles */
        int n = (int)(sizeof(multipliersFin_t) / sizeof(double));
        if (n) memcpy(out, &s->o.multipliers.f, sizeof(double) * n);
        return n;
    }
    if (!strcmp(field, "mult_t")) {
        int n = (int)(sizeof(multipliersEl_t) / sizeof(double));
        for (k = 0; k < T && n; k++) memcpy(out + (size_t)k * n, &s->o.multipliers.t[k], sizeof(double) * n);
        return T * n;
    }
    if (!strcmp(field, "log_linesearch")) {
        for (k = 0; k < s->o.max_iter; k++) out[k] = (double)s->o.log_linesearch[k];
        return s->o.max_iter;
    }
    return -1;
}

#if !defined(H_NO_PHASES) && !defined(H_NO_SPY_FUNCS)
/* ---- spies ------------------------------------------------------------------------------------------------ */
int h_spy_line_search(tOptSet *o, int iter)
{
    HSolver *s = g_cur;
    double *r = NULL;
    int ok;
    if (s && &s->o == o) {
        if (s->n_ls == s->cap_ls) {
            s->cap_ls = s->cap_ls ? 2 * s->cap_ls : 256;
            s->ls = (double *)realloc(s->ls, sizeof(double) * LS_FIELDS * s->cap_ls);
        }
        r = s->ls + (size_t)LS_FIELDS * s->n_ls++;
        r[0] = o->lambda; r[1] = o->g_norm; r[2] = o->dV[0]; r[3] = o->dV[1]; r[4] = o->cost;
    }
    ok = line_search(o, iter);
    if (r) {
        r[5] = ok; r[6] = o->new_cost; r[7] = o->dcost; r[8] = o->expected;
        r[9] = o->log_linesearch ? (double)o->log_linesearch[iter] : -1.0;
        r[10] = iter;
    }
    return ok;
}

int h_spy_back_pass(tOptSet *o)
{
    HSolver *s = g_cur;
    double lam = o->lambda;
    int res = back_pass(o);
    if (s && &s->o == o) {
        if (s->n_bp == s->cap_bp) {
            s->cap_bp = s->cap_bp ? 2 * s->cap_bp : 256;
            s->bp = (double *)realloc(s->bp, sizeof(double) * 2 * s->cap_bp);
        }
        s->bp[2 * s->n_bp] = lam;
        s->bp[2 * s->n_bp + 1] = res;
        s->n_bp++;
    }
    return res;
}

int h_spy_boxQP(double *H, const double *g, const double *lower, const double *upper, double *x, double *Hfree,
                double *L, double *grad, double *grad_clamped, double *search, int *is_clamped, int *n_free_,
                double *invHfree, const int n)
{
    HSolver *s = g_cur;
    int res = boxQP(H, g, lower, upper, x, Hfree, L, grad, grad_clamped, search, is_clamped, n_free_, invHfree, n);
    if (s && s->qp_on && s->n_qp < s->cap_qp) {
        int i;
        s->qp_ret[s->n_qp] = res;
        s->qp_nfree[s->n_qp] = n_free_[0];
        for (i = 0; i < N_U; i++) s->qp_clamped[(size_t)s->n_qp * N_U + i] = is_clamped[i];
        s->n_qp++;
    }
    return res;
}

#endif

static const char *LS_NAMES[LS_FIELDS] = {"lambda", "g_norm", "dV0", "dV1", "cost", "success", "new_cost",
                                          "dcost", "expected", "alpha_idx", "iter"};

int h_trace_len(HSolver *s) { return s->n_ls; }
int h_trace(HSolver *s, const char *what, double *out)
{
    int f, i;
    for (f = 0; f < LS_FIELDS; f++)
        if (!strcmp(what, LS_NAMES[f])) {
            for (i = 0; i < s->n_ls; i++) out[i] = s->ls[(size_t)i * LS_FIELDS + f];
            return s->n_ls;
        }
    return -1;
}
int h_bp_trace_len(HSolver *s) { return s->n_bp; }
int h_bp_trace(HSolver *s, const char *what, double *out)
{
    int f = !strcmp(what, "lambda") ? 0 : (!strcmp(what, "result") ? 1 : -1), i;
    if (f < 0) return -1;
    for (i = 0; i < s->n_bp; i++) out[i] = s->bp[2 * i + f];
    return s->n_bp;
}
void h_qp_trace_enable(HSolver *s, int cap)
{
    s->qp_on = cap > 0;
    s->cap_qp = cap;
    s->n_qp = 0;
    s->qp_ret = (int *)realloc(s->qp_ret, sizeof(int) * (cap > 0 ? cap : 1));
    s->qp_nfree = (int *)realloc(s->qp_nfree, sizeof(int) * (cap > 0 ? cap : 1));
    s->qp_clamped = (int *)realloc(s->qp_clamped, sizeof(int) * (cap > 0 ? cap : 1) * N_U);
}
int h_qp_trace_len(HSolver *s) { return s->n_qp; }
int h_qp_trace(HSolver *s, int *ret_code, int *n_free, int *is_clamped)
{
    memcpy(ret_code, s->qp_ret, sizeof(int) * s->n_qp);
    memcpy(n_free, s->qp_nfree, sizeof(int) * s->n_qp);
    memcpy(is_clamped, s->qp_clamped, sizeof(int) * s->n_qp * N_U);
    return s->n_qp;
}

/* ---- batch on host threads ------------------------------------------------------------------------------------ */
typedef struct {
    int B, n_hor, n_opts, next;
    const double *x0, *u0, *params_flat;
    const char *const *opt_names;
    const double *opt_vals;
    double *cost_out, *x_out, *u_out;
    int *iter_out, *nls_out, *result_out;
    pthread_mutex_t mu;
} batch_t;

static void *batch_worker(void *arg)
{
    batch_t *b = (batch_t *)arg;
    HSolver *s = h_create(b->n_hor);
    int i, off = 0;
    for (i = 0; i < b->n_opts; i++) h_set_opt(s, b->opt_names[i], b->opt_vals + i, 1);
    for (i = 0; i < n_params; i++) {
        memcpy(s->p[i], b->params_flat + off, sizeof(double) * s->p_len[i]);
        off += s->p_len[i];
    }
    for (;;) {
        int id, res;
        pthread_mutex_lock(&b->mu);
        id = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (id >= b->B) break;
        res = 0;
        if (h_init(s, b->x0 + (size_t)id * N_X, b->u0 + (size_t)id * b->n_hor * N_U))
            res = h_solve(s);
        else
            res = -1;
        b->cost_out[id] = s->o.cost;
        b->iter_out[id] = s->o.iterations;
        b->nls_out[id] = (int)h_scalar(s, "n_linesearch");
        b->result_out[id] = res;
        if (b->x_out) h_get(s, "x", b->x_out + (size_t)id * (b->n_hor + 1) * N_X);
        if (b->u_out) h_get(s, "u", b->u_out + (size_t)id * b->n_hor * N_U);
    }
    h_destroy(s);
    return NULL;
}

int h_solve_batch(int B, int n_hor, const double *x0, const double *u0, const double *params_flat,
                  const char *const *opt_names, const double *opt_vals, int n_opts, int n_threads,
                  double *cost_out, int *iter_out, int *nls_out, int *result_out, double *x_out, double *u_out)
{
    batch_t b;
    pthread_t *th;
    int i;
    if (n_threads < 1) n_threads = 1;
    memset(&b, 0, sizeof b);
    b.B = B; b.n_hor = n_hor; b.x0 = x0; b.u0 = u0; b.params_flat = params_flat;
    b.opt_names = opt_names; b.opt_vals = opt_vals; b.n_opts = n_opts;
    b.cost_out = cost_out; b.iter_out = iter_out; b.nls_out = nls_out; b.result_out = result_out;
    b.x_out = x_out; b.u_out = u_out;
    pthread_mutex_init(&b.mu, NULL);
    th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &b);
    for (i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&b.mu);
    return 0;
}
