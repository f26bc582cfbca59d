/* oracle/harness.h -- flat C interface around a reference-ABI iLQG build (TEST INFRASTRUCTURE).
 *
 * The same harness.c is linked either against the UNMODIFIED reference solver core compiled in place from
 * /root/reference (oracle/_ref/libref_*.so) or against oracle/port_*.c, the independent restatement
 * (oracle/_build/libport_*.so).  It plays the role of the reference's only shipped caller, iLQG_mex.c:19-144,
 * minus MATLAB: allocate, init_opt, copy u_nom, initial rollout, iLQG(), copy out.  Python tests and
 * bench.py's cpu_baseline leg load it with ctypes.  Nothing in the product path may use it. */
#ifndef ORACLE_HARNESS_H
#define ORACLE_HARNESS_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct HSolver HSolver;

/* static facts of this build */
int h_nx(void);
int h_nu(void);
int h_full_ddp(void);
int h_n_params(void);
const char *h_param_name(int i);
int h_param_size(int i);          /* 1, k>1, or -1 (= n_hor+1 entries) */
int h_has_spies(void);
const char *h_kind(void);         /* "reference" or "port" */

HSolver *h_create(int n_hor);
void h_destroy(HSolver *s);
const char *h_set_opt(HSolver *s, const char *name, const double *v, int n);   /* setOptParam */
int h_set_param(HSolver *s, int i, const double *v, int n);
int h_init(HSolver *s, const double *x0, const double *u0);  /* init_opt + initial rollout; 1 ok */
int h_solve(HSolver *s);                                     /* iLQG(); its return value */

/* single phases on the current nominal trajectory (public reference functions, iLQG.h:78-88) */
int h_calc_derivs(HSolver *s);
int h_back_pass(HSolver *s);
int h_line_search(HSolver *s, int iter);
int h_forward_pass(HSolver *s, double alpha, double *csum, int cost_only);  /* into candidates[0] */
void h_make_candidate_nominal(HSolver *s);
int h_update_multipliers(HSolver *s, int init);
void h_set_scalar(HSolver *s, const char *name, double v);

/* results: flat copies of the nominal trajectory, step-major ([k][i]); returns number of doubles written */
int h_get(HSolver *s, const char *field, double *out);
double h_scalar(HSolver *s, const char *name);

/* traces recorded by the spies (line_search / back_pass / boxQP interposed with -D, no source patch) */
int h_trace_len(HSolver *s);                 /* number of line searches */
int h_trace(HSolver *s, const char *what, double *out);   /* per line search: lambda, g_norm, dV0, dV1, cost,
                                                             success, new_cost, dcost, expected, alpha_idx, iter */
int h_bp_trace_len(HSolver *s);              /* number of back_pass calls */
int h_bp_trace(HSolver *s, const char *what, double *out);  /* lambda, result */
int h_qp_trace_len(HSolver *s);              /* number of boxQP calls recorded (capped) */
int h_qp_trace(HSolver *s, int *ret_code, int *n_free, int *is_clamped /* [len][nu] */);
void h_qp_trace_enable(HSolver *s, int cap);

/* batch solve on host threads, one solver per problem (single-threaded build is re-entrant) */
int h_solve_batch(int B, int n_hor, const double *x0 /*[B][nx]*/, const double *u0 /*[B][T][nu]*/,
                  const double *params_flat, const char *const *opt_names, const double *opt_vals, int n_opts,
                  int n_threads, double *cost_out, int *iter_out, int *nls_out, int *result_out,
                  double *x_out /* [B][T+1][nx] or NULL */, double *u_out /* [B][T][nu] or NULL */);

#ifdef __cplusplus
}
#endif
#endif
