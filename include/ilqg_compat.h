/* ilqg_compat.h -- source/ABI-compatible declarations of the reference's single-problem solver interface.
 *
 * A translation unit written against the reference's headers (iLQG.h, matMult.h, back_pass.h, line_search.h,
 * boxQP.h, cholesky.h -- thin shims with those names sit next to this file) compiles unchanged against this
 * one and links either to the B200 library (libilqg_b200_<problem>.so, GPU execution) or to the CPU oracle.
 * Struct member order is part of the ABI and therefore identical to the reference (iLQG.h:31-76); the
 * problem-sized types (trajEl_t, trajFin_t, traj_t, multipliers_t, N_X, N_U, sizeofQ*) come from the
 * generated iLQG_problem.h exactly as in the reference (iLQG.h:8).
 *
 * Entry points and what they replace (reference file:line):
 *   standard_parameters  iLQG.c:57-78      option defaults
 *   setOptParam          iLQG.c:91-216     named option with validation; returns NULL or a static message
 *   iLQG                 iLQG.c:224-379    the solve loop; 1 = converged, 0 = not
 *   makeCandidateNominal iLQG.c:381-386
 *   back_pass            back_pass.c:38-257
 *   line_search          line_search.c:33-78
 *   boxQP                boxQP.c:39-238
 *   cholesky_tri(_inv)   cholesky.c:6-27, 51-74
 *   addMulVec/addSquareTri/addMul2Tri  matMult.c:3-72
 *   forward_pass, calc_derivs, init_opt, update_multipliers, clampU, get_g_size, calcG
 *                        generated problem code (iLQG_func.tem:68-221, 402-521)
 */
#ifndef ILQG_COMPAT_H
#define ILQG_COMPAT_H

/* compile-time switches, same names and defaults as the reference (iLQG.h:4-6, 12-25) */
#ifndef FULL_DDP
#define FULL_DDP 1
#endif
#ifndef PRNT
#define PRNT printf
#endif
#ifndef MULTI_THREADED
#define MULTI_THREADED 0
#endif
#ifndef NUMBER_OF_THREADS
#define NUMBER_OF_THREADS 1
#endif
#if MULTI_THREADED
#include <pthread.h>
#endif

#include "iLQG_problem.h"

/* `#if PREFIX1(FLAG)==1` detects a flag defined without a value (iLQG.h:27-28) */
#define DO_PREFIX1(VAL) 1##VAL
#define PREFIX1(VAL) DO_PREFIX1(VAL)

/* matrix index helpers (matMult.h:4-9): column-major dense, packed upper triangle column by column */
#define MAT_IDX(r, c, nr) ((r) + (c) * (nr))
#define UTRI_MAT_IDX(r, c) (((c) * ((c) + 1)) / 2 + (r))
#define SYMTRI_MAT_IDX(r, c) (((r) > (c)) ? UTRI_MAT_IDX(c, r) : UTRI_MAT_IDX(r, c))

#ifdef __cplusplus
extern "C" {
#endif

typedef struct paramDesc {
    char *name; /* name of the field in the caller's parameter struct */
    int size;   /* 1, fixed length k > 1, or -1 for one value per timestep (n_hor + 1) */
    int is_var;
} tParamDesc;

typedef struct optSet {
    /* problem instance */
    int n_hor;
    int debug_level;
    double *x0, new_cost, cost, dcost, lambda, g_norm, expected;
    double **p;
    /* options */
    const double *alpha;
    int n_alpha;
    double lambdaMax;
    double lambdaMin;
    double lambdaInit;
    double dlambdaInit;
    double lambdaFactor;
    int max_iter;
    double tolGrad;
    double tolFun;
    double tolConstraint;
    double zMin;
    int regType;
    /* results and logs */
    int iterations;
    int *log_linesearch;
    double *log_z;
    double *log_cost;
    double dV[2];
    /* augmented-Lagrangian penalty weights */
    double w_pen_l;
    double w_pen_f;
    double w_pen_max_l;
    double w_pen_max_f;
    double w_pen_init_l;
    double w_pen_init_f;
    double w_pen_fact1;
    double w_pen_fact2;
    /* trajectory storage (caller-allocated, iLQG_mex.c:100-103) */
    traj_t *nominal;
    traj_t *candidates[NUMBER_OF_THREADS];
    traj_t trajectories[NUMBER_OF_THREADS + 1];
    multipliers_t multipliers;
} tOptSet;

#define INIT_OPTSET {0}

void printParams(double **p, int k);
void standard_parameters(tOptSet *o);
int iLQG(tOptSet *o);
char *setOptParam(tOptSet *o, const char *name, const double *value, const int n);
void makeCandidateNominal(tOptSet *o, int idx);
int back_pass(tOptSet *o);
int line_search(tOptSet *o, int iter);

/* generated per problem */
int forward_pass(traj_t *c, tOptSet *o, double alpha, double *csum, int cost_only);
int calc_derivs(tOptSet *o);
int init_opt(tOptSet *o);
int update_multipliers(tOptSet *o, int init);
void clampU(double *u, trajEl_t *t, int k, double **p, int N);
int get_g_size();
int calcG(double g[], trajEl_t *t, int k, double *p[]);
extern int n_params;
extern int n_vars;
extern tParamDesc *paramdesc[];

/* dense helpers */
int boxQP(double *H, const double *g, const double *lower, const double *upper, double *x, double *Hfree,
          double *L, double *grad, double *grad_clamped, double *search, int *is_clamped, int *n_free_,
          double *invHfree, const int n);
int cholesky_tri(const double *A, int n, double *L);
void cholesky_tri_inv(const double *L_, double *invA, const int n, double *x);
void addMulVec(double base[], const double a[], const double b[], const int n_r, const int n_c);
void addSquareTri(double base[], const double b[], const double a[], const int n_r, const int n_c, double ba[]);
void addMul2Tri(double base[], const double b[], const double a[], const int n_ra, const int n_ca,
                const double c[], const int n_rc, const int n_cc, double bc[]);

#if MULTI_THREADED
extern pthread_mutex_t step_mutex;
extern pthread_cond_t next_step_condition;
extern int step_calc_done;
#endif

#ifdef __cplusplus
}
#endif

/* the generated code calls these unqualified (iLQG.h:90-96); NaN-propagation follows the comparison */
static inline double max(double a, double b) { return (a > b) ? a : b; }
static inline double min(double a, double b) { return (a < b) ? a : b; }

#endif /* ILQG_COMPAT_H */
