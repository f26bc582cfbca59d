/* iLQG_mex_b200.c -- MATLAB / Octave gateway over the batched B200 solver (include/ilqg_b200.h).
 *
 *     [success, x, u, cost] = iLQG_b200(x0, u_nom, params, opt)
 *
 * Same call, argument checks, error identifiers and outputs as the reference's gateway (iLQG_mex.c:19-144): x0 holds the n
 * initial states, u_nom is m x (N-1), params is a scalar struct with one vector per generated parameter (length 1, k, or
 * N for a [k]-indexed one), opt a scalar struct of solver options (names and validation of setOptParam, iLQG.c:91-216).
 * success is 1 / 0 as returned by iLQG(), x is n x N, u is m x (N-1), cost the final cost.
 *
 * Batched extension (what the GPU is for): u_nom may be m x (N-1) x B; then x0 holds n*B values (n x B), the outputs
 * become 1 x B, n x N x B, m x (N-1) x B, 1 x B, and a parameter may be given as a k x B matrix (one column per problem,
 * not for [k]-indexed parameters) instead of a vector shared by all problems.  B = 1 is exactly the reference call.
 *
 * Build (Octave):  mkoctfile --mex iLQG_mex_b200.c -I<repo>/include -L<repo>/ddp-generator_b200/lib -lilqg_b200_<problem>_ddp<d>
 * Build (MATLAB):  mex iLQG_mex_b200.c -I<repo>/include -L<repo>/ddp-generator_b200/lib -lilqg_b200_<problem>_ddp<d>
 * There is no CPU path: without a CUDA device the call fails with the identifier iLQG:gpu. */
#include <stdio.h>

#include "mex.h"

#include "ilqg_b200.h"

#define MAX_GATEWAY_PARAMS 256

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    const mxArray *mxParams, *mxOpt, *f;
    const mxArray *pfield[MAX_GATEWAY_PARAMS];
    int per_problem[MAX_GATEWAY_PARAMS];
    const mwSize *du;
    const char *err, *name;
    ilqgb_handle *h;
    double *success, *x, *u, *cost;
    int *result;
    int n, m, T, B, i, k, b, si, np;
    size_t rows, cols;

    if (nrhs != 4) {
        mexErrMsgIdAndTxt("MATLAB:minrhs", "wrong number of arguments (expected: x0, u_nom, params, opt_params)");
        return;
    }
    if (nlhs != 4) {
        mexErrMsgIdAndTxt("MATLAB:minlhs", "wrong number of return values (expected: success, x_new, u_new, new_cost)");
        return;
    }

    /* dimensions: iLQG_mex.c:36-44, plus the optional third (batch) dimension of u_nom */
    du = mxGetDimensions(prhs[1]);
    m = (int)du[0];
    T = (int)du[1];
    B = mxGetNumberOfDimensions(prhs[1]) > 2 ? (int)du[2] : 1;
    if (B < 1 || mxGetNumberOfDimensions(prhs[1]) > 3) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "u_nom must be m x (N-1) or m x (N-1) x B");
        return;
    }
    n = (int)(mxGetNumberOfElements(prhs[0]) / (size_t)B);
    if (n != ilqgb_nx() || mxGetNumberOfElements(prhs[0]) != (size_t)n * (size_t)B) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "wrong number of states (%d expected)", ilqgb_nx());
        return;
    }
    if (m != ilqgb_nu()) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "wrong number of inputs (%d expected)", ilqgb_nu());
        return;
    }
    if (T < 1) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "wrong number of elements in u_nom (%dx%d expected)", m, 1);
        return;
    }
    if (!mxIsDouble(prhs[0]) || !mxIsDouble(prhs[1]) || mxIsSparse(prhs[0]) || mxIsSparse(prhs[1])) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "x0 and u_nom must be full double arrays");
        return;
    }

    mxParams = prhs[2];
    if (!mxIsStruct(mxParams) || mxGetNumberOfElements(mxParams) != 1) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "Input 3 must be a scalar struct.\n");
        return;
    }
    mxOpt = prhs[3];
    if (!mxIsStruct(mxOpt) || mxGetNumberOfElements(mxOpt) != 1) {
        mexErrMsgIdAndTxt("MATLAB:dimagree", "Input 4 must be a scalar struct of optimization parameters.\n");
        return;
    }

    /* options are checked before anything is allocated (iLQG_mex.c:59-67; the check itself is setOptParam's) */
    for (i = 0; i < mxGetNumberOfFields(mxOpt); i++) {
        f = mxGetFieldByNumber(mxOpt, 0, i);
        name = mxGetFieldNameByNumber(mxOpt, i);
        if (!mxIsDouble(f) || mxIsSparse(f)) {
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Error setting optimization parameter '%s': %s.\n", name, "value must be a full double array");
            return;
        }
        err = ilqgb_validate_opt(name, mxGetPr(f), (int)mxGetNumberOfElements(f));
        if (err) {
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Error setting optimization parameter '%s': %s.\n", name, err);
            return;
        }
    }

    /* parameters by name (iLQG_mex.c:70-84) */
    np = ilqgb_n_params();
    if (np > MAX_GATEWAY_PARAMS) {
        mexErrMsgIdAndTxt("iLQG:gateway", "too many parameters (%d)", np);
        return;
    }
    for (i = 0; i < np; i++) {
        name = ilqgb_param_name(i);
        si = ilqgb_param_size(i) == -1 ? T + 1 : ilqgb_param_size(i);
        f = mxGetField(mxParams, 0, name);
        if (f == NULL) {
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Parameter name '%s' is not member of parameters struct.\n", name);
            return;
        }
        rows = mxGetM(f);
        cols = mxGetN(f);
        per_problem[i] = 0;
        if (!mxIsSparse(f) && mxIsDouble(f) && B > 1 && ilqgb_param_size(i) != -1 && rows == (size_t)si && cols == (size_t)B)
            per_problem[i] = 1; /* one column per problem */
        else if (mxIsSparse(f) || !mxIsDouble(f) || (rows != 1 && cols != 1) || rows * cols != (size_t)si) {
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Parameter name '%s' must be a vector length %d.\n", name, si);
            return;
        }
        pfield[i] = f;
    }

    h = ilqgb_create(0, B, T, 0, NULL);
    if (h == NULL) {
        mexErrMsgIdAndTxt("iLQG:gpu", "%s", ilqgb_last_error(NULL));
        return;
    }
    for (i = 0; i < mxGetNumberOfFields(mxOpt); i++) {
        f = mxGetFieldByNumber(mxOpt, 0, i);
        ilqgb_set_opt(h, mxGetFieldNameByNumber(mxOpt, i), mxGetPr(f), (int)mxGetNumberOfElements(f));
    }
    for (i = 0; i < np; i++) {
        si = ilqgb_param_size(i) == -1 ? T + 1 : ilqgb_param_size(i);
        if (per_problem[i] ? ilqgb_set_param_batch(h, i, mxGetPr(pfield[i]), si) : ilqgb_set_param(h, i, mxGetPr(pfield[i]), si)) {
            static char msg[512];
            snprintf(msg, sizeof msg, "%s", ilqgb_last_error(h));
            ilqgb_destroy(h);
            mexErrMsgIdAndTxt("iLQG:gpu", "%s", msg);
            return;
        }
    }

    /* outputs (iLQG_mex.c:87-98) */
    plhs[0] = mxCreateDoubleMatrix(1, (mwSize)B, mxREAL);
    plhs[3] = mxCreateDoubleMatrix(1, (mwSize)B, mxREAL);
    if (B == 1) {
        plhs[1] = mxCreateDoubleMatrix((mwSize)n, (mwSize)T + 1, mxREAL);
        plhs[2] = mxCreateDoubleMatrix((mwSize)m, (mwSize)T, mxREAL);
    } else {
        mwSize dx[3], dv[3];
        dx[0] = (mwSize)n; dx[1] = (mwSize)T + 1; dx[2] = (mwSize)B;
        dv[0] = (mwSize)m; dv[1] = (mwSize)T; dv[2] = (mwSize)B;
        plhs[1] = mxCreateNumericArray(3, dx, mxDOUBLE_CLASS, mxREAL);
        plhs[2] = mxCreateNumericArray(3, dv, mxDOUBLE_CLASS, mxREAL);
    }
    success = mxGetPr(plhs[0]);
    x = mxGetPr(plhs[1]);
    u = mxGetPr(plhs[2]);
    cost = mxGetPr(plhs[3]);
    result = (int *)mxMalloc(sizeof(int) * (size_t)B);

    /* init_opt, initial rollout, iLQG(), copy-out (iLQG_mex.c:106-137), for the whole batch on the GPU */
    if (ilqgb_solve_host(h, mxGetPr(prhs[0]), mxGetPr(prhs[1]), x, u, cost, NULL, result, NULL)) {
        static char msg[512];
        snprintf(msg, sizeof msg, "%s", ilqgb_last_error(h));
        mxFree(result);
        ilqgb_destroy(h);
        mexErrMsgIdAndTxt("iLQG:gpu", "%s", msg);
        return;
    }
    for (b = 0; b < B; b++) {
        success[b] = result[b] == 1 ? 1.0 : 0.0;
        if (result[b] < 0) { /* non-finite initial rollout: the reference returns success 0 and leaves x, u zero (iLQG_mex.c:116-118) */
            for (k = 0; k < n * (T + 1); k++) x[(size_t)b * n * (T + 1) + k] = 0.0;
            for (k = 0; k < m * T; k++) u[(size_t)b * m * T + k] = 0.0;
        }
    }
    mexPrintf("iLQG (B200): %d problem(s), horizon %d, %ld kernel launches\n", B, T, ilqgb_launch_count(h));
    mxFree(result);
    ilqgb_destroy(h);
}
