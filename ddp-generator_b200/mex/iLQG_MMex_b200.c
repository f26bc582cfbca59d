/* iLQG_MMex_b200.c -- MATLAB / Octave gateway for SINGLE evaluations of the generated problem functions on the GPU:
 *
 *     out = iLQG<Name>MMex(x, u, params, mode, k, n_hor)
 *
 * the interface of the reference's iLQG_MMex.tem (generated per problem by make_iLQG_MMex.mac; it exists to cross-check the
 * generated functions and derivatives against other implementations).  Same six inputs, argument checks and messages
 * (iLQG_MMex.tem:36-79), same 17 modes and output shapes (81-226): 0 f, 1 L, 2 F, 3 Fx, 4 Fxx, 5 Lx, 6 Lu, 7 Lxx, 8 Luu, 9 Lxu,
 * 10 fx, 11 fu, 12 fxx, 13 fuu, 14 fxu, 15 y (0 x 1), 16 the clamped u.  k is 1-based as in MATLAB.  The evaluation runs in the
 * CUDA library (ilqgb_eval, include/ilqg_b200.h); there is no CPU path.
 *
 * Build (Octave):  mkoctfile --mex iLQG_MMex_b200.c -I<repo>/include -L<repo>/ddp-generator_b200/lib -lilqg_b200_<problem>_ddp<d> */
#include <math.h>
#include <stdio.h>

#include "mex.h"

#include "ilqg_b200.h"

static ilqgb_handle *g_h = NULL;
static int g_N = -1;

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    const int nx = ilqgb_nx(), nu = ilqgb_nu(), n_params = ilqgb_n_params();
    const mxArray *mxParam, *mxParams;
    int mode, i, k, N, si, n_out;
    size_t m_, n_;
    mwSize dims[3];
    double *x, *u, *out;

    if (nrhs != 6) { mexErrMsgTxt("wrong number of arguments (6 expected)"); return; }
    if (nlhs != 1) { mexErrMsgTxt("wrong number of return values (1 expected)"); return; }
    if (mxGetNumberOfElements(prhs[0]) != (size_t)nx) { mexErrMsgIdAndTxt("", "wrong number of elements in x (%d expected)", nx); return; }
    if (mxGetNumberOfElements(prhs[1]) != (size_t)nu) { mexErrMsgIdAndTxt("", "wrong number of elements in u (%d expected)", nu); return; }
    if (mxGetNumberOfElements(prhs[2]) != 1) { mexErrMsgTxt("wrong number of elements in params (1 expected)"); return; }
    if (mxGetNumberOfElements(prhs[3]) != 1) { mexErrMsgTxt("wrong number of elements in mode (1 expected)"); return; }
    if (mxGetNumberOfElements(prhs[4]) != 1) { mexErrMsgTxt("wrong number of elements in k (1 expected)"); return; }
    if (mxGetNumberOfElements(prhs[5]) != 1) { mexErrMsgTxt("wrong number of elements in n_hor (1 expected)"); return; }

    mode = (int)mxGetScalar(prhs[3]);
    k = (int)mxGetScalar(prhs[4]) - 1;
    N = (int)mxGetScalar(prhs[5]);
    x = mxGetPr(prhs[0]);
    u = mxGetPr(prhs[1]);

    mxParams = prhs[2];
    if (!mxIsStruct(mxParams)) mexErrMsgIdAndTxt("MATLAB:dimagree", "Input 3 must be a struct.\n");
    if (N < 1) mexErrMsgIdAndTxt("MATLAB:dimagree", "n_hor must be at least 1.\n");

    if (!g_h || g_N != N) { /* one evaluation point, horizon N (it sizes the [k]-indexed parameters) */
        if (g_h) ilqgb_destroy(g_h);
        g_h = ilqgb_create(0, 1, N, ILQGB_CHUNKS(1), NULL);
        g_N = g_h ? N : -1;
        if (!g_h) mexErrMsgIdAndTxt("iLQG:gpu", "%s\n", ilqgb_last_error(NULL));
    }
    for (i = 0; i < n_params; i++) {
        si = (ilqgb_param_size(i) == -1) ? N + 1 : ilqgb_param_size(i);
        if ((mxParam = mxGetField(mxParams, 0, ilqgb_param_name(i))) == NULL)
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Parameter name '%s' is not member of parameters struct.\n", ilqgb_param_name(i));
        m_ = mxGetM(mxParam);
        n_ = mxGetN(mxParam);
        if (mxIsSparse(mxParam) || !mxIsDouble(mxParam) || (m_ != 1 && n_ != 1) || (m_ * n_ != (size_t)si))
            mexErrMsgIdAndTxt("MATLAB:dimagree", "Parameter name '%s' must be a vector length %d.\n", ilqgb_param_name(i), si);
        if (ilqgb_set_param(g_h, i, mxGetPr(mxParam), si)) mexErrMsgIdAndTxt("iLQG:gpu", "%s\n", ilqgb_last_error(g_h));
    }

    n_out = ilqgb_eval_size(mode);
    if (n_out < 0) { /* the reference's switch has no default: an unknown mode returns without creating the output */
        plhs[0] = mxCreateDoubleMatrix(0, 0, mxREAL);
        return;
    }
    switch (mode) {
    case 0: plhs[0] = mxCreateDoubleMatrix(nx, 1, mxREAL); break;
    case 1: case 2: plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL); break;
    case 3: case 5: plhs[0] = mxCreateDoubleMatrix(1, nx, mxREAL); break;
    case 6: plhs[0] = mxCreateDoubleMatrix(1, nu, mxREAL); break;
    case 4: case 7: case 10: plhs[0] = mxCreateDoubleMatrix(nx, nx, mxREAL); break;
    case 8: plhs[0] = mxCreateDoubleMatrix(nu, nu, mxREAL); break;
    case 9: case 11: plhs[0] = mxCreateDoubleMatrix(nx, nu, mxREAL); break;
    case 12: dims[0] = nx; dims[1] = nx; dims[2] = nx; plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL); break;
    case 13: dims[0] = nu; dims[1] = nu; dims[2] = nx; plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL); break;
    case 14: dims[0] = nx; dims[1] = nu; dims[2] = nx; plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL); break;
    case 15: plhs[0] = mxCreateDoubleMatrix(0, 1, mxREAL); break;
    default: plhs[0] = mxCreateDoubleMatrix(nu, 1, mxREAL); break; /* 16 */
    }
    if (n_out == 0) return;
    out = mxGetPr(plhs[0]);
    if (k < 0 || k > N) mexErrMsgIdAndTxt("MATLAB:dimagree", "k must be in 1..n_hor+1.\n");
    if (ilqgb_eval(g_h, mode, k, x, u, out)) mexErrMsgIdAndTxt("iLQG:gpu", "%s\n", ilqgb_last_error(g_h));
    if ((mode == 1 || mode == 2) && isnan(out[0])) out[0] = INFINITY; /* iLQG_MMex.tem:93, 103 */
}
