/* Stand-in for MATLAB/Octave's mex.h.  TEST INFRASTRUCTURE (oracle/).
 *
 * Two uses:
 *  1. the reference solver core includes "mex.h" only for mxIsNaN/mxIsInf/mxGetInf (iLQG_problem.tem:6,11-12;
 *     iLQG.c:16-19; back_pass.c:15-18) -- the three macros below are all it needs;
 *  2. gateway tests: the subset of the mex C API that the reference's gateway (iLQG_mex.c:19-144) and this repo's
 *     gateway (ddp-generator_b200/mex/iLQG_mex_b200.c) call, implemented by fake_mex.c (libfakemex.so) so that a
 *     mexFunction can be driven from a test without MATLAB or Octave.
 * Compile with -DHAVE_OCTAVE so matrix.h is not requested. */
#ifndef ORACLE_MEX_STUB_H
#define ORACLE_MEX_STUB_H
#include <math.h>
#include <stddef.h>
#include <stdio.h>

#define mxIsNaN(v) isnan(v)
#define mxIsInf(v) isinf(v)
#define mxGetInf() ((double)INFINITY)

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxSTRUCT_CLASS = 2, mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12 } mxClassID;

/* queries */
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a); /* product of all dimensions but the first, as in MATLAB */
size_t mxGetNumberOfElements(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
int mxIsStruct(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsSparse(const mxArray *a);
int mxGetNumberOfFields(const mxArray *a);
mxArray *mxGetFieldByNumber(const mxArray *a, mwIndex index, int field);
const char *mxGetFieldNameByNumber(const mxArray *a, int field);
mxArray *mxGetField(const mxArray *a, mwIndex index, const char *name);
double mxGetScalar(const mxArray *a);

/* creation / memory */
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
void mxDestroyArray(mxArray *a);
void *mxMalloc(size_t n);
void *mxCalloc(size_t n, size_t size);
void mxFree(void *p);

/* host interaction */
int mexPrintf(const char *fmt, ...);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...); /* does not return (longjmp into fm_call) */
void mexErrMsgTxt(const char *msg);                           /* the same without an identifier */
void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...);

/* test-side helpers (not part of the mex API) */
typedef void (*fm_mexfunction)(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
mxArray *fm_new_double(int ndim, const size_t *dims, const double *data); /* copies data (NULL = zeros) */
mxArray *fm_new_struct(void);                                             /* 1x1 struct without fields */
void fm_set_field(mxArray *s, const char *name, mxArray *value);          /* the struct takes ownership */
void fm_mark_sparse(mxArray *a, int sparse);
int fm_call(fm_mexfunction fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs); /* 0 ok, 1 = mexErrMsg... */
const char *fm_error_id(void);
const char *fm_error_msg(void);
const char *fm_printed(void); /* tail of what the last call printed through mexPrintf */
long fm_live_allocs(void);    /* mxMalloc'ed blocks not yet freed */

#ifdef __cplusplus
}
#endif
#endif
