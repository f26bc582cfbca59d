/* fake_mex.c -- TEST INFRASTRUCTURE (oracle/): a small in-process implementation of the mex C API subset declared in
 * mex.h, enough to drive a mexFunction gateway (the reference's iLQG_mex.c or this repo's iLQG_mex_b200.c) from a
 * test.  Built as oracle/_build/libfakemex.so and loaded RTLD_GLOBAL before the gateway library, which leaves the mx and mex
 * symbols undefined exactly as a real mex file does.  mexErrMsgIdAndTxt unwinds to fm_call with longjmp, as MATLAB's does
 * to the interpreter. */
#include "mex.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#define FM_MAX_DIMS 4
#define FM_MAX_FIELDS 64

struct mxArray_tag {
    mxClassID cls;
    int sparse;
    mwSize ndim;
    mwSize dims[FM_MAX_DIMS];
    void *data; /* doubles or int32 */
    int n_fields;
    char *field_name[FM_MAX_FIELDS];
    mxArray *field_val[FM_MAX_FIELDS];
};

static jmp_buf fm_jmp;
static int fm_jmp_armed = 0;
static char fm_err_id[256], fm_err_msg[1024], fm_out[4096];
static size_t fm_out_len = 0;
static long fm_allocs = 0;

static size_t numel(const mxArray *a)
{
    size_t n = 1;
    for (mwSize i = 0; i < a->ndim; i++) n *= a->dims[i];
    return n;
}

static mxArray *new_array(mwSize ndim, const mwSize *dims, mxClassID cls)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->cls = cls;
    a->ndim = ndim < 2 ? 2 : ndim;
    for (mwSize i = 0; i < a->ndim; i++) a->dims[i] = i < ndim ? dims[i] : 1;
    size_t el = cls == mxINT32_CLASS ? 4 : 8, n = numel(a);
    a->data = cls == mxSTRUCT_CLASS ? NULL : calloc(n ? n : 1, el);
    return a;
}

size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a)
{
    size_t n = 1;
    for (mwSize i = 1; i < a->ndim; i++) n *= a->dims[i];
    return n;
}
size_t mxGetNumberOfElements(const mxArray *a) { return numel(a); }
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->ndim; }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims; }
double *mxGetPr(const mxArray *a) { return a->cls == mxDOUBLE_CLASS ? (double *)a->data : NULL; }
void *mxGetData(const mxArray *a) { return a->data; }
int mxIsStruct(const mxArray *a) { return a->cls == mxSTRUCT_CLASS; }
int mxIsDouble(const mxArray *a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsSparse(const mxArray *a) { return a->sparse; }
int mxGetNumberOfFields(const mxArray *a) { return a->cls == mxSTRUCT_CLASS ? a->n_fields : 0; }
mxArray *mxGetFieldByNumber(const mxArray *a, mwIndex index, int field)
{
    if (a->cls != mxSTRUCT_CLASS || index != 0 || field < 0 || field >= a->n_fields) return NULL;
    return a->field_val[field];
}
const char *mxGetFieldNameByNumber(const mxArray *a, int field)
{
    if (a->cls != mxSTRUCT_CLASS || field < 0 || field >= a->n_fields) return NULL;
    return a->field_name[field];
}
mxArray *mxGetField(const mxArray *a, mwIndex index, const char *name)
{
    if (a->cls != mxSTRUCT_CLASS || index != 0) return NULL;
    for (int i = 0; i < a->n_fields; i++)
        if (!strcmp(a->field_name[i], name)) return a->field_val[i];
    return NULL;
}

mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c)
{
    (void)c;
    mwSize d[2] = {m, n};
    return new_array(2, d, mxDOUBLE_CLASS);
}
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c)
{
    (void)c;
    return new_array(ndim, dims, cls);
}
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c)
{
    (void)c;
    mwSize d[2] = {m, n};
    return new_array(2, d, cls);
}
void mxDestroyArray(mxArray *a)
{
    if (!a) return;
    for (int i = 0; i < a->n_fields; i++) {
        free(a->field_name[i]);
        mxDestroyArray(a->field_val[i]);
    }
    free(a->data);
    free(a);
}
void *mxMalloc(size_t n)
{
    fm_allocs++;
    return malloc(n ? n : 1);
}
void *mxCalloc(size_t n, size_t size)
{
    fm_allocs++;
    return calloc(n ? n : 1, size ? size : 1);
}
void mxFree(void *p)
{
    if (p) fm_allocs--;
    free(p);
}

int mexPrintf(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    size_t len = strlen(buf);
    if (fm_out_len + len >= sizeof fm_out) { /* keep the tail */
        size_t keep = sizeof fm_out / 2;
        if (fm_out_len > keep) {
            memmove(fm_out, fm_out + fm_out_len - keep, keep);
            fm_out_len = keep;
        }
    }
    if (fm_out_len + len < sizeof fm_out) {
        memcpy(fm_out + fm_out_len, buf, len);
        fm_out_len += len;
    }
    fm_out[fm_out_len] = 0;
    return n;
}

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(fm_err_msg, sizeof fm_err_msg, fmt, ap);
    va_end(ap);
    snprintf(fm_err_id, sizeof fm_err_id, "%s", id ? id : "");
    if (fm_jmp_armed) longjmp(fm_jmp, 1);
    fprintf(stderr, "mexErrMsgIdAndTxt outside fm_call: %s: %s\n", fm_err_id, fm_err_msg);
    abort();
}

void mexErrMsgTxt(const char *msg) { mexErrMsgIdAndTxt("", "%s", msg); }

double mxGetScalar(const mxArray *a) { return (a && a->cls == mxDOUBLE_CLASS && numel(a) > 0) ? ((const double *)a->data)[0] : 0.0; }

void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    (void)id;
    (void)fmt;
}

/* ---- test-side helpers ---- */
mxArray *fm_new_double(int ndim, const size_t *dims, const double *data)
{
    mxArray *a = new_array((mwSize)ndim, dims, mxDOUBLE_CLASS);
    if (data) memcpy(a->data, data, numel(a) * sizeof(double));
    return a;
}
mxArray *fm_new_struct(void)
{
    mwSize d[2] = {1, 1};
    return new_array(2, d, mxSTRUCT_CLASS);
}
void fm_set_field(mxArray *s, const char *name, mxArray *value)
{
    if (s->n_fields >= FM_MAX_FIELDS) abort();
    s->field_name[s->n_fields] = strdup(name);
    s->field_val[s->n_fields] = value;
    s->n_fields++;
}
void fm_mark_sparse(mxArray *a, int sparse) { a->sparse = sparse; }

int fm_call(fm_mexfunction fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs)
{
    fm_err_id[0] = fm_err_msg[0] = 0;
    fm_out_len = 0;
    fm_out[0] = 0;
    fm_allocs = 0;
    if (setjmp(fm_jmp)) {
        fm_jmp_armed = 0;
        return 1;
    }
    fm_jmp_armed = 1;
    fn(nlhs, plhs, nrhs, prhs);
    fm_jmp_armed = 0;
    return 0;
}
const char *fm_error_id(void) { return fm_err_id; }
const char *fm_error_msg(void) { return fm_err_msg; }
const char *fm_printed(void) { return fm_out; }
long fm_live_allocs(void) { return fm_allocs; }
