/* A mock of the part of MATLAB's MEX runtime the gateways under <package>/mex use (TEST INFRASTRUCTURE, not a MEX
 * implementation and not part of the product): mxArray with double / logical / struct classes, mexCallMATLAB("randn"),
 * mexErrMsgIdAndTxt as a long jump back to the caller of mexFunction, mexPrintf into a log, mexLock / mexAtExit records.
 * Compiled together with ONE gateway source into a shared object that tests/test_mex_gateway.py drives through ctypes,
 * so that the gateways are executed -- argument validation, the order and sizes of the randn calls
 * (triple_decomp_ADMM.m:23), marshalling of the outputs, error identifiers -- and not merely syntax-checked.
 * The declarations come from the same stub mex.h the gateways are compiled against. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mex.h"

enum { CLS_DOUBLE = 0, CLS_LOGICAL = 1, CLS_STRUCT = 2, CLS_SINGLE = 3 };
#define MAXF 32
struct mxArray_tag {
    int cls, is_complex, is_sparse;
    mwSize ndim, dims[4];
    void* data;
    int nfields;
    char* names[MAXF];
    mxArray* vals[MAXF];
};

static size_t numel(const mxArray* a) {
    size_t n = 1;
    for (mwSize i = 0; i < a->ndim; ++i) n *= a->dims[i];
    return n;
}
static mxArray* new_array(int cls, mwSize ndim, const mwSize* dims, size_t elsize) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->cls = cls;
    a->ndim = ndim < 2 ? 2 : ndim;
    for (mwSize i = 0; i < 4; ++i) a->dims[i] = 1;
    for (mwSize i = 0; i < ndim && i < 4; ++i) a->dims[i] = dims[i];
    /* like MATLAB: trailing singleton dimensions beyond the second are dropped (an n1 x n2 x 1 array is 2-D) */
    while (a->ndim > 2 && a->dims[a->ndim - 1] == 1) a->ndim--;
    const size_t n = numel(a);
    a->data = elsize ? calloc(n ? n : 1, elsize) : NULL;
    return a;
}

/* ---- the mx / mex API of stub/mex.h ------------------------------------------------------------------------------ */
int mxIsDouble(const mxArray* a) { return a->cls == CLS_DOUBLE; }
int mxIsComplex(const mxArray* a) { return a->is_complex; }
int mxIsSparse(const mxArray* a) { return a->is_sparse; }
int mxIsStruct(const mxArray* a) { return a->cls == CLS_STRUCT; }
int mxIsLogical(const mxArray* a) { return a->cls == CLS_LOGICAL; }
int mxIsEmpty(const mxArray* a) { return numel(a) == 0; }
mxLogical* mxGetLogicals(const mxArray* a) { return a->cls == CLS_LOGICAL ? (mxLogical*)a->data : NULL; }
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->ndim; }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims; }
size_t mxGetNumberOfElements(const mxArray* a) { return numel(a); }
double* mxGetPr(const mxArray* a) { return a->cls == CLS_DOUBLE ? (double*)a->data : NULL; }
double mxGetScalar(const mxArray* a) {
    if (numel(a) == 0) return 0.0;                 /* (MATLAB: undefined; the gateways must not get here) */
    if (a->cls == CLS_DOUBLE) return ((double*)a->data)[0];
    if (a->cls == CLS_LOGICAL) return (double)((mxLogical*)a->data)[0];
    if (a->cls == CLS_SINGLE) return (double)((float*)a->data)[0];
    return 0.0;
}
mxArray* mxGetField(const mxArray* s, size_t idx, const char* name) {
    if (s->cls != CLS_STRUCT || idx != 0) return NULL;
    for (int i = 0; i < s->nfields; ++i)
        if (strcmp(s->names[i], name) == 0) return s->vals[i];
    return NULL;
}
void mxSetField(mxArray* s, size_t idx, const char* name, mxArray* v) {
    if (s->cls != CLS_STRUCT || idx != 0) return;
    for (int i = 0; i < s->nfields; ++i)
        if (strcmp(s->names[i], name) == 0) { s->vals[i] = v; return; }
    if (s->nfields < MAXF) { s->names[s->nfields] = strdup(name); s->vals[s->nfields++] = v; }
}
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nf, const char** names) {
    const mwSize d[2] = {m, n};
    mxArray* s = new_array(CLS_STRUCT, 2, d, 0);
    for (int i = 0; i < nf && i < MAXF; ++i) { s->names[i] = strdup(names[i]); s->vals[i] = NULL; }
    s->nfields = nf < MAXF ? nf : MAXF;
    return s;
}
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity c) {
    (void)cls;
    mxArray* a = new_array(CLS_DOUBLE, ndim, dims, sizeof(double));
    a->is_complex = c == mxCOMPLEX;
    return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    const mwSize d[2] = {m, n};
    return mxCreateNumericArray(2, d, mxDOUBLE_CLASS, c);
}
mxArray* mxCreateDoubleScalar(double v) {
    mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
    ((double*)a->data)[0] = v;
    return a;
}
void mxDestroyArray(mxArray* a) {
    if (!a) return;
    free(a->data);
    for (int i = 0; i < a->nfields; ++i) free(a->names[i]);
    free(a);
}
void* mxMalloc(size_t n) { return malloc(n ? n : 1); }
void mxFree(void* p) { free(p); }

static jmp_buf g_jmp;
static int g_in_call = 0;
static char g_err_id[128], g_err_msg[1024];
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    snprintf(g_err_id, sizeof g_err_id, "%s", id);
    vsnprintf(g_err_msg, sizeof g_err_msg, fmt, ap);
    va_end(ap);
    if (g_in_call) longjmp(g_jmp, 1);
    fprintf(stderr, "mexmock: error outside a call: %s: %s\n", g_err_id, g_err_msg);
    abort();
}

static char g_printed[1 << 16];
static size_t g_printed_len = 0;
int mexPrintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(g_printed + g_printed_len, sizeof g_printed - g_printed_len, fmt, ap);
    va_end(ap);
    if (n > 0) g_printed_len = g_printed_len + (size_t)n < sizeof g_printed ? g_printed_len + (size_t)n : sizeof g_printed - 1;
    return n;
}

static int g_locks = 0;
static void (*g_atexit)(void) = NULL;
void mexLock(void) { g_locks++; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }

/* randn: the numbers come, in call order, from a buffer the test supplied (so the test knows A0, B0, C0) */
static const double* g_rand_src = NULL;
static size_t g_rand_len = 0, g_rand_pos = 0;
static int g_randn_calls = 0;
static double g_randn_dims[16][3];
int mexCallMATLAB(int nlhs, mxArray** plhs, int nrhs, mxArray** prhs, const char* fn) {
    if (strcmp(fn, "randn") != 0 || nlhs != 1 || nrhs != 1 || !mxIsDouble(prhs[0]) || numel(prhs[0]) != 3) return 1;
    const double* d = (const double*)prhs[0]->data;
    const mwSize dims[3] = {(mwSize)d[0], (mwSize)d[1], (mwSize)d[2]};
    if (g_randn_calls < 16) memcpy(g_randn_dims[g_randn_calls], d, 3 * sizeof(double));
    g_randn_calls++;
    mxArray* a = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
    const size_t n = numel(a);
    if (g_rand_pos + n > g_rand_len) return 1;      /* the test did not supply enough numbers */
    memcpy(a->data, g_rand_src + g_rand_pos, n * sizeof(double));
    g_rand_pos += n;
    plhs[0] = a;
    return 0;
}

/* ---- what the test drives ---------------------------------------------------------------------------------------- */
mxArray* mock_double(int ndim, const size_t* dims, const double* data) {
    mwSize d[4] = {1, 1, 1, 1};
    for (int i = 0; i < ndim && i < 4; ++i) d[i] = dims[i];
    mxArray* a = mxCreateNumericArray((mwSize)ndim, d, mxDOUBLE_CLASS, mxREAL);
    if (data && numel(a)) memcpy(a->data, data, numel(a) * sizeof(double));
    return a;
}
mxArray* mock_logical(int ndim, const size_t* dims, const unsigned char* data) {
    mwSize d[4] = {1, 1, 1, 1};
    for (int i = 0; i < ndim && i < 4; ++i) d[i] = dims[i];
    mxArray* a = new_array(CLS_LOGICAL, (mwSize)ndim, d, 1);
    if (data && numel(a)) memcpy(a->data, data, numel(a));
    return a;
}
mxArray* mock_single_scalar(float v) {
    const mwSize d[2] = {1, 1};
    mxArray* a = new_array(CLS_SINGLE, 2, d, sizeof(float));
    ((float*)a->data)[0] = v;
    return a;
}
mxArray* mock_struct(void) { return mxCreateStructMatrix(1, 1, 0, NULL); }
void mock_set_flags(mxArray* a, int is_complex, int is_sparse) { a->is_complex = is_complex; a->is_sparse = is_sparse; }
void mock_set_randn_source(const double* src, size_t n) { g_rand_src = src; g_rand_len = n; g_rand_pos = 0; }
void mock_reset(void) {
    g_err_id[0] = g_err_msg[0] = 0;
    g_printed_len = 0; g_printed[0] = 0;
    g_randn_calls = 0; g_rand_pos = 0;
}
/* returns 0 when mexFunction returned, 1 when it left through mexErrMsgIdAndTxt */
int mock_call(int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs) {
    g_in_call = 1;
    if (setjmp(g_jmp)) { g_in_call = 0; return 1; }
    mexFunction(nlhs, plhs, nrhs, prhs);
    g_in_call = 0;
    return 0;
}
const char* mock_err_id(void) { return g_err_id; }
const char* mock_err_msg(void) { return g_err_msg; }
const char* mock_printed(void) { g_printed[g_printed_len] = 0; return g_printed; }
int mock_randn_calls(void) { return g_randn_calls; }
void mock_randn_dims(int i, double* out3) { memcpy(out3, g_randn_dims[i], 3 * sizeof(double)); }
int mock_locks(void) { return g_locks; }
int mock_run_atexit(void) { if (!g_atexit) return 0; g_atexit(); return 1; }
int mock_ndim(const mxArray* a) { return (int)a->ndim; }
size_t mock_dim(const mxArray* a, int i) { return a->dims[i]; }
