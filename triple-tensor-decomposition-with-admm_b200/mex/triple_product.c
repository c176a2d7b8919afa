/* MEX gateway: Xhat = triple_product(A, B, C)
 * Drop-in for fast_robust_triple_tensor/triple_product.m:1-8, which the reference's callers run
 * right after the solve (traffic_triple_comparison.m:62).  Compiled against stub/mex.h, executed against the mock MEX
 * runtime of tests/mex_mock (tests/test_mex_gateway.py, tests/test_zz_mex_gateway_gpu.py). */
#include "mex.h"
#include "tritd.h"

static tritd_ctx* g_ctx = NULL;
static void at_exit(void) { if (g_ctx) { tritd_destroy(g_ctx); g_ctx = NULL; } }

static void dims3(const mxArray* a, mwSize d[3]) {
    const mwSize nd = mxGetNumberOfDimensions(a);
    const mwSize* dd = mxGetDimensions(a);
    if (!mxIsDouble(a) || mxIsComplex(a) || mxIsSparse(a) || nd > 3)
        mexErrMsgIdAndTxt("tritd:arg", "A, B, C must be full real double arrays with at most 3 dimensions.");
    d[0] = dd[0]; d[1] = nd >= 2 ? dd[1] : 1; d[2] = nd >= 3 ? dd[2] : 1;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    mwSize a[3], b[3], c[3];
    if (nrhs != 3 || nlhs > 1) mexErrMsgIdAndTxt("tritd:nargin", "Usage: Xhat = triple_product(A, B, C)");
    dims3(prhs[0], a); dims3(prhs[1], b); dims3(prhs[2], c);
    const mwSize r = a[1];
    if (a[2] != r || b[0] != r || b[2] != r || c[0] != r || c[1] != r)
        mexErrMsgIdAndTxt("tritd:arg", "Expected A: n1 x r x r, B: r x n2 x r, C: r x r x n3.");
    if (!g_ctx) {
        if (tritd_create(0, &g_ctx) != TRITD_OK) mexErrMsgIdAndTxt("tritd:cuda", "%s", tritd_last_error());
        mexLock();
        mexAtExit(at_exit);
    }
    const mwSize dX[3] = {a[0], b[1], c[2]};
    plhs[0] = mxCreateNumericArray(3, dX, mxDOUBLE_CLASS, mxREAL);
    if (tritd_triple_product_f64(g_ctx, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), (int64_t)a[0],
                                 (int64_t)b[1], (int64_t)c[2], (int)r, mxGetPr(plhs[0])) != TRITD_OK)
        mexErrMsgIdAndTxt("tritd:solve", "%s", tritd_last_error());
}
