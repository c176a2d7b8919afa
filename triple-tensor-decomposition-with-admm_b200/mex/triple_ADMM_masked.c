/* MEX gateway: [A,B,C,O,E,Out] = triple_ADMM_masked(Y, mask, r, opts)
 *
 * The completion variant the reference's drivers name in a comment but do not ship
 * (traffic_triple_comparison.m:53, video_triple_comparison.m:52:
 *     [A, B, C, O, E, Out] = triple_ADMM_masked(Y, ~mask_missing, r, opts);  errHist = Out.errHist;)
 * mask is logical, size of Y, true = observed.  Semantics: DESIGN.md 4.6 (no reference code exists for it): on the
 * observed entries the statements of triple_decomp_ADMM.m:33-65, unobserved entries carry no constraint and are
 * imputed with the low-rank estimate.  opts as for triple_decomp_ADMM (+ A0, B0, C0, device).
 * This image has neither MATLAB nor Octave: the file is compiled against stub/mex.h and executed against the mock MEX
 * runtime of tests/mex_mock (tests/test_mex_gateway.py). */
#include <string.h>

#include "mex.h"
#include "tritd.h"

static tritd_ctx* g_ctx = NULL;
static int g_dev = 0, g_locked = 0;
static void at_exit(void) { if (g_ctx) { tritd_destroy(g_ctx); g_ctx = NULL; } }
static void to_matlab_console(const char* line, void* user) { (void)user; mexPrintf("%s", line); }

static double req_field(const mxArray* opts, const char* name) {
    const mxArray* f = mxGetField(opts, 0, name);
    if (!f) mexErrMsgIdAndTxt("MATLAB:nonExistentField", "Unrecognized field name \"%s\".", name);
    if (!(mxIsDouble(f) || mxIsLogical(f)) || mxIsComplex(f) || mxIsSparse(f) || mxGetNumberOfElements(f) != 1)
        mexErrMsgIdAndTxt("tritd:opts", "opts.%s must be a real scalar.", name);
    return mxGetScalar(f);
}

static const double* opt_factor(const mxArray* opts, const char* name, size_t numel) {
    const mxArray* f = mxGetField(opts, 0, name);
    if (!f || mxIsEmpty(f)) return NULL;
    if (!mxIsDouble(f) || mxIsComplex(f) || mxGetNumberOfElements(f) != numel)
        mexErrMsgIdAndTxt("tritd:opts", "opts.%s has the wrong size or class.", name);
    return mxGetPr(f);
}

static mxArray* randn3(mwSize a, mwSize b, mwSize c) {
    mxArray* dims = mxCreateDoubleMatrix(1, 3, mxREAL);
    mxArray* out = NULL;
    double* d = mxGetPr(dims);
    d[0] = (double)a; d[1] = (double)b; d[2] = (double)c;
    if (mexCallMATLAB(1, &out, 1, &dims, "randn") != 0) mexErrMsgIdAndTxt("tritd:randn", "randn failed.");
    mxDestroyArray(dims);
    return out;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 4) mexErrMsgIdAndTxt("tritd:nargin", "Usage: [A,B,C,O,E,Out] = triple_ADMM_masked(Y, mask, r, opts)");
    if (nlhs > 6) mexErrMsgIdAndTxt("MATLAB:TooManyOutputs", "Too many output arguments.");
    const mxArray* Ym = prhs[0];
    if (!mxIsDouble(Ym) || mxIsComplex(Ym) || mxIsSparse(Ym)) mexErrMsgIdAndTxt("tritd:Y", "Y must be a full real double array.");
    const mwSize nd = mxGetNumberOfDimensions(Ym);
    const mwSize* dd = mxGetDimensions(Ym);
    if (nd < 2 || nd > 3) mexErrMsgIdAndTxt("tritd:Y", "Y must be n1 x n2 x n3.");
    const mwSize n1 = dd[0], n2 = dd[1], n3 = nd == 3 ? dd[2] : 1;
    if (n1 == 0 || n2 == 0 || n3 == 0) mexErrMsgIdAndTxt("tritd:Y", "Y must not be empty.");
    if (!mxIsLogical(prhs[1]) || mxGetNumberOfElements(prhs[1]) != (size_t)n1 * n2 * n3)
        mexErrMsgIdAndTxt("tritd:mask", "mask must be a logical array of the size of Y (true = observed).");
    if (mxGetNumberOfElements(prhs[2]) != 1) mexErrMsgIdAndTxt("tritd:r", "r must be a scalar.");
    const int r = (int)mxGetScalar(prhs[2]);
    if (r < 1 || (double)r != mxGetScalar(prhs[2])) mexErrMsgIdAndTxt("tritd:r", "r must be a positive integer.");
    if (!mxIsStruct(prhs[3])) mexErrMsgIdAndTxt("tritd:opts", "opts must be a struct.");
    const mxArray* om = prhs[3];

    tritd_opts o;
    o.mu = req_field(om, "mu");
    o.rho = req_field(om, "rho");
    o.lambda_ = req_field(om, "lambda");
    o.lambda2 = req_field(om, "lambda2");
    o.maxIter = (int32_t)req_field(om, "maxIter");
    o.tol = req_field(om, "tol");
    o.disp = req_field(om, "disp") != 0.0;

    const size_t R = (size_t)r * r;
    mxArray *A0m = NULL, *B0m = NULL, *C0m = NULL;
    const double* A0 = opt_factor(om, "A0", n1 * R);
    const double* B0 = opt_factor(om, "B0", n2 * R);
    const double* C0 = opt_factor(om, "C0", n3 * R);
    if (!A0) { A0m = randn3(n1, r, r); A0 = mxGetPr(A0m); }
    if (!B0) { B0m = randn3(r, n2, r); B0 = mxGetPr(B0m); }
    if (!C0) { C0m = randn3(r, r, n3); C0 = mxGetPr(C0m); }

    {
        /* the context is cached across calls; a call that names another opts.device gets a new one */
        const mxArray* dv = mxGetField(om, 0, "device");
        const int dev = (dv && !mxIsEmpty(dv)) ? (int)mxGetScalar(dv) : 0;
        if (g_ctx && dev != g_dev) { tritd_destroy(g_ctx); g_ctx = NULL; }
        if (!g_ctx) {
            if (tritd_create(dev, &g_ctx) != TRITD_OK) mexErrMsgIdAndTxt("tritd:cuda", "%s", tritd_last_error());
            g_dev = dev;
            if (!g_locked) { mexLock(); mexAtExit(at_exit); g_locked = 1; }
        }
    }
    tritd_set_print(to_matlab_console, NULL);

    const mwSize dA[3] = {n1, (mwSize)r, (mwSize)r}, dB[3] = {(mwSize)r, n2, (mwSize)r}, dC[3] = {(mwSize)r, (mwSize)r, n3};
    const mwSize dO[3] = {n1, n2, n3};
    mxArray* Am = mxCreateNumericArray(3, dA, mxDOUBLE_CLASS, mxREAL);
    mxArray* Bm = mxCreateNumericArray(3, dB, mxDOUBLE_CLASS, mxREAL);
    mxArray* Cm = mxCreateNumericArray(3, dC, mxDOUBLE_CLASS, mxREAL);
    mxArray* Om = nlhs >= 4 ? mxCreateNumericArray(3, dO, mxDOUBLE_CLASS, mxREAL) : NULL;
    mxArray* Em = nlhs >= 5 ? mxCreateNumericArray(3, dO, mxDOUBLE_CLASS, mxREAL) : NULL;
    double* eh = (double*)mxMalloc(sizeof(double) * (size_t)(o.maxIter > 0 ? o.maxIter : 1));
    int32_t iters = 0;
    const int st = tritd_admm_ex_f64(g_ctx, mxGetPr(Ym), (const unsigned char*)mxGetLogicals(prhs[1]), (int64_t)n1, (int64_t)n2,
                                     (int64_t)n3, r, &o, A0, B0, C0, mxGetPr(Am), mxGetPr(Bm), mxGetPr(Cm),
                                     Om ? mxGetPr(Om) : NULL, Em ? mxGetPr(Em) : NULL, NULL, eh, &iters, NULL);
    if (A0m) mxDestroyArray(A0m);
    if (B0m) mxDestroyArray(B0m);
    if (C0m) mxDestroyArray(C0m);
    if (st != TRITD_OK) {
        mxFree(eh);
        mexErrMsgIdAndTxt("tritd:solve", "%s", tritd_last_error());
    }
    plhs[0] = Am;
    if (nlhs >= 2) plhs[1] = Bm; else mxDestroyArray(Bm);
    if (nlhs >= 3) plhs[2] = Cm; else mxDestroyArray(Cm);
    if (nlhs >= 4) plhs[3] = Om;
    if (nlhs >= 5) plhs[4] = Em;
    if (nlhs >= 6) {                                /* Out.errHist */
        const char* fields[1] = {"errHist"};
        mxArray* hist = mxCreateDoubleMatrix((mwSize)iters, 1, mxREAL);
        memcpy(mxGetPr(hist), eh, sizeof(double) * (size_t)iters);
        plhs[5] = mxCreateStructMatrix(1, 1, 1, fields);
        mxSetField(plhs[5], 0, "errHist", hist);
    }
    mxFree(eh);
}
