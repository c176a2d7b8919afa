/* MEX gateway: [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts)
 *
 * Drop-in for fast_robust_triple_tensor/triple_decomp_ADMM.m:1 -- same name, same three
 * inputs, same five outputs; a triple_decomp_ADMM.mex* earlier on the MATLAB path shadows
 * the .m file.  All numerical work happens in libtritd (CUDA, sm_100a); this file only
 * validates arguments, draws the initial factors with MATLAB's own randn in the reference's
 * order (:23 -- A, B, C -- so rng(0) in the caller gives the reference's factors) and moves
 * mxArrays in and out.  Optional extension fields of opts: A0, B0, C0 (injected
 * initial factors), device (CUDA ordinal) / devices (list of ordinals) / ngpu (use devices 0..ngpu-1: one MATLAB
 * process drives several GPUs, D sharded along mode 3), mask (logical n1 x n2 x n3, true = observed: the completion
 * variant, see triple_ADMM_masked.c).  A sixth output, when requested, is L = triple_product(A,B,C), a seventh is
 * E ("O,E : sparse components (clone E)", triple_decomp_ADMM.m:12).  The progress line of opts.disp (:60-62) is
 * printed through mexPrintf.
 *
 * Build (on a machine with MATLAB + CUDA):
 *   mex -I../../include triple_decomp_ADMM.c -L../tritd -ltritd
 * This image has neither MATLAB nor Octave: the file is compiled against stub/mex.h and executed against the mock MEX
 * runtime of tests/mex_mock (tests/test_mex_gateway.py, tests/test_zz_mex_gateway_gpu.py).
 */
#include <string.h>

#include "mex.h"
#include "tritd.h"

static tritd_ctx* g_ctx = NULL;
static int g_devs[8], g_ndev = 0, g_locked = 0;   /* the devices of the cached context; mexLock taken once */

static void at_exit(void) {
    if (g_ctx) { tritd_destroy(g_ctx); g_ctx = NULL; }
}

static void to_matlab_console(const char* line, void* user) {
    (void)user;
    mexPrintf("%s", line);
}

/* a required scalar field of opts; logical is accepted (opts.disp = true) */
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
    if (nrhs != 3) mexErrMsgIdAndTxt("tritd:nargin", "Usage: [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts)");
    if (nlhs > 7) mexErrMsgIdAndTxt("MATLAB:TooManyOutputs", "Too many output arguments.");
    const mxArray* Dm = prhs[0];
    if (!mxIsDouble(Dm) || mxIsComplex(Dm) || mxIsSparse(Dm)) mexErrMsgIdAndTxt("tritd:D", "D must be a full real double array.");
    const mwSize nd = mxGetNumberOfDimensions(Dm);
    const mwSize* dd = mxGetDimensions(Dm);
    if (nd < 2 || nd > 3) mexErrMsgIdAndTxt("tritd:D", "D must be n1 x n2 x n3.");
    const mwSize n1 = dd[0], n2 = dd[1], n3 = nd == 3 ? dd[2] : 1;   /* n3 = 1 arrives as 2-D */
    if (n1 == 0 || n2 == 0 || n3 == 0) mexErrMsgIdAndTxt("tritd:D", "D must not be empty.");
    if (mxGetNumberOfElements(prhs[1]) != 1) mexErrMsgIdAndTxt("tritd:r", "r must be a scalar.");
    const int r = (int)mxGetScalar(prhs[1]);
    if (r < 1 || (double)r != mxGetScalar(prhs[1])) mexErrMsgIdAndTxt("tritd:r", "r must be a positive integer.");
    if (!mxIsStruct(prhs[2])) mexErrMsgIdAndTxt("tritd:opts", "opts must be a struct.");
    const mxArray* om = prhs[2];

    tritd_opts o;   /* the seven fields the reference reads at :16-20; anything else is ignored */
    o.mu = req_field(om, "mu");
    o.rho = req_field(om, "rho");
    o.lambda_ = req_field(om, "lambda");
    o.lambda2 = req_field(om, "lambda2");
    o.maxIter = (int32_t)req_field(om, "maxIter");
    o.tol = req_field(om, "tol");
    o.disp = req_field(om, "disp") != 0.0;

    /* initial factors: randn(n1,r,r), randn(r,n2,r), randn(r,r,n3) in this order (:23) */
    const size_t R = (size_t)r * r;
    mxArray *A0m = NULL, *B0m = NULL, *C0m = NULL;
    const double* A0 = opt_factor(om, "A0", n1 * R);
    const double* B0 = opt_factor(om, "B0", n2 * R);
    const double* C0 = opt_factor(om, "C0", n3 * R);
    if (!A0) { A0m = randn3(n1, r, r); A0 = mxGetPr(A0m); }
    if (!B0) { B0m = randn3(r, n2, r); B0 = mxGetPr(B0m); }
    if (!C0) { C0m = randn3(r, r, n3); C0 = mxGetPr(C0m); }

    {
        /* opts.devices = [0 1 2 3] (CUDA ordinals) or opts.ngpu = 4 (devices 0..3) or opts.device = 2; default: device 0.
         * Several devices: one MATLAB process drives them all (tritd_create_devices), D is sharded along mode 3.
         * The context is cached across calls; a call that asks for OTHER devices than the cached context has
         * (opts.ngpu = 8 after a single-GPU call) gets a new one. */
        int devs[8], ndev = 0;
        const mxArray* dl = mxGetField(om, 0, "devices");
        const mxArray* ng = mxGetField(om, 0, "ngpu");
        const mxArray* dv = mxGetField(om, 0, "device");
        if (dl && !mxIsEmpty(dl)) {
            if (!mxIsDouble(dl) || mxGetNumberOfElements(dl) > 8) mexErrMsgIdAndTxt("tritd:opts", "opts.devices must list at most 8 device ordinals.");
            for (ndev = 0; ndev < (int)mxGetNumberOfElements(dl); ++ndev) devs[ndev] = (int)mxGetPr(dl)[ndev];
        } else if (ng && !mxIsEmpty(ng)) {
            const int n = (int)mxGetScalar(ng);
            if (n < 1 || n > 8) mexErrMsgIdAndTxt("tritd:opts", "opts.ngpu must be 1..8.");
            for (ndev = 0; ndev < n; ++ndev) devs[ndev] = ndev;
        } else {
            devs[0] = (dv && !mxIsEmpty(dv)) ? (int)mxGetScalar(dv) : 0;
            ndev = 1;
        }
        if (g_ctx && (ndev != g_ndev || memcmp(devs, g_devs, sizeof(int) * (size_t)ndev) != 0)) {
            tritd_destroy(g_ctx);
            g_ctx = NULL;
        }
        if (!g_ctx) {
            if (tritd_create_devices(devs, ndev, &g_ctx) != TRITD_OK)
                mexErrMsgIdAndTxt("tritd:cuda", "%s", tritd_last_error());
            memcpy(g_devs, devs, sizeof(int) * (size_t)ndev);
            g_ndev = ndev;
            if (!g_locked) { mexLock(); mexAtExit(at_exit); g_locked = 1; }
        }
    }
    tritd_set_print(to_matlab_console, NULL);

    /* optional completion mask: logical, size of D, true = observed */
    const unsigned char* mask = NULL;
    {
        const mxArray* mk = mxGetField(om, 0, "mask");
        if (mk && !mxIsEmpty(mk)) {
            if (!mxIsLogical(mk) || mxGetNumberOfElements(mk) != (size_t)n1 * n2 * n3)
                mexErrMsgIdAndTxt("tritd:opts", "opts.mask must be a logical array of the size of D.");
            mask = (const unsigned char*)mxGetLogicals(mk);
        }
    }

    const mwSize dA[3] = {n1, (mwSize)r, (mwSize)r}, dB[3] = {(mwSize)r, n2, (mwSize)r}, dC[3] = {(mwSize)r, (mwSize)r, n3};
    const mwSize dO[3] = {n1, n2, n3};
    mxArray* Am = mxCreateNumericArray(3, dA, mxDOUBLE_CLASS, mxREAL);
    mxArray* Bm = mxCreateNumericArray(3, dB, mxDOUBLE_CLASS, mxREAL);
    mxArray* Cm = mxCreateNumericArray(3, dC, mxDOUBLE_CLASS, mxREAL);
    mxArray* Om = nlhs >= 4 ? mxCreateNumericArray(3, dO, mxDOUBLE_CLASS, mxREAL) : NULL;
    mxArray* Lm = nlhs >= 6 ? mxCreateNumericArray(3, dO, mxDOUBLE_CLASS, mxREAL) : NULL;
    mxArray* Em = nlhs >= 7 ? mxCreateNumericArray(3, dO, mxDOUBLE_CLASS, mxREAL) : NULL;
    double* eh = (double*)mxMalloc(sizeof(double) * (size_t)(o.maxIter > 0 ? o.maxIter : 1));
    int32_t iters = 0;

    /* inputs are MATLAB-owned and only read; the library emits the reference's progress line
     * ("Iter %d, errL=%.2e, errO=%.2e", :60-62) through the print sink when opts.disp is set */
    const int st = tritd_admm_ex_f64(g_ctx, mxGetPr(Dm), mask, (int64_t)n1, (int64_t)n2, (int64_t)n3, r, &o, A0, B0, C0,
                                     mxGetPr(Am), mxGetPr(Bm), mxGetPr(Cm), Om ? mxGetPr(Om) : NULL,
                                     Em ? mxGetPr(Em) : NULL, Lm ? mxGetPr(Lm) : NULL, eh, &iters, NULL);
    if (A0m) mxDestroyArray(A0m);
    if (B0m) mxDestroyArray(B0m);
    if (C0m) mxDestroyArray(C0m);
    if (st != TRITD_OK) {
        mxFree(eh);
        mexErrMsgIdAndTxt("tritd:solve", "%s", tritd_last_error());   /* long-jumps; mxArrays are reclaimed by MATLAB */
    }

    plhs[0] = Am;                                   /* nlhs == 0 still returns ans = A */
    if (nlhs >= 2) plhs[1] = Bm; else mxDestroyArray(Bm);
    if (nlhs >= 3) plhs[2] = Cm; else mxDestroyArray(Cm);
    if (nlhs >= 4) plhs[3] = Om;
    if (nlhs >= 5) {                                /* errHist = errHist(1:k)  (:68) */
        plhs[4] = mxCreateDoubleMatrix((mwSize)iters, 1, mxREAL);
        memcpy(mxGetPr(plhs[4]), eh, sizeof(double) * (size_t)iters);
    }
    if (nlhs >= 6) plhs[5] = Lm;
    if (nlhs >= 7) plhs[6] = Em;
    mxFree(eh);
}
