/* A test double of the few libtritd entry points the MEX gateways call (TEST INFRASTRUCTURE; never shipped, never a
 * fallback: the product library has no CPU path).  It lets tests/test_mex_gateway.py execute the gateways' marshalling
 * AFTER the solve on a machine without a GPU: outputs are simple, recognisable functions of the inputs, the arguments the
 * gateway passed are recorded.  The GPU variant of the test links the same gateway + mock against the real libtritd. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tritd.h"

struct tritd_ctx { int ndev; int devs[8]; };
static char g_err[256] = "";
static tritd_print_fn g_print = NULL;
static void* g_print_user = NULL;

/* what the last calls saw (read back by the test) */
int fake_created = 0, fake_destroyed = 0, fake_ndev = 0, fake_devs[8];
int fake_fail_solve = 0;          /* != 0: the next solve returns this status */
int fake_fail_create = 0;
int fake_iters = 7;               /* iterations "executed" (capped by maxIter) */
int fake_mask_seen = 0; long fake_mask_sum = 0;
int fake_have_O = 0, fake_have_E = 0, fake_have_L = 0;
tritd_opts fake_opts;
long long fake_shape[4];

int tritd_create_devices(const int* devices, int ndev, tritd_ctx** out) {
    if (fake_fail_create) { snprintf(g_err, sizeof g_err, "no CUDA device (fake)"); return TRITD_ERR_CUDA; }
    tritd_ctx* c = (tritd_ctx*)calloc(1, sizeof *c);
    c->ndev = ndev;
    fake_ndev = ndev;
    for (int i = 0; i < ndev && i < 8; ++i) c->devs[i] = fake_devs[i] = devices[i];
    fake_created++;
    *out = c;
    return TRITD_OK;
}
int tritd_create(int device, tritd_ctx** out) { return tritd_create_devices(&device, 1, out); }
void tritd_destroy(tritd_ctx* c) { if (c) { free(c); fake_destroyed++; } }
const char* tritd_last_error(void) { return g_err; }
void tritd_set_print(tritd_print_fn fn, void* user) { g_print = fn; g_print_user = user; }

static int solve_common(const double* D, const unsigned char* mask, int64_t n1, int64_t n2, int64_t n3, int r, int32_t maxIter,
                        int32_t disp, const double* A0, const double* B0, const double* C0, double* A, double* B, double* C,
                        double* O, double* E, double* L, double* errHist, int32_t* iters_out, const char* line_fmt, int every) {
    if (fake_fail_solve) { snprintf(g_err, sizeof g_err, "ridge system contains NaN / Inf (fake)"); return fake_fail_solve; }
    const size_t R = (size_t)r * r, N = (size_t)n1 * n2 * n3;
    fake_shape[0] = n1; fake_shape[1] = n2; fake_shape[2] = n3; fake_shape[3] = r;
    for (size_t i = 0; i < (size_t)n1 * R; ++i) A[i] = 2.0 * A0[i];
    for (size_t i = 0; i < (size_t)n2 * R; ++i) B[i] = 3.0 * B0[i];
    for (size_t i = 0; i < (size_t)n3 * R; ++i) C[i] = 4.0 * C0[i];
    fake_have_O = O != NULL; fake_have_E = E != NULL; fake_have_L = L != NULL;
    for (size_t i = 0; i < N; ++i) {
        if (O) O[i] = D[i] + 1.0;
        if (E) E[i] = D[i] + 2.0;
        if (L) L[i] = D[i] + 3.0;
    }
    fake_mask_seen = mask != NULL; fake_mask_sum = 0;
    if (mask) for (size_t i = 0; i < N; ++i) fake_mask_sum += mask[i] != 0;
    const int k = fake_iters < maxIter ? fake_iters : maxIter;
    for (int i = 0; i < k; ++i) {
        errHist[i] = 1.0 / (i + 1);
        if (disp && g_print && (i + 1) % every == 0) {
            char line[128];
            snprintf(line, sizeof line, line_fmt, i + 1, errHist[i], errHist[i]);
            g_print(line, g_print_user);
        }
    }
    *iters_out = k;
    return TRITD_OK;
}

int tritd_admm_ex_f64(tritd_ctx* ctx, const double* D, const unsigned char* mask, int64_t n1, int64_t n2, int64_t n3, int r,
                      const tritd_opts* o, const double* A0, const double* B0, const double* C0, double* A, double* B, double* C,
                      double* O, double* E, double* L, double* errHist, int32_t* iters_out, tritd_timing* tm) {
    (void)ctx; (void)tm;
    fake_opts = *o;
    return solve_common(D, mask, n1, n2, n3, r, o->maxIter, o->disp, A0, B0, C0, A, B, C, O, E, L, errHist, iters_out,
                        "Iter %d, errL=%.2e, errO=%.2e\n", 10);
}
int tritd_als_f64(tritd_ctx* ctx, const double* X, int64_t n1, int64_t n2, int64_t n3, int r, int32_t maxIter, double tol,
                  int32_t disp, const double* A0, const double* B0, const double* C0, double* A, double* B, double* C,
                  double* errHist, int32_t* iters_out) {
    (void)ctx;
    memset(&fake_opts, 0, sizeof fake_opts);
    fake_opts.maxIter = maxIter; fake_opts.tol = tol; fake_opts.disp = disp;
    return solve_common(X, NULL, n1, n2, n3, r, maxIter, disp, A0, B0, C0, A, B, C, NULL, NULL, NULL, errHist, iters_out,
                        "Iteration %d, relative error = %.4e\n", 5);
}
int tritd_triple_product_f64(tritd_ctx* ctx, const double* A, const double* B, const double* C, int64_t n1, int64_t n2,
                             int64_t n3, int r, double* Xhat) {
    (void)ctx;
    if (fake_fail_solve) { snprintf(g_err, sizeof g_err, "failed (fake)"); return fake_fail_solve; }
    fake_shape[0] = n1; fake_shape[1] = n2; fake_shape[2] = n3; fake_shape[3] = r;
    /* the definition itself: Xhat(i,j,t) = sum_{p,s} A(i,p,s) B(p,j,s) C(p,s,t)  (triple_product.m:6 with buildF.m:17-21); small sizes only */
    for (int64_t t = 0; t < n3; ++t)
        for (int64_t j = 0; j < n2; ++j)
            for (int64_t i = 0; i < n1; ++i) {
                double v = 0.0;
                for (int s = 0; s < r; ++s)
                    for (int p = 0; p < r; ++p)
                        v += A[i + n1 * (p + (int64_t)r * s)] * B[p + r * (j + n2 * s)] * C[p + r * (s + (int64_t)r * t)];
                Xhat[i + n1 * (j + n2 * t)] = v;
            }
    return TRITD_OK;
}
