/* Plain-C smoke test of the drop-in boundary: links -ltritd, no Python in the loop.
 *   gcc -std=c99 -Iinclude tests/c_abi_smoke.c -Ltriple-tensor-decomposition-with-admm_b200/tritd -ltritd \
 *       -Wl,-rpath,$PWD/triple-tensor-decomposition-with-admm_b200/tritd -lm -o c_abi_smoke
 * With a B200: runs [A,B,C,O,errHist] = triple_decomp_ADMM(D, 3, opts) on the 7x6x5 case through tritd_admm_f64 and
 * compares errHist and A with the outputs of the reference's own .m source (tests/c_abi_smoke_data.h), then
 * triple_product through tritd_triple_product_f64.  Without a CUDA device: checks that the library refuses with
 * TRITD_ERR_CUDA (there is no CPU fallback) and exits 0.  Exit code != 0 on any mismatch. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tritd.h"
#include "c_abi_smoke_data.h"

static int n_lines = 0;
static void count_lines(const char* line, void* user) { (void)user; n_lines += line != NULL; }

static double rel(const double* x, const double* ref, size_t n) {
    double d = 0.0, s = 0.0;
    for (size_t i = 0; i < n; ++i) { d += (x[i] - ref[i]) * (x[i] - ref[i]); s += ref[i] * ref[i]; }
    return sqrt(d) / sqrt(s);
}

int main(void) {
    tritd_ctx* ctx = NULL;
    int64_t t0 = -1, t1 = -1;
    if (tritd_slab_bounds(300, 8, 4, &t0, &t1) != TRITD_OK || t0 != 152 || t1 != 189) { printf("slab_bounds wrong\n"); return 1; }
    printf("%s\n", tritd_version());
    int st = tritd_create(0, &ctx);
    if (st == TRITD_ERR_CUDA) {
        printf("no CUDA device: library refuses (%s) -- ok, no CPU fallback\n", tritd_last_error());
        return strstr(tritd_last_error(), "no CPU fallback") ? 0 : 1;
    }
    if (st != TRITD_OK) { printf("tritd_create failed: %s\n", tritd_last_error()); return 1; }

    enum { n1 = SMOKE_N1, n2 = SMOKE_N2, n3 = SMOKE_N3, r = SMOKE_R, R = SMOKE_R * SMOKE_R, K = SMOKE_ITERS };
    static double A[n1 * R], B[n2 * R], C[n3 * R], O[n1 * n2 * n3], L[n1 * n2 * n3], L2[n1 * n2 * n3], eh[K > 10 ? K : 10];
    tritd_opts o;
    memset(&o, 0, sizeof(o));
    o.mu = SMOKE_MU; o.rho = SMOKE_RHO; o.lambda_ = SMOKE_LAMBDA; o.lambda2 = SMOKE_LAMBDA2; o.tol = 0.0; o.maxIter = K; o.disp = 0;
    int32_t iters = 0;
    tritd_timing tm;
    st = tritd_admm_f64(ctx, smoke_D, n1, n2, n3, r, &o, smoke_A0, smoke_B0, smoke_C0, A, B, C, O, L, eh, &iters, &tm);
    if (st != TRITD_OK) { printf("tritd_admm_f64 failed: %s\n", tritd_last_error()); return 1; }
    const double e1 = rel(eh, smoke_errHist_ref, K), e2 = rel(A, smoke_A_ref, n1 * R);
    printf("iters %d launches %d | rel err vs the reference source: errHist %.2e, A %.2e\n", iters, tm.launches, e1, e2);
    if (iters != K || !(e1 < 1e-8) || !(e2 < 1e-8) || tm.launches <= 0) return 1;

    st = tritd_triple_product_f64(ctx, A, B, C, n1, n2, n3, r, L2);
    if (st != TRITD_OK) { printf("tritd_triple_product_f64 failed: %s\n", tritd_last_error()); return 1; }
    const double e3 = rel(L2, L, n1 * n2 * n3);
    printf("triple_product vs the solver's L output: %.2e\n", e3);
    if (!(e3 < 1e-13)) return 1;

    /* progress lines go through the replaceable sink (what the MEX gateway points at mexPrintf) */
    tritd_set_print(count_lines, NULL);
    o.disp = 1; o.maxIter = 10;
    st = tritd_admm_f64(ctx, smoke_D, n1, n2, n3, r, &o, smoke_A0, smoke_B0, smoke_C0, A, B, C, NULL, NULL, eh, &iters, NULL);
    tritd_set_print(NULL, NULL);
    if (st != TRITD_OK || iters != 10 || n_lines != 1) { printf("print sink: st %d iters %d lines %d\n", st, iters, n_lines); return 1; }

    /* error behaviour: bad arguments are reported, not crashed on */
    if (tritd_admm_f64(ctx, smoke_D, n1, n2, n3, 9, &o, smoke_A0, smoke_B0, smoke_C0, A, B, C, NULL, NULL, eh, &iters, NULL) != TRITD_ERR_UNSUPPORTED) return 1;
    if (tritd_admm_f64(ctx, NULL, n1, n2, n3, r, &o, smoke_A0, smoke_B0, smoke_C0, A, B, C, NULL, NULL, eh, &iters, NULL) != TRITD_ERR_INVALID) return 1;
    tritd_destroy(ctx);
    printf("c_abi_smoke: ok\n");
    return 0;
}
