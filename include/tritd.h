/* libtritd -- C ABI of the B200-native TriTD-ADMM hot path.
 *
 * This is the drop-in boundary for ONE path of the reference
 * (dangnq2501/Triple-Tensor-Decomposition-with-ADMM):
 *
 *   [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts)
 *       fast_robust_triple_tensor/triple_decomp_ADMM.m:1-70
 *
 * plus the L2 helpers that path calls and its callers use afterwards
 * (triple_product.m:1-8, unfold.m:1-14, buildF.m/buildG.m/buildH.m,
 * soft_threshold.m:1-3).  The reference has no FFI of its own (it is pure
 * MATLAB); the entry points below are what a MEX gateway named
 * triple_decomp_ADMM.mex* binds -- see INTEGRATION.md and
 * triple-tensor-decomposition-with-admm_b200/mex/.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns a tritd_status
 *    (0 = OK) and never throws; tritd_last_error() gives the message of the
 *    last failure on the calling thread.
 *  - all matrices/tensors are IEEE-754 float64 in MATLAB (column-major)
 *    layout: D, O, L are n1 x n2 x n3; A is n1 x r x r, B is r x n2 x r,
 *    C is r x r x n3; errHist has room for opts->maxIter doubles.
 *  - "_host" pointers are host memory (pageable or pinned), "_dev" pointers
 *    are device memory on the context's GPU.
 *  - there is no CPU fallback: without a CUDA device every compute entry
 *    point fails with TRITD_ERR_CUDA.
 *  - a context is bound to one calling thread at a time; calls on it are
 *    serialised by the caller.
 *
 * Multi-GPU: the tensor is sharded along its slowest mode (t, mode 3) into
 * contiguous slabs, one per rank (tritd_slab_bounds).  Either one process drives all GPUs
 * (tritd_create_devices: full tensors in and out, the MEX gateway's mode) or: a context created
 * with tritd_create_rank() is one rank of an NCCL communicator (one process
 * per GPU; the 128-byte unique id is exchanged by the caller, e.g. through
 * torch.distributed); every rank then passes only ITS slab of D / C0 and
 * receives its slab of O / C, while A and B are replicated (bitwise equal on all ranks).
 * With 2..8 ranks on one node the per-iteration partials are exchanged through NVLink peer
 * mailboxes mapped with CUDA IPC (NCCL is then used for set-up only); TRITD_XCHG_NCCL=1 in the
 * environment selects NCCL all-reduces instead.  tritd_problem_create(), tritd_problem_init() and
 * the solver calls are collective: every rank must make them in the same order.
 */
#ifndef TRITD_H
#define TRITD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tritd_ctx tritd_ctx;
typedef struct tritd_problem tritd_problem;

typedef enum {
    TRITD_OK = 0,
    TRITD_ERR_INVALID = 1,      /* bad argument (NULL, non-positive size, r out of range, ...) */
    TRITD_ERR_CUDA = 2,         /* CUDA runtime/driver failure or no device */
    TRITD_ERR_NCCL = 3,         /* NCCL missing or a collective failed */
    TRITD_ERR_NUMERIC = 4,      /* a ridge system contains NaN / Inf (MATLAB's pinv raises an error there too) */
    TRITD_ERR_UNSUPPORTED = 5,  /* r > TRITD_MAX_R, or an operation the context kind does not offer */
    TRITD_ERR_TIMEOUT = 6       /* a bounded device-side wait gave up (a peer rank died / broken exchange); results invalid */
} tritd_status;

#define TRITD_MAX_R 8           /* triple rank r (R = r^2 <= 64 columns per factor) */
#define TRITD_NCCL_ID_BYTES 128

/* opts struct of triple_decomp_ADMM.m:16-20 (field names as in the reference; `lambda_`
 * because lambda is reserved in some bindings).  mu_max = 1e6*mu and the fixed 1e-9 ridge of
 * update_C (:93) are part of the algorithm, not options. */
typedef struct {
    double mu;        /* opts.mu      initial penalty (muL = muO)          */
    double rho;       /* opts.rho     penalty growth factor                */
    double lambda_;   /* opts.lambda  weight of ||E||_1                    */
    double lambda2;   /* opts.lambda2 ridge of the A and B updates         */
    double tol;       /* opts.tol     relative-change stopping tolerance   */
    int32_t maxIter;  /* opts.maxIter                                      */
    int32_t disp;     /* opts.disp    print "Iter %d, errL=%.2e, errO=%.2e" every 10th iteration */
} tritd_opts;

/* Wall-clock / device timing of one tritd_admm_f64 call, milliseconds. */
typedef struct {
    double h2d_ms;        /* host -> device copies of D and the initial factors */
    double iterate_ms;    /* the ADMM loop, CUDA-event time on the compute stream */
    double d2h_ms;        /* device -> host copies of the outputs */
    double total_ms;      /* entry to return, host clock */
    int32_t iters;        /* iterations executed (== length of errHist) */
    int32_t launches;     /* kernels this library launched during the call */
} tritd_timing;

/* ---- contexts ---------------------------------------------------------- */

/* Single-rank context on CUDA device `device`. */
int tritd_create(int device, tritd_ctx** out);

/* One rank of an `nranks`-way mode-3 sharded solve (one process per GPU).
 * `nccl_id` is the TRITD_NCCL_ID_BYTES blob produced by tritd_nccl_unique_id()
 * on rank 0 and distributed by the caller. */
int tritd_nccl_unique_id(void* id_out);
int tritd_create_rank(int device, int rank, int nranks, const void* nccl_id, tritd_ctx** out);

/* Single-PROCESS multi-GPU context over `ndev` (1..8) distinct devices with mutual peer access -- what a MEX gateway
 * needs, MATLAB being one process (traffic_triple_comparison.m:55).  tritd_admm_f64 / tritd_admm_ex_f64 on such a
 * context take and return FULL tensors: device g owns the mode-3 slab tritd_slab_bounds(n3, ndev, g) (n3 >= ndev), the
 * per-iteration partials travel through the same NVLink peer mailboxes as with tritd_create_rank (mapped by peer
 * access: no IPC, no NCCL), A and B are taken from device 0.  The staged tritd_problem_* solver API needs a
 * single-device context; the standalone helpers run on the first device. */
int tritd_create_devices(const int* devices, int ndev, tritd_ctx** out);

void tritd_destroy(tritd_ctx* ctx);
const char* tritd_last_error(void);
const char* tritd_version(void);

/* Sink of the progress lines ("Iter %d, errL=%.2e, errO=%.2e\n" every 10th iteration when opts.disp,
 * triple_decomp_ADMM.m:60-62; "Iteration %d, relative error = %.4e\n" of triple_decomp_ALS.m:17-19).  Default: stdout.
 * A MEX gateway points it at mexPrintf (the MATLAB desktop does not show the process's stdout); NULL restores the
 * default.  Process-wide; called on the thread that runs the solver. */
typedef void (*tritd_print_fn)(const char* line, void* user);
void tritd_set_print(tritd_print_fn fn, void* user);

/* Use an existing CUDA stream (a cudaStream_t passed as void*) for all work of this
 * context, so the caller can bracket it with its own events; NULL restores the
 * context's private stream. */
int tritd_set_stream(tritd_ctx* ctx, void* cuda_stream);

/* Slab [t0,t1) of rank `rank` when n3 slices are split over nranks ranks. */
int tritd_slab_bounds(int64_t n3, int nranks, int rank, int64_t* t0, int64_t* t1);

/* ---- the solver, one call (what the MEX gateway binds) ----------------- */

/* [A,B,C,O,errHist] = triple_decomp_ADMM(D, r, opts) with injected initial factors
 * (the reference draws them with randn at :23 in the order A, B, C; the gateway
 * does that in MATLAB and passes them in so the caller's rng stream is preserved).
 * n3 is the number of slices held by THIS rank (== global n3 on a single-rank
 * context); D_host/C0/C/O/L are this rank's slab.  O and L_or_null may be NULL
 * (then they are not copied back).  errHist must hold opts->maxIter doubles;
 * *iters_out receives the executed iteration count. */
int tritd_admm_f64(tritd_ctx* ctx, const double* D_host, int64_t n1, int64_t n2, int64_t n3, int r,
                   const tritd_opts* opts, const double* A0, const double* B0, const double* C0,
                   double* A, double* B, double* C, double* O, double* L_or_null,
                   double* errHist, int32_t* iters_out, tritd_timing* timing_or_null);

/* The same call with the two optional extensions of SURVEY 8f rank 2 (neither changes the default path):
 *  - E_or_null: the auxiliary sparse variable E ("O,E : sparse components (clone E)", triple_decomp_ADMM.m:12), a
 *    7th output of the gateway;
 *  - mask_or_null: the completion variant the reference's drivers name but do not ship
 *    ([A,B,C,O,E,Out] = triple_ADMM_masked(Y, ~mask_missing, r, opts), traffic_triple_comparison.m:53):
 *    dense n1 x n2 x n3 bytes, non-zero = observed.  Unobserved entries carry no data-fit constraint: there
 *    O = E = Y_L = Y_O = 0, they do not enter ||D||, resL or resO, and the low-rank target is imputed, T = L
 *    (zero for the first iteration, the drivers' own zero-fill).  On the observed entries every statement of
 *    :33-:65 is unchanged; an all-ones mask reproduces tritd_admm_f64 bit for bit.  Specified in DESIGN.md 4.6,
 *    restated in oracle/tritd_oracle.py (triple_ADMM_masked). */
int tritd_admm_ex_f64(tritd_ctx* ctx, const double* D_host, const unsigned char* mask_or_null, int64_t n1, int64_t n2,
                      int64_t n3, int r, const tritd_opts* opts, const double* A0, const double* B0, const double* C0,
                      double* A, double* B, double* C, double* O, double* E_or_null, double* L_or_null,
                      double* errHist, int32_t* iters_out, tritd_timing* timing_or_null);

/* tritd_admm_f64 keeps its device state (the N-sized arrays, tensor maps, the captured iteration graph) in the
 * context and reuses it when the next call has the same shape and rank; tritd_trim() releases it (tritd_destroy()
 * does so too). */
int tritd_trim(tritd_ctx* ctx);

/* ---- the solver, staged (device-resident state; used by benchmarks and by
 *      callers that keep L/O on the GPU) ------------------------------------ */

int tritd_problem_create(tritd_ctx* ctx, int64_t n1, int64_t n2, int64_t n3_local, int r, tritd_problem** out);
void tritd_problem_destroy(tritd_problem* p);

/* Load this rank's slab of D (dense column-major n1 x n2 x n3_local). */
int tritd_problem_set_D_host(tritd_problem* p, const double* D_host);
int tritd_problem_set_D_dev(tritd_problem* p, const double* D_dev);
/* Completion variant (see tritd_admm_ex_f64): mark the entries with mask == 0 as unobserved; call after set_D and
 * before init.  set_D clears the mask again. */
int tritd_problem_set_mask_host(tritd_problem* p, const unsigned char* mask_host);
int tritd_problem_set_mask_dev(tritd_problem* p, const unsigned char* mask_dev);

/* Reset the ADMM state (O = E = Y_L = Y_O = 0, mu = opts.mu, k = 0), install the
 * initial factors (host pointers, MATLAB 3-D shapes; C0 is this rank's slab) and
 * compute ||D|| (all-reduced over ranks). */
int tritd_problem_init(tritd_problem* p, const tritd_opts* opts, const double* A0, const double* B0,
                       const double* C0);

/* Run up to `max_more` further iterations (<= remaining maxIter); stops early when the
 * reference's stopping rule fires.  Asynchronous work is complete on return.
 * *iters_total receives the number of iterations executed since init. */
int tritd_problem_iterate(tritd_problem* p, int32_t max_more, int32_t* iters_total);

/* Enqueue exactly `n` iterations on the context's stream WITHOUT synchronising or
 * checking the stopping rule on the host (the device-side rule still turns later
 * iterations into no-ops).  For benchmarks that time with their own CUDA events. */
int tritd_problem_enqueue(tritd_problem* p, int32_t n);
int tritd_problem_sync(tritd_problem* p);

/* Copy results to host (any pointer may be NULL).  errHist/errL/errO receive
 * `iters` doubles each. */
int tritd_problem_get(tritd_problem* p, double* A, double* B, double* C, double* O, double* L,
                      double* errHist, double* errL, double* errO, int32_t* iters);
/* E of the last finished iteration (host / device destination). */
int tritd_problem_get_E(tritd_problem* p, double* E_host);
int tritd_problem_get_E_dev(tritd_problem* p, double* E_dev);
/* Device-side views (dense column-major copies written to caller-owned device memory). */
int tritd_problem_get_O_dev(tritd_problem* p, double* O_dev);
int tritd_problem_get_L_dev(tritd_problem* p, double* L_dev);
/* [rmse, nrmse] = evaluate(triple_product(A,B,C), gt, mask) with the factors the solver holds on the device. */
int tritd_problem_evaluate(tritd_problem* p, const double* gt_host, const unsigned char* mask_host, double* rmse,
                           double* nrmse);
/* pinv semantics of the ridge solves (:78/:86/:93): the solves invert the SPD system directly; when a pivot is not
 * positive or min pivot / max pivot < 16 R eps they run a Jacobi eigen-decomposition and zero the singular values
 * <= R * eps(sigma_max) exactly like MATLAB's pinv.  *fallbacks = solves of this run that took that path,
 * *truncated = singular values they zeroed in total. */
int tritd_problem_pinv_stats(tritd_problem* p, int32_t* fallbacks, int32_t* truncated);
/* Per-phase device timing: when enabled, every enqueued iteration is launched kernel by kernel (no graph replay)
 * and bracketed by CUDA events on the context's stream at its phase boundaries.  tritd_problem_phase_ms() synchronises, adds up
 * the elapsed milliseconds of each phase over all iterations recorded since the last call into
 * ms_out[TRITD_NPHASE] and reports how many iterations that was.  Phases:
 *   0 mode-1 MTTKRP (first iteration only; later the partials come from phase 4 and are summed in phase 1)
 *   1 update A: RHS reduction (+exchange), ridge inverse, apply, Gram(A)   2 shared pass P = T x_1 A
 *   3 updates B and C: RHS reductions (+exchange), ridge inverses, apply, Gram(B), Gram(C)
 *   4 fused L/O/E/dual/T/norm kernel, on the peer-exchange and single-rank paths including errHist/mu/stopping rule
 *   5 NCCL path only: all-reduce of the residual sums + errHist/mu/stopping rule */
#define TRITD_NPHASE 6
int tritd_problem_set_profiling(tritd_problem* p, int enable);
int tritd_problem_phase_ms(tritd_problem* p, double* ms_out, int32_t* iters_out);
/* Number of kernels launched by this library on this context so far. */
int64_t tritd_launch_count(const tritd_ctx* ctx);
/* Measured FP64 tensor-core (DMMA.8x8x4) peak of the context's GPU in TFLOP/s: about `ms_budget` milliseconds of
 * back-to-back probe launches.  The denominator of the FP64-tensor roofline fractions bench.py reports. */
int tritd_measure_dmma_peak(tritd_ctx* ctx, double ms_budget, double* tflops);

/* ---- the ALS solver -------------------------------- */

/* [A,B,C,errHist] = triple_decomp_ALS(X, r, opts)      triple_decomp_ALS.m:1-40
 * opts fields used by the reference: maxIter, tol (:2-3).  The relative error ||X - Xhat|| / ||X|| is taken
 * BEFORE the updates of an iteration (:15-16); when the relative-change rule fires (:20-23) the updates of that
 * iteration are skipped; all three ridge terms are 1e-9 (:27,:32,:37).  The reference prints
 * "Iteration %d, relative error = %.4e" every 5th iteration unconditionally (:17-19); here only when disp != 0.
 * Initial factors are injected like in tritd_admm_f64.  Single-rank contexts only. */
int tritd_als_f64(tritd_ctx* ctx, const double* X_host, int64_t n1, int64_t n2, int64_t n3, int r, int32_t maxIter,
                  double tol, int32_t disp, const double* A0, const double* B0, const double* C0,
                  double* A, double* B, double* C, double* errHist, int32_t* iters_out);

/* ---- standalone L2 helpers (host pointers, MATLAB shapes) --------------- */

/* Xhat = triple_product(A,B,C)                      triple_product.m:1-8   */
int tritd_triple_product_f64(tritd_ctx* ctx, const double* A, const double* B, const double* C,
                             int64_t n1, int64_t n2, int64_t n3, int r, double* Xhat);
/* the same with the result left on the device (dense column-major, caller-owned device memory) */
int tritd_triple_product_dev_f64(tritd_ctx* ctx, const double* A, const double* B, const double* C,
                                 int64_t n1, int64_t n2, int64_t n3, int r, double* Xhat_dev);
/* Xn = unfold(X, mode), mode in {1,2,3}             unfold.m:1-14          */
int tritd_unfold_f64(tritd_ctx* ctx, const double* X, int64_t n1, int64_t n2, int64_t n3, int mode, double* Xn);
int tritd_unfold_dev_f64(tritd_ctx* ctx, const double* X_dev, int64_t n1, int64_t n2, int64_t n3, int mode, double* Xn_dev);
/* F = buildF(B,C)  r^2 x (n2 n3)                    buildF.m:17-21         */
int tritd_buildF_f64(tritd_ctx* ctx, const double* B, const double* C, int64_t n2, int64_t n3, int r, double* F);
/* G = buildG(A,C)  r^2 x (n1 n3)                    buildG.m:17-21         */
int tritd_buildG_f64(tritd_ctx* ctx, const double* A, const double* C, int64_t n1, int64_t n3, int r, double* G);
/* H = buildH(A,B)  r^2 x (n1 n2)                    buildH.m:17-21         */
int tritd_buildH_f64(tritd_ctx* ctx, const double* A, const double* B, int64_t n1, int64_t n2, int r, double* H);
/* device-pointer form of buildF/G/H: which = 0 F(B,C), 1 G(A,C), 2 H(A,B); U, V the MATLAB 3-D factor arrays,
 * na / nb the two mode sizes (n2,n3 / n1,n3 / n1,n2); out r^2 x (na*nb) column-major */
int tritd_build_design_dev_f64(tritd_ctx* ctx, int which, const double* U_dev, const double* V_dev, int64_t na, int64_t nb,
                               int r, double* out_dev);
/* O = soft_threshold(X, lam), n elements            soft_threshold.m:2     */
int tritd_soft_threshold_f64(tritd_ctx* ctx, const double* X, int64_t n, double lam, double* out);
int tritd_soft_threshold_dev_f64(tritd_ctx* ctx, const double* X_dev, int64_t n, double lam, double* out_dev);
/* Design matrices and product of the ORIGINAL (Qi) triple decomposition, the model the README's
 * RPAS claim describes; origin_triple_tensor/buildF.m:4-6, buildG.m:9-11, buildH.m:9-11, triple_product.m.
 *   which = 0: F = buildF(B,C), U = B (r x n2 x r), V = C (r x r x n3), na = n2, nb = n3,
 *              F(q+(s-1)r, j+(t-1)n2) = sum_p B(p,j,s) C(p,q,t)
 *   which = 1: G = buildG(A,C), U = A (n1 x r x r), V = C, na = n1, nb = n3,
 *              G(p+(s-1)r, i+(t-1)n1) = sum_q A(i,q,s) C(p,q,t)
 *   which = 2: H = buildH(A,B), U = A, V = B, na = n1, nb = n2,
 *              H(p+(q-1)r, i+(j-1)n1) = sum_s A(i,q,s) B(p,j,s)
 * out is r^2 x (na*nb) column-major.  Only these helpers exist for the Qi model; the solver implements the
 * rank-r^2 CP form of fast_robust_triple_tensor/ (SURVEY fact 1). */
int tritd_design_qi_f64(tritd_ctx* ctx, int which, const double* U, const double* V, int64_t na, int64_t nb, int r,
                        double* out);
/* Xhat(i,j,t) = sum_{p,q,s} A(i,q,s) B(p,j,s) C(p,q,t)     origin_triple_tensor/triple_product.m */
int tritd_triple_product_qi_f64(tritd_ctx* ctx, const double* A, const double* B, const double* C, int64_t n1,
                                int64_t n2, int64_t n3, int r, double* Xhat);
/* [rmse, nrmse] = evaluate(Xhat, gt, mask) of the reference's drivers (traffic_triple_comparison.m:194-202) with
 * Xhat = triple_product(A,B,C) formed on the device (the reconstruction never crosses PCIe).
 * gt: dense n1 x n2 x n3 ground truth (entries outside the mask are ignored); mask: dense n1 x n2 x n3 bytes
 * (non-zero = evaluated) or NULL for all entries, which gives the drivers' RRE ||Xhat - gt|| / ||gt||. */
int tritd_evaluate_f64(tritd_ctx* ctx, const double* A, const double* B, const double* C, int64_t n1, int64_t n2,
                       int64_t n3, int r, const double* gt_host, const unsigned char* mask_host, double* rmse,
                       double* nrmse);
/* The three contractions of one sweep at fixed factors (what update_A/B/C feed to pinv):
 * rhsA = X1*F' (n1 x r^2), rhsB = X2*G' (n2 x r^2), rhsC = X3*H' (n3 x r^2), column-major. */
int tritd_mttkrp_f64(tritd_ctx* ctx, const double* X, const double* A, const double* B, const double* C,
                     int64_t n1, int64_t n2, int64_t n3, int r, int mode, double* rhs);

/* One factor update in isolation -- what update_A/B/C do after the contraction (:77-78 / :86 / :93):
 *   X = rhs * pinv(S1 o S2 + alpha*I)          rhs n x r^2, S1, S2 r^2 x r^2 (column-major), o = Hadamard product
 * through the solver's own update kernel.  Optional outputs: the pseudo-inverse (r^2 x r^2), X'X (r^2 x r^2),
 * info[0] = 1 when the truncating pinv path ran, info[1] = singular values it zeroed. */
int tritd_factor_update_f64(tritd_ctx* ctx, const double* rhs, int64_t n, int r, const double* S1, const double* S2,
                            double alpha, double* X, double* Ginv_or_null, double* XtX_or_null, int32_t* info_or_null);

#ifdef __cplusplus
}
#endif
#endif /* TRITD_H */
