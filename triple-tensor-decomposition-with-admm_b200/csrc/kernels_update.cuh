// k_upd: one factor update of the Gauss-Seidel sweep as ONE kernel
// (reference: fast_robust_triple_tensor/triple_decomp_ADMM.m:73-95):
//     RHS = X_(k) * M'                       -- the last reduction step of the MTTKRP (sources below)
//     X   = RHS * pinv(M*M' + alpha*I)       -- :78 / :86 / :93, M*M' = S1 o S2 (Hadamard identity)
//     S   = X'X                              -- the small Gram the next two updates need
// CTA roles (256 threads each):
//   block 0      inverts the R x R ridge system (in-place Gauss-Jordan, SPD => no pivoting, the pivots are the
//                Cholesky pivots L_kk^2; a non-positive / non-finite pivot is reported through IterState::status)
//                WHILE the row CTAs reduce their RHS rows, publishes inv(G) and raises flags[0];
//   blocks 1..   8/wpr rows each: fixed-order reduction of the RHS rows from the source, wait for flags[0],
//                apply the inverse like the reference applies pinv(G), write X (and its transpose for TMA);
//   blocks 0..G-1 finally form S = X'X straight from the freshly written rows once every row CTA has
//                signalled flags[1] (one entry per thread, rows summed in a fixed order => deterministic).
// The inverse is off the critical path, nothing is launched between "RHS ready" and "factor ready", and no
// partial-Gram buffers exist.  Row CTAs wait only for block 0 (scheduled first, waits for nobody); the Gram
// phase waits for the row CTAs, which never wait for it: no cyclic dependency, at most G CTAs spin.
//
// With apply == 0 the kernel only reduces: rows go to rhs_out (N>1: the all-reduce comes next).
#pragma once
#include "common.cuh"
#include "kernels_fused.cuh"
#include "kernels_xchg.cuh"

namespace tritd {

constexpr int kStatusNumeric = 1;     // a ridge system contained NaN / Inf (MATLAB's pinv raises an error there too)

// The RHS source is a strided sum: row `row` of the RHS is  sum_{m < count} w[m*wstride + k] * v[row_base + m*stride + k]
// with row_base = (row / tile_h) * tile_stride + (row % tile_h) * row_stride, so ONE kernel body (one set of
// instruction addresses, which stays warm in the instruction cache across the three updates of an iteration) serves
//   direct      count = 1, v = rhs, w = 1                        (first iteration; N>1 after the all-reduce)
//   mode-1      v = k_admm's per-CTA partials of X1*F', w = 1    (update_A)
//   P over t    v = P[t][row][k], w = C3[t][k]                   (update_B, :86)
//   P over j    v = P[row][j][k], w = B2[j][k]                   (update_C, :93)
struct UpdArgs {
    const double* v;          // items
    const double* w;          // weights (a vector of ones with wstride 0 when the source is a plain sum)
    long stride, wstride;     // item stride of v / w in doubles
    long tile_stride, row_stride;
    int tile_h;               // rows per tile of the source (INT_MAX when rows are simply row_stride apart)
    int count;                // items per row
    const int* tile_cnt;      // optional: items per row for the rows of source tile q (else `count` for all rows)
    int wpr;                  // warps that share a row: 1 (8 rows per CTA) or 8 (1 row per CTA)
    const double *S1, *S2;    // [RS][RS] small Grams of the two other factors; S2 may be a stack of ns2 matrices
    int ns2;                  // (the per-rank partials of C3'C3 in the exchange mailbox), summed in order
    long s2stride;
    // N>1 peer exchange inside the kernel (xmerge): every row CTA writes its reduced rows into this rank's slot of every
    // rank's mailbox as self-validating words (value + epoch, kernels_xchg.cuh), waits until the words of all ranks in
    // its own mailbox show the epoch and sums the slots in rank order.  The rows are partitioned over the CTAs
    // identically on every rank, so a thread only needs what its counterparts on the other ranks pushed: no grid-wide
    // step, no CTA waits for a CTA of the same grid.
    // C3'C3 (the one small Gram with per-rank partials) travels EARLY: update C pushes its local partial from the end
    // of its Gram phase (sc_push) into a parity-double-buffered region of every mailbox, so that the ridge inverses
    // of the next iteration's updates A and B (sc_wait: S2 = the stack of the ranks' partials) start at once instead of
    // after an exchange of their own.
    int xmerge;
    int sc_wait, sc_push;     // this launch consumes / produces the exchanged C3'C3
    long s2par_stride;        // doubles between the two parity buffers of the S2 stack
    const unsigned* sc_flags; // own mailbox: [2 parities][8] epochs of the C3'C3 pushes
    long sc_off, sc_flag_off; // offsets inside a mailbox: doubles to the C3'C3 region, u32 index of its flags (from the flag area)
    long flag_area_off;       // doubles from the mailbox base to the flag area
    int gram_cap;             // CTAs that may take part (and spin) in the Gram phase
    int inv_here;             // block 0 inverts the ridge system (else an earlier kernel published it and raised flags[0])
    double* const* peers;     // [nranks] mailbox bases
    long push_off;            // offset in doubles inside a mailbox: this rank's slot of the exchange (2 doubles per element)
    int rank, nranks;
    const double* xbox;       // slot 0 of the exchange region in the own mailbox
    long xslot;               // slot stride
    unsigned xbase;           // epoch of iteration k = xbase + k + 1
    double alpha;
    double* Minv;             // [R][RS] scratch: inv(S1 o S2 + alpha I), written by block 0
    double* rhs_out;          // apply == 0: reduced rows [n][RS]
    double* X;                // [n][RS]
    double* XT;               // [RS][ldt] or nullptr
    double* gram_out;         // [RS][RS] = X'X over the rows of this rank
    double* gram_part;        // [gr][RS][RS] scratch: row-slice partials
    unsigned* gram_cnt;       // [<= 64] per entry-slice arrival counters (zero between launches)
    int gr;                   // row slices of the Gram phase
    IterState* st;
    unsigned* flags;          // [0] inverse published, [1] row CTAs done, [2] Gram CTAs done, [3] pre-computation punted; all zero between launches
    int apply;
    int n, R, RS, ldt;
    long long* dbg;           // optional [16] globaltimer stamps (diagnostics; TRITD_DEBUG_STAMPS)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_acq_rel_add_u32(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
// CTA-wide wait until *p == want: thread 0 polls with acquire loads, the barrier extends the acquire to the CTA
// (release/acquire are cumulative over bar.sync, so no full fence.sc is needed on either side).
__device__ __forceinline__ void cta_wait_eq(const unsigned* p, unsigned want, int* status) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_u32(p) != want) {
            __nanosleep(20);
            if (++spins > kSpinLimit) { atomicExch(status, kStatusHang); break; }     // bounded: see kernels_xchg.cuh
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double rcp_newton(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// In-place Gauss-Jordan inverse of G = S1 o S2 + alpha*I by one CTA of 256 threads, matrix in registers:
// thread (ty,tx) of a 16 x 16 grid owns entries (ty + 16p, tx + 16q), p,q < PQ.  Per step only the pivot row and
// column pass through shared memory (double-buffered, published by their owners as they are produced): one
// barrier, 2*PQ+1 shared loads, one reciprocal, PQ*PQ FMAs.  Returns 0 when the inverse can be trusted to agree with
// the reference's pinv, 1 when a pivot is non-positive / non-finite or min pivot / max pivot < 16 R eps (pinv's cutoff
// max(size) * eps(sigma_max) may truncate: the caller runs pinv_jacobi instead), 2 in the threads whose entries of the
// system itself are NaN / Inf (the caller votes).
constexpr int kRidgeOk = 0, kRidgeIll = 1, kRidgeNonFinite = 2;
// BAR = 0: the calling CTA has exactly 256 threads (__syncthreads); BAR > 0: the first 256 threads of a larger CTA call
// it and synchronise on named barrier BAR.
template <int BAR> __device__ __forceinline__ void bar256() {
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, 256;" ::"n"(BAR) : "memory");
}
template <int PQ, int BAR = 0>
__device__ int invert_ridge_system(const double* S1, const double* S2, int ns2, long s2stride, double alpha, int R, int RS, double* out,
                                    double* sm /* >= 256 doubles */, long long* dbg = nullptr) {
    double* prow = sm;        // [2][64]
    double* pcol = sm + 128;  // [2][64]
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double gq[PQ][PQ];
    bool input_nonfinite = false;         // NaN / Inf in THIS thread's entries of the system (the caller votes over the CTA)
    {
        // all loads first (2*PQ*PQ in flight per thread), then the Hadamard product and the ridge
        double t1[PQ][PQ], t2[PQ][PQ];
#pragma unroll
        for (int p = 0; p < PQ; ++p)
#pragma unroll
            for (int q = 0; q < PQ; ++q) {
                const int i = ty + 16 * p, j = tx + 16 * q;
                const bool ok = i < R && j < R;
                t1[p][q] = ok ? S1[i * RS + j] : 0.0;
                t2[p][q] = ok ? S2[i * RS + j] : 0.0;
            }
        if (ns2 > 1) {                      // S2 is a stack of per-rank partials: add the others in rank order
#pragma unroll
            for (int p = 0; p < PQ; ++p)
#pragma unroll
                for (int q = 0; q < PQ; ++q) {
                    const int i = ty + 16 * p, j = tx + 16 * q;
                    if (i < R && j < R) {
                        double t[7];
#pragma unroll
                        for (int u = 0; u < 7; ++u) t[u] = u + 1 < ns2 ? S2[(u + 1) * s2stride + i * RS + j] : 0.0;
#pragma unroll
                        for (int u = 0; u < 7; ++u) t2[p][q] += t[u];
                    }
                }
        }
#pragma unroll
        for (int p = 0; p < PQ; ++p)
#pragma unroll
            for (int q = 0; q < PQ; ++q) {
                const int i = ty + 16 * p, j = tx + 16 * q;
                double v = t1[p][q] * t2[p][q];
                if (i == j && i < R) v += alpha;
                input_nonfinite = input_nonfinite || !isfinite(v);
                if (i == 0) prow[j] = v;
                if (j == 0) pcol[i] = v;
                gq[p][q] = v;
            }
    }
    bar256<BAR>();
    if (dbg && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[13] = t_; }
    bool bad = false;
    double pmin = 1.7976931348623157e308, pmax = 0.0;
    // Steps k = 16*KP + kl with the 16-block KP of the pivot a compile-time constant (unrolled), so the pivot row and
    // column are fixed registers: one generic FMA per entry, then per-step fix-ups of row k / column k.
#pragma unroll
    for (int KP = 0; KP < PQ; ++KP) {
#pragma unroll 1
        for (int kl = 0; kl < 16; ++kl) {
            const int k = 16 * KP + kl;
            if (k >= R) break;
            const double* pr = prow + (k & 1) * 64;
            const double* pc = pcol + (k & 1) * 64;
            double* prn = prow + ((k + 1) & 1) * 64;
            double* pcn = pcol + ((k + 1) & 1) * 64;
            const double piv = pr[k];
            bad = bad || !(piv > 0.0) || !isfinite(piv);      // (a zero pivot makes the later ones Inf / NaN: still "ill", not an input error)
            pmin = fmin(pmin, piv); pmax = fmax(pmax, piv);
            const double inv = rcp_newton(piv);
            double prj[PQ], pci[PQ];     // entries outside R x R see zeros here and stay zero
#pragma unroll
            for (int q = 0; q < PQ; ++q) prj[q] = (tx + 16 * q < R) ? pr[tx + 16 * q] * inv : 0.0;
#pragma unroll
            for (int p = 0; p < PQ; ++p) pci[p] = (ty + 16 * p < R) ? pc[ty + 16 * p] : 0.0;
#pragma unroll
            for (int p = 0; p < PQ; ++p)
#pragma unroll
                for (int q = 0; q < PQ; ++q) gq[p][q] = fma(-pci[p], prj[q], gq[p][q]);
            const bool colk = tx == kl, rowk = ty == kl;
            if (colk) {                                  // column k: -pc[i] * inv
#pragma unroll
                for (int p = 0; p < PQ; ++p) gq[p][KP] = -pci[p] * inv;
            }
            if (rowk) {                                  // row k: pr[j] * inv, and inv at the pivot
#pragma unroll
                for (int q = 0; q < PQ; ++q) gq[KP][q] = prj[q];
                if (colk) gq[KP][KP] = inv;
            }
            // publish the next pivot row and column
            if (kl < 15) {
                if (ty == kl + 1) {
#pragma unroll
                    for (int q = 0; q < PQ; ++q) prn[tx + 16 * q] = gq[KP][q];
                }
                if (tx == kl + 1) {
#pragma unroll
                    for (int p = 0; p < PQ; ++p) pcn[ty + 16 * p] = gq[p][KP];
                }
            } else if (KP + 1 < PQ) {
                const int NP = KP + 1 < PQ ? KP + 1 : PQ - 1;
                if (ty == 0) {
#pragma unroll
                    for (int q = 0; q < PQ; ++q) prn[tx + 16 * q] = gq[NP][q];
                }
                if (tx == 0) {
#pragma unroll
                    for (int p = 0; p < PQ; ++p) pcn[ty + 16 * p] = gq[p][NP];
                }
            }
            bar256<BAR>();
        }
    }
#pragma unroll
    for (int p = 0; p < PQ; ++p)
#pragma unroll
        for (int q = 0; q < PQ; ++q) {
            const int i = ty + 16 * p, j = tx + 16 * q;
            if (i < R && j < R) out[i * RS + j] = gq[p][q];
        }
    if (input_nonfinite) return kRidgeNonFinite;
    return (bad || pmin < pmax * (16.0 * 2.220446049250313e-16) * R) ? kRidgeIll : kRidgeOk;
}

// pinv(G) of the symmetric ridge system G = S1 o S2 + alpha*I exactly as the reference forms it (:78/:86/:93):
// MATLAB's pinv is SVD based and zeroes the singular values <= max(size(G)) * eps(norm(G)).  For a symmetric
// matrix the singular values are |eigenvalues|, so: cyclic Jacobi eigen-decomposition G = V diag(lam) V' (parallel
// round-robin ordering, R/2 disjoint rotations per round, matrix and eigenvectors in shared memory), then
// pinv = V diag(1/lam_k if |lam_k| > R * eps(max|lam|) else 0) V'.  Only block 0 runs this, and only when
// invert_ridge_system reported kRidgeIll (rank-deficient or nearly so: duplicated factor columns, lambda2 = 0, the
// 1e-9 ridge of update_C against sigma_max > ~1e5): a rare path, ~1 ms.  Returns the number of truncated values.
template <int BAR = 0>
__device__ int pinv_jacobi(const double* S1, const double* S2, int ns2, long s2stride, double alpha, int R, int RS, double* out,
                           double* sm /* >= 2*R*(R+1) + 3*64 doubles */) {
    const int tid = threadIdx.x, P = R + 1;
    constexpr int NTH = 256;
    double* G = sm;                       // [R][P]
    double* V = sm + R * P;               // [R][P]
    double* cs = V + R * P;               // [32 pairs][2] rotations of the round, then [64] weights
    __shared__ int s_rot, s_trunc;
    for (int e = tid; e < R * R; e += NTH) {
        const int i = e / R, j = e - i * R;
        double t2 = S2[i * RS + j];
        for (int u = 1; u < ns2; ++u) t2 += S2[u * s2stride + i * RS + j];
        double v = S1[i * RS + j] * t2;
        if (i == j) v += alpha;
        G[i * P + j] = v;
        V[i * P + j] = i == j ? 1.0 : 0.0;
    }
    bar256<BAR>();
    const int n = (R + 1) & ~1, half = n >> 1;       // players of the round-robin tournament (a dummy one when R is odd)
    for (int sweep = 0; sweep < 30; ++sweep) {
        if (tid == 0) s_rot = 0;
        bar256<BAR>();
        for (int rd = 0; rd < n - 1; ++rd) {
            // pair q of round rd: (n-1, rd) for q = 0, else ((rd+q) mod (n-1), (rd-q) mod (n-1))
            if (tid < half) {
                int p_ = tid == 0 ? n - 1 : (rd + tid) % (n - 1), q_ = tid == 0 ? rd : (rd - tid + (n - 1)) % (n - 1);
                if (p_ > q_) { const int t_ = p_; p_ = q_; q_ = t_; }
                double c = 1.0, sn = 0.0;
                if (q_ < R) {
                    const double gpp = G[p_ * P + p_], gqq = G[q_ * P + q_], gpq = G[p_ * P + q_];
                    if (gpq != 0.0 && fabs(gpq) > 1e-18 * sqrt(fabs(gpp) * fabs(gqq))) {
                        const double tau = (gqq - gpp) / (2.0 * gpq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        sn = t * c;
                        atomicAdd(&s_rot, 1);
                    }
                }
                cs[2 * tid] = c; cs[2 * tid + 1] = sn;
            }
            bar256<BAR>();
            // columns p,q of G and V for every row: (x_p, x_q) <- (c x_p - s x_q, s x_p + c x_q)
            for (int e = tid; e < R * half; e += NTH) {
                const int i = e / half, pr = e - i * half;
                int p_ = pr == 0 ? n - 1 : (rd + pr) % (n - 1), q_ = pr == 0 ? rd : (rd - pr + (n - 1)) % (n - 1);
                if (p_ > q_) { const int t_ = p_; p_ = q_; q_ = t_; }
                const double c = cs[2 * pr], sn = cs[2 * pr + 1];
                if (q_ < R && sn != 0.0) {
                    const double gp = G[i * P + p_], gq = G[i * P + q_];
                    G[i * P + p_] = c * gp - sn * gq; G[i * P + q_] = sn * gp + c * gq;
                    const double vp = V[i * P + p_], vq = V[i * P + q_];
                    V[i * P + p_] = c * vp - sn * vq; V[i * P + q_] = sn * vp + c * vq;
                }
            }
            bar256<BAR>();
            // rows p,q of G for every column
            for (int e = tid; e < R * half; e += NTH) {
                const int j = e / half, pr = e - j * half;
                int p_ = pr == 0 ? n - 1 : (rd + pr) % (n - 1), q_ = pr == 0 ? rd : (rd - pr + (n - 1)) % (n - 1);
                if (p_ > q_) { const int t_ = p_; p_ = q_; q_ = t_; }
                const double c = cs[2 * pr], sn = cs[2 * pr + 1];
                if (q_ < R && sn != 0.0) {
                    const double gp = G[p_ * P + j], gq = G[q_ * P + j];
                    G[p_ * P + j] = c * gp - sn * gq; G[q_ * P + j] = sn * gp + c * gq;
                }
            }
            bar256<BAR>();
        }
        if (s_rot == 0) break;            // (uniform: read after the barrier that ended the last round)
        bar256<BAR>();
    }
    // weights 1/lam_k above MATLAB's cutoff max(size(G)) * eps(norm(G)), eps(x) = 2^(floor(log2 x) - 52)
    if (tid == 0) {
        double lmax = 0.0;
        for (int k = 0; k < R; ++k) lmax = fmax(lmax, fabs(G[k * P + k]));
        const double tol = lmax > 0.0 ? (double)R * ldexp(1.0, ilogb(lmax) - 52) : 0.0;
        int nt = 0;
        for (int k = 0; k < R; ++k) {
            const double l = G[k * P + k];
            const bool keep = fabs(l) > tol;
            cs[k] = keep ? 1.0 / l : 0.0;
            nt += keep ? 0 : 1;
        }
        s_trunc = nt;
    }
    bar256<BAR>();
    for (int e = tid; e < R * R; e += NTH) {
        const int i = e / R, j = e - i * R;
        double acc = 0.0;
        for (int k = 0; k < R; ++k) acc = fma(V[i * P + k] * cs[k], V[j * P + k], acc);
        out[i * RS + j] = acc;
    }
    bar256<BAR>();
    return s_trunc;
}

// The whole ridge-system step for a group of 256 threads (k_upd's block 0, or the first 256 threads of the first
// k_admm / k_ppass CTA to finish, which pre-compute the NEXT update's inverse while their grid drains): direct inverse,
// the pinv path when the reference's pinv would truncate, status and statistics.  `sm` must hold
// max(256, 2 R (R+1) + 192) doubles.
struct RidgeJob {
    const double *S1, *S2;    // [RS][RS]; S2 may be a stack of ns2 matrices (per-rank partials), summed in order
    int ns2;
    long s2stride;
    double alpha;
    double* Minv;             // [R][RS] out
    int enable;
    // N>1: S2 = the exchanged C3'C3, two parity buffers; wait for the ranks' pushes (see UpdArgs::sc_wait)
    int sc_wait;
    long s2par_stride;
    const unsigned* sc_flags;
    int nranks;
    unsigned xbase;
    unsigned* done_flag;      // done_flag[0] = 1 (release) when Minv is published; done_flag[3] = 1 instead when a
                              // pre-computing kernel punts (ill-conditioned / non-finite system: the update kernel's block 0,
                              // which has the shared memory for the pinv path, then does the whole job itself)
};

// FULL: with the pinv path (needs max(256, 2 R (R+1) + 192) doubles of scratch); else 256 doubles suffice and an
// ill-conditioned system is left to the update kernel (done_flag[3]).
template <int PQ, int BAR, bool FULL>
__device__ __forceinline__ void ridge_job_run(const RidgeJob& j, IterState* st, int k_of_consumer, int R, int RS, double* sm, long long* dbg = nullptr) {
    const int tid = threadIdx.x;
    const double* S2 = j.S2;
    if (j.sc_wait) {
        // the ranks' C3'C3 partials pushed by the previous update C (or the initialisation): parity (k+1)&1, epoch xbase + k,
        // k = iteration index the CONSUMING update runs in
        const unsigned want = j.xbase + (unsigned)k_of_consumer;
        const int par = (k_of_consumer + 1) & 1;
        if (tid < j.nranks)
            spin_until_epoch(j.sc_flags + par * 8 + tid, want, &st->status);
        bar256<BAR>();
        S2 += par * j.s2par_stride;
    }
    const int cond = invert_ridge_system<PQ, BAR>(j.S1, S2, j.ns2, j.s2stride, j.alpha, R, RS, j.Minv, sm, dbg);
    // (`cond` is uniform except for kRidgeNonFinite, which every thread decides from its own entries: vote)
    __shared__ int s_vote[2];
    if (tid == 0) { s_vote[0] = 0; s_vote[1] = 0; }
    bar256<BAR>();
    if (cond == kRidgeNonFinite) s_vote[0] = 1;
    if (cond == kRidgeIll) s_vote[1] = 1;
    bar256<BAR>();
    if (!FULL) {
        if (tid == 0) st_release_u32((s_vote[0] || s_vote[1]) ? j.done_flag + 3 : j.done_flag, 1u);
        return;
    }
    if (s_vote[0]) {
        if (tid == 0) atomicExch(&st->status, kStatusNumeric);
    } else if (s_vote[1]) {
        // the reference's pinv would (or might) truncate: do exactly what it does
        const int nt = pinv_jacobi<BAR>(j.S1, S2, j.ns2, j.s2stride, j.alpha, R, RS, j.Minv, sm);
        if (tid == 0) { atomicAdd(&st->pinv_fallbacks, 1); atomicAdd(&st->pinv_truncated, nt); }
    }
    bar256<BAR>();
    if (tid == 0) st_release_u32(j.done_flag, 1u);
}

// S[a][b] = sum_i X[i][a] * X[i][b] over rows [0,n) of a row-major n x RS factor; S is RS x RS.
// One thread per (a,b); rows are summed in order in 4 interleaved chains (deterministic).
__global__ void __launch_bounds__(256) k_small_gram(const double* X, int n, int RS, double* S, const int* stop) {
    if (stop && *stop) return;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= RS * RS) return;
    const int aa = idx / RS, bb = idx - aa * RS;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = 0;
    for (; i + 3 < n; i += 4) {
        s0 = fma(X[(size_t)i * RS + aa], X[(size_t)i * RS + bb], s0);
        s1 = fma(X[(size_t)(i + 1) * RS + aa], X[(size_t)(i + 1) * RS + bb], s1);
        s2 = fma(X[(size_t)(i + 2) * RS + aa], X[(size_t)(i + 2) * RS + bb], s2);
        s3 = fma(X[(size_t)(i + 3) * RS + aa], X[(size_t)(i + 3) * RS + bb], s3);
    }
    for (; i < n; ++i) s0 = fma(X[(size_t)i * RS + aa], X[(size_t)i * RS + bb], s0);
    S[idx] = (s0 + s1) + (s2 + s3);
}

constexpr int kUpdThreads = 256;

constexpr int kUpdMaxGramCtas = 64;   // CTAs that may spin in the Gram phase (<< 148 SMs x resident CTAs)
__host__ __device__ inline size_t upd_smem_bytes(int RS) {
    const int ms = RS * RS > 64 * RS ? RS * RS : 64 * RS;      // inv(G) tile / Gram row chunk [64][RS]
    const int jac = 2 * RS * (RS + 1) + 3 * 64;                // pinv_jacobi: G, V and the rotations (block 0 only)
    const int tot = 3 * 8 * 64 + ms;
    return (size_t)(tot > jac ? tot : jac) * sizeof(double);
}

template <int PQ>   // ceil(R / 16): the register patch of the inversion; also fixes the columns per lane (1 for RS <= 32, else 2)
__global__ void __launch_bounds__(kUpdThreads) k_upd(const UpdArgs a) {
    pdl_wait();
    pdl_trigger();
    if (a.st->stop) return;
    constexpr int KPL = PQ <= 2 ? 1 : 2;
    constexpr int CH = 8;                         // independent accumulation chains (= loads in flight) per lane and column
    extern __shared__ double sm[];
    double* red = sm;                             // [8 warps][64]  (block 0: pivot row/column buffers)
    double* rhs_s = sm + 512;                     // [rows][64]
    double* Ms = sm + 1536;                       // [R][RS] inverse, later the Gram row chunk [64][RS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.R, RS = a.RS;
    const int wpr = a.wpr, rows = 8 / wpr;        // rows per row CTA; wpr warps share a row
    const int nrowcta = (a.n + rows - 1) / rows;
#define TRITD_STAMP(blk, q)                                                         \
    if (a.dbg && blockIdx.x == (blk) && tid == 0) {                                 \
        long long t_;                                                               \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                      \
        a.dbg[q] = t_;                                                              \
    }
    TRITD_STAMP(0, 0)
    TRITD_STAMP(1, 4)
    const unsigned epoch = a.xbase + (unsigned)a.st->k + 1u;

    if (blockIdx.x == 0 && !a.apply) {
        return;
    } else if (blockIdx.x == 0) {
        // ---------------- the ridge system, inverted while the row CTAs reduce ----------------
        // (unless the previous k_admm / k_ppass already did it while its grid drained: inv_here == 0)
        if (a.inv_here || ld_acquire_u32(&a.flags[3]) != 0u) {       // (a pre-computing kernel has long finished: plain stream order)
            RidgeJob j;
            j.S1 = a.S1; j.S2 = a.S2; j.ns2 = a.ns2; j.s2stride = a.s2stride; j.alpha = a.alpha; j.Minv = a.Minv; j.enable = 1;
            j.sc_wait = a.sc_wait; j.s2par_stride = a.s2par_stride; j.sc_flags = a.sc_flags; j.nranks = a.nranks; j.xbase = a.xbase;
            j.done_flag = &a.flags[0];
            ridge_job_run<PQ, 0, true>(j, a.st, a.st->k, R, RS, sm, a.dbg);
        }
        TRITD_STAMP(0, 1)
    } else {
        // ---------------- RHS rows: fixed-order strided sum ----------------
        const int row0 = (blockIdx.x - 1) * rows;
        const int r = warp / wpr, sub = warp - r * wpr;
        const int row = row0 + r;
        double acc[KPL][CH];
#pragma unroll
        for (int q = 0; q < KPL; ++q)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[q][c] = 0.0;
        bool kok[KPL];
#pragma unroll
        for (int q = 0; q < KPL; ++q) kok[q] = lane + 32 * q < RS;
        if (row < a.n) {
            const int it = row / a.tile_h, il = row - it * a.tile_h;
            const double* vb = a.v + it * a.tile_stride + il * a.row_stride + lane;
            const double* wb = a.w + lane;
            const int count = a.tile_cnt ? a.tile_cnt[it] : a.count;
            // items m = sub + wpr * (c + CH * round); the loads of a round are issued together, then accumulated;
            // out-of-range items of the last round read item 0 with weight 0
            for (int m0 = sub; m0 < count; m0 += wpr * CH) {
                double v[KPL][CH], w[KPL][CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int m = m0 + c * wpr;
                    const bool ok = m < count;
                    const long mv = ok ? m * a.stride : 0, mw = ok ? m * a.wstride : 0;
#pragma unroll
                    for (int q = 0; q < KPL; ++q) {
                        v[q][c] = kok[q] ? vb[mv + 32 * q] : 0.0;
                        w[q][c] = (ok && kok[q]) ? wb[mw + 32 * q] : 0.0;
                    }
                }
#pragma unroll
                for (int c = 0; c < CH; ++c)
#pragma unroll
                    for (int q = 0; q < KPL; ++q) acc[q][c] = fma(w[q][c], v[q][c], acc[q][c]);
            }
        }
#pragma unroll
        for (int q = 0; q < KPL; ++q)
            red[warp * 64 + lane + 32 * q] =
                ((acc[q][0] + acc[q][1]) + (acc[q][2] + acc[q][3])) + ((acc[q][4] + acc[q][5]) + (acc[q][6] + acc[q][7]));
        __syncthreads();
        // combine the wpr warps of a row in warp order: thread (rr = warp, k = lane + 32q) for rr < rows
        if (warp < rows) {
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                const int k = lane + 32 * q;
                if (k < RS) {
                    double v = red[(warp * wpr) * 64 + k];
                    for (int s2 = 1; s2 < wpr; ++s2) v += red[(warp * wpr + s2) * 64 + k];
                    rhs_s[warp * 64 + k] = v;
                    if (row0 + warp < a.n) {
                        if (!a.apply) a.rhs_out[(size_t)(row0 + warp) * RS + k] = v;
                        else if (a.xmerge)
                            for (int r = 0; r < a.nranks; ++r)
                                ll_store(a.peers[(a.rank + 1 + r) % a.nranks] + a.push_off + 2 * ((size_t)(row0 + warp) * RS + k), v, epoch);
                    }
                }
            }
        }
        if (!a.apply) return;
        if (a.xmerge) {
            // the all-reduced rows: every thread waits for ITS element from every rank (the words validate themselves,
            // kernels_xchg.cuh) and sums the ranks' slots in rank order -- the same sum on every rank
            if (warp < rows && row0 + warp < a.n) {
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    const int k = lane + 32 * q;
                    if (k < RS)
                        rhs_s[warp * 64 + k] = ll_sum_ranks(a.xbox + 2 * ((size_t)(row0 + warp) * RS + k), a.xslot, a.nranks, epoch, &a.st->status);
                }
            }
        }
    }
    if (blockIdx.x != 0) {
        const int row0 = (blockIdx.x - 1) * rows;
        TRITD_STAMP(1, 5)

        // ---------------- apply the inverse: X[row][:] = RHS[row][:] * inv(G) ----------------
        cta_wait_eq(&a.flags[0], 1u, &a.st->status);              // (also orders the rhs_s writes above)
        TRITD_STAMP(1, 6)
        for (int e = tid; e < R * RS; e += kUpdThreads) Ms[e] = __ldcg(a.Minv + e);
        __syncthreads();
        TRITD_STAMP(1, 8)
        if (warp < rows) {
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                const int k = lane + 32 * q;
                if (k < RS) {
                    double v = 0.0;
                    if (k < R) {
                        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                        const double* rr = rhs_s + warp * 64;
                        int m = 0;
                        for (; m + 3 < R; m += 4) {
                            s0 = fma(rr[m], Ms[m * RS + k], s0);
                            s1 = fma(rr[m + 1], Ms[(m + 1) * RS + k], s1);
                            s2 = fma(rr[m + 2], Ms[(m + 2) * RS + k], s2);
                            s3 = fma(rr[m + 3], Ms[(m + 3) * RS + k], s3);
                        }
                        for (; m < R; ++m) s0 = fma(rr[m], Ms[m * RS + k], s0);
                        v = (s0 + s1) + (s2 + s3);
                    }
                    if (row0 + warp < a.n) {
                        a.X[(size_t)(row0 + warp) * RS + k] = v;
                        if (a.XT) a.XT[(size_t)k * a.ldt + row0 + warp] = v;
                    }
                }
            }
        }
        TRITD_STAMP(1, 9)
        __syncthreads();
        TRITD_STAMP(1, 10)
        if (tid == 0) red_release_add_u32(&a.flags[1], 1u);
        TRITD_STAMP(1, 7)
    }

    // ---------------- S = X'X once every row is written ----------------
    // Slices of ES entries x one row range, one per CTA: the 256 threads are RG row groups x ES entries (4 x 64 for
    // R <= 32, so a 64-row chunk costs 16 MACs per thread); the groups are combined in group order through shared memory and the
    // gr row-range partials by the last slice to arrive, in slice order (deterministic either way).
    const int RR = R * R;
    const int ES = RR <= 1024 ? 64 : 256, RG = kUpdThreads / ES;       // entries per slice x row groups = 256 threads
    const int G = (RR + ES - 1) / ES;
    const int nslice = G * a.gr;
    // Only the first `npart` CTAs take part (and spin): the row CTAs behind them must be able to become resident,
    // so the number of waiting CTAs stays well below the number of CTA slots of the GPU.
    const int npart = min(min(nslice, (int)gridDim.x), a.gram_cap);
    if ((int)blockIdx.x >= npart) return;
    cta_wait_eq(&a.flags[1], (unsigned)nrowcta, &a.st->status);
    TRITD_STAMP(0, 2)
    const int el = tid % ES, grp = tid / ES, rpg = 64 / RG;
    __shared__ int s_lastslice;
    for (int sl = blockIdx.x; sl < nslice; sl += npart) {
        const int es = sl % G, rs = sl / G;
        const int r0 = (int)((long)a.n * rs / a.gr), r1 = (int)((long)a.n * (rs + 1) / a.gr);
        const int e = es * ES + el;
        const bool ok = e < RR;
        const int aa = ok ? e / R : 0, bb = ok ? e - aa * R : 0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int i0 = r0; i0 < r1; i0 += 64) {
            __syncthreads();
            const int nr = min(64, r1 - i0);
            const int nq = nr * RS;
            const double* src = a.X + (size_t)i0 * RS;
            for (int q0 = 0; q0 < 64 * RS; q0 += 8 * kUpdThreads) {      // 8 loads in flight per thread
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int q = q0 + u * kUpdThreads + tid; t[u] = q < nq ? __ldcg(src + q) : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int q = q0 + u * kUpdThreads + tid; if (q < 64 * RS) Ms[q] = t[u]; }
            }
            __syncthreads();
            TRITD_STAMP(0, 11)
            const double* m = Ms + grp * rpg * RS;                       // rows rpg*grp .. rpg*(grp+1)-1 of the chunk
#pragma unroll 4
            for (int i = 0; i < rpg; i += 4) {
                s0 = fma(m[i * RS + aa], m[i * RS + bb], s0);
                s1 = fma(m[(i + 1) * RS + aa], m[(i + 1) * RS + bb], s1);
                s2 = fma(m[(i + 2) * RS + aa], m[(i + 2) * RS + bb], s2);
                s3 = fma(m[(i + 3) * RS + aa], m[(i + 3) * RS + bb], s3);
            }
        }
        __syncthreads();
        red[grp * ES + el] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (grp == 0 && ok) {
            double v = red[el];
            for (int g2 = 1; g2 < RG; ++g2) v += red[g2 * ES + el];
            if (a.gr == 1) a.gram_out[aa * RS + bb] = v;
            else a.gram_part[(size_t)rs * RS * RS + aa * RS + bb] = v;
        }
        if (a.gr > 1) {
            // the last row slice to arrive for this entry slice sums the gr partials in slice order
            __syncthreads();
            if (tid == 0) s_lastslice = atom_acq_rel_add_u32(&a.gram_cnt[es], 1u) == (unsigned)a.gr - 1;
            __syncthreads();
            if (s_lastslice) {
                if (grp == 0 && ok) {
                    double t[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = u < a.gr ? __ldcg(a.gram_part + (size_t)u * RS * RS + aa * RS + bb) : 0.0;
                    double v = t[0];
#pragma unroll
                    for (int u = 1; u < 8; ++u) v += t[u];
                    a.gram_out[aa * RS + bb] = v;
                }
                if (tid == 0) a.gram_cnt[es] = 0u;
            }
        }
    }
    __syncthreads();
    TRITD_STAMP(0, 12)
    __shared__ int s_lastgram;
    if (tid == 0) {
        s_lastgram = atom_acq_rel_add_u32(&a.flags[2], 1u) == (unsigned)npart - 1;
        if (s_lastgram) { a.flags[0] = 0u; a.flags[1] = 0u; a.flags[2] = 0u; a.flags[3] = 0u; }  // last Gram CTA: leave the flags zero for the next launch
    }
    if (a.sc_push) {
        // update C, N>1: the finished local C3'C3 goes to every rank's mailbox now (parity k&1, epoch xbase + k + 1);
        // the next iteration's updates A and B find it there
        __syncthreads();
        if (s_lastgram) {
            const int k = a.st->k;
            const long dst_off = a.sc_off + (long)(k & 1) * a.s2par_stride + (long)a.rank * RS * RS;
            for (int r = 0; r < a.nranks; ++r) {
                double* dst = a.peers[(a.rank + 1 + r) % a.nranks] + dst_off;
                for (int e = tid; e < RS * RS; e += kUpdThreads) dst[e] = __ldcg(a.gram_out + e);
            }
            __syncthreads();
            if (tid < a.nranks)
                st_release_sys_u32(reinterpret_cast<unsigned*>(a.peers[tid] + a.flag_area_off) + a.sc_flag_off + (k & 1) * 8 + a.rank,
                                   a.xbase + (unsigned)k + 1u);
        }
    }
    TRITD_STAMP(0, 3)
#undef TRITD_STAMP
}

}  // namespace tritd
