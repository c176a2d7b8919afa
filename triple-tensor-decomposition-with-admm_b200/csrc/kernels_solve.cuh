// Ridge solves of the factor updates (reference: triple_decomp_ADMM.m:77-78, :86, :93):
//     X = RHS * pinv(M*M' + alpha*I),   M in {F, G, H}.
// Gram matrices come from the Hadamard identity  F*F' = (B2'B2) o (C3'C3),
// G*G' = (A1'A1) o (C3'C3), H*H' = (A1'A1) o (B2'B2)  (SURVEY.md fact 1), i.e. from the
// R x R "small Grams" of the factor matrices; the ridge makes them SPD, so the
// pseudo-inverse is the inverse and a Cholesky solve agrees with the SVD-based pinv
// to cond*eps.  A non-positive / non-finite pivot is reported through IterState::status
// (no silent fallback).
#pragma once
#include "common.cuh"
#include "kernels_fused.cuh"

namespace tritd {

constexpr int kStatusCholesky = 1;

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

// X[row][:] = RHS[row][:] * inv(S1 o S2 + alpha*I), 32 rows per CTA of 1024 threads.
// Every CTA inverts the R x R ridge system itself by Gauss-Jordan elimination (SPD, so no pivoting;
// the pivots are the Cholesky pivots d_k = L_kk^2, and a non-positive or non-finite one is reported
// through IterState::status).  The matrix lives in registers, a 2 x 2 patch per thread; only the
// pivot row and column pass through shared memory, one barrier per step -- R <= 64 steps of a few
// hundred cycles, a few microseconds, no extra launch on the critical path.  The
// inverse is then applied to the CTA's rows like the reference applies pinv(G) (:78).  The CTA also
// forms the partial small Gram X'X of its rows; the last CTA to finish (ticket counter) sums the
// partials in CTA order, so S_out = X'X is deterministic and needs no separate kernel.  Optionally
// writes the transposed factor XT[k][row] that k_ppass streams with TMA.
struct SolveArgs {
    const double* rhs;     // [n][RS]
    const double *S1, *S2; // [RS][RS]
    double alpha;
    double* X;             // [n][RS]
    double* XT;            // [RS][ldt] or nullptr
    double* gram_part;     // [grid][R*R]
    double* gram_out;      // [RS][RS]  (= X'X over the rows of this rank)
    unsigned* ticket;
    IterState* st;
    long long* dbg;        // optional: clock64 stamps of CTA 0 (diagnostics)
    int n, R, RS, ldt;
};

constexpr int kSolveRows = 32;
constexpr int kSolveThreads = 1024;

template <int PQ>   // PQ x PQ register patch per thread: 1 for R <= 32, 2 for R <= 64
__global__ void __launch_bounds__(kSolveThreads) k_solve(const SolveArgs a) {
    if (a.st->stop) return;
    extern __shared__ double sm[];
    const int R = a.R, P = R | 1, RR = R * R;   // odd pitch
    double* Gb = sm;                       // [R][P]     inv(G) (written once, after the elimination)
    double* prow = Gb + R * P;             // [2][64]    published pivot row    (double-buffered)
    double* pcol = prow + 128;             // [2][64]    published pivot column (double-buffered)
    double* rows = pcol + 128;             // [32][P]    RHS rows
    double* xr = rows + kSolveRows * P;    // [32][P]    solved rows
    __shared__ int s_last, s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    const int row0 = blockIdx.x * kSolveRows;
#define TRITD_STAMP(q) if (a.dbg && blockIdx.x == 0 && tid == 0) a.dbg[q] = clock64();
    TRITD_STAMP(0)

    // this thread's entries e = tid + m*1024 of an R x R (or 32 x R) index space
    int ei[4], ej[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int e = tid + m * kSolveThreads;
        ei[m] = e / R; ej[m] = e - ei[m] * R;
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
        if (tid + m * kSolveThreads < kSolveRows * R)
            rows[ei[m] * P + ej[m]] = (row0 + ei[m] < a.n) ? a.rhs[(size_t)(row0 + ei[m]) * a.RS + ej[m]] : 0.0;

    // In-place Gauss-Jordan with the matrix in REGISTERS: thread (ty,tx) of a 32 x 32 grid owns entries
    // (ty + 32p, tx + 32q), p,q in {0,1}.  Per step only the pivot row and pivot column go through shared
    // memory (double-buffered, published by their owners as they are produced), so a step is: one barrier,
    // five shared loads, one reciprocal, four FMAs.  A bad pivot only raises a flag (checked once).
    auto fast_rcp = [](double x) {
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
        e = fma(-x, y, 1.0);
        return fma(y, e, y);
    };
    const int ty = tid >> 5, tx = tid & 31;
    double gq[PQ][PQ];
#pragma unroll
    for (int p2 = 0; p2 < PQ; ++p2)
#pragma unroll
        for (int q2 = 0; q2 < PQ; ++q2) {
            const int i = ty + 32 * p2, j = tx + 32 * q2;
            double v = 0.0;
            if (i < R && j < R) {
                v = a.S1[i * a.RS + j] * a.S2[i * a.RS + j];
                if (i == j) v += a.alpha;
                if (i == 0) prow[j] = v;
                if (j == 0) pcol[i] = v;
            }
            gq[p2][q2] = v;
        }
    __syncthreads();
    TRITD_STAMP(1)
    bool bad = false;
    // only the warps that own rows take part in the elimination (named barrier over those warps)
    const int nwarp_act = PQ == 1 ? R : 32;
    if (ty < nwarp_act) {
        for (int k = 0; k < R; ++k) {
            const double* pr = prow + (k & 1) * 64;
            const double* pc = pcol + (k & 1) * 64;
            double* prn = prow + ((k + 1) & 1) * 64;
            double* pcn = pcol + ((k + 1) & 1) * 64;
            const double piv = pr[k];
            bad = bad || !(piv > 0.0) || !isfinite(piv);
            const double inv = fast_rcp(piv);
#pragma unroll
            for (int p2 = 0; p2 < PQ; ++p2)
#pragma unroll
                for (int q2 = 0; q2 < PQ; ++q2) {
                    const int i = ty + 32 * p2, j = tx + 32 * q2;
                    if (i < R && j < R) {
                        double v;
                        if (i == k) v = (j == k) ? inv : pr[j] * inv;
                        else if (j == k) v = -pc[i] * inv;
                        else v = fma(-pc[i], pr[j] * inv, gq[p2][q2]);
                        gq[p2][q2] = v;
                        if (i == k + 1) prn[j] = v;
                        if (j == k + 1) pcn[i] = v;
                    }
                }
            asm volatile("bar.sync 1, %0;" ::"r"(nwarp_act * 32));
        }
#pragma unroll
        for (int p2 = 0; p2 < PQ; ++p2)
#pragma unroll
            for (int q2 = 0; q2 < PQ; ++q2) {
                const int i = ty + 32 * p2, j = tx + 32 * q2;
                if (i < R && j < R) Gb[i * P + j] = gq[p2][q2];
            }
        if (bad && tid == 0) s_bad = 1;
    }
    __syncthreads();
    bad = s_bad != 0;
    if (bad) {                                         // uniform (every thread of the elimination saw the same pivots)
        if (tid == 0) atomicExch(&a.st->status, kStatusCholesky);
        return;
    }
    const double* Gi = Gb;
    TRITD_STAMP(2)

    // apply: xr[rr][k] = sum_m rows[rr][m] * inv(G)[m][k]
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        if (tid + m * kSolveThreads < kSolveRows * R) {
            const int rr = ei[m], k = ej[m];
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int q = 0;
            for (; q + 3 < R; q += 4) {
                s0 = fma(rows[rr * P + q], Gi[q * P + k], s0);
                s1 = fma(rows[rr * P + q + 1], Gi[(q + 1) * P + k], s1);
                s2 = fma(rows[rr * P + q + 2], Gi[(q + 2) * P + k], s2);
                s3 = fma(rows[rr * P + q + 3], Gi[(q + 3) * P + k], s3);
            }
            for (; q < R; ++q) s0 = fma(rows[rr * P + q], Gi[q * P + k], s0);
            xr[rr * P + k] = (s0 + s1) + (s2 + s3);
        }
    }
    __syncthreads();
    TRITD_STAMP(3)
    for (int e = tid; e < kSolveRows * a.RS; e += kSolveThreads) {
        const int rr = e / a.RS, k = e - rr * a.RS;
        if (row0 + rr < a.n) a.X[(size_t)(row0 + rr) * a.RS + k] = (k < R) ? xr[rr * P + k] : 0.0;
    }
    if (a.XT) {
        for (int e = tid; e < kSolveRows * a.RS; e += kSolveThreads) {
            const int k = e / kSolveRows, rr = e - k * kSolveRows;
            if (row0 + rr < a.n) a.XT[(size_t)k * a.ldt + row0 + rr] = (k < R) ? xr[rr * P + k] : 0.0;
        }
    }
    TRITD_STAMP(4)
    // partial small Gram of this CTA's rows (rows beyond n are zero)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (tid + m * kSolveThreads < RR) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int rr = 0; rr < kSolveRows; rr += 4) {
                s0 = fma(xr[rr * P + ei[m]], xr[rr * P + ej[m]], s0);
                s1 = fma(xr[(rr + 1) * P + ei[m]], xr[(rr + 1) * P + ej[m]], s1);
                s2 = fma(xr[(rr + 2) * P + ei[m]], xr[(rr + 2) * P + ej[m]], s2);
                s3 = fma(xr[(rr + 3) * P + ei[m]], xr[(rr + 3) * P + ej[m]], s3);
            }
            a.gram_part[(size_t)blockIdx.x * RR + tid + m * kSolveThreads] = (s0 + s1) + (s2 + s3);
        }
    }
    TRITD_STAMP(5)
    __threadfence();
    __syncthreads();
    TRITD_STAMP(6)
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int e = tid + m * kSolveThreads;
            if (e < RR) {
                double s0 = 0.0, s1 = 0.0;
                unsigned c = 0;
                for (; c + 1 < gridDim.x; c += 2) {
                    s0 += a.gram_part[(size_t)c * RR + e];
                    s1 += a.gram_part[(size_t)(c + 1) * RR + e];
                }
                if (c < gridDim.x) s0 += a.gram_part[(size_t)c * RR + e];
                a.gram_out[ei[m] * a.RS + ej[m]] = s0 + s1;
            }
        }
        if (tid == 0) *a.ticket = 0u;
    }
    TRITD_STAMP(7)
#undef TRITD_STAMP
}

}  // namespace tritd
