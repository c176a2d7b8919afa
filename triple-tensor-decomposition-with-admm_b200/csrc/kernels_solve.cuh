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

// X[row][:] = RHS[row][:] * inv(S1 o S2 + alpha*I), 32 rows per CTA, 256 threads.
// Every CTA inverts the R x R ridge system itself by Gauss-Jordan elimination on [G | I] in shared
// memory (SPD, so no pivoting; pivots are the Cholesky pivots d_k = L_kk^2 and a non-positive or
// non-finite one is reported through IterState::status) -- R <= 64, a few microseconds, and no
// extra launch on the critical path -- and then applies it to its rows like the reference applies
// pinv(G) (:78).  The CTA also forms the partial small Gram X'X of its rows; the last CTA to finish
// (ticket counter) sums the partials in CTA order, so S_out = X'X is deterministic and needs no
// separate kernel.  Optionally writes the transposed factor XT[k][row] that k_ppass streams with TMA.
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
    int n, R, RS, ldt;
};

constexpr int kSolveRows = 32;

__global__ void __launch_bounds__(256) k_solve(const SolveArgs a) {
    if (a.st->stop) return;
    extern __shared__ double sm[];
    const int R = a.R, W2 = 2 * R, PW = W2 | 1, P = R | 1;
    double* Wm = sm;                       // [R][PW]   [G | I] -> [I | inv(G)]
    double* rowk = Wm + R * PW;            // [2R]
    double* colk = rowk + W2;              // [R]
    double* rows = colk + R;               // [32][P]   RHS rows
    double* xr = rows + kSolveRows * P;    // [32][P]   solved rows
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * kSolveRows;

    for (int e = tid; e < R * W2; e += 256) {
        const int i = e / W2, j = e - i * W2;
        double v;
        if (j < R) {
            v = a.S1[i * a.RS + j] * a.S2[i * a.RS + j];
            if (i == j) v += a.alpha;
        } else {
            v = (j - R == i) ? 1.0 : 0.0;
        }
        Wm[i * PW + j] = v;
    }
    for (int e = tid; e < kSolveRows * R; e += 256) {
        const int rr = e / R, k = e - rr * R;
        rows[rr * P + k] = (row0 + rr < a.n) ? a.rhs[(size_t)(row0 + rr) * a.RS + k] : 0.0;
    }
    __syncthreads();

    for (int k = 0; k < R; ++k) {
        const double piv = Wm[k * PW + k];
        if (!(piv > 0.0) || !isfinite(piv)) {          // uniform across the CTA
            if (tid == 0) atomicExch(&a.st->status, kStatusCholesky);
            return;
        }
        const double inv = 1.0 / piv;
        if (tid < W2) rowk[tid] = Wm[k * PW + tid] * inv;
        else if (tid - W2 < R) colk[tid - W2] = Wm[(tid - W2) * PW + k];
        if (W2 + R > 256)                               // R = 64: 192 threads' worth of work, fits; kept general
            for (int q = 256 + tid; q < W2 + R; q += 256) colk[q - W2] = Wm[(q - W2) * PW + k];
        __syncthreads();
        for (int e = tid; e < R * W2; e += 256) {
            const int i = e / W2, j = e - i * W2;
            Wm[i * PW + j] = (i == k) ? rowk[j] : Wm[i * PW + j] - colk[i] * rowk[j];
        }
        __syncthreads();
    }

    // apply: xr[rr][k] = sum_m rows[rr][m] * inv(G)[m][k]
    for (int e = tid; e < kSolveRows * R; e += 256) {
        const int rr = e / R, k = e - rr * R;
        double s0 = 0.0, s1 = 0.0;
        int m = 0;
        for (; m + 1 < R; m += 2) {
            s0 = fma(rows[rr * P + m], Wm[m * PW + R + k], s0);
            s1 = fma(rows[rr * P + m + 1], Wm[(m + 1) * PW + R + k], s1);
        }
        if (m < R) s0 = fma(rows[rr * P + m], Wm[m * PW + R + k], s0);
        xr[rr * P + k] = s0 + s1;
    }
    __syncthreads();
    for (int e = tid; e < kSolveRows * a.RS; e += 256) {
        const int rr = e / a.RS, k = e - rr * a.RS;
        if (row0 + rr < a.n) a.X[(size_t)(row0 + rr) * a.RS + k] = (k < R) ? xr[rr * P + k] : 0.0;
    }
    if (a.XT) {
        for (int e = tid; e < kSolveRows * a.RS; e += 256) {
            const int k = e / kSolveRows, rr = e - k * kSolveRows;
            if (row0 + rr < a.n) a.XT[(size_t)k * a.ldt + row0 + rr] = (k < R) ? xr[rr * P + k] : 0.0;
        }
    }
    // partial small Gram of this CTA's rows (rows beyond n are zero)
    for (int e = tid; e < R * R; e += 256) {
        const int aa = e / R, bb = e - aa * R;
        double s = 0.0;
#pragma unroll 4
        for (int rr = 0; rr < kSolveRows; ++rr) s = fma(xr[rr * P + aa], xr[rr * P + bb], s);
        a.gram_part[(size_t)blockIdx.x * R * R + e] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int e = tid; e < R * R; e += 256) {
            double s = 0.0;
            for (unsigned c = 0; c < gridDim.x; ++c) s += a.gram_part[(size_t)c * R * R + e];
            const int aa = e / R, bb = e - aa * R;
            a.gram_out[aa * a.RS + bb] = s;
        }
        if (tid == 0) *a.ticket = 0u;
    }
}

}  // namespace tritd
