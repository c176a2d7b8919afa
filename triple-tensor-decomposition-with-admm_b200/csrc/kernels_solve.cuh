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

// X[row][:] = RHS[row][:] * inv(S1 o S2 + alpha*I) for 64 rows per CTA.
// Every CTA factors the R x R system itself (R <= 64: a few microseconds, no extra launch
// on the critical path), then each of 64 threads runs forward/back substitution on its row.
// Optionally also writes the transposed factor XT[k][row] (leading dimension ldt) that
// k_ppass streams with TMA.
struct SolveArgs {
    const double* rhs;     // [n][RS]
    const double *S1, *S2; // [RS][RS]
    double alpha;
    double* X;             // [n][RS]
    double* XT;            // [RS][ldt] or nullptr
    IterState* st;
    int n, R, RS, ldt;
};

__global__ void __launch_bounds__(256) k_solve(const SolveArgs a) {
    if (a.st->stop) return;
    extern __shared__ double sm[];
    const int R = a.R, P = R | 1;   // odd pitch: 64-bit row accesses of a half-warp hit 16 distinct bank pairs
    double* Lm = sm;              // [R][P]  lower Cholesky factor
    double* rows = sm + R * P;    // [64][P]
    const int tid = threadIdx.x;

    for (int e = tid; e < R * R; e += 256) {
        const int i = e / R, j = e - i * R;
        double v = a.S1[i * a.RS + j] * a.S2[i * a.RS + j];
        if (i == j) v += a.alpha;
        Lm[i * P + j] = v;
    }
    const int row0 = blockIdx.x * 64;
    for (int e = tid; e < 64 * R; e += 256) {
        const int rr = e / R, k = e - rr * R;
        rows[rr * P + k] = (row0 + rr < a.n) ? a.rhs[(size_t)(row0 + rr) * a.RS + k] : 0.0;
    }
    __syncthreads();

    // right-looking Cholesky on the lower triangle
    for (int k = 0; k < R; ++k) {
        const double d = Lm[k * P + k];
        if (!(d > 0.0) || !isfinite(d)) {          // uniform across the CTA
            if (tid == 0) atomicExch(&a.st->status, kStatusCholesky);
            return;
        }
        const double sd = sqrt(d);
        __syncthreads();
        for (int i = k + tid; i < R; i += 256) Lm[i * P + k] = (i == k) ? sd : Lm[i * P + k] / sd;
        __syncthreads();
        const int m = R - k - 1;
        for (int e = tid; e < m * m; e += 256) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) Lm[i * P + j] -= Lm[i * P + k] * Lm[j * P + k];
        }
        __syncthreads();
    }

    if (tid < 64) {
        double* b = rows + tid * P;
        for (int k = 0; k < R; ++k) {              // L y = b
            double s = b[k];
            for (int m = 0; m < k; ++m) s -= Lm[k * P + m] * b[m];
            b[k] = s / Lm[k * P + k];
        }
        for (int k = R - 1; k >= 0; --k) {         // L' x = y
            double s = b[k];
            for (int m = k + 1; m < R; ++m) s -= Lm[m * P + k] * b[m];
            b[k] = s / Lm[k * P + k];
        }
    }
    __syncthreads();
    for (int e = tid; e < 64 * a.RS; e += 256) {
        const int rr = e / a.RS, k = e - rr * a.RS;
        if (row0 + rr < a.n) a.X[(size_t)(row0 + rr) * a.RS + k] = (k < R) ? rows[rr * P + k] : 0.0;
    }
    if (a.XT) {
        for (int e = tid; e < 64 * a.RS; e += 256) {
            const int k = e / 64, rr = e - k * 64;
            if (row0 + rr < a.n) a.XT[(size_t)k * a.ldt + row0 + rr] = (k < R) ? rows[rr * P + k] : 0.0;
        }
    }
}

}  // namespace tritd
