// Standalone L2 helpers of the reference (API completeness; the solver itself never
// materialises these): unfold.m:1-14, buildF/G/H.m:17-21, soft_threshold.m:2.
// All are single-pass, coalesced, HBM-bound kernels.
#pragma once
#include "common.cuh"

namespace tritd {

// Batched 2-D transpose through a padded 32x32 shared tile:
//   out[b][c][r] = in[b][r][c]   (r fastest on input, c fastest on output)
// unfold(X,2): rows = n1, cols = n2, batch = n3;  unfold(X,3): rows = n1*n2, cols = n3, batch = 1.
__global__ void __launch_bounds__(256) k_transpose(const double* in, double* out, long rows, long cols) {
    __shared__ double tile[32][33];
    const size_t boff = (size_t)blockIdx.z * rows * cols;
    const long r0 = (long)blockIdx.x * 32, c0 = (long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const long r = r0 + tx, c = c0 + k;
        if (r < rows && c < cols) tile[k][tx] = in[boff + (size_t)c * rows + r];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const long c = c0 + tx, r = r0 + k;
        if (r < rows && c < cols) out[boff + (size_t)r * cols + c] = tile[tx][k];
    }
}

// Transposed Khatri-Rao product, the common form of buildF/G/H:
//   out[k + R*(a + na*b)] = Xa[a][k] * Xb[b][k]      (Xa: na x RS, Xb: nb x RS row-major)
// buildF: (Xa,Xb) = (B2,C3); buildG: (A1,C3); buildH: (A1,B2).
__global__ void __launch_bounds__(256) k_khatri_rao_t(const double* Xa, const double* Xb, double* out, long na,
                                                      long nb, int R, int RS) {
    const size_t total = (size_t)R * na * nb;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const int k = (int)(e % R);
        const size_t col = e / R;
        const long aa = (long)(col % na), bb = (long)(col / na);
        out[e] = Xa[(size_t)aa * RS + k] * Xb[(size_t)bb * RS + k];
    }
}

__global__ void __launch_bounds__(256) k_soft_threshold(const double* x, double* out, size_t n, double lam) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const double v = x[i];
        const double mx = fmax(fabs(v) - lam, 0.0);
        out[i] = v > 0.0 ? mx : (v < 0.0 ? -mx : 0.0);
    }
}

// Design matrices of the ORIGINAL (Qi) triple decomposition, origin_triple_tensor/buildF.m:4-6, buildG.m:9-11,
// buildH.m:9-11 -- one summed index per entry.  Inputs are the MATLAB 3-D factor arrays
// (column-major: A(i,q,s) at i + n1*(q + r*s), B(p,j,s) at p + r*(j + n2*s), C(p,q,t) at p + r*(q + r*t)):
//   which = 0:  F(q + r*s, j + n2*t) = sum_p B(p,j,s) * C(p,q,t)        na = n2, nb = n3
//   which = 1:  G(p + r*s, i + n1*t) = sum_q A(i,q,s) * C(p,q,t)        na = n1, nb = n3
//   which = 2:  H(p + r*q, i + n1*j) = sum_s A(i,q,s) * B(p,j,s)        na = n1, nb = n2
// One thread per output entry, entries of a column (r^2 consecutive doubles) written by consecutive threads.
__global__ void __launch_bounds__(256) k_design_qi(const double* __restrict__ U, const double* __restrict__ V, double* __restrict__ out,
                                                   long na, long nb, int r, int which) {
    const int R = r * r;
    const size_t total = (size_t)R * na * nb;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const int k = (int)(e % R), k0 = k % r, k1 = k / r;
        const size_t col = e / R;
        const long a = (long)(col % na), b = (long)(col / na);
        double s = 0.0;
        if (which == 0) {           // U = B (r x n2 x r), V = C; k0 = q, k1 = s; a = j, b = t
            for (int p = 0; p < r; ++p) s = fma(U[p + (size_t)r * (a + na * k1)], V[p + (size_t)r * (k0 + (size_t)r * b)], s);
        } else if (which == 1) {    // U = A (n1 x r x r), V = C; k0 = p, k1 = s; a = i, b = t
            for (int q = 0; q < r; ++q) s = fma(U[a + na * (q + (size_t)r * k1)], V[k0 + (size_t)r * (q + (size_t)r * b)], s);
        } else {                    // U = A, V = B (r x n2 x r); k0 = p, k1 = q; a = i, b = j
            for (int t = 0; t < r; ++t) s = fma(U[a + na * (k1 + (size_t)r * t)], V[k0 + (size_t)r * (b + nb * t)], s);
        }
        out[e] = s;
    }
}

// Xhat = A_(1) * F for a materialised r^2 x ncols design matrix (the Qi triple product
// X(i,j,t) = sum_{p,q,s} A(i,q,s) B(p,j,s) C(p,q,t), origin_triple_tensor/triple_product.m): thread per (i, col),
// i fastest, so the loads of A(:,k) and the stores are coalesced and F(k,col) is a broadcast.
__global__ void __launch_bounds__(256) k_unfold1_times(const double* __restrict__ A, const double* __restrict__ F,
                                                       double* __restrict__ X, long n1, size_t ncols, int R) {
    const size_t total = (size_t)n1 * ncols;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const long i = (long)(e % n1);
        const size_t col = e / n1;
        double s = 0.0;
        for (int k = 0; k < R; ++k) s = fma(A[i + n1 * (size_t)k], F[col * R + k], s);
        X[e] = s;
    }
}

// evaluate() of the reference's drivers (traffic_triple_comparison.m:194-202): per-CTA partials of
//   sum_{mask} (Xhat - gt)^2  and  sum_{mask} gt^2   over padded column-major arrays (ld1 rows per column);
// mask is dense n1 x n2 x n3 bytes (nullptr = all true).  Fixed-order partial sums (deterministic).
__global__ void __launch_bounds__(256) k_evaluate(const double* __restrict__ Xhat, const double* __restrict__ gt,
                                                  const unsigned char* __restrict__ mask, int n1, int ld1, size_t ncols,
                                                  double* part) {
    __shared__ double red[64];
    double a = 0.0, b = 0.0;
    for (size_t col = blockIdx.x; col < ncols; col += gridDim.x)
        for (int i = threadIdx.x; i < n1; i += 256) {
            if (mask && !mask[col * (size_t)n1 + i]) continue;
            const double g = gt[col * (size_t)ld1 + i], d = Xhat[col * (size_t)ld1 + i] - g;
            a = fma(d, d, a);
            b = fma(g, g, b);
        }
    block_sum2(a, b, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b; }
}

}  // namespace tritd
