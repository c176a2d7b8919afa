// Iteration scalars (IterState), the scalar step that ends an iteration, and k_fused: the reconstruction
// L = triple_product(A,B,C) formed on the fly with DMMA (fast_robust_triple_tensor/triple_product.m:6-7), either
// stored (MODE 1: tritd_triple_product_f64, L output, evaluate) or only compared with a tensor X
// (MODE 2: sum((X - Xhat).^2) of triple_decomp_ALS.m:15-16).  The ADMM element-wise block itself lives in
// kernels_admm.cuh (k_admm).
#pragma once
#include "common.cuh"

namespace tritd {

// Iteration scalars, kept on the device so a whole iteration replays without the host.
struct IterState {
    double muL, muO, muL_max, muO_max, rhoL, rhoO, lambda, tol, normD;
    double rmuL, rmuO, thr, musum, rmuL_next;   // (1/muL), (1/muO), lambda/muO, muL+muO, 1/muL of the next iteration
    int k;         // iterations completed
    int stop;      // stopping rule fired or maxIter reached: all later launches are no-ops
    int status;    // != 0: numerical failure (NaN / Inf in a ridge system), see tritd.h
    int maxIter;
    int pinv_fallbacks;   // ridge solves that went through the truncating pseudo-inverse (pinv_jacobi) ...
    int pinv_truncated;   // ... and the number of singular values they zeroed, like MATLAB's pinv
    int masked;           // completion variant: NaN in D marks an unobserved entry
    int pad_;
    double muO_prev, thr_prev;   // muO and lambda/muO of the iteration that produced the stored Z = R3 (k_admm)
};

__host__ __device__ inline void iter_state_derive(IterState& s) {
    s.rmuL = 1.0 / s.muL;
    s.rmuO = 1.0 / s.muO;
    s.thr = s.lambda / s.muO;
    s.musum = s.muL + s.muO;
    const double nL = s.muL * s.rhoL;
    s.rmuL_next = 1.0 / (nL < s.muL_max ? nL : s.muL_max);
}

struct FusedArgs {
    const double* D;                   // MODE 2: the tensor X compared with the reconstruction
    double* O;                         // MODE 1: output L
    const double *A1, *B2, *C3;        // [n][RS]
    const IterState* st;
    double* norm_part;                 // [grid][2]
    int n1, n2, n3, ld1, RS;
    int n_it, n_jc, gi;                // i-tiles (128), j-chunks (32), CTAs per i-tile
};

template <int KS> struct FusedCfg {
    static constexpr int PL = (KS & 1) ? 4 * KS : 4 * KS + 4;   // pitch of the B2 chunk: == 4 (mod 8) doubles
    static constexpr int kMinBlocks = KS <= 8 ? 2 : 1;
};

// CTA = 8 warps x 16 rows i (i-tile of 128).  CTA c owns i-tile c % n_it and a contiguous
// range of column blocks v = jc * n3 + t (32 columns j of slice t).  Each warp walks the
// block in four groups of 8 columns; per group it forms a 16 x 8 patch of L with 2*KS DMMAs
// (M = i, two m-tiles = even/odd i; N = j; K = k) whose accumulator layout is exactly the
// 16-byte-vector layout of the column-major tensor, so it is touched with 128-byte-per-row coalesced v2
// accesses straight from/to registers.
template <int KS, int MODE>
__global__ void __launch_bounds__(256, FusedCfg<KS>::kMinBlocks) k_fused(const FusedArgs a) {
    constexpr int PL = FusedCfg<KS>::PL;
    if (MODE != 1 && a.st->stop) return;
    __shared__ double B2s[32 * PL];
    __shared__ double red[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int it = blockIdx.x % a.n_it, x = blockIdx.x / a.n_it;
    const long V = (long)a.n_jc * a.n3;
    const long v0 = V * x / a.gi, v1 = V * (x + 1) / a.gi;
    const int i0 = it * 128 + warp * 16 + 2 * g;          // this lane's even row; it also owns i0 + 1
    const bool i_ok = i0 < a.ld1;

    // rows of A1 this lane feeds into the A fragments (clamped; rows >= n1 are zeroed below)
    const double* a1r0 = a.A1 + (size_t)min(i0, a.n1 - 1) * a.RS + tig;
    const double* a1r1 = a.A1 + (size_t)min(i0 + 1, a.n1 - 1) * a.RS + tig;
    const double z0 = (i0 < a.n1) ? 1.0 : 0.0, z1 = (i0 + 1 < a.n1) ? 1.0 : 0.0;

    struct Buf { double2 d[2]; };
    auto load = [&](Buf& b, long v, int jg) {
        const int jc = (int)(v / a.n3), t = (int)(v - (long)jc * a.n3);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int j = jc * 32 + jg * 8 + 2 * tig + c;
            const bool ok = i_ok && j < a.n2;
            const size_t off = ((size_t)t * a.n2 + j) * a.ld1 + i0;
            if (MODE != 1) b.d[c] = ok ? ldg_stream2(a.D + off) : make_double2(0.0, 0.0);
        }
    };

    double sL = 0.0, sO = 0.0;
    int cur_jc = -1;
    Buf cur, nxt;
    if (MODE != 1 && v0 < v1) load(cur, v0, 0);

    for (long v = v0; v < v1; ++v) {
        const int jc = (int)(v / a.n3), t = (int)(v - (long)jc * a.n3);
        if (jc != cur_jc) {
            __syncthreads();
            for (int e = threadIdx.x; e < 32 * 4 * KS; e += 256) {
                const int j = e / (4 * KS), k = e - j * (4 * KS);
                const int jj = jc * 32 + j;
                B2s[j * PL + k] = (jj < a.n2) ? a.B2[(size_t)jj * a.RS + k] : 0.0;
            }
            __syncthreads();
            cur_jc = jc;
        }
        // A fragments with C3[t,:] folded in: aS[m][s] = A1(i0+m, 4s+tig) * C3(t, 4s+tig)  (L1/L2-resident factors)
        double aS[2][KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const double c3 = __ldg(a.C3 + (size_t)t * a.RS + 4 * s + tig);
            aS[0][s] = __ldg(a1r0 + 4 * s) * c3 * z0;
            aS[1][s] = __ldg(a1r1 + 4 * s) * c3 * z1;
        }

#pragma unroll
        for (int jg = 0; jg < 4; ++jg) {
            if (MODE != 1) {
                if (jg < 3) load(nxt, v, jg + 1);
                else if (v + 1 < v1) load(nxt, v + 1, 0);
            }
            double l[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double b = B2s[(jg * 8 + g) * PL + 4 * s + tig];
                dmma884(l[0][0], l[0][1], aS[0][s], b);
                dmma884(l[1][0], l[1][1], aS[1][s], b);
            }
            // l[m][c] = L(i0 + m, jc*32 + jg*8 + 2*tig + c, t)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int j = jc * 32 + jg * 8 + 2 * tig + c;
                if (!(i_ok && j < a.n2)) continue;
                const size_t off = ((size_t)t * a.n2 + j) * a.ld1 + i0;
                if (MODE == 1) {
                    stg_stream2(a.O + off, make_double2(l[0][c], l[1][c]));
                } else if (MODE == 2) {                 // sum((X - Xhat).^2) only (triple_decomp_ALS.m:15-16)
                    const double e0 = __dsub_rn(cur.d[c].x, l[0][c]), e1 = __dsub_rn(cur.d[c].y, l[1][c]);
                    sL = fma(e0, e0, sL);
                    sL = fma(e1, e1, sL);
                }
            }
            if (MODE != 1) cur = nxt;
        }
    }
    if (MODE != 1) {
        block_sum2(sL, sO, red);
        if (threadIdx.x == 0) { a.norm_part[2 * blockIdx.x] = sL; a.norm_part[2 * blockIdx.x + 1] = sO; }
    }
}

// Fixed-order sum of `n` pairs (per-CTA partials) into out[0..1]; one CTA.
__global__ void __launch_bounds__(256) k_sum_pairs(const double* part, int n, double* out, const int* stop) {
    if (stop && *stop) return;
    __shared__ double red[64];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { a += part[2 * i]; b += part[2 * i + 1]; }
    block_sum2(a, b, red);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

// Per-CTA partial of sum(x^2) over a padded N-array (pad entries are zero).
// (masked: NaN marks an unobserved entry, which does not count -- the norm of the observed data)
__global__ void __launch_bounds__(256) k_sumsq_part(const double* x, size_t n, double* part, int masked) {
    __shared__ double red[64];
    double s = 0.0, z = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const double v = x[i];
        if (!(masked && v != v)) s = fma(v, v, s);
    }
    block_sum2(s, z, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = s; part[2 * blockIdx.x + 1] = 0.0; }
}

// The scalar step that ends an iteration (triple_decomp_ADMM.m:56-65): mu schedule BEFORE errHist, the
// relative-change stopping rule, maxIter.  a, b = sum(resL.^2), sum(resO.^2) over the whole tensor.
__device__ __forceinline__ void iter_finalize(IterState* st, double a, double b, double* errHist, double* errL, double* errO) {
    const int k = st->k;
    st->muO_prev = st->muO; st->thr_prev = st->thr;
    st->muL = fmin(st->muL * st->rhoL, st->muL_max);
    st->muO = fmin(st->muO * st->rhoO, st->muO_max);
    iter_state_derive(*st);
    const double eL = sqrt(a) / st->normD, eO = sqrt(b) / st->normD;
    errL[k] = eL; errO[k] = eO; errHist[k] = eL + eO;
    st->k = k + 1;
    if (k >= 1 && fabs(errHist[k] - errHist[k - 1]) < st->tol * errHist[k - 1]) st->stop = 1;
    if (k + 1 >= st->maxIter) st->stop = 1;
    if (st->status != 0) st->stop = 1;     // a factor update reported a bad pivot: later launches become no-ops
}

// errHist / mu schedule / stopping rule (triple_decomp_ADMM.m:56-65).  One CTA: first the fixed-order
// sum of the per-CTA partials of sum(resL^2), sum(resO^2) (npart pairs; with `reduced` set the pair
// in norms[] was already summed -- and all-reduced over the ranks), then one thread does the scalars.
__global__ void __launch_bounds__(256) k_finalize(IterState* st, const double* part, int npart, double* norms, int reduced,
                                                  double* errHist, double* errL, double* errO) {
    if (st->stop) return;
    __shared__ double red[64];
    double a = 0.0, b = 0.0;
    if (!reduced) {
        for (int i = threadIdx.x; i < npart; i += 256) { a += part[2 * i]; b += part[2 * i + 1]; }
        block_sum2(a, b, red);
    }
    if (threadIdx.x != 0) return;
    if (reduced) { a = norms[0]; b = norms[1]; }
    iter_finalize(st, a, b, errHist, errL, errO);
}

// ALS (triple_decomp_ALS.m:14-22): errHist(k) = ||X - Xhat|| / ||X|| with the factors BEFORE the updates of
// iteration k, then the relative-change stopping rule; when it fires the update kernels of this iteration see
// the stop flag and do nothing, exactly like the reference's `break` before the updates.
__global__ void __launch_bounds__(256) k_finalize_als(IterState* st, const double* part, int npart, double* errHist) {
    if (st->stop) return;
    __shared__ double red[64];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < npart; i += 256) a += part[2 * i];
    block_sum2(a, b, red);
    if (threadIdx.x != 0) return;
    const int k = st->k;
    errHist[k] = sqrt(a) / st->normD;
    st->k = k + 1;
    if (k >= 1 && fabs(errHist[k] - errHist[k - 1]) < st->tol * errHist[k - 1]) st->stop = 1;
    if (st->status != 0) st->stop = 1;
}

// E of the last finished iteration from the state the iteration keeps: Z = R3 of that iteration, E =
// soft_threshold(R3, lambda/muO) with that iteration's muO (triple_decomp_ADMM.m:46-47; see k_admm).
__global__ void __launch_bounds__(256) k_E_from_Z(const double* __restrict__ Z, const IterState* st, double* __restrict__ E, size_t n) {
    const double thr = st->thr_prev;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const double z = Z[i];
        E[i] = z - fmin(fmax(z, -thr), thr);          // = soft_threshold(z, thr), as k_admm forms it
    }
}

// O of the last finished iteration, recovered from the state the iteration keeps:
// k_admm forms T' = (D - O) + (1/muL')*Y_L' and never stores O (nothing inside the loop reads it,
// triple_decomp_ADMM.m:41-43), so O = D - (T' - (1/muL')*Y_L') with the same product as the forward pass;
// exact up to one rounding of the sum, i.e. ~1 ulp of max(|D|,|Y_L/muL|) per entry.
__global__ void __launch_bounds__(256) k_recover_O(const double* __restrict__ D, const double* __restrict__ T,
                                                   const double* __restrict__ YL, const IterState* st,
                                                   double* __restrict__ O, size_t n) {
    const double rmu = st->rmuL;
    const bool masked = st->masked != 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const double d = D[i];
        O[i] = (masked && d != d) ? 0.0 : __dsub_rn(d, __dsub_rn(T[i], __dmul_rn(rmu, YL[i])));
    }
}

// Completion variant: D(i) <- NaN where mask(i) == 0 (dense n1 x ncols byte mask, non-zero = observed), padded D.
__global__ void __launch_bounds__(256) k_apply_mask(double* D, const unsigned char* mask, int n1, int ld1, size_t ncols) {
    const double nan_ = __longlong_as_double(0x7ff8000000000000LL);
    const size_t total = (size_t)n1 * ncols;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const size_t col = e / n1;
        const int i = (int)(e - col * n1);
        if (!mask[e]) D[col * ld1 + i] = nan_;
    }
}
// first target of the completion variant: T = D on the observed entries, 0 on the unobserved ones
__global__ void __launch_bounds__(256) k_fill_unobserved(const double* __restrict__ D, double* __restrict__ T, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const double d = D[i];
        T[i] = d != d ? 0.0 : d;
    }
}

__global__ void k_set_normD(IterState* st, const double* sumsq) {
    if (threadIdx.x == 0 && blockIdx.x == 0) st->normD = sqrt(sumsq[0]);
}

}  // namespace tritd
