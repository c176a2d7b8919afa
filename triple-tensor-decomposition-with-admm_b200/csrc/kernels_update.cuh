// k_upd: one factor update of the Gauss-Seidel sweep as ONE kernel
// (reference: fast_robust_triple_tensor/triple_decomp_ADMM.m:73-95):
//     RHS = X_(k) * M'                       -- the last reduction step of the MTTKRP (sources below)
//     X   = RHS * pinv(M*M' + alpha*I)       -- :78 / :86 / :93, M*M' = S1 o S2 (Hadamard identity)
//     S   = X'X                              -- the small Gram the next two updates need
// CTA roles (256 threads each):
//   block 0      inverts the R x R ridge system (in-place Gauss-Jordan, SPD => no pivoting, the pivots are the
//                Cholesky pivots L_kk^2; a non-positive / non-finite pivot is reported through IterState::status)
//                WHILE the row CTAs reduce their RHS rows, publishes inv(G) and raises flags[0];
//   blocks 1..   8/WPR rows each: fixed-order reduction of the RHS rows from the source, wait for flags[0],
//                apply the inverse like the reference applies pinv(G), write X (and its transpose for TMA);
//   blocks 0..G-1 finally form S = X'X straight from the freshly written rows once every row CTA has
//                signalled flags[1] (one entry per thread, rows summed in a fixed order => deterministic).
// The inverse is off the critical path, nothing is launched between "RHS ready" and "factor ready", and no
// partial-Gram buffers exist.  Row CTAs wait only for block 0 (scheduled first, waits for nobody); the Gram
// phase waits for the row CTAs, which never wait for it: no cyclic dependency, at most G CTAs spin.
//
// RHS sources:
//   kSrcDirect  rhs[row][k] as given (first iteration; N>1 after the all-reduce)
//   kSrcPartF   sum over the k_admm CTAs of the row's i-tile of their mode-1 partials (update_A)
//   kSrcPB      sum_t C3[t][k] * P[t][row][k]                          (update_B, :86)
//   kSrcPC      sum_j B2[j][k] * P[row][j][k]                          (update_C, :93)
// With apply == 0 the kernel only reduces: rows go to rhs_out (N>1: the all-reduce comes next).
#pragma once
#include "common.cuh"
#include "kernels_fused.cuh"

namespace tritd {

constexpr int kStatusCholesky = 1;

enum UpdSrc { kSrcDirect = 0, kSrcPartF = 1, kSrcPB = 2, kSrcPC = 3 };

struct UpdArgs {
    const double* rhs;        // kSrcDirect: [n][RS]
    const double* part;       // kSrcPartF: [gridA][128][RS]; the CTAs of i-tile `it` are c = it + q*nit, q < part_count
    int part_count, nit, tile_h;
    const double* P;          // kSrcPB / kSrcPC: [n3][n2][RS]
    const double* W;          // kSrcPB: C3 [n3][RS]; kSrcPC: B2 [n2][RS]
    int n2, n3;
    const double *S1, *S2;    // small Grams of the two other factors, each a stack of ns1 / ns2 partial [RS][RS] matrices
    int ns1, ns2;
    double alpha;
    double* Minv;             // [R][RS] scratch: inv(S1 o S2 + alpha I), written by block 0
    double* rhs_out;          // apply == 0: reduced rows [n][RS]
    double* X;                // [n][RS]
    double* XT;               // [RS][ldt] or nullptr
    double* gram_out;         // [gr][RS][RS]: X'X over the rows of this rank as gr row-slice partials (consumers sum them)
    int gr;
    IterState* st;
    unsigned* flags;          // [0] inverse published, [1] row CTAs done, [2] Gram CTAs done; all zero between launches
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
__device__ __forceinline__ void cta_wait_eq(const unsigned* p, unsigned want) {
    if (threadIdx.x == 0)
        while (ld_acquire_u32(p) != want) __nanosleep(20);
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
// barrier, 2*PQ+1 shared loads, one reciprocal, PQ*PQ FMAs.  Returns true when a pivot was bad.
template <int PQ>
__device__ bool invert_ridge_system(const double* S1, int ns1, const double* S2, int ns2, double alpha, int R, int RS, double* out,
                                    double* sm /* >= 256 doubles */) {
    double* prow = sm;        // [2][64]
    double* pcol = sm + 128;  // [2][64]
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double gq[PQ][PQ];
#pragma unroll
    for (int p = 0; p < PQ; ++p)
#pragma unroll
        for (int q = 0; q < PQ; ++q) {
            const int i = ty + 16 * p, j = tx + 16 * q;
            double v = 0.0;
            if (i < R && j < R) {
                // partial stacks (<= 8 slices) are summed in slice order; all loads of an entry are issued together
                double t1[8], t2[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    t1[s] = s < ns1 ? S1[s * RS * RS + i * RS + j] : 0.0;
                    t2[s] = s < ns2 ? S2[s * RS * RS + i * RS + j] : 0.0;
                }
                double s1 = t1[0], s2 = t2[0];
#pragma unroll
                for (int s = 1; s < 8; ++s) { s1 += t1[s]; s2 += t2[s]; }
                v = s1 * s2;
                if (i == j) v += alpha;
                if (i == 0) prow[j] = v;
                if (j == 0) pcol[i] = v;
            }
            gq[p][q] = v;
        }
    __syncthreads();
    bool bad = false;
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
            bad = bad || !(piv > 0.0) || !isfinite(piv);
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
            __syncthreads();
        }
    }
#pragma unroll
    for (int p = 0; p < PQ; ++p)
#pragma unroll
        for (int q = 0; q < PQ; ++q) {
            const int i = ty + 16 * p, j = tx + 16 * q;
            if (i < R && j < R) out[i * RS + j] = gq[p][q];
        }
    return bad;
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
__host__ __device__ inline size_t upd_smem_bytes(int RS) {
    const int ms = RS * RS > 64 * RS ? RS * RS : 64 * RS;      // inv(G) tile / Gram row chunk [64][RS]
    return (size_t)(3 * 8 * 64 + ms) * sizeof(double);
}

template <int SRC, int WPR, int KPL>   // KPL = columns per lane: 1 (RS <= 32) or 2 (RS <= 64)
__global__ void __launch_bounds__(kUpdThreads) k_upd(const UpdArgs a) {
    if (a.st->stop) return;
    constexpr int ROWS = 8 / WPR;                 // rows per row CTA; WPR warps share a row
    constexpr int CH = 8;                         // independent accumulation chains per lane and column
    extern __shared__ double sm[];
    double* red = sm;                             // [8 warps][64]  (block 0: pivot row/column buffers)
    double* rhs_s = sm + 512;                     // [ROWS][64]
    double* xs = sm + 1024;                       // [ROWS][64]
    double* Ms = sm + 1536;                       // [R][RS] inverse, later the Gram row chunk [32][RS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.R, RS = a.RS;
    const int nrowcta = (a.n + ROWS - 1) / ROWS;
#define TRITD_STAMP(blk, q)                                                         \
    if (a.dbg && blockIdx.x == (blk) && tid == 0) {                                 \
        long long t_;                                                               \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                      \
        a.dbg[q] = t_;                                                              \
    }
    TRITD_STAMP(0, 0)
    TRITD_STAMP(1, 4)

    if (blockIdx.x == 0) {
        if (!a.apply) return;
        // ---------------- the ridge system, inverted while the row CTAs reduce ----------------
        bool bad;
        if (R <= 16) bad = invert_ridge_system<1>(a.S1, a.ns1, a.S2, a.ns2, a.alpha, R, RS, a.Minv, red);
        else if (R <= 32) bad = invert_ridge_system<2>(a.S1, a.ns1, a.S2, a.ns2, a.alpha, R, RS, a.Minv, red);
        else if (R <= 48) bad = invert_ridge_system<3>(a.S1, a.ns1, a.S2, a.ns2, a.alpha, R, RS, a.Minv, red);
        else bad = invert_ridge_system<4>(a.S1, a.ns1, a.S2, a.ns2, a.alpha, R, RS, a.Minv, red);
        if (bad && tid == 0) atomicExch(&a.st->status, kStatusCholesky);
        __syncthreads();
        if (tid == 0) st_release_u32(&a.flags[0], 1u);
        TRITD_STAMP(0, 1)
    } else {
        // ---------------- RHS rows: fixed-order reduction from the source ----------------
        const int row0 = (blockIdx.x - 1) * ROWS;
        const int r = warp / WPR, sub = warp - r * WPR;
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
            if (SRC == kSrcDirect) {
                if (sub == 0)
#pragma unroll
                    for (int q = 0; q < KPL; ++q)
                        if (kok[q]) acc[q][0] = a.rhs[(size_t)row * RS + lane + 32 * q];
            } else {
                const int count = SRC == kSrcPartF ? a.part_count : (SRC == kSrcPB ? a.n3 : a.n2);
                const double* base;           // item m lives at base + m * stride (+ k)
                size_t stride;
                const double* wbase = a.W + lane;
                if (SRC == kSrcPartF) {
                    const int it = row / a.tile_h, il = row - it * a.tile_h;
                    base = a.part + ((size_t)it * 128 + il) * RS + lane;
                    stride = (size_t)a.nit * 128 * RS;
                } else if (SRC == kSrcPB) {
                    base = a.P + (size_t)row * RS + lane;
                    stride = (size_t)a.n2 * RS;
                } else {
                    base = a.P + (size_t)row * a.n2 * RS + lane;
                    stride = (size_t)RS;
                }
                // the loads of a round are issued together (CH x KPL per lane in flight), then accumulated
                int m0 = sub;
                for (; m0 + (CH - 1) * WPR < count; m0 += WPR * CH) {
                    double v[KPL][CH], w[KPL][CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const size_t m = (size_t)(m0 + c * WPR);
#pragma unroll
                        for (int q = 0; q < KPL; ++q) {
                            v[q][c] = kok[q] ? base[m * stride + 32 * q] : 0.0;
                            if (SRC != kSrcPartF) w[q][c] = kok[q] ? wbase[m * RS + 32 * q] : 0.0;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CH; ++c)
#pragma unroll
                        for (int q = 0; q < KPL; ++q) {
                            if (SRC == kSrcPartF) acc[q][c] += v[q][c];
                            else acc[q][c] = fma(w[q][c], v[q][c], acc[q][c]);
                        }
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int m = m0 + c * WPR;
                    if (m < count) {
#pragma unroll
                        for (int q = 0; q < KPL; ++q) {
                            const double v = kok[q] ? base[(size_t)m * stride + 32 * q] : 0.0;
                            if (SRC == kSrcPartF) acc[q][c] += v;
                            else acc[q][c] = fma(kok[q] ? wbase[(size_t)m * RS + 32 * q] : 0.0, v, acc[q][c]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const double s = ((acc[q][0] + acc[q][1]) + (acc[q][2] + acc[q][3])) + ((acc[q][4] + acc[q][5]) + (acc[q][6] + acc[q][7]));
            red[warp * 64 + lane + 32 * q] = s;
        }
        __syncthreads();
        for (int e = tid; e < ROWS * RS; e += kUpdThreads) {
            const int rr = e / RS, k = e - rr * RS;
            double v = red[(rr * WPR) * 64 + k];
#pragma unroll
            for (int s = 1; s < WPR; ++s) v += red[(rr * WPR + s) * 64 + k];
            rhs_s[rr * 64 + k] = v;
            if (!a.apply && row0 + rr < a.n) a.rhs_out[(size_t)(row0 + rr) * RS + k] = v;
        }
        if (!a.apply) return;
        TRITD_STAMP(1, 5)

        // ---------------- apply the inverse: X[row][:] = RHS[row][:] * inv(G) ----------------
        cta_wait_eq(&a.flags[0], 1u);              // (also orders the rhs_s writes above)
        TRITD_STAMP(1, 6)
        for (int e = tid; e < R * RS; e += kUpdThreads) Ms[e] = __ldcg(a.Minv + e);
        __syncthreads();
        TRITD_STAMP(1, 8)
        for (int e = tid; e < ROWS * RS; e += kUpdThreads) {
            const int rr = e / RS, k = e - rr * RS;
            double v = 0.0;
            if (k < R) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int m = 0;
                for (; m + 3 < R; m += 4) {
                    s0 = fma(rhs_s[rr * 64 + m], Ms[m * RS + k], s0);
                    s1 = fma(rhs_s[rr * 64 + m + 1], Ms[(m + 1) * RS + k], s1);
                    s2 = fma(rhs_s[rr * 64 + m + 2], Ms[(m + 2) * RS + k], s2);
                    s3 = fma(rhs_s[rr * 64 + m + 3], Ms[(m + 3) * RS + k], s3);
                }
                for (; m < R; ++m) s0 = fma(rhs_s[rr * 64 + m], Ms[m * RS + k], s0);
                v = (s0 + s1) + (s2 + s3);
            }
            xs[rr * 64 + k] = v;
            if (row0 + rr < a.n) a.X[(size_t)(row0 + rr) * RS + k] = v;
        }
        TRITD_STAMP(1, 9)
        if (a.XT) {
            __syncthreads();
            for (int e = tid; e < ROWS * RS; e += kUpdThreads) {
                const int k = e / ROWS, rr = e - k * ROWS;
                if (row0 + rr < a.n) a.XT[(size_t)k * a.ldt + row0 + rr] = xs[rr * 64 + k];
            }
        }
        __syncthreads();
        TRITD_STAMP(1, 10)
        if (tid == 0) red_release_add_u32(&a.flags[1], 1u);
        TRITD_STAMP(1, 7)
    }

    // ---------------- S = X'X once every row is written ----------------
    // G entry slices (256 entries, one per thread) x gr row slices, one (entry, row) slice per CTA: every slice is a
    // couple of 64-row chunks, so the phase is one or two L2 round trips; the gr partial matrices are summed by the
    // consumer (block 0 of the next updates) in slice order.
    const int RR = R * R;
    const int G = (RR + kUpdThreads - 1) / kUpdThreads;
    const int nslice = G * a.gr;
    if ((int)blockIdx.x >= nslice) return;
    cta_wait_eq(&a.flags[1], (unsigned)nrowcta);
    TRITD_STAMP(0, 2)
    for (int sl = blockIdx.x; sl < nslice; sl += gridDim.x) {
        const int es = sl % G, rs = sl / G;
        const int r0 = (int)((long)a.n * rs / a.gr), r1 = (int)((long)a.n * (rs + 1) / a.gr);
        const int e = es * kUpdThreads + tid;
        const bool ok = e < RR;
        const int aa = ok ? e / R : 0, bb = ok ? e - aa * R : 0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int i0 = r0; i0 < r1; i0 += 64) {
            __syncthreads();
            const int nr = min(64, r1 - i0);
            for (int q = tid; q < 64 * RS; q += kUpdThreads) Ms[q] = q < nr * RS ? __ldcg(a.X + (size_t)i0 * RS + q) : 0.0;
            __syncthreads();
            TRITD_STAMP(0, 11)
            const int nr4 = (nr + 3) & ~3;
            for (int i = 0; i < nr4; i += 4) {
                s0 = fma(Ms[i * RS + aa], Ms[i * RS + bb], s0);
                s1 = fma(Ms[(i + 1) * RS + aa], Ms[(i + 1) * RS + bb], s1);
                s2 = fma(Ms[(i + 2) * RS + aa], Ms[(i + 2) * RS + bb], s2);
                s3 = fma(Ms[(i + 3) * RS + aa], Ms[(i + 3) * RS + bb], s3);
            }
        }
        if (ok) a.gram_out[(size_t)rs * RS * RS + aa * RS + bb] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    TRITD_STAMP(0, 12)
    if (tid == 0) {
        const unsigned parts = (unsigned)min(nslice, (int)gridDim.x);
        if (atom_acq_rel_add_u32(&a.flags[2], 1u) == parts - 1) {      // last Gram CTA: leave the flags zero for the next launch
            a.flags[0] = 0u; a.flags[1] = 0u; a.flags[2] = 0u;
        }
    }
    TRITD_STAMP(0, 3)
#undef TRITD_STAMP
}

}  // namespace tritd
