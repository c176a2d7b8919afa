// Mode-wise contractions X_(k) * M^T of the TriTD-ADMM factor updates
// (reference: fast_robust_triple_tensor/triple_decomp_ADMM.m:73-95, the dgemm
// calls at :78, :86, :93), as FP64 DMMA kernels fed by TMA.  The design
// matrices F/G/H (buildF/G/H, :132-160) and the permuted unfold copies
// (:97-109) are never materialised: with A1 = unfold(A,1), B2 = unfold(B,2),
// C3 = unfold(C,3) (n x R, R = r^2)
//     X1*F' [i,k] = sum_{j,t} T(i,j,t) B2(j,k) C3(t,k)            (k_mttkrp1)
//     P[t][j][k]  = sum_i     T(i,j,t) A1(i,k)                     (k_ppass)
//     X2*G' [j,k] = sum_t C3(t,k) P[t][j][k]                       (k_upd, strided-sum source over t)
//     X3*H' [t,k] = sum_j B2(j,k) P[t][j][k]                       (k_upd, strided-sum source over j)
// P depends on T and the *new* A only, so one pass over T serves both the B and
// the C update (Gauss-Seidel order is preserved: the sum over j runs after B is solved).
//
// Device layout: every N-sized array is column-major n1 x n2 x n3 with leading
// dimension ld1 = n1 rounded up to even (16-byte columns for TMA / v2 access);
// factor matrices are row-major n x RS with RS = R rounded up to a multiple of 8,
// zero padded.
#pragma once
#include "common.cuh"
#include "kernels_update.cuh"

namespace tritd {

constexpr int kCW = 8;            // consumer warps per contraction CTA
constexpr int kBoxRows = 32;      // rows (j) of one TMA box
constexpr int kBoxBytes = kBoxRows * 16 * 8;   // [32][16] doubles = 4 KB
constexpr int kStages = 4;
constexpr int kPJ = kBoxRows + 8; // pitch of the transposed factor chunk; == 8 (mod 16) -> conflict-free LDS.128

// ---------------------------------------------------------------------------
// k_mttkrp1: partial RHS of the A update.
// CTA = 8 consumer warps x 16 rows i (one 128-row i-tile) + 1 TMA producer warp.
// Work unit = (i-tile, j-chunk of 32, slice t): 8 TMA boxes [32 j][16 i], 128B-swizzled.
// Units are linearised u = (itile * n_jc + jc) * n3 + t and each CTA takes one
// contiguous range, so the split is balanced to one unit.  A CTA's range touches at
// most two i-tiles; it writes one partial per touched tile into slot 0 / slot 1.
// MMA roles: M = i (two m-tiles: even/odd i of the warp's 16), K = j, N = k.
// ---------------------------------------------------------------------------
struct Mttkrp1Args {
    const double* B2;   // [n2][RS]
    const double* C3;   // [n3][RS]
    double* part;       // [grid][2][128][RS]
    const int* stop;
    int n1, n2, n3, RS;
    int n_it, n_jc;     // tiles along i (128) and j (32)
    long units;         // n_it * n_jc * n3
};

template <int NT>
__global__ void __launch_bounds__((kCW + 1) * 32, 1)
k_mttkrp1(const __grid_constant__ CUtensorMap mapT, const Mttkrp1Args a) {
    if (*a.stop) return;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);                       // kStages * 8 boxes
    double* B2T = reinterpret_cast<double*>(smem_raw + kStages * kCW * kBoxBytes);  // [NT*8][kPJ]
    uint64_t* full = reinterpret_cast<uint64_t*>(B2T + NT * 8 * kPJ);
    uint64_t* empty = full + kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const long u0 = a.units * blockIdx.x / gridDim.x;
    const long u1 = a.units * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kCW); }
        mbar_fence_init();
    }
    __syncthreads();

    const long per_it = (long)a.n_jc * a.n3;
    if (warp == kCW) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            tma_prefetch_desc(&mapT);
            int s = 0; uint32_t ph = 0;
            for (long u = u0; u < u1; ++u) {
                const int it = (int)(u / per_it);
                const long rem = u - it * per_it;
                const int jc = (int)(rem / a.n3), t = (int)(rem - (long)jc * a.n3);
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], kCW * kBoxBytes);
                double* dst = stage_base + (size_t)s * kCW * (kBoxBytes / 8);
#pragma unroll
                for (int w = 0; w < kCW; ++w)
                    tma_load_3d(dst + w * (kBoxBytes / 8), &mapT, &full[s], it * 128 + w * 16, jc * kBoxRows, t);
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    double acc[2][NT][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;

    int s = 0; uint32_t ph = 0;
    int cur_it = -1, cur_jc = -1, slot = 0;
    double* part_cta = a.part + (size_t)blockIdx.x * 2 * 128 * a.RS;

    auto flush = [&](int sl) {
        double* p = part_cta + (size_t)sl * 128 * a.RS;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int il = warp * 16 + 2 * g + m;
                *reinterpret_cast<double2*>(p + (size_t)il * a.RS + 8 * n + 2 * tig) =
                    make_double2(acc[m][n][0], acc[m][n][1]);
                acc[m][n][0] = acc[m][n][1] = 0.0;
            }
    };

    for (long u = u0; u < u1; ++u) {
        const int it = (int)(u / per_it);
        const long rem = u - it * per_it;
        const int jc = (int)(rem / a.n3), t = (int)(rem - (long)jc * a.n3);
        if (it != cur_it) {
            if (cur_it >= 0) { flush(slot); slot = 1; }
            cur_it = it; cur_jc = -1;
        }
        if (jc != cur_jc) {
            // (re)load the transposed chunk B2T[k][j] = B2[jc*32 + j][k]; consumer-only barrier
            asm volatile("bar.sync 1, %0;" ::"n"(kCW * 32));
            for (int e = threadIdx.x; e < NT * 8 * kBoxRows; e += kCW * 32) {
                const int j = e / (NT * 8), k = e - j * (NT * 8);
                const int jj = jc * kBoxRows + j;
                B2T[k * kPJ + j] = (jj < a.n2) ? a.B2[(size_t)jj * a.RS + k] : 0.0;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kCW * 32));
            cur_jc = jc;
        }
        double c3s[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) c3s[n] = __ldg(a.C3 + (size_t)t * a.RS + 8 * n + g);

        mbar_wait(&full[s], ph);
        const double* box = stage_base + (size_t)s * kCW * (kBoxBytes / 8) + warp * (kBoxBytes / 8);
        if (it * 128 + warp * 16 < a.n1) {
#pragma unroll
            for (int jg = 0; jg < kBoxRows / 8; ++jg) {
                // k-step e covers j = jg*8 + 2*tig + e; lane's two doubles are i = 2g (m-tile 0), 2g+1 (m-tile 1)
                const double2 a0 = lds_swz128(box, jg * 8 + 2 * tig, g);
                const double2 a1 = lds_swz128(box, jg * 8 + 2 * tig + 1, g);
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const double2 b = *reinterpret_cast<const double2*>(B2T + (8 * n + g) * kPJ + jg * 8 + 2 * tig);
                    const double b0 = b.x * c3s[n], b1 = b.y * c3s[n];
                    dmma884(acc[0][n][0], acc[0][n][1], a0.x, b0);
                    dmma884(acc[1][n][0], acc[1][n][1], a0.y, b0);
                    dmma884(acc[0][n][0], acc[0][n][1], a1.x, b1);
                    dmma884(acc[1][n][0], acc[1][n][1], a1.y, b1);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == kStages) { s = 0; ph ^= 1; }
    }
    if (cur_it >= 0) flush(slot);
}

// Sum the per-CTA partials of k_mttkrp1 in a fixed order (deterministic).  256 threads = 8 CTA-lanes x
// 32 consecutive outputs (i,k); lane l adds the contributing CTAs c = l, l+8, ... and the 8 lanes are
// combined by a fixed tree.  tile0[c]/tile1[c] (host-computed) are the first/last i-tile CTA c touched:
// it contributes to tile `it` through slot it - tile0[c] of its `cta_stride`-sized partial block.
__global__ void __launch_bounds__(256) k_mttkrp1_reduce(const double* part, size_t cta_stride, int tile_h, double* rhs, int n1, int RS,
                                                       const int* tile0, const int* tile1, int grid_m,
                                                       const int* stop) {
    if (*stop) return;
    __shared__ double red[8][33];
    const int ol = threadIdx.x & 31, cl = threadIdx.x >> 5;
    const long idx = (long)blockIdx.x * 32 + ol;
    const bool ok = idx < (long)n1 * RS;
    double sum = 0.0;
    if (ok) {
        const int i = (int)(idx / RS), k = (int)(idx - (long)i * RS);
        const int it = i / tile_h, il = i - it * tile_h;
        double s1 = 0.0;
        int c = cl;
        for (; c + 8 < grid_m; c += 16) {      // two independent chains: twice the loads in flight
            const int t0 = tile0[c], u0 = tile0[c + 8];
            const bool h0 = it >= t0 && it <= tile1[c], h1 = it >= u0 && it <= tile1[c + 8];
            const double x0 = h0 ? part[(size_t)c * cta_stride + ((size_t)(it - t0) * 128 + il) * RS + k] : 0.0;
            const double x1 = h1 ? part[(size_t)(c + 8) * cta_stride + ((size_t)(it - u0) * 128 + il) * RS + k] : 0.0;
            sum += x0; s1 += x1;
        }
        for (; c < grid_m; c += 8) {
            const int t0 = tile0[c];
            if (it >= t0 && it <= tile1[c]) sum += part[(size_t)c * cta_stride + ((size_t)(it - t0) * 128 + il) * RS + k];
        }
        sum += s1;
    }
    red[cl][ol] = sum;
    __syncthreads();
    if (cl == 0 && ok)
        rhs[idx] = ((red[0][ol] + red[1][ol]) + (red[2][ol] + red[3][ol])) + ((red[4][ol] + red[5][ol]) + (red[6][ol] + red[7][ol]));
}

// ---------------------------------------------------------------------------
// k_ppass: P[t][j][k] = sum_i T(i,j,t) * A1(i,k)      (shared by the B and C updates)
// CTA = 16 consumer warps + 1 TMA producer warp.  Row blocks rb = t * n_jb + jb (32 rows j
// of slice t); a pass is 8 consecutive row blocks, two warps (16 rows each) per block.  Per K-chunk of 16 i
// a stage holds 8 T boxes [32 j][16 i] and one factor box A1T[RS k][16 i], all 128B-swizzled.
// MMA roles: M = j (2 m-tiles per warp, rows permuted by rho8), K = i, N = k (columns permuted by rho8).
// ---------------------------------------------------------------------------
struct PpassArgs {
    double* P;          // [n3][n2][RS]
    const int* stop;
    int n1, n2, n3, RS;
    int n_jb;           // ceil(n2 / 32)
    long n_rb;          // n3 * n_jb
    long units;         // ceil(n_rb / 8) (grid sizing)
    RidgeJob inv;       // update B's ridge inverse (S_A o S_C + lambda2 I, S_A final since update A): CTA 0, before its row blocks
    int inv_rb;         // ... and CTA 0 gets this many row blocks less in return
    IterState* st;
    int R;
};

constexpr int kPW = 16;           // k_ppass consumer warps: two per row block (m-tiles 0,1 / 2,3), 4 per scheduler
// ring depth of k_ppass: as many (8 T boxes + factor box) stages as 227 KB hold, at most 6
__host__ __device__ constexpr int ppass_stages(int NT) {
    return (220 * 1024) / (kCW * kBoxBytes + NT * 8 * 128) > 6 ? 6 : (220 * 1024) / (kCW * kBoxBytes + NT * 8 * 128);
}

// R = 8 (NT - 1) + 1 (r = 3, 5, 7): the last n-tile would hold ONE real column.  That column is formed with scalar DFMA
// from the A fragments already in registers (8 DFMA instead of 8 DMMA per warp and stage: 16 instead of 128 cycles of
// the shared FP64 pipe) and reduced over the quad at the end of the pass: a quarter less DMMA work at r = 5.
__host__ __device__ constexpr bool ppass_scalar_col(int NT, int KS) {
    return (NT == 2 && KS == 3) || (NT == 4 && KS == 7) || (NT == 7 && KS == 13);
}

template <int NT, bool SC = false>
__global__ void __launch_bounds__((kPW + 1) * 32, 1)
k_ppass(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapA1T, const PpassArgs a) {
    constexpr int kStageDoubles = kCW * (kBoxBytes / 8) + NT * 8 * 16;
    constexpr int kStages = ppass_stages(NT);       // (shadows the 4 of k_mttkrp1)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + (size_t)kStages * kStageDoubles);
    uint64_t* empty = full + kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int nkc = (a.n1 + 15) >> 4;
    // each CTA owns a contiguous range of row blocks, balanced to ONE row block (the DMMA pipe is the limit, so
    // time follows the row-block count, not the number of 8-block passes); a pass = 8 consecutive row blocks
    const long drb = a.inv.enable ? a.inv_rb : 0;       // CTA 0's ridge-inverse job counts like drb row blocks in front of its range
    const long rbA = max(0L, (a.n_rb + drb) * blockIdx.x / gridDim.x - drb);
    const long rbB = max(0L, (a.n_rb + drb) * (blockIdx.x + 1) / gridDim.x - drb);
    const long npass = (rbB - rbA + kCW - 1) / kCW;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kPW); }
        mbar_fence_init();
        tma_prefetch_desc(&mapT); tma_prefetch_desc(&mapA1T);
    }
    __syncthreads();
    pdl_wait();                          // everything above overlapped the tail of the previous kernel
    pdl_trigger();
    if (*a.stop) return;

    if (warp == kPW) {
        // Producer warp: lane w < 8 owns slot w of every stage (its row block's (t, jb) is worked out once per pass, not
        // once per chunk), lane 8 the factor box; lane 0 waits for the slot and arms the barrier, then the nine lanes
        // issue their TMA loads side by side -- one thread doing all of it (eight 64-bit divisions and nine issues
        // per stage) was the slowest link of the kernel.
        int s = 0; uint32_t ph = 0;
        for (long u = 0; u < npass; ++u) {
            const long rb0 = rbA + u * kCW;
            const int nvalid = (int)min((long)kCW, rbB - rb0);
            int t = 0, jb = 0;
            if (lane < nvalid) {
                const unsigned rb = (unsigned)(rb0 + lane);
                t = (int)(rb / (unsigned)a.n_jb);
                jb = (int)(rb - (unsigned)t * (unsigned)a.n_jb);
            }
            for (int kc = 0; kc < nkc; ++kc) {
                double* dst = stage_base + (size_t)s * kStageDoubles;
                if (lane == 0) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], nvalid * kBoxBytes + NT * 8 * 128);
                }
                __syncwarp();
                if (lane < nvalid) tma_load_3d(dst + lane * (kBoxBytes / 8), &mapT, &full[s], kc * 16, jb * kBoxRows, t);
                else if (lane == kCW) tma_load_2d(dst + kCW * (kBoxBytes / 8), &mapA1T, &full[s], kc * 16, 0);
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    if (a.inv.enable && blockIdx.x == 0 && threadIdx.x < 256) {
        // update B's ridge system (see k_admm: same scheme), consumer warps 0..7 of CTA 0 before their first pass
        __shared__ double inv_scratch[256];
        ridge_job_run<(NT + 1) / 2, 3, false>(a.inv, a.st, a.st->k, a.R, a.RS, inv_scratch);
    }
    int s = 0; uint32_t ph = 0;
    const int rg = rho8(g);
    // Row block of the pass and half of its 32 rows (m-tiles 2mh, 2mh+1): warps 2b and 2b+1 share block b, so the valid
    // blocks of a partial pass spread over the four schedulers (5 blocks: 3+3+2+2 warps, 3/4 of a full pass's DMMA time;
    // with blocks b and b+8... on one scheduler a 5-block pass cost as much as a full one)
    const int bw = warp >> 1, mh = warp & 1;
    for (long u = 0; u < npass; ++u) {
        const long rb = rbA + u * kCW + bw;
        const bool valid = rb < rbB;
        constexpr int NTD = SC ? NT - 1 : NT;             // n-tiles formed with DMMA
        double acc[2][NTD][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NTD; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
        double sacc[2] = {0.0, 0.0};                      // SC: this lane's share of column 8 * NTD for its two rows

        for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&full[s], ph);
            const double* st = stage_base + (size_t)s * kStageDoubles;
            const double* box = st + bw * (kBoxBytes / 8);
            const double* fbox = st + kCW * (kBoxBytes / 8);
            if (valid) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // k-steps (h,e) cover i = kc*16 + 8h + 2*tig + e
                    double2 af[2];
#pragma unroll
                    for (int m = 0; m < 2; ++m) af[m] = lds_swz128(box, 8 * (2 * mh + m) + rg, 4 * h + tig);
#pragma unroll
                    for (int n = 0; n < NTD; ++n) {
                        const double2 b = lds_swz128(fbox, 8 * n + rg, 4 * h + tig);
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            dmma884(acc[m][n][0], acc[m][n][1], af[m].x, b.x);
                            dmma884(acc[m][n][0], acc[m][n][1], af[m].y, b.y);
                        }
                    }
                    if (SC) {
                        const double2 b = lds_swz128(fbox, 8 * NTD, 4 * h + tig);     // A1(i, 8 NTD) for this lane's two i
#pragma unroll
                        for (int m = 0; m < 2; ++m) sacc[m] = fma(af[m].y, b.y, fma(af[m].x, b.x, sacc[m]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == kStages) { s = 0; ph ^= 1; }
        }
        if (valid) {
            const int t = (int)(rb / a.n_jb), jb = (int)(rb - (long)t * a.n_jb);
            // C fragment: row = 8m' + rho8(g); columns n = 2*tig + c map to k = 8n' + rho8(2*tig + c) = 8n' + tig + 4c
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int j = jb * kBoxRows + 8 * (2 * mh + m) + rg;
                if (j < a.n2) {
                    double* p = a.P + ((size_t)t * a.n2 + j) * a.RS;
#pragma unroll
                    for (int n = 0; n < NTD; ++n) {
                        p[8 * n + tig] = acc[m][n][0];
                        p[8 * n + tig + 4] = acc[m][n][1];
                    }
                }
            }
        }
        if (SC) {             // (all lanes take part in the shuffles; rows of an invalid block hold zeros)
            const int t = valid ? (int)(rb / a.n_jb) : 0, jb = valid ? (int)(rb - (long)t * a.n_jb) : 0;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                double v = sacc[m];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                const int j = jb * kBoxRows + 8 * (2 * mh + m) + rg;
                // columns 8 NTD + 1 .. RS - 1 are padding: zero (the update kernels read whole rows)
                if (valid && j < a.n2) {
                    double* p = a.P + ((size_t)t * a.n2 + j) * a.RS + 8 * NTD;
                    p[tig] = tig == 0 ? v : 0.0;
                    p[tig + 4] = 0.0;
                }
            }
        }
    }
}

}  // namespace tritd
