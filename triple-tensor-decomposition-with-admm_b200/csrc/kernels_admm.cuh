// k_admm: the fused kernel of one TriTD-ADMM iteration, TMA in / DMMA / TMA out.
//
// Reference statements covered (fast_robust_triple_tensor/triple_decomp_ADMM.m):
//   :38      L = triple_product(A,B,C)            -- formed on the fly with DMMA, never stored
//   :41-43   R1, R2, O                            :46-47  R3, E = soft_threshold(R3, lambda/muO)
//   :50-53   resL, resO, Y_L, Y_O                 :59     ||resL||, ||resO|| (per-CTA partial sums)
//   :33      T = D - O + (1/muL)*Y_L of the NEXT iteration (muL already advanced, :56)
//   :74-78   X1*F' of the NEXT iteration's update_A (mode-1 MTTKRP of the new T), accumulated
//            from registers so T is read from HBM only once per iteration (by k_ppass).
//
// Data movement: every warp owns 16 rows i and runs a private S-stage ring of 1 KB TMA boxes
// [8 j][16 i] (128B-swizzled) for D, Y_L, E, Y_O; results are written back in place in shared memory
// (T over D) plus one extra box for O and leave through TMA stores.  No thread touches HBM with a
// load/store instruction; out-of-range rows/columns are zero-filled on load and clipped on store by
// the tensor maps.  Algorithmic traffic: 4 reads + 5 writes = 72 bytes per element.
#pragma once
#include "common.cuh"
#include "kernels_contract.cuh"
#include "kernels_fused.cuh"

namespace tritd {

struct AdmmMaps { CUtensorMap D, YL, E, YO, T, O; };

struct AdmmArgs {
    const double *A1, *B2, *C3;        // [n][RS]
    const IterState* st;
    double* norm_part;                 // [grid][2]
    double* partM;                     // [grid][128][RS]: this CTA's partial of the next X1*F'
    int n1, n2, n3, RS;
    int n_it, n_jc, gi;                // i-tiles (128), j-chunks (32), CTAs per i-tile
};

template <int KS, int NT, bool WRITE_O> struct AdmmCfg {
    static constexpr int PL = FusedCfg<KS>::PL;
    static constexpr int NB = WRITE_O ? 5 : 4;                       // boxes per stage
    static constexpr int kFixed = (32 * PL + NT * 8 * kPJ + 64) * 8 + 1024;
    static constexpr int kAvail = 227 * 1024 - kFixed;
    static constexpr int S = (kAvail / (8 * NB * 1024)) > 6 ? 6 : (kAvail / (8 * NB * 1024));
    static constexpr size_t kSmem = (size_t)8 * S * NB * 1024 + kFixed;
    static_assert(S >= 3, "ring too shallow");
};

template <int KS, int NT, bool WRITE_O>
__global__ void __launch_bounds__(256, 1) k_admm(const __grid_constant__ AdmmMaps maps, const AdmmArgs a) {
    using Cfg = AdmmCfg<KS, NT, WRITE_O>;
    constexpr int PL = Cfg::PL, NB = Cfg::NB, S = Cfg::S, PD = S - 2;
    if (a.st->stop) return;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);              // [8 warps][S][NB][128]
    double* B2s = ring + 8 * S * NB * 128;                           // [32 j][PL]      B-operand of L
    double* B2T = B2s + 32 * PL;                                     // [NT*8 k][kPJ]   B-operand of the MTTKRP
    double* red = B2T + NT * 8 * kPJ;                                // [64]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 64);          // [8][S]
    __shared__ IterState prm_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int it = blockIdx.x % a.n_it, x = blockIdx.x / a.n_it;
    const long V = (long)a.n_jc * a.n3;
    const long v0 = V * x / a.gi, v1 = V * (x + 1) / a.gi;
    const long nq = (v1 - v0) * 4;
    const int iw = it * 128 + warp * 16;
    const bool active = iw < a.n1;
    double* wring = ring + (size_t)warp * S * NB * 128;
    uint64_t* wfull = full + warp * S;

    if (threadIdx.x == 0) prm_s = *a.st;
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&wfull[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue_loads = [&](long q) {          // lane 0 of an active warp
        const long v = v0 + (q >> 2);
        const int jc = (int)(v / a.n3), t = (int)(v - (long)jc * a.n3), j0 = jc * 32 + (int)(q & 3) * 8;
        const int s = (int)(q % S);
        double* st = wring + s * NB * 128;
        mbar_expect_tx(&wfull[s], 4 * 1024);
        tma_load_3d(st, &maps.D, &wfull[s], iw, j0, t);
        tma_load_3d(st + 128, &maps.YL, &wfull[s], iw, j0, t);
        tma_load_3d(st + 256, &maps.E, &wfull[s], iw, j0, t);
        tma_load_3d(st + 384, &maps.YO, &wfull[s], iw, j0, t);
    };
    if (active && lane == 0) {
        tma_prefetch_desc(&maps.D); tma_prefetch_desc(&maps.YL); tma_prefetch_desc(&maps.E);
        tma_prefetch_desc(&maps.YO); tma_prefetch_desc(&maps.T);
        if (WRITE_O) tma_prefetch_desc(&maps.O);
        for (long q = 0; q < PD && q < nq; ++q) issue_loads(q);
    }

    const int i0 = iw + 2 * g;
    const double* a1r0 = a.A1 + (size_t)min(i0, a.n1 - 1) * a.RS + tig;
    const double* a1r1 = a.A1 + (size_t)min(i0 + 1, a.n1 - 1) * a.RS + tig;
    const double z0 = (i0 < a.n1) ? 1.0 : 0.0, z1 = (i0 + 1 < a.n1) ? 1.0 : 0.0;

    double acc[2][NT][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
    double aS[2][KS];
    double c3s[NT];
    double sL = 0.0, sO = 0.0;
    int cur_jc = -1;

    for (long q = 0; q < nq; ++q) {
        const long v = v0 + (q >> 2);
        const int jc = (int)(v / a.n3), t = (int)(v - (long)jc * a.n3);
        const int jg = (int)(q & 3), j0 = jc * 32 + jg * 8;
        if (jg == 0) {
            if (jc != cur_jc) {               // uniform over the CTA: all warps walk the same q sequence
                __syncthreads();
                for (int e = threadIdx.x; e < 32 * 4 * KS; e += 256) {
                    const int j = e / (4 * KS), k = e - j * (4 * KS);
                    const int jj = jc * 32 + j;
                    B2s[j * PL + k] = (jj < a.n2) ? a.B2[(size_t)jj * a.RS + k] : 0.0;
                }
                for (int e = threadIdx.x; e < NT * 8 * 32; e += 256) {
                    const int j = e / (NT * 8), k = e - j * (NT * 8);
                    const int jj = jc * 32 + j;
                    B2T[k * kPJ + j] = (jj < a.n2) ? a.B2[(size_t)jj * a.RS + k] : 0.0;
                }
                __syncthreads();
                cur_jc = jc;
            }
            // A fragments of L with C3[t,:] folded in; column scales of the MTTKRP B fragments
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const double c3 = __ldg(a.C3 + (size_t)t * a.RS + 4 * s + tig);
                aS[0][s] = __ldg(a1r0 + 4 * s) * c3 * z0;
                aS[1][s] = __ldg(a1r1 + 4 * s) * c3 * z1;
            }
#pragma unroll
            for (int n = 0; n < NT; ++n) c3s[n] = __ldg(a.C3 + (size_t)t * a.RS + 8 * n + g);
        }
        if (!active) continue;

        if (lane == 0 && q + PD < nq) {
            tma_store_wait_read<1>();         // the slot's previous stores (group q-2) have left shared memory
            issue_loads(q + PD);
        }
        const int s = (int)(q % S);
        double* st = wring + s * NB * 128;
        mbar_wait(&wfull[s], (uint32_t)((q / S) & 1));

        // L patch: l[m][c] = L(i0 + m, j0 + 2*tig + c, t)
        double l[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const double b = B2s[(jg * 8 + g) * PL + 4 * ks + tig];
            dmma884(l[0][0], l[0][1], aS[0][ks], b);
            dmma884(l[1][0], l[1][1], aS[1][ks], b);
        }
        double2 tn[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int row = 2 * tig + c;
            const double2 d = lds_swz128(st, row, g);
            double2 yl = lds_swz128(st + 128, row, g);
            double2 e = lds_swz128(st + 256, row, g);
            double2 yo = lds_swz128(st + 384, row, g);
            double2 o;
            admm_point(prm_s, d.x, l[0][c], yl.x, e.x, yo.x, o.x, tn[c].x, sL, sO);
            admm_point(prm_s, d.y, l[1][c], yl.y, e.y, yo.y, o.y, tn[c].y, sL, sO);
            sts_swz128(st, row, g, tn[c]);
            sts_swz128(st + 128, row, g, yl);
            sts_swz128(st + 256, row, g, e);
            sts_swz128(st + 384, row, g, yo);
            if (WRITE_O) sts_swz128(st + 512, row, g, o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_3d(&maps.T, st, iw, j0, t);
            tma_store_3d(&maps.YL, st + 128, iw, j0, t);
            tma_store_3d(&maps.E, st + 256, iw, j0, t);
            tma_store_3d(&maps.YO, st + 384, iw, j0, t);
            if (WRITE_O) tma_store_3d(&maps.O, st + 512, iw, j0, t);
            tma_store_commit();
        }
        // next iteration's X1*F' : acc[m][n] += T'(i, j) * B2(j, k) * C3(t, k); k-step c covers j = j0 + 2*tig + c
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const double2 b = *reinterpret_cast<const double2*>(B2T + (8 * n + g) * kPJ + jg * 8 + 2 * tig);
            const double b0 = b.x * c3s[n], b1 = b.y * c3s[n];
            dmma884(acc[0][n][0], acc[0][n][1], tn[0].x, b0);
            dmma884(acc[1][n][0], acc[1][n][1], tn[0].y, b0);
            dmma884(acc[0][n][0], acc[0][n][1], tn[1].x, b1);
            dmma884(acc[1][n][0], acc[1][n][1], tn[1].y, b1);
        }
    }
    if (active && lane == 0) tma_store_wait_all<0>();

    double* p = a.partM + (size_t)blockIdx.x * 128 * a.RS;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
            *reinterpret_cast<double2*>(p + (size_t)(warp * 16 + 2 * g + m) * a.RS + 8 * n + 2 * tig) =
                make_double2(acc[m][n][0], acc[m][n][1]);

    block_sum2(sL, sO, red);
    if (threadIdx.x == 0) { a.norm_part[2 * blockIdx.x] = sL; a.norm_part[2 * blockIdx.x + 1] = sO; }
}

}  // namespace tritd
