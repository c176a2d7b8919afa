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
// State kept in HBM between iterations: D, Y_L, the next target T and ONE array Z for the sparse pair (E, Y_O):
// Z = R3 = O + (1/muO)*Y_O of the last iteration (:46).  E = soft_threshold(Z, lambda/muO) is re-derived where it is
// needed (:47), and the dual follows from the reference's own update (:53), Y_O' = Y_O + muO*(O - E) = muO*(R3 - E) =
// muO*(Z - E) -- an identity in exact arithmetic, so the iterates are the reference's up to a few roundings of
// |Y_O| <= lambda per element and iteration (measured against the oracle: DESIGN 4.1).  That removes one N-sized read and
// one N-sized write per iteration: 48 instead of 64 bytes per element.
//
// Structure: CTA = 8 consumer warps x 16 rows i (an i-tile of <= 128 rows) + 1 warp that drives TMA.
// A stage is JG (1 or 2) groups of 8 columns j of slice t for the whole i-tile, i.e. three 8*JG KB boxes
// (D, Y_L, Z) fetched by ONE TMA op each through a 4-D view (i_lo=16, j, i_hi, t) of the
// column-major arrays, which lands as [warp][8 j][16 i] with the 128B swizzle pattern the DMMA
// accumulator layout reads conflict-free.  Consumers update the boxes in place (T over D; O is not stored
// inside the loop, k_recover_O rebuilds it on demand); the TMA
// warp then streams the boxes back with one TMA store each and refills the slot.  No thread touches HBM with a
// load/store instruction; rows/columns outside the tensor are zero-filled on load and clipped on store by the
// tensor maps (ld1 is a multiple of 16 so the padded rows exist and stay zero).  Algorithmic traffic:
// 3 reads (D, Y_L, Z) + 3 writes (T', Y_L', Z') = 48 bytes per element.
#pragma once
#include "common.cuh"
#include "kernels_contract.cuh"
#include "kernels_fused.cuh"
#include "kernels_xchg.cuh"
#include "kernels_update.cuh"

namespace tritd {

struct AdmmMaps { CUtensorMap D, YL, Z, T, O; };

struct AdmmArgs {
    const double *A1, *B2, *C3;        // [n][RS]
    IterState* st;
    double* norm_part;                 // [grid][2]
    double* norms;                     // [2]: fixed-order sum of norm_part (left for the all-reduce when !finalize)
    double *errHist, *errL, *errO;     // history (written by the last CTA when finalize)
    unsigned* ticket;                  // CTA completion counter (zero between launches)
    int finalize;                      // 1: the last CTA also runs the errHist / mu / stopping-rule step (:56-65)
    // N>1 peer exchange (kernels_xchg.cuh): the last CTA writes the pair into this rank's slot of every mailbox
    double* const* peers;              // nullptr: pair left in norms[] (NCCL path)
    long norm_off;                     // offset in doubles inside a mailbox: this rank's pair (two 16-byte words)
    const double* nslots;              // own mailbox: the ranks' pairs, 8 doubles apart
    int rank, nranks;
    unsigned xbase;
    double* partM;                     // [i-tile][part_slots][128][RS]: this CTA's partial of the next X1*F' at (its tile, its index in the tile)
    RidgeJob inv;                      // update A's ridge inverse of the NEXT iteration: CTA 0 computes it before its streaming work
    int inv_stages;                    // ... and gets this many stages less than its siblings in return
    int R;                             // r^2
    long long* dbg;                    // optional [grid][2] globaltimer stamps: CTA start / end (diagnostics, TRITD_DEBUG_STAMPS)
    int part_slots;                    // slots per tile (= the largest number of CTAs any tile has)
    const int* cta_tab;                // [grid][3]: i-tile, index within the tile's CTAs, CTAs of that tile
    int n1, n2, n3, RS;
    int n1s;                           // rows covered by 16-row strips (n1, or the padded leading dimension)
    int n_jc;                          // j-chunks (32 columns)
    int tile_h;                        // rows per i-tile: 16 * (consumer warps used), <= 128
};

// JGP: column groups per stage asked for (1 or 2).  Two-group stages give every warp two independent dependency
// chains between barriers; one-group stages are half as large, so twice as many fit the ring and twice as many loads
// are in flight per SM -- which wins depends on the shape (chosen at problem set-up, tritd.cu).
template <int KS, int NT, int JGP = 2> struct AdmmCfg {
    static constexpr int PL = FusedCfg<KS>::PL;
    static constexpr int NB = 3;                                                // boxes per stage (D, Y_L, Z)
    // large R: the B operand of L is read from the transposed chunk too (2-way bank conflict on a small share of
    // the shared-memory traffic) so that the second layout's 17 KB buy a two-group stage
    static constexpr bool kShareB = KS > 8;
    static constexpr int kFixed = ((kShareB ? 0 : 32 * PL) + NT * 8 * kPJ + 64) * 8 + 1024;
    static constexpr int kAvail = 227 * 1024 - kFixed;
    // a stage holds JG groups of 8 columns: two when three such stages fit (more independent work per warp
    // between barriers), else one
    static constexpr int JG = (JGP >= 2 && (kAvail / (NB * 2 * 8 * 128 * 8 + 1024)) >= 3) ? 2 : 1;
    static constexpr int kBoxD = 8 * JG * 128;                                  // doubles per array per stage: [8 warps][8*JG j][16 i]
    static constexpr int kStageBytes = NB * kBoxD * 8 + 1024;                   // + the C3 row of the slice; keeps boxes 1 KB aligned
    static constexpr int S = (kAvail / kStageBytes) > 8 ? 8 : (kAvail / kStageBytes);
    static constexpr size_t kSmem = (size_t)S * kStageBytes + kFixed;
    static_assert(S >= 3, "ring too shallow");
};

// 1-D bulk copy global -> shared with mbarrier completion (the C3 row of a slice).  SASS: UBLKCP.
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// Correctly rounded a/b for normal-range operands from y = RN(1/b): q0 = RN(a*y), r = a - q0*b (exact in
// one FMA), q = RN(q0 + r*y) (Markstein).  Replaces the ~35-instruction IEEE division sequence in the O update.
__device__ __forceinline__ double div_by(double a, double b, double y) {
    const double q0 = __dmul_rn(a, y);
    const double r = __fma_rn(-q0, b, a);
    return __fma_rn(r, y, q0);
}

// One element of triple_decomp_ADMM.m:41-53 and the next T (:33); same association as the MATLAB
// expressions, every operation an explicit round-to-nearest intrinsic (no FMA contraction).
struct AdmmPrm { double muL, muO, rmuL, thr, musum, rmusum, rmuL_next, yo_scale, thr_prev; };

// x - soft_threshold(x, thr) = x clipped to [-thr, thr]; soft_threshold(x, thr) = sign(x).*max(|x|-thr,0)
// (soft_threshold.m:2) is then x - clip(x): the same subtraction |x| - thr for |x| > thr (negation is exact), x - x = 0
// inside, and NaN where x is NaN (fmax / fmin drop the NaN, NaN - finite = NaN: MATLAB's sign(NaN)*0 = NaN).  Two DMNMX
// and one DADD instead of a subtraction, a maximum, three comparisons and the selects.
__device__ __forceinline__ double clip_thr(double x, double thr) { return fmin(fmax(x, -thr), thr); }

template <bool MASKED>
__device__ __forceinline__ void admm_point2(const AdmmPrm& p, double d, double l, double& yl, double& z, double& tn,
                                            double& sL, double& sO) {
    if (MASKED && d != d) {              // unobserved entry (NaN in D): no constraint, impute the low-rank estimate
        yl = 0.0; z = 0.0; tn = l;
        return;
    }
    // The sparse pair of the previous iteration from Z = its R3 (:46): E = soft_threshold(R3, lambda/muO) = Z - clip(Z) (:47)
    // and Y_O = Y_O_old + muO*(O - E) = muO*(R3 - E) = muO*clip(Z) (:53), so (1/muO')*Y_O = (muO/muO')*clip(Z).
    // (Z = 0 before the first iteration: E = Y_O = 0.)
    const double w = clip_thr(z, p.thr_prev);
    const double e = __dsub_rn(z, w);
    const double my = __dmul_rn(p.yo_scale, w);                          // (1/muO)*Y_O
    const double dl = __dsub_rn(d, l);                                   // D - L
    const double r1 = __dadd_rn(dl, __dmul_rn(p.rmuL, yl));              // R1 = D - L + (1/muL)*Y_L
    const double r2 = __dsub_rn(e, my);                                  // R2 = E - (1/muO)*Y_O
    const double o = div_by(__dadd_rn(__dmul_rn(p.muL, r1), __dmul_rn(p.muO, r2)), p.musum, p.rmusum);
    const double r3 = __dadd_rn(o, my);                                  // R3 = O + (1/muO)*Y_O
    const double en = __dsub_rn(r3, clip_thr(r3, p.thr));                // E = soft_threshold(R3, lambda/muO)
    const double resL = __dsub_rn(dl, o);                                // D - L - O
    const double resO = __dsub_rn(o, en);                                // O - E
    yl = __dadd_rn(yl, __dmul_rn(p.muL, resL));
    z = r3;
    tn = __dadd_rn(__dsub_rn(d, o), __dmul_rn(p.rmuL_next, yl));         // next T = D - O + (1/muL')*Y_L
    sL = fma(resL, resL, sL);
    sO = fma(resO, resO, sO);
}

// Block = 3 warpgroups: two of consumers (8 warps), one whose first lane drives TMA.  The kernel is
// compiled for 168 registers/thread (65536 / 384); the TMA warpgroup gives most of its share back and
// the consumers grow to 224 with setmaxnreg, so the DMMA accumulators and fragments never spill.
constexpr int kAdmmThreads = 384;

// MASKED: the opt-in completion variant (tritd_admm_masked_f64): unobserved entries are stored as NaN
// in D; there O = E = Y_L = Y_O = 0, the residuals do not count and the next target is T' = L (imputation).
template <int KS, int NT, bool MASKED, int JGP = 2>
__global__ void __launch_bounds__(kAdmmThreads, 1) k_admm(const __grid_constant__ AdmmMaps maps_full, const __grid_constant__ AdmmMaps maps_last,
                                                          const AdmmArgs a) {
    using Cfg = AdmmCfg<KS, NT, JGP>;
    constexpr int PL = Cfg::PL, NB = Cfg::NB, S = Cfg::S, JG = Cfg::JG, kBoxD = Cfg::kBoxD;
    constexpr int kStageD = Cfg::kStageBytes / 8;
    constexpr int SPU = 4 / JG;           // stages per unit (32 columns of one slice)
    constexpr bool kFoldA = KS <= 8;      // fold C3[t,:] into the A fragments once per slice (else into B per use)
    constexpr int kPartH = 128;           // rows of a CTA's X1*F' partial
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);              // [S][NB boxes + C3 row]
    double* B2s = ring + (size_t)S * kStageD;                        // [32 j][PL]      B-operand of L
    double* B2T = B2s + (Cfg::kShareB ? 0 : 32 * PL);                // [NT*8 k][kPJ]   B-operand of the MTTKRP
    double* red = B2T + NT * 8 * kPJ;                                // [64]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 64);          // [S] TMA landed
    uint64_t* done = full + S;                                       // [S] consumers finished writing back

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int it = a.cta_tab[3 * blockIdx.x], x = a.cta_tab[3 * blockIdx.x + 1], gi = a.cta_tab[3 * blockIdx.x + 2];
    // The CTAs of an i-tile split its stages (a stage = 8*JG columns of one slice; SPU stages = one unit of 32 columns)
    // evenly: a contiguous range of stage indices q in [q0, q1), unit = q / SPU, column group = q % SPU.
    // CTA 0 additionally inverts the ridge system of the next update A (below), which takes about as long as
    // `inv_stages` stages: the split of ITS tile treats that job like stages in front of CTA 0's range.
    const long Q = (long)a.n_jc * a.n3 * SPU;
    const long dq = (a.inv.enable && it == a.cta_tab[0]) ? a.inv_stages : 0;
    const long q0 = max(0L, (Q + dq) * x / gi - dq), q1 = max(0L, (Q + dq) * (x + 1) / gi - dq);
    const long nq = q1 - q0;
    const int nwf = a.tile_h >> 4;                                   // 16-row strips (= consumer warps) of a full i-tile
    const int nact = min(nwf, (a.n1s - it * a.tile_h + 15) >> 4);     // ... of THIS tile (the last tile may have fewer)
    // TMA box depth in i_hi = strips of this tile: the last tile has its own tensor maps, so that no box ever
    // reaches past the tensor in i (boxes clipped by the out-of-bounds logic measured ~7 % slower)
    const int nw = nact;
    const AdmmMaps& maps = nact < nwf ? maps_last : maps_full;
    const int i_hi0 = it * nwf;                                      // first strip of this tile
    const long u0 = q0 / SPU;
    const int jc0 = (int)(u0 / a.n3), t0 = (int)(u0 - (long)jc0 * a.n3), sg0 = (int)(q0 - u0 * SPU);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], nact); }
        mbar_fence_init();
        if (a.dbg) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.dbg[2 * blockIdx.x] = t_; }
        tma_prefetch_desc(&maps.D); tma_prefetch_desc(&maps.YL); tma_prefetch_desc(&maps.Z); tma_prefetch_desc(&maps.T);
    }
    __syncthreads();
    pdl_wait();                           // everything above overlapped the tail of the previous kernel
    pdl_trigger();
    if (a.st->stop) return;
    const int k_start = a.st->k;          // iteration index of this launch (the last CTA advances it at the very end)

    if (warp >= 8) {
        // ---------------- TMA warpgroup: loads, stores, slot recycling ----------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 8 && lane == 0 && nq > 0) {
            int ljc = jc0, lt = t0, ljg = sg0, ls = 0;               // load cursor
            auto issue_load = [&]() {
                double* st = ring + (size_t)ls * kStageD;
                const int j0 = ljc * 32 + ljg * (8 * JG);
                mbar_expect_tx(&full[ls], NB * nw * JG * 1024 + NT * 8 * 8);
                tma_load_4d(st, &maps.D, &full[ls], 0, j0, i_hi0, lt);
                tma_load_4d(st + kBoxD, &maps.YL, &full[ls], 0, j0, i_hi0, lt);
                tma_load_4d(st + 2 * kBoxD, &maps.Z, &full[ls], 0, j0, i_hi0, lt);
                bulk_load_1d(st + NB * kBoxD, a.C3 + (size_t)lt * a.RS, NT * 8 * 8, &full[ls]);   // C3 row of slice t
                if (++ljg == SPU) { ljg = 0; if (++lt == a.n3) { lt = 0; ++ljc; } }
                if (++ls == S) ls = 0;
            };
            long loaded = 0;
            for (; loaded < S && loaded < nq; ++loaded) issue_load();
            int sjc = jc0, stt = t0, sjg = sg0, ss = 0; uint32_t sph = 0;   // store cursor
            for (long q = 0; q < nq; ++q) {
                mbar_wait(&done[ss], sph);
                double* st = ring + (size_t)ss * kStageD;
                const int j0 = sjc * 32 + sjg * (8 * JG);
                tma_store_4d(&maps.T, st, 0, j0, i_hi0, stt);
                tma_store_4d(&maps.YL, st + kBoxD, 0, j0, i_hi0, stt);
                tma_store_4d(&maps.Z, st + 2 * kBoxD, 0, j0, i_hi0, stt);
                tma_store_commit();
                if (++sjg == SPU) { sjg = 0; if (++stt == a.n3) { stt = 0; ++sjc; } }
                if (++ss == S) { ss = 0; sph ^= 1; }
                if (q >= 1 && loaded < nq) {          // stage q-1 has left shared memory: refill its slot
                    tma_store_wait_read<1>();
                    issue_load();
                    ++loaded;
                }
            }
            tma_store_wait_all<0>();
        }
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
      if (a.inv.enable && blockIdx.x == 0) {
        // The ridge system of the NEXT update A (S_B o S_C + lambda2 I, both final since update C) is inverted here, by
        // the eight consumer warps of CTA 0 before they start on their (correspondingly shorter) range of stages: the
        // inverse is ready long before update A launches, which then applies it the moment its right-hand sides are
        // summed.  Scratch: the factor-chunk buffer (not loaded yet).  An ill-conditioned system is left to update A.
        ridge_job_run<(NT + 1) / 2, 3, false>(a.inv, a.st, k_start + 1, a.R, a.RS, B2s);
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      if (warp < nact) {
        // ---------------- consumers ----------------
        const int g = lane >> 2, tig = lane & 3;
        AdmmPrm prm;
        {
            const IterState& S0 = *a.st;
            prm.muL = S0.muL; prm.muO = S0.muO; prm.rmuL = S0.rmuL; prm.thr = S0.thr;
            prm.musum = S0.musum; prm.rmusum = 1.0 / S0.musum; prm.rmuL_next = S0.rmuL_next;
            prm.yo_scale = S0.muO_prev * S0.rmuO; prm.thr_prev = S0.thr_prev;
        }
        double sL = 0.0, sO = 0.0;
        const int nthr = nact * 32;
        const int i0 = it * a.tile_h + warp * 16 + 2 * g;
        double aF[2][KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            aF[0][s] = (i0 < a.n1) ? __ldg(a.A1 + (size_t)i0 * a.RS + 4 * s + tig) : 0.0;
            aF[1][s] = (i0 + 1 < a.n1) ? __ldg(a.A1 + (size_t)(i0 + 1) * a.RS + 4 * s + tig) : 0.0;
        }
        double acc[2][NT][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
        double aS[2][kFoldA ? KS : 1];
        // swizzled offsets (in double2 units) of this lane's two columns c = 0, 1 of the first column group inside its
        // warp's [8*JG j][16 i] block; the second group is 64 further (8 rows of 128 B, same swizzle phase)
        const int off0 = warp * (64 * JG) + (2 * tig) * 8 + (g ^ (2 * tig));
        const int off1 = warp * (64 * JG) + (2 * tig + 1) * 8 + (g ^ (2 * tig + 1));
        int slot = 0; uint32_t ph = 0;
        int jc = jc0, t = t0, cur_jc = -1;

        const long u_end = nq > 0 ? (q1 - 1) / SPU + 1 : u0;
        for (long u = u0; u < u_end; ++u) {
            if (jc != cur_jc) {             // uniform over the consumers: all walk the same sequence
                asm volatile("bar.sync 1, %0;" ::"r"(nthr));
                if (!Cfg::kShareB)
                    for (int e = threadIdx.x; e < 32 * 4 * KS; e += nthr) {
                        const int j = e / (4 * KS), k = e - j * (4 * KS);
                        const int jj = jc * 32 + j;
                        B2s[j * PL + k] = (jj < a.n2) ? a.B2[(size_t)jj * a.RS + k] : 0.0;
                    }
                for (int e = threadIdx.x; e < NT * 8 * 32; e += nthr) {
                    const int j = e / (NT * 8), k = e - j * (NT * 8);
                    const int jj = jc * 32 + j;
                    B2T[k * kPJ + j] = (jj < a.n2) ? a.B2[(size_t)jj * a.RS + k] : 0.0;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(nthr));
                cur_jc = jc;
            }
#pragma unroll
            for (int sg = 0; sg < SPU; ++sg) {
                const long qa = u * SPU + sg;                // (the first and the last unit of a CTA may be partial)
                if (qa < q0 || qa >= q1) continue;
                double* st = ring + (size_t)slot * kStageD;
                mbar_wait(&full[slot], ph);
                const double* c3row = st + NB * kBoxD;       // C3(t, :), landed with this stage
                if (kFoldA && (sg == 0 || qa == q0)) {
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        const double c3 = c3row[4 * s + tig];
                        aS[0][s] = aF[0][s] * c3;
                        aS[1][s] = aF[1][s] * c3;
                    }
                }
                double c3s[NT];                              // column scales of the MTTKRP B fragments
#pragma unroll
                for (int n = 0; n < NT; ++n) c3s[n] = c3row[8 * n + g];
                double2* s2 = reinterpret_cast<double2*>(st);
                // the JG column groups of a stage are independent: one basic block, so their DMMA and
                // element-wise dependency chains interleave
#pragma unroll
                for (int h = 0; h < JG; ++h) {
                    const int jg = sg * JG + h;
                    // L patch: l[m][c] = L(i0 + m, j0 + 2*tig + c, t)
                    double l[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        double b = Cfg::kShareB ? B2T[(4 * ks + tig) * kPJ + jg * 8 + g] : B2s[(jg * 8 + g) * PL + 4 * ks + tig];
                        if (!kFoldA) b *= c3row[4 * ks + tig];
                        if (kFoldA) {
                            dmma884(l[0][0], l[0][1], aS[0][ks], b);
                            dmma884(l[1][0], l[1][1], aS[1][ks], b);
                        } else {
                            dmma884(l[0][0], l[0][1], aF[0][ks], b);
                            dmma884(l[1][0], l[1][1], aF[1][ks], b);
                        }
                    }
                    double2 tn[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int off = (c ? off1 : off0) + 64 * h;
                        const double2 d = s2[off];
                        double2 yl = s2[kBoxD / 2 + off];
                        double2 z = s2[2 * (kBoxD / 2) + off];
                        admm_point2<MASKED>(prm, d.x, l[0][c], yl.x, z.x, tn[c].x, sL, sO);
                        admm_point2<MASKED>(prm, d.y, l[1][c], yl.y, z.y, tn[c].y, sL, sO);
                        s2[off] = tn[c];
                        s2[kBoxD / 2 + off] = yl;
                        s2[2 * (kBoxD / 2) + off] = z;
                    }
                    // next iteration's X1*F': acc[m][n] += T'(i,j) B2(j,k) C3(t,k); k-step c covers j = j0 + 2*tig + c
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
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&done[slot]);
                if (++slot == S) { slot = 0; ph ^= 1; }
            }
            if (++t == a.n3) { t = 0; ++jc; }
        }
        double* p = a.partM + ((size_t)it * a.part_slots + x) * kPartH * a.RS;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
                *reinterpret_cast<double2*>(p + (size_t)(warp * 16 + 2 * g + m) * a.RS + 8 * n + 2 * tig) =
                    make_double2(acc[m][n][0], acc[m][n][1]);
        sL = warp_sum(sL);
        sO = warp_sum(sO);
        if (lane == 0) { red[warp] = sL; red[8 + warp] = sO; }
      } else {
        // consumer warp whose 16 rows lie entirely outside the tensor: contributes zeros
        double* p = a.partM + ((size_t)it * a.part_slots + x) * kPartH * a.RS;
        for (int e = lane; e < 16 * a.RS; e += 32) p[(size_t)warp * 16 * a.RS + e] = 0.0;
        if (lane == 0) { red[warp] = 0.0; red[8 + warp] = 0.0; }
      }
    }
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        double sl = 0.0, so = 0.0;
        for (int w = 0; w < 8; ++w) { sl += red[w]; so += red[8 + w]; }
        a.norm_part[2 * blockIdx.x] = sl; a.norm_part[2 * blockIdx.x + 1] = so;
        if (a.dbg) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.dbg[2 * blockIdx.x + 1] = t_; }
        __threadfence();
        s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // Last CTA to finish (every CTA has read the iteration scalars long ago): fixed-order sum of the per-CTA
    // residual sums, then -- single rank -- the scalar step of the iteration; with several ranks the pair is
    // left in norms[] for the all-reduce and k_finalize does the scalar step.
    __threadfence();
    double sa = 0.0, sb = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kAdmmThreads) { sa += __ldcg(a.norm_part + 2 * i); sb += __ldcg(a.norm_part + 2 * i + 1); }
    block_sum2(sa, sb, red);
    if (threadIdx.x == 0) {
        *a.ticket = 0u;
        if (a.finalize) iter_finalize(a.st, sa, sb, a.errHist, a.errL, a.errO);
        else { a.norms[0] = sa; a.norms[1] = sb; red[0] = sa; red[1] = sb; }
    }
    if (!a.finalize && a.peers) {
        // peer exchange of the residual sums by this (last) CTA as self-validating words (kernels_xchg.cuh); then wait
        // for all ranks, sum in rank order -- every rank takes the same stopping decision -- and finalise the iteration
        __syncthreads();
        const unsigned epoch = a.xbase + (unsigned)a.st->k + 1u;
        if (threadIdx.x < (unsigned)a.nranks) {
            double* box = a.peers[threadIdx.x];
            ll_store(box + a.norm_off, red[0], epoch);
            ll_store(box + a.norm_off + 2, red[1], epoch);
        }
        if (threadIdx.x == 0) {
            const double ta = ll_sum_ranks(a.nslots, 8, a.nranks, epoch, &a.st->status);
            const double tb = ll_sum_ranks(a.nslots + 2, 8, a.nranks, epoch, &a.st->status);
            iter_finalize(a.st, ta, tb, a.errHist, a.errL, a.errO);
        }
    }
}

}  // namespace tritd
