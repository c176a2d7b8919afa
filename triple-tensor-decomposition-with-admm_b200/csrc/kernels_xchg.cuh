// Exchange of the small per-iteration partials between the ranks of a mode-3 sharded solve over NVLink peer
// memory (one process per GPU; mailboxes mapped with CUDA IPC), replacing NCCL all-reduces whose latency
// (20-30 us each at 8 ranks) would dominate a 100-300 us iteration.
//
// Every rank owns a mailbox  [A: nranks x slotA | B: nranks x slotB | N: nranks x 8 | flags: 3 rows of 8 u32].
// k_xchg_push copies a rank's partial ([RHS_A ; C3'C3], RHS_B or the two residual sums) into ITS slot of every
// rank's mailbox (remote stores through NVLink / NVSwitch, the local copy included) and then raises its flag
// in every mailbox to the epoch of the iteration (release at system scope).  Consumers (k_upd, k_finalize) wait
// until all nranks flags of their mailbox reached the epoch (acquire at system scope) and sum the nranks slots
// in RANK ORDER from local memory: every rank forms bit-identical sums, so the replicated factors A and B stay
// bitwise equal across ranks without a broadcast.  Slot reuse is safe without double buffering: between two
// pushes into the same region every rank passes the two other exchanges of the iteration, each of which waits
// for all ranks.
#pragma once
#include "common.cuh"
#include "kernels_fused.cuh"

namespace tritd {

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// CTA-wide wait until the `n` flags at `flags` (one per source rank) are all >= epoch; lanes of warp 0 poll one
// rank each (n <= 32).  Epochs only grow, so ">=" also covers a source rank that is already one exchange ahead.
__device__ __forceinline__ void cta_wait_ranks(const unsigned* flags, int n, unsigned epoch) {
    if (threadIdx.x < (unsigned)n)
        while ((int)(ld_acquire_sys_u32(flags + threadIdx.x) - epoch) < 0) __nanosleep(40);
    __syncthreads();
}

struct XchgPushArgs {
    const double* src;        // local payload, n doubles (n even)
    long n;
    double* const* peers;     // [nranks] mailbox base of every rank (own mailbox included), device array
    long dst_off;             // offset in doubles of this rank's slot inside a mailbox
    long flag_off;            // offset in doubles of the flag words of this exchange inside a mailbox
    int rank, nranks;
    unsigned xbase;           // epoch of iteration k of this solve = xbase + k + 1
    const IterState* st;
    unsigned* ticket;
};

__global__ void __launch_bounds__(256) k_xchg_push(const XchgPushArgs a) {
    if (a.st->stop) return;
    const unsigned epoch = a.xbase + (unsigned)a.st->k + 1u;
    const long n2 = a.n >> 1;
    const double2* src = reinterpret_cast<const double2*>(a.src);
    for (int q = 0; q < a.nranks; ++q) {
        const int r = (a.rank + 1 + q) % a.nranks;          // start with the neighbour: spreads the NVLink traffic
        double2* dst = reinterpret_cast<double2*>(a.peers[r] + a.dst_off);
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n2; i += (long)gridDim.x * 256) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < (unsigned)a.nranks) {
        unsigned* f = reinterpret_cast<unsigned*>(a.peers[threadIdx.x] + a.flag_off) + a.rank;
        st_release_sys_u32(f, epoch);
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
}

// errHist / mu / stopping rule after the residual sums of all ranks arrived (N>1 with the peer exchange):
// the pairs are summed in rank order, so every rank takes the same decision.
__global__ void __launch_bounds__(32) k_finalize_xchg(IterState* st, const double* slots, const unsigned* flags, int nranks,
                                                      unsigned xbase, double* errHist, double* errL, double* errO) {
    if (st->stop) return;
    const unsigned epoch = xbase + (unsigned)st->k + 1u;
    cta_wait_ranks(flags, nranks, epoch);
    if (threadIdx.x != 0) return;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < nranks; ++r) { a += __ldcg(slots + 8 * r); b += __ldcg(slots + 8 * r + 1); }
    iter_finalize(st, a, b, errHist, errL, errO);
}

}  // namespace tritd
