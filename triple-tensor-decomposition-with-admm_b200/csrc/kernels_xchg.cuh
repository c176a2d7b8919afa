// Exchange of the small per-iteration partials between the ranks of a mode-3 sharded solve over NVLink peer
// memory (one process per GPU; mailboxes mapped with CUDA IPC), replacing NCCL all-reduces whose latency
// (20-30 us each at 8 ranks) would dominate a 100-300 us iteration.
//
// Every rank owns a mailbox  [A: nranks x slotA | B: nranks x slotB | N: nranks x 8 | S: C3'C3 partials | flags].
// The exchanging kernels (k_upd for updates A and B, the last CTA of k_admm) copy a rank's partial (RHS_A, RHS_B or the
// two residual sums) into ITS slot of every rank's mailbox (remote stores through NVLink / NVSwitch, the local copy
// included) as self-validating 16-byte words (ll_store below: value + the epoch of the iteration); the same threads then
// poll the nranks slots of their own mailbox until every word shows the epoch and sum them in RANK ORDER from local
// memory: every rank forms bit-identical sums, so the replicated factors A and B stay bitwise equal across ranks
// without a broadcast.  (C3'C3, which nobody waits for on the spot, travels as plain data + a release flag.)  Slot reuse is safe without double buffering: between two
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

// Every spin in this library is bounded: a wait that has not been satisfied after kSpinLimit polls (several seconds --
// a peer rank died, or a protocol error) raises IterState::status = kStatusHang and falls through, so that the solve
// ends with TRITD_ERR_TIMEOUT instead of hanging the GPU.  The results of such a solve are meaningless.
constexpr int kStatusHang = 2;
constexpr unsigned kSpinLimit = 10u * 1000u * 1000u;   // x (40 ns sleep + one acquire load) ~ 10 s
__device__ __forceinline__ void spin_until_epoch(const unsigned* flag, unsigned epoch, int* status) {
    unsigned spins = 0;
    while ((int)(ld_acquire_sys_u32(flag) - epoch) < 0) {
        __nanosleep(40);
        if (++spins > kSpinLimit) { atomicExch(status, kStatusHang); break; }
    }
}

// CTA-wide wait until the `n` flags at `flags` (one per source rank) are all >= epoch; lanes of warp 0 poll one
// rank each (n <= 32).  Epochs only grow, so ">=" also covers a source rank that is already one exchange ahead.
__device__ __forceinline__ void cta_wait_ranks(const unsigned* flags, int n, unsigned epoch, int* status) {
    if (threadIdx.x < (unsigned)n) spin_until_epoch(flags + threadIdx.x, epoch, status);
    __syncthreads();
}

// Low-latency payload words (the scheme of NCCL's LL protocol): a double travels as ONE 16-byte store
// {lo, epoch, hi, epoch}, each 8-byte half carrying the epoch of the iteration, so the receiver needs no separate flag,
// and the sender no system-scope release (which would first wait for the acknowledgement of all its remote stores):
// a word is valid the moment both of its halves show the expected epoch.  Epochs grow by one per iteration and every
// region is rewritten only after all ranks have consumed it (see above), so the test is for equality.
__device__ __forceinline__ void ll_store(double* slot, double v, unsigned epoch) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(slot), "r"((unsigned)b), "r"(epoch), "r"((unsigned)(b >> 32))
                 : "memory");
}
__device__ __forceinline__ bool ll_try_load(const double* slot, unsigned epoch, double& v) {
    unsigned x, y, z, w;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(slot) : "memory");
    v = __longlong_as_double((long long)(((unsigned long long)z << 32) | x));
    return y == epoch && w == epoch;
}
// the `n` (<= 8) words slot0 + r * stride, r = 0..n-1, all valid; summed in rank order
__device__ __forceinline__ double ll_sum_ranks(const double* slot0, long stride, int n, unsigned epoch, int* status) {
    double t[8];
    unsigned spins = 0;
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int r = 0; r < 8; ++r) { t[r] = 0.0; if (r < n) ok &= ll_try_load(slot0 + r * stride, epoch, t[r]); }
        if (ok) break;
        __nanosleep(20);
        if (++spins > kSpinLimit) { atomicExch(status, kStatusHang); break; }
    }
    double v = t[0];
#pragma unroll
    for (int r = 1; r < 8; ++r) v += t[r];
    return v;
}

// The initial C3'C3 partial (tritd_problem_init) goes to every mailbox the same way update C's later ones do:
// parity 1, epoch xbase (the first update A waits for exactly that).
__global__ void __launch_bounds__(256) k_push_sc(const double* src, double* const* peers, long dst_off, long flag_area_off, long flag_idx,
                                                 int nranks, int n, unsigned epoch) {
    for (int r = 0; r < nranks; ++r) {
        double* dst = peers[r] + dst_off;
        for (int e = threadIdx.x; e < n; e += 256) dst[e] = src[e];
    }
    __syncthreads();
    if (threadIdx.x < (unsigned)nranks)
        st_release_sys_u32(reinterpret_cast<unsigned*>(peers[threadIdx.x] + flag_area_off) + flag_idx, epoch);
}

}  // namespace tritd
