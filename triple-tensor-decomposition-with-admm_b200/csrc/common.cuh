// Shared device helpers for libtritd (sm_100a only): DMMA, mbarrier, TMA, reductions.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace tritd {

// ---------------------------------------------------------------------------
// FP64 tensor-core MMA (SASS: DMMA.8x8x4).  Fragment ownership, lane = 4*g + tig:
//   A (8x4, row)  a  = A[g][tig]
//   B (4x8, col)  b  = B[tig][g]
//   C (8x8)       c0 = C[g][2*tig], c1 = C[g][2*tig+1]
// ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // not volatile: a pure function of its operands, so the scheduler may interleave independent chains
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// The permutation used wherever 8 MMA rows (or columns) are mapped onto 8 rows of a
// 128B-swizzled shared-memory box so that a quarter-warp of LDS.128 hits 8 distinct
// 16-byte bank groups: rho = {0,4,1,5,2,6,3,7}.
__device__ __forceinline__ int rho8(int g) { return (g >> 1) | ((g & 1) << 2); }

// Programmatic dependent launch: the kernels of the iteration are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs become resident and run their
// prologue (barrier set-up, descriptor prefetch) while this grid drains.  pdl_wait() returns once every grid this one
// depends on has completed and its memory is visible (a no-op for a plainly launched kernel); EVERY thread calls it
// before it reads anything an earlier kernel wrote and before it exits, so completion stays transitive along the
// chain.  pdl_trigger() lets the dependents be scheduled; they still wait for this grid's completion themselves.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------
// shared-memory addresses, mbarrier, TMA
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 3-D tiled TMA load (global -> shared), completion on an mbarrier.  SASS: UTMALDG.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 3-D tiled TMA store (shared -> global), bulk-group completion.  SASS: UTMASTG.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still have to READ their shared-memory source
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before a bulk store
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Read one 16-byte chunk (two doubles) of a [rows][16 doubles] box written by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B (box base 1024-byte aligned): chunk c of row r sits at c ^ (r & 7).
__device__ __forceinline__ double2 lds_swz128(const double* box, int row, int chunk) {
    const double2* p = reinterpret_cast<const double2*>(box) + row * 8 + (chunk ^ (row & 7));
    return *p;
}
__device__ __forceinline__ void sts_swz128(double* box, int row, int chunk, double2 v) {
    double2* p = reinterpret_cast<double2*>(box) + row * 8 + (chunk ^ (row & 7));
    *p = v;
}

// ---------------------------------------------------------------------------
// streaming global access: read-once / write-once data must not displace the
// L1-resident factor tiles
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream2(double* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---------------------------------------------------------------------------
// deterministic reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of two values; result valid in thread 0.  `red` holds 2*32 doubles.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) { red[w] = a; red[32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        double x = l < nw ? red[l] : 0.0, y = l < nw ? red[32 + l] : 0.0;
        a = warp_sum(x);
        b = warp_sum(y);
    }
}

}  // namespace tritd
