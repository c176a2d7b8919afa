// B200 micro-measurements that SURVEY.md says nobody has written down for this pool:
// FP64 DMMA (mma.sync m8n8k4) and DFMA peak, and plain FP64 streaming patterns of the fused kernel
// (4 reads + 5 writes).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dmma(double* out, int iters) {
    double c[CH][2];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_dfma(double* out, int iters) {
    double c[CH];
    for (int i = 0; i < CH; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA in the same warp: CH DMMAs + NF DFMAs per iteration, all chains independent.  If the two share one
// datapath the time is the SUM of the two pure loops, if they are separate pipes it is the MAX.
template <int CH, int NF>
__global__ void k_mix(double* out, int iters) {
    double c[CH > 0 ? CH : 1][2], f[NF > 0 ? NF : 1];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    for (int i = 0; i < NF; ++i) f[i] = i;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4, fa = 1.0 + threadIdx.x * 1e-9, fb = threadIdx.x * 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (CH > NF ? CH : NF); ++i) {
            if (i < CH) dmma(c[i][0], c[i][1], a, b);
            if (i < NF) f[i] = fma(f[i], fa, fb);
        }
    }
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_layout(const double* A, const double* B, double* C) {
    // C(8x8) = A(8x4, row-major) * B(4x8, stored B[k][n]) using the documented fragment ownership
    const int lane = threadIdx.x, g = lane >> 2, tig = lane & 3;
    double c0 = 0, c1 = 0;
    dmma(c0, c1, A[g * 4 + tig], B[tig * 8 + g]);
    C[g * 8 + 2 * tig] = c0; C[g * 8 + 2 * tig + 1] = c1;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void k_r4w5(const double2* __restrict__ a, double2* b, double2* c, double2* d, double2* o, double2* t, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 x = a[i], y = b[i], z = c[i], w = d[i];
        double2 s; s.x = x.x + y.x * 0.5 - z.x + w.x; s.y = x.y + y.y * 0.5 - z.y + w.y;
        b[i] = s; c[i] = make_double2(s.x * 0.5, s.y * 0.5); d[i] = make_double2(s.x - 1, s.y - 1); o[i] = make_double2(x.x - s.x, x.y - s.y);
        t[i] = make_double2(x.x + s.x, x.y + s.y);
    }
}
__global__ void k_readsum(const double2* __restrict__ in, double* out, size_t n2) {
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) { double2 v = in[i]; s += v.x + v.y; }
    if (s == 1.2345) out[0] = s;
}

template <typename F> float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    // layout check
    {
        std::vector<double> A(32), B(32), C(64), R(64, 0.0);
        for (int i = 0; i < 32; ++i) { A[i] = i + 1; B[i] = 0.5 * i - 3; }
        for (int m = 0; m < 8; ++m) for (int n = 0; n < 8; ++n) for (int k = 0; k < 4; ++k) R[m * 8 + n] += A[m * 4 + k] * B[k * 8 + n];
        double *dA, *dB, *dC; CK(cudaMalloc(&dA, 256)); CK(cudaMalloc(&dB, 256)); CK(cudaMalloc(&dC, 512));
        CK(cudaMemcpy(dA, A.data(), 256, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), 256, cudaMemcpyHostToDevice));
        k_layout<<<1, 32>>>(dA, dB, dC); CK(cudaMemcpy(C.data(), dC, 512, cudaMemcpyDeviceToHost));
        double md = 0; for (int i = 0; i < 64; ++i) md = fmax(md, fabs(C[i] - R[i]));
        printf("dmma m8n8k4 fragment layout check: max |diff| = %g  (%s)\n", md, md == 0 ? "OK" : "MISMATCH");
    }
    double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int blocks_per_sm : {1}) {
            int grid = p.multiProcessorCount * blocks_per_sm;
            float ms = time_ms([&] { k_dmma<8><<<grid, warps * 32>>>(out, iters); });
            double fl = (double)grid * warps * iters * 8 * 512.0;
            printf("DMMA 8 chains  warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
            ms = time_ms([&] { k_dmma<2><<<grid, warps * 32>>>(out, iters); });
            fl = (double)grid * warps * iters * 2 * 512.0;
            printf("DMMA 2 chains  warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
            ms = time_ms([&] { k_dfma<8><<<grid, warps * 32>>>(out, iters); });
            fl = (double)grid * warps * 32.0 * iters * 8 * 2.0;
            printf("DFMA 8 chains  warps/SM=%2d : %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        }
    }
    // do DMMA and DFMA share a datapath?  8 DMMA (8*512 flop/warp) + 8..64 DFMA (64 flop/warp each) per iteration
    {
        int grid = p.multiProcessorCount, warps = 16;
        auto report = [&](const char* nm, float ms, double nd, double nf) {
            double fd = (double)grid * warps * iters * nd * 512.0, ff = (double)grid * warps * 32.0 * iters * nf * 2.0;
            printf("MIX %-18s: %.3f ms  DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", nm, ms, fd / ms * 1e-9, ff / ms * 1e-9, (fd + ff) / ms * 1e-9);
        };
        report("8 dmma + 0 dfma", time_ms([&] { k_mix<8, 0><<<grid, warps * 32>>>(out, iters); }), 8, 0);
        report("0 dmma + 16 dfma", time_ms([&] { k_mix<0, 16><<<grid, warps * 32>>>(out, iters); }), 0, 16);
        report("8 dmma + 8 dfma", time_ms([&] { k_mix<8, 8><<<grid, warps * 32>>>(out, iters); }), 8, 8);
        report("8 dmma + 16 dfma", time_ms([&] { k_mix<8, 16><<<grid, warps * 32>>>(out, iters); }), 8, 16);
        report("8 dmma + 32 dfma", time_ms([&] { k_mix<8, 32><<<grid, warps * 32>>>(out, iters); }), 8, 32);
        report("4 dmma + 32 dfma", time_ms([&] { k_mix<4, 32><<<grid, warps * 32>>>(out, iters); }), 4, 32);
    }
    // sustained DMMA (about 2 s) to see clocks under power
    {
        int grid = p.multiProcessorCount;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int r = 0; r < 40; ++r) k_dmma<8><<<grid, 512>>>(out, iters * 4);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 40.0 * grid * 16 * (iters * 4.0) * 8 * 512.0;
        printf("DMMA sustained %.0f ms: %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    // streaming
    const size_t N = 23040000;     // 240 x 320 x 300
    double *a, *b, *c, *d, *o, *t;
    for (double** q : {&a, &b, &c, &d, &o, &t}) { CK(cudaMalloc(q, N * 8)); CK(cudaMemset(*q, 0, N * 8)); }
    for (int grid_mult : {2, 4, 8, 16}) {
        int grid = p.multiProcessorCount * grid_mult;
        float ms = time_ms([&] { k_copy<<<grid, 256>>>((double2*)a, (double2*)b, N / 2); }, 10);
        printf("copy  f64 N=23.04M grid=%4d: %.3f ms  %.0f GB/s\n", grid, ms, 16.0 * N / ms * 1e-6);
        ms = time_ms([&] { k_r4w5<<<grid, 256>>>((double2*)a, (double2*)b, (double2*)c, (double2*)d, (double2*)o, (double2*)t, N / 2); }, 10);
        printf("r4w5  f64 N=23.04M grid=%4d: %.3f ms  %.0f GB/s\n", grid, ms, 72.0 * N / ms * 1e-6);
        ms = time_ms([&] { k_readsum<<<grid, 256>>>((double2*)a, out, N / 2); }, 10);
        printf("read  f64 N=23.04M grid=%4d: %.3f ms  %.0f GB/s\n", grid, ms, 8.0 * N / ms * 1e-6);
    }
    return 0;
}
