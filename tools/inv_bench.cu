// Diagnostics: cycles of the ridge-system inversion alone, DFMA / bar.sync / rcp latencies.
#include <cstdio>
#include <vector>
#include "../triple-tensor-decomposition-with-admm_b200/csrc/kernels_update.cuh"
using namespace tritd;

template <int PQ> __global__ void k_inv_test(const double* S1, const double* S2, int R, int RS, double* out, long long* cyc) {
    __shared__ double sm[256];
    long long t0 = clock64();
    int bad = invert_ridge_system<PQ>(S1, S2, 1, 0, 1e-3, R, RS, out, sm);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = bad; }
}
__device__ long long g_tdbg[128];
template <int NBMAX> __global__ void k_inv_blocked(const double* S1, const double* S2, int R, int RS, double* out, long long* cyc) {
    extern __shared__ double smb[];
    long long t0 = clock64();
    if (threadIdx.x == 0) g_tdbg[127] = t0;
    int bad = invert_ridge_blocked<NBMAX>(S1, S2, 1, 0, 1e-3, R, RS, out, smb, g_tdbg);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = bad; }
}
__global__ void k_lat(double* out, long long* cyc, int n) {
    double x = out[threadIdx.x], y = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, y, 1e-9);
    long long t1 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t2 = clock64();
    double z = x;
    for (int i = 0; i < n; ++i) z = rcp_newton(z) + 1.5;
    long long t3 = clock64();
    __shared__ double s[64];
    for (int i = 0; i < n; ++i) { s[(threadIdx.x + i) & 63] = z; __syncthreads(); z += s[(i * 7) & 63]; }
    long long t4 = clock64();
    out[threadIdx.x] = x + z;
    if (threadIdx.x == 0) { cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main() {
    const int RS = 64;
    std::vector<double> h(RS * RS, 0.0);
    for (int i = 0; i < RS; ++i) for (int j = 0; j < RS; ++j) h[i * RS + j] = (i == j) ? 3.0 : 1.0 / (1 + abs(i - j));
    double *S1, *S2, *out; long long* cyc;
    cudaMalloc(&S1, RS * RS * 8); cudaMalloc(&S2, RS * RS * 8); cudaMalloc(&out, RS * RS * 8); cudaMalloc(&cyc, 64);
    double* out2; cudaMalloc(&out2, RS * RS * 8);
    cudaFuncSetAttribute(k_inv_blocked<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaMemcpy(S1, h.data(), RS * RS * 8, cudaMemcpyHostToDevice); cudaMemcpy(S2, h.data(), RS * RS * 8, cudaMemcpyHostToDevice);
    long long c[8];
    for (int rep = 0; rep < 3; ++rep)
    for (int R : {9, 16, 25, 36, 49, 64}) {
        if (R <= 16) k_inv_test<1><<<1, 256>>>(S1, S2, R, RS, out, cyc);
        else if (R <= 32) k_inv_test<2><<<1, 256>>>(S1, S2, R, RS, out, cyc);
        else if (R <= 48) k_inv_test<3><<<1, 256>>>(S1, S2, R, RS, out, cyc);
        else k_inv_test<4><<<1, 256>>>(S1, S2, R, RS, out, cyc);
        cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
        if (rep == 2) printf("R=%d: %lld cycles (%.0f / step) bad=%lld\n", R, c[0], (double)c[0] / R, c[1]);
        cudaMemset(out2, 0, RS * RS * 8);
        if (R <= 16) k_inv_blocked<2><<<1, 256, ridge_blocked_smem_doubles(R) * 8>>>(S1, S2, R, RS, out2, cyc);
        else if (R <= 32) k_inv_blocked<4><<<1, 256, ridge_blocked_smem_doubles(R) * 8>>>(S1, S2, R, RS, out2, cyc);
        else if (R <= 48) k_inv_blocked<6><<<1, 256, ridge_blocked_smem_doubles(R) * 8>>>(S1, S2, R, RS, out2, cyc);
        else k_inv_blocked<8><<<1, 256, ridge_blocked_smem_doubles(R) * 8>>>(S1, S2, R, RS, out2, cyc);
        cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
        if (rep == 2) {
            std::vector<double> a(RS * RS), b(RS * RS);
            cudaMemcpy(a.data(), out, RS * RS * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), out2, RS * RS * 8, cudaMemcpyDeviceToHost);
            double md = 0, mx = 0;
            for (int i = 0; i < R; ++i) for (int j = 0; j < R; ++j) { md = fmax(md, fabs(a[i * RS + j] - b[i * RS + j])); mx = fmax(mx, fabs(a[i * RS + j])); }
            long long td[128]; cudaMemcpyFromSymbol(td, g_tdbg, sizeof(td));
            printf("   stamps (cycles since start; warp0 / warp1): ");
            for (int sl = 0; sl < 10; ++sl) printf("[%d] %lld/%lld ", sl, td[sl * 8] - td[127], td[sl * 8 + 1] - td[127]);
            printf("\n");
            printf("   blocked: %lld cycles (%.0f / block step) cond=%lld  max|scalar - blocked| = %.3e (max |inv| %.3e)  %s\n", c[0], (double)c[0] / ((R + 7) / 8), c[1], md, mx, cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int t : {32, 256, 1024}) {
        k_lat<<<1, t>>>(out, cyc, 1000);
        cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
        printf("threads=%d: dfma chain %.1f clk, __syncthreads %.1f clk, rcp_newton+add %.1f clk, sts+bar+lds+add %.1f clk\n", t, c[0] / 1000.0, c[1] / 1000.0, c[2] / 1000.0, c[3] / 1000.0);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
