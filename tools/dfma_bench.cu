// DFMA / LDS issue-rate microbenchmark for sm_100a (decides the FMA-vs-DMMA question of config C5).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(512, 1) dfma_kernel(double *out, int iters, double x) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    const double m = x, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *out;
    cudaMalloc(&out, 148 * 4 * 512 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 20000;
        dfma_kernel<16><<<148, warps * 32>>>(out, 100, 0.999);
        cudaEventRecord(e0);
        dfma_kernel<16><<<148, warps * 32>>>(out, iters, 0.999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * warps * 32 * 16.0 * iters;
        printf("warps/SM %2d: %.2f TFMA/s = %.2f TFLOP/s, %.1f DFMA/clk/SM at 1.965 GHz\n", warps, fma / ms / 1e9, 2 * fma / ms / 1e9,
               fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    {   // dependent-issue latency: one warp per SM, one accumulator
        const int iters = 200000;
        dfma_kernel<1><<<148, 32>>>(out, 100, 0.999);
        cudaEventRecord(e0);
        dfma_kernel<1><<<148, 32>>>(out, iters, 0.999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("dependent DFMA latency: %.1f cycles at 1.965 GHz\n", ms * 1e-3 * 1.965e9 / iters);
        dfma_kernel<2><<<148, 32>>>(out, iters, 0.999);
        cudaEventRecord(e0);
        dfma_kernel<2><<<148, 32>>>(out, iters, 0.999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("2 independent DFMA per iteration: %.1f cycles per iteration\n", ms * 1e-3 * 1.965e9 / iters);
        dfma_kernel<8><<<148, 32>>>(out, iters, 0.999);
        cudaEventRecord(e0);
        dfma_kernel<8><<<148, 32>>>(out, iters, 0.999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("8 independent DFMA per iteration (1 warp): %.1f cycles per iteration\n", ms * 1e-3 * 1.965e9 / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
