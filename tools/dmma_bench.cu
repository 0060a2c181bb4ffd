// FP64 tensor-core (DMMA, mma.sync ... f64) vs FP64 FMA-pipe microbenchmark for sm_100a.
//
// north_star: "FP64 DMMA tensor cores are used only if ncu shows they beat the FP64 FMA pipe at the config's block
// size".  This tool measures the two issue rates the decision rests on (tools/dfma_bench.cu measures DFMA alone):
//   * mma.sync.aligned.m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 .f64 throughput with 1..16 warps per SM and 1..8
//     independent accumulator tiles per warp, and the dependent-issue latency of one accumulator chain,
//   * DFMA throughput in the same harness,
//   * both interleaved in one warp (do they share a pipe?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_bench tools/dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void __launch_bounds__(512, 1) k884(double *out, int iters, double x) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    const double a = x, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int K>
__global__ void __launch_bounds__(512, 1) k168(double *out, int iters, double x) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x + i * 1e-12;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-9 * threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if constexpr (K == 4) { double aa[2] = {a[0], a[1]}; dmma1684(c[i], aa, b[0]); }
            else if constexpr (K == 8) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; dmma1688(c[i], aa, bb); }
            else dmma16816(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void __launch_bounds__(512, 1) kfma(double *out, int iters, double x) {
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
// NF DFMA and NM m8n8k4 DMMA per iteration, independent chains, same warp
template <int NF, int NM>
__global__ void __launch_bounds__(512, 1) kmix(double *out, int iters, double x) {
    double f[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < NM; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    const double a = x, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
            if (i < NM) dmma884(c0[i], c1[i], a, b);
            if (i < NF) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static cudaEvent_t e0, e1;
template <class F>
static float timed(F launch) {
    launch(50);
    cudaEventRecord(e0);
    launch(0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double *out;
    cudaMalloc(&out, 148 * 512 * 8);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int clk_khz = 1965000;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("SM clock (attribute) %.3f GHz\n", ghz);
    const int iters = 20000;
#define REPORT(name, fma_per_instr, ilp, launchexpr)                                                                      \
    for (int warps = 1; warps <= 16; warps *= 2) {                                                                        \
        const float ms = timed([&](int it) { const int n_it = it ? it : iters; launchexpr; });                            \
        const double instr = 148.0 * warps * (double)(ilp) * iters;                                                       \
        const double fma = instr * (fma_per_instr);                                                                       \
        printf("%-22s ilp %d warps/SM %2d: %7.2f TFLOP/s  %6.1f FMA/clk/SM  %6.2f clk per warp-instr per SM\n", name, ilp, warps,   \
               2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / (ghz * 1e9), (ms * 1e-3) * ghz * 1e9 / (instr / 148.0));     \
    }
    REPORT("dfma", 32.0, 8, (kfma<8><<<148, warps * 32>>>(out, n_it, 0.999)))
    REPORT("dmma m8n8k4", 256.0, 1, (k884<1><<<148, warps * 32>>>(out, n_it, 0.999)))
    REPORT("dmma m8n8k4", 256.0, 4, (k884<4><<<148, warps * 32>>>(out, n_it, 0.999)))
    REPORT("dmma m8n8k4", 256.0, 8, (k884<8><<<148, warps * 32>>>(out, n_it, 0.999)))
    REPORT("dmma m16n8k4", 512.0, 4, (k168<4, 4><<<148, warps * 32>>>(out, n_it, 0.999)))
    REPORT("dmma m16n8k8", 1024.0, 4, (k168<4, 8><<<148, warps * 32>>>(out, n_it, 0.999)))
    REPORT("dmma m16n8k16", 2048.0, 4, (k168<4, 16><<<148, warps * 32>>>(out, n_it, 0.999)))
    {   // latency of a dependent chain, one warp per SM
        float ms = timed([&](int it) { k884<1><<<148, 32>>>(out, it ? it : 200000, 0.999); });
        printf("dependent m8n8k4 latency: %.1f cycles\n", ms * 1e-3 * ghz * 1e9 / 200000 * (50.0 / 50.0));
        ms = timed([&](int it) { kfma<1><<<148, 32>>>(out, it ? it : 200000, 0.999); });
        printf("dependent dfma latency:   %.1f cycles\n", ms * 1e-3 * ghz * 1e9 / 200000);
    }
    // shared pipe?  8 warps/SM: 8 DFMA alone, 4 DMMA alone, both together
    for (int warps = 4; warps <= 16; warps *= 2) {
        const float a = timed([&](int it) { kmix<8, 0><<<148, warps * 32>>>(out, it ? it : iters, 0.999); });
        const float b = timed([&](int it) { kmix<0, 4><<<148, warps * 32>>>(out, it ? it : iters, 0.999); });
        const float c = timed([&](int it) { kmix<8, 4><<<148, warps * 32>>>(out, it ? it : iters, 0.999); });
        printf("mix warps/SM %2d: 8 dfma %.3f ms, 4 dmma884 %.3f ms, both %.3f ms (sum %.3f)\n", warps, a, b, c, a + b);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
