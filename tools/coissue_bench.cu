// coissue_bench.cu -- does a scheduler that streams FP64 tensor-core instructions still issue other warps' instructions?
//
// One CTA per SM.  Warps 0-3 (one per scheduler) run a "victim": a dependent integer chain (IMAD), a dependent
// shared-memory round trip chain (STS -> LDS), or a dependent warp-reduce chain (REDUX).  Warps 4..4+4*H-1 are "hogs":
// back-to-back independent DMMA m8n8k4 (8 accumulator tiles) or DFMA streams.  The victim's time per step with and without
// the hogs says how much of the scheduler a math stream leaves to a latency-bound co-resident warp (bigd.cuh's panel warps).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/coissue_bench tools/coissue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// victim: 0 = integer chain, 1 = STS/LDS chain, 2 = REDUX chain, 3 = dependent DFMA, 4 = independent DFMA, 5 = pivot-like step;  hog: 0 = none, 1 = DMMA, 2 = DFMA
template <int VICTIM, int HOG>
__global__ void __launch_bounds__(1024, 1) k(long long *cycles, double *sink, int steps, int hogs_per_sched, double x) {
    __shared__ volatile int flag;
    __shared__ int box[4][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    if (warp < 4) {
        unsigned v = lane + 1;
        double dv = 1.0 + lane, dw[3] = {1.0, 2.0, 3.0};
        const long long t0 = clock64();
        for (int i = 0; i < steps; ++i) {
            if constexpr (VICTIM == 0) {
#pragma unroll
                for (int u = 0; u < 16; ++u) v = (v ^ (v >> 3)) * 1664525u + (unsigned)i;
            } else if constexpr (VICTIM == 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    box[warp][lane] = (int)v;
                    __syncwarp();
                    v = (unsigned)box[warp][(lane + 1) & 31] + 1u;
                    __syncwarp();
                }
            } else if constexpr (VICTIM == 2) {
#pragma unroll
                for (int u = 0; u < 4; ++u) v = __reduce_max_sync(0xffffffffu, v ^ (unsigned)u) + (unsigned)lane;
            } else if constexpr (VICTIM == 3) {      // dependent DFMA chain
#pragma unroll
                for (int u = 0; u < 4; ++u) dv = fma(dv, x, 1e-9);
            } else if constexpr (VICTIM == 4) {      // 4 independent DFMA per step
                dv = fma(dv, x, 1e-9); dw[0] = fma(dw[0], x, 1e-9); dw[1] = fma(dw[1], x, 1e-9); dw[2] = fma(dw[2], x, 1e-9);
            } else {                                 // a pivot-like step: REDUX -> STS -> LDS -> 2 dependent DFMA
                v = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(dv) ^ (unsigned)lane);
                box[warp][lane] = (int)v;
                __syncwarp();
                const int r = box[warp][v & 31];
                __syncwarp();
                dv = fma(dv, x, (double)r * 1e-30);
                dv = fma(dv, x, 1e-9);
            }
        }
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x * 4 + warp] = t1 - t0;
        sink[blockIdx.x * blockDim.x + threadIdx.x] = (double)v + dv + dw[0] + dw[1] + dw[2];
        __threadfence_block();
        if (lane == 0) atomicAdd((int *)&flag, 1);
    } else if (warp < 4 + 4 * hogs_per_sched && HOG != 0) {
        double c0[8], c1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
        const double a = x, b = 1e-9 * threadIdx.x;
        while (flag < 4) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if constexpr (HOG == 1) dmma884(c0[i], c1[i], a, b);
                    else { c0[i] = fma(c0[i], a, b); c1[i] = fma(c1[i], a, b); }
                }
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
        sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

template <int VICTIM, int HOG>
static double run(int hogs, int per_step) {
    const int steps = 20000;
    long long *cyc; double *sink;
    cudaMalloc(&cyc, 148 * 4 * sizeof(long long)); cudaMalloc(&sink, 148 * 1024 * sizeof(double));
    k<VICTIM, HOG><<<148, 1024>>>(cyc, sink, steps, hogs, 1.0000001);
    cudaDeviceSynchronize();
    long long h[148 * 4]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (long long v : h) s += (double)v;
    cudaFree(cyc); cudaFree(sink);
    return s / (148 * 4) / steps / per_step;
}

int main() {
    printf("cycles per dependent victim operation (one victim warp per scheduler), by co-resident math stream\n");
    printf("%-34s %10s %10s %10s %10s %10s\n", "victim \\ hogs per scheduler", "none", "1 DMMA", "4 DMMA", "1 DFMA", "4 DFMA");
    printf("%-34s %10.1f %10.1f %10.1f %10.1f %10.1f\n", "IMAD (dependent)", run<0, 0>(0, 16), run<0, 1>(1, 16), run<0, 1>(4, 16), run<0, 2>(1, 16), run<0, 2>(4, 16));
    printf("%-34s %10.1f %10.1f %10.1f %10.1f %10.1f\n", "STS -> LDS round trip", run<1, 0>(0, 4), run<1, 1>(1, 4), run<1, 1>(4, 4), run<1, 2>(1, 4), run<1, 2>(4, 4));
    printf("%-34s %10.1f %10.1f %10.1f %10.1f %10.1f\n", "REDUX (dependent)", run<2, 0>(0, 4), run<2, 1>(1, 4), run<2, 1>(4, 4), run<2, 2>(1, 4), run<2, 2>(4, 4));
    printf("%-34s %10.1f %10.1f %10.1f %10.1f %10.1f\n", "DFMA (dependent)", run<3, 0>(0, 4), run<3, 1>(1, 4), run<3, 1>(4, 4), run<3, 2>(1, 4), run<3, 2>(4, 4));
    printf("%-34s %10.1f %10.1f %10.1f %10.1f %10.1f\n", "DFMA (4 independent, per instr)", run<4, 0>(0, 4), run<4, 1>(1, 4), run<4, 1>(4, 4), run<4, 2>(1, 4), run<4, 2>(4, 4));
    printf("%-34s %10.1f %10.1f %10.1f %10.1f %10.1f\n", "pivot step (REDUX,STS,LDS,2 DFMA)", run<5, 0>(0, 1), run<5, 1>(1, 1), run<5, 1>(4, 1), run<5, 2>(1, 1), run<5, 2>(4, 1));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
