// bigd_probe.cu -- standalone timing harness of the nv = 64 factor kernel (bigd.cuh) on random block rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo [-DBIGD_TIMING] [-DBIGD_NO_DMMA] \
//        -I rustedscithe_b200/csrc -o gpurun_out/bigd_probe tools/bigd_probe.cu
// Experiments only (what-if switches change the arithmetic); the product path is the library.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "bigd.cuh"

using namespace rst;

int main(int argc, char **argv) {
    const int waves = argc > 1 ? atoi(argv[1]) : 8;
    const int S = 16, NV = 64;
    const int64_t nslabs = 148LL * waves, n = nslabs * S;
    std::vector<double> hA((size_t)n * NV * NV), hB((size_t)n * NV * NV);
    unsigned long long s = 88172645463325252ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0 - 0.5; };
    for (auto &v : hA) v = rnd();
    for (auto &v : hB) v = rnd();
    for (int64_t g = 0; g < n; ++g)
        for (int i = 0; i < NV; ++i) { hA[(g * NV + i) * NV + i] += 4.0; hB[(g * NV + i) * NV + i] -= 4.0; }
    AbdLevel L{};
    L.n = n; L.nslabs = nslabs; L.S = S; L.nv = NV; L.prefetch = 0;
    double *A, *B, *An, *Bn, *Lm, *UEF; int8_t *piv; unsigned *flags;
    cudaMalloc(&A, hA.size() * 8); cudaMalloc(&B, hB.size() * 8);
    cudaMalloc(&An, nslabs * NV * NV * 8); cudaMalloc(&Bn, nslabs * NV * NV * 8);
    cudaMalloc(&Lm, (size_t)n * NV * 2 * NV * 8); cudaMalloc(&UEF, (size_t)n * 3 * NV * NV * 8); cudaMalloc(&piv, (size_t)n * 2 * NV);
    cudaMalloc(&flags, 16); cudaMemset(flags, 0, 16);
    cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(B, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice);
    L.A = A; L.B = B; L.A_next = An; L.B_next = Bn; L.Lm = Lm; L.UEF = UEF; L.piv = piv;
    const int smem = (int)sizeof(BigdSmem);
    auto kern = bigd_factor_kernel<false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        kern<<<(unsigned)nslabs, BigdCfg::THREADS + 128, smem>>>(L, flags);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    unsigned hf[4]; cudaMemcpy(hf, flags, 16, cudaMemcpyDeviceToHost);
    std::vector<double> hAn(nslabs * NV * NV); cudaMemcpy(hAn.data(), An, hAn.size() * 8, cudaMemcpyDeviceToHost);
    double cs = 0; for (double v : hAn) cs += v;
    printf("err %s  slabs %lld  S %d  best %.3f ms  -> %.1f us per slab-wave, scaled to 15625 slabs: %.1f ms   flags %u  checksum %.12e\n",
           cudaGetErrorString(cudaGetLastError()), (long long)nslabs, S, best, best * 1e3 / waves, best / waves * (15625.0 / 148.0), hf[0], cs);
    return 0;
}
