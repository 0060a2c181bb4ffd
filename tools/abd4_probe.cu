// abd4_probe.cu -- stand-alone timing harness of the nv = 12 level-0 factor kernel (abd4.cuh) on random block rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo [-DABD4_X_...] \
//        -I rustedscithe_b200/csrc -I rustedscithe_b200/csrc/generated -o .probe/abd4_probe tools/abd4_probe.cu
// Experiments only (what-if switches change the arithmetic); the product path is the library.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "abd4.cuh"

using namespace rst;

__global__ void fill_kernel(double *A, int64_t n, int nv) {
    const int64_t total = n * nv * nv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long s = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
        s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
        double v = (double)(s >> 11) / 9007199254740992.0 - 0.5;
        const int e = (int)(i % (nv * nv));
        if (e / nv == e % nv) v -= 3.0;      // -(I + h J)-like: dominant diagonal next to the identity B
        A[i] = v;
    }
}

int main(int argc, char **argv) {
    constexpr int NV = 12;
    using Cf = Abd4Cfg<NV>;
    using Pk = AbdPack<NV>;
    const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000;
    const int S = 16;
    const int64_t nslabs = (n + S - 1) / S;
    AbdLevel L{};
    L.n = n; L.nslabs = nslabs; L.S = S; L.nv = NV; L.prefetch = 0;
    double *A, *An, *Bn, *Lm, *UEF; int8_t *piv; unsigned *flags;
    cudaMalloc(&A, (size_t)n * NV * NV * 8);
    cudaMalloc(&An, (size_t)nslabs * NV * NV * 8); cudaMalloc(&Bn, (size_t)nslabs * NV * NV * 8);
    cudaMalloc(&Lm, (size_t)n * Cf::LT * 8); cudaMalloc(&UEF, (size_t)n * Pk::UEF * 8); cudaMalloc(&piv, (size_t)n * Cf::PIV_STRIDE);
    cudaMalloc(&flags, 16); cudaMemset(flags, 0, 16);
    fill_kernel<<<148 * 8, 256>>>(A, n, NV);
    L.A = A; L.B = nullptr; L.A_next = An; L.B_next = Bn; L.Lm = Lm; L.UEF = UEF; L.piv = piv;
    // right-hand side riding along (RHS variant) against the forward sweep on the stored factors
    double *r, *ev1, *rn1, *ev2, *rn2;
    cudaMalloc(&r, (size_t)n * NV * 8); cudaMalloc(&ev1, (size_t)n * NV * 8); cudaMalloc(&ev2, (size_t)n * NV * 8);
    cudaMalloc(&rn1, (size_t)nslabs * NV * 8); cudaMalloc(&rn2, (size_t)nslabs * NV * 8);
    cudaMemset(ev1, 0, (size_t)n * NV * 8); cudaMemset(ev2, 0, (size_t)n * NV * 8);
    fill_kernel<<<148 * 8, 256>>>(r, n, 1);     // nv = 1: n * 1 * 1 ... fills n values; repeat for the rest
    fill_kernel<<<148 * 8, 256>>>(r + n, (n * (NV - 1)), 1);
    L.r = r; L.e = ev1; L.r_next = rn1;
    const int smem = 4 * (int)sizeof(Abd4Smem<NV, true>);
#ifdef ABD4_PROBE_RHS
    auto kern = abd_factor4_kernel<NV, true, true>;
#else
    auto kern = abd_factor4_kernel<NV, true>;
#endif
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const unsigned blocks = (unsigned)((nslabs + 3) / 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        kern<<<blocks, 128, smem>>>(L, flags);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    unsigned hf[4]; cudaMemcpy(hf, flags, 16, cudaMemcpyDeviceToHost);
    std::vector<double> h((size_t)nslabs * NV * NV);
    double cs = 0;
    cudaMemcpy(h.data(), An, h.size() * 8, cudaMemcpyDeviceToHost); for (double v : h) cs += v;
    cudaMemcpy(h.data(), Bn, h.size() * 8, cudaMemcpyDeviceToHost); for (double v : h) cs += v * 0.5;
    std::vector<double> hl((size_t)1000 * Cf::LT); cudaMemcpy(hl.data(), Lm + (size_t)5000 * Cf::LT, hl.size() * 8, cudaMemcpyDeviceToHost);
    double cl = 0; for (double v : hl) cl += v;
#ifdef ABD4_PROBE_RHS
    {
        AbdLevel L2 = L; L2.e = ev2; L2.r_next = rn2;
        abd_forward4_kernel<NV><<<(unsigned)((nslabs + 3) / 4), 128>>>(L2);
        cudaDeviceSynchronize();
        std::vector<double> a((size_t)n * NV), b((size_t)n * NV);
        cudaMemcpy(a.data(), ev1, a.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), ev2, b.size() * 8, cudaMemcpyDeviceToHost);
        double md = 0, mx = 0; size_t nb = 0;
        for (size_t i = 0; i < a.size(); ++i) { double d = fabs(a[i] - b[i]); if (d > md) md = d; if (fabs(b[i]) > mx) mx = fabs(b[i]); if (a[i] != b[i]) ++nb; }
        std::vector<double> c((size_t)nslabs * NV), d2((size_t)nslabs * NV);
        cudaMemcpy(c.data(), rn1, c.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(d2.data(), rn2, d2.size() * 8, cudaMemcpyDeviceToHost);
        double md2 = 0; size_t nb2 = 0;
        for (size_t i = 0; i < c.size(); ++i) { double d = fabs(c[i] - d2[i]); if (d > md2) md2 = d; if (c[i] != d2[i]) ++nb2; }
        printf("rhs column vs forward sweep: e max |diff| %.3e (max |e| %.3e, %zu of %zu words differ), r_next max |diff| %.3e (%zu of %zu differ)\n",
               md, mx, nb, a.size(), md2, nb2, c.size());
    }
#endif
    printf("err %s  n %lld  S %d  best %.4f ms   flags %u  checksum %.12e  Lt %.12e\n", cudaGetErrorString(cudaGetLastError()), (long long)n, S, best,
           hf[0], cs, cl);
    return 0;
}
