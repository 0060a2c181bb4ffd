// btd.cuh -- device side of the BlockTridiagonal drop-in.
//
// The reference's BlockTridiagonalLu{,Consistent} (src/somelinalg/banded/block_tridiagonal_lu_consistent.rs:98-218)
// is block Thomas with pivoting INSIDE diagonal blocks only and rejects singular diagonal blocks
// (:1267-1286).  Here a general block-tridiagonal system
//     L_{k-1} x_{k-1} + D_k x_k + U_k x_{k+1} = b_k ,  k = 0..nb-1          (block_tridiagonal.rs:3-19)
// is embedded in the block-bidiagonal form the pivoted condensation solver (abd.cuh) consumes, with nodes
// Y_j = [x_j ; x_{j+1}] (size nv = 2 bs), j = 0..nb-2, and block rows for j = 0..nb-3:
//     [ 0    -I      ] Y_j + [ I   0       ] Y_{j+1} = [ 0       ]     (consistency of the shared x_{j+1})
//     [ L_j  D_{j+1} ]       [ 0   U_{j+1} ]           [ b_{j+1} ]     (block row j+1)
// closed by the general boundary rows  [D_0 U_0] Y_0 = b_0  and  [L_{nb-2} D_{nb-1}] Y_{nb-2} = b_{nb-1}.
// Partial pivoting then runs across neighbouring block rows, so singular diagonal blocks are fine as long as the
// matrix itself is non-singular.
#pragma once
#include "rst_common.cuh"

namespace rst {

// lower/upper: [nb-1][bs][bs], diag: [nb][bs][bs] row-major.  A, B: [nb-2][nv][nv]
__global__ void __launch_bounds__(128) btd_embed_kernel(const double *__restrict__ lower, const double *__restrict__ diag,
                                                        const double *__restrict__ upper, double *__restrict__ A,
                                                        double *__restrict__ B, int64_t n_int, int bs) {
    const int nv = 2 * bs;
    const int64_t total = n_int * (int64_t)nv * nv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx / (nv * nv);
        const int e = (int)(idx - j * nv * nv);
        const int i = e / nv, c = e % nv;
        double a, b;
        if (i < bs) {  // consistency rows
            a = (c >= bs && c - bs == i) ? -1.0 : 0.0;
            b = (c < bs && c == i) ? 1.0 : 0.0;
        } else {       // block row j + 1
            const int ii = i - bs;
            a = c < bs ? lower[(j * bs + ii) * bs + c] : diag[((j + 1) * bs + ii) * bs + (c - bs)];
            b = c < bs ? 0.0 : upper[((j + 1) * bs + ii) * bs + (c - bs)];
        }
        A[idx] = a;
        B[idx] = b;
    }
}

// boundary rows: Ba = [D_0 U_0] (bs x nv), Bb = [L_{nb-2} D_{nb-1}] (bs x nv)
__global__ void btd_boundary_kernel(const double *lower, const double *diag, const double *upper, double *Ba, double *Bb,
                                    int64_t nb, int bs) {
    const int nv = 2 * bs;
    for (int idx = threadIdx.x; idx < bs * nv; idx += blockDim.x) {
        const int i = idx / nv, c = idx % nv;
        Ba[idx] = c < bs ? diag[i * bs + c] : upper[i * bs + (c - bs)];
        Bb[idx] = c < bs ? lower[((nb - 2) * bs + i) * bs + c] : diag[((nb - 1) * bs + i) * bs + (c - bs)];
    }
}

// rhs: r[j] = [0 ; b_{j+1}], ba = b_0, bb = b_{nb-1}
__global__ void __launch_bounds__(128) btd_rhs_kernel(const double *__restrict__ b, double *__restrict__ r,
                                                      double *__restrict__ ba, double *__restrict__ bb, int64_t nb,
                                                      int bs) {
    const int nv = 2 * bs;
    const int64_t total = (nb - 2) * (int64_t)nv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = idx / nv;
        const int i = (int)(idx - j * nv);
        r[idx] = i < bs ? 0.0 : b[(j + 1) * bs + (i - bs)];
    }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < bs; i += blockDim.x) { ba[i] = b[i]; bb[i] = b[(nb - 1) * bs + i]; }
}

// x_j = Y_j[0:bs] (j <= nb-2), x_{nb-1} = Y_{nb-2}[bs:]
__global__ void __launch_bounds__(128) btd_extract_kernel(const double *__restrict__ Z, double *__restrict__ x,
                                                          int64_t nb, int bs) {
    const int nv = 2 * bs;
    const int64_t total = nb * (int64_t)bs;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = idx / bs;
        const int i = (int)(idx - k * bs);
        x[idx] = k <= nb - 2 ? Z[k * nv + i] : Z[(nb - 2) * nv + bs + i];
    }
}

}  // namespace rst
