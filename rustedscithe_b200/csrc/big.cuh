// big.cuh -- "block per slab" kernels of the pivoted condensation solver for wide blocks (16 < nv <= 64, config C5).
//
// Same algorithm and the SAME factor layout as abd.cuh (Lm [g][k][row], UEF [g][c][k], piv [g][row]); what changes is
// the mapping: the stacked 2nv x 3nv panel [M | P | Q] of one slab (128 x 192 doubles = 192 KB at nv = 64) lives in the
// shared memory of ONE CTA, nv is a run-time value, every elimination step is a block-wide arg-max followed by a
// rank-1 update of the remaining panel.  This is the first correct wide-block path (north_star config 5); it is
// shared-memory-bandwidth bound (3 smem accesses per FMA) -- register tiling and FP64 DMMA for the rank-nv update are the
// next step.
#pragma once
#include "abd.cuh"

namespace rst {

constexpr int BIG_THREADS = 256;
constexpr int BIG_MAX_NV = 64;

__host__ __device__ inline size_t big_factor_smem_bytes(int nv) {
    const int R2 = 2 * nv, WP = 3 * nv + 1;
    return (size_t)R2 * WP * sizeof(double) + (size_t)R2 * sizeof(double) + (size_t)(3 * R2 + 64) * sizeof(int);
}

// block-wide arg-max over `val[r]`, r < R2 (val < 0 = not a candidate); result (row) returned to every thread
__device__ __forceinline__ int block_argmax(double v, int r, double *s_val, int *s_idx) {
    int idx = r;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_xor_sync(RST_FULL_MASK, v, off);
        const int oi = __shfl_xor_sync(RST_FULL_MASK, idx, off);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) { s_val[warp] = v; s_idx[warp] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = s_val[0];
        int bi = s_idx[0];
        for (int w = 1; w < nw; ++w)
            if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
        s_idx[32] = bi;
        s_val[32] = bv;
    }
    __syncthreads();
    return s_idx[32];
}

template <bool IDENT_B>
__global__ void __launch_bounds__(BIG_THREADS) big_factor_kernel(AbdLevel L, int nv, unsigned *flags) {
    extern __shared__ __align__(16) unsigned char big_smem[];
    const int R2 = 2 * nv, W = 3 * nv, WP = W + 1;
    double *panel = reinterpret_cast<double *>(big_smem);            // [R2][WP]
    double *lcol = panel + (size_t)R2 * WP;                           // [R2] multipliers of the current pivot
    int *pv = reinterpret_cast<int *>(lcol + R2);                     // [R2] pivot step of a row, -1 = survivor
    int *carry = pv + R2;                                             // [R2] 1 = row carries the condensed row
    int *slot_of = carry + R2;                                        // [nv] row slot that receives new block row t
    __shared__ double s_val[33];
    __shared__ int s_idx[33];
    const int tid = threadIdx.x;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);

    for (int idx = tid; idx < R2 * WP; idx += blockDim.x) panel[idx] = 0.0;
    for (int r = tid; r < R2; r += blockDim.x) carry[r] = r < nv ? 1 : 0;
    __syncthreads();
    for (int idx = tid; idx < nv * nv; idx += blockDim.x) {
        const int r = idx / nv, c = idx - r * nv;
        panel[r * WP + nv + c] = L.A[g0 * nv * nv + idx];                                           // P = A
        panel[r * WP + c] = IDENT_B ? (r == c ? 1.0 : 0.0) : L.B[g0 * nv * nv + idx];               // M = B
    }
    __syncthreads();

    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        if (tid == 0) {
            int t = 0;
            for (int r = 0; r < R2; ++r)
                if (!carry[r]) slot_of[t++] = r;
        }
        for (int r = tid; r < R2; r += blockDim.x) pv[r] = -1;
        __syncthreads();
        for (int idx = tid; idx < nv * nv; idx += blockDim.x) {
            const int t = idx / nv, c = idx - t * nv;
            const int r = slot_of[t];
            panel[r * WP + c] = L.A[g * nv * nv + idx];                                             // M = A
            panel[r * WP + nv + c] = 0.0;                                                           // P = 0
            panel[r * WP + 2 * nv + c] = IDENT_B ? (t == c ? 1.0 : 0.0) : L.B[g * nv * nv + idx];   // Q = B
        }
        __syncthreads();
        for (int k = 0; k < nv; ++k) {
            const double v = (tid < R2 && pv[tid] < 0) ? fabs(panel[tid * WP + k]) : -1.0;
            const int p = block_argmax(v, tid < R2 ? tid : R2, s_val, s_idx);
            const double pivval = panel[p * WP + k];
            if (tid == 0 && !(fabs(pivval) > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
            if (tid < R2) {
                double l = 0.0;
                if (tid != p && pv[tid] < 0) {
                    l = panel[tid * WP + k] / pivval;
                    panel[tid * WP + k] = l;
                }
                lcol[tid] = l;
            }
            __syncthreads();
            if (tid == 0) pv[p] = k;
            const int ncol = W - k - 1;
            for (int idx = tid; idx < R2 * ncol; idx += blockDim.x) {
                const int r = idx / ncol, c = k + 1 + (idx - r * ncol);
                const double l = lcol[r];
                if (l != 0.0) panel[r * WP + c] = fma(-l, panel[p * WP + c], panel[r * WP + c]);
            }
            __syncthreads();
        }
        // factors of this step (same layout as abd.cuh)
        double *lm = L.Lm + g * (int64_t)(nv * R2);
        for (int idx = tid; idx < nv * R2; idx += blockDim.x) {
            const int k = idx / R2, r = idx - k * R2;
            st_cs(lm + idx, panel[r * WP + k]);
        }
        for (int r = tid; r < R2; r += blockDim.x) L.piv[g * R2 + r] = (int8_t)pv[r];
        double *uef = L.UEF + g * (int64_t)(3 * nv * nv);
        for (int idx = tid; idx < R2 * W; idx += blockDim.x) {
            const int r = idx / W, c = idx - r * W;
            const int k = pv[r];
            if (k >= 0) st_cs(uef + (int64_t)c * nv + k, panel[r * WP + c]);
        }
        __syncthreads();
        // survivors carry the condensed row: M <- Q, Q <- 0
        for (int idx = tid; idx < R2 * nv; idx += blockDim.x) {
            const int r = idx / nv, c = idx - r * nv;
            if (pv[r] < 0) {
                panel[r * WP + c] = panel[r * WP + 2 * nv + c];
                panel[r * WP + 2 * nv + c] = 0.0;
            }
        }
        for (int r = tid; r < R2; r += blockDim.x) carry[r] = pv[r] < 0 ? 1 : 0;
        __syncthreads();
    }
    if (tid == 0) {
        int t = 0;
        for (int r = 0; r < R2; ++r)
            if (carry[r]) slot_of[t++] = r;
    }
    __syncthreads();
    for (int idx = tid; idx < nv * nv; idx += blockDim.x) {
        const int t = idx / nv, c = idx - t * nv;
        const int r = slot_of[t];
        L.A_next[slab * nv * nv + idx] = panel[r * WP + nv + c];
        L.B_next[slab * nv * nv + idx] = panel[r * WP + c];
    }
}

// forward sweep: apply the recorded eliminations to the right-hand side (one CTA per slab)
__global__ void __launch_bounds__(128) big_forward_kernel(AbdLevel L, int nv) {
    __shared__ double w[2 * BIG_MAX_NV];
    __shared__ int pv[2 * BIG_MAX_NV], carry[2 * BIG_MAX_NV], slot_of[BIG_MAX_NV], row_of[BIG_MAX_NV];
    const int R2 = 2 * nv;
    const int tid = threadIdx.x;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    if (tid < R2) {
        carry[tid] = tid < nv ? 1 : 0;
        w[tid] = tid < nv ? L.r[g0 * nv + tid] : 0.0;
    }
    __syncthreads();
    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        if (tid == 0) {
            int t = 0;
            for (int r = 0; r < R2; ++r)
                if (!carry[r]) slot_of[t++] = r;
        }
        if (tid < R2) {
            pv[tid] = (int)L.piv[g * R2 + tid];
            if (pv[tid] >= 0) row_of[pv[tid]] = tid;
        }
        __syncthreads();
        if (tid < nv) w[slot_of[tid]] = L.r[g * nv + tid];
        __syncthreads();
        const double *lm = L.Lm + g * (int64_t)(nv * R2);
        for (int k = 0; k < nv; ++k) {
            const double wp = w[row_of[k]];
            __syncthreads();
            if (tid < R2 && (pv[tid] < 0 || pv[tid] > k)) w[tid] = fma(-ld_cs(lm + (int64_t)k * R2 + tid), wp, w[tid]);
            __syncthreads();
        }
        if (tid < R2) {
            if (pv[tid] >= 0) L.e[g * nv + pv[tid]] = w[tid];
            carry[tid] = pv[tid] < 0 ? 1 : 0;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int t = 0;
        for (int r = 0; r < R2; ++r)
            if (carry[r]) slot_of[t++] = r;
    }
    __syncthreads();
    if (tid < nv) L.r_next[slab * nv + tid] = w[slot_of[tid]];
}

// backward sweep: interior nodes of a slab from its end nodes (one CTA per slab, thread k = pivot row k)
__global__ void __launch_bounds__(64) big_backward_kernel(AbdLevel L, int nv) {
    __shared__ double Ys[BIG_MAX_NV], Yn[BIG_MAX_NV], ybc;
    const int k = threadIdx.x;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    if (k < nv) {
        Ys[k] = L.Z_next[slab * nv + k];
        Yn[k] = L.Z_next[(slab + 1) * nv + k];
        L.Z[g0 * nv + k] = Ys[k];
        if (slab == (int64_t)L.nslabs - 1) L.Z[L.n * nv + k] = Yn[k];
    }
    __syncthreads();
    for (int i = len - 1; i >= 1; --i) {
        const int64_t g = g0 + i;
        const double *u = L.UEF + g * (int64_t)(3 * nv * nv) + k;
        double t = 0.0, inv = 1.0;
        if (k < nv) {
            t = L.e[g * nv + k];
            for (int c = 0; c < nv; ++c) {
                t = fma(-ld_cs(u + (int64_t)(nv + c) * nv), Ys[c], t);
                t = fma(-ld_cs(u + (int64_t)(2 * nv + c) * nv), Yn[c], t);
            }
            inv = 1.0 / u[(int64_t)k * nv];
        }
        __syncthreads();  // every thread has finished reading the old Yn
        for (int kk = nv - 1; kk >= 0; --kk) {
            if (k == kk) { ybc = t * inv; Yn[kk] = ybc; }
            __syncthreads();
            if (k < kk) t = fma(-u[(int64_t)kk * nv], ybc, t);
            __syncthreads();
        }
        if (k < nv) L.Z[g * nv + k] = Yn[k];
        __syncthreads();
    }
}

// closure with pinned boundary components, run-time nv (abd_top_solve_kernel with K sized for nv <= 64)
__global__ void __launch_bounds__(32) big_top_solve_kernel(const double *C, const double *D, const double *g, double *Z,
                                                           int nv, unsigned long long left_mask,
                                                           unsigned long long right_mask, unsigned *flags) {
    __shared__ double K[BIG_MAX_NV][BIG_MAX_NV + 1];
    __shared__ int colmap[BIG_MAX_NV];
    const int lane = threadIdx.x;
    if (lane == 0) {
        int nc = 0;
        for (int c = 0; c < nv; ++c)
            if (!((left_mask >> c) & 1ull)) { if (nc < nv) colmap[nc] = c; ++nc; }
        for (int c = 0; c < nv; ++c)
            if (!((right_mask >> c) & 1ull)) { if (nc < nv) colmap[nc] = nv + c; ++nc; }
        if (nc != nv) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
    }
    __syncwarp();
    for (int idx = lane; idx < nv * nv; idx += 32) {
        const int row = idx / nv, cc = idx % nv;
        const int cm = colmap[cc];
        K[row][cc] = cm < nv ? C[row * nv + cm] : D[row * nv + (cm - nv)];
    }
    for (int row = lane; row < nv; row += 32) K[row][nv] = g[row];
    __syncwarp();
    for (int k = 0; k < nv; ++k) {
        double best = -1.0;
        int bi = k;
        for (int row = k + lane; row < nv; row += 32) {
            const double v = fabs(K[row][k]);
            if (v > best) { best = v; bi = row; }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ob = __shfl_xor_sync(RST_FULL_MASK, best, off);
            const int oi = __shfl_xor_sync(RST_FULL_MASK, bi, off);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (!(best > 0.0) && lane == 0) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
        if (bi != k)
            for (int c = lane; c <= nv; c += 32) { const double tmp = K[k][c]; K[k][c] = K[bi][c]; K[bi][c] = tmp; }
        __syncwarp();
        const double pivval = K[k][k];
        for (int row = k + 1 + lane; row < nv; row += 32) {
            const double l = K[row][k] / pivval;
            for (int c = k + 1; c <= nv; ++c) K[row][c] = fma(-l, K[k][c], K[row][c]);
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (int row = nv - 1; row >= 0; --row) {
            double s = K[row][nv];
            for (int c = row + 1; c < nv; ++c) s = fma(-K[row][c], K[c][nv], s);
            K[row][nv] = s / K[row][row];
        }
        for (int c = 0; c < 2 * nv; ++c) Z[c] = 0.0;
        for (int cc = 0; cc < nv; ++cc) Z[colmap[cc]] = K[cc][nv];
    }
}

}  // namespace rst
