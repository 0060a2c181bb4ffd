// assemble.cuh -- fused residual + Jacobian-block assembly over all mesh intervals (north_star
// subsystem 1).  One thread owns one interval j; the per-point mathematics comes from the struct the
// CUDA-C emitter generated (rustedscithe_b200/emit/cuda_emitter.py), the discretisation is restated
// from the reference:
//   forward    r[j,i] = Y[j+1][i] - Y[j][i] - h_j f_i(t_j, Y[j])
//              (src/symbolic/symbolic_functions_BVP.rs:5486-5490, numeric_discretization.rs:236)
//              A_j = dr_j/dY_j = -I - h_j f_y(t_j, Y_j),  B_j = dr_j/dY_{j+1} = +I   (:329-341)
//   trapezoid  r = Y[j+1] - Y[j] - h/2 (f(t_j,Y_j) + f(t*,Y_{j+1})), t* = t_j on the symbolic route
//              (symbolic_functions_BVP.rs:5250-5253), t_{j+1} on the closure route
//              (numeric_discretization.rs:223-232);  A_j = -I - h/2 f_y(Y_j), B_j = +I - h/2 f_y(Y_{j+1})
// State is the FULL (N+1) x nv node-major array with boundary slots pinned (SURVEY.md hard part 2),
// J is written directly in the block layout the solver consumes: A[j][nv][nv] row-major (B implicit
// identity for the forward scheme).
#pragma once
#include "rst_common.cuh"

namespace rst {

struct AsmArgs {
    const double *Y;       // [(n+1)][nv]
    double *F;             // [n][nv]
    double *A;             // [n][nv][nv]   (nullptr when only F is wanted)
    double *B;             // [n][nv][nv]   trapezoid only
    const double *mesh;    // [n+1] or nullptr for a uniform mesh
    const double *params;  // [np] device
    double t0, h;          // uniform mesh: t_j = t0 + (j + j_offset) * h  (multiplication, :5307)
    int64_t n;             // local intervals
    int64_t j_offset;      // global index of local interval 0 (mesh slabs across GPUs)
    int trap_time;         // trapezoid second-term time: 0 = t_j, 1 = t_{j+1}
};

template <int NV>
__device__ __forceinline__ void load_node(const double *src, double *dst) {
    if constexpr (NV % 2 == 0) {
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
#pragma unroll
        for (int c = 0; c < NV / 2; ++c) {
            const double2 v = s2[c];
            dst[2 * c] = v.x;
            dst[2 * c + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < NV; ++c) dst[c] = src[c];
    }
}

// block = identity_sign * I - hs * fy, streamed out (written once, read once by the factor kernel)
template <int NV>
__device__ __forceinline__ void store_block(double *dst, const double *fy, double hs, double idsign) {
    if constexpr ((NV * NV) % 2 == 0) {
        double2 *d2 = reinterpret_cast<double2 *>(dst);
#pragma unroll
        for (int e = 0; e < NV * NV / 2; ++e) {
            const int i0 = (2 * e) / NV, c0 = (2 * e) % NV, i1 = (2 * e + 1) / NV, c1 = (2 * e + 1) % NV;
            double2 v;
            v.x = fma(-hs, fy[2 * e], (i0 == c0) ? idsign : 0.0);
            v.y = fma(-hs, fy[2 * e + 1], (i1 == c1) ? idsign : 0.0);
            st_cs2(d2 + e, v);
        }
    } else {
#pragma unroll
        for (int e = 0; e < NV * NV; ++e) st_cs(dst + e, fma(-hs, fy[e], ((e / NV) == (e % NV)) ? idsign : 0.0));
    }
}

template <class P, bool WITH_J, bool TRAPEZOID>
__global__ void __launch_bounds__(128) assemble_kernel(AsmArgs a) {
    constexpr int NV = P::NV;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    double y0[NV], y1[NV], f0[NV];
    load_node<NV>(a.Y + j * NV, y0);
    load_node<NV>(a.Y + (j + 1) * NV, y1);
    double t, t1, h;
    if (a.mesh) {
        t = a.mesh[j];
        t1 = a.mesh[j + 1];
        h = t1 - t;
    } else {
        t = a.t0 + (double)(j + a.j_offset) * a.h;
        t1 = a.t0 + (double)(j + a.j_offset + 1) * a.h;
        h = a.h;
    }
    double *r = a.F + j * NV;
    if constexpr (!TRAPEZOID) {
        if constexpr (WITH_J) {
            double fy[NV * NV];
            P::eval_f_fy(t, y0, a.params, f0, fy);
            store_block<NV>(a.A + j * (int64_t)(NV * NV), fy, h, -1.0);
        } else {
            P::eval_f(t, y0, a.params, f0);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) r[i] = y1[i] - y0[i] - h * f0[i];
    } else {
        double f1[NV];
        const double tt = a.trap_time ? t1 : t;
        if constexpr (WITH_J) {
            {
                double fy[NV * NV];
                P::eval_f_fy(t, y0, a.params, f0, fy);
                store_block<NV>(a.A + j * (int64_t)(NV * NV), fy, 0.5 * h, -1.0);
            }
            {
                double fy[NV * NV];
                P::eval_f_fy(tt, y1, a.params, f1, fy);
                store_block<NV>(a.B + j * (int64_t)(NV * NV), fy, 0.5 * h, 1.0);
            }
        } else {
            P::eval_f(t, y0, a.params, f0);
            P::eval_f(tt, y1, a.params, f1);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) r[i] = y1[i] - y0[i] - 0.5 * h * (f0[i] + f1[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// Wide blocks (nv > 16, config C5): a dense per-thread fy[nv*nv] cannot live in registers, so the emitter hands the
// structural non-zeros of df/dy to a sink that stores  idsign*delta - hs*value  straight into the block.  The blocks
// are initialised ONCE (rst_b200_create) with idsign*I; structural zeros are never written afterwards, so one assembly
// pass writes NNZ + nv doubles per interval instead of nv*nv.
// ------------------------------------------------------------------------------------------------
template <int NV>
struct BlockSink {
    double *dst;
    double hs, idsign;
    __device__ __forceinline__ void put(int idx, double v) const {
        st_cs(dst + idx, fma(-hs, v, (idx / NV) == (idx % NV) ? idsign : 0.0));
    }
};
struct NullSink {
    __device__ __forceinline__ void put(int, double) const {}
};

template <class P, bool WITH_J, bool TRAPEZOID>
__global__ void __launch_bounds__(64) assemble_wide_kernel(AsmArgs a) {
    constexpr int NV = P::NV;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    double y0[NV], f0[NV];
    load_node<NV>(a.Y + j * NV, y0);
    double t, t1, h;
    if (a.mesh) {
        t = a.mesh[j];
        t1 = a.mesh[j + 1];
        h = t1 - t;
    } else {
        t = a.t0 + (double)(j + a.j_offset) * a.h;
        t1 = a.t0 + (double)(j + a.j_offset + 1) * a.h;
        h = a.h;
    }
    double *r = a.F + j * NV;
    const double *yn = a.Y + (j + 1) * NV;
    if constexpr (!TRAPEZOID) {
        if constexpr (WITH_J) {
            BlockSink<NV> sink{a.A + j * (int64_t)(NV * NV), h, -1.0};
            P::eval_f_nz(t, y0, a.params, f0, sink);
        } else {
            P::eval_f(t, y0, a.params, f0);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) r[i] = yn[i] - y0[i] - h * f0[i];
    } else {
        double y1[NV], f1[NV];
        load_node<NV>(yn, y1);
        const double tt = a.trap_time ? t1 : t;
        if constexpr (WITH_J) {
            BlockSink<NV> sa{a.A + j * (int64_t)(NV * NV), 0.5 * h, -1.0};
            P::eval_f_nz(t, y0, a.params, f0, sa);
            BlockSink<NV> sb{a.B + j * (int64_t)(NV * NV), 0.5 * h, 1.0};
            P::eval_f_nz(tt, y1, a.params, f1, sb);
        } else {
            P::eval_f(t, y0, a.params, f0);
            P::eval_f(tt, y1, a.params, f1);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) r[i] = y1[i] - y0[i] - 0.5 * h * (f0[i] + f1[i]);
    }
}

// blocks[j] = idsign * I  (run once at create for the wide path)
__global__ void init_identity_blocks_kernel(double *blocks, int64_t n, int nv, double idsign) {
    const int64_t total = n * nv * nv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx % (nv * nv));
        blocks[idx] = (e / nv) == (e % nv) ? idsign : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Fused F + J assembly with shared-memory staged block stores (nv >= 6).
// In assemble_kernel every thread streams its own nv*nv block (1152 B at nv = 12) with 16-byte stores, i.e. a
// warp store instruction touches 32 different 128-byte lines with half a sector each (measured 18 % of the HBM
// roofline at nv = 12, profiles/r01_a_summary.md).  Here one warp owns 32 consecutive intervals = ONE contiguous
// 32*nv*nv*8-byte region of A: each lane stages its block into a padded (odd pitch -> conflict-free) shared tile,
// then the warp streams the tile out with fully coalesced 512-byte store instructions.
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void stage_block(double *tile_row, const double *fy, double hs, double idsign) {
#pragma unroll
    for (int e = 0; e < NV * NV; ++e) tile_row[e] = fma(-hs, fy[e], ((e / NV) == (e % NV)) ? idsign : 0.0);
}

template <int NV>
__device__ __forceinline__ void flush_tile(const double *tile, double *dst, int nvalid, int lane) {
    constexpr int BL = NV * NV, PITCH = BL | 1;
    static_assert(BL % 2 == 0, "staged assembly needs an even block size");
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    const int total2 = nvalid * (BL / 2);
#pragma unroll 4
    for (int q2 = lane; q2 < total2; q2 += 32) {
        const int jj = q2 / (BL / 2);
        const int e = 2 * (q2 - jj * (BL / 2));
        const double *s = tile + jj * PITCH + e;
        st_cs2(d2 + q2, make_double2(s[0], s[1]));
    }
}

template <class P, bool TRAPEZOID>
__global__ void __launch_bounds__(32) assemble_j_staged_kernel(AsmArgs a) {
    constexpr int NV = P::NV, BL = NV * NV, PITCH = BL | 1;
    __shared__ double tile[32 * PITCH];
    const int lane = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    const int64_t j = j0 + lane;
    const bool valid = j < a.n;
    const int nvalid = (int)((a.n - j0) < 32 ? (a.n - j0) : 32);
    double y0[NV], y1[NV], f0[NV], f1[NV];
    double t = 0.0, t1 = 0.0, h = a.h;
    if (valid) {
        load_node<NV>(a.Y + j * NV, y0);
        load_node<NV>(a.Y + (j + 1) * NV, y1);
        if (a.mesh) {
            t = a.mesh[j];
            t1 = a.mesh[j + 1];
            h = t1 - t;
        } else {
            t = a.t0 + (double)(j + a.j_offset) * a.h;
            t1 = a.t0 + (double)(j + a.j_offset + 1) * a.h;
        }
        double fy[BL];
        P::eval_f_fy(t, y0, a.params, f0, fy);
        stage_block<NV>(tile + lane * PITCH, fy, TRAPEZOID ? 0.5 * h : h, -1.0);
    }
    __syncwarp();
    flush_tile<NV>(tile, a.A + j0 * (int64_t)BL, nvalid, lane);
    if constexpr (TRAPEZOID) {
        __syncwarp();
        if (valid) {
            double fy[BL];
            P::eval_f_fy(a.trap_time ? t1 : t, y1, a.params, f1, fy);
            stage_block<NV>(tile + lane * PITCH, fy, 0.5 * h, 1.0);
        }
        __syncwarp();
        flush_tile<NV>(tile, a.B + j0 * (int64_t)BL, nvalid, lane);
    }
    if (valid) {
        double *r = a.F + j * NV;
        if constexpr (TRAPEZOID) {
#pragma unroll
            for (int i = 0; i < NV; ++i) r[i] = y1[i] - y0[i] - 0.5 * h * (f0[i] + f1[i]);
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) r[i] = y1[i] - y0[i] - h * f0[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Residual-only assembly with shared-memory staged state and residual (nv <= 16).
// assemble_kernel<P, false> reads Y_j / Y_{j+1} and writes r_j with a lane stride of nv doubles: every 16-byte access
// of a warp touches 32 different lines (measured 0.45 of the HBM roofline at nv = 12, profiles/r01_c_summary.md; the
// kernel runs twice per damped Newton iteration).  Here a CTA owns 128 consecutive intervals: their 129 nodes come in
// with fully coalesced loads into an odd-pitch tile (each lane then reads its two nodes conflict-free), the residual
// rows leave through a second tile with fully coalesced stores.
// ------------------------------------------------------------------------------------------------
template <class P, bool TRAPEZOID>
__global__ void __launch_bounds__(128) assemble_f_staged_kernel(AsmArgs a) {
    constexpr int NV = P::NV, PITCH = NV | 1, TJ = 128;
    __shared__ double ysm[(TJ + 1) * PITCH];
    __shared__ double rsm[TJ * PITCH];
    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * TJ;
    const int nj = (int)((a.n - j0) < TJ ? (a.n - j0) : TJ);       // intervals of this CTA
    {
        const double *src = a.Y + j0 * NV;
        const int total = (nj + 1) * NV;
        for (int idx = tid; idx < total; idx += TJ) ysm[(idx / NV) * PITCH + (idx % NV)] = src[idx];
    }
    __syncthreads();
    if (tid < nj) {
        const int64_t j = j0 + tid;
        double y0[NV], y1[NV], f0[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) { y0[c] = ysm[tid * PITCH + c]; y1[c] = ysm[(tid + 1) * PITCH + c]; }
        double t, t1, h;
        if (a.mesh) {
            t = a.mesh[j];
            t1 = a.mesh[j + 1];
            h = t1 - t;
        } else {
            t = a.t0 + (double)(j + a.j_offset) * a.h;
            t1 = a.t0 + (double)(j + a.j_offset + 1) * a.h;
            h = a.h;
        }
        P::eval_f(t, y0, a.params, f0);
        if constexpr (!TRAPEZOID) {
#pragma unroll
            for (int i = 0; i < NV; ++i) rsm[tid * PITCH + i] = y1[i] - y0[i] - h * f0[i];
        } else {
            double f1[NV];
            P::eval_f(a.trap_time ? t1 : t, y1, a.params, f1);
#pragma unroll
            for (int i = 0; i < NV; ++i) rsm[tid * PITCH + i] = y1[i] - y0[i] - 0.5 * h * (f0[i] + f1[i]);
        }
    }
    __syncthreads();
    {
        double *dst = a.F + j0 * NV;
        const int total = nj * NV;
        for (int idx = tid; idx < total; idx += TJ) st_cs(dst + idx, rsm[(idx / NV) * PITCH + (idx % NV)]);
    }
}

}  // namespace rst
