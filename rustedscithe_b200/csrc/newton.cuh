// newton.cuh -- device-resident pieces of the damped Newton driver (north_star subsystem 3): the
// state, residual, Jacobian blocks, factors and steps stay in HBM; only a handful of scalars per
// damping trial cross PCIe.
//
// Reference semantics restated here (src/numerical/BVP_Damp/):
//   y_trial = y - lambda * step                      NR_Damp_solver_damped.rs:1934-1935
//   s       = ||step||_2                             :1924, :1947 (nalgebra norm, BVP_traits.rs:283)
//   conv    = max(sum_i |y_i| rtol_i, atol)          BVP_utils_damped.rs:209-222 (unknowns only)
//   fbound  = bound_step_Cantera2(y, step, bounds)   BVP_utils_damped.rs:39-61 (evaluates y + step)
//   NaN scans of F and of the step                   NR_Damp_solver_damped.rs:1821-1826, :1839-1844
// The full (N+1) x nv node-major state keeps boundary slots pinned; `left_mask` / `right_mask` mark the
// pinned components of node 0 / node N so reductions run over the reference's unknowns only.
#pragma once
#include "rst_common.cuh"
#include "p2p.cuh"

namespace rst {

struct ReduceArgs {
    const double *Y;      // state the tolerance sum / bound guard refer to   [(n+1)*nv]
    const double *dY;     // step                                             [(n+1)*nv]
    const double *F;      // residual (NaN scan only, may be nullptr)          [n*nv]
    const double *lo, *hi, *rtol;  // per variable [nv]
    int64_t n_nodes;      // nodes this rank owns
    int64_t n_rows;       // residual block rows this rank owns (NaN scan of F)
    int nv;
    unsigned long long left_mask, right_mask;
    int has_left, has_right;  // this rank owns global node 0 / global node N
    double *partials;     // [gridDim.x][4]: sumsq(step), sum |y| rtol, min fbound, nan flags
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(RST_FULL_MASK, v, off);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmin(v, __shfl_xor_sync(RST_FULL_MASK, v, off));
    return v;
}

// one pass over the step: all scalars a damping trial needs
template <int NV>
__global__ void __launch_bounds__(256) newton_reduce_kernel(ReduceArgs a) {
    double ss = 0.0, tol = 0.0, fb = 1.0, nanf = 0.0;
    __shared__ double s_lo[NV], s_hi[NV], s_rt[NV];
    if (threadIdx.x < NV) {
        s_lo[threadIdx.x] = a.lo[threadIdx.x];
        s_hi[threadIdx.x] = a.hi[threadIdx.x];
        s_rt[threadIdx.x] = a.rtol[threadIdx.x];
    }
    __syncthreads();
    const int64_t total = a.n_nodes * NV;
    const int64_t right0 = (a.n_nodes - 1) * NV;  // first element of the last owned node
    auto visit = [&](int64_t idx, double y, double s) {
        if (s != s) nanf = 2.0;
        const int v = (int)(idx % NV);  // NV is a compile-time constant: multiply-shift, no division
        const bool pinned = (a.has_left && idx < NV && ((a.left_mask >> v) & 1ull)) ||
                            (a.has_right && idx >= right0 && ((a.right_mask >> v) & 1ull));
        if (pinned) return;
        ss = fma(s, s, ss);
        tol = fma(fabs(y), s_rt[v], tol);
        const double nv_ = y + s;  // the reference evaluates y + step although it applies y - lambda*step
        if (nv_ > s_hi[v]) fb = fmax(fmin(fb, (s_hi[v] - y) / (nv_ - y)), 0.0);
        else if (nv_ < s_lo[v]) fb = fmin(fb, (y - s_lo[v]) / (y - nv_));
    };
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if constexpr (NV % 2 == 0) {  // fully coalesced 16-byte loads over the flat state
        const double2 *y2 = reinterpret_cast<const double2 *>(a.Y);
        const double2 *d2 = reinterpret_cast<const double2 *>(a.dY);
        for (int64_t i2 = tid; i2 < total / 2; i2 += nthr) {
            const double2 yy = y2[i2], sv = d2[i2];
            visit(2 * i2, yy.x, sv.x);
            visit(2 * i2 + 1, yy.y, sv.y);
        }
    } else {
        for (int64_t idx = tid; idx < total; idx += nthr) visit(idx, a.Y[idx], a.dY[idx]);
    }
    if (a.F) {
        const int64_t nf = a.n_rows * NV;
        if constexpr (NV % 2 == 0) {
            const double2 *f2 = reinterpret_cast<const double2 *>(a.F);
            for (int64_t i2 = tid; i2 < nf / 2; i2 += nthr) {
                const double2 f = f2[i2];
                if (f.x != f.x || f.y != f.y) nanf = fmax(nanf, 1.0);
            }
        } else {
            for (int64_t idx = tid; idx < nf; idx += nthr) {
                const double f = a.F[idx];
                if (f != f) nanf = fmax(nanf, 1.0);
            }
        }
    }
    __shared__ double sh[4][8];
    ss = warp_sum(ss);
    tol = warp_sum(tol);
    fb = warp_min(fb);
    nanf = -warp_min(-nanf);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = ss; sh[1][w] = tol; sh[2][w] = fb; sh[3][w] = nanf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 1.0, a3 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            a0 += sh[0][i]; a1 += sh[1][i]; a2 = fmin(a2, sh[2][i]); a3 = fmax(a3, sh[3][i]);
        }
        double *p = a.partials + 4 * (int64_t)blockIdx.x;
        p[0] = a0; p[1] = a1; p[2] = a2; p[3] = a3;
    }
}

// run-time nv variant (wide blocks): same reductions, one element per thread iteration
__global__ void __launch_bounds__(256) newton_reduce_generic_kernel(ReduceArgs a) {
    double ss = 0.0, tol = 0.0, fb = 1.0, nanf = 0.0;
    const int64_t total = a.n_nodes * a.nv;
    const int64_t right0 = (a.n_nodes - 1) * a.nv;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = tid; idx < total; idx += nthr) {
        const double s = a.dY[idx], y = a.Y[idx];
        if (s != s) nanf = 2.0;
        const int v = (int)(idx % a.nv);
        const bool pinned = (a.has_left && idx < a.nv && ((a.left_mask >> v) & 1ull)) ||
                            (a.has_right && idx >= right0 && ((a.right_mask >> v) & 1ull));
        if (pinned) continue;
        ss = fma(s, s, ss);
        tol = fma(fabs(y), a.rtol[v], tol);
        const double nv_ = y + s;
        if (nv_ > a.hi[v]) fb = fmax(fmin(fb, (a.hi[v] - y) / (nv_ - y)), 0.0);
        else if (nv_ < a.lo[v]) fb = fmin(fb, (y - a.lo[v]) / (y - nv_));
    }
    if (a.F) {
        const int64_t nf = a.n_rows * a.nv;
        for (int64_t idx = tid; idx < nf; idx += nthr) {
            const double f = a.F[idx];
            if (f != f) nanf = fmax(nanf, 1.0);
        }
    }
    __shared__ double sh[4][8];
    ss = warp_sum(ss); tol = warp_sum(tol); fb = warp_min(fb); nanf = -warp_min(-nanf);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = ss; sh[1][w] = tol; sh[2][w] = fb; sh[3][w] = nanf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 1.0, a3 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            a0 += sh[0][i]; a1 += sh[1][i]; a2 = fmin(a2, sh[2][i]); a3 = fmax(a3, sh[3][i]);
        }
        double *p = a.partials + 4 * (int64_t)blockIdx.x;
        p[0] = a0; p[1] = a1; p[2] = a2; p[3] = a3;
    }
}

// deterministic second stage: out[0..3] = {sumsq, tolsum, fbound, nanflag}; out[4] = device error flags.
// Multi-GPU with mapped peers: the 8-scalar all-gather over NVLink and the combination run in the same launch.
__global__ void __launch_bounds__(256) newton_reduce_final_kernel(const double *partials, int nblocks, double *out,
                                                                  unsigned *flags, P2pPeers peers, int rank, int world,
                                                                  int pay, unsigned long long seq, double *gscal) {
    double a0 = 0.0, a1 = 0.0, a2 = 1.0, a3 = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        a0 += partials[4 * i]; a1 += partials[4 * i + 1];
        a2 = fmin(a2, partials[4 * i + 2]); a3 = fmax(a3, partials[4 * i + 3]);
    }
    __shared__ double sh[4][8];
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_min(a2); a3 = -warp_min(-a3);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = a0; sh[1][w] = a1; sh[2][w] = a2; sh[3][w] = a3; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double b0 = 0.0, b1 = 0.0, b2 = 1.0, b3 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            b0 += sh[0][i]; b1 += sh[1][i]; b2 = fmin(b2, sh[2][i]); b3 = fmax(b3, sh[3][i]);
        }
        out[0] = b0; out[1] = b1; out[2] = b2; out[3] = b3;
        out[4] = (double)(flags[0] | flags[2]);   // flags[2]: sticky (peer timeout poisons the handle)
        out[5] = 0.0; out[6] = 0.0; out[7] = 0.0;
    }
    if (world > 1) {
        p2p_allgather_block(peers, out, 8, gscal, rank, world, pay, seq, flags);
        if (threadIdx.x == 0) {
            double c0 = 0.0, c1 = 0.0, c2 = 1.0, c3 = 0.0;
            unsigned c4 = flags[0] | flags[2];
            for (int p = 0; p < world; ++p) {
                c0 += gscal[8 * p]; c1 += gscal[8 * p + 1]; c2 = fmin(c2, gscal[8 * p + 2]);
                c3 = fmax(c3, gscal[8 * p + 3]);
                const double f = gscal[8 * p + 4];
                c4 |= (f == f && f >= 0.0 && f < 4.0e9) ? (unsigned)f : RST_FLAG_PEER_TIMEOUT;   // a NaN word = poisoned exchange
            }
            out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
            out[4] = (double)c4;
        }
    }
}

// last node of the damped-step sequence: both scalar sets go straight into mapped pinned host memory, followed by a ticket
// the host spins on (a stream synchronise costs ~10 us of turnaround per Newton iteration; the ticket counter lives on the
// device so the launch is graph-replayable)
__global__ void __launch_bounds__(32) publish_scalars_kernel(const double *__restrict__ d_scal, double *h_scal,
                                                             unsigned long long *d_ticket, volatile unsigned long long *h_ticket) {
    if (threadIdx.x < 16) h_scal[threadIdx.x] = d_scal[threadIdx.x];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) {
        const unsigned long long t = *d_ticket + 1ull;
        *d_ticket = t;
        *h_ticket = t;
        __threadfence_system();
    }
}

// y_trial = y - lambda * step   (NR_Damp_solver_damped.rs:1934-1935: mul_float, then subtraction)
__global__ void __launch_bounds__(256) newton_trial_kernel(const double *__restrict__ Y, const double *__restrict__ dY,
                                                           double *__restrict__ Yt, double lambda, int64_t total) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x)
        Yt[idx] = Y[idx] - dY[idx] * lambda;
}

// same with lambda = 1.0 * fbound read from the device-resident reduction result (scal[2]): lets the first damping
// trial be enqueued without a host round trip
__global__ void __launch_bounds__(256) newton_trial_dev_kernel(const double *__restrict__ Y, const double *__restrict__ dY,
                                                               double *__restrict__ Yt, const double *__restrict__ scal,
                                                               int64_t total) {
    const double lambda = 1.0 * scal[2];
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x)
        Yt[idx] = Y[idx] - dY[idx] * lambda;
}

// First trial of an iteration whose undamped step is ALREADY KNOWN: when the previous iteration accepted a trial (state Yt, step
// dY1 = J^-1 F(Yt)) and the Jacobian is kept, the reference's next iteration recomputes F(y) and J^-1 F(y) at y = Yt with the same
// factors (NR_Damp_solver_damped.rs:1862-1880) -- bit for bit the trial's dY1.  The host takes it over by swapping the roles of the
// two step buffers (so `step` below IS the accepted trial's step); this kernel forms the next trial state from it and moves the
// accepted trial's scalars into slot 0 when they are in slot 1 (speculative first trial; a later trial reduces into slot 0).
__global__ void __launch_bounds__(256) newton_trial_carry_kernel(const double *__restrict__ Y, const double *__restrict__ step,
                                                                 double *__restrict__ Yt, double *__restrict__ scal, int64_t total,
                                                                 int src_slot) {
    const double lambda = 1.0 * scal[8 * src_slot + 2];
    if (src_slot == 1 && blockIdx.x == 0 && threadIdx.x < 8) scal[threadIdx.x] = scal[8 + threadIdx.x];   // nobody reads slot 0 in this kernel
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x)
        Yt[idx] = Y[idx] - step[idx] * lambda;
}

// ---- reduced ordering <-> full node-major state (numeric_discretization.rs:80-106, BVP_utils.rs:534-558) ----
// reduced index k -> full position: the nl pinned slots of node 0 and the nr pinned slots of node N are deleted
struct LayoutArgs {
    int64_t n_steps;
    int nv;
    int nl, nr;
    int left_free[64];   // free components of node 0, ascending
    int right_free[64];  // free components of node N, ascending
};

__device__ __forceinline__ int64_t reduced_to_full(const LayoutArgs &a, int64_t k) {
    const int64_t first = a.nv - a.nl;
    if (k < first) return a.left_free[k];
    const int64_t interior_end = first + (a.n_steps - 1) * (int64_t)a.nv;
    if (k < interior_end) return k + a.nl;
    return a.n_steps * (int64_t)a.nv + a.right_free[k - interior_end];
}

__global__ void __launch_bounds__(256) scatter_reduced_kernel(LayoutArgs a, const double *__restrict__ reduced,
                                                              double *__restrict__ full) {
    const int64_t n = a.n_steps * (int64_t)a.nv;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        full[reduced_to_full(a, k)] = reduced[k];
}
__global__ void __launch_bounds__(256) gather_reduced_kernel(LayoutArgs a, const double *__restrict__ full,
                                                             double *__restrict__ reduced) {
    const int64_t n = a.n_steps * (int64_t)a.nv;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        reduced[k] = full[reduced_to_full(a, k)];
}

// Jacobian value stream in the reference's entry order (sorted (row, col), symbolic_functions_BVP.rs:1343):
// src[k] >= 0: element offset into A; -1: the +1 identity entry of B (forward); <= -2: offset -(src+2) into B
__global__ void __launch_bounds__(256) gather_values_kernel(const int64_t *__restrict__ src, const double *__restrict__ A,
                                                            const double *__restrict__ B, double *__restrict__ out,
                                                            int64_t nnz) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = src[k];
        out[k] = s >= 0 ? A[s] : (s == -1 ? 1.0 : B[-(s + 2)]);
    }
}

// residual check r = rhs - J x in block form (used by the guarded refinement, lapack_style_banded.rs:1227-1298)
template <int NV>
__global__ void __launch_bounds__(128) block_residual_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                             const double *__restrict__ X, const double *__restrict__ rhs,
                                                             double *__restrict__ out, int64_t n) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * NV) return;
    const int64_t j = idx / NV;
    const int i = (int)(idx - j * NV);
    const double *a = A + (j * NV + i) * NV;
    double s = rhs[idx];
#pragma unroll
    for (int c = 0; c < NV; ++c) s = fma(-a[c], X[j * NV + c], s);
    if (B) {
        const double *b = B + (j * NV + i) * NV;
#pragma unroll
        for (int c = 0; c < NV; ++c) s = fma(-b[c], X[(j + 1) * NV + c], s);
    } else {
        s -= X[(j + 1) * NV + i];
    }
    out[idx] = s;
}

// run-time block size (wide blocks): one thread per residual row
__global__ void __launch_bounds__(128) block_residual_generic_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                                     const double *__restrict__ X, const double *__restrict__ rhs,
                                                                     double *__restrict__ out, int64_t n, int nv) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * nv) return;
    const int64_t j = idx / nv;
    const int i = (int)(idx - j * nv);
    const double *a = A + (j * nv + i) * (int64_t)nv;
    double s = rhs[idx];
    for (int c = 0; c < nv; ++c) s = fma(-a[c], X[j * nv + c], s);
    if (B) {
        const double *b = B + (j * nv + i) * (int64_t)nv;
        for (int c = 0; c < nv; ++c) s = fma(-b[c], X[(j + 1) * nv + c], s);
    } else {
        s -= X[(j + 1) * nv + i];
    }
    out[idx] = s;
}

// max |v| as the bit pattern of a non-negative double (orders like an unsigned integer); NaN / inf map to all ones
__global__ void __launch_bounds__(256) absmax_kernel(const double *__restrict__ v, int64_t total, unsigned long long *out) {
    unsigned long long m = 0ull;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const double x = fabs(v[idx]);
        const unsigned long long b = (x == x && x < 1.0e308 * 10.0) ? (unsigned long long)__double_as_longlong(x) : ~0ull;
        m = b > m ? b : m;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(RST_FULL_MASK, m, off);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__global__ void __launch_bounds__(256) axpy_kernel(double *__restrict__ x, const double *__restrict__ d, double alpha,
                                                   int64_t total) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x)
        x[idx] = fma(alpha, d[idx], x[idx]);
}

}  // namespace rst
