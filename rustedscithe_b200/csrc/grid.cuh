// grid.cuh -- adaptive grid refinement hand-off on the device.
//
// Restates `new_grid` (src/numerical/BVP_Damp/grid_api.rs:154-189) and the algorithms it dispatches to
// (src/numerical/BVP_Damp/adaptive_grid_basic.rs): DoublePoints (:187-203), Easy (:257-313, parallel variant),
// Pearson (:469-540, parallel variant incl. the shifted-mark merge :50-54 and the buffer pass :517-533),
// GrcarSmooke (:565-741, the sequential variant `new_grid` calls), Sci (:885-924 with modify_mesh :926-945 and
// interpolate_solution :947-989), and the common constructor `construct_refined_grid_from_marks` (:56-105).
//
// The reference works on an nv x (N+1) matrix of the full solution; here the state already is the full node-major
// (N+1) x nv array in HBM, so refinement is: per-variable extrema (one pass) -> marks (one pass per variable where the
// reference's merge is order dependent) -> exclusive scan of 1 + mark -> scatter of old nodes + linear interpolation.
// All arithmetic that reaches the new mesh / guess uses explicitly rounded operations (no FMA contraction) so the result
// is bit-identical to the reference's `y + dy * theta`.
#pragma once
#include "rst_common.cuh"

namespace rst {

struct GridArgs {
    const double *Y;      // [(N+1)][nv] node-major full solution
    const double *mesh;   // [N+1] or nullptr (uniform: x_i = t0 + i * h, solver_common.rs:43-46)
    double t0, h;
    int64_t n_nodes;      // N + 1
    int nv;
};

__device__ __forceinline__ double grid_x(const GridArgs &g, int64_t i) {
    return g.mesh ? g.mesh[i] : __dadd_rn(g.t0, __dmul_rn((double)i, g.h));
}

// order-preserving map double -> uint64 (so integer atomics give min / max)
__device__ __forceinline__ unsigned long long ord_encode(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double ord_decode(unsigned long long e) {
    const unsigned long long u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof(d));
    return d;
#endif
}

__global__ void grid_extrema_init_kernel(unsigned long long *ext, int nv) {
    for (int i = threadIdx.x; i < 4 * nv; i += blockDim.x) ext[i] = (i & 1) ? 0ull : ~0ull;   // min slots ~0, max slots 0
}

// ext[v][0..3] = encoded (min y, max y, min dy/dx, max dy/dx) of variable v (grid_extrema_init_kernel first)
__global__ void __launch_bounds__(256) grid_extrema_kernel(GridArgs g, unsigned long long *ext, int64_t step) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = g.n_nodes * g.nv;
    const int v = (int)(tid % g.nv);   // step is a multiple of nv: a thread stays on one variable
    double ymin = 0, ymax = 0, smin = 0, smax = 0;
    bool have = false, have_s = false;
    for (int64_t idx = tid; tid < step && idx < total; idx += step) {
        const int64_t i = idx / g.nv;
        const double y = g.Y[idx];
        if (!have) { ymin = ymax = y; have = true; }
        else { ymin = fmin(ymin, y); ymax = fmax(ymax, y); }
        if (i + 1 < g.n_nodes) {
            const double hi = __dsub_rn(grid_x(g, i + 1), grid_x(g, i));
            const double s = __ddiv_rn(__dsub_rn(g.Y[idx + g.nv], y), hi);
            if (!have_s) { smin = smax = s; have_s = true; }
            else { smin = fmin(smin, s); smax = fmax(smax, s); }
        }
    }
    // block-level combine in shared memory first: a few hundred thousand threads on 4 nv global addresses would serialise
    __shared__ unsigned long long sm[4 * 64];
    for (int i = threadIdx.x; i < 4 * g.nv; i += blockDim.x) sm[i] = (i & 1) ? 0ull : ~0ull;
    __syncthreads();
    if (have) {
        atomicMin(sm + 4 * v + 0, ord_encode(ymin));
        atomicMax(sm + 4 * v + 1, ord_encode(ymax));
    }
    if (have_s) {
        atomicMin(sm + 4 * v + 2, ord_encode(smin));
        atomicMax(sm + 4 * v + 3, ord_encode(smax));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * g.nv; i += blockDim.x) {
        if (i & 1) atomicMax(ext + i, sm[i]);
        else atomicMin(ext + i, sm[i]);
    }
}

// Rust `as i32` on f64: truncate, saturate, NaN -> 0
__device__ __forceinline__ int rust_f64_as_i32(double x) {
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (-2147483647 - 1);
    return (int)x;
}

// DoublePoints: every interval gets one point (:187-191)
__global__ void grid_mark_all_kernel(int *mark, int64_t n_nodes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) mark[i] = i + 1 < n_nodes ? 1 : 0;
}

// Easy (:257-305): mark[i-1] = 1 when |y[i] - y[i-1]| > tol (max - min) for any variable, i = 1 .. len-2
__global__ void grid_mark_easy_kernel(GridArgs g, const unsigned long long *ext, double tol, int *mark) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // mark index = i - 1
    if (m >= g.n_nodes) return;
    int out = 0;
    const int64_t i = m + 1;
    if (i >= 1 && i + 1 < g.n_nodes) {
        for (int v = 0; v < g.nv; ++v) {
            const double delta = __dmul_rn(tol, __dsub_rn(ord_decode(ext[4 * v + 1]), ord_decode(ext[4 * v + 0])));
            const double tau = fabs(__dsub_rn(g.Y[i * g.nv + v], g.Y[(i - 1) * g.nv + v]));
            if (tau > delta) out = 1;
        }
    }
    mark[m] = out;
}

// Pearson / Grcar-Smooke marking for ONE variable v (the merge across variables is order dependent, :50-54 and :668-672):
//   trigger at i (1 <= i <= len-2) with count N writes mark[i-1] = N when N > mark[i] (the value before this variable's pass)
// gather form: out[m] = (trigger(m+1) && N > in[m+1]) ? N : in[m]
__global__ void grid_mark_jump_kernel(GridArgs g, const unsigned long long *ext, int v, double d, double gpar, int use_eta,
                                      const int *in, int *out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.n_nodes) return;
    int val = in[m];
    const int64_t i = m + 1;
    if (i >= 1 && i + 1 < g.n_nodes) {
        const double ym = g.Y[(i - 1) * g.nv + v], y0 = g.Y[i * g.nv + v], yp = g.Y[(i + 1) * g.nv + v];
        const double delta = __dmul_rn(d, __dsub_rn(ord_decode(ext[4 * v + 1]), ord_decode(ext[4 * v + 0])));
        const double tau = fabs(__dsub_rn(y0, ym));
        const bool not_small = (fabs(y0) > 1e-4) & (fabs(yp) > 1e-4);
        bool trig = tau > delta && not_small;
        if (use_eta) {
            const double gamma = __dmul_rn(gpar, __dsub_rn(ord_decode(ext[4 * v + 3]), ord_decode(ext[4 * v + 2])));
            const double xm = grid_x(g, i - 1), x0 = grid_x(g, i), xp = grid_x(g, i + 1);
            const double s1 = __ddiv_rn(__dsub_rn(yp, y0), __dsub_rn(xp, x0));
            const double s0 = __ddiv_rn(__dsub_rn(y0, ym), __dsub_rn(x0, xm));
            const double eta = fabs(__dsub_rn(s1, s0));
            trig = trig || (eta > gamma && not_small);
        }
        if (trig) {
            int N = rust_f64_as_i32(__ddiv_rn(tau, delta));
            if (N < 1) N = 1;
            if (N > in[i]) val = N;
        }
    }
    out[m] = val;
}

// buffer pass (:517-533 / :703-733): bounded ratio of adjacent intervals.  Step i writes only mark[i-1] and reads the
// untouched mark[i], so the sequential loop is a gather.
__global__ void grid_mark_buffer_kernel(GridArgs g, double C, const int *in, int *out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= g.n_nodes) return;
    int val = in[m];
    const int64_t i = m + 1;
    if (i >= 1 && i + 1 < g.n_nodes) {
        const double h1 = __dsub_rn(grid_x(g, i + 1), grid_x(g, i)), h0 = __dsub_rn(grid_x(g, i), grid_x(g, i - 1));
        const double ratio = __ddiv_rn(h1, h0);
        const bool c1 = ratio <= C, c2 = ratio >= __ddiv_rn(1.0, C);
        if (!c1 && (i - 1) != 0) {
            if (1 > val) val = 1;
        }
        if (!c2) {
            if (1 > in[i]) val = 1;
        }
    }
    out[m] = val;
}

// Sci (:896-903): residual thresholds; index j of the residual vector is used as the interval index (reference quirk)
__global__ void grid_mark_sci_kernel(const double *res, int64_t m_res, double tol, int *mark, int64_t n_nodes, unsigned *oob) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int val = 0;
    if (j + 1 < m_res) {
        const double r = res[j];
        if (r > tol && r < __dmul_rn(100.0, tol)) val = 1;
        else if (r >= __dmul_rn(100.0, tol)) val = 2;
    }
    if (j < n_nodes) {
        if (j + 1 >= n_nodes && val) atomicOr(oob, 1u);   // x[i+1] out of bounds: the reference panics
        else mark[j] = val;
    } else if (val && j < m_res) {
        atomicOr(oob, 1u);
    }
}

// ---- exclusive scan of cnt[i] = 1 + (i < len-1 ? mark[i] : 0) ----
constexpr int GRID_SCAN_BLOCK = 1024;
__global__ void __launch_bounds__(GRID_SCAN_BLOCK) grid_scan_local_kernel(const int *mark, int64_t n_nodes, int64_t *pos,
                                                                         int64_t *block_sum) {
    __shared__ int64_t sh[GRID_SCAN_BLOCK];
    const int64_t i = (int64_t)blockIdx.x * GRID_SCAN_BLOCK + threadIdx.x;
    int64_t c = 0;
    if (i < n_nodes) c = 1 + (i + 1 < n_nodes ? (int64_t)mark[i] : 0);
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int off = 1; off < GRID_SCAN_BLOCK; off <<= 1) {
        int64_t t = 0;
        if ((int)threadIdx.x >= off) t = sh[threadIdx.x - off];
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n_nodes) pos[i] = sh[threadIdx.x] - c;
    if (threadIdx.x == GRID_SCAN_BLOCK - 1) block_sum[blockIdx.x] = sh[threadIdx.x];
}
__global__ void grid_scan_blocks_kernel(int64_t *block_sum, int64_t nblocks, int64_t *total) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t acc = 0;
        for (int64_t b = 0; b < nblocks; ++b) {
            const int64_t t = block_sum[b];
            block_sum[b] = acc;
            acc += t;
        }
        *total = acc;
    }
}
__global__ void __launch_bounds__(GRID_SCAN_BLOCK) grid_scan_add_kernel(int64_t *pos, int64_t n_nodes, const int64_t *block_sum) {
    const int64_t i = (int64_t)blockIdx.x * GRID_SCAN_BLOCK + threadIdx.x;
    if (i < n_nodes) pos[i] += block_sum[blockIdx.x];
}

// construct_refined_grid_from_marks (:56-105) / modify_mesh + interpolate_solution for Sci (:926-989)
__global__ void __launch_bounds__(256) grid_construct_kernel(GridArgs g, const int *mark, const int64_t *pos, int sci,
                                                            double *mesh_new, double *Y_new) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.n_nodes * g.nv) return;
    const int64_t i = idx / g.nv;
    const int v = (int)(idx - i * g.nv);
    const int64_t p = pos[i];
    const double y = g.Y[idx];
    Y_new[p * g.nv + v] = y;
    const double xi = grid_x(g, i);
    if (v == 0) mesh_new[p] = xi;
    const int n = (i + 1 < g.n_nodes) ? mark[i] : 0;
    if (n > 0) {
        const double xn = grid_x(g, i + 1);
        const double hi = __dsub_rn(xn, xi);
        const double dy = __dsub_rn(g.Y[idx + g.nv], y);
        for (int k = 1; k <= n; ++k) {
            const double theta = __ddiv_rn((double)k, __dadd_rn((double)n, 1.0));
            Y_new[(p + k) * g.nv + v] = __dadd_rn(y, __dmul_rn(dy, theta));
            if (v == 0) {
                double xnew;
                if (!sci) xnew = __dadd_rn(xi, __dmul_rn(hi, theta));
                else if (n == 1) xnew = __dmul_rn(0.5, __dadd_rn(xi, xn));
                else if (k == 1) xnew = __ddiv_rn(__dadd_rn(__dmul_rn(2.0, xi), xn), 3.0);
                else xnew = __ddiv_rn(__dadd_rn(xi, __dmul_rn(2.0, xn)), 3.0);
                mesh_new[p + k] = xnew;
            }
        }
    }
}

}  // namespace rst
