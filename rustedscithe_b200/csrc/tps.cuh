// tps.cuh -- "thread per slab" kernels of the pivoted condensation solver for tiny blocks (nv = 2: configs C1/C2).
//
// Same algorithm as abd.cuh (pivoted elimination of the node shared by the running condensed row and the next block
// row, slab by slab, recursively), but for nv = 2 a 4 x 6 panel fits in one thread's registers: no shuffles, no shared
// memory, slab length S is a compile-time constant so the whole slab is unrolled and its loads are issued up front.
// The round-1 launch list (profiles/r01_a_summary.md) showed the lanes-per-slab kernels spend as long in the upper
// levels (latency chains of S dependent steps behind every launch) as in level 0, so here
//   * levels with more than TPS_TAIL_ROWS rows run one launch each,
//   * all remaining levels, the boundary closure and their back-substitution run inside ONE single-CTA "tail" kernel
//     (levels separated by __syncthreads, data L2 resident).
// Pivoting: at elimination step k the rows k..2nv-1 are scanned with conditional swaps (static register indices); the
// swap decisions are recorded as a bit mask and replayed on the right-hand side by the forward sweep.
//
// Factor storage (thread-interleaved, every access coalesced):   fac[((i-1)*NE + e) * nslabs + slab]
//   e: [ L (k, row>k) | per k: 1/pivot, U(k, c>k) | E (k,c) | F (k,c) ],  bits[(i-1)*nslabs + slab], e_rhs likewise.
#pragma once
#include "rst_common.cuh"
#include "p2p.cuh"

namespace rst {

constexpr int TPS_S = 8;            // rows per slab (compile-time: the slab loop is fully unrolled)
constexpr int TPS_TAIL_ROWS = 4096;
constexpr int TPS_MAX_TAIL_LEVELS = 8;

struct TpsLevel {
    int64_t n, nslabs;
    const double *A, *B;      // [n][nv][nv]; B == nullptr -> identity
    double *A_next, *B_next;  // [nslabs][nv][nv]
    double *fac;
    unsigned *bits;
    const double *r;          // [n][nv]
    double *e;                // [(S-1)][nv][nslabs]
    double *r_next;           // [nslabs][nv]
    double *Z;                // [n+1][nv]
    const double *Z_next;     // [nslabs+1][nv]
};

struct TpsTail {
    int nlevels;
    TpsLevel lv[TPS_MAX_TAIL_LEVELS];
    // boundary closure (world == 1): C = last A_next, D = last B_next, g = last r_next
    unsigned long long left_mask, right_mask;
    double *Ztop;             // [2][nv]
    int do_top;
};

// damping reductions fused into the level-0 backward sweep (the sweep already holds every component of the step):
// ||step||^2, sum |y| rtol and the Cantera bound factor (newton.cuh) without re-reading the step or the residual
struct TpsReduce {
    const double *Y;              // state the tolerance sum / bound guard refer to
    const double *lo, *hi, *rtol; // per variable
    unsigned long long left_mask, right_mask;
    int has_left, has_right;
    int64_t n_own;                // local nodes [0, n_own) belong to this rank
    double *partials;             // [gridDim.x][4]
};

struct TpsRedAcc {
    double ss, tol, fb, nanf;
};

template <int NV>
__device__ __forceinline__ void tps_red_visit(const TpsReduce &R, TpsRedAcc &a, int64_t node, const double *step,
                                              const double *y, const double *lo, const double *hi, const double *rt) {
    if (node >= R.n_own) return;
    unsigned long long pin = 0ull;
    if (R.has_left && node == 0) pin |= R.left_mask;
    if (R.has_right && node == R.n_own - 1) pin |= R.right_mask;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        const double s = step[c], yy = y[c];
        if (s != s) a.nanf = 2.0;
        if ((pin >> c) & 1ull) continue;
        a.ss = fma(s, s, a.ss);
        a.tol = fma(fabs(yy), rt[c], a.tol);
        const double nv_ = yy + s;
        if (nv_ > hi[c]) a.fb = fmax(fmin(a.fb, (hi[c] - yy) / (nv_ - yy)), 0.0);
        else if (nv_ < lo[c]) a.fb = fmin(a.fb, (yy - lo[c]) / (yy - nv_));
    }
}

template <int NV>
struct TpsCfg {
    static constexpr int R2 = 2 * NV;
    static constexpr int NL = NV * (3 * NV - 1) / 2;  // sum_k (2nv-1-k)
    static constexpr int NU = NV * (NV + 1) / 2;
    static constexpr int NE = NL + NU + 2 * NV * NV;
};

// FULL = the slab has all S rows: every load is unconditional, so the unrolled body issues the whole slab's loads up
// front instead of one dependent load per step (only the last slab of a level can be ragged).
template <int NV, int S, bool IDENT_B, bool FULL>
__device__ __forceinline__ void tps_factor_slab_impl(const TpsLevel &L, int64_t slab, unsigned *flags) {
    using Cfg = TpsCfg<NV>;
    constexpr int R2 = Cfg::R2, W = 3 * NV;
    const int64_t g0 = slab * S;
    const int len = FULL ? S : (int)((L.n - g0) < (int64_t)S ? (L.n - g0) : (int64_t)S);
    double R[R2][W];
#pragma unroll
    for (int i = 0; i < R2; ++i)
#pragma unroll
        for (int c = 0; c < W; ++c) R[i][c] = 0.0;
    // carry rows <- first block row of the slab: P = A, M = B
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            R[i][NV + c] = L.A[(g0 * NV + i) * NV + c];
            R[i][c] = IDENT_B ? ((i == c) ? 1.0 : 0.0) : L.B[(g0 * NV + i) * NV + c];
        }
    // all block rows of the slab are loaded before the dependent elimination chain starts
    double An[S - 1][NV][NV], Bn[IDENT_B ? 1 : S - 1][NV][NV];
#pragma unroll
    for (int st = 1; st < S; ++st) {
        const bool ok = FULL || st < len;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                An[st - 1][i][c] = ok ? L.A[((g0 + st) * NV + i) * NV + c] : 0.0;
                if (!IDENT_B) Bn[st - 1][i][c] = ok ? L.B[((g0 + st) * NV + i) * NV + c] : 0.0;
            }
    }
#pragma unroll
    for (int st = 1; st < S; ++st) {
        if (FULL || st < len) {
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c) {
                    R[NV + i][c] = An[st - 1][i][c];
                    R[NV + i][NV + c] = 0.0;
                    R[NV + i][2 * NV + c] = IDENT_B ? ((i == c) ? 1.0 : 0.0) : Bn[IDENT_B ? 0 : st - 1][i][c];
                }
            double *fac = L.fac + ((int64_t)(st - 1) * Cfg::NE) * L.nslabs + slab;
            unsigned bits = 0;
            int bit = 0, le = 0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
#pragma unroll
                for (int ii = k + 1; ii < R2; ++ii) {
                    const bool sw = fabs(R[ii][k]) > fabs(R[k][k]);
                    bits |= (sw ? 1u : 0u) << bit;
                    ++bit;
#pragma unroll
                    for (int c = k; c < W; ++c) {
                        const double a = R[k][c], b = R[ii][c];
                        R[k][c] = sw ? b : a;
                        R[ii][c] = sw ? a : b;
                    }
                }
                const double piv = R[k][k];
                if (!(fabs(piv) > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
                const double rc = 1.0 / piv;
                R[k][k] = rc;  // the diagonal slot keeps the reciprocal pivot
#pragma unroll
                for (int ii = k + 1; ii < R2; ++ii) {
                    const double l = R[ii][k] * rc;
                    fac[(int64_t)le * L.nslabs] = l;
                    ++le;
#pragma unroll
                    for (int c = k + 1; c < W; ++c) R[ii][c] = fma(-l, R[k][c], R[ii][c]);
                }
            }
            L.bits[(int64_t)(st - 1) * L.nslabs + slab] = bits;
            int ue = Cfg::NL;
#pragma unroll
            for (int k = 0; k < NV; ++k)
#pragma unroll
                for (int c = k; c < NV; ++c) { fac[(int64_t)ue * L.nslabs] = R[k][c]; ++ue; }
#pragma unroll
            for (int k = 0; k < NV; ++k)
#pragma unroll
                for (int c = 0; c < NV; ++c) { fac[(int64_t)ue * L.nslabs] = R[k][NV + c]; ++ue; }
#pragma unroll
            for (int k = 0; k < NV; ++k)
#pragma unroll
                for (int c = 0; c < NV; ++c) { fac[(int64_t)ue * L.nslabs] = R[k][2 * NV + c]; ++ue; }
            // survivors become the carry rows: M <- Q', P <- P', Q <- 0
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int c = 0; c < NV; ++c) {
                    R[i][NV + c] = R[NV + i][NV + c];
                    R[i][c] = R[NV + i][2 * NV + c];
                    R[i][2 * NV + c] = 0.0;
                }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            L.A_next[(slab * NV + i) * NV + c] = R[i][NV + c];
            L.B_next[(slab * NV + i) * NV + c] = R[i][c];
        }
}

template <int NV, int S, bool IDENT_B>
__device__ __forceinline__ void tps_factor_slab(const TpsLevel &L, int64_t slab, unsigned *flags) {
    if ((slab + 1) * S <= L.n) tps_factor_slab_impl<NV, S, IDENT_B, true>(L, slab, flags);
    else tps_factor_slab_impl<NV, S, IDENT_B, false>(L, slab, flags);
}

template <int NV, int S, bool FULL>
__device__ __forceinline__ void tps_forward_slab_impl(const TpsLevel &L, int64_t slab) {
    using Cfg = TpsCfg<NV>;
    constexpr int R2 = Cfg::R2;
    const int64_t g0 = slab * S;
    const int len = FULL ? S : (int)((L.n - g0) < (int64_t)S ? (L.n - g0) : (int64_t)S);
    double w[R2];
    // every load of the slab is issued before the dependent elimination chain starts
    double lf[S - 1][Cfg::NL], rr[S - 1][NV];
    unsigned bt[S - 1];
#pragma unroll
    for (int st = 1; st < S; ++st) {
        const bool ok = FULL || st < len;
        const double *fac = L.fac + ((int64_t)(st - 1) * Cfg::NE) * L.nslabs + slab;
#pragma unroll
        for (int e = 0; e < Cfg::NL; ++e) lf[st - 1][e] = ok ? __ldcs(fac + (int64_t)e * L.nslabs) : 0.0;
        bt[st - 1] = ok ? __ldcs(L.bits + (int64_t)(st - 1) * L.nslabs + slab) : 0u;
#pragma unroll
        for (int i = 0; i < NV; ++i) rr[st - 1][i] = ok ? L.r[(g0 + st) * NV + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) { w[i] = L.r[g0 * NV + i]; w[NV + i] = 0.0; }
#pragma unroll
    for (int st = 1; st < S; ++st) {
        if (FULL || st < len) {
#pragma unroll
            for (int i = 0; i < NV; ++i) w[NV + i] = rr[st - 1][i];
            const double *fac = lf[st - 1];
            const unsigned bits = bt[st - 1];
            int bit = 0, le = 0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
#pragma unroll
                for (int ii = k + 1; ii < R2; ++ii) {
                    const bool sw = (bits >> bit) & 1u;
                    ++bit;
                    const double a = w[k], b = w[ii];
                    w[k] = sw ? b : a;
                    w[ii] = sw ? a : b;
                }
#pragma unroll
                for (int ii = k + 1; ii < R2; ++ii) {
                    w[ii] = fma(-fac[le], w[k], w[ii]);
                    ++le;
                }
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                L.e[((int64_t)(st - 1) * NV + k) * L.nslabs + slab] = w[k];
                w[k] = w[NV + k];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) L.r_next[slab * NV + i] = w[i];
}

template <int NV, int S>
__device__ __forceinline__ void tps_forward_slab(const TpsLevel &L, int64_t slab) {
    if ((slab + 1) * S <= L.n) tps_forward_slab_impl<NV, S, true>(L, slab);
    else tps_forward_slab_impl<NV, S, false>(L, slab);
}

template <int NV, int S, bool FULL, bool REDUCE = false>
__device__ __forceinline__ void tps_backward_slab_impl(const TpsLevel &L, int64_t slab, const TpsReduce *R = nullptr,
                                                       TpsRedAcc *acc = nullptr) {
    using Cfg = TpsCfg<NV>;
    const int64_t g0 = slab * S;
    const int len = FULL ? S : (int)((L.n - g0) < (int64_t)S ? (L.n - g0) : (int64_t)S);
    double Ys[NV], Yn[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        Ys[c] = L.Z_next[slab * NV + c];
        Yn[c] = L.Z_next[(slab + 1) * NV + c];
        L.Z[g0 * NV + c] = Ys[c];
    }
    if (slab == L.nslabs - 1) {
#pragma unroll
        for (int c = 0; c < NV; ++c) L.Z[L.n * NV + c] = Yn[c];
    }
    double rlo[NV], rhi[NV], rrt[NV], ystate[S][NV];
    if constexpr (REDUCE) {
#pragma unroll
        for (int c = 0; c < NV; ++c) { rlo[c] = R->lo[c]; rhi[c] = R->hi[c]; rrt[c] = R->rtol[c]; }
#pragma unroll
        for (int st = 0; st < S; ++st)
#pragma unroll
            for (int c = 0; c < NV; ++c) ystate[st][c] = (FULL || st < len) ? R->Y[(g0 + st) * NV + c] : 0.0;
        tps_red_visit<NV>(*R, *acc, g0, Ys, ystate[0], rlo, rhi, rrt);
        if (slab == L.nslabs - 1) {
            double yl[NV];
#pragma unroll
            for (int c = 0; c < NV; ++c) yl[c] = R->Y[L.n * NV + c];
            tps_red_visit<NV>(*R, *acc, L.n, Yn, yl, rlo, rhi, rrt);
        }
    }
    // all loads of the slab first; the part of the right-hand side that only needs Y_s is folded in as data arrives,
    // so the dependent chain per step is just F * Y_next and the nv x nv triangular solve
    double Us[S - 1][NV][NV], Fs[S - 1][NV][NV], ts[S - 1][NV];
#pragma unroll
    for (int st = S - 1; st >= 1; --st) {
        const bool ok = FULL || st < len;
        const double *fac = L.fac + ((int64_t)(st - 1) * Cfg::NE) * L.nslabs + slab;
        int ue = Cfg::NL;
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int c = k; c < NV; ++c) { Us[st - 1][k][c] = ok ? __ldcs(fac + (int64_t)ue * L.nslabs) : 1.0; ++ue; }
#pragma unroll
        for (int k = 0; k < NV; ++k)
            ts[st - 1][k] = ok ? __ldcs(L.e + ((int64_t)(st - 1) * NV + k) * L.nslabs + slab) : 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                const double ev = ok ? __ldcs(fac + (int64_t)ue * L.nslabs) : 0.0;
                ts[st - 1][k] = fma(-ev, Ys[c], ts[st - 1][k]);
                ++ue;
            }
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int c = 0; c < NV; ++c) { Fs[st - 1][k][c] = ok ? __ldcs(fac + (int64_t)ue * L.nslabs) : 0.0; ++ue; }
    }
#pragma unroll
    for (int st = S - 1; st >= 1; --st) {
        if (FULL || st < len) {
            double t[NV];
            const double (*U)[NV] = Us[st - 1];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                t[k] = ts[st - 1][k];
#pragma unroll
                for (int c = 0; c < NV; ++c) t[k] = fma(-Fs[st - 1][k][c], Yn[c], t[k]);
            }
#pragma unroll
            for (int k = NV - 1; k >= 0; --k) {
                double s = t[k];
#pragma unroll
                for (int c = k + 1; c < NV; ++c) s = fma(-U[k][c], Yn[c], s);
                Yn[k] = s * U[k][k];  // U[k][k] holds 1/pivot
            }
            const int64_t g = g0 + st;
#pragma unroll
            for (int c = 0; c < NV; ++c) L.Z[g * NV + c] = Yn[c];
            if constexpr (REDUCE) tps_red_visit<NV>(*R, *acc, g, Yn, ystate[st], rlo, rhi, rrt);
        }
    }
}

template <int NV, int S>
__device__ __forceinline__ void tps_backward_slab(const TpsLevel &L, int64_t slab) {
    if ((slab + 1) * S <= L.n) tps_backward_slab_impl<NV, S, true>(L, slab);
    else tps_backward_slab_impl<NV, S, false>(L, slab);
}

// ---- one launch per level (big levels) ------------------------------------------------------------
template <int NV, int S, bool IDENT_B>
__global__ void __launch_bounds__(128) tps_factor_kernel(TpsLevel L, unsigned *flags) {
    const int64_t slab = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slab < L.nslabs) tps_factor_slab<NV, S, IDENT_B>(L, slab, flags);
}
template <int NV, int S>
__global__ void __launch_bounds__(128) tps_forward_kernel(TpsLevel L) {
    pdl_launch_dependents();   // programmatic dependent launch (rst_common.cuh): scheduled under the predecessor's tail
    pdl_wait();
    const int64_t slab = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slab < L.nslabs) tps_forward_slab<NV, S>(L, slab);
}
template <int NV, int S>
__global__ void __launch_bounds__(128) tps_backward_kernel(TpsLevel L) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t slab = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slab < L.nslabs) tps_backward_slab<NV, S>(L, slab);
}

// level-0 backward sweep with the damping reductions fused in; one partial result per block
template <int NV, int S>
__global__ void __launch_bounds__(128) tps_backward_reduce_kernel(TpsLevel L, TpsReduce R) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t slab = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TpsRedAcc acc{0.0, 0.0, 1.0, 0.0};
    if (slab < L.nslabs) {
        if ((slab + 1) * S <= L.n) tps_backward_slab_impl<NV, S, true, true>(L, slab, &R, &acc);
        else tps_backward_slab_impl<NV, S, false, true>(L, slab, &R, &acc);
    }
    __shared__ double sh[4][4];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        acc.ss += __shfl_xor_sync(RST_FULL_MASK, acc.ss, off);
        acc.tol += __shfl_xor_sync(RST_FULL_MASK, acc.tol, off);
        acc.fb = fmin(acc.fb, __shfl_xor_sync(RST_FULL_MASK, acc.fb, off));
        acc.nanf = fmax(acc.nanf, __shfl_xor_sync(RST_FULL_MASK, acc.nanf, off));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][w] = acc.ss; sh[1][w] = acc.tol; sh[2][w] = acc.fb; sh[3][w] = acc.nanf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 1.0, a3 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            a0 += sh[0][i]; a1 += sh[1][i]; a2 = fmin(a2, sh[2][i]); a3 = fmax(a3, sh[3][i]);
        }
        double *p = R.partials + 4 * (int64_t)blockIdx.x;
        p[0] = a0; p[1] = a1; p[2] = a2; p[3] = a3;
    }
}

// ---- the tail: every remaining level inside one CTA -----------------------------------------------
template <int NV, int S>
__global__ void __launch_bounds__(256) tps_tail_factor_kernel(TpsTail T, P2pIface I, unsigned long long seq,
                                                              unsigned *flags) {
    for (int l = 0; l < T.nlevels; ++l) {
        const TpsLevel &L = T.lv[l];
        for (int64_t slab = threadIdx.x; slab < L.nslabs; slab += blockDim.x) {
            // one generic (predicated-load) variant per case keeps the single CTA's straight-line code -- and with it
            // the cold instruction-fetch cost that dominates a one-CTA kernel -- small
            if (L.B == nullptr) tps_factor_slab_impl<NV, S, true, false>(L, slab, flags);  // small meshes: level 0 in the tail
            else tps_factor_slab_impl<NV, S, false, false>(L, slab, flags);
        }
        __syncthreads();
    }
    // multi-GPU: gather every rank's interface row over NVLink and factor the interface system (same launch)
    if (I.world > 1) iface_factor_block<NV>(I, seq, flags);
}

// boundary closure, serial in one thread (nv x nv, nv tiny): C Z0 + D Z1 = g with pinned components removed
template <int NV>
__device__ void tps_top_solve(const double *C, const double *D, const double *g, double *Z,
                              unsigned long long left_mask, unsigned long long right_mask, unsigned *flags) {
    double K[NV][NV + 1];
    int colmap[NV];
    int nc = 0;
    for (int c = 0; c < NV; ++c)
        if (!((left_mask >> c) & 1ull)) { if (nc < NV) colmap[nc] = c; ++nc; }
    for (int c = 0; c < NV; ++c)
        if (!((right_mask >> c) & 1ull)) { if (nc < NV) colmap[nc] = NV + c; ++nc; }
    if (nc != NV) { atomicOr(flags, RST_FLAG_ZERO_PIVOT); return; }
    for (int row = 0; row < NV; ++row) {
        for (int cc = 0; cc < NV; ++cc) {
            const int cm = colmap[cc];
            K[row][cc] = cm < NV ? C[row * NV + cm] : D[row * NV + (cm - NV)];
        }
        K[row][NV] = g[row];
    }
    for (int k = 0; k < NV; ++k) {
        int p = k;
        double best = fabs(K[k][k]);
        for (int row = k + 1; row < NV; ++row)
            if (fabs(K[row][k]) > best) { best = fabs(K[row][k]); p = row; }
        if (!(best > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
        if (p != k)
            for (int c = 0; c <= NV; ++c) { const double tmp = K[k][c]; K[k][c] = K[p][c]; K[p][c] = tmp; }
        for (int row = k + 1; row < NV; ++row) {
            const double l = K[row][k] / K[k][k];
            for (int c = k + 1; c <= NV; ++c) K[row][c] = fma(-l, K[k][c], K[row][c]);
        }
    }
    for (int row = NV - 1; row >= 0; --row) {
        double s = K[row][NV];
        for (int c = row + 1; c < NV; ++c) s = fma(-K[row][c], K[c][NV], s);
        K[row][NV] = s / K[row][row];
    }
    for (int c = 0; c < 2 * NV; ++c) Z[c] = 0.0;
    for (int cc = 0; cc < NV; ++cc) Z[colmap[cc]] = K[cc][NV];
}

// phases: bit 0 forward sweeps, bit 1 boundary closure (single GPU), bit 3 interface exchange + solve (multi GPU),
// bit 2 backward sweeps
template <int NV, int S>
__global__ void __launch_bounds__(256) tps_tail_solve_kernel(TpsTail T, P2pIface I, unsigned long long seq, int phases,
                                                             unsigned *flags) {
    pdl_launch_dependents();
    // The tail runs right after the level-0..2 sweeps have streamed > 100 MB through the L2: its own factors (a few
    // hundred KB) are gone by then and every slab step would pay a dependent DRAM round trip.  Pull them back in one burst.
    if (phases & 5) {
        for (int l = 0; l < T.nlevels; ++l) {
            const TpsLevel &L = T.lv[l];
            const size_t fac_bytes = (size_t)(S - 1) * TpsCfg<NV>::NE * L.nslabs * sizeof(double);
            const char *pf = reinterpret_cast<const char *>(L.fac);
            for (size_t off = (size_t)threadIdx.x * 128; off < fac_bytes; off += (size_t)blockDim.x * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + off));
            const size_t bit_bytes = (size_t)(S - 1) * L.nslabs * sizeof(unsigned);
            const char *pb = reinterpret_cast<const char *>(L.bits);
            for (size_t off = (size_t)threadIdx.x * 128; off < bit_bytes; off += (size_t)blockDim.x * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + off));
        }
    }
    pdl_wait();                // the factors travel while the predecessor finishes
    if (phases & 1) {
        for (int l = 0; l < T.nlevels; ++l) {
            const TpsLevel &L = T.lv[l];
            for (int64_t slab = threadIdx.x; slab < L.nslabs; slab += blockDim.x) tps_forward_slab_impl<NV, S, false>(L, slab);
            __syncthreads();
        }
    }
    if (phases & 2) {
        if (threadIdx.x == 0) {
            const TpsLevel &L = T.lv[T.nlevels - 1];
            tps_top_solve<NV>(L.A_next, L.B_next, L.r_next, T.Ztop, T.left_mask, T.right_mask, flags);
        }
        __syncthreads();
    }
    if (phases & 8) iface_solve_block<NV>(I, seq, flags);
    if (phases & 4) {
        for (int l = T.nlevels - 1; l >= 0; --l) {
            const TpsLevel &L = T.lv[l];
            for (int64_t slab = threadIdx.x; slab < L.nslabs; slab += blockDim.x) tps_backward_slab_impl<NV, S, false>(L, slab);
            __syncthreads();
        }
    }
}

}  // namespace rst
