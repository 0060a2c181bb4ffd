// abd.cuh -- pivoted recursive condensation solver for the block-bidiagonal ("almost block diagonal")
// Newton systems of the discretised BVP (north_star subsystem 2).
//
// Replaces, on the GPU, the reference's serial scalar banded LU
//   src/somelinalg/banded/lapack_style_banded.rs:386-475 (factor), :1163-1225 (solve)
// and its block-tridiagonal LU (block_tridiagonal_lu_consistent.rs:98-218).  The linear system is
//   A_j dY_j + B_j dY_{j+1} = r_j ,  j = 0..n-1   (A_j = -I - h f_y, B_j = I for the forward scheme,
//   src/numerical/BVP_Damp/numeric_discretization.rs:328-342) plus nv pinned boundary components.
// Diagonal blocks of the reduced-ordering matrix are structurally singular,
// so every elimination uses row partial pivoting across the two block rows that share a node:
//
//   slab of S consecutive block rows -> march through the slab; at each step the shared node Y_j of the
//   running condensed row [C | D] (couples slab-start node Y_s and Y_j) and the next row [A_j | B_j] is
//   eliminated by LU with partial pivoting on the stacked 2nv x nv column [D; A_j].  The nv pivot rows
//   are kept for back-substitution (U Y_j = e - E Y_s - F Y_{j+1}), the nv surviving rows become the new
//   condensed row.  Each slab ends as ONE block row coupling its two end nodes -> a block-bidiagonal
//   system S times smaller -> recurse (levels) until one row is left, which is closed with the boundary
//   conditions (abd_top_solve_kernel).  Factor and solve are separate so the damped loop re-uses factors.
//
// Thread mapping ("lanes per slab"): one lane per row of the stacked 2nv-row panel, G = pow2 >= 2nv
// lanes per slab (nv <= 16), 32/G slabs per warp; a lane keeps its row [M | P | Q] (coefficients of the
// node being eliminated, of the slab-start node, of the next node) in registers; pivot search is a
// warp-shuffle arg-max, the pivot row is broadcast with shuffles.  Rows never move between lanes: the
// nv surviving lanes carry the condensed row, the nv pivot lanes are re-used to load the next block row.
//
// HBM layout of the factors (per eliminated node g, lane-fastest so every access is coalesced):
//   Lm [g][k][r]   k < nv, r < 2nv : multiplier of row r at elimination step k (valid while unpivoted)
//   UEF[g][c][k]   c < 3nv, k < nv : pivot row k: U (c < nv, valid for c >= k) | E (Y_s coeffs) | F (Y_{j+1} coeffs)
//   piv[g][r]      int8           : elimination step at which row r became the pivot row, -1 = survivor
#pragma once
#include "rst_common.cuh"

namespace rst {

struct AbdLevel {
    int64_t n;         // block rows at this level
    int64_t nslabs;    // ceil(n / S) = block rows of the next level
    int S;             // slab length (rows per slab)
    int nv;            // block size (used by the run-time-nv kernels of big.cuh)
    int prefetch;      // != 0: the sweeps prefetch the next block row's factors into L2 while the current one is processed
    const double *A;   // [n][nv][nv] row-major
    const double *B;   // [n][nv][nv] or nullptr (identity)
    double *A_next;    // [nslabs][nv][nv]
    double *B_next;    // [nslabs][nv][nv]
    double *Lm;        // [n][nv][2nv]
    double *UEF;       // [n][3nv][nv]
    int8_t *piv;       // [n][2nv]
    const double *r;   // [n][nv]      right-hand side of this level
    double *e;         // [n][nv]      transformed rhs of the pivot rows
    double *r_next;    // [nslabs][nv]
    double *Z;         // [n+1][nv]    node values of this level (level 0: dY)
    const double *Z_next;  // [nslabs+1][nv]
};

// rank of this lane among the set bits of `bits` restricted to its group (bits already shifted to the group)
__device__ __forceinline__ int rank_in(unsigned gbits, int r) { return __popc(gbits & ((1u << r) - 1u)); }

template <int G>
__device__ __forceinline__ unsigned group_bits(unsigned ballot, int gw) {
    if constexpr (G == 32) return ballot;
    else return (ballot >> (gw * G)) & ((1u << G) - 1u);
}

// Packed factor layout of the lanes-per-slab kernels (per block row g):
//   Lm  [k][rank]   only the rows that are updated at step k (not yet pivoted, not the pivot row) store a multiplier, compacted
//                   by their rank among those rows: 2nv - k - 1 entries at offset k (2nv - 1) - k (k - 1) / 2
//                   -> 1.5 nv^2 - nv/2 doubles instead of 2 nv^2
//   UEF U part [c][k <= c]  upper triangle incl. diagonal at offset c (c + 1) / 2 + k; E, F full [c][k] behind it
//                   -> 2.5 nv^2 + nv/2 doubles instead of 3 nv^2
// 20 % less solve traffic than the dense layout (which the run-time-nv kernels of big.cuh keep, in the same buffers).
template <int NV>
struct AbdPack {
    static constexpr int R2 = 2 * NV;
    static constexpr int LM = NV * (R2 - 1) - NV * (NV - 1) / 2;
    static constexpr int TRI = NV * (NV + 1) / 2;
    static constexpr int UEF = TRI + 2 * NV * NV;
    __host__ __device__ static constexpr int lm_off(int k) { return k * (R2 - 1) - k * (k - 1) / 2; }
    __host__ __device__ static constexpr int u_off(int c) { return c * (c + 1) / 2; }
};

// factors of one block row from the register panel (all lanes of the warp must call: ballots inside)
template <int NV, int G>
__device__ __forceinline__ void abd_store_factors(const AbdLevel &L, int64_t g, int r, int gw, bool live, int pv,
                                                  const double *M, const double *P, const double *Q) {
    using Pk = AbdPack<NV>;
    double *lm = L.Lm + g * (int64_t)Pk::LM;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const bool upd = live && (pv < 0 || pv > k);
        const unsigned m = group_bits<G>(__ballot_sync(RST_FULL_MASK, upd), gw);
        if (upd) st_cs(lm + Pk::lm_off(k) + rank_in(m, r), M[k]);
    }
    if (live) {
        L.piv[g * Pk::R2 + r] = (int8_t)pv;
        if (pv >= 0) {
            double *u = L.UEF + g * (int64_t)Pk::UEF + pv;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                if (c >= pv) st_cs(u + Pk::u_off(c), M[c]);
                st_cs(u + Pk::TRI + c * NV, P[c]);
                st_cs(u + Pk::TRI + NV * NV + c * NV, Q[c]);
            }
        }
    }
}

// arg-max of |v| over the lanes of a group that are still candidates; returns the group-relative lane.
// Key = bits(|v|) with the low 5 mantissa bits replaced by (31 - r): ties resolve to the lowest row,
// the comparison ignores the last 5 of 52 mantissa bits (irrelevant for stability).
template <int G>
__device__ __forceinline__ int group_argmax(double v, bool cand, int r) {
    long long key = cand ? ((__double_as_longlong(fabs(v)) & ~31LL) | (long long)(31 - r)) : -1LL;
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
        const long long o = __shfl_xor_sync(RST_FULL_MASK, key, off);
        key = o > key ? o : key;
    }
    return 31 - (int)(key & 31LL);
}

// ------------------------------------------------------------------------------------------------
// factor: one level of condensation
// ------------------------------------------------------------------------------------------------
template <int NV, bool IDENT_B>
__global__ void __launch_bounds__(128) abd_factor_kernel(AbdLevel L, unsigned *flags) {
    constexpr int R2 = 2 * NV;
    constexpr int G = next_pow2(R2);
    static_assert(G <= 32, "lanes-per-slab kernel supports nv <= 16");
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int r = lane & (G - 1);
    const int gw = lane / G;
    const int64_t slab = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + gw;
    const bool slab_ok = slab < L.nslabs;
    const bool row_ok = r < R2;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = slab_ok ? (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S) : 0;

    double M[NV], P[NV], Q[NV];
    bool carry = slab_ok && r < NV;
#pragma unroll
    for (int c = 0; c < NV; ++c) { M[c] = 0.0; P[c] = 0.0; Q[c] = 0.0; }
    if (carry) {
        const double *a = L.A + (g0 * NV + r) * NV;
#pragma unroll
        for (int c = 0; c < NV; ++c) P[c] = a[c];
        if (IDENT_B) {
#pragma unroll
            for (int c = 0; c < NV; ++c) M[c] = (c == r) ? 1.0 : 0.0;
        } else {
            const double *b = L.B + (g0 * NV + r) * NV;
#pragma unroll
            for (int c = 0; c < NV; ++c) M[c] = b[c];
        }
    }

    for (int i = 1; i < L.S; ++i) {
        const bool active = i < len;  // uniform per group; every lane still joins the shuffles
        const int64_t g = g0 + i;
        const bool is_free = row_ok && !carry;
        const unsigned fb = group_bits<G>(__ballot_sync(RST_FULL_MASK, is_free), gw);
        if (active && is_free) {
            const int t = rank_in(fb, r);
            const double *a = L.A + (g * NV + t) * NV;
#pragma unroll
            for (int c = 0; c < NV; ++c) { M[c] = a[c]; P[c] = 0.0; }
            if (IDENT_B) {
#pragma unroll
                for (int c = 0; c < NV; ++c) Q[c] = (c == t) ? 1.0 : 0.0;
            } else {
                const double *b = L.B + (g * NV + t) * NV;
#pragma unroll
                for (int c = 0; c < NV; ++c) Q[c] = b[c];
            }
        }
        int pv = -1;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const bool cand = active && row_ok && pv < 0;
            const int p = group_argmax<G>(M[k], cand, r);
            const double pivval = shfl_d(M[k], p, G);
            const bool is_piv = cand && (r == p);
            if (is_piv) {
                pv = k;
                if (!(fabs(pivval) > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
            }
            const bool upd = cand && !is_piv;
            const double l = upd ? M[k] / pivval : 0.0;
            if (upd) M[k] = l;
#pragma unroll
            for (int c = k + 1; c < NV; ++c) {
                const double mp = shfl_d(M[c], p, G);
                M[c] = fma(-l, mp, M[c]);
            }
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                const double pp = shfl_d(P[c], p, G);
                P[c] = fma(-l, pp, P[c]);
                const double qp = shfl_d(Q[c], p, G);
                Q[c] = fma(-l, qp, Q[c]);
            }
        }
        abd_store_factors<NV, G>(L, g, r, gw, active && row_ok, pv, M, P, Q);
        if (active) {
            carry = row_ok && pv < 0;
            if (carry) {
#pragma unroll
                for (int c = 0; c < NV; ++c) { M[c] = Q[c]; Q[c] = 0.0; }
            }
        }
    }
    const unsigned cb = group_bits<G>(__ballot_sync(RST_FULL_MASK, carry), gw);
    if (slab_ok && carry) {
        const int t = rank_in(cb, r);
        double *a = L.A_next + (slab * NV + t) * NV;
        double *b = L.B_next + (slab * NV + t) * NV;
#pragma unroll
        for (int c = 0; c < NV; ++c) { a[c] = P[c]; b[c] = M[c]; }
    }
}

// ------------------------------------------------------------------------------------------------
// solve, forward sweep of one level: apply the recorded eliminations to the right-hand side
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(128) abd_forward_kernel(AbdLevel L) {
    using Pk = AbdPack<NV>;
    constexpr int R2 = 2 * NV;
    constexpr int G = next_pow2(R2);
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int r = lane & (G - 1);
    const int gw = lane / G;
    const int64_t slab = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + gw;
    const bool slab_ok = slab < L.nslabs;
    const bool row_ok = r < R2;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = slab_ok ? (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S) : 0;

    bool carry = slab_ok && r < NV;
    double w = carry ? L.r[g0 * NV + r] : 0.0;

    for (int i = 1; i < L.S; ++i) {
        const bool active = i < len;
        const int64_t g = g0 + i;
        const bool is_free = row_ok && !carry;
        const unsigned fb = group_bits<G>(__ballot_sync(RST_FULL_MASK, is_free), gw);
        if (L.prefetch && i + 1 < len && r * 128 < Pk::LM * 8)   // one line per lane; hides the per-block-row round trip
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(L.Lm + (g + 1) * (int64_t)Pk::LM) + r * 128));
        int pv = -1;
        double lrow[NV];
        const bool live = active && row_ok;
        if (live) {
            if (is_free) w = L.r[g * NV + rank_in(fb, r)];
            pv = (int)L.piv[g * R2 + r];
        }
        {
            const double *lm = L.Lm + g * (int64_t)Pk::LM;
#pragma unroll
            for (int k = 0; k < NV; ++k) {   // packed multipliers: rank among the rows updated at step k
                const bool upd = live && (pv < 0 || pv > k);
                const unsigned m = group_bits<G>(__ballot_sync(RST_FULL_MASK, upd), gw);
                lrow[k] = upd ? ld_cs(lm + Pk::lm_off(k) + rank_in(m, r)) : 0.0;
            }
        }
        // (deriving the ranks from the pivot ballots of the chain below saves 12 ballots, but ptxas then interleaves the
        // loads with the chain and exposes one memory round trip per step: measured 1.56 ms vs 1.15 ms per C3 solve)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const unsigned pb = group_bits<G>(__ballot_sync(RST_FULL_MASK, pv == k), gw);
            const int p = pb ? (__ffs(pb) - 1) : 0;
            const double wp = shfl_d(w, p, G);
            if (pv < 0 || pv > k) w = fma(-lrow[k], wp, w);
        }
        if (active && row_ok) {
            if (pv >= 0) L.e[g * NV + pv] = w;
            carry = pv < 0;
        }
    }
    const unsigned cb = group_bits<G>(__ballot_sync(RST_FULL_MASK, carry), gw);
    if (slab_ok && carry) L.r_next[slab * NV + rank_in(cb, r)] = w;
}

// ------------------------------------------------------------------------------------------------
// solve, backward sweep of one level: recover the interior nodes of every slab from its end nodes
// ------------------------------------------------------------------------------------------------
// body for one warp: `warp_index` counts warps over the level (32 / G slabs each)
template <int NV>
__device__ __forceinline__ void abd_backward_warp(const AbdLevel &L, int64_t warp_index, int lane) {
    using Pk = AbdPack<NV>;
    constexpr int G = next_pow2(NV);
    constexpr int GPW = 32 / G;
    const int k = lane & (G - 1);
    const int gw = lane / G;
    const int64_t slab = warp_index * GPW + gw;
    const bool slab_ok = slab < L.nslabs;
    const bool k_ok = slab_ok && k < NV;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = slab_ok ? (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S) : 0;

    double Ys[NV], Yn[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        Ys[c] = slab_ok ? L.Z_next[slab * NV + c] : 0.0;
        Yn[c] = slab_ok ? L.Z_next[(slab + 1) * NV + c] : 0.0;
    }
    if (k_ok) {
        double ys = 0.0, yn = 0.0;
#pragma unroll
        for (int c = 0; c < NV; ++c) { ys = (c == k) ? Ys[c] : ys; yn = (c == k) ? Yn[c] : yn; }
        L.Z[g0 * NV + k] = ys;
        if (slab == L.nslabs - 1) L.Z[L.n * NV + k] = yn;
    }
    for (int i = L.S - 1; i >= 1; --i) {
        const bool active = i < len;
        const int64_t g = g0 + i;
        double U[NV];
        double t = 0.0, inv = 1.0;
#pragma unroll
        for (int c = 0; c < NV; ++c) U[c] = 0.0;
        if (L.prefetch && slab_ok && i - L.prefetch >= 1) {   // L.prefetch = distance in block rows
            const char *nx = reinterpret_cast<const char *>(L.UEF + (g - L.prefetch) * (int64_t)Pk::UEF);
            for (int off = (lane & (G - 1)) * 128; off < Pk::UEF * 8; off += G * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + off));
        }
        if (active && k_ok) {
            const double *u = L.UEF + g * (int64_t)Pk::UEF + k;
            t = L.e[g * NV + k];
            double diag = 1.0;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                U[c] = (c >= k) ? ld_cs(u + Pk::u_off(c)) : 0.0;   // upper triangle only
                t = fma(-ld_cs(u + Pk::TRI + c * NV), Ys[c], t);
                t = fma(-ld_cs(u + Pk::TRI + NV * NV + c * NV), Yn[c], t);
                diag = (c == k) ? U[c] : diag;
            }
            inv = 1.0 / diag;
        }
#pragma unroll
        for (int kk = NV - 1; kk >= 0; --kk) {
            const double y = shfl_d(t * inv, kk, G);
            if (k < kk) t = fma(-U[kk], y, t);
            if (active) Yn[kk] = y;
        }
        if (active && k_ok) {
            double mine = 0.0;
#pragma unroll
            for (int c = 0; c < NV; ++c) mine = (c == k) ? Yn[c] : mine;
            L.Z[g * NV + k] = mine;
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(128) abd_backward_kernel(AbdLevel L) {
    const int64_t warp_index = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    {   // U | E | F of the last block row of this lane group's slab
        constexpr int G = next_pow2(NV);
        const int64_t slab = warp_index * (32 / G) + lane / G;
        if (slab < L.nslabs) {
            const int64_t g0 = slab * (int64_t)L.S;
            const int64_t last = g0 + ((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S) - 1;
            const char *p = reinterpret_cast<const char *>(L.UEF + last * (int64_t)AbdPack<NV>::UEF);
            for (int off = (lane & (G - 1)) * 128; off < AbdPack<NV>::UEF * 8; off += G * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
        }
    }
    pdl_wait();
    abd_backward_warp<NV>(L, warp_index, lane);
}

// ------------------------------------------------------------------------------------------------
// top of the recursion: the single remaining block row  C Z0 + D Z1 = g  closed with the boundary
// conditions (pinned components have zero correction): an nv x nv dense system in the free components
// of Z0 and Z1 (numeric_discretization.rs:80-106 deletes exactly those columns), solved by one warp
// with partial pivoting in shared memory.
// ------------------------------------------------------------------------------------------------
template <int NV>
struct AbdTopSmem {
    int colmap[NV];
};

// one warp, lane = row of the nv x nv closure system, the row lives in registers: pivot search with one REDUX per column
// (rows never move; a retired row simply stops being a candidate), pivot row broadcast by shuffles, substitution by
// shuffles.  (The first version kept K in shared memory and let lane 0 substitute: 17 us per linear solve at nv = 12.)
template <int NV>
__device__ __forceinline__ void abd_top_solve_warp(AbdTopSmem<NV> &sm, const double *C, const double *D, const double *g,
                                                   double *Z /* [2][nv] */, unsigned long long left_mask,
                                                   unsigned long long right_mask, unsigned *flags, int lane) {
    static_assert(NV <= 32, "one row per lane");
    int *colmap = sm.colmap;
    if (lane == 0) {
        int nc = 0;
        for (int c = 0; c < NV; ++c)
            if (!((left_mask >> c) & 1ull)) colmap[nc++] = c;
        for (int c = 0; c < NV; ++c)
            if (!((right_mask >> c) & 1ull)) { if (nc < NV) colmap[nc] = NV + c; ++nc; }
        if (nc != NV) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
    }
    __syncwarp();
    const bool row_ok = lane < NV;
    const int row = row_ok ? lane : 0;
    double a[NV], b = row_ok ? g[row] : 0.0;
#pragma unroll
    for (int cc = 0; cc < NV; ++cc) {
        const int cm = colmap[cc];
        a[cc] = row_ok ? (cm < NV ? C[row * NV + cm] : D[row * NV + (cm - NV)]) : 0.0;
    }
    bool cand = row_ok;
    int ps[NV];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        // exact arg-max over the candidate rows: high word, then low word among the rows that tie on it
        const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(a[k]));
        const unsigned hi = cand ? (unsigned)(bits >> 32) + 1u : 0u;
        const unsigned mhi = __reduce_max_sync(RST_FULL_MASK, hi);
        const bool top = cand && hi == mhi;
        const unsigned mlo = __reduce_max_sync(RST_FULL_MASK, top ? (unsigned)bits : 0u);
        const unsigned win = __ballot_sync(RST_FULL_MASK, top && (unsigned)bits == mlo);
        const int p = win ? (__ffs(win) - 1) : 0;
        ps[k] = p;
        const double piv = __shfl_sync(RST_FULL_MASK, a[k], p);
        bad = bad || !(fabs(piv) > 0.0);
        const bool upd = cand && lane != p;
        const double l = upd ? a[k] / piv : 0.0;
#pragma unroll
        for (int c = k + 1; c < NV; ++c) a[c] = fma(-l, __shfl_sync(RST_FULL_MASK, a[c], p), a[c]);
        b = fma(-l, __shfl_sync(RST_FULL_MASK, b, p), b);
        if (lane == p) cand = false;
    }
    if (bad && lane == 0) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
    // back-substitution in pivot order: row ps[k] determines unknown k
    double x[NV];
    bool done = !row_ok;
#pragma unroll
    for (int k = NV - 1; k >= 0; --k) {
        const double xk = __shfl_sync(RST_FULL_MASK, b / a[k], ps[k]);
        x[k] = xk;
        if (lane == ps[k]) done = true;
        if (!done) b = fma(-a[k], xk, b);
    }
    if (lane == 0) {
        for (int c = 0; c < 2 * NV; ++c) Z[c] = 0.0;
#pragma unroll
        for (int cc = 0; cc < NV; ++cc) Z[colmap[cc]] = x[cc];
    }
}

template <int NV>
__global__ void __launch_bounds__(32) abd_top_solve_kernel(const double *C, const double *D, const double *g,
                                                           double *Z /* [2][nv] */, unsigned long long left_mask,
                                                           unsigned long long right_mask, unsigned *flags) {
    __shared__ AbdTopSmem<NV> sm;
    abd_top_solve_warp<NV>(sm, C, D, g, Z, left_mask, right_mask, flags, threadIdx.x);
}

}  // namespace rst

namespace rst {

// ------------------------------------------------------------------------------------------------
// Closure with GENERAL separated boundary rows (used by the block-tridiagonal drop-in, btd.cuh):
//   Ba Z0 = ba (na rows),  C Z0 + D Z1 = g (nv rows),  Bb Z1 = bb (nb rows),  na + nb = nv
// -> one dense 2nv x 2nv system, partial pivoting in shared memory, one warp.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(32) abd_top_general_kernel(const double *C, const double *D, const double *g,
                                                             const double *Ba, const double *ba, int na,
                                                             const double *Bb, const double *bb, int nb, double *Z,
                                                             unsigned *flags) {
    constexpr int N2 = 2 * NV;
    __shared__ double K[N2][N2 + 1];
    const int lane = threadIdx.x;
    for (int idx = lane; idx < N2 * (N2 + 1); idx += 32) (&K[0][0])[idx] = 0.0;
    __syncwarp();
    for (int idx = lane; idx < na * NV; idx += 32) K[idx / NV][idx % NV] = Ba[idx];
    for (int idx = lane; idx < na; idx += 32) K[idx][N2] = ba[idx];
    for (int idx = lane; idx < NV * NV; idx += 32) {
        K[na + idx / NV][idx % NV] = C[idx];
        K[na + idx / NV][NV + idx % NV] = D[idx];
    }
    for (int idx = lane; idx < NV; idx += 32) K[na + idx][N2] = g[idx];
    for (int idx = lane; idx < nb * NV; idx += 32) K[na + NV + idx / NV][NV + idx % NV] = Bb[idx];
    for (int idx = lane; idx < nb; idx += 32) K[na + NV + idx][N2] = bb[idx];
    __syncwarp();
    for (int k = 0; k < N2; ++k) {
        double best = -1.0;
        int bi = k;
        for (int row = k + lane; row < N2; row += 32) {
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
            for (int c = lane; c <= N2; c += 32) { const double tmp = K[k][c]; K[k][c] = K[bi][c]; K[bi][c] = tmp; }
        __syncwarp();
        const double pivval = K[k][k];
        for (int row = k + 1 + lane; row < N2; row += 32) {
            const double l = K[row][k] / pivval;
            for (int c = k + 1; c <= N2; ++c) K[row][c] = fma(-l, K[k][c], K[row][c]);
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (int row = N2 - 1; row >= 0; --row) {
            double s = K[row][N2];
            for (int c = row + 1; c < N2; ++c) s = fma(-K[row][c], K[c][N2], s);
            K[row][N2] = s / K[row][row];
        }
        for (int c = 0; c < N2; ++c) Z[c] = K[c][N2];
    }
}

}  // namespace rst
