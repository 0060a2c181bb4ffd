// abd3.cuh -- third-generation factor kernel of the pivoted condensation solver (algorithm and factor layout:
// abd.cuh; it pairs with abd_forward_kernel / abd_backward_kernel of abd.cuh).
//
// Same one-panel-row-per-lane mapping as abd_factor_kernel (2nv lanes per slab, 122 registers -> 16 warps/SM),
// with the three costs the ncu captures singled out removed (profiles/r01_a_summary.md, r01_b_summary.md):
//   * pivot search: two REDUX.MAX (hi / lo word of |value|) + one ballot instead of a 5-round 64-bit shuffle
//     butterfly; exact 64-bit comparison, ties to the lowest row;
//   * pivot-row broadcast: the owning lane publishes [M(k..) | P | Q] to shared memory with STS.128, all lanes read
//     it back as LDS.128 broadcasts (double-buffered: one __syncwarp per pivot) -- 30 LSU instructions per pivot
//     instead of ~70 shuffles;
//   * multipliers: one correctly rounded reciprocal of the pivot + a multiply instead of a full FP64 division.
#pragma once
#include "abd.cuh"

namespace rst {

template <int G>
__device__ __forceinline__ unsigned group_mask(int gw) {
    if constexpr (G == 32) return RST_FULL_MASK;
    else return ((1u << G) - 1u) << (gw * G);
}

// exact arg-max of |v| over candidate lanes of the group; returns group-relative lane (0 when there is none)
template <int G>
__device__ __forceinline__ int group_argmax_redux(double v, bool cand, int gw, unsigned gmask) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(v));
    // +1 so that a candidate holding an exact zero still beats a non-candidate
    const unsigned hi = cand ? (unsigned)(bits >> 32) + 1u : 0u;
    const unsigned mhi = __reduce_max_sync(gmask, hi);
    const bool top = cand && hi == mhi;
    const unsigned lo = top ? (unsigned)bits : 0u;
    const unsigned mlo = __reduce_max_sync(gmask, lo);
    const unsigned win = group_bits<G>(__ballot_sync(gmask, top && lo == mlo), gw);
    return win ? (__ffs(win) - 1) : 0;
}

template <int NV, bool IDENT_B>
__global__ void __launch_bounds__(128, 4) abd_factor3_kernel(AbdLevel L, unsigned *flags) {
    constexpr int R2 = 2 * NV;
    constexpr int G = next_pow2(R2);
    static_assert(G <= 32 && NV % 2 == 0, "kernel supports even nv <= 16");
    constexpr int GPW = 32 / G;
    constexpr int WP = 3 * NV;  // even
    __shared__ __align__(16) double bc[4][GPW][2][WP];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int r = lane & (G - 1);
    const int gw = lane / G;
    const unsigned gmask = group_mask<G>(gw);
    const int64_t slab = ((int64_t)blockIdx.x * 4 + warp) * GPW + gw;
    const bool slab_ok = slab < L.nslabs;
    const bool row_ok = r < R2;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = slab_ok ? (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S) : 0;

    double M[NV], P[NV], Q[NV];
    bool carry = slab_ok && r < NV;
#pragma unroll
    for (int c = 0; c < NV; ++c) { M[c] = 0.0; P[c] = 0.0; Q[c] = 0.0; }
    if (carry) {
        const double2 *a = reinterpret_cast<const double2 *>(L.A + (g0 * NV + r) * NV);
#pragma unroll
        for (int c = 0; c < NV; c += 2) { const double2 v = a[c >> 1]; P[c] = v.x; P[c + 1] = v.y; }
        if (IDENT_B) {
#pragma unroll
            for (int c = 0; c < NV; ++c) M[c] = (c == r) ? 1.0 : 0.0;
        } else {
            const double2 *b = reinterpret_cast<const double2 *>(L.B + (g0 * NV + r) * NV);
#pragma unroll
            for (int c = 0; c < NV; c += 2) { const double2 v = b[c >> 1]; M[c] = v.x; M[c + 1] = v.y; }
        }
    }

    for (int i = 1; i < L.S; ++i) {
        const bool active = i < len;
        const int64_t g = g0 + i;
        const bool is_free = row_ok && !carry;
        const unsigned fb = group_bits<G>(__ballot_sync(RST_FULL_MASK, is_free), gw);
        if (active) {
            if (i + 1 < len && r * 16 < NV * NV)  // L2 prefetch of the next block row, one line per lane
                asm volatile("prefetch.global.L2 [%0];" ::"l"(L.A + (g + 1) * (int64_t)(NV * NV) + r * 16));
            if (is_free) {
                const int t = rank_in(fb, r);
                const double2 *a = reinterpret_cast<const double2 *>(L.A + (g * NV + t) * NV);
#pragma unroll
                for (int c = 0; c < NV; c += 2) {
                    const double2 v = ld_cs2(a + (c >> 1));
                    M[c] = v.x; M[c + 1] = v.y;
                    P[c] = 0.0; P[c + 1] = 0.0;
                }
                if (IDENT_B) {
#pragma unroll
                    for (int c = 0; c < NV; ++c) Q[c] = (c == t) ? 1.0 : 0.0;
                } else {
                    const double2 *b = reinterpret_cast<const double2 *>(L.B + (g * NV + t) * NV);
#pragma unroll
                    for (int c = 0; c < NV; c += 2) { const double2 v = ld_cs2(b + (c >> 1)); Q[c] = v.x; Q[c + 1] = v.y; }
                }
            }
        }
        int pv = -1;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const bool cand = active && row_ok && pv < 0;
            const int p = group_argmax_redux<G>(M[k], cand, gw, gmask);
            const bool is_piv = cand && (r == p);
            double2 *b2 = reinterpret_cast<double2 *>(&bc[warp][gw][k & 1][0]);
            if (r == p) {  // publish the pivot row (one predicated STS.128 sequence, no divergence between groups)
#pragma unroll
                for (int c = (k & ~1); c < NV; c += 2) b2[c >> 1] = make_double2(M[c], M[c + 1]);
#pragma unroll
                for (int c = 0; c < NV; c += 2) {
                    b2[(NV + c) >> 1] = make_double2(P[c], P[c + 1]);
                    b2[(2 * NV + c) >> 1] = make_double2(Q[c], Q[c + 1]);
                }
            }
            __syncwarp();
            const double2 vk = b2[k >> 1];
            const double pivval = (k & 1) ? vk.y : vk.x;
            if (is_piv) {
                pv = k;
                if (!(fabs(pivval) > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
            }
            const bool upd = cand && !is_piv;
            const double l = upd ? M[k] * __drcp_rn(pivval) : 0.0;
            if (upd) M[k] = l;
            if ((k & 1) == 0 && k + 1 < NV) M[k + 1] = fma(-l, vk.y, M[k + 1]);
#pragma unroll
            for (int c = (k & ~1) + 2; c < NV; c += 2) {
                const double2 v = b2[c >> 1];
                M[c] = fma(-l, v.x, M[c]);
                M[c + 1] = fma(-l, v.y, M[c + 1]);
            }
#pragma unroll
            for (int c = 0; c < NV; c += 2) {
                const double2 v = b2[(NV + c) >> 1];
                P[c] = fma(-l, v.x, P[c]);
                P[c + 1] = fma(-l, v.y, P[c + 1]);
                const double2 q = b2[(2 * NV + c) >> 1];
                Q[c] = fma(-l, q.x, Q[c]);
                Q[c + 1] = fma(-l, q.y, Q[c + 1]);
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

}  // namespace rst
