// abd2.cuh -- second-generation factor / forward kernels of the pivoted condensation solver (see abd.cuh for
// the algorithm and the factor layout).  Same mathematics, different mapping, driven by the round-1 ncu capture
// (profiles/r01_a_summary.md: the first kernel was bound by the shuffle pipe that broadcasts the pivot row):
//   * TWO panel rows per lane (nv lanes per slab instead of 2nv): every broadcast value feeds two FMAs and
//     two slabs share one warp instruction stream (nv = 12: 16-lane groups, 2 slabs per warp);
//   * the pivot row [M | P | Q] is broadcast through shared memory (one lane writes 3nv doubles with STS.128,
//     all lanes read them back as LDS.128 broadcasts, double-buffered so one __syncwarp per pivot suffices);
//   * the next block row is prefetched into L2 while the current one is eliminated.
// Panel row id = 2 * lane + slot; factors are stored exactly as abd.cuh documents with r := row id.
#pragma once
#include "abd.cuh"

namespace rst {

template <int NV>
struct Abd2Cfg {
    static constexpr int R2 = 2 * NV;
    static constexpr int G = next_pow2(NV);   // lanes per slab
    static constexpr int GPW = 32 / G;        // slabs per warp
    static constexpr int W = 3 * NV;          // broadcast row length
    static constexpr int WP = (W + 1) & ~1;   // padded to an even count (16-byte rows)
};

__device__ __forceinline__ unsigned below(int r) { return (1u << r) - 1u; }

template <int NV, bool IDENT_B>
__global__ void __launch_bounds__(128) abd_factor2_kernel(AbdLevel L, unsigned *flags) {
    using Cfg = Abd2Cfg<NV>;
    constexpr int R2 = Cfg::R2, G = Cfg::G, GPW = Cfg::GPW, WP = Cfg::WP;
    static_assert(NV % 2 == 0 && NV <= 16, "two-rows-per-lane kernel supports even nv <= 16");
    __shared__ __align__(16) double bc[4][GPW][2][WP];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int r = lane & (G - 1);
    const int gw = lane / G;
    const int64_t slab = ((int64_t)blockIdx.x * 4 + warp) * GPW + gw;
    const bool slab_ok = slab < L.nslabs;
    const bool lane_ok = r < NV;
    const unsigned gmask = G == 32 ? RST_FULL_MASK : (((1u << (G & 31)) - 1u) << (gw * G));
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = slab_ok ? (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S) : 0;

    double M[2][NV], P[2][NV], Q[2][NV];
    bool carry[2];
    carry[0] = slab_ok && lane_ok;
    carry[1] = false;
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int c = 0; c < NV; ++c) { M[s][c] = 0.0; P[s][c] = 0.0; Q[s][c] = 0.0; }
    if (carry[0]) {
        const double *a = L.A + (g0 * NV + r) * NV;
#pragma unroll
        for (int c = 0; c < NV; ++c) P[0][c] = a[c];
        if (IDENT_B) {
#pragma unroll
            for (int c = 0; c < NV; ++c) M[0][c] = (c == r) ? 1.0 : 0.0;
        } else {
            const double *b = L.B + (g0 * NV + r) * NV;
#pragma unroll
            for (int c = 0; c < NV; ++c) M[0][c] = b[c];
        }
    }

    for (int i = 1; i < L.S; ++i) {
        const bool active = i < len;
        const int64_t g = g0 + i;
        const bool free0 = lane_ok && !carry[0], free1 = lane_ok && !carry[1];
        const unsigned f0 = group_bits<G>(__ballot_sync(RST_FULL_MASK, free0), gw);
        const unsigned f1 = group_bits<G>(__ballot_sync(RST_FULL_MASK, free1), gw);
        if (active) {
            const int base = __popc(f0 & below(r)) + __popc(f1 & below(r));
            // L2 prefetch of the block row after this one (one 128-byte line per lane)
            if (i + 1 < len && r * 16 < NV * NV) {
                const double *nx = L.A + (g + 1) * (int64_t)(NV * NV) + r * 16;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const bool fr = s == 0 ? free0 : free1;
                if (fr) {
                    const int t = base + ((s == 1 && free0) ? 1 : 0);
                    const double *a = L.A + (g * NV + t) * NV;
#pragma unroll
                    for (int c = 0; c < NV; ++c) { M[s][c] = a[c]; P[s][c] = 0.0; }
                    if (IDENT_B) {
#pragma unroll
                        for (int c = 0; c < NV; ++c) Q[s][c] = (c == t) ? 1.0 : 0.0;
                    } else {
                        const double *b = L.B + (g * NV + t) * NV;
#pragma unroll
                        for (int c = 0; c < NV; ++c) Q[s][c] = b[c];
                    }
                }
            }
        }
        int pv[2] = {-1, -1};
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const bool cand0 = active && lane_ok && pv[0] < 0;
            const bool cand1 = active && lane_ok && pv[1] < 0;
            // pivot search: 64-bit key = |value| bits with the low 5 mantissa bits replaced by (31 - row id); two REDUX.MAX
            // (hi word, then lo word among the hi-word winners) give every lane the unique maximum key
            const long long k0 = cand0 ? ((__double_as_longlong(fabs(M[0][k])) & ~31LL) | (long long)(31 - 2 * r)) : -1LL;
            const long long k1 = cand1 ? ((__double_as_longlong(fabs(M[1][k])) & ~31LL) | (long long)(30 - 2 * r)) : -1LL;
            const long long kk = k0 > k1 ? k0 : k1;
            const bool cany = cand0 || cand1;
            const unsigned hi = cany ? (unsigned)((unsigned long long)kk >> 32) + 1u : 0u;
            const unsigned mhi = __reduce_max_sync(gmask, hi);
            const unsigned lo = (cany && hi == mhi) ? (unsigned)kk : 0u;
            const unsigned mlo = __reduce_max_sync(gmask, lo);
            const int pid = 31 - (int)(mlo & 31u);
            const int pl = pid >> 1, ps = pid & 1;
            double2 *b2 = reinterpret_cast<double2 *>(&bc[warp][gw][k & 1][0]);
            if (r == pl) {  // owner of the pivot row publishes it (one STS.128 sequence, slot chosen by selects)
#pragma unroll
                for (int c = (k & ~1); c < NV; c += 2)
                    b2[c >> 1] = make_double2(ps ? M[1][c] : M[0][c], ps ? M[1][c + 1] : M[0][c + 1]);
#pragma unroll
                for (int c = 0; c < NV; c += 2) {
                    b2[(NV + c) >> 1] = make_double2(ps ? P[1][c] : P[0][c], ps ? P[1][c + 1] : P[0][c + 1]);
                    b2[(2 * NV + c) >> 1] = make_double2(ps ? Q[1][c] : Q[0][c], ps ? Q[1][c + 1] : Q[0][c + 1]);
                }
            }
            __syncwarp();
            const bool piv0 = cand0 && r == pl && ps == 0;
            const bool piv1 = cand1 && r == pl && ps == 1;
            const double2 vk = b2[k >> 1];  // LDS.128 broadcast
            const double pivval = (k & 1) ? vk.y : vk.x;
            if (piv0 || piv1) {
                if (!(fabs(pivval) > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
            }
            if (piv0) pv[0] = k;
            if (piv1) pv[1] = k;
            const bool u0 = cand0 && !piv0, u1 = cand1 && !piv1;
            const double rc = __drcp_rn(pivval);
            const double l0 = u0 ? M[0][k] * rc : 0.0;
            const double l1 = u1 ? M[1][k] * rc : 0.0;
            if (u0) M[0][k] = l0;
            if (u1) M[1][k] = l1;
            if ((k & 1) == 0 && k + 1 < NV) {
                M[0][k + 1] = fma(-l0, vk.y, M[0][k + 1]);
                M[1][k + 1] = fma(-l1, vk.y, M[1][k + 1]);
            }
#pragma unroll
            for (int c = (k & ~1) + 2; c < NV; c += 2) {
                const double2 v = b2[c >> 1];
                M[0][c] = fma(-l0, v.x, M[0][c]);
                M[1][c] = fma(-l1, v.x, M[1][c]);
                M[0][c + 1] = fma(-l0, v.y, M[0][c + 1]);
                M[1][c + 1] = fma(-l1, v.y, M[1][c + 1]);
            }
#pragma unroll
            for (int c = 0; c < NV; c += 2) {
                const double2 v = b2[(NV + c) >> 1];
                P[0][c] = fma(-l0, v.x, P[0][c]);
                P[1][c] = fma(-l1, v.x, P[1][c]);
                P[0][c + 1] = fma(-l0, v.y, P[0][c + 1]);
                P[1][c + 1] = fma(-l1, v.y, P[1][c + 1]);
                const double2 q = b2[(2 * NV + c) >> 1];
                Q[0][c] = fma(-l0, q.x, Q[0][c]);
                Q[1][c] = fma(-l1, q.x, Q[1][c]);
                Q[0][c + 1] = fma(-l0, q.y, Q[0][c + 1]);
                Q[1][c + 1] = fma(-l1, q.y, Q[1][c + 1]);
            }
        }
        if (active && lane_ok) {
            double2 *lm = reinterpret_cast<double2 *>(L.Lm + g * (int64_t)(NV * R2)) + r;
#pragma unroll
            for (int k = 0; k < NV; ++k) st_cs2(lm + k * NV, make_double2(M[0][k], M[1][k]));
            char2 pp;
            pp.x = (signed char)pv[0];
            pp.y = (signed char)pv[1];
            reinterpret_cast<char2 *>(L.piv + g * R2)[r] = pp;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (pv[s] >= 0) {
                    double *u = L.UEF + g * (int64_t)(3 * NV * NV) + pv[s];
#pragma unroll
                    for (int c = 0; c < NV; ++c) {
                        st_cs(u + c * NV, M[s][c]);
                        st_cs(u + (NV + c) * NV, P[s][c]);
                        st_cs(u + (2 * NV + c) * NV, Q[s][c]);
                    }
                }
                carry[s] = pv[s] < 0;
                if (carry[s]) {
#pragma unroll
                    for (int c = 0; c < NV; ++c) { M[s][c] = Q[s][c]; Q[s][c] = 0.0; }
                }
            }
        }
    }
    const unsigned c0 = group_bits<G>(__ballot_sync(RST_FULL_MASK, carry[0]), gw);
    const unsigned c1 = group_bits<G>(__ballot_sync(RST_FULL_MASK, carry[1]), gw);
    if (slab_ok && lane_ok) {
        const int base = __popc(c0 & below(r)) + __popc(c1 & below(r));
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (carry[s]) {
                const int t = base + ((s == 1 && carry[0]) ? 1 : 0);
                double *a = L.A_next + (slab * NV + t) * NV;
                double *b = L.B_next + (slab * NV + t) * NV;
#pragma unroll
                for (int c = 0; c < NV; ++c) { a[c] = P[s][c]; b[c] = M[s][c]; }
            }
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(128) abd_forward2_kernel(AbdLevel L) {
    using Cfg = Abd2Cfg<NV>;
    constexpr int R2 = Cfg::R2, G = Cfg::G, GPW = Cfg::GPW;
    const int lane = threadIdx.x & 31;
    const int r = lane & (G - 1);
    const int gw = lane / G;
    const int64_t slab = ((int64_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * GPW + gw;
    const bool slab_ok = slab < L.nslabs;
    const bool lane_ok = r < NV;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = slab_ok ? (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S) : 0;

    bool carry[2];
    carry[0] = slab_ok && lane_ok;
    carry[1] = false;
    double w[2];
    w[0] = carry[0] ? L.r[g0 * NV + r] : 0.0;
    w[1] = 0.0;

    for (int i = 1; i < L.S; ++i) {
        const bool active = i < len;
        const int64_t g = g0 + i;
        const bool free0 = lane_ok && !carry[0], free1 = lane_ok && !carry[1];
        const unsigned f0 = group_bits<G>(__ballot_sync(RST_FULL_MASK, free0), gw);
        const unsigned f1 = group_bits<G>(__ballot_sync(RST_FULL_MASK, free1), gw);
        int pv[2] = {-1, -1};
        double lrow[2][NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) { lrow[0][k] = 0.0; lrow[1][k] = 0.0; }
        if (active && lane_ok) {
            const int base = __popc(f0 & below(r)) + __popc(f1 & below(r));
            if (free0) w[0] = L.r[g * NV + base];
            if (free1) w[1] = L.r[g * NV + base + (free0 ? 1 : 0)];
            const char2 pp = reinterpret_cast<const char2 *>(L.piv + g * R2)[r];
            pv[0] = pp.x;
            pv[1] = pp.y;
            const double2 *lm = reinterpret_cast<const double2 *>(L.Lm + g * (int64_t)(NV * R2)) + r;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const double2 v = ld_cs2(lm + k * NV);
                lrow[0][k] = v.x;
                lrow[1][k] = v.y;
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const unsigned pb = group_bits<G>(__ballot_sync(RST_FULL_MASK, pv[0] == k || pv[1] == k), gw);
            const int p = pb ? (__ffs(pb) - 1) : 0;
            const double wsel = (pv[1] == k) ? w[1] : w[0];
            const double wp = shfl_d(wsel, p, G);
            if (pv[0] < 0 || pv[0] > k) w[0] = fma(-lrow[0][k], wp, w[0]);
            if (pv[1] < 0 || pv[1] > k) w[1] = fma(-lrow[1][k], wp, w[1]);
        }
        if (active && lane_ok) {
            if (pv[0] >= 0) L.e[g * NV + pv[0]] = w[0];
            if (pv[1] >= 0) L.e[g * NV + pv[1]] = w[1];
            carry[0] = pv[0] < 0;
            carry[1] = pv[1] < 0;
        }
    }
    const unsigned c0 = group_bits<G>(__ballot_sync(RST_FULL_MASK, carry[0]), gw);
    const unsigned c1 = group_bits<G>(__ballot_sync(RST_FULL_MASK, carry[1]), gw);
    if (slab_ok && lane_ok) {
        const int base = __popc(c0 & below(r)) + __popc(c1 & below(r));
        if (carry[0]) L.r_next[slab * NV + base] = w[0];
        if (carry[1]) L.r_next[slab * NV + base + (carry[0] ? 1 : 0)] = w[1];
    }
}

}  // namespace rst
