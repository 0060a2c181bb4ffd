// abd4.cuh -- fourth-generation factor + forward kernels of the pivoted condensation solver for block sizes
// nv = 8, 12, 16 (config C3: nv = 12).  Algorithm and the U|E|F pivot-row layout: abd.cuh.
//
// What changed against abd3.cuh (ncu, profiles/r01_b_summary.md: MIO-queue bound, 1481 warp instructions per block row,
// one LDS.128 pivot-row broadcast per two DFMA, FP64 pipe 21 % busy):
//   * the stacked 2nv x 3nv panel [M | P | Q] lives in registers in the accumulator layout of the FP64 tensor-core
//     instruction (mma.sync m8n8k4: lane = (row & 7, column pair)), all 32 lanes busy;
//   * elimination is BLOCKED by 4 pivots: the 2nv x 4 pivot-column panel is gathered to one row per lane, factored
//     with partial pivoting there (REDUX arg-max, 4 shuffles per pivot), and the rank-4 update of the remaining
//     columns is ONE sweep of DMMA m8n8k4 instructions (tools/dmma_bench.cu: DMMA and DFMA share the FP64 pipe and
//     have the same peak, 63.8 FMA/clk/SM, but one DMMA replaces 8 DFMA + their operand broadcasts);
//   * the multipliers of a block are folded with the inverse of the block's 4 x 4 unit lower triangle,
//     Lt = L * L11^-1, so the update uses the RAW pivot rows (no triangular solve on wide rows):
//         X <- X - Lt * X[pivot rows]      (for the pivot rows themselves this yields exactly their U rows)
//     and the same Lt is what the forward sweep applies to a right-hand side: 4 independent broadcasts + 4 FMA
//     per block instead of 4 dependent (ballot, shuffle, FMA) steps;
//   * the next block row is staged in shared memory with cp.async while the current one is eliminated.
//
// Factor layout written here (per block row g):
//   Lt   [b][j][rank]   block b = 0..nv/4-1, column j = 0..3, rank = rank of the slot among the 2nv - 4b slots that
//                       were not pivoted before block b            -> 4 (2nv) (nv/4) - 8 (nv/4)(nv/4 - 1) doubles
//   piv  [k]            int8: slot (panel row) that became pivot k  (16 bytes per block row)
//   UEF                 as abd.cuh (packed upper triangle of U, E, F), consumed by abd_backward_kernel unchanged
#pragma once
#include <type_traits>
#include "abd3.cuh"

namespace rst {

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

template <int NV>
struct Abd4Cfg {
    static_assert(NV % 4 == 0 && NV >= 8 && NV <= 16, "blocked DMMA kernels: nv = 8, 12, 16");
    static constexpr int R2 = 2 * NV;          // panel rows (slots)
    static constexpr int RT = R2 / 8;          // row tiles
    static constexpr int NC = 3 * NV;          // panel columns [M | P | Q]
    static constexpr int CT = (NC + 7) / 8;    // column tiles (last one may be half padding)
    static constexpr int NB = NV / 4;          // pivot blocks
    static constexpr int SH = NV / 4;          // tile distance between the M and the Q part (2 nv columns)
    static constexpr int PITCH = 8 * CT + 4;   // pivot-row buffer pitch: conflict-free B-fragment reads
    static constexpr int PIV_STRIDE = 16;      // bytes of pivot-slot record per block row
    __host__ __device__ static constexpr int lt_off(int b) { return 4 * R2 * b - 8 * b * (b - 1); }
    static constexpr int LT = 4 * R2 * NB - 8 * NB * (NB - 1);
    static constexpr unsigned ALL = R2 == 32 ? 0xffffffffu : ((1u << R2) - 1u);
};

__device__ __forceinline__ void dmma_m8n8k4(double2 &c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ unsigned lanes_below(int lane) { return (1u << lane) - 1u; }

// part (0 = M, 1 = P, 2 = Q, 3 = padding) of the column pair a lane holds in column tile TC.  nv is a multiple of 4, so a
// tile crosses at most one part boundary, in its middle: lanes with (lane & 3) < 2 hold the low half.
template <int NV, int TC>
struct Abd4Tile {
    static constexpr int LO = (8 * TC) / NV, HI = (8 * TC + 4) / NV;
    static constexpr bool SPLIT = LO != HI;
};

template <int NV, bool IDENT_B>
__global__ void __launch_bounds__(128, 4) abd_factor4_kernel(AbdLevel L, unsigned *flags) {
    using Cf = Abd4Cfg<NV>;
    using Pk = AbdPack<NV>;
    constexpr int R2 = Cf::R2, RT = Cf::RT, CT = Cf::CT, NB = Cf::NB, SH = Cf::SH, PITCH = Cf::PITCH;
    constexpr int BL = NV * NV;
    constexpr int STG = IDENT_B ? BL : 2 * BL;   // doubles staged per block row
    __shared__ __align__(16) double s_pan[4][R2 * 4];
    __shared__ __align__(16) double s_lb[4][R2 * 4];
    __shared__ __align__(16) double s_ub[4][4 * PITCH];
    __shared__ __align__(16) double s_stage[4][2][STG];
    __shared__ int s_pv[4][R2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, q = lane & 3;
    const bool hi = q >= 2;
    const int64_t slab = (int64_t)blockIdx.x * 4 + warp;
    if (slab >= L.nslabs) return;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    double *pan = s_pan[warp], *lb = s_lb[warp], *ub = s_ub[warp];
    int *pvs = s_pv[warp];

    auto stage = [&](int64_t g, int buf) {
        const char *srcA = reinterpret_cast<const char *>(L.A + g * (int64_t)BL);
        char *dst = reinterpret_cast<char *>(s_stage[warp][buf]);
        for (int c = lane; c < BL / 2; c += 32) cp_async16(dst + c * 16, srcA + c * 16);
        if constexpr (!IDENT_B) {
            const char *srcB = reinterpret_cast<const char *>(L.B + g * (int64_t)BL);
            for (int c = lane; c < BL / 2; c += 32) cp_async16(dst + BL * 8 + c * 16, srcB + c * 16);
        }
        cp_async_commit();
    };
    if (len > 1) stage(g0 + 1, 1);

    // panel in accumulator layout: X[T][TC] = (row 8T + gid, columns 8TC + 2q, 8TC + 2q + 1)
    double2 X[RT][CT];
    // carry rows of the slab start: slots 0..nv-1 = block row g0 as [ . | P = A | Q = B ]
    unsigned carrymask = (1u << NV) - 1u;
#pragma unroll
    for (int T = 0; T < RT; ++T) {
        const int slot = 8 * T + gid;
        static_for<0, CT>([&](auto tc) {
            constexpr int TC = decltype(tc)::value;
            using Tl = Abd4Tile<NV, TC>;
            const int part = (Tl::SPLIT && hi) ? Tl::HI : Tl::LO;
            const int idx = 8 * TC + 2 * q - part * NV;
            double2 v = make_double2(0.0, 0.0);
            if (slot < NV) {
                if (part == 1) v = *reinterpret_cast<const double2 *>(L.A + (g0 * NV + slot) * NV + idx);
                if (part == 2) {
                    if constexpr (IDENT_B) v = make_double2(idx == slot ? 1.0 : 0.0, idx + 1 == slot ? 1.0 : 0.0);
                    else v = *reinterpret_cast<const double2 *>(L.B + (g0 * NV + slot) * NV + idx);
                }
            }
            X[T][TC] = v;
        });
    }

    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        cp_async_wait_all();
        __syncwarp();
        const double *stA = s_stage[warp][i & 1];
        // ---- carry rows: M <- Q, Q <- 0; free slots: next block row [M = A | 0 | Q = B] ---------------------------
        const unsigned freemask = ~carrymask & Cf::ALL;
#pragma unroll
        for (int T = 0; T < RT; ++T) {
            const int slot = 8 * T + gid;
            const bool fr = (freemask >> slot) & 1u;
            const int t = __popc(freemask & lanes_below(slot));
            static_for<0, CT>([&](auto tc) {
                constexpr int TC = decltype(tc)::value;
                using Tl = Abd4Tile<NV, TC>;
                const int part = (Tl::SPLIT && hi) ? Tl::HI : Tl::LO;
                const int idx = 8 * TC + 2 * q - part * NV;
                double2 nv_ = make_double2(0.0, 0.0);   // value of a new row
                double2 cv = make_double2(0.0, 0.0);    // value of a carry row
                auto for_part = [&](int pt) {
                    if (pt == 0) {
                        if (fr) nv_ = *reinterpret_cast<const double2 *>(stA + t * NV + idx);
                        if constexpr (TC + SH < CT) cv = X[T][TC + SH];
                    } else if (pt == 1) {
                        cv = X[T][TC];
                    } else if (pt == 2) {
                        if constexpr (IDENT_B) nv_ = make_double2(idx == t ? 1.0 : 0.0, idx + 1 == t ? 1.0 : 0.0);
                        else if (fr) nv_ = *reinterpret_cast<const double2 *>(stA + BL + t * NV + idx);
                    }
                };
                if constexpr (Tl::SPLIT) {
                    if (!hi) for_part(Tl::LO); else for_part(Tl::HI);
                } else {
                    for_part(Tl::LO);
                }
                X[T][TC] = fr ? nv_ : cv;
            });
        }
        __syncwarp();                       // every lane has read its staged rows: the other buffer may be refilled
        if (i + 1 < len) stage(g + 1, (i + 1) & 1);

        // ---- blocked elimination of the nv columns of M ------------------------------------------------------
        unsigned cand_mask = Cf::ALL;       // slots not pivoted yet in this block row
        int pv = -1;                        // row layout (lane = slot): pivot index of my slot
        unsigned pword[NB];                 // pivot slots, 4 per word
        double *lt_g = L.Lm + g * (int64_t)Cf::LT;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int TCB = (4 * b) / 8;          // column tile of the block's columns
            const bool blk_hi = ((4 * b) % 8) != 0;
            // (a) gather the 2nv x 4 pivot-column panel: one row per lane
#pragma unroll
            for (int T = 0; T < RT; ++T)
                if (hi == blk_hi) *reinterpret_cast<double2 *>(pan + (8 * T + gid) * 4 + 2 * (q & 1)) = X[T][TCB];
            __syncwarp();
            double a[4];
            {
                const int row = lane < R2 ? lane : 0;
                const double2 v0 = *reinterpret_cast<const double2 *>(pan + row * 4);
                const double2 v1 = *reinterpret_cast<const double2 *>(pan + row * 4 + 2);
                a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
            }
            // (b) 4 pivots with partial pivoting over the candidate slots
            const unsigned cand0 = cand_mask;
            bool cand = (lane < R2) && ((cand_mask >> lane) & 1u);
            const bool cand_at_start = cand;
            int mypiv = -1;                 // 0..3: my slot became pivot `mypiv` of this block
            int ps[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int p = group_argmax_redux<32>(a[ii], cand, 0, RST_FULL_MASK);
                ps[ii] = p;
                double ap[4];
#pragma unroll
                for (int j = ii; j < 4; ++j) ap[j] = __shfl_sync(RST_FULL_MASK, a[j], p);
                if (!(fabs(ap[ii]) > 0.0) && lane == 0) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
                const double rinv = __drcp_rn(ap[ii]);
                const bool is_piv = lane == p;
                if (cand && !is_piv) {
                    const double l = a[ii] * rinv;
                    a[ii] = l;
#pragma unroll
                    for (int j = ii + 1; j < 4; ++j) a[j] = fma(-l, ap[j], a[j]);
                }
                if (is_piv) { mypiv = ii; pv = 4 * b + ii; cand = false; }
                cand_mask &= ~(1u << p);
            }
            pword[b] = (unsigned)ps[0] | ((unsigned)ps[1] << 8) | ((unsigned)ps[2] << 16) | ((unsigned)ps[3] << 24);
            // multipliers of my slot: candidates keep all four, pivot ii keeps those of the earlier pivots, others none
            double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
            if (cand_at_start) {
                const int lim = mypiv < 0 ? 4 : mypiv;
                l0 = lim > 0 ? a[0] : 0.0; l1 = lim > 1 ? a[1] : 0.0; l2 = lim > 2 ? a[2] : 0.0; l3 = lim > 3 ? a[3] : 0.0;
            }
            // (c) inverse of the block's unit lower triangle (strict part n_ij = multiplier of pivot row i at pivot j)
            const double n10 = __shfl_sync(RST_FULL_MASK, l0, ps[1]);
            const double n20 = __shfl_sync(RST_FULL_MASK, l0, ps[2]);
            const double n21 = __shfl_sync(RST_FULL_MASK, l1, ps[2]);
            const double n30 = __shfl_sync(RST_FULL_MASK, l0, ps[3]);
            const double n31 = __shfl_sync(RST_FULL_MASK, l1, ps[3]);
            const double n32 = __shfl_sync(RST_FULL_MASK, l2, ps[3]);
            const double i10 = -n10;
            const double i21 = -n21, i20 = -fma(n21, i10, n20);
            const double i32 = -n32, i31 = -fma(n32, i21, n31), i30 = -fma(n32, i20, fma(n31, i10, n30));
            // Lt = L * L11^-1
            const double t3 = l3;
            const double t2 = fma(l3, i32, l2);
            const double t1 = fma(l3, i31, fma(l2, i21, l1));
            const double t0 = fma(l3, i30, fma(l2, i20, fma(l1, i10, l0)));
            // (d) Lt: shared memory (A fragments) + HBM (forward sweep)
            if (lane < R2) {
                *reinterpret_cast<double2 *>(lb + lane * 4) = make_double2(-t0, -t1);
                *reinterpret_cast<double2 *>(lb + lane * 4 + 2) = make_double2(-t2, -t3);
                if (b == NB - 1) pvs[lane] = pv;
            }
            if (cand_at_start) {
                const int ncand = R2 - 4 * b;
                double *dst = lt_g + Cf::lt_off(b) + __popc(cand0 & lanes_below(lane));
                st_cs(dst, t0); st_cs(dst + ncand, t1); st_cs(dst + 2 * ncand, t2); st_cs(dst + 3 * ncand, t3);
            }
            // (e) raw pivot rows -> shared memory (B fragments)
#pragma unroll
            for (int T = 0; T < RT; ++T) {
                const int slot = 8 * T + gid;
                int which = -1;
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) which = (slot == ps[ii]) ? ii : which;
                if (which >= 0) {
#pragma unroll
                    for (int TC = TCB; TC < CT; ++TC)
                        *reinterpret_cast<double2 *>(ub + which * PITCH + 8 * TC + 2 * q) = X[T][TC];
                }
            }
            __syncwarp();
            // (f) + (g): X <- X - Lt * X[pivot rows] on the FP64 tensor cores
            double af[RT];
#pragma unroll
            for (int T = 0; T < RT; ++T) af[T] = lb[32 * T + lane];
#pragma unroll
            for (int TC = TCB; TC < CT; ++TC) {
                const double bf = ub[q * PITCH + 8 * TC + gid];
#pragma unroll
                for (int T = 0; T < RT; ++T) dmma_m8n8k4(X[T][TC], af[T], bf);
            }
        }
        __syncwarp();   // pvs visible
        // ---- pivot rows -> U | E | F (layout of abd.cuh), pivot slots -> piv --------------------------------------
        if (lane < NB) {
            unsigned w = pword[0];
#pragma unroll
            for (int b = 1; b < NB; ++b) w = (lane == b) ? pword[b] : w;
            reinterpret_cast<unsigned *>(L.piv + g * (int64_t)Cf::PIV_STRIDE)[lane] = w;
        }
        double *uef = L.UEF + g * (int64_t)Pk::UEF;
#pragma unroll
        for (int T = 0; T < RT; ++T) {
            const int k = pvs[8 * T + gid];
            if (k >= 0) {
                static_for<0, CT>([&](auto tc) {
                    constexpr int TC = decltype(tc)::value;
                    using Tl = Abd4Tile<NV, TC>;
                    const int part = (Tl::SPLIT && hi) ? Tl::HI : Tl::LO;
                    const int idx = 8 * TC + 2 * q - part * NV;
                    const double2 v = X[T][TC];
                    if (part == 0) {
                        if (idx >= k) st_cs(uef + Pk::u_off(idx) + k, v.x);
                        if (idx + 1 >= k) st_cs(uef + Pk::u_off(idx + 1) + k, v.y);
                    } else if (part == 1) {
                        st_cs(uef + Pk::TRI + idx * NV + k, v.x);
                        st_cs(uef + Pk::TRI + (idx + 1) * NV + k, v.y);
                    } else if (part == 2) {
                        st_cs(uef + Pk::TRI + BL + idx * NV + k, v.x);
                        st_cs(uef + Pk::TRI + BL + (idx + 1) * NV + k, v.y);
                    }
                });
            }
        }
        carrymask = cand_mask;
    }
    // ---- condensed row of the slab: A_next = P part, B_next = Q part of the surviving slots -------------------------
#pragma unroll
    for (int T = 0; T < RT; ++T) {
        const int slot = 8 * T + gid;
        if ((carrymask >> slot) & 1u) {
            const int t = __popc(carrymask & lanes_below(slot));
            static_for<0, CT>([&](auto tc) {
                constexpr int TC = decltype(tc)::value;
                using Tl = Abd4Tile<NV, TC>;
                const int part = (Tl::SPLIT && hi) ? Tl::HI : Tl::LO;
                const int idx = 8 * TC + 2 * q - part * NV;
                if (part == 1) *reinterpret_cast<double2 *>(L.A_next + (slab * NV + t) * NV + idx) = X[T][TC];
                if (part == 2) *reinterpret_cast<double2 *>(L.B_next + (slab * NV + t) * NV + idx) = X[T][TC];
            });
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward sweep for the blocked factors: lane = slot, w = right-hand side entry of the slot.  Per block row: 4-byte
// pivot records (warp-uniform), per block 4 coalesced Lt loads, 4 independent broadcasts of the raw pivot entries, 4 FMA.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(128) abd_forward4_kernel(AbdLevel L) {
    using Cf = Abd4Cfg<NV>;
    constexpr int R2 = Cf::R2, NB = Cf::NB;
    const int lane = threadIdx.x & 31;
    const int64_t slab = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (slab >= L.nslabs) return;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    const unsigned below = lanes_below(lane);
    unsigned carrymask = (1u << NV) - 1u;
    double w = lane < NV ? L.r[g0 * NV + lane] : 0.0;
    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        const double *lt_g = L.Lm + g * (int64_t)Cf::LT;
        if (L.prefetch && i + 1 < len && lane * 128 < Cf::LT * 8)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(lt_g + Cf::LT) + lane * 128));
        const unsigned *pw = reinterpret_cast<const unsigned *>(L.piv + g * (int64_t)Cf::PIV_STRIDE);
        unsigned pword[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) pword[b] = __ldg(pw + b);
        const unsigned freemask = ~carrymask & Cf::ALL;
        if ((freemask >> lane) & 1u) w = L.r[g * NV + __popc(freemask & below)];
        // all multiplier loads of the block row up front (independent of the chain)
        unsigned cm[NB + 1];
        cm[0] = Cf::ALL;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            unsigned m = cm[b];
#pragma unroll
            for (int j = 0; j < 4; ++j) m &= ~(1u << ((pword[b] >> (8 * j)) & 31u));
            cm[b + 1] = m;
        }
        double lt[NB][4];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const bool c = (cm[b] >> lane) & 1u;
            const int ncand = R2 - 4 * b;
            const double *src = lt_g + Cf::lt_off(b) + __popc(cm[b] & below);
#pragma unroll
            for (int j = 0; j < 4; ++j) lt[b][j] = c ? ld_cs(src + j * ncand) : 0.0;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            double wp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wp[j] = __shfl_sync(RST_FULL_MASK, w, (int)((pword[b] >> (8 * j)) & 31u));
#pragma unroll
            for (int j = 0; j < 4; ++j) w = fma(-lt[b][j], wp[j], w);
        }
        // transformed right-hand side of the pivot rows, in pivot order
        {
            unsigned wsel = pword[0];
#pragma unroll
            for (int b = 1; b < NB; ++b) wsel = ((lane >> 2) == b) ? pword[b] : wsel;
            const int src = (int)((wsel >> (8 * (lane & 3))) & 31u);
            const double e = __shfl_sync(RST_FULL_MASK, w, src);
            if (lane < NV) L.e[g * NV + lane] = e;
        }
        carrymask = cm[NB];
    }
    if ((carrymask >> lane) & 1u) L.r_next[slab * NV + __popc(carrymask & below)] = w;
}

}  // namespace rst
