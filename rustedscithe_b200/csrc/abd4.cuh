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
    static constexpr int PITCH = 8 * CT + 4;   // pitch of the 4 raw pivot rows of a block: conflict-free B-fragment reads
    static constexpr int PITCH_W = 8 * CT + 2; // pitch of the nv final pivot rows: the transposing U|E|F write-out reads columns
    static constexpr int PIV_STRIDE = 16;      // bytes of pivot-slot record per block row
    __host__ __device__ static constexpr int lt_off(int b) { return 4 * R2 * b - 8 * b * (b - 1); }
    static constexpr int LT = 4 * R2 * NB - 8 * NB * (NB - 1);
    static constexpr unsigned ALL = R2 == 32 ? 0xffffffffu : ((1u << R2) - 1u);
    // nv = 12: the last column tile is half padding (columns 36..39).  Column NC can carry the right-hand side through the
    // elimination for free (abd_factor4_kernel<.., RHS = true>): the factorisation then delivers what the forward sweep would.
    static constexpr bool HAS_RHS_COL = (NC % 8) == 4;
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
// predicated 16-byte shared-memory store (no branch, no reconvergence point)
__device__ __forceinline__ void sts128_if(bool pred, double *p, double2 v) {
    asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %3, 0;\n\t@pp st.shared.v2.f64 [%0], {%1, %2};\n\t}"
                 ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v.x), "d"(v.y), "r"((int)pred) : "memory");
}
// part (0 = M, 1 = P, 2 = Q, 3 = padding) of the column pair a lane holds in column tile TC.  nv is a multiple of 4, so a
// tile crosses at most one part boundary, in its middle: lanes with (lane & 3) < 2 hold the low half.
template <int NV, int TC>
struct Abd4Tile {
    static constexpr int LO = (8 * TC) / NV, HI = (8 * TC + 4) / NV;
    static constexpr bool SPLIT = LO != HI;
};

// shared memory of one warp of the factor kernel (dynamic: 4 warps need 46..90 KB)
template <int NV, bool IDENT_B>
struct Abd4Smem {
    using Cf = Abd4Cfg<NV>;
    static constexpr int STG_RHS = (IDENT_B ? 1 : 2) * NV * NV;   // where the staged right-hand side of the block row starts
    static constexpr int STG = STG_RHS + (Cf::HAS_RHS_COL ? ((NV + 1) / 2) * 2 : 0);   // doubles staged per block row
    double pan[4 * Cf::R2];             // pivot-column panel [column plane][slot]; planes ordered 0, 2, 1, 3 (bank spread)
    double lb[4 * Cf::R2 + 8];          // -Lt of the block (A fragments): columns (0,1) of every slot, then (+8: other banks) columns (2,3)
    double pr[4 * 4];                   // pivot row of each of the block's 4 pivots (its own entry replaced by 1/pivot)
    double xs[NV * Cf::PITCH_W];        // 4 raw pivot rows of a block (B fragments) / nv final pivot rows (U|E|F write-out)
    double stage[2][STG];               // next block row(s), filled by cp.async
    int pv[Cf::R2];                     // pivot index of each slot in the current block row
};

#ifndef ABD4_MIN_BLOCKS
#define ABD4_MIN_BLOCKS 4
#endif
// RHS = true (nv = 12 only): the right-hand side L.r rides along as panel column 3 nv, so that the factorisation also
// produces the transformed right-hand side of the pivot rows (L.e) and of the condensed row (L.r_next) -- exactly what
// abd_forward4_kernel computes from the stored Lt -- and the first linear solve after a refactorisation needs no forward sweep.
template <int NV, bool IDENT_B, bool RHS = false>
__global__ void __launch_bounds__(128, ABD4_MIN_BLOCKS) abd_factor4_kernel(AbdLevel L, unsigned *flags) {
    using Cf = Abd4Cfg<NV>;
    static_assert(!RHS || Cf::HAS_RHS_COL, "the right-hand side column needs a half-padded last tile");
    using Pk = AbdPack<NV>;
    using Sm = Abd4Smem<NV, IDENT_B>;
    constexpr int R2 = Cf::R2, RT = Cf::RT, CT = Cf::CT, NB = Cf::NB, SH = Cf::SH, PITCH = Cf::PITCH, PITCH_W = Cf::PITCH_W;
    constexpr int BL = NV * NV;
    constexpr int NM = (Pk::UEF + 31) / 32;   // U|E|F doubles written per lane and block row
    extern __shared__ __align__(16) unsigned char abd4_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = lane >> 2, q = lane & 3;
    const bool hi = q >= 2;
    const int64_t slab = (int64_t)blockIdx.x * 4 + warp;
    if (slab >= L.nslabs) return;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    Sm &sm = reinterpret_cast<Sm *>(abd4_smem_raw)[warp];
    double *pan = sm.pan, *lb = sm.lb, *xs = sm.xs, *pr = sm.pr;
    int *pvs = sm.pv;

    auto stage = [&](int64_t g, int buf) {
        const char *srcA = reinterpret_cast<const char *>(L.A + g * (int64_t)BL);
        char *dst = reinterpret_cast<char *>(sm.stage[buf]);
        for (int c = lane; c < BL / 2; c += 32) cp_async16(dst + c * 16, srcA + c * 16);
        if constexpr (!IDENT_B) {
            const char *srcB = reinterpret_cast<const char *>(L.B + g * (int64_t)BL);
            for (int c = lane; c < BL / 2; c += 32) cp_async16(dst + BL * 8 + c * 16, srcB + c * 16);
        }
#ifndef ABD4_X_NORSTAGE
        if constexpr (RHS) {
            if (lane < NV / 2) cp_async16(dst + Sm::STG_RHS * 8 + lane * 16, reinterpret_cast<const char *>(L.r + g * NV) + lane * 16);
        }
#endif
        cp_async_commit();
    };
    if (len > 1) stage(g0 + 1, 1);

    // U|E|F write-out map: output double i = lane + 32 m of a block row comes from pivot row k(i), panel column col(i);
    // byte offsets into xs (rows by pivot index), two per register
    unsigned wmap[(NM + 1) / 2];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const int i = lane + 32 * m;
        int k = 0, col = 0;
        if (i < Pk::TRI) {
            int c = 0;
            while ((c + 1) * (c + 2) / 2 <= i) ++c;
            k = i - c * (c + 1) / 2;
            col = c;
        } else if (i < Pk::UEF) {
            const int j = i - Pk::TRI;
            k = j % NV;
            col = NV + j / NV;
        }
        const unsigned off = (unsigned)((k * PITCH_W + col) * 8);
        if (m % 2 == 0) wmap[m / 2] = off; else wmap[m / 2] |= off << 16;
    }

    // panel in accumulator layout: X[T][TC] = (row 8T + gid, columns 8TC + 2q, 8TC + 2q + 1)
    double2 X[RT][CT];
    // carry rows of the slab start: slots 0..nv-1 = block row g0 as [ . | P = A | Q = B ]
    unsigned carrymask = (1u << NV) - 1u;
    unsigned bad = 0u;
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
                if constexpr (RHS) {
                    if (part == 3 && q == 2) v = make_double2(L.r[g0 * NV + slot], 0.0);
                }
            }
            X[T][TC] = v;
        });
    }

    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        cp_async_wait_all();
        __syncwarp();
        const double *stA = sm.stage[i & 1];
        // ---- carry rows: M <- Q, Q <- 0; free slots: next block row [M = A | 0 | Q = B] ---------------------------
        const unsigned freemask = ~carrymask & Cf::ALL;
#pragma unroll
        for (int T = 0; T < RT; ++T) {
            const int slot = 8 * T + gid;
            const bool fr = (freemask >> slot) & 1u;
            const int t = __popc(freemask & lanes_below(slot));
            if (fr) {
                static_for<0, CT>([&](auto tc) {
                    constexpr int TC = decltype(tc)::value;
                    using Tl = Abd4Tile<NV, TC>;
                    const int part = (Tl::SPLIT && hi) ? Tl::HI : Tl::LO;
                    const int idx = 8 * TC + 2 * q - part * NV;
                    double2 v = make_double2(0.0, 0.0);
                    if (Tl::LO == 0 || Tl::HI == 0) {
                        if (part == 0) v = *reinterpret_cast<const double2 *>(stA + t * NV + idx);
                    }
                    if (Tl::LO == 2 || Tl::HI == 2) {
                        if (part == 2) {
                            if constexpr (IDENT_B) v = make_double2(idx == t ? 1.0 : 0.0, idx + 1 == t ? 1.0 : 0.0);
                            else v = *reinterpret_cast<const double2 *>(stA + BL + t * NV + idx);
                        }
                    }
                    if constexpr (RHS) {
                        if (part == 3 && q == 2) v = make_double2(stA[Sm::STG_RHS + t], 0.0);
                    }
                    X[T][TC] = v;
                });
            } else {
                static_for<0, CT>([&](auto tc) {
                    constexpr int TC = decltype(tc)::value;
                    using Tl = Abd4Tile<NV, TC>;
                    if constexpr (!Tl::SPLIT) {
                        if constexpr (Tl::LO == 0) X[T][TC] = X[T][TC + SH];
                        else if constexpr (Tl::LO != 1) X[T][TC] = make_double2(0.0, 0.0);
                    } else {
                        // low half / high half of the tile belong to different parts
                        double2 lo_v = X[T][TC], hi_v = X[T][TC];
                        if constexpr (Tl::LO == 0) { if constexpr (TC + SH < CT) lo_v = X[T][TC + SH]; }
                        else if constexpr (Tl::LO != 1) lo_v = make_double2(0.0, 0.0);
                        if constexpr (Tl::HI != 1 && !(RHS && Tl::HI == 3)) hi_v = make_double2(0.0, 0.0);   // RHS: column 3 nv stays
                        X[T][TC] = hi ? hi_v : lo_v;
                    }
                });
            }
        }
        __syncwarp();                       // every lane has read its staged rows: the other buffer may be refilled
        if (i + 1 < len) stage(g + 1, (i + 1) & 1);

        // ---- blocked elimination of the nv columns of M ------------------------------------------------------
        unsigned cand_mask = Cf::ALL;       // slots not pivoted yet in this block row
        int pv = -1;                        // row layout (lane = slot): pivot index of my slot
        unsigned mypword = 0u;              // lane b: the pivot slots of block b, 4 per word
        double *lt_g = L.Lm + g * (int64_t)Cf::LT;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int TCB = (4 * b) / 8;          // column tile of the block's columns
            const bool blk_hi = ((4 * b) % 8) != 0;
            // (a) gather the 2nv x 4 pivot-column panel: one row per lane
#pragma unroll
            for (int T = 0; T < RT; ++T) {
                if (hi == blk_hi) {
                    pan[(q & 1) * R2 + 8 * T + gid] = X[T][TCB].x;          // column 2 (q & 1)     -> plane (q & 1)
                    pan[(2 + (q & 1)) * R2 + 8 * T + gid] = X[T][TCB].y;    // column 2 (q & 1) + 1 -> plane 2 + (q & 1)
                }
            }
            __syncwarp();
            double a[4];
            {
                const int row = lane < R2 ? lane : 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) a[j] = pan[((j >> 1) | ((j & 1) << 1)) * R2 + row];
            }
            // (b) 4 pivots with partial pivoting over the candidate slots
            const unsigned cand0 = cand_mask;
            bool cand = (lane < R2) && ((cand_mask >> lane) & 1u);
            const bool cand_at_start = cand;
            int mypiv = 4;                  // 0..3: my slot became pivot `mypiv` of this block
            int ps[4];
#ifdef ABD4_X_DIVFREE
            // what-if (tools/abd4_probe.cu), measured SLOWER here (2.35 vs 1.98 ms: this kernel is bound by instruction and
            // shared-memory-pipe throughput, and the exponent arithmetic + 20 shuffles cost more than the shorter FP64 chain saves).
            // DIVISION-FREE elimination inside the block (as bigd.cuh): live rows are updated as
            //   a'[c] = (pi s) a[c] - (a[ii] s) ap[c]
            // with pi the pivot and s the power of two that brings pi into [1, 2), applied through the exponent field on the
            // integer pipe: the per-pivot chain has 2 dependent FP64 operations instead of 6 (reciprocal, multiplier, update),
            // which matters because co-resident warps stream DMMAs (profiles/r02_coissue.md: a dependent DFMA next to a DMMA stream
            // takes 61 cycles, not 8).  Every live row carries the same factor, so the arg-max is unchanged; the multipliers
            // l = a[j] / piv[j] are scale free and are formed after the four pivots, reciprocals and L11 entries once per block,
            // spread over 10 lanes.
            double piv[4], num[6];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const unsigned hw = (unsigned)__double2hiint(a[ii]) & 0x7fffffffu;
                const unsigned key = cand ? (((hw & ~31u) | (unsigned)(31 - lane)) + 1u) : 0u;
                const unsigned mx = __reduce_max_sync(RST_FULL_MASK, key);
                const int p = 31 - (int)((mx - 1u) & 31u);
                ps[ii] = p;
                if (lane == p) {
                    *reinterpret_cast<double2 *>(pr + 4 * ii) = make_double2(a[0], a[1]);
                    *reinterpret_cast<double2 *>(pr + 4 * ii + 2) = make_double2(a[2], a[3]);
                }
                __syncwarp();
                double ap[4];
                {
                    const double2 r0 = *reinterpret_cast<const double2 *>(pr + 4 * ii);
                    const double2 r1 = *reinterpret_cast<const double2 *>(pr + 4 * ii + 2);
                    ap[0] = r0.x; ap[1] = r0.y; ap[2] = r1.x; ap[3] = r1.y;
                }
                piv[ii] = ap[ii];
                if (ii == 1) { num[0] = ap[0]; }
                if (ii == 2) { num[1] = ap[0]; num[2] = ap[1]; }
                if (ii == 3) { num[3] = ap[0]; num[4] = ap[1]; num[5] = ap[2]; }
                bad |= (mx <= 32u || mx > 0x7ff00000u) ? 1u : 0u;   // zero / denormal or non-finite pivot
                const bool is_piv = lane == p;
                const bool upd = cand && !is_piv;
                if (ii < 3) {
                    const int phi = __double2hiint(ap[ii]);
                    const int ex = (phi >> 20) & 0x7ff;
                    const int d = ex != 0 ? 1023 - ex : 0;
                    const double pis = __hiloint2double(ex != 0 ? ((phi & 0x800fffff) | 0x3ff00000) : phi, __double2loint(ap[ii]));
                    const int ahi = __double2hiint(a[ii]);
                    const int ae = (ahi >> 20) & 0x7ff;
                    const bool ok = ae != 0 && ae + d > 0;
                    const double ais = __hiloint2double(ok ? ahi + d * (1 << 20) : (ahi & (int)0x80000000), ok ? __double2loint(a[ii]) : 0);
                    if (upd) {
#pragma unroll
                        for (int j = ii + 1; j < 4; ++j) a[j] = fma(-ais, ap[j], pis * a[j]);
                    }
                }
                if (is_piv) { mypiv = ii; pv = 4 * b + ii; cand = false; }
                cand_mask &= ~(1u << p);
            }
            if (lane == b) mypword = (unsigned)ps[0] | ((unsigned)ps[1] << 8) | ((unsigned)ps[2] << 16) | ((unsigned)ps[3] << 24);
            // warp-uniform values, one per lane: lanes 0..5  n_k = num[k] / piv[j(k)]  (j = 0 0 1 0 1 2), lanes 6..9  1 / piv[lane - 6]
            double n10, n20, n21, n30, n31, n32, rinv[4];
            {
                const int jk = (int)((0x3210210100ull >> (4 * (lane < 10 ? lane : 0))) & 3ull);
                double pj = piv[0];
                pj = jk == 1 ? piv[1] : pj; pj = jk == 2 ? piv[2] : pj; pj = jk == 3 ? piv[3] : pj;
                double nk = 1.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) nk = lane == k ? num[k] : nk;
                const double v = fast_rcp(pj) * nk;
                n10 = __shfl_sync(RST_FULL_MASK, v, 0); n20 = __shfl_sync(RST_FULL_MASK, v, 1); n21 = __shfl_sync(RST_FULL_MASK, v, 2);
                n30 = __shfl_sync(RST_FULL_MASK, v, 3); n31 = __shfl_sync(RST_FULL_MASK, v, 4); n32 = __shfl_sync(RST_FULL_MASK, v, 5);
#pragma unroll
                for (int j = 0; j < 4; ++j) rinv[j] = __shfl_sync(RST_FULL_MASK, v, 6 + j);
            }
            // multipliers of my slot: candidates keep all four, pivot ii keeps those of the earlier pivots, others none
            const int lim = cand_at_start ? mypiv : 0;
            const double l0 = lim > 0 ? a[0] * rinv[0] : 0.0, l1 = lim > 1 ? a[1] * rinv[1] : 0.0, l2 = lim > 2 ? a[2] * rinv[2] : 0.0,
                         l3 = lim > 3 ? a[3] * rinv[3] : 0.0;
            // (c) inverse of the block's unit lower triangle: i10 = -n10, i21 = -n21, i32 = -n32 through operand negation
            const double i20 = fma(n21, n10, -n20);
            const double i31 = fma(n32, n21, -n31);
            const double i30 = fma(-n32, i20, fma(n31, n10, -n30));
            // Lt = L * L11^-1
            const double t3 = l3;
            const double t2 = fma(-l3, n32, l2);
            const double t1 = fma(l3, i31, fma(-l2, n21, l1));
            const double t0 = fma(l3, i30, fma(l2, i20, fma(-l1, n10, l0)));
#else
            double n10 = 0.0, n20 = 0.0, n21 = 0.0, n30 = 0.0, n31 = 0.0, n32 = 0.0;   // strict lower triangle of the block
#ifdef ABD4_X_SHORTPANEL
            ps[1] = 1; ps[2] = 2; ps[3] = 3;
#pragma unroll
            for (int ii = 0; ii < 1; ++ii) {
#else
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
#endif
                // arg-max of |a[ii]| over the candidates: ONE 32-bit REDUX on (high word of |a| with its low 5 bits replaced
                // by 31 - lane): ties and values equal in their leading 15 mantissa bits resolve to the lowest slot
                const unsigned hw = (unsigned)__double2hiint(a[ii]) & 0x7fffffffu;
                const unsigned key = cand ? (((hw & ~31u) | (unsigned)(31 - lane)) + 1u) : 0u;
                // (the compiler predicates the Newton steps on the pivot lane and issues them after the REDUX; forcing them ahead
                // of it with a volatile block measures the same: this kernel is throughput bound, not chain bound)
                const double myr = fast_rcp(a[ii]);
                const unsigned mx = __reduce_max_sync(RST_FULL_MASK, key);
                const int p = 31 - (int)((mx - 1u) & 31u);
                ps[ii] = p;
                // the pivot lane publishes its row (multipliers of the earlier pivots | 1/pivot | entries to the right) through
                // shared memory: 4 wavefronts instead of 8-10 shuffles per pivot, and the triangle of (c) comes for free
                if (lane == p) {
                    *reinterpret_cast<double2 *>(pr + 4 * ii) = make_double2(ii == 0 ? myr : a[0], ii == 1 ? myr : a[1]);
                    *reinterpret_cast<double2 *>(pr + 4 * ii + 2) = make_double2(ii == 2 ? myr : a[2], ii == 3 ? myr : a[3]);
                }
                __syncwarp();
                double ap[4];
                {
                    const double2 r0 = *reinterpret_cast<const double2 *>(pr + 4 * ii);
                    const double2 r1 = *reinterpret_cast<const double2 *>(pr + 4 * ii + 2);
                    ap[0] = r0.x; ap[1] = r0.y; ap[2] = r1.x; ap[3] = r1.y;
                }
                const double rinv = ap[ii];
                if (ii == 1) { n10 = ap[0]; }
                if (ii == 2) { n20 = ap[0]; n21 = ap[1]; }
                if (ii == 3) { n30 = ap[0]; n31 = ap[1]; n32 = ap[2]; }
                bad |= (mx <= 32u || mx > 0x7ff00000u) ? 1u : 0u;   // zero / denormal or non-finite pivot
                const bool is_piv = lane == p;
                const bool upd = cand && !is_piv;
                const double l = upd ? a[ii] * rinv : 0.0;
                if (upd) a[ii] = l;
#pragma unroll
                for (int j = ii + 1; j < 4; ++j) a[j] = fma(-l, ap[j], a[j]);
                if (is_piv) { mypiv = ii; pv = 4 * b + ii; cand = false; }
                cand_mask &= ~(1u << p);
            }
            if (lane == b) mypword = (unsigned)ps[0] | ((unsigned)ps[1] << 8) | ((unsigned)ps[2] << 16) | ((unsigned)ps[3] << 24);
            // multipliers of my slot: candidates keep all four, pivot ii keeps those of the earlier pivots, others none
            const int lim = cand_at_start ? mypiv : 0;
            const double l0 = lim > 0 ? a[0] : 0.0, l1 = lim > 1 ? a[1] : 0.0, l2 = lim > 2 ? a[2] : 0.0, l3 = lim > 3 ? a[3] : 0.0;
            // (c) inverse of the block's unit lower triangle (strict part n_ij = multiplier of pivot row i at pivot j)
            const double i10 = -n10;
            const double i21 = -n21, i20 = -fma(n21, i10, n20);
            const double i32 = -n32, i31 = -fma(n32, i21, n31), i30 = -fma(n32, i20, fma(n31, i10, n30));
            // Lt = L * L11^-1
            const double t3 = l3;
            const double t2 = fma(l3, i32, l2);
            const double t1 = fma(l3, i31, fma(l2, i21, l1));
            const double t0 = fma(l3, i30, fma(l2, i20, fma(l1, i10, l0)));
#endif
            // (d) Lt: shared memory (A fragments) + HBM (forward sweep)
            if (lane < R2) {
                *reinterpret_cast<double2 *>(lb + lane * 2) = make_double2(-t0, -t1);
                *reinterpret_cast<double2 *>(lb + 2 * R2 + 8 + lane * 2) = make_double2(-t2, -t3);
                if (b == NB - 1) pvs[lane] = pv;
            }
#ifndef ABD4_X_NOLTSTORE
            if (cand_at_start) {
                const int ncand = R2 - 4 * b;
                double *dst = lt_g + Cf::lt_off(b) + __popc(cand0 & lanes_below(lane));
                st_cs(dst, t0); st_cs(dst + ncand, t1); st_cs(dst + 2 * ncand, t2); st_cs(dst + 3 * ncand, t3);
            }
#endif
            // (e) raw rows of the block's 4 pivots -> shared memory (B fragments); which pivot my slot is comes from its lane
#pragma unroll
            for (int T = 0; T < RT; ++T) {
                const int which = __shfl_sync(RST_FULL_MASK, mypiv, 8 * T + gid);
#pragma unroll
#ifndef ABD4_X_NOPOST
                for (int TC = TCB; TC < CT; ++TC) sts128_if(which < 4, xs + which * PITCH + 8 * TC + 2 * q, X[T][TC]);
#endif
            }
            __syncwarp();
            // (f): X <- X - Lt * X[pivot rows] on the FP64 tensor cores
            double af[RT];
#pragma unroll
            for (int T = 0; T < RT; ++T) af[T] = lb[(q >> 1) * (2 * R2 + 8) + (8 * T + gid) * 2 + (q & 1)];
#pragma unroll
            for (int TC = TCB; TC < CT; ++TC) {
                const double bf = xs[q * PITCH + 8 * TC + gid];
#ifndef ABD4_X_NODMMA   // what-if switch of tools/abd4_probe.cu
#pragma unroll
                for (int T = 0; T < RT; ++T) dmma_m8n8k4(X[T][TC], af[T], bf);
#else
                X[0][TC].x += bf * af[0] == 123.0 ? 1.0 : 0.0;
#endif
            }
            __syncwarp();   // xs / lb / pan are rewritten by the next block (or the write-out below)
        }
        // ---- pivot rows -> U | E | F (layout of abd.cuh), pivot slots -> piv --------------------------------------
        if (lane < NB) {
            reinterpret_cast<unsigned *>(L.piv + g * (int64_t)Cf::PIV_STRIDE)[lane] = mypword;
        }
#ifndef ABD4_X_NOWRITE
#pragma unroll
        for (int T = 0; T < RT; ++T) {      // final pivot rows, ordered by pivot index
            const int k = pvs[8 * T + gid];
#pragma unroll
            for (int TC = 0; TC < CT; ++TC) sts128_if(k >= 0, xs + k * PITCH_W + 8 * TC + 2 * q, X[T][TC]);
        }
        __syncwarp();
#ifndef ABD4_X_NOESTORE
        if constexpr (RHS) {                // transformed right-hand side of the pivot rows, in pivot order
            if (lane < NV) L.e[g * NV + lane] = xs[lane * PITCH_W + Cf::NC];
        }
#endif
        {   // coalesced write-out: lane handles doubles lane, lane + 32, ... of the record
            double *uef = L.UEF + g * (int64_t)Pk::UEF + lane;
            const char *xb = reinterpret_cast<const char *>(xs);
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const unsigned off = (m % 2 == 0) ? (wmap[m / 2] & 0xffffu) : (wmap[m / 2] >> 16);
                const double v = *reinterpret_cast<const double *>(xb + off);
                if (32 * m + 31 < Pk::UEF || lane + 32 * m < Pk::UEF) st_cs(uef + 32 * m, v);
            }
        }
#endif
        carrymask = cand_mask;
        // (the next block row's first __syncwarp orders these xs reads before its writes)
    }
    if (bad && lane == 0) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
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
                if constexpr (RHS) {
                    if (part == 3 && q == 2) L.r_next[slab * NV + t] = X[T][TC].x;
                }
            });
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward sweep for the blocked factors: lane = slot, w = right-hand side entry of the slot.  Per block row: 4-byte
// pivot records (warp-uniform), per block 4 coalesced Lt loads, 4 independent broadcasts of the raw pivot entries, 4 FMA.
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void abd_forward4_slab(const AbdLevel &L, int64_t slab, int lane) {
    using Cf = Abd4Cfg<NV>;
    constexpr int R2 = Cf::R2, NB = Cf::NB;
    if (slab >= L.nslabs) return;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    const unsigned below = lanes_below(lane);
    unsigned carrymask = (1u << NV) - 1u;
    double w = lane < NV ? L.r[g0 * NV + lane] : 0.0;
    // pivot records run one block row ahead of the chain: the multiplier addresses depend on them, so fetching them inside
    // the step would put two dependent memory round trips on every step of the (latency-bound) small levels
    unsigned pnext[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) pnext[b] = len > 1 ? __ldg(reinterpret_cast<const unsigned *>(L.piv + (g0 + 1) * (int64_t)Cf::PIV_STRIDE) + b) : 0u;
    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        const double *lt_g = L.Lm + g * (int64_t)Cf::LT;
        if (L.prefetch && i + L.prefetch < len && lane * 128 < Cf::LT * 8)   // L.prefetch = distance in block rows
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(lt_g + (int64_t)L.prefetch * Cf::LT) + lane * 128));
        unsigned pword[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) pword[b] = pnext[b];
        if (i + 1 < len) {
            const unsigned *pw = reinterpret_cast<const unsigned *>(L.piv + (g + 1) * (int64_t)Cf::PIV_STRIDE);
#pragma unroll
            for (int b = 0; b < NB; ++b) pnext[b] = __ldg(pw + b);
        }
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

template <int NV>
__global__ void __launch_bounds__(128) abd_forward4_kernel(AbdLevel L) {
    const int64_t slab = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();   // programmatic dependent launch, see abd.cuh
    if (slab < L.nslabs && slab * (int64_t)L.S + 1 < L.n) {   // Lt and pivot record of the slab's first elimination step
        const int64_t g1 = slab * (int64_t)L.S + 1;
        if (lane * 128 < Abd4Cfg<NV>::LT * 8)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(L.Lm + g1 * (int64_t)Abd4Cfg<NV>::LT) + lane * 128));
        if (lane == 31) asm volatile("prefetch.global.L2 [%0];" ::"l"(L.piv + g1 * (int64_t)Abd4Cfg<NV>::PIV_STRIDE));
    }
    pdl_wait();
    abd_forward4_slab<NV>(L, slab, lane);
}

// ------------------------------------------------------------------------------------------------
// Tail of the level hierarchy: every level with at most ABD_TAIL_ROWS block rows runs inside ONE single-CTA launch per
// sweep direction (levels separated by block barriers; data moves through HBM / L2 as between launches).  At 10^6 rows
// the small levels are pure latency: six launches of 10-18 us each per sweep, 20 % of a linear solve
// (profiles/r02_summary.md).  The backward launch starts with the closure of the last block row (boundary conditions)
// when the handle owns the whole mesh.
// ------------------------------------------------------------------------------------------------
constexpr int ABD_TAIL_ROWS = 64;    // one slab per warp at the largest tail level: a single CTA cannot hide DRAM latency with several slabs per warp
constexpr int ABD_TAIL_MAX_LEVELS = 6;
constexpr int ABD_TAIL_THREADS = 512;

struct AbdTail {
    int nlev;
    int do_top;
    AbdLevel lv[ABD_TAIL_MAX_LEVELS];   // fine -> coarse
    const double *C, *D, *g;            // last block row (closure)
    double *Ztop;
    unsigned long long left_mask, right_mask;
};

// every thread of the CTA: pull [p, p + bytes) into L2 (the factors of the small levels were written a whole
// factorisation ago: without this every elimination step of the tail waits for a DRAM round trip)
__device__ __forceinline__ void prefetch_l2_range(const void *p, size_t bytes) {
    const char *c = reinterpret_cast<const char *>(p);
    for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)blockDim.x * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}

template <int NV>
__global__ void __launch_bounds__(ABD_TAIL_THREADS) abd_tail_forward_kernel(AbdTail T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    pdl_launch_dependents();
    for (int l = 0; l < T.nlev; ++l) {
        prefetch_l2_range(T.lv[l].Lm, (size_t)T.lv[l].n * Abd4Cfg<NV>::LT * sizeof(double));
        prefetch_l2_range(T.lv[l].piv, (size_t)T.lv[l].n * Abd4Cfg<NV>::PIV_STRIDE);
    }
    pdl_wait();                // the factors travel while the predecessor finishes (programmatic dependent launch, abd.cuh)
    for (int l = 0; l < T.nlev; ++l) {
        const AbdLevel &L = T.lv[l];
        for (int64_t slab = warp; slab < L.nslabs; slab += nwarps) abd_forward4_slab<NV>(L, slab, lane);
        __syncthreads();   // r_next of this level is the right-hand side of the next one
    }
}

template <int NV>
__global__ void __launch_bounds__(ABD_TAIL_THREADS) abd_tail_backward_kernel(AbdTail T, unsigned *flags) {
    __shared__ AbdTopSmem<NV> top_sm;
    constexpr int GPW = 32 / next_pow2(NV);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    pdl_launch_dependents();
    for (int l = T.nlev - 1; l >= 0; --l) prefetch_l2_range(T.lv[l].UEF, (size_t)T.lv[l].n * AbdPack<NV>::UEF * sizeof(double));
    pdl_wait();
    for (int l = T.nlev - 1; l >= 0; --l) prefetch_l2_range(T.lv[l].e, (size_t)T.lv[l].n * NV * sizeof(double));
    if (T.do_top) {
        if (warp == 0) abd_top_solve_warp<NV>(top_sm, T.C, T.D, T.g, T.Ztop, T.left_mask, T.right_mask, flags, lane);
        __syncthreads();
    }
    for (int l = T.nlev - 1; l >= 0; --l) {
        const AbdLevel &L = T.lv[l];
        const int64_t nw = (L.nslabs + GPW - 1) / GPW;
        for (int64_t w = warp; w < nw; w += nwarps) abd_backward_warp<NV>(L, w, lane);
        __syncthreads();   // Z of this level holds the slab end values of the next finer one
    }
}

}  // namespace rst
