// bigd.cuh -- wide-block (nv = 64, config C5) factor kernel with the rank-4 updates on the FP64 tensor cores.
//
// north_star: "FP64 DMMA tensor cores are used only if ncu shows they beat the FP64 FMA pipe at the config's block size".
// Measured (tools/dmma_bench.cu, profiles/r02_fp64_peaks.md): DMMA m8n8k4 and DFMA share ONE pipe with ONE peak
// (63.8 FMA/clk/SM = 37.1 TFLOP/s), so DMMA cannot raise the ceiling -- but one DMMA does the work of 8 DFMA warp
// instructions and needs no per-FMA operand broadcast.  bigr_factor_kernel (big.cuh, register-tiled DFMA, one block
// barrier per elimination step) sits at 18 % of the pipe because every step is a publish / barrier / retire / update
// chain in which ALU, LSU and FP64 phases never overlap.  This kernel keeps bigr's data placement (the 128 x 192 panel
// [M | P | Q] fills the register file of one 16-warp CTA) and bigr's factor layout (Lm [g][k][row], UEF [g][c][k], piv),
// and restructures the elimination exactly like abd4.cuh:
//   * warp w owns panel rows 8w .. 8w+7 as 24 accumulator tiles of mma.sync m8n8k4 (48 registers per lane);
//   * elimination is blocked by 4 pivots: the 128 x 4 pivot-column panel goes through shared memory to ONE warp
//     (4 rows per lane) that factors it without any block barrier (REDUX arg-max, pivot row through a 32-byte
//     shared-memory broadcast);
//   * the block's multipliers are folded with the inverse of its 4 x 4 unit lower triangle, Lt = L L11^-1, so the
//     trailing update X <- X - Lt X[pivot rows] uses the RAW pivot rows: 24 - b/2 DMMA per warp and block, 3 block
//     barriers per 4 pivots instead of 4;
//   * the sequential multipliers L (what the sweeps of big.cuh apply) are stored by the panel warp, the final pivot
//     rows are transposed through shared memory into UEF [c][k] with coalesced stores.
// Block sizes other than 64 keep bigr_factor_kernel.
//
// Where the time goes (tools/bigd_probe.cu = this kernel with the BIGD_X_* / BIGD_NO_DMMA / BIGD_PW1 what-if switches below,
// tools/coissue_bench.cu, profiles/r02_coissue.md): time = 20 + 3.6 per pivot round + 17.5 ms for the trailing DMMAs at C5's
// size, exactly additive -- the look-ahead does NOT overlap, because a scheduler's one FP64 pipe serves DMMA before DFMA: a
// dependent DFMA takes 8 cycles alone, 61 next to one DMMA-streaming warp and is starved next to four, so the panel warps'
// first FP64 instruction waits for the update warps of their scheduler to finish the block.  All panel warps on one
// scheduler (BIGD_X_PANEL_ONE_SCHEDULER, 52.7 ms), one panel warp with 4 rows per lane there (BIGD_PW1, 58.2 ms), a 16-warp
// hybrid at 128 registers (51.2 ms) and an instruction-lean chain (this version) all measure the same or worse.
#pragma once
#include "abd4.cuh"
#include "big.cuh"

namespace rst {

struct BigdCfg {
    static constexpr int NV = 64, R2 = 128, NC = 192, CT = 24, NB = 16, WARPS = 16, THREADS = 512;
    static constexpr int PAN_PITCH = 136;   // pivot-column plane pitch (== 8 mod 16: the two planes a store hits differ in bank)
    static constexpr int LB_PLANE = 2 * R2 + 8;
    static constexpr int XS_PITCH = 200;    // == 8 mod 16: B-fragment reads of 4 rows x 8 columns hit every bank twice (minimum)
    static constexpr int UST_PITCH = 193;   // odd: the transposing read of the write-out is conflict free
};

struct BigdSmem {
    double pan[4][BigdCfg::PAN_PITCH];
    double lb[2][BigdCfg::LB_PLANE];
    double xs[4][BigdCfg::XS_PITCH];
    double slot[2][4][4];             // candidate pivot row of each panel warp (its own entry replaced by 1 / pivot), by pivot parity
    unsigned slot_key[2][4];
    double ust[BigdCfg::NV][BigdCfg::UST_PITCH];
    unsigned masks[4];
    int tmap[BigdCfg::R2];            // free row -> index of the block row it receives; carried row -> its rank; else -1
    signed char pv[BigdCfg::R2];      // pivot index of the row in the current block row, -1 = survivor
    unsigned char wh[BigdCfg::R2];    // 0..3: the row is that pivot of the current block, 4: it is not
};

__device__ __forceinline__ void bar_named(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// One block (4 pivots) of the 128 x 4 pivot-column panel, one panel row per lane of the 4 calling warps (pw = 0..3): partial
// pivoting with ONE named barrier (id 3, 128 threads) per pivot -- every warp publishes its best candidate row speculatively --
// sequential multipliers L -> HBM (what the sweeps apply), -Lt = -L L11^-1 -> shared memory (A fragments), pivot identities.
#ifdef BIGD_TIMING
#define BIGD_TICK(k) { const long long t_now = clock64(); tseg[k] += t_now - t_last; t_last = t_now; }
#else
#define BIGD_TICK(k)
#endif
__device__ __forceinline__ void bigd_panel_block(BigdSmem &sm, const AbdLevel &L, int64_t g, int b, int pw, int lane, bool &cand,
                                                 unsigned &bad
#ifdef BIGD_TIMING
                                                 , long long *tseg
#endif
                                                 ) {
#ifdef BIGD_TIMING
    long long t_last = clock64();
#endif
    constexpr int NV = BigdCfg::NV, R2 = BigdCfg::R2, PAN = 128;
    const int myrow = 32 * pw + lane;
    double a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] = sm.pan[(c >> 1) | ((c & 1) << 1)][myrow];
    const bool cand0 = cand;
    int mypiv = 4;
    // The panel chain runs concurrently with the update warps' DMMAs, and every FP64 instruction a panel warp issues queues
    // behind them in its scheduler's one FP64 pipe (~40 cycles each, dependent or not; tools/bigd_probe.cu: the chain costs
    // 32.6 ms alone and 50.6 ms with the DMMAs, i.e. the two add up).  So the chain is written for the fewest FP64 INSTRUCTIONS:
    //  * DIVISION-FREE elimination inside the block: live rows are updated as  a'[c] = (pi s) a[c] - (a[ii] s) ap[c]  with pi the
    //    pivot and s the power of two that brings pi into [1, 2), applied through the exponent field on the integer pipe.
    //    Every live row (and the next pivot candidates) carries the same factor, so the arg-max is unchanged, and the
    //    multipliers  l = a[ii] / pi  are scale free -- they are formed after the four pivots.  2 FP64 instructions per column.
    //  * the warp-uniform values of the block (4 reciprocals, the 6 entries of L11) are computed ONCE, spread over 10 lanes,
    //    and broadcast by shuffles: 5 FP64 instructions instead of 26.
#ifdef BIGD_X_SHORTPANEL   // what-if: one pivot round instead of four (wrong numbers, timing only)
    constexpr int NPIV = 1;
    double piv[4] = {1.0, 1.0, 1.0, 1.0};
    double num[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#else
    constexpr int NPIV = 4;
    double piv[4];            // pivot values (in the block's running scale)
    double num[6];            // entries of the later pivot rows under the earlier pivots: n_ij = num / piv[j]
#endif
#pragma unroll
    for (int ii = 0; ii < NPIV; ++ii) {
        // key = (high word of |a| with its low 7 bits replaced by 127 - row) + 1: threshold pivoting (2^-13
        // relative), ties to the lowest row.  Every warp publishes its best row speculatively: ONE barrier per pivot.
        const unsigned hw = (unsigned)__double2hiint(a[ii]) & 0x7fffffffu;
        const unsigned key = cand ? (((hw & ~127u) | (unsigned)(127 - myrow)) + 1u) : 0u;
        const unsigned mxw = __reduce_max_sync(RST_FULL_MASK, key);
        if (lane == 31 - (int)((mxw - 1u) & 31u)) {
            double *sl = &sm.slot[ii & 1][pw][0];
            *reinterpret_cast<double2 *>(sl) = make_double2(a[0], a[1]);
            *reinterpret_cast<double2 *>(sl + 2) = make_double2(a[2], a[3]);
            sm.slot_key[ii & 1][pw] = mxw;
        }
        BIGD_TICK(0)
        bar_named(3, PAN);
        BIGD_TICK(1)
        unsigned mx = sm.slot_key[ii & 1][0];
#pragma unroll
        for (int w = 1; w < 4; ++w) mx = max(mx, sm.slot_key[ii & 1][w]);
        bad |= (mx <= 128u || mx > 0x7ff00000u) ? 1u : 0u;
        const int prow = 127 - (int)((mx - 1u) & 127u);
        double ap[4];
        {
            const double *sl = &sm.slot[ii & 1][prow >> 5][0];
            const double2 p0 = *reinterpret_cast<const double2 *>(sl);
            const double2 p1 = *reinterpret_cast<const double2 *>(sl + 2);
            ap[0] = p0.x; ap[1] = p0.y; ap[2] = p1.x; ap[3] = p1.y;
        }
        piv[ii] = ap[ii];
        if (ii == 1) { num[0] = ap[0]; }
        if (ii == 2) { num[1] = ap[0]; num[2] = ap[1]; }
        if (ii == 3) { num[3] = ap[0]; num[4] = ap[1]; num[5] = ap[2]; }
        const bool is_piv = myrow == prow;
        const bool upd = cand && !is_piv;
        if (ii < 3) {
            // s = 2^d, d = -(exponent of the pivot): pi s keeps the pivot's mantissa, a[ii] s is an exponent subtraction
            // (|a[ii]| <= (1 + 2^-13) |pi|: no overflow; underflow -> 0).  Denormal pivots are not scaled.
            const int phi = __double2hiint(ap[ii]);
            const int ex = (phi >> 20) & 0x7ff;
            const int d = ex != 0 ? 1023 - ex : 0;
            const double pis = __hiloint2double(ex != 0 ? ((phi & 0x800fffff) | 0x3ff00000) : phi, __double2loint(ap[ii]));
            const int ahi = __double2hiint(a[ii]);
            const int ae = (ahi >> 20) & 0x7ff;
            const bool ok = ae != 0 && ae + d > 0;
            const double ais = __hiloint2double(ok ? ahi + d * (1 << 20) : (ahi & (int)0x80000000), ok ? __double2loint(a[ii]) : 0);
#ifndef BIGD_X_NOPIVFP
            if (upd) {
#pragma unroll
                for (int c = ii + 1; c < 4; ++c) a[c] = fma(-ais, ap[c], pis * a[c]);
            }
#else
            if (upd && ais == 123.0) a[ii + 1] = pis;
#endif
        }
        if (is_piv) { mypiv = ii; cand = false; }
        BIGD_TICK(2)
    }
    // warp-uniform values, one per lane: lanes 0..5  n_k = num[k] / piv[j(k)]  (j = 0 0 1 0 1 2), lanes 6..9  1 / piv[lane - 6]
    double n10, n20, n21, n30, n31, n32, rinv[4];
#ifdef BIGD_X_NOTAILFP
    n10 = num[0]; n20 = num[1]; n21 = num[2]; n30 = num[3]; n31 = num[4]; n32 = num[5];
    rinv[0] = piv[0]; rinv[1] = piv[1]; rinv[2] = piv[2]; rinv[3] = piv[3];
    const int lim = cand0 ? mypiv : 0;
    const double l0 = lim > 0 ? a[0] : 0.0, l1 = lim > 1 ? a[1] : 0.0, l2 = lim > 2 ? a[2] : 0.0, l3 = lim > 3 ? a[3] : 0.0;
    const double nt3 = l3, nt2 = l2, nt1 = lim > 1 ? n10 : l1, nt0 = lim > 2 ? n32 : l0;
#else
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
    // L11^-1 (unit lower triangle, entries i_ij): i10 = -n10, i21 = -n21, i32 = -n32 are used through operand negation
    const double i20 = fma(n21, n10, -n20);
    const double i31 = fma(n32, n21, -n31);
    const double i30 = fma(-n32, i20, fma(n31, n10, -n30));
    const int lim = cand0 ? mypiv : 0;   // pivot ii keeps the multipliers of the earlier pivots
    // multipliers l_j = a[j] / piv[j] (a[j] is untouched after step j)
    const double l0 = lim > 0 ? a[0] * rinv[0] : 0.0, l1 = lim > 1 ? a[1] * rinv[1] : 0.0, l2 = lim > 2 ? a[2] * rinv[2] : 0.0,
                 l3 = lim > 3 ? a[3] * rinv[3] : 0.0;
    // -Lt = -L L11^-1: what the tensor-core update adds to the raw pivot rows' multiples
    const double nt3 = __hiloint2double(__double2hiint(l3) ^ (int)0x80000000, __double2loint(l3));
    const double nt2 = fma(l3, n32, -l2);
    const double nt1 = fma(-l3, i31, fma(l2, n21, -l1));
    const double nt0 = fma(-l3, i30, fma(-l2, i20, fma(l1, n10, -l0)));
#endif
    *reinterpret_cast<double2 *>(&sm.lb[0][2 * myrow]) = make_double2(nt0, nt1);
    *reinterpret_cast<double2 *>(&sm.lb[1][2 * myrow]) = make_double2(nt2, nt3);
    sm.wh[myrow] = (unsigned char)mypiv;
    if (mypiv < 4) sm.pv[myrow] = (signed char)(4 * b + mypiv);
    BIGD_TICK(3)
    // sequential multipliers: what the sweeps apply (layout of big.cuh) -- after the hand-over data, off the chain
    double *lm = L.Lm + g * (int64_t)(NV * R2) + (int64_t)(4 * b) * R2 + myrow;
    st_cs(lm, l0); st_cs(lm + R2, l1); st_cs(lm + 2 * R2, l2); st_cs(lm + 3 * R2, l3);
    BIGD_TICK(4)
}

// The same block of 4 pivots factored by ONE warp, 4 panel rows per lane (row = lane + 32 j): no barrier and no shared-memory
// round trip inside -- arg-max = per-lane maximum over 4 rows + one REDUX, pivot row = 4 shuffles.  Same arithmetic, bit for bit.
__device__ __forceinline__ void bigd_panel_block_w1(BigdSmem &sm, const AbdLevel &L, int64_t g, int b, int lane, unsigned &candbits,
                                                    unsigned &bad) {
    constexpr int NV = BigdCfg::NV, R2 = BigdCfg::R2;
    double a[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) a[j][c] = sm.pan[(c >> 1) | ((c & 1) << 1)][lane + 32 * j];
    const unsigned cand0 = candbits;
    int whj[4] = {4, 4, 4, 4};
    double piv[4], num[6];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        unsigned kbest = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned hw = (unsigned)__double2hiint(a[j][ii]) & 0x7fffffffu;
            const unsigned key = ((candbits >> j) & 1u) ? (((hw & ~127u) | (unsigned)(127 - (lane + 32 * j))) + 1u) : 0u;
            kbest = max(kbest, key);
        }
        const unsigned mx = __reduce_max_sync(RST_FULL_MASK, kbest);
        bad |= (mx <= 128u || mx > 0x7ff00000u) ? 1u : 0u;
        const int prow = 127 - (int)((mx - 1u) & 127u);
        const int lp = prow & 31, jp = prow >> 5;      // warp-uniform
        double ap[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double r = a[0][c];
            r = jp == 1 ? a[1][c] : r;
            r = jp == 2 ? a[2][c] : r;
            r = jp == 3 ? a[3][c] : r;
            ap[c] = __shfl_sync(RST_FULL_MASK, r, lp);
        }
        piv[ii] = ap[ii];
        if (ii == 1) { num[0] = ap[0]; }
        if (ii == 2) { num[1] = ap[0]; num[2] = ap[1]; }
        if (ii == 3) { num[3] = ap[0]; num[4] = ap[1]; num[5] = ap[2]; }
        if (ii < 3) {
            const int phi = __double2hiint(ap[ii]);
            const int ex = (phi >> 20) & 0x7ff;
            const int d = ex != 0 ? 1023 - ex : 0;
            const double pis = __hiloint2double(ex != 0 ? ((phi & 0x800fffff) | 0x3ff00000) : phi, __double2loint(ap[ii]));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool upd = ((candbits >> j) & 1u) && !((lane == lp) && (jp == j));
                const int ahi = __double2hiint(a[j][ii]);
                const int ae = (ahi >> 20) & 0x7ff;
                const bool ok = ae != 0 && ae + d > 0;
                const double ais = __hiloint2double(ok ? ahi + d * (1 << 20) : (ahi & (int)0x80000000), ok ? __double2loint(a[j][ii]) : 0);
                if (upd) {
#pragma unroll
                    for (int c = ii + 1; c < 4; ++c) a[j][c] = fma(-ais, ap[c], pis * a[j][c]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((lane == lp) && (jp == j)) whj[j] = ii;
        if (lane == lp) candbits &= ~(1u << jp);
    }
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
    const double i20 = fma(n21, n10, -n20);
    const double i31 = fma(n32, n21, -n31);
    const double i30 = fma(-n32, i20, fma(n31, n10, -n30));
    double *lm = L.Lm + g * (int64_t)(NV * R2) + (int64_t)(4 * b) * R2 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = lane + 32 * j;
        const int lim = ((cand0 >> j) & 1u) ? whj[j] : 0;
        const double l0 = lim > 0 ? a[j][0] * rinv[0] : 0.0, l1 = lim > 1 ? a[j][1] * rinv[1] : 0.0,
                     l2 = lim > 2 ? a[j][2] * rinv[2] : 0.0, l3 = lim > 3 ? a[j][3] * rinv[3] : 0.0;
        const double nt3 = __hiloint2double(__double2hiint(l3) ^ (int)0x80000000, __double2loint(l3));
        const double nt2 = fma(l3, n32, -l2);
        const double nt1 = fma(-l3, i31, fma(l2, n21, -l1));
        const double nt0 = fma(-l3, i30, fma(-l2, i20, fma(l1, n10, -l0)));
        *reinterpret_cast<double2 *>(&sm.lb[0][2 * r]) = make_double2(nt0, nt1);
        *reinterpret_cast<double2 *>(&sm.lb[1][2 * r]) = make_double2(nt2, nt3);
        sm.wh[r] = (unsigned char)whj[j];
        if (whj[j] < 4) sm.pv[r] = (signed char)(4 * b + whj[j]);
        st_cs(lm + 32 * j, l0); st_cs(lm + R2 + 32 * j, l1); st_cs(lm + 2 * R2 + 32 * j, l2); st_cs(lm + 3 * R2 + 32 * j, l3);
    }
}

// 16 update warps (accumulator tiles) + 4 panel warps (one panel row per lane).  Five warps per scheduler cap the kernel
// at 96 registers per thread -- exactly the 24 tiles of an update warp; the panel warps need 8.
template <bool IDENT_B>
__global__ void __launch_bounds__(BigdCfg::THREADS + 128, 1) bigd_factor_kernel(AbdLevel L, unsigned *flags) {
    using C = BigdCfg;
    constexpr int NV = C::NV, R2 = C::R2, CT = C::CT, NB = C::NB;
    constexpr int MT = NV / 8;   // column tiles per part: M = tiles 0..7, P = 8..15, Q = 16..23
#ifdef BIGD_PW1
    constexpr int ALL = C::THREADS + 32, UPD = C::THREADS;    // barriers 1 / 2
#else
    constexpr int ALL = C::THREADS + 128, UPD = C::THREADS;   // barriers 1 / 2 (3: among the four panel warps)
#endif
    extern __shared__ __align__(16) unsigned char bigd_smem_raw[];
    BigdSmem &sm = *reinterpret_cast<BigdSmem *>(bigd_smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#if defined(BIGD_PW1)
    // ONE panel warp (4 rows per lane) alone with one update warp on scheduler 0: hardware warp 0 = panel, 4 / 8 / 12 retire at
    // once, the other sixteen are the update warps (5 on each of the schedulers 1-3, 1 on scheduler 0).
    if (wid == 4 || wid == 8 || wid == 12) return;
    const bool is_pan = wid == 0;
    const int warp = is_pan ? 0 : (wid < 16 ? wid - (wid >> 2) - 1 : wid - 4);
#elif !defined(BIGD_X_PANEL_ONE_SCHEDULER)
    // one panel warp per scheduler: hardware warps 16..19
    const bool is_pan = wid >= C::WARPS;
    const int warp = is_pan ? wid - C::WARPS : wid;
#else
    // what-if (tools/bigd_probe.cu): all four panel warps on scheduler 0 (hardware warps 0, 4, 8, 12), which then hosts only
    // ONE update warp.  Measured: the DMMAs now cost +6 ms instead of +17 ms on top of the chain, but the four lock-stepped
    // panel warps serialise their ~290 instructions per block on one issue port (+12 ms): 52.7 vs 52.0 ms, no gain.
    const bool is_pan = wid < 16 && (wid & 3) == 0;
    const int warp = is_pan ? wid >> 2 : (wid < 16 ? wid - (wid >> 2) - 1 : wid - 4);   // panel warp 0..3 / update warp 0..15
#endif
    const int tid = 32 * warp + lane;   // thread index within the role
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);

    if (is_pan) {
        // ================= panel warps: factor the 128 x 4 pivot-column panel of every block, one row per lane =================
        const int pw = warp;
        unsigned bad = 0u;
#ifdef BIGD_TIMING
        long long t_wait = 0, t_lu = 0, tseg[5] = {0, 0, 0, 0, 0};
#endif
        for (int istep = 1; istep < len; ++istep) {
            const int64_t g = g0 + istep;
            bool cand = true;               // my row has not been pivoted in this block row
            unsigned candbits = 0xFu;       // BIGD_PW1: bit j = row lane + 32 j has not been pivoted
#pragma unroll 1
            for (int b = 0; b < NB; ++b) {
#ifdef BIGD_TIMING
                const long long tp0 = clock64();
#endif
                bar_named(1, ALL);          // A(b): the pivot-column panel of block b is in shared memory
#ifdef BIGD_TIMING
                const long long tp1 = clock64();
#endif
#ifdef BIGD_TIMING
                bigd_panel_block(sm, L, g, b, pw, lane, cand, bad, tseg);
                t_wait += tp1 - tp0; t_lu += clock64() - tp1;
#elif defined(BIGD_PW1)
                bigd_panel_block_w1(sm, L, g, b, lane, candbits, bad);
#else
                bigd_panel_block(sm, L, g, b, pw, lane, cand, bad);
#endif
                bar_named(1, ALL);          // B(b): Lt and the pivot identities of block b are published
            }
        }
        if (bad && lane == 0) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
#ifdef BIGD_TIMING
        if (blockIdx.x == 0 && pw == 0 && lane == 0)
            printf("panel warp: per block  wait for A %lld  LU %lld cycles = publish %lld + barrier %lld + read/update %lld (4 pivots) + tail %lld + Lm stores %lld\n",
                   t_wait / ((len - 1) * NB), t_lu / ((len - 1) * NB), tseg[0] / ((len - 1) * NB), tseg[1] / ((len - 1) * NB),
                   tseg[2] / ((len - 1) * NB), tseg[3] / ((len - 1) * NB), tseg[4] / ((len - 1) * NB));
#endif
        return;
    }

    // ================= update warps: warp w owns panel rows 8w .. 8w+7 as 24 accumulator tiles =================
    const int gid = lane >> 2, q = lane & 3;
    const bool hi = q >= 2;
    const int row = 8 * warp + gid;                 // my panel row in accumulator layout
#ifdef BIGD_TIMING
    long long t_serial = 0, t_dmma = 0, t_waitb = 0, t_step0 = clock64();
#endif
    double2 X[CT];
    // carried rows of the slab start: rows 0..63 = block row g0 as [ . | P = A | Q = B ]; rows 64..127 are free
#pragma unroll
    for (int tc = 0; tc < CT; ++tc) X[tc] = make_double2(0.0, 0.0);
    if (row < NV) {
#pragma unroll
        for (int tc = 0; tc < MT; ++tc) {
            X[MT + tc] = *reinterpret_cast<const double2 *>(L.A + (g0 * NV + row) * NV + 8 * tc + 2 * q);
            if constexpr (IDENT_B) X[2 * MT + tc] = make_double2(8 * tc + 2 * q == row ? 1.0 : 0.0, 8 * tc + 2 * q + 1 == row ? 1.0 : 0.0);
            else X[2 * MT + tc] = *reinterpret_cast<const double2 *>(L.B + (g0 * NV + row) * NV + 8 * tc + 2 * q);
        }
    }
    if (tid < R2) {
        sm.tmap[tid] = tid < NV ? -1 : tid - NV;   // next block row's row t goes to free row 64 + t
        sm.pv[tid] = tid < NV ? -1 : 0;            // >= 0: free
    }
    bar_named(2, UPD);

    // pivot columns of block b (column tile b / 2, half b & 1) -> shared memory, one plane per column
    auto post_panel = [&](const double2 &v, bool blk_hi) {
        if (hi == blk_hi) {
            sm.pan[q & 1][row] = v.x;          // column 2 (q & 1)     -> plane (q & 1)
            sm.pan[2 + (q & 1)][row] = v.y;    // column 2 (q & 1) + 1 -> plane 2 + (q & 1)
        }
    };

    for (int istep = 1; istep < len; ++istep) {
        const int64_t g = g0 + istep;
        // ---- carried rows: M <- Q, Q <- 0; free rows: next block row [M = A | 0 | Q = B] ----------------------------
        {
            const int t = sm.tmap[row];
            const bool fr = sm.pv[row] >= 0;
            if (fr) {
                const double *arow = L.A + (g * NV + t) * NV + 2 * q;
#pragma unroll
                for (int tc = 0; tc < MT; ++tc) {
                    X[tc] = ld_cs2(reinterpret_cast<const double2 *>(arow + 8 * tc));
                    X[MT + tc] = make_double2(0.0, 0.0);
                    if constexpr (IDENT_B) X[2 * MT + tc] = make_double2(8 * tc + 2 * q == t ? 1.0 : 0.0, 8 * tc + 2 * q + 1 == t ? 1.0 : 0.0);
                    else X[2 * MT + tc] = ld_cs2(reinterpret_cast<const double2 *>(L.B + (g * NV + t) * NV + 8 * tc + 2 * q));
                }
            } else {
#pragma unroll
                for (int tc = 0; tc < MT; ++tc) {
                    X[tc] = X[2 * MT + tc];
                    X[2 * MT + tc] = make_double2(0.0, 0.0);
                }
            }
            if (istep + 1 < len && tid < NV * NV * 8 / 128)   // L2 prefetch of the next block row (32 KB)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(L.A + (g + 1) * (int64_t)(NV * NV)) + tid * 128));
        }
        bar_named(2, UPD);   // tmap / pv of the previous block row are consumed
        if (tid < R2) sm.pv[tid] = -1;
        post_panel(X[0], false);
        bar_named(1, ALL);   // A(0)
        bar_named(1, ALL);   // B(0)

#pragma unroll
        for (int b = 0; b < NB; ++b) {            // fully unrolled: every tile index below is a compile-time constant
            const int TCB = b >> 1;
            const int TN = (b + 1) >> 1;           // column tile that holds the pivot columns of block b + 1
            const bool more = b + 1 < NB;
#ifdef BIGD_TIMING
            const long long tu0 = clock64();
#endif
            // ---- raw rows of the block's 4 pivots -> shared memory (B fragments) ---------------------------------------
            const int which = sm.wh[row];
            if (which < 4) {
                double *dst = &sm.xs[which][2 * q];
#pragma unroll
                for (int tc = 0; tc < CT; ++tc)
                    if (tc >= TCB) *reinterpret_cast<double2 *>(dst + 8 * tc) = X[tc];
            }
            const double af = sm.lb[q >> 1][2 * row + (q & 1)];
            bar_named(2, UPD);
            const double *bsrc = &sm.xs[q][gid];
            // ---- look-ahead: update the tile with the next block's pivot columns first and hand them to the panel warps ----
            if (more) {
                dmma_m8n8k4(X[TN], af, bsrc[8 * TN]);
                post_panel(X[TN], ((b + 1) & 1) != 0);
                bar_named(1, ALL);   // A(b + 1): the panel warps factor block b + 1 while the rest of block b is applied
            }
#ifdef BIGD_TIMING
            const long long tu1 = clock64();
#endif
            // ---- X <- X - Lt X[pivot rows] on the remaining tiles ------------------------------------------------------------
#ifndef BIGD_NO_DMMA   // what-if switch of tools/bigd_probe.cu: the panel chain alone
#pragma unroll
            for (int tc = 0; tc < CT; ++tc)
                if (tc >= TCB && !(more && tc == TN)) dmma_m8n8k4(X[tc], af, bsrc[8 * tc]);
#endif
#ifdef BIGD_TIMING
            const long long tu2 = clock64();
#endif
#if !defined(BIGD_X_NOWRITE) && !defined(BIGD_X_WRITE_AT_END)
            // U | E | F of the PREVIOUS block row leave shared memory here, two of its 24 store rounds per block, in the time the
            // update warps would otherwise spend waiting for the panel warps (the staging buffer is rewritten after block 15)
            if (istep > 1 && b < 3 * NV * NV / C::THREADS / 2) {
                double *uef = L.UEF + (g - 1) * (int64_t)(3 * NV * NV);
#pragma unroll
                for (int m = 2 * b; m < 2 * b + 2; ++m) {
                    const int idx = tid + C::THREADS * m;      // = c * 64 + k
                    st_cs(uef + idx, sm.ust[idx & (NV - 1)][idx >> 6]);
                }
            }
#endif
            if (more) bar_named(1, ALL);   // B(b + 1)
#ifdef BIGD_TIMING
            t_serial += tu1 - tu0; t_dmma += tu2 - tu1; t_waitb += clock64() - tu2;
#endif
        }
        bar_named(2, UPD);
        // ---- factors of this block row: pivot rows -> UEF [c][k] through shared memory, piv, row map of the next step ----
        {
            const int k = sm.pv[row];
            if (k >= 0) {
                double *dst = &sm.ust[k][2 * q];
#pragma unroll
                for (int tc = 0; tc < CT; ++tc) { dst[8 * tc] = X[tc].x; dst[8 * tc + 1] = X[tc].y; }
            }
        }
        if (tid < R2) {
            const int k = sm.pv[tid];
            L.piv[g * R2 + tid] = (int8_t)k;
            // rank of the row among the free rows (next block row) resp. among the survivors (condensed row)
            const unsigned mf = __ballot_sync(RST_FULL_MASK, k >= 0);
            if (lane == 0) sm.masks[warp] = mf;
        }
        bar_named(2, UPD);
#if !defined(BIGD_X_NOWRITE)
#if !defined(BIGD_X_WRITE_AT_END)
        if (istep == len - 1)      // the last block row of the slab has no successor to hide its write-out behind
#endif
        {
            double *uef = L.UEF + g * (int64_t)(3 * NV * NV);
#pragma unroll 4
            for (int m = 0; m < 3 * NV * NV / C::THREADS; ++m) {
                const int idx = tid + C::THREADS * m;      // = c * 64 + k
                st_cs(uef + idx, sm.ust[idx & (NV - 1)][idx >> 6]);
            }
        }
#endif
        if (tid < R2) {
            const bool fr = sm.pv[tid] >= 0;
            int t = 0;
            for (int w = 0; w < warp; ++w) t += fr ? __popc(sm.masks[w]) : 32 - __popc(sm.masks[w]);
            const unsigned mine = fr ? sm.masks[warp] : ~sm.masks[warp];
            t += __popc(mine & ((1u << lane) - 1u));
            sm.tmap[tid] = t;
        }
        bar_named(2, UPD);
    }
#ifdef BIGD_TIMING
    if (blockIdx.x == 0 && tid == 32 * 5 && len > 1)
        printf("update warp 5: per block  hand-over %lld  DMMA %lld  wait for B %lld;  per block row %lld cycles\n", t_serial / ((len - 1) * NB),
               t_dmma / ((len - 1) * NB), t_waitb / ((len - 1) * NB), (clock64() - t_step0) / (len - 1));
#endif
    // ---- condensed row of the slab: A_next = P part, B_next = Q part of the surviving rows, in ascending row order ----
    if (sm.pv[row] < 0) {
        // before the first elimination tmap of a carried row is -1: the survivors are rows 0..63 in order
        const int t = len > 1 ? sm.tmap[row] : row;
#pragma unroll
        for (int tc = 0; tc < MT; ++tc) {
            *reinterpret_cast<double2 *>(L.A_next + (slab * NV + t) * NV + 8 * tc + 2 * q) = X[MT + tc];
            *reinterpret_cast<double2 *>(L.B_next + (slab * NV + t) * NV + 8 * tc + 2 * q) = X[2 * MT + tc];
        }
    }
}

}  // namespace rst
