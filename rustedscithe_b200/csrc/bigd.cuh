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
// Where the time goes (clock64 instrumentation behind BIGD_TIMING, profiles/r02_summary.md): per block of 4 pivots the update
// warps need ~320 cycles for the hand-over (post pivot rows, look-ahead tile, post panel) and 1 150 - 3 700 cycles for their
// DMMAs (the B-fragment LDS -> DMMA pairs cannot be prefetched: 96 registers = the accumulator tiles), the panel warps
// ~3 600 - 4 800 cycles for the 4 pivots: their chain of ~8 dependent FP64 operations per pivot (reciprocal, multiplier,
// update of the next column) queues behind the update warps' DMMAs in the one FP64 pipe of each scheduler.  A 16-warp
// hybrid (bigh_factor_kernel: warps 0-3 own tiles AND factor the panel, 128 registers, no spills) measures the same
// (51.2 vs 49.8 ms): the panel chain, not the registers, is the limit.
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
__device__ __forceinline__ void bigd_panel_block(BigdSmem &sm, const AbdLevel &L, int64_t g, int b, int pw, int lane, bool &cand,
                                                 unsigned &bad) {
    constexpr int NV = BigdCfg::NV, R2 = BigdCfg::R2, PAN = 128;
    const int myrow = 32 * pw + lane;
    double a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] = sm.pan[(c >> 1) | ((c & 1) << 1)][myrow];
    const bool cand0 = cand;
    int mypiv = 4;
    double n10 = 0.0, n20 = 0.0, n21 = 0.0, n30 = 0.0, n31 = 0.0, n32 = 0.0;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        // key = (high word of |a| with its low 7 bits replaced by 127 - row) + 1: threshold pivoting (2^-13
        // relative), ties to the lowest row.  Every warp publishes its best row speculatively: ONE barrier per pivot.
        const unsigned hw = (unsigned)__double2hiint(a[ii]) & 0x7fffffffu;
        const unsigned key = cand ? (((hw & ~127u) | (unsigned)(127 - myrow)) + 1u) : 0u;
        const double myr = fast_rcp(a[ii]);
        const unsigned mxw = __reduce_max_sync(RST_FULL_MASK, key);
        if (lane == 31 - (int)((mxw - 1u) & 31u)) {
            double *sl = &sm.slot[ii & 1][pw][0];
            *reinterpret_cast<double2 *>(sl) = make_double2(ii == 0 ? myr : a[0], ii == 1 ? myr : a[1]);
            *reinterpret_cast<double2 *>(sl + 2) = make_double2(ii == 2 ? myr : a[2], ii == 3 ? myr : a[3]);
            sm.slot_key[ii & 1][pw] = mxw;
        }
        bar_named(3, PAN);
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
        const double rinv = ap[ii];
        if (ii == 1) { n10 = ap[0]; }
        if (ii == 2) { n20 = ap[0]; n21 = ap[1]; }
        if (ii == 3) { n30 = ap[0]; n31 = ap[1]; n32 = ap[2]; }
        const bool is_piv = myrow == prow;
        const bool upd = cand && !is_piv;
        const double l = upd ? a[ii] * rinv : 0.0;
        if (upd) a[ii] = l;
#pragma unroll
        for (int c = ii + 1; c < 4; ++c) a[c] = fma(-l, ap[c], a[c]);
        if (is_piv) { mypiv = ii; cand = false; }
    }
    const double i10 = -n10;
    const double i21 = -n21, i20 = -fma(n21, i10, n20);
    const double i32 = -n32, i31 = -fma(n32, i21, n31), i30 = -fma(n32, i20, fma(n31, i10, n30));
    const int lim = cand0 ? mypiv : 0;   // pivot ii keeps the multipliers of the earlier pivots
    const double l0 = lim > 0 ? a[0] : 0.0, l1 = lim > 1 ? a[1] : 0.0, l2 = lim > 2 ? a[2] : 0.0, l3 = lim > 3 ? a[3] : 0.0;
    // sequential multipliers: what the sweeps apply (layout of big.cuh)
    double *lm = L.Lm + g * (int64_t)(NV * R2) + (int64_t)(4 * b) * R2 + myrow;
    st_cs(lm, l0); st_cs(lm + R2, l1); st_cs(lm + 2 * R2, l2); st_cs(lm + 3 * R2, l3);
    // Lt = L L11^-1: what the tensor-core update applies to the raw pivot rows
    const double t3 = l3;
    const double t2 = fma(l3, i32, l2);
    const double t1 = fma(l3, i31, fma(l2, i21, l1));
    const double t0 = fma(l3, i30, fma(l2, i20, fma(l1, i10, l0)));
    *reinterpret_cast<double2 *>(&sm.lb[0][2 * myrow]) = make_double2(-t0, -t1);
    *reinterpret_cast<double2 *>(&sm.lb[1][2 * myrow]) = make_double2(-t2, -t3);
    sm.wh[myrow] = (unsigned char)mypiv;
    if (mypiv < 4) sm.pv[myrow] = (signed char)(4 * b + mypiv);
}

// 16 update warps (accumulator tiles) + 4 panel warps (one panel row per lane).  Five warps per scheduler cap the kernel
// at 96 registers per thread -- exactly the 24 tiles of an update warp; the panel warps need 8.
template <bool IDENT_B>
__global__ void __launch_bounds__(BigdCfg::THREADS + 128, 1) bigd_factor_kernel(AbdLevel L, unsigned *flags) {
    using C = BigdCfg;
    constexpr int NV = C::NV, R2 = C::R2, CT = C::CT, NB = C::NB;
    constexpr int MT = NV / 8;   // column tiles per part: M = tiles 0..7, P = 8..15, Q = 16..23
    constexpr int ALL = C::THREADS + 128, UPD = C::THREADS, PAN = 128;   // barriers 1 / 2 / 3
    extern __shared__ __align__(16) unsigned char bigd_smem_raw[];
    BigdSmem &sm = *reinterpret_cast<BigdSmem *>(bigd_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);

    if (warp >= C::WARPS) {
        // ================= panel warps: factor the 128 x 4 pivot-column panel of every block, one row per lane =================
        const int pw = warp - C::WARPS;
        unsigned bad = 0u;
        for (int istep = 1; istep < len; ++istep) {
            const int64_t g = g0 + istep;
            bool cand = true;               // my row has not been pivoted in this block row
#pragma unroll 1
            for (int b = 0; b < NB; ++b) {
#ifdef BIGD_TIMING
                const long long tp0 = clock64();
#endif
                bar_named(1, ALL);          // A(b): the pivot-column panel of block b is in shared memory
#ifdef BIGD_TIMING
                const long long tp1 = clock64();
#endif
                bigd_panel_block(sm, L, g, b, pw, lane, cand, bad);
#ifdef BIGD_TIMING
                if (blockIdx.x == 0 && pw == 0 && lane == 0 && istep == 2) printf("panel b=%d waitA %lld LU %lld\n", b, tp1 - tp0, clock64() - tp1);
#endif
                bar_named(1, ALL);          // B(b): Lt and the pivot identities of block b are published
            }
        }
        if (bad && lane == 0) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
        return;
    }

    // ================= update warps: warp w owns panel rows 8w .. 8w+7 as 24 accumulator tiles =================
    const int gid = lane >> 2, q = lane & 3;
    const bool hi = q >= 2;
    const int row = 8 * warp + gid;                 // my panel row in accumulator layout
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
            {
                const int which = sm.wh[row];
                if (which < 4) {
                    double *dst = &sm.xs[which][2 * q];
#pragma unroll
                    for (int tc = 0; tc < CT; ++tc)
                        if (tc >= TCB) *reinterpret_cast<double2 *>(dst + 8 * tc) = X[tc];
                }
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
#pragma unroll
            for (int tc = 0; tc < CT; ++tc)
                if (tc >= TCB && !(more && tc == TN)) dmma_m8n8k4(X[tc], af, bsrc[8 * tc]);
#ifdef BIGD_TIMING
            const long long tu2 = clock64();
#endif
            if (more) bar_named(1, ALL);   // B(b + 1)
#ifdef BIGD_TIMING
            if (blockIdx.x == 0 && tid == 32 * 5 && istep == 2) printf("upd b=%d serial %lld dmma %lld waitB %lld\n", b, tu1 - tu0, tu2 - tu1, clock64() - tu2);
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
        {
            double *uef = L.UEF + g * (int64_t)(3 * NV * NV);
#pragma unroll 4
            for (int m = 0; m < 3 * NV * NV / C::THREADS; ++m) {
                const int idx = tid + C::THREADS * m;      // = c * 64 + k
                st_cs(uef + idx, sm.ust[idx & (NV - 1)][idx >> 6]);
            }
        }
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

// Hybrid variant: 16 warps only (4 per scheduler -> 128 registers per thread: the 96 tile registers + room for the panel
// entries and for B-fragment prefetch, no spills).  Warps 0-3 own accumulator tiles like the others AND factor the pivot-column
// panel of the next block (one panel row per lane) right after the look-ahead tile is posted; the other twelve warps apply
// their tiles of the current block meanwhile, warps 0-3 apply theirs afterwards.
template <bool IDENT_B>
__global__ void __launch_bounds__(BigdCfg::THREADS, 1) bigh_factor_kernel(AbdLevel L, unsigned *flags) {
    using C = BigdCfg;
    constexpr int NV = C::NV, R2 = C::R2, CT = C::CT, NB = C::NB;
    constexpr int MT = NV / 8;
    constexpr int UPD = C::THREADS;
    extern __shared__ __align__(16) unsigned char bigd_smem_raw[];
    BigdSmem &sm = *reinterpret_cast<BigdSmem *>(bigd_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_pan = warp < 4;
    unsigned bad = 0u;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);

    // ================= update warps: warp w owns panel rows 8w .. 8w+7 as 24 accumulator tiles =================
    const int gid = lane >> 2, q = lane & 3;
    const bool hi = q >= 2;
    const int row = 8 * warp + gid;                 // my panel row in accumulator layout
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
        bool cand = true;    // panel duty (warps 0-3): my panel row 32 warp + lane has not been pivoted in this block row
        post_panel(X[0], false);
        bar_named(2, UPD);   // A(0)
        if (is_pan) bigd_panel_block(sm, L, g, 0, warp, lane, cand, bad);
        bar_named(2, UPD);   // B(0)

#pragma unroll
        for (int b = 0; b < NB; ++b) {            // fully unrolled: every tile index below is a compile-time constant
            const int TCB = b >> 1;
            const int TN = (b + 1) >> 1;           // column tile that holds the pivot columns of block b + 1
            const bool more = b + 1 < NB;
            // ---- raw rows of the block's 4 pivots -> shared memory (B fragments) ---------------------------------------
            {
                const int which = sm.wh[row];
                if (which < 4) {
                    double *dst = &sm.xs[which][2 * q];
#pragma unroll
                    for (int tc = 0; tc < CT; ++tc)
                        if (tc >= TCB) *reinterpret_cast<double2 *>(dst + 8 * tc) = X[tc];
                }
            }
            const double af = sm.lb[q >> 1][2 * row + (q & 1)];
            bar_named(2, UPD);
            const double *bsrc = &sm.xs[q][gid];
            // ---- look-ahead: update the tile with the next block's pivot columns first and hand them to the panel warps ----
            if (more) {
                dmma_m8n8k4(X[TN], af, bsrc[8 * TN]);
                post_panel(X[TN], ((b + 1) & 1) != 0);
                bar_named(2, UPD);   // A(b + 1)
                // warps 0-3 factor block b + 1 now (their own remaining tiles of block b follow); the other twelve warps keep the
                // FP64 pipe busy with theirs meanwhile
                if (is_pan) bigd_panel_block(sm, L, g, b + 1, warp, lane, cand, bad);
            }
            // ---- X <- X - Lt X[pivot rows] on the remaining tiles ------------------------------------------------------------
#pragma unroll
            for (int tc = 0; tc < CT; ++tc)
                if (tc >= TCB && !(more && tc == TN)) dmma_m8n8k4(X[tc], af, bsrc[8 * tc]);
            if (more) bar_named(2, UPD);   // B(b + 1)
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
        {
            double *uef = L.UEF + g * (int64_t)(3 * NV * NV);
#pragma unroll 4
            for (int m = 0; m < 3 * NV * NV / C::THREADS; ++m) {
                const int idx = tid + C::THREADS * m;      // = c * 64 + k
                st_cs(uef + idx, sm.ust[idx & (NV - 1)][idx >> 6]);
            }
        }
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
    if (is_pan && bad && lane == 0) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
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
