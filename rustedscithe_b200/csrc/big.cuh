// big.cuh -- "block per slab" kernels of the pivoted condensation solver for wide blocks (16 < nv <= 64, config C5).
//
// Same algorithm and the SAME factor layout as abd.cuh (Lm [g][k][row], UEF [g][c][k], piv [g][row]); what changes is
// the mapping: the stacked 2nv x 3nv panel [M | P | Q] of one slab (128 x 192 doubles = 192 KB at nv = 64) lives in the
// shared memory of ONE CTA, nv is a run-time value, every elimination step is a block-wide arg-max followed by a
// rank-1 update of the remaining panel.  This is the first correct wide-block path (north_star config 5); it is
// shared-memory-bandwidth bound (3 smem accesses per FMA) -- register tiling and FP64 DMMA for the rank-nv update are the
// next step.
#pragma once
#include "abd.cuh"

namespace rst {

constexpr int BIG_THREADS = 256;
constexpr int BIG_MAX_NV = 64;

__host__ __device__ inline size_t big_factor_smem_bytes(int nv) {
    const int R2 = 2 * nv, WP = 3 * nv + 1;
    return (size_t)R2 * WP * sizeof(double) + (size_t)R2 * sizeof(double) + (size_t)(3 * R2 + 64) * sizeof(int);
}

// block-wide arg-max over `val[r]`, r < R2 (val < 0 = not a candidate); result (row) returned to every thread
__device__ __forceinline__ int block_argmax(double v, int r, double *s_val, int *s_idx) {
    int idx = r;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_xor_sync(RST_FULL_MASK, v, off);
        const int oi = __shfl_xor_sync(RST_FULL_MASK, idx, off);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) { s_val[warp] = v; s_idx[warp] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bv = s_val[0];
        int bi = s_idx[0];
        for (int w = 1; w < nw; ++w)
            if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
        s_idx[32] = bi;
        s_val[32] = bv;
    }
    __syncthreads();
    return s_idx[32];
}

template <bool IDENT_B>
__global__ void __launch_bounds__(BIG_THREADS) big_factor_kernel(AbdLevel L, int nv, unsigned *flags) {
    extern __shared__ __align__(16) unsigned char big_smem[];
    const int R2 = 2 * nv, W = 3 * nv, WP = W + 1;
    double *panel = reinterpret_cast<double *>(big_smem);            // [R2][WP]
    double *lcol = panel + (size_t)R2 * WP;                           // [R2] multipliers of the current pivot
    int *pv = reinterpret_cast<int *>(lcol + R2);                     // [R2] pivot step of a row, -1 = survivor
    int *carry = pv + R2;                                             // [R2] 1 = row carries the condensed row
    int *slot_of = carry + R2;                                        // [nv] row slot that receives new block row t
    __shared__ double s_val[33];
    __shared__ int s_idx[33];
    const int tid = threadIdx.x;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);

    for (int idx = tid; idx < R2 * WP; idx += blockDim.x) panel[idx] = 0.0;
    for (int r = tid; r < R2; r += blockDim.x) carry[r] = r < nv ? 1 : 0;
    __syncthreads();
    for (int idx = tid; idx < nv * nv; idx += blockDim.x) {
        const int r = idx / nv, c = idx - r * nv;
        panel[r * WP + nv + c] = L.A[g0 * nv * nv + idx];                                           // P = A
        panel[r * WP + c] = IDENT_B ? (r == c ? 1.0 : 0.0) : L.B[g0 * nv * nv + idx];               // M = B
    }
    __syncthreads();

    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        if (tid == 0) {
            int t = 0;
            for (int r = 0; r < R2; ++r)
                if (!carry[r]) slot_of[t++] = r;
        }
        for (int r = tid; r < R2; r += blockDim.x) pv[r] = -1;
        __syncthreads();
        for (int idx = tid; idx < nv * nv; idx += blockDim.x) {
            const int t = idx / nv, c = idx - t * nv;
            const int r = slot_of[t];
            panel[r * WP + c] = L.A[g * nv * nv + idx];                                             // M = A
            panel[r * WP + nv + c] = 0.0;                                                           // P = 0
            panel[r * WP + 2 * nv + c] = IDENT_B ? (t == c ? 1.0 : 0.0) : L.B[g * nv * nv + idx];   // Q = B
        }
        __syncthreads();
        for (int k = 0; k < nv; ++k) {
            const double v = (tid < R2 && pv[tid] < 0) ? fabs(panel[tid * WP + k]) : -1.0;
            const int p = block_argmax(v, tid < R2 ? tid : R2, s_val, s_idx);
            const double pivval = panel[p * WP + k];
            if (tid == 0 && !(fabs(pivval) > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
            if (tid < R2) {
                double l = 0.0;
                if (tid != p && pv[tid] < 0) {
                    l = panel[tid * WP + k] / pivval;
                    panel[tid * WP + k] = l;
                }
                lcol[tid] = l;
            }
            __syncthreads();
            if (tid == 0) pv[p] = k;
            const int ncol = W - k - 1;
            for (int idx = tid; idx < R2 * ncol; idx += blockDim.x) {
                const int r = idx / ncol, c = k + 1 + (idx - r * ncol);
                const double l = lcol[r];
                if (l != 0.0) panel[r * WP + c] = fma(-l, panel[p * WP + c], panel[r * WP + c]);
            }
            __syncthreads();
        }
        // factors of this step (same layout as abd.cuh)
        double *lm = L.Lm + g * (int64_t)(nv * R2);
        for (int idx = tid; idx < nv * R2; idx += blockDim.x) {
            const int k = idx / R2, r = idx - k * R2;
            st_cs(lm + idx, panel[r * WP + k]);
        }
        for (int r = tid; r < R2; r += blockDim.x) L.piv[g * R2 + r] = (int8_t)pv[r];
        double *uef = L.UEF + g * (int64_t)(3 * nv * nv);
        for (int idx = tid; idx < R2 * W; idx += blockDim.x) {
            const int r = idx / W, c = idx - r * W;
            const int k = pv[r];
            if (k >= 0) st_cs(uef + (int64_t)c * nv + k, panel[r * WP + c]);
        }
        __syncthreads();
        // survivors carry the condensed row: M <- Q, Q <- 0
        for (int idx = tid; idx < R2 * nv; idx += blockDim.x) {
            const int r = idx / nv, c = idx - r * nv;
            if (pv[r] < 0) {
                panel[r * WP + c] = panel[r * WP + 2 * nv + c];
                panel[r * WP + 2 * nv + c] = 0.0;
            }
        }
        for (int r = tid; r < R2; r += blockDim.x) carry[r] = pv[r] < 0 ? 1 : 0;
        __syncthreads();
    }
    if (tid == 0) {
        int t = 0;
        for (int r = 0; r < R2; ++r)
            if (carry[r]) slot_of[t++] = r;
    }
    __syncthreads();
    for (int idx = tid; idx < nv * nv; idx += blockDim.x) {
        const int t = idx / nv, c = idx - t * nv;
        const int r = slot_of[t];
        L.A_next[slab * nv * nv + idx] = panel[r * WP + nv + c];
        L.B_next[slab * nv * nv + idx] = panel[r * WP + c];
    }
}

// ------------------------------------------------------------------------------------------------
// Register-tiled factor kernel (generation 2 of the wide path).
// The stacked panel [M | P | Q] lives in REGISTERS: 16 warps, warp `ty` owns rows r = ty + 16 i (i < RT = NVP/8), lane
// `tx` owns the padded columns tx + 32 j (j < CT = 3 NVP/32); block b of the panel starts at padded column b NVP, so the
// M <- Q rotation of a carried row stays inside a lane.  One elimination step costs ONE block barrier:
//   * the lane that owns column k finds its warp's best candidate row, every warp publishes that row (padded 3 NVP
//     doubles) and |pivot| in a parity-double-buffered shared slot;
//   * after the barrier every warp reduces the 16 candidates, reads the winning row from the winner's slot and applies
//     the rank-1 update to its 8 x 6 register tile (predicated on column > k; rows with multiplier 0 are skipped).
// Multipliers stay in the eliminated column of the tile.  Pivot rows and the multiplier tile are staged in shared memory
// and written with coalesced stores in the layout of abd.cuh (Lm [g][k][row], UEF [g][c][k]), so the sweeps and the
// narrow-block kernels read the same factors.  FP64-pipe bound: 48 DFMA per thread and step.
// ------------------------------------------------------------------------------------------------
// NW = 16: 8 x 6 tile per thread, pivot rows staged in smem for a coalesced UEF store.
// NW = 32: 4 x 6 tile per thread (64 registers, twice the warps to hide the latency of the per-step chain), pivot rows
//          stored straight to global (the L2 merges the four 8-byte words of a sector written by consecutive steps).
template <int NVP, int NW_>
struct BigrCfg {
    static constexpr int WARPS = NW_, THREADS = 32 * NW_;
    static constexpr int RT = 2 * NVP / WARPS, CT = 3 * NVP / 32, JB = NVP / 32;
    static constexpr bool DIRECT_U = NW_ > 16;
    static constexpr int SLOTW = 3 * NVP;                       // padded panel width
    static constexpr int PU = NVP + 1;                          // stage_u[c_padded][k] pitch
    static constexpr int PL = 2 * NVP + 1;                      // stage_l[k][row] pitch
    static constexpr int slot_doubles = 2 * WARPS * SLOTW;
    static constexpr int stage_u_doubles = DIRECT_U ? 0 : SLOTW * PU;
    static constexpr int stage_l_doubles = NVP * PL;
    static constexpr size_t smem_bytes = (size_t)(slot_doubles + stage_u_doubles + stage_l_doubles + 2 * WARPS) * sizeof(double) +
                                         (2 * WARPS + 2 * NVP + 8) * sizeof(int) + 2 * NVP;
};

template <int NVP, int NW_, bool IDENT_B>
__global__ void __launch_bounds__(32 * NW_, 1) bigr_factor_kernel(AbdLevel L, int nv, unsigned *flags) {
    using C = BigrCfg<NVP, NW_>;
    static_assert(C::RT >= 1 && C::RT <= 8, "rows per thread");
    constexpr int RT = C::RT, CT = C::CT, JB = C::JB, SLOTW = C::SLOTW, PU = C::PU, PL = C::PL, NW = C::WARPS;
    extern __shared__ __align__(16) unsigned char big_smem[];
    double *slot = reinterpret_cast<double *>(big_smem);                 // [2][16][SLOTW] candidate rows
    double *stage_u = slot + C::slot_doubles;                            // [SLOTW][PU] pivot rows of this block row
    double *stage_l = stage_u + C::stage_u_doubles;                      // [NVP][PL]  multiplier columns of this block row
    double *part_val = stage_l + C::stage_l_doubles;                     // [2][16] signed candidate pivot
    int *part_key = reinterpret_cast<int *>(part_val + 2 * NW);          // [2][16] (|pivot| high word, warp, row) keys
    int *tmap = part_key + 2 * NW;                                       // [2 NVP] row -> index of the new block row / carried row
    unsigned *masks = reinterpret_cast<unsigned *>(tmap + 2 * NVP);      // [8] ballot words
    unsigned char *s_pv = reinterpret_cast<unsigned char *>(masks + 8);  // [2 NVP] row state

    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int R2 = 2 * nv;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    constexpr int FREE = 64, DEAD = 0x7F, ACTIVE = 0xFF;

    double a[RT][CT];
    int st[RT];         // ACTIVE = in the panel and not pivoted in this block row, 0..63 = pivoted at that k, FREE / DEAD
    unsigned act = 0;   // bit i: st[i] == ACTIVE
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int r = ty + NW * i;
#pragma unroll
        for (int j = 0; j < CT; ++j) a[i][j] = 0.0;
        st[i] = r < nv ? ACTIVE : (r < R2 ? 0 : DEAD);
        if (r < nv) {
            act |= 1u << i;
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const int c = tx + 32 * j;
                if (c < nv) {
                    a[i][j] = IDENT_B ? (r == c ? 1.0 : 0.0) : L.B[(g0 * nv + r) * nv + c];   // M = B
                    a[i][JB + j] = L.A[(g0 * nv + r) * nv + c];                               // P = A
                }
            }
        }
    }
    // keys: |pivot| high word with the low 7 bits replaced by (15 - warp, 7 - row) -> ties go to the lowest row; the
    // pivot is the largest candidate up to 2^-13 relative, i.e. threshold pivoting with a threshold of 0.9999
    const int keytag = ((NW - 1 - ty) << 3);
    constexpr int KEYMASK = NW > 16 ? 0x7fffff00 : 0x7fffff80;

    for (int istep = 1; istep < len; ++istep) {
        const int64_t g = g0 + istep;
        // ---- new block row t goes into the t-th free row (ascending), as the sweeps assume ----
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < RT; ++i) s_pv[ty + NW * i] = (unsigned char)st[i];
        }
        __syncthreads();
        if (tid < 2 * NVP) {
            const unsigned m = __ballot_sync(RST_FULL_MASK, s_pv[tid] < FREE);
            if (tx == 0) masks[ty] = m;
        }
        __syncthreads();
        if (tid < 2 * NVP) {
            int t = __popc(masks[ty] & ((1u << tx) - 1u));
            for (int w = 0; w < ty; ++w) t += __popc(masks[w]);
            tmap[tid] = s_pv[tid] < FREE ? t : -1;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            const int t = tmap[ty + NW * i];
            if (t >= 0) {
                st[i] = ACTIVE;
                act |= 1u << i;
#pragma unroll
                for (int j = 0; j < JB; ++j) {
                    const int c = tx + 32 * j;
                    const bool in = c < nv;
                    a[i][j] = in ? L.A[(g * nv + t) * nv + c] : 0.0;                                          // M = A
                    a[i][JB + j] = 0.0;                                                                       // P = 0
                    a[i][2 * JB + j] = in ? (IDENT_B ? (t == c ? 1.0 : 0.0) : L.B[(g * nv + t) * nv + c]) : 0.0;  // Q = B
                }
            }
        }
        // ---- nv elimination steps, one block barrier each ----
        // Rows that are not candidates (pivoted earlier in this block row, padding) hold ZEROS in registers: their pivot
        // key is minimal and their multiplier is 0, so neither the pivot search nor the update needs a row mask.
#pragma unroll
        for (int jk = 0; jk < JB; ++jk) {
            for (int kk = 0; kk < 32; ++kk) {
                const int k = 32 * jk + kk;
                if (k >= nv) break;
                const int par = k & 1;
                // warp-local candidate of column k (held by lane kk in register column jk)
                int key = -1;
#pragma unroll
                for (int i = 0; i < RT; ++i) key = max(key, ((__double2hiint(a[i][jk]) & KEYMASK) | keytag) + (7 - i));
                key = __shfl_sync(RST_FULL_MASK, key, kk);
                // column k of this warp's rows, broadcast to every lane (becomes the multipliers after the barrier)
                double l[RT];
#pragma unroll
                for (int i = 0; i < RT; ++i) l[i] = __shfl_sync(RST_FULL_MASK, a[i][jk], kk);
                const int bi = 7 - (key & 7);
                double *myslot = slot + (par * NW + ty) * SLOTW + tx;
                double cand = 0.0;
                switch (bi) {
#define BIGR_PUBLISH(I)                                                                             \
    case I:                                                                                         \
        if (I < RT) {                                                                               \
            _Pragma("unroll") for (int j = jk; j < CT; ++j) myslot[32 * j] = a[I < RT ? I : 0][j];  \
            cand = l[I < RT ? I : 0];                                                               \
        }                                                                                           \
        break;
                    BIGR_PUBLISH(0) BIGR_PUBLISH(1) BIGR_PUBLISH(2) BIGR_PUBLISH(3)
                    BIGR_PUBLISH(4) BIGR_PUBLISH(5) BIGR_PUBLISH(6) BIGR_PUBLISH(7)
#undef BIGR_PUBLISH
                }
                if (tx == kk) {
                    part_key[par * NW + ty] = key;
                    // reciprocal of the candidate pivot; a zero candidate (sparse column: only zeros in this warp's rows) must not
                    // enter the division's slow path
                    part_val[par * NW + ty] = cand != 0.0 ? 1.0 / cand : __longlong_as_double(0x7ff0000000000000ll);
                }
                __syncthreads();
                // winner over the 16 warps
                const int wkey = __reduce_max_sync(RST_FULL_MASK, part_key[par * NW + (tx & (NW - 1))]);
                const int wp = NW - 1 - ((wkey >> 3) & (NW - 1)), ip = 7 - (wkey & 7);
                const double rp = part_val[par * NW + wp];
                const double *prow = slot + (par * NW + wp) * SLOTW + tx;
                if (ty == wp) {
                    // the winning row leaves the panel: record it, stage it for the UEF store, clear its registers
                    if (tx == 0 && !(fabs(rp) <= 1.79769313486231570e308)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
                    // pivot row k -> UEF [c][k]: staged for a coalesced store, or straight to global (DIRECT_U)
                    double *su = C::DIRECT_U ? L.UEF + g * (int64_t)(3 * nv * nv) + k : stage_u + tx * PU + k;
                    switch (ip) {
#define BIGR_RETIRE(I)                                                                              \
    case I:                                                                                         \
        if (I < RT) {                                                                               \
            st[I < RT ? I : 0] = k;                                                                 \
            l[I < RT ? I : 0] = 0.0;                                                                \
            _Pragma("unroll") for (int j = jk; j < CT; ++j) {                                       \
                if (C::DIRECT_U) {                                                                  \
                    const int cc = tx + 32 * (j % JB);                                              \
                    if (cc < nv) su[((j / JB) * nv + cc) * nv] = a[I < RT ? I : 0][j];              \
                } else {                                                                            \
                    su[32 * j * PU] = a[I < RT ? I : 0][j];                                         \
                }                                                                                   \
            }                                                                                       \
            _Pragma("unroll") for (int j = 0; j < CT; ++j) a[I < RT ? I : 0][j] = 0.0;              \
        }                                                                                           \
        break;
                        BIGR_RETIRE(0) BIGR_RETIRE(1) BIGR_RETIRE(2) BIGR_RETIRE(3)
                        BIGR_RETIRE(4) BIGR_RETIRE(5) BIGR_RETIRE(6) BIGR_RETIRE(7)
#undef BIGR_RETIRE
                    }
                    act &= ~(1u << ip);
                }
#pragma unroll
                for (int i = 0; i < RT; ++i) l[i] *= rp;
                if (tx == kk) {   // multiplier column k -> Lm staging
                    double *lcol = stage_l + k * PL + ty;
#pragma unroll
                    for (int i = 0; i < RT; ++i) lcol[NW * i] = l[i];
                }
#pragma unroll
                for (int j = jk; j < CT; ++j) {
                    double pr = prow[32 * j];
                    if (j == jk && tx <= kk) pr = 0.0;   // columns <= k of this block are finished
#pragma unroll
                    for (int i = 0; i < RT; ++i) a[i][j] = fma(-l[i], pr, a[i][j]);
                }
            }
        }
        __syncthreads();
        // ---- factors of this block row (layout of abd.cuh): UEF [c][k], Lm [k][row], piv ----
        {
            if (!C::DIRECT_U) {
                double *uef = L.UEF + g * (int64_t)(3 * nv * nv);
                for (int b = 0; b < 3; ++b)
                    for (int cc = ty; cc < nv; cc += NW) {
                        const double *src = stage_u + (b * NVP + cc) * PU;
                        double *dst = uef + (b * nv + cc) * nv;
                        for (int k = tx; k < nv; k += 32) st_cs(dst + k, src[k]);
                    }
            }
            double *lm = L.Lm + g * (int64_t)(nv * R2);
            for (int k = ty; k < nv; k += NW)
                for (int r = tx; r < R2; r += 32) st_cs(lm + k * R2 + r, stage_l[k * PL + r]);
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const int r = ty + NW * i;
                if (r < R2) L.piv[g * R2 + r] = (int8_t)st[i];
            }
        }
        // ---- survivors carry the condensed row: M <- Q, Q <- 0 ----
#pragma unroll
        for (int i = 0; i < RT; ++i) {
            if ((act >> i) & 1u) {
#pragma unroll
                for (int j = 0; j < JB; ++j) {
                    a[i][j] = a[i][2 * JB + j];
                    a[i][2 * JB + j] = 0.0;
                }
            }
        }
        __syncthreads();   // staging buffers are reused by the next block row
    }
    // ---- condensed row of the slab: carried rows in ascending order ----
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < RT; ++i) s_pv[ty + NW * i] = (unsigned char)st[i];
    }
    __syncthreads();
    if (tid < 2 * NVP) {
        const unsigned m = __ballot_sync(RST_FULL_MASK, s_pv[tid] == ACTIVE);
        if (tx == 0) masks[ty] = m;
    }
    __syncthreads();
    if (tid < 2 * NVP) {
        int t = __popc(masks[ty] & ((1u << tx) - 1u));
        for (int w = 0; w < ty; ++w) t += __popc(masks[w]);
        tmap[tid] = s_pv[tid] == ACTIVE ? t : -1;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int t = tmap[ty + NW * i];
        if (t >= 0) {
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const int c = tx + 32 * j;
                if (c < nv) {
                    L.A_next[(slab * nv + t) * nv + c] = a[i][JB + j];
                    L.B_next[(slab * nv + t) * nv + c] = a[i][j];
                }
            }
        }
    }
}

// forward sweep: apply the recorded eliminations to the right-hand side (one CTA per slab)
__global__ void __launch_bounds__(128) big_forward_kernel(AbdLevel L, int nv) {
    __shared__ double w[2 * BIG_MAX_NV];
    __shared__ int pv[2 * BIG_MAX_NV], carry[2 * BIG_MAX_NV], slot_of[BIG_MAX_NV], row_of[BIG_MAX_NV];
    const int R2 = 2 * nv;
    const int tid = threadIdx.x;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    if (tid < R2) {
        carry[tid] = tid < nv ? 1 : 0;
        w[tid] = tid < nv ? L.r[g0 * nv + tid] : 0.0;
    }
    __syncthreads();
    for (int i = 1; i < len; ++i) {
        const int64_t g = g0 + i;
        if (tid == 0) {
            int t = 0;
            for (int r = 0; r < R2; ++r)
                if (!carry[r]) slot_of[t++] = r;
        }
        if (tid < R2) {
            pv[tid] = (int)L.piv[g * R2 + tid];
            if (pv[tid] >= 0) row_of[pv[tid]] = tid;
        }
        __syncthreads();
        if (tid < nv) w[slot_of[tid]] = L.r[g * nv + tid];
        __syncthreads();
        const double *lm = L.Lm + g * (int64_t)(nv * R2);
        for (int k = 0; k < nv; ++k) {
            const double wp = w[row_of[k]];
            __syncthreads();
            if (tid < R2 && (pv[tid] < 0 || pv[tid] > k)) w[tid] = fma(-ld_cs(lm + (int64_t)k * R2 + tid), wp, w[tid]);
            __syncthreads();
        }
        if (tid < R2) {
            if (pv[tid] >= 0) L.e[g * nv + pv[tid]] = w[tid];
            carry[tid] = pv[tid] < 0 ? 1 : 0;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int t = 0;
        for (int r = 0; r < R2; ++r)
            if (carry[r]) slot_of[t++] = r;
    }
    __syncthreads();
    if (tid < nv) L.r_next[slab * nv + tid] = w[slot_of[tid]];
}

// backward sweep: interior nodes of a slab from its end nodes (one CTA per slab, thread k = pivot row k)
__global__ void __launch_bounds__(64) big_backward_kernel(AbdLevel L, int nv) {
    __shared__ double Ys[BIG_MAX_NV], Yn[BIG_MAX_NV], ybc;
    const int k = threadIdx.x;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    if (k < nv) {
        Ys[k] = L.Z_next[slab * nv + k];
        Yn[k] = L.Z_next[(slab + 1) * nv + k];
        L.Z[g0 * nv + k] = Ys[k];
        if (slab == (int64_t)L.nslabs - 1) L.Z[L.n * nv + k] = Yn[k];
    }
    __syncthreads();
    for (int i = len - 1; i >= 1; --i) {
        const int64_t g = g0 + i;
        const double *u = L.UEF + g * (int64_t)(3 * nv * nv) + k;
        double t = 0.0, inv = 1.0;
        if (k < nv) {
            t = L.e[g * nv + k];
            for (int c = 0; c < nv; ++c) {
                t = fma(-ld_cs(u + (int64_t)(nv + c) * nv), Ys[c], t);
                t = fma(-ld_cs(u + (int64_t)(2 * nv + c) * nv), Yn[c], t);
            }
            inv = 1.0 / u[(int64_t)k * nv];
        }
        __syncthreads();  // every thread has finished reading the old Yn
        for (int kk = nv - 1; kk >= 0; --kk) {
            if (k == kk) { ybc = t * inv; Yn[kk] = ybc; }
            __syncthreads();
            if (k < kk) t = fma(-u[(int64_t)kk * nv], ybc, t);
            __syncthreads();
        }
        if (k < nv) L.Z[g * nv + k] = Yn[k];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Pipelined sweeps of the wide path (generation 2).  The factors of one block row (Lm: 16 nv^2 bytes, UEF: 24 nv^2 bytes;
// 64 KB + 96 KB at nv = 64) are streamed through a shared-memory ring with 1-D bulk async copies (cp.async.bulk +
// mbarrier complete_tx): one elected thread keeps NS stages in flight, the CTA consumes them in order.  The arithmetic
// is a dependent chain (nv steps per block row), so the ring is what keeps HBM busy: 3 CTAs per SM x 3 stages in flight.
// Requires 16-byte aligned stages: always true for Lm, true for UEF when 3 nv^2 is even (else the generation-1 kernels run).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();   // a lost copy must not hang the GPU
    }
}

constexpr int BIGP_KC = 8;       // elimination steps (forward) per stage
constexpr int BIGP_NS = 3;       // ring depth
constexpr int BIGP_THREADS = 128;
__host__ __device__ inline size_t bigp_forward_smem(int nv) { return (size_t)BIGP_NS * BIGP_KC * 2 * nv * sizeof(double); }

// forward sweep: w <- (eliminations of every block row of the slab) w; thread r owns panel row r
__global__ void __launch_bounds__(BIGP_THREADS) bigp_forward_kernel(AbdLevel L, int nv) {
    extern __shared__ __align__(128) unsigned char big_smem[];
    double *ring = reinterpret_cast<double *>(big_smem);   // [NS][KC][R2]
    __shared__ unsigned long long full[BIGP_NS];
    __shared__ double vbuf[2];
    __shared__ unsigned masks[4];
    const int R2 = 2 * nv;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    const int nh = (nv + BIGP_KC - 1) / BIGP_KC;                 // stages per block row
    const int total = (len - 1) * nh;
    const unsigned stage_doubles = BIGP_KC * R2;

    auto issue = [&](int sidx) {                                  // thread 0 only
        const int is = 1 + sidx / nh, h = sidx - (is - 1) * nh;
        const int kc = min(BIGP_KC, nv - h * BIGP_KC);
        const unsigned bytes = (unsigned)(kc * R2 * sizeof(double));
        const int slot = sidx % BIGP_NS;
        mbar_expect_tx(&full[slot], bytes);
        bulk_g2s(ring + (size_t)slot * stage_doubles, L.Lm + ((g0 + is) * (int64_t)nv + h * BIGP_KC) * R2, bytes, &full[slot]);
    };
    if (tid == 0) {
        for (int i = 0; i < BIGP_NS; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < BIGP_NS && i < total; ++i) issue(i);

    // row state: pvr = elimination step at which this row is the pivot (0..nv-1), 127 = carries on, 126 = padding
    double w = 0.0;
    int pvr = tid < nv ? 127 : (tid < R2 ? 0 : 126);              // before the first block row: rows nv..2nv-1 are free
    if (tid < nv) w = L.r[g0 * nv + tid];
    // rank of this row among the free rows = index of the new block row it receives
    auto free_rank = [&](bool fr) {
        const unsigned m = __ballot_sync(RST_FULL_MASK, fr);
        if (lane == 0) masks[wid] = m;
        __syncthreads();
        int t = __popc(m & ((1u << lane) - 1u));
        for (int q = 0; q < wid; ++q) t += __popc(masks[q]);
        __syncthreads();
        return t;
    };
    int t_new = free_rank(pvr < nv);
    double r_new = (len > 1 && pvr < nv) ? L.r[(g0 + 1) * nv + t_new] : 0.0;
    int pv_new = (len > 1 && tid < R2) ? (int)L.piv[(g0 + 1) * R2 + tid] : 126;

    int sidx = 0;
    for (int istep = 1; istep < len; ++istep) {
        const int64_t g = g0 + istep;
        if (pvr < nv) w = r_new;                                  // free row receives block row t_new of the rhs
        if (tid < R2) pvr = pv_new < 0 ? 127 : pv_new;
        // prefetch the next block row's pivots / rhs (the rows pivoted now are the free rows then)
        t_new = free_rank(pvr < nv);
        if (istep + 1 < len) {
            if (pvr < nv) r_new = L.r[(g + 1) * nv + t_new];
            if (tid < R2) pv_new = (int)L.piv[(g + 1) * R2 + tid];
        }
        for (int h = 0; h < nh; ++h, ++sidx) {
            const int slot = sidx % BIGP_NS;
            mbar_wait(&full[slot], (unsigned)((sidx / BIGP_NS) & 1));
            const double *tile = ring + (size_t)slot * stage_doubles + tid;
            const int k0 = h * BIGP_KC, kc = min(BIGP_KC, nv - k0);
            for (int kl = 0; kl < kc; ++kl) {
                const int k = k0 + kl;
                if (pvr == k) vbuf[k & 1] = w;
                __syncthreads();
                const double v = vbuf[k & 1];
                if (pvr > k && tid < R2) w = fma(-tile[kl * R2], v, w);
            }
            __syncthreads();                                      // every thread is done with this slot
            if (tid == 0 && sidx + BIGP_NS < total) issue(sidx + BIGP_NS);
        }
        if (pvr < nv) L.e[g * nv + pvr] = w;
    }
    const int t_c = free_rank(pvr == 127);
    if (pvr == 127) L.r_next[slab * nv + t_c] = w;
}

constexpr int BIGP_CC = 16;      // UEF rows (c) per stage in the backward sweep (must divide 32: a chunk's unknowns live in one warp)
__host__ __device__ inline size_t bigp_backward_smem(int nv) { return (size_t)BIGP_NS * BIGP_CC * nv * sizeof(double); }

// backward sweep: interior nodes of the slab from its end nodes.  Stage order per block row (descending):
// E chunks, F chunks (matrix-vector part, all 128 threads), then U chunks from the last column block down (back substitution,
// threads 0..nv-1 with a 64-thread named barrier per step).
__global__ void __launch_bounds__(BIGP_THREADS) bigp_backward_kernel(AbdLevel L, int nv) {
    extern __shared__ __align__(128) unsigned char big_smem[];
    double *ring = reinterpret_cast<double *>(big_smem);   // [NS][CC][nv]
    __shared__ unsigned long long full[BIGP_NS];
    __shared__ double Ys[BIG_MAX_NV], Yn[BIG_MAX_NV], part[BIG_MAX_NV], xb[2];
    const int tid = threadIdx.x;
    const int k = tid & 63, grp = tid >> 6;
    const int64_t slab = blockIdx.x;
    const int64_t g0 = slab * (int64_t)L.S;
    const int len = (int)((L.n - g0) < (int64_t)L.S ? (L.n - g0) : (int64_t)L.S);
    const int nch = (nv + BIGP_CC - 1) / BIGP_CC;                 // chunks per nv columns
    const int nst = 3 * nch;                                      // stages per block row: E.., F.., U.. (descending)
    const int total = (len - 1) * nst;
    const unsigned stage_doubles = BIGP_CC * nv;

    // stage sidx -> (block row, first UEF row c0, rows)
    auto stage_of = [&](int sidx, int &is, int &c0, int &cc) {
        const int q = sidx / nst, j = sidx - q * nst;
        is = len - 1 - q;
        if (j < 2 * nch) {                                        // E then F, ascending
            const int b = j / nch, ch = j - b * nch;
            c0 = (1 + b) * nv + ch * BIGP_CC;
            cc = min(BIGP_CC, nv - ch * BIGP_CC);
        } else {                                                  // U, descending chunks
            const int ch = nch - 1 - (j - 2 * nch);
            c0 = ch * BIGP_CC;
            cc = min(BIGP_CC, nv - ch * BIGP_CC);
        }
    };
    auto issue = [&](int sidx) {
        int is, c0, cc;
        stage_of(sidx, is, c0, cc);
        const unsigned bytes = (unsigned)(cc * nv * sizeof(double));
        const int slot = sidx % BIGP_NS;
        mbar_expect_tx(&full[slot], bytes);
        bulk_g2s(ring + (size_t)slot * stage_doubles, L.UEF + (g0 + is) * (int64_t)(3 * nv * nv) + (int64_t)c0 * nv, bytes, &full[slot]);
    };
    if (tid == 0) {
        for (int i = 0; i < BIGP_NS; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < BIGP_NS && i < total; ++i) issue(i);
    if (tid < nv) {
        Ys[tid] = L.Z_next[slab * nv + tid];
        Yn[tid] = L.Z_next[(slab + 1) * nv + tid];
        L.Z[g0 * nv + tid] = Ys[tid];
        if (slab == (int64_t)L.nslabs - 1) L.Z[L.n * nv + tid] = Yn[tid];
    }
    double e_next = (len > 1 && tid < nv) ? L.e[(g0 + len - 1) * nv + tid] : 0.0;
    __syncthreads();

    int sidx = 0;
    for (int istep = len - 1; istep >= 1; --istep) {
        const int64_t g = g0 + istep;
        const double e_cur = e_next;
        if (istep > 1 && tid < nv) e_next = L.e[(g - 1) * nv + tid];
        // ---- t_k = e_k - sum_c E[c][k] Ys[c] - sum_c F[c][k] Yn[c]; the two thread groups split the rows of a chunk ----
        double acc0 = 0.0, acc1 = 0.0;
        for (int j = 0; j < 2 * nch; ++j, ++sidx) {
            const int slot = sidx % BIGP_NS;
            mbar_wait(&full[slot], (unsigned)((sidx / BIGP_NS) & 1));
            const double *tile = ring + (size_t)slot * stage_doubles + k;
            const int b = j / nch, ch = j - b * nch;
            const int cc = min(BIGP_CC, nv - ch * BIGP_CC);
            const double *y = (b == 0 ? Ys : Yn) + ch * BIGP_CC;
            if (k < nv) {
                int c = grp;
                for (; c + 2 < cc; c += 4) {
                    acc0 = fma(tile[c * nv], y[c], acc0);
                    acc1 = fma(tile[(c + 2) * nv], y[c + 2], acc1);
                }
                for (; c < cc; c += 2) acc0 = fma(tile[c * nv], y[c], acc0);
            }
            __syncthreads();
            if (tid == 0 && sidx + BIGP_NS < total) issue(sidx + BIGP_NS);
        }
        if (grp == 1 && k < nv) part[k] = acc0 + acc1;
        __syncthreads();
        double t = 0.0;
        if (grp == 0 && k < nv) t = e_cur - ((acc0 + acc1) + part[k]);
        // ---- back substitution with U, last column block first ----
        for (int j = 0; j < nch; ++j, ++sidx) {
            const int slot = sidx % BIGP_NS;
            mbar_wait(&full[slot], (unsigned)((sidx / BIGP_NS) & 1));
            const int ch = nch - 1 - j;
            const int c0 = ch * BIGP_CC, cc = min(BIGP_CC, nv - c0);
            if (grp == 0) {
                const double *tile = ring + (size_t)slot * stage_doubles + k;
                // the reciprocal of this thread's diagonal entry is formed when its chunk arrives, off the serial chain below
                // (a division there costs more than the barrier + FMA of a substitution step)
                double myinv = 0.0;
                if (k >= c0 && k < c0 + cc) myinv = fast_rcp(tile[(k - c0) * nv]);
                // the chunk's <= 16 unknowns sit in one warp (BIGP_CC = 16 divides 32): their triangular solve runs on shuffles,
                // no barrier and no shared-memory round trip per unknown; the threads below the chunk apply the chunk's
                // unknowns afterwards in the same order (same rounding as updating step by step)
                const bool in_chunk = k >= c0 && k < c0 + cc;
                if ((k >> 5) == (c0 >> 5)) {
                    for (int cl = cc - 1; cl >= 0; --cl) {
                        const int kk = c0 + cl;
                        const double x = __shfl_sync(RST_FULL_MASK, t * myinv, kk & 31);
                        if (k == kk) Yn[kk] = x;
                        else if (in_chunk && k < kk) t = fma(-tile[cl * nv], x, t);
                    }
                }
                asm volatile("bar.sync 1, 64;" ::: "memory");
                if (k < c0)
                    for (int cl = cc - 1; cl >= 0; --cl) t = fma(-tile[cl * nv], Yn[c0 + cl], t);
            }
            __syncthreads();
            if (tid == 0 && sidx + BIGP_NS < total) issue(sidx + BIGP_NS);
        }
        if (tid < nv) L.Z[g * nv + tid] = Yn[tid];
    }
}

// closure with pinned boundary components, run-time nv (abd_top_solve_kernel with K sized for nv <= 64)
__global__ void __launch_bounds__(32) big_top_solve_kernel(const double *C, const double *D, const double *g, double *Z,
                                                           int nv, unsigned long long left_mask,
                                                           unsigned long long right_mask, unsigned *flags) {
    __shared__ double K[BIG_MAX_NV][BIG_MAX_NV + 1];
    __shared__ int colmap[BIG_MAX_NV];
    const int lane = threadIdx.x;
    if (lane == 0) {
        int nc = 0;
        for (int c = 0; c < nv; ++c)
            if (!((left_mask >> c) & 1ull)) { if (nc < nv) colmap[nc] = c; ++nc; }
        for (int c = 0; c < nv; ++c)
            if (!((right_mask >> c) & 1ull)) { if (nc < nv) colmap[nc] = nv + c; ++nc; }
        if (nc != nv) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
    }
    __syncwarp();
    for (int idx = lane; idx < nv * nv; idx += 32) {
        const int row = idx / nv, cc = idx % nv;
        const int cm = colmap[cc];
        K[row][cc] = cm < nv ? C[row * nv + cm] : D[row * nv + (cm - nv)];
    }
    for (int row = lane; row < nv; row += 32) K[row][nv] = g[row];
    __syncwarp();
    for (int k = 0; k < nv; ++k) {
        double best = -1.0;
        int bi = k;
        for (int row = k + lane; row < nv; row += 32) {
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
            for (int c = lane; c <= nv; c += 32) { const double tmp = K[k][c]; K[k][c] = K[bi][c]; K[bi][c] = tmp; }
        __syncwarp();
        const double pivval = K[k][k];
        for (int row = k + 1 + lane; row < nv; row += 32) {
            const double l = K[row][k] / pivval;
            for (int c = k + 1; c <= nv; ++c) K[row][c] = fma(-l, K[k][c], K[row][c]);
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (int row = nv - 1; row >= 0; --row) {
            double s = K[row][nv];
            for (int c = row + 1; c < nv; ++c) s = fma(-K[row][c], K[c][nv], s);
            K[row][nv] = s / K[row][row];
        }
        for (int c = 0; c < 2 * nv; ++c) Z[c] = 0.0;
        for (int cc = 0; cc < nv; ++cc) Z[colmap[cc]] = K[cc][nv];
    }
}

}  // namespace rst
