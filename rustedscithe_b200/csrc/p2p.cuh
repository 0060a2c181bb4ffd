// p2p.cuh -- latency-bound all-gather of the mesh-slab interface data over NVLink peer memory.
//
// The only data-path exchange of the multi-GPU path is an all-gather of a few hundred
// doubles per rank: the condensed interface row (2 nv^2) per factorisation, its right-hand side (nv) per solve and
// 8 damping scalars per trial.  Through NCCL + a host callback each of them costs ~25 us of launch/host latency,
// which is a quarter of a 1 ms Newton solve at C2.  Here every rank owns a "mailbox" in its HBM that its peers map
// through CUDA IPC; the gather is ONE small kernel per rank on the solver's own stream:
//   1. store my payload into slot [parity][my_rank] of every peer's mailbox (peer stores travel over NVLink),
//   2. spin on my own mailbox until every word carries the tag of this exchange, keep the values.
// Wire format (the "LL" protocol of collective libraries): PTX guarantees single-copy atomicity per 8-byte element
// only, so EVERY 8-byte word carries its own flag -- a double travels as two words
//       word0 = (low 32 bits of the value)  | (tag << 32)      word1 = (high 32 bits) | (tag << 32)
// with tag = low 32 bits of the exchange sequence number.  A word whose tag matches is complete by itself: no fence, no
// separate flag round trip, and a torn 16-byte store can at worst delay one half, never pair a stale half with a fresh
// tag.  (Round 1 sent (value, tag) as one 16-byte store and relied on it arriving untorn.)
// One process per GPU, so the kernels of different ranks always run concurrently (each on its own device); slots are
// double-buffered by the parity of the sequence number: a rank cannot run two gathers ahead of a peer because it
// waits for that peer's words of the gather in between.
#pragma once
#include "rst_common.cuh"

namespace rst {

constexpr int P2P_MAX_WORLD = 16;

struct P2pPeers {
    double *mailbox[P2P_MAX_WORLD];  // peer mailboxes (own entry = local pointer)
    // exchange sequence number, kept on the DEVICE (every rank runs the same exchanges in stream order, so the local
    // counters agree): kernel arguments stay constant from launch to launch and the sequences can be replayed as CUDA graphs
    unsigned long long *seq_dev;
    // spin budget in clock cycles before a missing peer is reported (RST_FLAG_PEER_TIMEOUT); the handle is then poisoned:
    // the received values are NaN, never stale data
    long long timeout_cycles;
};

// mailbox layout: words[2 parities][world][pay][2] of 8 bytes
__host__ __device__ inline size_t p2p_mailbox_doubles(int world, int pay) { return (size_t)2 * 2 * world * pay; }

__device__ __forceinline__ void st_ll_pair(double *p, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long w0 = (b & 0xffffffffull) | t, w1 = (b >> 32) | t;
    asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ld_ll_pair(const double *p, unsigned tag, double &v) {
    unsigned long long w0, w1;
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    return (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag;
}

// The gather as a block-wide device function (every thread of ONE CTA calls it): used stand-alone below and fused into
// the single-CTA tail / reduction kernels so a multi-GPU solve needs no extra launches.
__device__ __forceinline__ void p2p_allgather_block(const P2pPeers &peers, const double *send, int count, double *recv,
                                                    int my_rank, int world, int pay, unsigned long long seq,
                                                    unsigned *flags) {
    (void)seq;   // legacy argument: the sequence number lives in peers.seq_dev
    __shared__ unsigned long long seq_sh;
    __syncthreads();  // `send` may have been produced by other threads of this CTA
    if (threadIdx.x == 0) seq_sh = *peers.seq_dev + 1ull;
    __syncthreads();
    seq = seq_sh;
    const unsigned tag = (unsigned)seq;
    const int parity = (int)(seq & 1ull);
    const size_t slot = ((size_t)parity * world + my_rank) * pay * 2;          // in doubles: two words per value
    for (int p = 0; p < world; ++p) {
        double *dst = peers.mailbox[p] + slot;
        for (int i = threadIdx.x; i < count; i += blockDim.x) st_ll_pair(dst + 2 * i, send[i], tag);
    }
    const double *mine = peers.mailbox[my_rank] + (size_t)parity * world * pay * 2;
    const long long budget = peers.timeout_cycles > 0 ? peers.timeout_cycles : 60000000000ll;   // default ~30 s
    for (int idx = threadIdx.x; idx < world * count; idx += blockDim.x) {
        const int p = idx / count, i = idx - p * count;
        const double *src = mine + ((size_t)p * pay + i) * 2;
        double v;
        const long long t0 = clock64();
        for (;;) {
            if (ld_ll_pair(src, tag, v)) break;
            if (clock64() - t0 > budget) {  // a peer never arrived: poison, do not hand out stale data
                atomicOr(flags, RST_FLAG_PEER_TIMEOUT);
                atomicOr(flags + 2, RST_FLAG_PEER_TIMEOUT);   // sticky word: survives the per-factorisation reset of flags[0]
                v = __longlong_as_double(0x7ff8000000000000ll);
                break;
            }
        }
        recv[idx] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) *peers.seq_dev = seq;
    __syncthreads();
}

__global__ void __launch_bounds__(256) p2p_allgather_kernel(P2pPeers peers, const double *__restrict__ send, int count,
                                                            double *__restrict__ recv, int my_rank, int world, int pay,
                                                            unsigned long long seq, unsigned *flags) {
    p2p_allgather_block(peers, send, count, recv, my_rank, world, pay, seq, flags);
}

// ------------------------------------------------------------------------------------------------
// Self-test of the exchange protocol on ONE device: CTA p plays rank p (own mailbox, own sequence counter, the same
// p2p_allgather_block the solver uses), `rounds` back-to-back gathers with round-dependent payloads and sizes, every
// received word checked bit for bit.  One cooperative launch, so the CTAs are co-resident by construction (spin-waiting
// kernels must never be separate launches on one GPU).  `absent` >= 0: that rank never shows up -> the others must report
// RST_FLAG_PEER_TIMEOUT and receive NaN for its words instead of stale data.
// ------------------------------------------------------------------------------------------------
struct P2pSelfTest {
    P2pPeers peers[P2P_MAX_WORLD];   // per emulated rank (they differ in seq_dev)
    double *send;                    // [world][pay]
    double *recv;                    // [world][world * pay]
    unsigned *flags;                 // [world][4]
    unsigned long long *mismatches;  // [1]
    unsigned long long *nan_words;   // [1] words received as NaN (timeout path)
    int world, pay, rounds, absent;
    int counts[3];
};

__device__ __forceinline__ double p2p_selftest_value(int rank, int round, int i) {
    // distinct bit patterns in both 32-bit halves
    const unsigned long long x = 0x9E3779B97F4A7C15ull * (unsigned long long)(1 + rank + 17 * (i + 131 * round)) + 0x1234567ull * i;
    return __longlong_as_double((long long)((x & 0x800fffffffffffffull) | 0x3ff0000000000000ull));
}

__global__ void __launch_bounds__(256) p2p_selftest_kernel(P2pSelfTest t) {
    const int rank = blockIdx.x;
    if (rank == t.absent) return;
    double *send = t.send + (size_t)rank * t.pay;
    double *recv = t.recv + (size_t)rank * t.world * t.pay;
    for (int r = 0; r < t.rounds; ++r) {
        const int count = t.counts[r % 3];
        for (int i = threadIdx.x; i < count; i += blockDim.x) send[i] = p2p_selftest_value(rank, r, i);
        p2p_allgather_block(t.peers[rank], send, count, recv, rank, t.world, t.pay, 0ull, t.flags + 4 * rank);
        for (int idx = threadIdx.x; idx < t.world * count; idx += blockDim.x) {
            const int p = idx / count, i = idx - p * count;
            const double got = recv[idx], want = p2p_selftest_value(p, r, i);
            if (got != got) atomicAdd(t.nan_words, 1ull);
            else if (__double_as_longlong(got) != __double_as_longlong(want)) atomicAdd(t.mismatches, 1ull);
        }
        __syncthreads();
        if (t.flags[4 * rank + 2] != 0u) return;   // poisoned: stop (every later gather would time out as well)
    }
}

// ------------------------------------------------------------------------------------------------
// Interface system of the mesh-slab partition, solved redundantly on every rank INSIDE the single-CTA kernels:
//   C_p Z_p + D_p Z_{p+1} = g_p (p = 0..world-1, nv rows each)  +  nv pinned boundary components of Z_0 / Z_world.
// Dense (world+1) nv square system: LU with partial pivoting at factor time, substitution per right-hand side.
// ------------------------------------------------------------------------------------------------
struct P2pIface {
    P2pPeers peers;
    int rank, world, pay;
    const double *rows_send;  // [2 nv^2]  (C | D) of this rank
    double *rows_gath;        // [world][2 nv^2]
    const double *rhs_send;   // [nv]
    double *rhs_gath;         // [world][nv]
    double *lu;               // [m][m] row-major, m = (world+1) nv
    int *perm;                // [m] row swapped with k at step k
    double *Zred;             // [world+1][nv]
    unsigned long long left_mask, right_mask;
};

template <int NV>
__device__ void iface_factor_block(const P2pIface &I, unsigned long long seq, unsigned *flags) {
    constexpr int MMAX = (P2P_MAX_WORLD + 1) * NV;
    __shared__ double K[MMAX][MMAX + 1];
    __shared__ int s_piv;
    const int m = (I.world + 1) * NV;
    p2p_allgather_block(I.peers, I.rows_send, 2 * NV * NV, I.rows_gath, I.rank, I.world, I.pay, seq, flags);
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) K[idx / m][idx % m] = 0.0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < I.world * NV * NV; idx += blockDim.x) {
        const int p = idx / (NV * NV), e = idx % (NV * NV), i = e / NV, c = e % NV;
        K[p * NV + i][p * NV + c] = I.rows_gath[(size_t)p * 2 * NV * NV + e];
        K[p * NV + i][(p + 1) * NV + c] = I.rows_gath[(size_t)p * 2 * NV * NV + NV * NV + e];
    }
    if (threadIdx.x == 0) {  // pinned components: unit rows
        int row = I.world * NV;
        for (int c = 0; c < NV; ++c)
            if ((I.left_mask >> c) & 1ull) { if (row < m) K[row][c] = 1.0; ++row; }
        for (int c = 0; c < NV; ++c)
            if ((I.right_mask >> c) & 1ull) { if (row < m) K[row][I.world * NV + c] = 1.0; ++row; }
        if (row != m) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        if (threadIdx.x == 0) {
            int p = k;
            double best = fabs(K[k][k]);
            for (int r = k + 1; r < m; ++r)
                if (fabs(K[r][k]) > best) { best = fabs(K[r][k]); p = r; }
            if (!(best > 0.0)) atomicOr(flags, RST_FLAG_ZERO_PIVOT);
            s_piv = p;
            I.perm[k] = p;
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k)
            for (int c = threadIdx.x; c < m; c += blockDim.x) { const double t = K[k][c]; K[k][c] = K[p][c]; K[p][c] = t; }
        __syncthreads();
        const double piv = K[k][k];
        for (int r = k + 1 + threadIdx.x; r < m; r += blockDim.x) {
            const double l = K[r][k] / piv;
            K[r][k] = l;
            for (int c = k + 1; c < m; ++c) K[r][c] = fma(-l, K[k][c], K[r][c]);
        }
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) I.lu[idx] = K[idx / m][idx % m];
    __syncthreads();
}

template <int NV>
__device__ void iface_solve_block(const P2pIface &I, unsigned long long seq, unsigned *flags) {
    constexpr int MMAX = (P2P_MAX_WORLD + 1) * NV;
    __shared__ double b[MMAX];
    __shared__ double lu_s[MMAX][MMAX + 1];
    __shared__ int perm_s[MMAX];
    const int m = (I.world + 1) * NV;
    // the factors of the interface system come from HBM once, BEFORE the exchange (their latency hides behind the wait for
    // the peers); substituting straight from global memory cost one dependent round trip per row (~18 us at 8 ranks)
    for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) lu_s[idx / m][idx % m] = I.lu[idx];
    for (int k = threadIdx.x; k < m; k += blockDim.x) perm_s[k] = I.perm[k];
    p2p_allgather_block(I.peers, I.rhs_send, NV, I.rhs_gath, I.rank, I.world, I.pay, seq, flags);
    for (int i = threadIdx.x; i < m; i += blockDim.x) b[i] = i < I.world * NV ? I.rhs_gath[i] : 0.0;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        if (lane == 0)
            for (int k = 0; k < m; ++k) {
                const int p = perm_s[k];
                if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
            }
        __syncwarp();
        if (m <= 32) {
            // lane = row: column-oriented substitution, one shuffle + one FMA per step, no reductions
            double bi = lane < m ? b[lane] : 0.0;
            for (int k = 0; k + 1 < m; ++k) {
                const double bk = __shfl_sync(RST_FULL_MASK, bi, k);
                if (lane > k && lane < m) bi = fma(-lu_s[lane][k], bk, bi);
            }
            for (int k = m - 1; k >= 0; --k) {
                if (lane == k) bi = bi / lu_s[k][k];
                const double xk = __shfl_sync(RST_FULL_MASK, bi, k);
                if (lane < k) bi = fma(-lu_s[lane][k], xk, bi);
            }
            if (lane < m) b[lane] = bi;
        } else {  // more than 32 interface unknowns: row-oriented with warp dot products (from shared memory)
            for (int i = 1; i < m; ++i) {
                double sacc = 0.0;
                for (int j = lane; j < i; j += 32) sacc = fma(lu_s[i][j], b[j], sacc);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) sacc += __shfl_xor_sync(RST_FULL_MASK, sacc, off);
                if (lane == 0) b[i] -= sacc;
                __syncwarp();
            }
            for (int i = m - 1; i >= 0; --i) {
                double sacc = 0.0;
                for (int j = i + 1 + lane; j < m; j += 32) sacc = fma(lu_s[i][j], b[j], sacc);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) sacc += __shfl_xor_sync(RST_FULL_MASK, sacc, off);
                if (lane == 0) b[i] = (b[i] - sacc) / lu_s[i][i];
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) I.Zred[i] = b[i];
    __syncthreads();
}

}  // namespace rst
