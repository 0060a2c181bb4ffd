// rst_b200.cu -- C ABI (include/rst_b200.h) over the sm_100a kernels: problem registry, device-resident
// solver handle, recursive condensation orchestration, damped Newton host loop.
//
// Host-side restatement of (paths relative to the reference root):
//   NRBVP::step / damped_step / try_main_loop_damped   src/numerical/BVP_Damp/NR_Damp_solver_damped.rs:1782-2156
//   jac_recalc                                          src/numerical/BVP_Damp/BVP_utils_damped.rs:185-207
//   unknown layout / BC elimination                     src/numerical/BVP_Damp/numeric_discretization.rs:27-136
// There is deliberately no CPU fallback anywhere in this file.
#include "../../include/rst_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "abd.cuh"
#include "abd3.cuh"
#include "abd4.cuh"
#include "tps.cuh"
#include "btd.cuh"
#include "p2p.cuh"
#include "big.cuh"
#include "bigd.cuh"
#include "grid.cuh"
#include "assemble.cuh"
#include "newton.cuh"
#include "generated_problems.cuh"  // from the -I generated directory (default: csrc/generated)

using namespace rst;

static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(RST_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

// ------------------------------------------------------------------------------------------------
// problem registry: one entry per struct the CUDA-C emitter generated
// ------------------------------------------------------------------------------------------------
struct ProblemVT {
    const char *key;
    int nv, np;
    const char *hash;
    const unsigned char *pattern;
    void (*assemble)(const AsmArgs &, bool with_j, bool trapezoid, cudaStream_t);
};

template <class P>
static void launch_assemble(const AsmArgs &a, bool with_j, bool trap, cudaStream_t st) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((a.n + threads - 1) / threads);
    if (blocks == 0) return;
    if constexpr (P::NV > 16) {  // wide blocks: sink-based emitter output, blocks pre-initialised with +-I at create
        const unsigned wb = (unsigned)((a.n + 63) / 64);
        if (!trap) {
            if (with_j) assemble_wide_kernel<P, true, false><<<wb, 64, 0, st>>>(a);
            else assemble_wide_kernel<P, false, false><<<wb, 64, 0, st>>>(a);
        } else {
            if (with_j) assemble_wide_kernel<P, true, true><<<wb, 64, 0, st>>>(a);
            else assemble_wide_kernel<P, false, true><<<wb, 64, 0, st>>>(a);
        }
        return;
    } else {
    // nv >= 6: blocks of A are staged through shared memory and streamed out coalesced (assemble.cuh)
    if constexpr (P::NV >= 6 && (P::NV % 2 == 0) && (32 * ((P::NV * P::NV) | 1) * 8 <= 48 * 1024)) {
        static const bool direct = std::getenv("RST_B200_ASSEMBLE_DIRECT") != nullptr;
        if (with_j && !direct) {
            const unsigned wblocks = (unsigned)((a.n + 31) / 32);
            if (trap) assemble_j_staged_kernel<P, true><<<wblocks, 32, 0, st>>>(a);
            else assemble_j_staged_kernel<P, false><<<wblocks, 32, 0, st>>>(a);
            return;
        }
    }
    // residual only, nv >= 4: state and residual staged through shared memory, coalesced both ways (assemble.cuh)
    if constexpr (P::NV >= 4) {
        static const bool direct_f = std::getenv("RST_B200_ASSEMBLE_DIRECT") != nullptr;
        if (!with_j && !direct_f) {
            if (trap) assemble_f_staged_kernel<P, true><<<blocks, 128, 0, st>>>(a);
            else assemble_f_staged_kernel<P, false><<<blocks, 128, 0, st>>>(a);
            return;
        }
    }
    if (!trap) {
        if (with_j) assemble_kernel<P, true, false><<<blocks, threads, 0, st>>>(a);
        else assemble_kernel<P, false, false><<<blocks, threads, 0, st>>>(a);
    } else {
        if (with_j) assemble_kernel<P, true, true><<<blocks, threads, 0, st>>>(a);
        else assemble_kernel<P, false, true><<<blocks, threads, 0, st>>>(a);
    }
    }
}

static const ProblemVT g_problems[] = {
#define RST_PROBLEM(P, H) {P::key(), P::NV, P::NP, H, P::pattern(), &launch_assemble<P>},
#include "generated_registry.inc"
#undef RST_PROBLEM
};
static const int g_nproblems = (int)(sizeof(g_problems) / sizeof(g_problems[0]));

static const ProblemVT *find_problem(const char *key) {
    if (!key) return nullptr;
    for (int i = 0; i < g_nproblems; ++i)
        if (std::strcmp(g_problems[i].key, key) == 0) return &g_problems[i];
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// block-size dispatch of the solver kernels
// ------------------------------------------------------------------------------------------------
struct SolverVT {
    int nv;
    void (*factor)(const AbdLevel &, bool ident_b, unsigned *flags, cudaStream_t);
    void (*forward)(const AbdLevel &, cudaStream_t);
    void (*backward)(const AbdLevel &, cudaStream_t);
    void (*top)(const double *C, const double *D, const double *g, double *Z, unsigned long long lm,
                unsigned long long rm, unsigned *flags, int nv, cudaStream_t);
    void (*residual)(const double *A, const double *B, const double *X, const double *rhs, double *out, int64_t n,
                     cudaStream_t);
    void (*reduce)(const ReduceArgs &, int blocks, cudaStream_t);
    void (*top_general)(const double *C, const double *D, const double *g, const double *Ba, const double *ba, int na,
                        const double *Bb, const double *bb, int nb, double *Z, unsigned *flags, cudaStream_t);
    // single-CTA tail of the level hierarchy (abd4.cuh); nullptr where the block size has no generation-4 kernels
    void (*tail_forward)(const AbdTail &, cudaStream_t);
    void (*tail_backward)(const AbdTail &, unsigned *flags, cudaStream_t);
    // factorisation with the level's right-hand side riding along as a panel column (abd4.cuh, nv = 12): also writes L.e and
    // L.r_next, i.e. what `forward` would compute from the stored factors, bit for bit; nullptr where there is no spare column
    void (*factor_rhs)(const AbdLevel &, bool ident_b, unsigned *flags, cudaStream_t);
};

// kernel generation: "v1" = one panel row per lane + shuffle broadcast (abd.cuh); default "v3" (abd3.cuh).  Both write
// the packed factor layout of abd.cuh.  (Generation 2 -- two rows per lane -- was slower and is retired.)
static int kernel_generation() {
    static int gen = [] {
        const char *e = std::getenv("RST_B200_KERNELS");
        if (e && std::strcmp(e, "v1") == 0) return 1;
        if (e && std::strcmp(e, "v3") == 0) return 3;  // abd3.cuh: one row per lane, REDUX pivot search, shared-memory broadcast
        return 4;  // abd4.cuh (nv = 8, 12, 16): blocked elimination, rank-4 updates on the FP64 tensor cores; abd3.cuh otherwise
    }();
    return gen;
}

// Launch with programmatic stream serialisation: the kernel may be scheduled while its predecessor in the stream is still
// running; it orders itself with griddepcontrol.wait (pdl_wait, abd.cuh).  Captured into CUDA graphs as a programmatic edge.
template <class... KArgs, class... Args>
static void launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
    static const bool on = std::getenv("RST_B200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int NV>
struct SolverImpl {
    static unsigned grid_for(int64_t nslabs, int lanes) {
        const int gpw = 32 / next_pow2(lanes);
        const int64_t per_block = 4 * gpw;
        return (unsigned)((nslabs + per_block - 1) / per_block);
    }
    // generation 4 (abd4.cuh) changes the multiplier layout: factor and forward sweep switch together
    static constexpr bool HAS_V4 = (NV % 4 == 0) && NV >= 8 && NV <= 16;
    static bool use_v4() { return HAS_V4 && kernel_generation() >= 4; }
    static void factor(const AbdLevel &L, bool ident_b, unsigned *flags, cudaStream_t st) {
        if constexpr (HAS_V4) {
            if (use_v4()) {
                const unsigned blocks = (unsigned)((L.nslabs + 3) / 4);
                // > 48 KB of dynamic shared memory: opt in on every launch (the attribute is per device and per loaded module)
                if (ident_b) {
                    constexpr int smem = 4 * (int)sizeof(Abd4Smem<NV, true>);
                    cudaFuncSetAttribute(abd_factor4_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    abd_factor4_kernel<NV, true><<<blocks, 128, smem, st>>>(L, flags);
                } else {
                    constexpr int smem = 4 * (int)sizeof(Abd4Smem<NV, false>);
                    cudaFuncSetAttribute(abd_factor4_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    abd_factor4_kernel<NV, false><<<blocks, 128, smem, st>>>(L, flags);
                }
                return;
            }
        }
        // REDUX with several sub-warp masks serialises per mask: generation 3 only where a slab spans >= 16 lanes
        if constexpr (NV % 2 == 0) {
            if (kernel_generation() >= 2 && NV >= 6) {
                const unsigned blocks = grid_for(L.nslabs, 2 * NV);
                if (ident_b) abd_factor3_kernel<NV, true><<<blocks, 128, 0, st>>>(L, flags);
                else abd_factor3_kernel<NV, false><<<blocks, 128, 0, st>>>(L, flags);
                return;
            }
        }
        const unsigned blocks = grid_for(L.nslabs, 2 * NV);
        if (ident_b) abd_factor_kernel<NV, true><<<blocks, 128, 0, st>>>(L, flags);
        else abd_factor_kernel<NV, false><<<blocks, 128, 0, st>>>(L, flags);
    }
    static constexpr bool HAS_RHS = []() { if constexpr (HAS_V4) return Abd4Cfg<NV>::HAS_RHS_COL; else return false; }();
    static void factor_rhs(const AbdLevel &L, bool ident_b, unsigned *flags, cudaStream_t st) {
        if constexpr (HAS_RHS) {
            const unsigned blocks = (unsigned)((L.nslabs + 3) / 4);
            if (ident_b) {
                constexpr int smem = 4 * (int)sizeof(Abd4Smem<NV, true>);
                cudaFuncSetAttribute(abd_factor4_kernel<NV, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                abd_factor4_kernel<NV, true, true><<<blocks, 128, smem, st>>>(L, flags);
            } else {
                constexpr int smem = 4 * (int)sizeof(Abd4Smem<NV, false>);
                cudaFuncSetAttribute(abd_factor4_kernel<NV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                abd_factor4_kernel<NV, false, true><<<blocks, 128, smem, st>>>(L, flags);
            }
        }
    }
    static void forward(const AbdLevel &L, cudaStream_t st) {
        if constexpr (HAS_V4) {
            if (use_v4()) {
                launch_pdl(abd_forward4_kernel<NV>, (unsigned)((L.nslabs + 3) / 4), 128u, st, L);
                return;
            }
        }
        abd_forward_kernel<NV><<<grid_for(L.nslabs, 2 * NV), 128, 0, st>>>(L);
    }
    static void backward(const AbdLevel &L, cudaStream_t st) {
        launch_pdl(abd_backward_kernel<NV>, grid_for(L.nslabs, NV), 128u, st, L);
    }
    static void top(const double *C, const double *D, const double *g, double *Z, unsigned long long lm,
                    unsigned long long rm, unsigned *flags, int, cudaStream_t st) {
        abd_top_solve_kernel<NV><<<1, 32, 0, st>>>(C, D, g, Z, lm, rm, flags);
    }
    static void residual(const double *A, const double *B, const double *X, const double *rhs, double *out, int64_t n,
                         cudaStream_t st) {
        const int64_t total = n * NV;
        block_residual_kernel<NV><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(A, B, X, rhs, out, n);
    }
    static void reduce(const ReduceArgs &a, int blocks, cudaStream_t st) {
        newton_reduce_kernel<NV><<<blocks, 256, 0, st>>>(a);
    }
    static void top_general(const double *C, const double *D, const double *g, const double *Ba, const double *ba, int na,
                            const double *Bb, const double *bb, int nb, double *Z, unsigned *flags, cudaStream_t st) {
        abd_top_general_kernel<NV><<<1, 32, 0, st>>>(C, D, g, Ba, ba, na, Bb, bb, nb, Z, flags);
    }
    static void tail_forward(const AbdTail &T, cudaStream_t st) {
        if constexpr (HAS_V4) launch_pdl(abd_tail_forward_kernel<NV>, 1u, (unsigned)ABD_TAIL_THREADS, st, T);
    }
    static void tail_backward(const AbdTail &T, unsigned *flags, cudaStream_t st) {
        if constexpr (HAS_V4) launch_pdl(abd_tail_backward_kernel<NV>, 1u, (unsigned)ABD_TAIL_THREADS, st, T, flags);
    }
    static SolverVT vt() {
        const bool tail = use_v4() && std::getenv("RST_B200_NO_TAIL") == nullptr;
        const bool rhs = HAS_RHS && use_v4() && std::getenv("RST_B200_NO_RHS_COLUMN") == nullptr;
        return SolverVT{NV, &factor, &forward, &backward, &top, &residual, &reduce, &top_general,
                        tail ? &tail_forward : nullptr, tail ? &tail_backward : nullptr, rhs ? &factor_rhs : nullptr};
    }
};

// wide blocks (16 < nv <= 64, or any nv >= 3 with RST_B200_FORCE_BIG): run-time-nv block-per-slab kernels (big.cuh)
struct BigImpl {
    template <int NVP, int NW>
    static void factor_tiled(const AbdLevel &L, bool ident_b, unsigned *flags, cudaStream_t st) {
        using C = BigrCfg<NVP, NW>;
        // the opt-in is per device and per loaded module: set it on every launch (cheap), never cache it per process
        cudaFuncSetAttribute(bigr_factor_kernel<NVP, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
        cudaFuncSetAttribute(bigr_factor_kernel<NVP, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
        if (ident_b) bigr_factor_kernel<NVP, NW, true><<<(unsigned)L.nslabs, C::THREADS, C::smem_bytes, st>>>(L, L.nv, flags);
        else bigr_factor_kernel<NVP, NW, false><<<(unsigned)L.nslabs, C::THREADS, C::smem_bytes, st>>>(L, L.nv, flags);
    }
    static void factor(const AbdLevel &L, bool ident_b, unsigned *flags, cudaStream_t st) {
        // default: register-tiled panel (FP64-pipe bound); RST_B200_BIG_V1 selects the first shared-memory-panel kernel
        const bool v1 = std::getenv("RST_B200_BIG_V1") != nullptr;   // read per call: tests toggle it
        // nv = 64: blocked elimination with the rank-4 updates on the FP64 tensor cores (bigd.cuh); RST_B200_BIG_DFMA keeps
        // the register-tiled DFMA kernel (same factor layout) for the A/B comparison the north star asks for
        if (!v1 && L.nv == BigdCfg::NV && std::getenv("RST_B200_BIG_DFMA") == nullptr && std::getenv("RST_B200_BIG_W32") == nullptr) {
            constexpr int smem = (int)sizeof(BigdSmem);
            if (ident_b) {
                cudaFuncSetAttribute(bigd_factor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                bigd_factor_kernel<true><<<(unsigned)L.nslabs, BigdCfg::THREADS + 128, smem, st>>>(L, flags);
            } else {
                cudaFuncSetAttribute(bigd_factor_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                bigd_factor_kernel<false><<<(unsigned)L.nslabs, BigdCfg::THREADS + 128, smem, st>>>(L, flags);
            }
            return;
        }
        if (!v1) {
            // RST_B200_BIG_W32: 32 warps x (4 x 6) tile at 64 registers instead of 16 warps x (8 x 6) at 128 -- measured
            // SLOWER (185 vs 112 ms per 2.5e5 rows: spills + uncoalesced pivot-row stores), kept for experiments
            const bool w32 = std::getenv("RST_B200_BIG_W32") != nullptr;
            if (L.nv <= 32) factor_tiled<32, 16>(L, ident_b, flags, st);
            else if (w32) factor_tiled<64, 32>(L, ident_b, flags, st);
            else factor_tiled<64, 16>(L, ident_b, flags, st);
            return;
        }
        const size_t smem = big_factor_smem_bytes(L.nv);
        cudaFuncSetAttribute(big_factor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_factor_smem_bytes(BIG_MAX_NV));
        cudaFuncSetAttribute(big_factor_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_factor_smem_bytes(BIG_MAX_NV));
        if (ident_b) big_factor_kernel<true><<<(unsigned)L.nslabs, BIG_THREADS, smem, st>>>(L, L.nv, flags);
        else big_factor_kernel<false><<<(unsigned)L.nslabs, BIG_THREADS, smem, st>>>(L, L.nv, flags);
    }
    // sweeps: generation 2 streams the factors through a shared-memory ring with bulk async copies (needs 16-byte
    // aligned stages, i.e. even nv); RST_B200_BIG_V1 or odd nv select the first kernels
    static bool pipelined(int nv) {
        const bool v1 = std::getenv("RST_B200_BIG_V1") != nullptr;   // read per call: tests toggle it
        return !v1 && (nv % 2 == 0);
    }
    static void forward(const AbdLevel &L, cudaStream_t st) {
        if (!pipelined(L.nv)) {
            big_forward_kernel<<<(unsigned)L.nslabs, 128, 0, st>>>(L, L.nv);
            return;
        }
        cudaFuncSetAttribute(bigp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bigp_forward_smem(BIG_MAX_NV));
        bigp_forward_kernel<<<(unsigned)L.nslabs, BIGP_THREADS, bigp_forward_smem(L.nv), st>>>(L, L.nv);
    }
    static void backward(const AbdLevel &L, cudaStream_t st) {
        if (!pipelined(L.nv)) {
            big_backward_kernel<<<(unsigned)L.nslabs, 64, 0, st>>>(L, L.nv);
            return;
        }
        cudaFuncSetAttribute(bigp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bigp_backward_smem(BIG_MAX_NV));
        bigp_backward_kernel<<<(unsigned)L.nslabs, BIGP_THREADS, bigp_backward_smem(L.nv), st>>>(L, L.nv);
    }
    static void top(const double *C, const double *D, const double *g, double *Z, unsigned long long lm,
                    unsigned long long rm, unsigned *flags, int nv, cudaStream_t st) {
        big_top_solve_kernel<<<1, 32, 0, st>>>(C, D, g, Z, nv, lm, rm, flags);
    }
    static void reduce(const ReduceArgs &a, int blocks, cudaStream_t st) {
        newton_reduce_generic_kernel<<<blocks, 256, 0, st>>>(a);
    }
    static SolverVT vt(int nv) { return SolverVT{nv, &factor, &forward, &backward, &top, nullptr, &reduce, nullptr, nullptr, nullptr}; }   // residual: generic kernel
};

static bool use_big_path(int nv) {
    return nv > 16 || (nv >= 3 && std::getenv("RST_B200_FORCE_BIG") != nullptr);  // env: test hook, read per create
}

static bool solver_vt_for(int nv, SolverVT *out, bool allow_big = true) {
    if (allow_big && use_big_path(nv)) {
        if (nv > BIG_MAX_NV) return false;
        *out = BigImpl::vt(nv);
        return true;
    }
    switch (nv) {
#define RST_NV(n) case n: *out = SolverImpl<n>::vt(); return true;
#include "generated_nv_list.inc"  // block sizes of the compiled-in problems + the BlockTridiagonal drop-in sizes
#undef RST_NV
        default: return false;
    }
}

// ------------------------------------------------------------------------------------------------
// solver handle
// ------------------------------------------------------------------------------------------------
struct LevelBuf {
    AbdLevel k;
    bool ident_b;
};

struct rst_b200_solver {
    const ProblemVT *pb = nullptr;
    SolverVT sv{};
    int nv = 0, np = 0, scheme = 0, trap_time = 0, device = 0;
    int rank = 0, world = 1;
    int64_t N = 0;         // global intervals
    int64_t n = 0;         // local intervals
    int64_t j_begin = 0;
    double t0 = 0, t_end = 1, h = 0;
    bool uniform = true;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // BC layout (global)
    unsigned long long left_mask = 0, right_mask = 0;
    std::vector<double> left_val, right_val;  // per component (valid where masked)
    LayoutArgs layout{};
    // device buffers
    double *Y = nullptr, *Yt = nullptr, *F = nullptr, *A = nullptr, *B = nullptr, *dY = nullptr;
    double *d_mesh = nullptr, *d_params = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_rtol = nullptr;
    double *dY_home = nullptr; // the buffer the undamped step starts in (dY and dY1 swap roles when a step is carried over)
    double *Y_home = nullptr; // the buffer Y starts in: the Newton drivers begin every solve there (stable CUDA-graph keys)
    double *dY1 = nullptr;    // step of a damping trial (the undamped step in dY must survive rejected trials)
    double *sol = nullptr;    // buffer the linear solve currently writes to (dY or dY1)
    double *Ysave = nullptr;  // device-side checkpoint of Y (rst_b200_checkpoint)
    double *d_tmp = nullptr;  // [n*nv + nv] staging for reduced-order host drop-ins
    double *ref_b = nullptr, *ref_x = nullptr, *ref_xt = nullptr;   // guarded refinement: rhs copy, iterate, trial iterate
    unsigned long long *ref_norm = nullptr;
    std::mutex aot_mtx;   // serialises the handle-less AOT symbols on this handle (the reference may call them from several rayon threads)
    // LinearSolverConfig.  ls_refine_steps < 0 = library default: the DirectLinearSolver drop-ins take ONE guarded refinement step
    // (tol 0: only the "residual must decrease" guard ends it) on meshes of >= 4096 intervals, where the forward error of the
    // condensation exceeds the 1e-10 contract; the Newton loop never refines
    int ls_policy = RST_B200_LS_AUTO, ls_fallback = RST_B200_FALLBACK_TO_FAER_SPARSE, ls_refine_steps = -1;
    std::vector<void *> owned;
    std::vector<LevelBuf> levels;      // local condensation levels (lanes-per-slab kernels)
    int tail_begin = -1;               // >= 0: levels[tail_begin..] run inside the single-CTA tail launches (abd4.cuh)
    bool use_tps = false;              // nv == 2: thread-per-slab kernels + single-CTA tail (tps.cuh)
    std::vector<TpsLevel> tps_big;     // levels launched one kernel each
    TpsTail tps_tail{};                // all remaining levels
    bool have_red = false;
    LevelBuf red{};                    // interface (reduced) level when world > 1
    double *iface_rows = nullptr;      // [2 nv^2]  (A | B) of the local condensed row
    double *iface_rhs = nullptr;       // [nv]
    double *gath_rows = nullptr;       // [world][2 nv^2]
    double *gath_rhs = nullptr;        // [world][nv]
    double *redA = nullptr, *redB = nullptr, *red_r = nullptr, *Zred = nullptr;
    double *topA = nullptr, *topB = nullptr, *top_r = nullptr, *Ztop = nullptr;
    // reductions
    int red_blocks = 0;
    double *d_partials = nullptr, *d_scal = nullptr, *d_gscal = nullptr;
    double *h_scal = nullptr;  // pinned (mapped): [0..15] two scalar sets, [16] ticket written by publish_scalars_kernel
    unsigned long long *d_ticket = nullptr;
    unsigned long long ticket_expected = 0;
    unsigned *d_flags = nullptr;
    // value-stream map for the legacy AOT ABI
    std::vector<int64_t> st_rows, st_cols;
    int st_kl = 0, st_ku = 0;
    int64_t *d_stream_src = nullptr;
    double *d_stream_out = nullptr;
    bool have_jac = false, factored = false;
    bool rhs_in_factors = false;   // the last factorisation carried F(y) along (SolverVT::factor_rhs): set by recalc_jacobian, consumed by the
                                   // damped step that follows it in the same driver call (its first solve skips the local forward sweeps)
    int fuse_reduce_state = -1;        // >= 0: the next solve fuses the damping reductions for that state (nv = 2 path)
    int scal_slot = 0;                 // result slot (0/1) the next reduction writes to
    int reduce_ready_state = -1;
    int reduce_ready = 0;              // > 0: partials of that many blocks are waiting for the final reduction stage
    // NVLink peer-memory all-gather (p2p.cuh)
    double *mailbox = nullptr;
    int p2p_pay = 0;
    bool p2p_enabled = false;
    P2pPeers peers{};
    std::vector<void *> peer_mapped;
    P2pIface iface{};                  // interface system solved inside the single-CTA kernels (p2p.cuh)
    double *iface_lu = nullptr;
    int *iface_perm = nullptr;
    rst_b200_allgather_fn allgather = nullptr;
    void *allgather_ctx = nullptr;
    int64_t launches = 0;
    // retained create() arguments (a refined-grid handle is created from them, rst_b200_create_refined)
    std::vector<int> k_bc_var, k_bc_side;
    std::vector<double> k_bc_val, k_params, k_lo, k_hi, k_rtol;
    int k_slab = 0;
    // adaptive grid refinement (grid.cuh): results of the last rst_b200_new_grid
    int *g_mark[2] = {nullptr, nullptr};
    int g_mark_cur = 0;
    int64_t *g_pos = nullptr, *g_block_sum = nullptr, *g_total = nullptr;
    unsigned long long *g_ext = nullptr;
    double *g_mesh_new = nullptr, *g_Y_new = nullptr;
    int64_t g_nodes_new = 0, g_added = -1, g_capacity = 0;
    // CUDA graphs of the speculative damped step (undamped step + first damping trial + scalar read-back), one per parity of
    // the Y / Y_trial buffers (accept_trial swaps them); multi-GPU with the NVLink mailboxes too (the exchange sequence
    // number lives on the device)
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; int seen = 0; };
    std::map<uintptr_t, GraphEntry> step_graphs;            // speculative damped step, key = Y buffer
    std::map<uintptr_t, GraphEntry> carry_graphs;           // the same with the undamped step carried over from the accepted trial
    std::map<uintptr_t, GraphEntry> jac_graphs;             // Jacobian assembly + factorisation of every level
    std::map<uintptr_t, GraphEntry> frozen_graphs;          // frozen-Jacobian iteration, key = Y buffer | refresh flag
};

// nv == 2 path with mapped peers: the interface exchange + solve is fused into the single-CTA tail kernels
// threads per CTA of the thread-per-slab kernels.  Small CTAs: the kernels hoist all loads of a slab, so a CTA runs in
// phases (load burst -> dependent chain -> stores); few big CTAs per SM move through them in lock-step and leave HBM idle.
static int tps_block() {
    static const int b = [] { const char *e = std::getenv("RST_B200_TPS_BLOCK"); const int v = e ? std::atoi(e) : 64; return (v == 32 || v == 64 || v == 128) ? v : 64; }();
    return b;
}
static bool fused_p2p(const rst_b200_solver *s) { return s->use_tps && s->p2p_enabled && s->world > 1; }

static P2pIface iface_args(const rst_b200_solver *s, bool fused) {
    P2pIface I = s->iface;
    if (!fused) I.world = 1;
    return I;
}

// Handle behind the reference's handle-less AOT symbols.  rst_b200_aot_bind binds for the CALLING THREAD and, the first
// time (or when the previous default goes away), process-wide: a thread that never bound anything -- a rayon worker the
// reference dispatches chunk calls to (symbolic_functions_BVP.rs:2308-2357) -- uses the library-wide default, threads that
// bound their own handle are isolated from each other.  Calls on one handle are serialised by its mutex.  Every AOT
// artefact (.so) has its own copy of these variables, i.e. its own binding.
static std::atomic<rst_b200_solver *> g_bound{nullptr};
static thread_local rst_b200_solver *t_bound = nullptr;
static rst_b200_solver *aot_current() { return t_bound ? t_bound : g_bound.load(std::memory_order_acquire); }

template <class H, class T>
static int dev_alloc(H *s, T **p, size_t count) {
    void *q = nullptr;
    if (count == 0) count = 1;
    CUDA_TRY(cudaMalloc(&q, count * sizeof(T)));
    s->owned.push_back(q);
    *p = (T *)q;
    return RST_B200_OK;
}
#define TRY(expr)                     \
    do {                              \
        int _rc = (expr);             \
        if (_rc != RST_B200_OK) return _rc; \
    } while (0)

static int default_slab(int64_t n, int requested) {
    if (requested >= 2) return requested;
    if (const char *e = std::getenv("RST_B200_SLAB_ALL")) {   // experiments: one slab length for every level
        const int v = std::atoi(e);
        if (v >= 2) return v;
    }
    // measured at nv = 12, 1e6 rows (solve ms): S = 8 1.090, 12 1.065, 16 1.049, 32 1.078, 64 1.107 -- longer slabs mean fewer
    // levels but a longer dependent chain per warp; nv = 64: 16 beats 32 as well
    if (n >= (1 << 12)) {
        if (n < (1 << 17))
            if (const char *e = std::getenv("RST_B200_SLAB_MID")) {   // experiment: slab length of the levels with 4096 <= n < 131072 rows
                const int v = std::atoi(e);
                if (v >= 2) return v;
            }
        return 16;
    }
    // small levels are latency bound (one dependent chain of S - 1 eliminations per slab): short slabs, more levels --
    // the levels below ABD_TAIL_ROWS rows share one launch per sweep (abd4.cuh), so extra levels are nearly free
    return 4;
}

template <class H>
static int build_level(H *s, LevelBuf *lv, int64_t n, int S, const double *A, const double *B,
                       const double *r, double *Z) {
    const int nv = s->nv;
    AbdLevel &k = lv->k;
    k.n = n;
    k.S = S;
    k.nv = nv;
    k.prefetch = 1;   // L2 prefetch distance of the sweeps in block rows; measured at nv = 12, 1e6 rows: solve 1.15 -> 1.08 ms, 2 / 4 not better
    if (const char *e = std::getenv("RST_B200_PREFETCH_MID")) {   // experiment: deeper prefetch on the latency-bound middle levels
        if (n < 200000) k.prefetch = std::atoi(e) > 0 ? std::atoi(e) : 1;
    }
    k.nslabs = (n + S - 1) / S;
    k.A = A;
    k.B = B;
    k.r = r;
    k.Z = Z;
    lv->ident_b = (B == nullptr);
    TRY(dev_alloc(s, &k.Lm, (size_t)n * nv * 2 * nv));
    TRY(dev_alloc(s, &k.UEF, (size_t)n * 3 * nv * nv));
    TRY(dev_alloc(s, &k.piv, (size_t)n * 2 * nv));
    TRY(dev_alloc(s, &k.e, (size_t)n * nv));
    return RST_B200_OK;
}

extern "C" int rst_b200_version(void) { return 1; }
extern "C" int rst_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
extern "C" int rst_b200_problem_count(void) { return g_nproblems; }
extern "C" const char *rst_b200_problem_key(int i) { return (i >= 0 && i < g_nproblems) ? g_problems[i].key : nullptr; }
extern "C" int rst_b200_problem_info(const char *key, int *nv, int *np, const char **hash) {
    const ProblemVT *p = find_problem(key);
    if (!p) return fail(RST_B200_ERR_UNSUPPORTED, "unknown problem key");
    if (nv) *nv = p->nv;
    if (np) *np = p->np;
    if (hash) *hash = p->hash;
    return RST_B200_OK;
}
extern "C" const char *rst_b200_last_error(void) { return g_last_error.c_str(); }

extern "C" int rst_b200_create(const rst_b200_desc *d, rst_b200_solver **out) {
    if (!d || !out) return fail(RST_B200_ERR_INVALID, "null descriptor");
    *out = nullptr;
    const ProblemVT *pb = find_problem(d->problem_key);
    if (!pb) return fail(RST_B200_ERR_UNSUPPORTED, std::string("problem key not compiled in: ") + (d->problem_key ? d->problem_key : "(null)"));
    const int nv = pb->nv;
    if (d->n_steps < 2) return fail(RST_B200_ERR_INVALID, "n_steps must be >= 2");  // numeric_discretization.rs:156
    if (d->n_bc != nv) return fail(RST_B200_ERR_INVALID, "expects exactly nv total boundary conditions");  // :60-67
    if (d->n_params != pb->np) return fail(RST_B200_ERR_INVALID, "parameter count mismatch");
    if (!d->lo || !d->hi || !d->rtol || !d->bc_var || !d->bc_side || !d->bc_val)
        return fail(RST_B200_ERR_INVALID, "missing bounds / tolerances / boundary conditions");
    const int world = d->world > 0 ? d->world : 1;
    if (d->rank < 0 || d->rank >= world) return fail(RST_B200_ERR_INVALID, "rank out of range");
    if (d->n_steps < 2 * (int64_t)world) return fail(RST_B200_ERR_INVALID, "fewer than 2 intervals per rank");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(RST_B200_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
    CUDA_TRY(cudaSetDevice(d->device));

    rst_b200_solver *s = new rst_b200_solver();
    s->pb = pb;
    s->nv = nv;
    s->np = pb->np;
    if (!solver_vt_for(nv, &s->sv)) {
        delete s;
        return fail(RST_B200_ERR_UNSUPPORTED, "block size not compiled in");
    }
    s->scheme = d->scheme;
    s->trap_time = d->trap_time;
    s->device = d->device;
    s->k_bc_var.assign(d->bc_var, d->bc_var + d->n_bc);
    s->k_bc_side.assign(d->bc_side, d->bc_side + d->n_bc);
    s->k_bc_val.assign(d->bc_val, d->bc_val + d->n_bc);
    if (d->params && pb->np > 0) s->k_params.assign(d->params, d->params + pb->np);
    s->k_lo.assign(d->lo, d->lo + nv);
    s->k_hi.assign(d->hi, d->hi + nv);
    s->k_rtol.assign(d->rtol, d->rtol + nv);
    s->k_slab = d->slab;
    s->rank = d->rank;
    s->world = world;
    s->N = d->n_steps;
    s->j_begin = (d->n_steps * (int64_t)d->rank) / world;
    const int64_t j_end = (d->n_steps * (int64_t)(d->rank + 1)) / world;
    s->n = j_end - s->j_begin;
    s->t0 = d->t0;
    s->t_end = d->t_end;
    s->h = (d->t_end - d->t0) / (double)d->n_steps;
    s->uniform = (d->mesh == nullptr);
    s->left_val.assign(nv, 0.0);
    s->right_val.assign(nv, 0.0);
    for (int k = 0; k < d->n_bc; ++k) {
        const int v = d->bc_var[k];
        if (v < 0 || v >= nv) { delete s; return fail(RST_B200_ERR_INVALID, "boundary variable index out of range"); }
        if (d->bc_side[k] == 0) {
            if ((s->left_mask >> v) & 1ull) { delete s; return fail(RST_B200_ERR_INVALID, "duplicate boundary condition"); }
            s->left_mask |= 1ull << v;
            s->left_val[v] = d->bc_val[k];
        } else if (d->bc_side[k] == 1) {
            if ((s->right_mask >> v) & 1ull) { delete s; return fail(RST_B200_ERR_INVALID, "duplicate boundary condition"); }
            s->right_mask |= 1ull << v;
            s->right_val[v] = d->bc_val[k];
        } else {
            delete s;
            return fail(RST_B200_ERR_INVALID, "boundary side flag must be 0 or 1");
        }
    }
    LayoutArgs &la = s->layout;
    la.n_steps = s->N;
    la.nv = nv;
    la.nl = la.nr = 0;
    {
        int a = 0, b = 0;
        for (int v = 0; v < nv; ++v) {
            if ((s->left_mask >> v) & 1ull) la.nl++; else la.left_free[a++] = v;
            if ((s->right_mask >> v) & 1ull) la.nr++; else la.right_free[b++] = v;
        }
    }
    if (d->stream) s->stream = (cudaStream_t)d->stream;
    else {
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return fail(RST_B200_ERR_CUDA, "stream"); }
        s->own_stream = true;
    }

    int rc = RST_B200_OK;
    auto build = [&]() -> int {
        const int64_t n = s->n;
        const size_t nodes = (size_t)(n + 1) * nv;
        TRY(dev_alloc(s, &s->Y, nodes));
        s->Y_home = s->Y;
        TRY(dev_alloc(s, &s->Yt, nodes));
        TRY(dev_alloc(s, &s->dY, nodes));
        s->dY_home = s->dY;
        TRY(dev_alloc(s, &s->dY1, nodes));
        s->sol = s->dY;
        TRY(dev_alloc(s, &s->F, (size_t)n * nv));
        TRY(dev_alloc(s, &s->A, (size_t)n * nv * nv));
        if (s->scheme == RST_B200_SCHEME_TRAPEZOID) TRY(dev_alloc(s, &s->B, (size_t)n * nv * nv));
        if (nv > 16) {  // wide path: assembly writes structural non-zeros only (assemble.cuh)
            init_identity_blocks_kernel<<<148 * 8, 256, 0, s->stream>>>(s->A, n, nv, -1.0);
            if (s->B) init_identity_blocks_kernel<<<148 * 8, 256, 0, s->stream>>>(s->B, n, nv, 1.0);
        }
        TRY(dev_alloc(s, &s->d_tmp, nodes + nv));
        TRY(dev_alloc(s, &s->d_params, (size_t)(s->np > 0 ? s->np : 1)));
        TRY(dev_alloc(s, &s->d_lo, (size_t)nv));
        TRY(dev_alloc(s, &s->d_hi, (size_t)nv));
        TRY(dev_alloc(s, &s->d_rtol, (size_t)nv));
        TRY(dev_alloc(s, &s->d_flags, (size_t)4));
        CUDA_TRY(cudaMemsetAsync(s->d_flags, 0, 4 * sizeof(unsigned), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->Y, 0, nodes * sizeof(double), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->Yt, 0, nodes * sizeof(double), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->dY, 0, nodes * sizeof(double), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->dY1, 0, nodes * sizeof(double), s->stream));
        if (s->np > 0) CUDA_TRY(cudaMemcpyAsync(s->d_params, d->params, s->np * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_lo, d->lo, nv * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_hi, d->hi, nv * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_rtol, d->rtol, nv * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        if (!s->uniform) {
            TRY(dev_alloc(s, &s->d_mesh, (size_t)(n + 1)));
            CUDA_TRY(cudaMemcpyAsync(s->d_mesh, d->mesh + s->j_begin, (n + 1) * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        }
        // interface + top buffers
        TRY(dev_alloc(s, &s->iface_rows, (size_t)2 * nv * nv));
        TRY(dev_alloc(s, &s->iface_rhs, (size_t)nv));
        TRY(dev_alloc(s, &s->Ztop, (size_t)2 * nv));
        if (world > 1) {
            TRY(dev_alloc(s, &s->gath_rows, (size_t)world * 2 * nv * nv));
            TRY(dev_alloc(s, &s->gath_rhs, (size_t)world * nv));
            TRY(dev_alloc(s, &s->redA, (size_t)world * nv * nv));
            TRY(dev_alloc(s, &s->redB, (size_t)world * nv * nv));
            TRY(dev_alloc(s, &s->red_r, (size_t)world * nv));
            TRY(dev_alloc(s, &s->Zred, (size_t)(world + 1) * nv));
            TRY(dev_alloc(s, &s->topA, (size_t)nv * nv));
            TRY(dev_alloc(s, &s->topB, (size_t)nv * nv));
            TRY(dev_alloc(s, &s->top_r, (size_t)nv));
            s->p2p_pay = 2 * nv * nv > 8 ? 2 * nv * nv : 8;
            const size_t mb = p2p_mailbox_doubles(world, s->p2p_pay);
            TRY(dev_alloc(s, &s->mailbox, mb));
            TRY(dev_alloc(s, &s->iface_lu, (size_t)((world + 1) * nv) * ((world + 1) * nv)));
            TRY(dev_alloc(s, &s->iface_perm, (size_t)(world + 1) * nv));
            CUDA_TRY(cudaMemsetAsync(s->mailbox, 0, mb * sizeof(double), s->stream));
        }
        s->use_tps = (nv == 2) && (std::getenv("RST_B200_NO_TPS") == nullptr);
        if (s->use_tps) {
            // thread-per-slab levels (S = TPS_S): big levels one launch each, the rest inside the tail kernel
            int64_t ln = n;
            const double *lA = s->A, *lB = s->B, *lr = s->F;
            double *lZ = s->dY;
            s->tps_tail.nlevels = 0;
            while (ln > 1) {
                TpsLevel lv{};
                lv.n = ln;
                lv.nslabs = (ln + TPS_S - 1) / TPS_S;
                lv.A = lA; lv.B = lB; lv.r = lr; lv.Z = lZ;
                TRY(dev_alloc(s, &lv.fac, (size_t)(TPS_S - 1) * TpsCfg<2>::NE * lv.nslabs));
                TRY(dev_alloc(s, &lv.bits, (size_t)(TPS_S - 1) * lv.nslabs));
                TRY(dev_alloc(s, &lv.e, (size_t)(TPS_S - 1) * nv * lv.nslabs));
                if (lv.nslabs > 1) {
                    TRY(dev_alloc(s, &lv.A_next, (size_t)lv.nslabs * nv * nv));
                    TRY(dev_alloc(s, &lv.B_next, (size_t)lv.nslabs * nv * nv));
                    TRY(dev_alloc(s, &lv.r_next, (size_t)lv.nslabs * nv));
                    double *zn = nullptr;
                    TRY(dev_alloc(s, &zn, (size_t)(lv.nslabs + 1) * nv));
                    lv.Z_next = zn;
                } else {
                    lv.A_next = s->iface_rows;
                    lv.B_next = s->iface_rows + nv * nv;
                    lv.r_next = s->iface_rhs;
                    lv.Z_next = (world > 1) ? (s->Zred + (size_t)s->rank * nv) : s->Ztop;
                }
                if (ln > TPS_TAIL_ROWS) s->tps_big.push_back(lv);
                else {
                    if (s->tps_tail.nlevels >= TPS_MAX_TAIL_LEVELS) return fail(RST_B200_ERR_INVALID, "too many tail levels");
                    s->tps_tail.lv[s->tps_tail.nlevels++] = lv;
                }
                lA = lv.A_next; lB = lv.B_next; lr = lv.r_next;
                lZ = const_cast<double *>(lv.Z_next);
                ln = lv.nslabs;
            }
            s->tps_tail.left_mask = s->left_mask;
            s->tps_tail.right_mask = s->right_mask;
            s->tps_tail.Ztop = s->Ztop;
            s->tps_tail.do_top = (world == 1);
        } else
        // local condensation levels: n -> ceil(n/S) -> ... -> 1
        {
            int64_t ln = n;
            const double *lA = s->A, *lB = s->B, *lr = s->F;
            double *lZ = s->dY;
            int depth = 0;
            while (ln > 1) {
                LevelBuf lv{};
                const int S = default_slab(ln, depth == 0 ? d->slab : 0);
                TRY(build_level(s, &lv, ln, S, lA, lB, lr, lZ));
                const int64_t nn = lv.k.nslabs;
                if (nn > 1) {
                    TRY(dev_alloc(s, &lv.k.A_next, (size_t)nn * nv * nv));
                    TRY(dev_alloc(s, &lv.k.B_next, (size_t)nn * nv * nv));
                    TRY(dev_alloc(s, &lv.k.r_next, (size_t)nn * nv));
                    double *zn = nullptr;
                    TRY(dev_alloc(s, &zn, (size_t)(nn + 1) * nv));
                    lv.k.Z_next = zn;
                } else {  // last local level feeds the interface row
                    lv.k.A_next = s->iface_rows;
                    lv.k.B_next = s->iface_rows + nv * nv;
                    lv.k.r_next = s->iface_rhs;
                    lv.k.Z_next = (world > 1) ? (s->Zred + (size_t)s->rank * nv) : s->Ztop;
                }
                s->levels.push_back(lv);
                lA = lv.k.A_next;
                lB = lv.k.B_next;
                lr = lv.k.r_next;
                lZ = const_cast<double *>(lv.k.Z_next);
                ln = nn;
                ++depth;
            }
        }
        if (!s->use_tps && s->sv.tail_forward && !s->levels.empty()) {
            int tb = (int)s->levels.size();
            while (tb > 0 && s->levels[tb - 1].k.n <= ABD_TAIL_ROWS && (int)s->levels.size() - (tb - 1) <= ABD_TAIL_MAX_LEVELS) --tb;
            if (tb < (int)s->levels.size()) s->tail_begin = tb;
        }
        if (world > 1) {
            LevelBuf lv{};
            TRY(build_level(s, &lv, world, world, s->redA, s->redB, s->red_r, s->Zred));
            lv.k.A_next = s->topA;
            lv.k.B_next = s->topB;
            lv.k.r_next = s->top_r;
            lv.k.Z_next = s->Ztop;
            s->red = lv;
            s->have_red = true;
        }
        // reductions
        s->red_blocks = 148 * 4;
        {
            size_t nb_part = (size_t)s->red_blocks;
            if (s->use_tps && !s->tps_big.empty()) nb_part = std::max(nb_part, (size_t)((s->tps_big[0].nslabs + 31) / 32));
            TRY(dev_alloc(s, &s->d_partials, nb_part * 4));
        }
        TRY(dev_alloc(s, &s->d_scal, (size_t)16));  // two result slots: undamped step / first damping trial
        TRY(dev_alloc(s, &s->d_gscal, (size_t)8 * world));
        CUDA_TRY(cudaMallocHost((void **)&s->h_scal, 32 * sizeof(double)));
        std::memset(s->h_scal, 0, 32 * sizeof(double));
        TRY(dev_alloc(s, &s->d_ticket, (size_t)1));
        CUDA_TRY(cudaMemsetAsync(s->d_ticket, 0, sizeof(unsigned long long), s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        return RST_B200_OK;
    };
    rc = build();
    if (rc != RST_B200_OK) {
        rst_b200_destroy(s);
        return rc;
    }
    *out = s;
    return RST_B200_OK;
}

extern "C" void rst_b200_destroy(rst_b200_solver *s) {
    if (!s) return;
    {
        rst_b200_solver *expect = s;
        g_bound.compare_exchange_strong(expect, nullptr);
        if (t_bound == s) t_bound = nullptr;
    }
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto *cache : {&s->step_graphs, &s->carry_graphs, &s->jac_graphs, &s->frozen_graphs})
        for (auto &kv : *cache)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (void *p : s->peer_mapped) cudaIpcCloseMemHandle(p);
    for (void *p : s->owned) cudaFree(p);
    if (s->h_scal) cudaFreeHost(s->h_scal);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int64_t rst_b200_n_unknowns(const rst_b200_solver *s) { return s ? s->n * s->nv : 0; }
extern "C" int rst_b200_memcpy(void *dst, const void *src, size_t bytes, int kind) {
    if (!dst || !src || kind < 1 || kind > 3) return fail(RST_B200_ERR_INVALID, "bad memcpy arguments");
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(dst, src, bytes, (cudaMemcpyKind)kind));
    return RST_B200_OK;
}
extern "C" int rst_b200_host_alloc(size_t bytes, void **out) {
    if (!out || bytes == 0) return fail(RST_B200_ERR_INVALID, "bad host_alloc arguments");
    *out = nullptr;
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return RST_B200_OK;
}
extern "C" int rst_b200_host_free(void *p) {
    if (!p) return RST_B200_OK;
    CUDA_TRY(cudaFreeHost(p));
    return RST_B200_OK;
}
extern "C" int64_t rst_b200_launch_count(const rst_b200_solver *s) { return s ? s->launches : 0; }
extern "C" int rst_b200_sync(rst_b200_solver *s) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(s->device));   // the caller may be a fresh worker thread (one handle per thread: pipeline.py)
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return RST_B200_OK;
}
extern "C" int rst_b200_set_params(rst_b200_solver *s, const double *params, int n_params) {
    if (!s || n_params != s->np || (n_params > 0 && !params)) return fail(RST_B200_ERR_INVALID, "parameter count mismatch");
    CUDA_TRY(cudaSetDevice(s->device));
    if (s->np > 0) {
        CUDA_TRY(cudaMemcpyAsync(s->d_params, params, s->np * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        s->k_params.assign(params, params + s->np);
    }
    s->have_jac = s->factored = false;
    return RST_B200_OK;
}
extern "C" int rst_b200_set_allgather(rst_b200_solver *s, rst_b200_allgather_fn fn, void *ctx) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    s->allgather = fn;
    s->allgather_ctx = ctx;
    return RST_B200_OK;
}
extern "C" int rst_b200_dist_range(const rst_b200_solver *s, int64_t *b, int64_t *e) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (b) *b = s->j_begin;
    if (e) *e = s->j_begin + s->n;
    return RST_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// adaptive grid refinement hand-off (grid.cuh; reference: grid_api.rs:154, NR_Damp_solver_damped.rs:2202-2340)
// ------------------------------------------------------------------------------------------------
static void grid_free(rst_b200_solver *s, void *p) {
    if (!p) return;
    for (auto it = s->owned.begin(); it != s->owned.end(); ++it)
        if (*it == p) { s->owned.erase(it); break; }
    cudaFree(p);
}

extern "C" int rst_b200_new_grid(rst_b200_solver *s, int method, const double *params, double abs_tolerance,
                                 int64_t *n_points_new, int64_t *n_added) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (s->world != 1) return fail(RST_B200_ERR_UNSUPPORTED, "grid refinement runs on the single-GPU path (regrid, then re-partition)");
    if (method < RST_B200_GRID_DOUBLE_POINTS || method > RST_B200_GRID_SCI)
        return fail(RST_B200_ERR_UNSUPPORTED, "grid refinement method not supported (TwoPoint is outside the hot-path scope)");
    if (method != RST_B200_GRID_DOUBLE_POINTS && method != RST_B200_GRID_SCI && !params)
        return fail(RST_B200_ERR_INVALID, "grid refinement parameters missing");
    CUDA_TRY(cudaSetDevice(s->device));
    const int nv = s->nv;
    const int64_t nn = s->n + 1;
    cudaStream_t st = s->stream;
    if (!s->g_mark[0]) {
        TRY(dev_alloc(s, &s->g_mark[0], (size_t)nn));
        TRY(dev_alloc(s, &s->g_mark[1], (size_t)nn));
        TRY(dev_alloc(s, &s->g_pos, (size_t)nn));
        TRY(dev_alloc(s, &s->g_block_sum, (size_t)((nn + GRID_SCAN_BLOCK - 1) / GRID_SCAN_BLOCK) + 1));
        TRY(dev_alloc(s, &s->g_total, (size_t)1));
        TRY(dev_alloc(s, &s->g_ext, (size_t)4 * nv));
    }
    GridArgs g{s->Y, s->uniform ? nullptr : s->d_mesh, s->t0, s->h, nn, nv};
    const unsigned nb = (unsigned)((nn + 255) / 256);
    int cur = 0;
    if (method == RST_B200_GRID_DOUBLE_POINTS) {
        grid_mark_all_kernel<<<nb, 256, 0, st>>>(s->g_mark[0], nn);
        s->launches += 1;
    } else if (method == RST_B200_GRID_SCI) {
        // residual of the current (converged) state, as create_new_grid does (NR_Damp_solver_damped.rs:2235-2250)
        TRY(rst_b200_eval_residual(s, 0));
        unsigned *oob = s->d_flags + 1;
        CUDA_TRY(cudaMemsetAsync(oob, 0, sizeof(unsigned), st));
        const int64_t m_res = s->n * nv;
        const int64_t nt = std::max(m_res, nn);
        grid_mark_sci_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(s->F, m_res, abs_tolerance, s->g_mark[0], nn, oob);
        s->launches += 1;
        unsigned h_oob = 0;
        CUDA_TRY(cudaMemcpyAsync(&h_oob, oob, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h_oob)
            return fail(RST_B200_ERR_INVALID, "Sci refinement: a residual row beyond the mesh length exceeds the tolerance "
                                              "(the reference indexes the mesh with the residual row and panics here)");
    } else {
        // per-variable extrema of y and dy/dx
        grid_extrema_init_kernel<<<1, 256, 0, st>>>(s->g_ext, nv);
        int64_t want = std::min<int64_t>(nn * nv, (int64_t)148 * 8 * 256);
        const int64_t step = ((want + nv - 1) / nv) * nv;
        grid_extrema_kernel<<<(unsigned)((step + 255) / 256), 256, 0, st>>>(g, s->g_ext, step);
        s->launches += 1;
        if (method == RST_B200_GRID_EASY) {
            grid_mark_easy_kernel<<<nb, 256, 0, st>>>(g, s->g_ext, params[0], s->g_mark[0]);
            s->launches += 1;
        } else {
            const bool grcar = method == RST_B200_GRID_GRCAR_SMOOKE;
            const double d = params[0], gp = grcar ? params[1] : 0.0, C = grcar ? params[2] : params[1];
            CUDA_TRY(cudaMemsetAsync(s->g_mark[0], 0, (size_t)nn * sizeof(int), st));
            for (int v = 0; v < nv; ++v) {   // the merge across variables is order dependent: one pass per variable
                grid_mark_jump_kernel<<<nb, 256, 0, st>>>(g, s->g_ext, v, d, gp, grcar ? 1 : 0, s->g_mark[cur], s->g_mark[cur ^ 1]);
                cur ^= 1;
            }
            grid_mark_buffer_kernel<<<nb, 256, 0, st>>>(g, C, s->g_mark[cur], s->g_mark[cur ^ 1]);
            cur ^= 1;
            s->launches += nv + 1;
        }
    }
    s->g_mark_cur = cur;
    const int64_t nblk = (nn + GRID_SCAN_BLOCK - 1) / GRID_SCAN_BLOCK;
    grid_scan_local_kernel<<<(unsigned)nblk, GRID_SCAN_BLOCK, 0, st>>>(s->g_mark[cur], nn, s->g_pos, s->g_block_sum);
    grid_scan_blocks_kernel<<<1, 32, 0, st>>>(s->g_block_sum, nblk, s->g_total);
    grid_scan_add_kernel<<<(unsigned)nblk, GRID_SCAN_BLOCK, 0, st>>>(s->g_pos, nn, s->g_block_sum);
    s->launches += 3;
    int64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, s->g_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (total > s->g_capacity) {   // grow-only buffers: cudaMalloc / cudaFree cost milliseconds and synchronise the device
        grid_free(s, s->g_mesh_new);
        grid_free(s, s->g_Y_new);
        s->g_mesh_new = s->g_Y_new = nullptr;
        s->g_capacity = 0;
        const int64_t cap = total + total / 4;
        TRY(dev_alloc(s, &s->g_mesh_new, (size_t)cap));
        TRY(dev_alloc(s, &s->g_Y_new, (size_t)cap * nv));
        s->g_capacity = cap;
    }
    const int64_t ne = nn * nv;
    grid_construct_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(g, s->g_mark[cur], s->g_pos, method == RST_B200_GRID_SCI ? 1 : 0,
                                                                         s->g_mesh_new, s->g_Y_new);
    s->launches += 1;
    CUDA_TRY(cudaGetLastError());
    s->g_nodes_new = total;
    s->g_added = total - nn;
    if (n_points_new) *n_points_new = total;
    if (n_added) *n_added = total - nn;
    return RST_B200_OK;
}

extern "C" int rst_b200_new_grid_fetch(rst_b200_solver *s, double *mesh_out, size_t mesh_len, double *y_full_out, size_t y_len,
                                       int32_t *marks_out, size_t marks_len) {
    if (!s || s->g_added < 0) return fail(RST_B200_ERR_INVALID, "no refined grid: call rst_b200_new_grid first");
    CUDA_TRY(cudaSetDevice(s->device));
    if (mesh_out) {
        if (mesh_len != (size_t)s->g_nodes_new) return fail(RST_B200_ERR_INVALID, "mesh buffer length mismatch");
        CUDA_TRY(cudaMemcpyAsync(mesh_out, s->g_mesh_new, mesh_len * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    if (y_full_out) {
        if (y_len != (size_t)s->g_nodes_new * s->nv) return fail(RST_B200_ERR_INVALID, "solution buffer length mismatch");
        CUDA_TRY(cudaMemcpyAsync(y_full_out, s->g_Y_new, y_len * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    if (marks_out) {
        if (marks_len != (size_t)(s->n + 1)) return fail(RST_B200_ERR_INVALID, "marks buffer length mismatch");
        CUDA_TRY(cudaMemcpyAsync(marks_out, s->g_mark[s->g_mark_cur], marks_len * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return RST_B200_OK;
}

extern "C" int rst_b200_create_refined(rst_b200_solver *s, rst_b200_solver **out) {
    if (!s || !out) return fail(RST_B200_ERR_INVALID, "null handle");
    if (s->g_added < 0) return fail(RST_B200_ERR_INVALID, "no refined grid: call rst_b200_new_grid first");
    CUDA_TRY(cudaSetDevice(s->device));
    // the mesh is a create() argument (host); 8 bytes per node cross PCIe, the state stays on the device
    std::vector<double> mesh((size_t)s->g_nodes_new);
    CUDA_TRY(cudaMemcpyAsync(mesh.data(), s->g_mesh_new, mesh.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    rst_b200_desc d{};
    d.problem_key = s->pb->key;
    d.n_steps = s->g_nodes_new - 1;
    d.scheme = s->scheme;
    d.trap_time = s->trap_time;
    d.t0 = mesh.front();
    d.t_end = mesh.back();
    d.mesh = mesh.data();
    d.n_bc = (int)s->k_bc_var.size();
    d.bc_var = s->k_bc_var.data();
    d.bc_side = s->k_bc_side.data();
    d.bc_val = s->k_bc_val.data();
    d.params = s->k_params.empty() ? nullptr : s->k_params.data();
    d.n_params = s->np;
    d.lo = s->k_lo.data();
    d.hi = s->k_hi.data();
    d.rtol = s->k_rtol.data();
    d.device = s->device;
    d.stream = s->own_stream ? nullptr : (void *)s->stream;
    d.slab = s->k_slab;
    d.rank = 0;
    d.world = 1;
    rst_b200_solver *t = nullptr;
    TRY(rst_b200_create(&d, &t));
    cudaStreamSynchronize(t->stream);
    // interpolated full solution -> state of the new handle (device to device); both streams drained around the copy
    cudaError_t e = cudaMemcpyAsync(t->Y, s->g_Y_new, (size_t)s->g_nodes_new * s->nv * sizeof(double), cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        rst_b200_destroy(t);
        return fail(RST_B200_ERR_CUDA, std::string("refined state copy: ") + cudaGetErrorString(e));
    }
    *out = t;
    return RST_B200_OK;
}

extern "C" void *rst_b200_device_ptr(rst_b200_solver *s, int which) {
    if (!s) return nullptr;
    switch (which) {
        case 0: return s->Y;
        case 1: return s->F;
        case 2: return s->A;
        case 3: return s->B;
        case 4: return s->dY;
        case 5: return s->Yt;
        case 6: return s->iface_rows;
        case 7: return s->iface_rhs;
        case 8: return s->gath_rows;
        case 9: return s->gath_rhs;
        case 10: return s->d_scal;
        case 11: return s->d_gscal;
        default: return nullptr;
    }
}

// ---- state movement ---------------------------------------------------------------------------
// global reduced index of (global node g, component v), or -1 for a pinned slot (numeric_discretization.rs:80-106)
static inline int64_t reduced_index(const rst_b200_solver *s, int64_t g, int v) {
    const int nv = s->nv;
    if (g == 0) {
        if ((s->left_mask >> v) & 1ull) return -1;
        int k = 0;
        for (int c = 0; c < v; ++c) k += !((s->left_mask >> c) & 1ull);
        return k;
    }
    if (g == s->N) {
        if ((s->right_mask >> v) & 1ull) return -1;
        int k = 0;
        for (int c = 0; c < v; ++c) k += !((s->right_mask >> c) & 1ull);
        return s->N * (int64_t)nv - s->layout.nl + k;
    }
    return g * (int64_t)nv + v - s->layout.nl;
}

// The reduced unknown vector is the full node-major state with the pinned slots of node 0 and node N deleted
// (numeric_discretization.rs:80-106), so every interior node is one contiguous run in both orderings: the state moves
// with ONE large copy straight between the caller's buffer and HBM plus two nv-sized boundary copies.
static int move_state(rst_b200_solver *s, double *y, bool to_device) {
    CUDA_TRY(cudaSetDevice(s->device));
    const int nv = s->nv;
    const int nl = s->layout.nl;
    const int64_t g_lo = s->j_begin, g_hi = s->j_begin + s->n;          // local nodes [g_lo, g_hi]
    const int64_t own_hi = (s->rank == s->world - 1) ? g_hi : g_hi - 1;  // nodes this rank reports back
    const int64_t i_lo = g_lo < 1 ? 1 : g_lo;
    const int64_t i_hi_set = g_hi > s->N - 1 ? s->N - 1 : g_hi;          // interior nodes (set: incl. halo)
    const int64_t i_hi_get = own_hi > s->N - 1 ? s->N - 1 : own_hi;
    const int64_t i_hi = to_device ? i_hi_set : i_hi_get;
    if (i_hi >= i_lo) {
        double *dev = s->Y + (i_lo - g_lo) * nv;
        double *host = y + (i_lo * nv - nl);
        const size_t bytes = (size_t)(i_hi - i_lo + 1) * nv * sizeof(double);
        if (to_device) CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s->stream));
        else CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s->stream));
    }
    double edge[128];
    const bool has0 = (g_lo == 0), hasN = to_device ? (g_hi == s->N) : (own_hi == s->N);
    if (to_device) {
        if (has0) {
            for (int v = 0, k = 0; v < nv; ++v) edge[v] = ((s->left_mask >> v) & 1ull) ? s->left_val[v] : y[k++];
            CUDA_TRY(cudaMemcpyAsync(s->Y, edge, nv * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        }
        if (hasN) {
            const double *src = y + (s->N * (int64_t)nv - nl);
            for (int v = 0, k = 0; v < nv; ++v) edge[64 + v] = ((s->right_mask >> v) & 1ull) ? s->right_val[v] : src[k++];
            CUDA_TRY(cudaMemcpyAsync(s->Y + (s->N - g_lo) * nv, edge + 64, nv * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    } else {
        if (has0) CUDA_TRY(cudaMemcpyAsync(edge, s->Y, nv * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        if (hasN) CUDA_TRY(cudaMemcpyAsync(edge + 64, s->Y + (s->N - g_lo) * nv, nv * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (has0)
            for (int v = 0, k = 0; v < nv; ++v)
                if (!((s->left_mask >> v) & 1ull)) y[k++] = edge[v];
        if (hasN) {
            double *dst = y + (s->N * (int64_t)nv - nl);
            for (int v = 0, k = 0; v < nv; ++v)
                if (!((s->right_mask >> v) & 1ull)) dst[k++] = edge[64 + v];
        }
    }
    return RST_B200_OK;
}

extern "C" int rst_b200_set_state(rst_b200_solver *s, const double *y, size_t len) {
    if (!s || !y) return fail(RST_B200_ERR_INVALID, "null argument");
    if ((int64_t)len != s->N * (int64_t)s->nv) return fail(RST_B200_ERR_INVALID, "reduced state length mismatch");
    TRY(move_state(s, const_cast<double *>(y), true));
    s->have_jac = s->factored = false;
    return RST_B200_OK;
}

// world > 1: only the unknowns of the nodes this rank owns are written
extern "C" int rst_b200_get_state(rst_b200_solver *s, double *y, size_t len) {
    if (!s || !y) return fail(RST_B200_ERR_INVALID, "null argument");
    if ((int64_t)len != s->N * (int64_t)s->nv) return fail(RST_B200_ERR_INVALID, "reduced state length mismatch");
    return move_state(s, y, false);
}

extern "C" int rst_b200_get_full_solution(rst_b200_solver *s, double *full, size_t len) {
    if (!s || !full) return fail(RST_B200_ERR_INVALID, "null argument");
    if ((int64_t)len != (s->N + 1) * (int64_t)s->nv) return fail(RST_B200_ERR_INVALID, "full solution length mismatch");
    const int64_t owned = (s->rank == s->world - 1) ? s->n + 1 : s->n;
    CUDA_TRY(cudaMemcpyAsync(full + s->j_begin * s->nv, s->Y, (size_t)owned * s->nv * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return RST_B200_OK;
}

// ---- hot path ------------------------------------------------------------------------------------
static int assemble(rst_b200_solver *s, int which_state, bool with_j) {
    AsmArgs a{};
    a.Y = which_state ? s->Yt : s->Y;
    a.F = s->F;
    a.A = s->A;
    a.B = s->B;
    a.mesh = s->uniform ? nullptr : s->d_mesh;
    a.params = s->d_params;
    a.t0 = s->t0;
    a.h = s->h;
    a.n = s->n;
    a.j_offset = s->j_begin;
    a.trap_time = s->trap_time;
    s->pb->assemble(a, with_j, s->scheme == RST_B200_SCHEME_TRAPEZOID, s->stream);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return RST_B200_OK;
}

extern "C" int rst_b200_eval_residual(rst_b200_solver *s, int which_state) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    return assemble(s, which_state, false);
}
extern "C" int rst_b200_eval_jacobian(rst_b200_solver *s, int which_state) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    TRY(assemble(s, which_state, true));
    s->have_jac = true;
    s->factored = false;
    return RST_B200_OK;
}

static int factor_local_impl(rst_b200_solver *s, bool with_rhs);
extern "C" int rst_b200_factor_local(rst_b200_solver *s) { return factor_local_impl(s, false); }

// with_rhs: every level's factor kernel also eliminates the level's right-hand side (level 0: s->F) -- SolverVT::factor_rhs
static int factor_local_impl(rst_b200_solver *s, bool with_rhs) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (!s->have_jac) return fail(RST_B200_ERR_NOT_FACTORIZED, "no Jacobian has been evaluated");
    CUDA_TRY(cudaMemsetAsync(s->d_flags, 0, sizeof(unsigned), s->stream));
    if (s->use_tps) {
        for (auto &lv : s->tps_big) {
            const int tb = tps_block();
            const unsigned blocks = (unsigned)((lv.nslabs + tb - 1) / tb);
            if (lv.B == nullptr) tps_factor_kernel<2, TPS_S, true><<<blocks, tb, 0, s->stream>>>(lv, s->d_flags);
            else tps_factor_kernel<2, TPS_S, false><<<blocks, tb, 0, s->stream>>>(lv, s->d_flags);
            s->launches++;
        }
        const bool fused = fused_p2p(s);
        tps_tail_factor_kernel<2, TPS_S><<<1, 256, 0, s->stream>>>(s->tps_tail, iface_args(s, fused), 0ull /* sequence number lives on the device (P2pPeers::seq_dev) */, s->d_flags);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return RST_B200_OK;
    }
    for (auto &lv : s->levels) {
        if (with_rhs) s->sv.factor_rhs(lv.k, lv.ident_b, s->d_flags, s->stream);
        else s->sv.factor(lv.k, lv.ident_b, s->d_flags, s->stream);
        s->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return RST_B200_OK;
}

__global__ void split_rows_kernel(const double *gath, double *A, double *B, int world, int nvnv) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= world * nvnv) return;
    const int p = idx / nvnv, e = idx % nvnv;
    A[idx] = gath[(size_t)p * 2 * nvnv + e];
    B[idx] = gath[(size_t)p * 2 * nvnv + nvnv + e];
}

extern "C" int rst_b200_factor_reduced(rst_b200_solver *s) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (s->have_red) {
        const int nvnv = s->nv * s->nv;
        split_rows_kernel<<<(s->world * nvnv + 127) / 128, 128, 0, s->stream>>>(s->gath_rows, s->redA, s->redB, s->world, nvnv);
        s->sv.factor(s->red.k, false, s->d_flags, s->stream);
        s->launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    s->factored = true;
    return RST_B200_OK;
}

// the levels of the single-CTA tail launches (built per call: set_solution_buffer may have redirected level 0's Z)
static AbdTail make_tail(const rst_b200_solver *s, bool with_top) {
    AbdTail T{};
    T.nlev = (int)s->levels.size() - s->tail_begin;
    for (int i = 0; i < T.nlev; ++i) T.lv[i] = s->levels[s->tail_begin + i].k;
    T.do_top = with_top ? 1 : 0;
    T.C = s->iface_rows;
    T.D = s->iface_rows + s->nv * s->nv;
    T.g = s->iface_rhs;
    T.Ztop = s->Ztop;
    T.left_mask = s->left_mask;
    T.right_mask = s->right_mask;
    return T;
}

// all-gather of `count` doubles per rank: NVLink peer-memory kernel when the peers are mapped, host callback otherwise
static int exchange(rst_b200_solver *s, const double *send, double *recv, int64_t count) {
    if (s->p2p_enabled) {
        p2p_allgather_kernel<<<1, 256, 0, s->stream>>>(s->peers, send, (int)count, recv, s->rank, s->world, s->p2p_pay,
                                                       0ull /* sequence number lives on the device (P2pPeers::seq_dev) */, s->d_flags);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return RST_B200_OK;
    }
    if (!s->allgather) return fail(RST_B200_ERR_INVALID, "world > 1 needs rst_b200_p2p_open or rst_b200_set_allgather");
    if (s->allgather(s->allgather_ctx, send, recv, count, s->stream) != 0)
        return fail(RST_B200_ERR_CUDA, "all-gather callback failed");
    return RST_B200_OK;
}

extern "C" int rst_b200_p2p_get_handle(rst_b200_solver *s, void *handle_out, size_t handle_bytes) {
    if (!s || !handle_out || handle_bytes < sizeof(cudaIpcMemHandle_t)) return fail(RST_B200_ERR_INVALID, "bad handle buffer");
    if (!s->mailbox) return fail(RST_B200_ERR_INVALID, "world == 1 has no mailbox");
    CUDA_TRY(cudaStreamSynchronize(s->stream));   // the mailbox must be zeroed before any peer can see it
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, s->mailbox));
    std::memcpy(handle_out, &h, sizeof(h));
    return RST_B200_OK;
}

// protocol self-test on one device (p2p.cuh): `world` emulated ranks in one cooperative launch
extern "C" int rst_b200_p2p_selftest(int device, int world, int rounds, int pay, int absent_rank, double timeout_ms,
                                     int64_t *mismatches, int64_t *nan_words, int *timeout_flags) {
    if (world < 2 || world > P2P_MAX_WORLD || rounds < 1 || pay < 8) return fail(RST_B200_ERR_INVALID, "bad self-test arguments");
    CUDA_TRY(cudaSetDevice(device));
    int coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) return fail(RST_B200_ERR_UNSUPPORTED, "device cannot launch cooperatively");
    P2pSelfTest t{};
    std::vector<void *> owned;
    auto alloc = [&](size_t bytes) -> void * {
        void *q = nullptr;
        if (cudaMalloc(&q, bytes) != cudaSuccess) return nullptr;
        cudaMemset(q, 0, bytes);
        owned.push_back(q);
        return q;
    };
    int khz = 1965000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    bool ok = true;
    const size_t mb = p2p_mailbox_doubles(world, pay);
    double *boxes[P2P_MAX_WORLD] = {};
    for (int p = 0; p < world; ++p) ok = ok && (boxes[p] = (double *)alloc(mb * sizeof(double)));
    unsigned long long *seqs = (unsigned long long *)alloc(sizeof(unsigned long long) * world);
    t.send = (double *)alloc(sizeof(double) * world * pay);
    t.recv = (double *)alloc(sizeof(double) * world * world * pay);
    t.flags = (unsigned *)alloc(sizeof(unsigned) * 4 * world);
    t.mismatches = (unsigned long long *)alloc(2 * sizeof(unsigned long long));
    ok = ok && seqs && t.send && t.recv && t.flags && t.mismatches;
    int rc = RST_B200_OK;
    if (!ok) rc = fail(RST_B200_ERR_CUDA, "self-test allocation failed");
    else {
        t.nan_words = t.mismatches + 1;
        for (int r = 0; r < world; ++r) {
            for (int p = 0; p < world; ++p) t.peers[r].mailbox[p] = boxes[p];
            t.peers[r].seq_dev = seqs + r;
            t.peers[r].timeout_cycles = (long long)(timeout_ms * (double)khz);
        }
        t.world = world; t.pay = pay; t.rounds = rounds; t.absent = absent_rank;
        t.counts[0] = 8; t.counts[1] = pay / 2 > 0 ? pay / 2 : 1; t.counts[2] = pay;
        void *args[] = {&t};
        cudaError_t e = cudaLaunchCooperativeKernel((void *)p2p_selftest_kernel, dim3(world), dim3(256), args, 0, nullptr);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(RST_B200_ERR_CUDA, std::string("p2p self-test: ") + cudaGetErrorString(e));
        else {
            unsigned long long h[2] = {0, 0};
            std::vector<unsigned> hf(4 * world);
            cudaMemcpy(h, t.mismatches, sizeof(h), cudaMemcpyDeviceToHost);
            cudaMemcpy(hf.data(), t.flags, sizeof(unsigned) * 4 * world, cudaMemcpyDeviceToHost);
            if (mismatches) *mismatches = (int64_t)h[0];
            if (nan_words) *nan_words = (int64_t)h[1];
            int tf = 0;
            for (int r = 0; r < world; ++r) tf += (hf[4 * r + 2] & RST_FLAG_PEER_TIMEOUT) ? 1 : 0;
            if (timeout_flags) *timeout_flags = tf;
        }
    }
    for (void *q : owned) cudaFree(q);
    return rc;
}

// handles: [world][handle_stride bytes] in rank order (own entry ignored)
extern "C" int rst_b200_p2p_open(rst_b200_solver *s, const void *handles, size_t handle_stride) {
    if (!s || !handles || handle_stride < sizeof(cudaIpcMemHandle_t)) return fail(RST_B200_ERR_INVALID, "bad handle array");
    if (s->world > P2P_MAX_WORLD || !s->mailbox) return fail(RST_B200_ERR_INVALID, "unsupported world size for peer access");
    for (int p = 0; p < s->world; ++p) {
        if (p == s->rank) { s->peers.mailbox[p] = s->mailbox; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char *)handles + (size_t)p * handle_stride, sizeof(h));
        void *ptr = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_mapped.push_back(ptr);
        s->peers.mailbox[p] = (double *)ptr;
    }
    {   // device-side exchange counter
        unsigned long long *d_seq = nullptr;
        TRY(dev_alloc(s, &d_seq, (size_t)1));
        CUDA_TRY(cudaMemset(d_seq, 0, sizeof(unsigned long long)));
        s->peers.seq_dev = d_seq;
    }
    {   // spin budget of the exchange (default 30 s: rank skew from lazy peer mapping, module load or a host-side regrid on
        // one rank must not raise a spurious timeout); RST_B200_PEER_TIMEOUT_MS overrides it
        const char *e = std::getenv("RST_B200_PEER_TIMEOUT_MS");
        const double ms = e ? std::atof(e) : 30000.0;
        int khz = 1965000;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, s->device);
        s->peers.timeout_cycles = (long long)(ms * (double)khz);
    }
    s->p2p_enabled = true;
    P2pIface &I = s->iface;
    I.peers = s->peers;
    I.rank = s->rank; I.world = s->world; I.pay = s->p2p_pay;
    I.rows_send = s->iface_rows; I.rows_gath = s->gath_rows;
    I.rhs_send = s->iface_rhs; I.rhs_gath = s->gath_rhs;
    I.lu = s->iface_lu; I.perm = s->iface_perm; I.Zred = s->Zred;
    I.left_mask = s->left_mask; I.right_mask = s->right_mask;
    return RST_B200_OK;
}

static int factor_impl(rst_b200_solver *s, bool with_rhs);
extern "C" int rst_b200_factor(rst_b200_solver *s) { return factor_impl(s, false); }
static bool can_factor_with_rhs(const rst_b200_solver *s) { return s->sv.factor_rhs != nullptr && !s->use_tps && !s->levels.empty(); }

static int factor_impl(rst_b200_solver *s, bool with_rhs) {
    TRY(factor_local_impl(s, with_rhs));
    if (fused_p2p(s)) {  // exchange + interface factorisation already ran inside the tail kernel
        s->factored = true;
        return RST_B200_OK;
    }
    if (s->world > 1) {
        TRY(exchange(s, s->iface_rows, s->gath_rows, 2 * (int64_t)s->nv * s->nv));
    }
    return rst_b200_factor_reduced(s);
}

extern "C" int rst_b200_solve_local_forward(rst_b200_solver *s) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (!s->factored) return fail(RST_B200_ERR_NOT_FACTORIZED, "solve before factor");
    if (s->use_tps) {
        for (auto &lv : s->tps_big) {
            launch_pdl(tps_forward_kernel<2, TPS_S>, (unsigned)((lv.nslabs + tps_block() - 1) / tps_block()), (unsigned)tps_block(), s->stream, lv);
            s->launches++;
        }
        if (s->world > 1 && !fused_p2p(s)) {  // otherwise the tail's forward sweep is fused with closure + backward sweep
            launch_pdl(tps_tail_solve_kernel<2, TPS_S>, 1u, 256u, s->stream, s->tps_tail, iface_args(s, false), 0ull, 1, s->d_flags);
            s->launches++;
        }
        CUDA_TRY(cudaGetLastError());
        return RST_B200_OK;
    }
    const int n_plain = s->tail_begin >= 0 ? s->tail_begin : (int)s->levels.size();
    for (int i = 0; i < n_plain; ++i) {
        s->sv.forward(s->levels[i].k, s->stream);
        s->launches++;
    }
    if (s->tail_begin >= 0) {
        s->sv.tail_forward(make_tail(s, false), s->stream);
        s->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return RST_B200_OK;
}

extern "C" int rst_b200_solve_reduced_backward(rst_b200_solver *s) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (s->use_tps && (s->world == 1 || fused_p2p(s))) {
        const bool fused = fused_p2p(s);
        // forward sweeps -> closure (1 GPU) or NVLink interface exchange + solve (multi GPU) -> backward sweeps: ONE launch
        launch_pdl(tps_tail_solve_kernel<2, TPS_S>, 1u, 256u, s->stream, s->tps_tail, iface_args(s, fused), 0ull /* sequence number lives on the device (P2pPeers::seq_dev) */,
                   fused ? 13 : 7, s->d_flags);
        s->launches++;
        for (int i = (int)s->tps_big.size() - 1; i >= 0; --i) {
            const TpsLevel &lv = s->tps_big[i];
            const int tb = tps_block();
            const unsigned blocks = (unsigned)((lv.nslabs + tb - 1) / tb);
            if (i == 0 && s->fuse_reduce_state >= 0) {  // level 0: damping reductions ride along with the sweep
                TpsReduce R{};
                R.Y = s->fuse_reduce_state ? s->Yt : s->Y;
                R.lo = s->d_lo; R.hi = s->d_hi; R.rtol = s->d_rtol;
                R.left_mask = s->left_mask; R.right_mask = s->right_mask;
                R.has_left = (s->rank == 0); R.has_right = (s->rank == s->world - 1);
                R.n_own = (s->rank == s->world - 1) ? s->n + 1 : s->n;
                R.partials = s->d_partials;
                launch_pdl(tps_backward_reduce_kernel<2, TPS_S>, blocks, (unsigned)tb, s->stream, lv, R);
                s->reduce_ready = (int)blocks;
                s->reduce_ready_state = s->fuse_reduce_state;
            } else {
                launch_pdl(tps_backward_kernel<2, TPS_S>, blocks, (unsigned)tb, s->stream, lv);
            }
            s->launches++;
        }
        s->fuse_reduce_state = -1;
        CUDA_TRY(cudaGetLastError());
        return RST_B200_OK;
    }
    const double *C, *D, *g;
    if (s->have_red) {
        CUDA_TRY(cudaMemcpyAsync(s->red_r, s->gath_rhs, (size_t)s->world * s->nv * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        s->sv.forward(s->red.k, s->stream);
        s->launches++;
        C = s->topA; D = s->topB; g = s->top_r;
    } else {
        C = s->iface_rows; D = s->iface_rows + s->nv * s->nv; g = s->iface_rhs;
    }
    const bool top_in_tail = s->tail_begin >= 0 && !s->have_red && !s->use_tps;   // closure + small levels: one launch
    if (!top_in_tail) {
        s->sv.top(C, D, g, s->Ztop, s->left_mask, s->right_mask, s->d_flags, s->nv, s->stream);
        s->launches++;
    }
    if (s->have_red) {
        s->sv.backward(s->red.k, s->stream);
        s->launches++;
    }
    if (s->use_tps) {
        launch_pdl(tps_tail_solve_kernel<2, TPS_S>, 1u, 256u, s->stream, s->tps_tail, iface_args(s, false), 0ull, 4, s->d_flags);
        s->launches++;
        for (int i = (int)s->tps_big.size() - 1; i >= 0; --i) {
            const TpsLevel &lv = s->tps_big[i];
            const int tb = tps_block();
            const unsigned blocks = (unsigned)((lv.nslabs + tb - 1) / tb);
            if (i == 0 && s->fuse_reduce_state >= 0) {  // level 0: damping reductions ride along with the sweep
                TpsReduce R{};
                R.Y = s->fuse_reduce_state ? s->Yt : s->Y;
                R.lo = s->d_lo; R.hi = s->d_hi; R.rtol = s->d_rtol;
                R.left_mask = s->left_mask; R.right_mask = s->right_mask;
                R.has_left = (s->rank == 0); R.has_right = (s->rank == s->world - 1);
                R.n_own = (s->rank == s->world - 1) ? s->n + 1 : s->n;
                R.partials = s->d_partials;
                launch_pdl(tps_backward_reduce_kernel<2, TPS_S>, blocks, (unsigned)tb, s->stream, lv, R);
                s->reduce_ready = (int)blocks;
                s->reduce_ready_state = s->fuse_reduce_state;
            } else {
                launch_pdl(tps_backward_kernel<2, TPS_S>, blocks, (unsigned)tb, s->stream, lv);
            }
            s->launches++;
        }
        s->fuse_reduce_state = -1;
        CUDA_TRY(cudaGetLastError());
        return RST_B200_OK;
    }
    if (s->tail_begin >= 0) {
        s->sv.tail_backward(make_tail(s, top_in_tail), s->d_flags, s->stream);
        s->launches++;
    }
    for (int i = (s->tail_begin >= 0 ? s->tail_begin : (int)s->levels.size()) - 1; i >= 0; --i) {
        // (a level-0 sweep with the damping reductions fused in, as the nv = 2 path has, was measured 80 us SLOWER per solve at
        // nv = 12 than sweep + newton_reduce_kernel: the extra state loads and registers cost the sweep more HBM efficiency
        // than the separate 58 us pass costs)
        s->sv.backward(s->levels[i].k, s->stream);
        s->launches++;
    }
    s->fuse_reduce_state = -1;
    CUDA_TRY(cudaGetLastError());
    return RST_B200_OK;
}

static int solve_impl(rst_b200_solver *s, bool skip_local_forward);
extern "C" int rst_b200_solve(rst_b200_solver *s) { return solve_impl(s, false); }

extern "C" int rst_b200_factor_solve(rst_b200_solver *s) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    const bool with_rhs = can_factor_with_rhs(s);
    TRY(factor_impl(s, with_rhs));
    return solve_impl(s, with_rhs);
}

// skip_local_forward: the factorisation that just ran eliminated this right-hand side as well (factor_rhs): e and r_next of
// every local level are already what the forward sweeps would write
static int solve_impl(rst_b200_solver *s, bool skip_local_forward) {
    if (skip_local_forward) {
        if (!s || !s->factored) return fail(RST_B200_ERR_NOT_FACTORIZED, "solve before factor");
    } else
    TRY(rst_b200_solve_local_forward(s));
    if (s->world > 1 && !fused_p2p(s)) {
        TRY(exchange(s, s->iface_rhs, s->gath_rhs, (int64_t)s->nv));
    }
    return rst_b200_solve_reduced_backward(s);
}

__global__ void combine_scalars_kernel(const double *g, int world, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double a0 = 0.0, a1 = 0.0, a2 = 1.0, a3 = 0.0, a4 = 0.0;
    for (int p = 0; p < world; ++p) {
        a0 += g[8 * p]; a1 += g[8 * p + 1]; a2 = fmin(a2, g[8 * p + 2]); a3 = fmax(a3, g[8 * p + 3]); a4 = fmax(a4, g[8 * p + 4]);
    }
    out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3; out[4] = a4;
}

extern "C" int rst_b200_reduce_local(rst_b200_solver *s, int which_state) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    ReduceArgs a{};
    a.Y = which_state ? s->Yt : s->Y;
    a.dY = s->sol;
    // NRBVP::step scans F and the step for NaN (NR_Damp_solver_damped.rs:1821-1826, :1839-1844).  Every entry of F is the
    // right-hand side of exactly one pivot row on some level and enters that row's solution component additively, and NaN
    // survives every arithmetic operation (also NaN * 0): a NaN in F is a NaN in the step, which the scan of the step reports
    // with the same status.  Reading F again (8 nv bytes per interval and solve) would add nothing.
    a.F = nullptr;
    a.lo = s->d_lo; a.hi = s->d_hi; a.rtol = s->d_rtol;
    // ranks own nodes [0, n) of their slab; the last rank also owns global node N
    a.n_nodes = (s->rank == s->world - 1) ? s->n + 1 : s->n;
    a.nv = s->nv;
    a.left_mask = s->left_mask; a.right_mask = s->right_mask;
    a.has_left = (s->rank == 0); a.has_right = (s->rank == s->world - 1);
    a.partials = s->d_partials;
    a.n_rows = s->n;
    int nblocks = s->red_blocks;
    if (s->reduce_ready > 0 && s->reduce_ready_state == (which_state ? 1 : 0)) {
        nblocks = s->reduce_ready;  // produced by the fused level-0 backward sweep of the solve that just ran
        s->launches--;
    } else {
        s->sv.reduce(a, s->red_blocks, s->stream);
    }
    s->reduce_ready = 0;
    const bool fused_red = s->p2p_enabled && s->world > 1;
    newton_reduce_final_kernel<<<1, 256, 0, s->stream>>>(s->d_partials, nblocks, s->d_scal + 8 * s->scal_slot, s->d_flags, s->peers, s->rank,
                                                         fused_red ? s->world : 1, s->p2p_pay, 0ull /* sequence number lives on the device (P2pPeers::seq_dev) */, s->d_gscal);
    s->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return RST_B200_OK;
}

extern "C" int rst_b200_reduce_finish(rst_b200_solver *s, rst_b200_scalars *out) {
    if (!s || !out) return fail(RST_B200_ERR_INVALID, "null argument");
    const double *src = s->d_scal;
    if (s->world > 1 && !s->p2p_enabled) {
        combine_scalars_kernel<<<1, 32, 0, s->stream>>>(s->d_gscal, s->world, s->d_scal);
        s->launches++;
    }
    CUDA_TRY(cudaMemcpyAsync(s->h_scal, src, 5 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    out->sumsq_step = s->h_scal[0];
    out->tol_sum = s->h_scal[1];
    out->fbound = s->h_scal[2];
    out->nan_flag = s->h_scal[3];
    out->device_flags = s->h_scal[4];
    return RST_B200_OK;
}

extern "C" int rst_b200_reduce(rst_b200_solver *s, int which_state, rst_b200_scalars *out) {
    TRY(rst_b200_reduce_local(s, which_state));
    if (s->world > 1 && !s->p2p_enabled) {
        TRY(exchange(s, s->d_scal, s->d_gscal, 8));
    }
    return rst_b200_reduce_finish(s, out);
}

extern "C" int rst_b200_make_trial(rst_b200_solver *s, double lambda) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    const int64_t total = (s->n + 1) * (int64_t)s->nv;
    newton_trial_kernel<<<s->red_blocks, 256, 0, s->stream>>>(s->Y, s->dY, s->Yt, lambda, total);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return RST_B200_OK;
}

extern "C" int rst_b200_accept_trial(rst_b200_solver *s) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    std::swap(s->Y, s->Yt);
    return RST_B200_OK;
}

extern "C" int rst_b200_checkpoint(rst_b200_solver *s, int restore) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    const size_t bytes = (size_t)(s->n + 1) * s->nv * sizeof(double);
    if (!s->Ysave) TRY(dev_alloc(s, &s->Ysave, (size_t)(s->n + 1) * s->nv));
    if (restore) {
        CUDA_TRY(cudaMemcpyAsync(s->Y, s->Ysave, bytes, cudaMemcpyDeviceToDevice, s->stream));
        s->have_jac = s->factored = false;
    } else {
        CUDA_TRY(cudaMemcpyAsync(s->Ysave, s->Y, bytes, cudaMemcpyDeviceToDevice, s->stream));
    }
    return RST_B200_OK;
}

// the level-0 node values ARE the solution of the linear solve: redirect them to `buf`
static void set_solution_buffer(rst_b200_solver *s, double *buf) {
    s->sol = buf;
    if (s->use_tps) {
        if (!s->tps_big.empty()) s->tps_big[0].Z = buf;
        else s->tps_tail.lv[0].Z = buf;
    } else if (!s->levels.empty()) {
        s->levels[0].k.Z = buf;
    }
}

// dY and dY1 swap roles whenever an accepted trial is carried into the next iteration; drivers that cache launch sequences by the
// state buffer only (frozen loop) and every new Newton solve start from the home assignment (the contents are dead by then)
static void normalize_step_buffers(rst_b200_solver *s) {
    if (s->dY != s->dY_home) {
        std::swap(s->dY, s->dY1);
        set_solution_buffer(s, s->dY);
    }
}

// ---- damped Newton -------------------------------------------------------------------------------
static int check_scalars(const rst_b200_scalars &sc) {
    if (sc.nan_flag != 0.0 || (((unsigned)sc.device_flags) & RST_FLAG_NAN_F))
        return fail(RST_B200_ERR_NAN, "NaN in residual or Newton step");
    if (((unsigned)sc.device_flags) & RST_FLAG_PEER_TIMEOUT) return fail(RST_B200_ERR_CUDA, "a peer rank never arrived at the interface exchange");
    if (((unsigned)sc.device_flags) & RST_FLAG_ZERO_PIVOT) return fail(RST_B200_ERR_ZERO_PIVOT, "zero pivot in block elimination");
    return RST_B200_OK;
}

// NRBVP::step (:1782-1847) for state `which`: dY = J^{-1} F(y)
// rhs_done: F(y) of this state was computed by the Jacobian pass and eliminated by the factorisation (rhs_in_factors)
static int newton_step(rst_b200_solver *s, int which, bool rhs_done = false) {   // enqueue-only; the caller counts the linear solve
    if (!rhs_done) TRY(rst_b200_eval_residual(s, which));
    s->fuse_reduce_state = which;  // the reduction that follows this solve refers to state `which`
    TRY(solve_impl(s, rhs_done));
    s->fuse_reduce_state = -1;
    return RST_B200_OK;
}

// Run `body` (enqueue-only work on s->stream with launch-invariant arguments) through a CUDA graph cache: the first time a
// key is seen the body runs directly, the second time it is captured, instantiated and launched, afterwards replayed.
template <class Body>
static int run_graphed(rst_b200_solver *s, std::map<uintptr_t, rst_b200_solver::GraphEntry> &cache, uintptr_t key, bool enabled,
                       bool *replayed, Body body) {
    *replayed = false;
    if (!enabled) return body();
    auto &ge = cache[key];
    if (ge.exec) {
        CUDA_TRY(cudaGraphLaunch(ge.exec, s->stream));
        s->launches += ge.launches;
        *replayed = true;
        return RST_B200_OK;
    }
    if (ge.seen < 0 || ge.seen++ < 1) return body();   // seen < 0: capture failed once on this key, stay on direct launches
    const int64_t launches0 = s->launches;
    if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        (void)cudaGetLastError();
        ge.seen = -1;
        return body();
    }
    const int rc = body();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    if (rc != RST_B200_OK) {   // a genuine error of the sequence itself (nothing has run)
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        return rc;
    }
    cudaError_t ei = e;
    if (e == cudaSuccess && graph) ei = cudaGraphInstantiate(&ge.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ei != cudaSuccess || !ge.exec) {   // not capturable here (e.g. a stream that cannot be captured): run it directly from now on
        (void)cudaGetLastError();
        ge.exec = nullptr;
        ge.seen = -1;
        s->launches = launches0;
        return body();
    }
    ge.launches = s->launches - launches0;
    CUDA_TRY(cudaGraphLaunch(ge.exec, s->stream));   // capture records, it does not run
    return RST_B200_OK;
}

static int recalc_jacobian(rst_b200_solver *s, rst_b200_newton_stats *st) {
    const bool graphs_on = std::getenv("RST_B200_NO_GRAPH") == nullptr;   // read per call: tests toggle it
    bool replayed = false;   // fixed launch sequence with fixed arguments per state buffer
    TRY(run_graphed(s, s->jac_graphs, (uintptr_t)s->Y, (s->world == 1 || s->p2p_enabled) && graphs_on, &replayed, [&]() -> int {
        TRY(rst_b200_eval_jacobian(s, 0));        // F(y) and J(y) in one pass
        TRY(factor_impl(s, can_factor_with_rhs(s)));
        return RST_B200_OK;
    }));
    if (replayed) s->have_jac = s->factored = true;
    s->rhs_in_factors = can_factor_with_rhs(s);
    st->jac_recalcs++;
    st->factorizations++;
    return RST_B200_OK;
}

// damped_step (:1862-2004).  Returns RST_B200_OK and the reference's status code in *status.
static void unpack_scalars(const double *h, rst_b200_scalars *sc) {
    sc->sumsq_step = h[0]; sc->tol_sum = h[1]; sc->fbound = h[2]; sc->nan_flag = h[3]; sc->device_flags = h[4];
}

// `carry_slot` (in, >= 0): the previous iteration accepted a trial and the factors are unchanged, so this iteration's undamped
// step and its scalars are that trial's (newton_trial_carry_kernel) and are not recomputed; the value says which device
// scalar slot holds them (1: speculative first trial, 0: a later trial of the damping loop).
// `*accepted_slot` (out): the same for the trial this iteration accepted, -1 if it cannot be carried.
static int damped_step(rst_b200_solver *s, const rst_b200_newton_opts *o, rst_b200_newton_stats *st, int *status, int carry_slot = -1,
                       int *accepted_slot = nullptr) {
    rst_b200_scalars sc{}, sc1{};
    if (accepted_slot) *accepted_slot = -1;
    bool carry = carry_slot >= 0;
    // trial steps go to dY1: the undamped step in dY is re-used by every damping trial (:1934, :1971)
    struct Restore {
        rst_b200_solver *s;
        ~Restore() { set_solution_buffer(s, s->dY); s->scal_slot = 0; }
    } restore{s};
    // Device-only exchange (1 GPU, or NVLink mailboxes): the first damping trial uses lambda = fbound, which is already
    // on the device, so undamped step + first trial are enqueued back to back and the host synchronises ONCE.
    const bool speculative = (s->world == 1 || s->p2p_enabled) && o->max_damp_iter > 0;
    carry = carry && speculative;
    const bool rhs_done = s->rhs_in_factors && !carry;   // first solve after recalc_jacobian: no residual pass, no forward sweeps
    s->rhs_in_factors = false;
    if (carry) std::swap(s->dY, s->dY1);                 // the accepted trial's step IS this iteration's undamped step
    const uintptr_t step_parity = s->dY != s->dY_home ? 2u : 0u;   // the graphs bake the buffer addresses in
    s->scal_slot = 0;
    // The speculative sequence is ~25-45 launches with fixed arguments: small meshes are launch bound, so it is captured
    // into a CUDA graph (per Y / Y_trial parity) the second time it runs and replayed with one launch afterwards.
    const bool graphs_on = std::getenv("RST_B200_NO_GRAPH") == nullptr;   // read per call: tests toggle it
    {
        bool replayed = false;
        TRY(run_graphed(s, carry ? s->carry_graphs : s->step_graphs, (uintptr_t)s->Y | step_parity | (rhs_done || (carry && carry_slot == 0) ? 1u : 0u), speculative && graphs_on, &replayed, [&]() -> int {
            const int64_t total = (s->n + 1) * (int64_t)s->nv;
            if (!carry) {
                TRY(newton_step(s, 0, rhs_done));
                if (!speculative) return RST_B200_OK;
                TRY(rst_b200_reduce_local(s, 0));
                newton_trial_dev_kernel<<<s->red_blocks, 256, 0, s->stream>>>(s->Y, s->dY, s->Yt, s->d_scal, total);
            } else {
                newton_trial_carry_kernel<<<s->red_blocks, 256, 0, s->stream>>>(s->Y, s->dY, s->Yt, s->d_scal, total, carry_slot);
            }
            s->launches++;
            set_solution_buffer(s, s->dY1);
            s->scal_slot = 1;
            TRY(newton_step(s, 1));
            TRY(rst_b200_reduce_local(s, 1));
            s->scal_slot = 0;
            publish_scalars_kernel<<<1, 32, 0, s->stream>>>(s->d_scal, s->h_scal, s->d_ticket,
                                                          reinterpret_cast<volatile unsigned long long *>(s->h_scal + 16));
            s->launches++;
            return RST_B200_OK;
        }));
        st->linear_solves += speculative ? 2 : 1;   // the reference's count: counted here, not in the (possibly captured / replayed) sequence
        if (carry) st->solves_reused++;
        if (replayed) {
            s->reduce_ready = 0;
            s->fuse_reduce_state = -1;
        }
    }
    if (speculative) {
        // wait for the ticket of this sequence instead of synchronising the stream
        const unsigned long long want = ++s->ticket_expected;
        volatile unsigned long long *tk = reinterpret_cast<volatile unsigned long long *>(s->h_scal + 16);
        for (unsigned spin = 0; *tk < want; ++spin) {
            if ((spin & 0x3fff) == 0x3fff) {   // every ~16k polls: has the stream died?
                const cudaError_t q = cudaStreamQuery(s->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady)
                    return fail(RST_B200_ERR_CUDA, std::string("damped step: ") + cudaGetErrorString(q));
                if (q == cudaSuccess && *tk < want) return fail(RST_B200_ERR_CUDA, "damped step finished without publishing its scalars");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        unpack_scalars(s->h_scal, &sc);
        unpack_scalars(s->h_scal + 8, &sc1);
    } else {
        TRY(rst_b200_reduce(s, 0, &sc));
    }
    TRY(check_scalars(sc));
    const double fbound = sc.fbound;
    st->last_fbound = fbound;
    if (std::isnan(fbound) || std::isinf(fbound)) return fail(RST_B200_ERR_NAN, "boundary step factor is not finite");
    double lambda = 1.0 * fbound;
    if (fbound < 1e-10) {
        if (speculative) st->linear_solves--;  // the reference never takes that trial (:1898-1905)
        *status = -3;
        return RST_B200_OK;
    }
    const double s0 = std::sqrt(sc.sumsq_step);
    st->last_s0 = s0;
    int k = 0;
    bool accepted = false;
    double s1 = 0.0, conv = 0.0;
    set_solution_buffer(s, s->dY1);
    while (k < o->max_damp_iter) {
        if (!(speculative && k == 0)) {
            TRY(rst_b200_make_trial(s, lambda));
            TRY(newton_step(s, 1));
            st->linear_solves++;
            TRY(rst_b200_reduce(s, 1, &sc1));
        }
        TRY(check_scalars(sc1));
        s1 = std::sqrt(sc1.sumsq_step);
        conv = sc1.tol_sum > o->abs_tol ? sc1.tol_sum : o->abs_tol;  // convergence_condition, BVP_utils_damped.rs:209-222
        st->last_lambda = lambda;
        if (s1 < 1.0 || s1 < s0) {                                // :1962
            accepted = true;
            if (accepted_slot) *accepted_slot = !speculative ? -1 : (k == 0 ? 1 : 0);   // later trials reduce into slot 0
            break;
        }
        lambda = lambda / std::pow(2.0, (double)k + o->damp_factor); // :1971
        k += 1;
    }
    st->last_s1 = s1;
    st->last_conv = conv;
    if (accepted) *status = (s1 > conv) ? 0 : 1;                   // :1983-1998
    else *status = -2;
    return RST_B200_OK;
}

extern "C" int rst_b200_newton_iterate(rst_b200_solver *s, const rst_b200_newton_opts *o, int recalc,
                                       rst_b200_newton_stats *st) {
    if (!s || !o || !st) return fail(RST_B200_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    if (recalc || !s->factored) TRY(recalc_jacobian(s, st));
    st->iterations++;
    int status = 0;
    TRY(damped_step(s, o, st, &status));
    if (status == 0 || status == 1) TRY(rst_b200_accept_trial(s));
    st->status = status;
    return RST_B200_OK;
}

extern "C" int rst_b200_newton_solve(rst_b200_solver *s, const rst_b200_newton_opts *o, rst_b200_newton_stats *st) {
    if (!s || !o || !st) return fail(RST_B200_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->device));
    std::memset(st, 0, sizeof(*st));
    const auto t_begin = std::chrono::steady_clock::now();
    // Y and Y_trial swap roles with every accepted iteration, and the CUDA-graph caches are keyed by the buffer Y lives in: after
    // a solve with an odd number of iterations the next one would meet every key in its other flavour (and pay the captures a
    // second time).  Start every solve from the home buffer: one device-to-device copy of the state when the roles are swapped.
    normalize_step_buffers(s);
    if (s->Y != s->Y_home && s->Yt == s->Y_home) {
        CUDA_TRY(cudaMemcpyAsync(s->Y_home, s->Y, (size_t)(s->n + 1) * s->nv * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
        std::swap(s->Y, s->Yt);
    }
    bool have_jac = false, recalc = true;
    // carry: the accepted first trial of the previous iteration IS this iteration's undamped step (same y, same factors)
    static const bool carry_on = std::getenv("RST_B200_NO_CARRY") == nullptr;
    int carry = -1;     // scalar slot of the accepted trial to carry over, -1: none
    int m = 0, n_jac_reeval = 0, it = 0, status = 0;
    while (it < o->max_iterations) {
        recalc = (!have_jac || m > o->max_jac) ? true : recalc;   // jac_recalc, BVP_utils_damped.rs:185-207
        if (recalc) {                                              // recalc_jacobian :1746-1771
            int rc = recalc_jacobian(s, st);
            if (rc != RST_B200_OK) { st->status = rc; return rc; }
            have_jac = true;
            m = 0;
            carry = -1;
        }
        m += 1;
        it += 1;
        st->iterations++;
        int accepted_slot = -1;
        int rc = damped_step(s, o, st, &status, carry_on ? carry : -1, &accepted_slot);
        carry = -1;
        if (rc != RST_B200_OK) { st->status = rc; return rc; }
        if (status == 0) {
            TRY(rst_b200_accept_trial(s));
            recalc = false;
            carry = accepted_slot;
        } else if (status == 1) {
            TRY(rst_b200_accept_trial(s));
            break;
        } else {                                                   // :2119-2136
            if (m > 1) {
                recalc = true;
                if (n_jac_reeval > 3) break;
                n_jac_reeval += 1;
            } else break;
        }
    }
    // every damped step ends with the host holding that step's scalars, i.e. the stream has drained: wall clock is exact and
    // saves an event record + synchronise per solve
    st->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    st->status = status;
    return RST_B200_OK;
}

// ---- frozen-Jacobian Newton -----------------------------------------------------
// NR_Damp_solver_frozen.rs:1035-1096 (iteration) and :1105-1148 (try_main_loop); strategies BVP_utils.rs:362-480.
static bool frozen_recalc(const rst_b200_frozen_opts *o, bool have_jac, int m, double error, double error_old) {
    const double *sp = o->strategy_params;
    switch (o->strategy) {
        case RST_B200_FROZEN_NAIVE: return true;
        case RST_B200_FROZEN_FROZEN_NAIVE: return !have_jac;
        case RST_B200_FROZEN_EVERY_M: return !have_jac || m > (int)sp[0];
        case RST_B200_FROZEN_AT_HIGH_NORM: return error > sp[0];
        case RST_B200_FROZEN_AT_LOW_SPEED: return sp[0] * error_old < error;
        case RST_B200_FROZEN_COMPLEX: return (error > sp[1]) || (m > (int)sp[0]) || (sp[2] * error_old < error);
        default: return true;
    }
}

extern "C" int rst_b200_newton_frozen(rst_b200_solver *s, const rst_b200_frozen_opts *o, rst_b200_newton_stats *st) {
    if (!s || !o || !st) return fail(RST_B200_ERR_INVALID, "null argument");
    if (o->strategy < 0 || o->strategy > RST_B200_FROZEN_COMPLEX) return fail(RST_B200_ERR_INVALID, "unknown strategy");
    normalize_step_buffers(s);
    std::memset(st, 0, sizeof(*st));
    const auto t_begin = std::chrono::steady_clock::now();
    bool jac_recalc = true, have_jac = false;
    int m = 0, it = 0, status = 0;
    double error_old = 0.0;
    rst_b200_scalars sc{};
    // device-only exchange (1 GPU or NVLink mailboxes): one iteration is an enqueue-only sequence with launch-invariant
    // arguments -> CUDA graph per (state buffer, refresh-or-not), scalars published to mapped host memory with a ticket
    const bool graphs_on = std::getenv("RST_B200_NO_GRAPH") == nullptr;   // read per call: tests toggle it
    const bool fast = s->world == 1 || s->p2p_enabled;
    while (it < o->max_iterations) {
        st->iterations++;
        if (jac_recalc) {
            have_jac = true;
            m = 0;
            st->jac_recalcs++;
            st->factorizations++;
        } else {
            m += 1;
        }
        st->linear_solves++;
        if (fast) {
            bool replayed = false;
            const bool refresh = jac_recalc;
            TRY(run_graphed(s, s->frozen_graphs, (uintptr_t)s->Y | (refresh ? 1u : 0u), graphs_on, &replayed, [&]() -> int {
                if (refresh) {  // F(y) and J(y) in one fused pass, then the (cached) factorisation
                    TRY(rst_b200_eval_jacobian(s, 0));
                    TRY(rst_b200_factor(s));
                } else {
                    TRY(rst_b200_eval_residual(s, 0));
                }
                TRY(rst_b200_solve(s));
                s->scal_slot = 0;
                TRY(rst_b200_reduce_local(s, 0));
                TRY(rst_b200_make_trial(s, 1.0));  // y_new = y - delta
                publish_scalars_kernel<<<1, 32, 0, s->stream>>>(s->d_scal, s->h_scal, s->d_ticket,
                                                              reinterpret_cast<volatile unsigned long long *>(s->h_scal + 16));
                s->launches++;
                return RST_B200_OK;
            }));
            if (replayed && refresh) s->have_jac = s->factored = true;
            const unsigned long long want = ++s->ticket_expected;
            volatile unsigned long long *tk = reinterpret_cast<volatile unsigned long long *>(s->h_scal + 16);
            for (unsigned spin = 0; *tk < want; ++spin) {
                if ((spin & 0x3fff) == 0x3fff) {
                    const cudaError_t q = cudaStreamQuery(s->stream);
                    if (q != cudaSuccess && q != cudaErrorNotReady)
                        return fail(RST_B200_ERR_CUDA, std::string("frozen iteration: ") + cudaGetErrorString(q));
                    if (q == cudaSuccess && *tk < want) return fail(RST_B200_ERR_CUDA, "frozen iteration finished without publishing its scalars");
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            unpack_scalars(s->h_scal, &sc);
        } else {
            if (jac_recalc) {
                TRY(rst_b200_eval_jacobian(s, 0));
                TRY(rst_b200_factor(s));
            } else {
                TRY(rst_b200_eval_residual(s, 0));
            }
            TRY(rst_b200_solve(s));
            TRY(rst_b200_reduce(s, 0, &sc));
        }
        TRY(check_scalars(sc));
        if (!fast) TRY(rst_b200_make_trial(s, 1.0));  // y_new = y - delta
        TRY(rst_b200_accept_trial(s));
        const double error = std::sqrt(sc.sumsq_step);  // ||y_new - y||
        jac_recalc = frozen_recalc(o, have_jac, m, error, error_old);
        error_old = error;
        st->last_s1 = error;
        if (error < o->tolerance) { status = 1; break; }
        it += 1;
    }
    if (!fast) CUDA_TRY(cudaStreamSynchronize(s->stream));
    st->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    st->status = status;
    return RST_B200_OK;
}

// ---- host-pointer drop-ins -----------------------------------------------------------------------
extern "C" int rst_b200_solve_in_place(rst_b200_solver *s, double *rhs, size_t len) {
    if (!s || !rhs) return fail(RST_B200_ERR_INVALID, "null argument");
    if (s->world != 1) return fail(RST_B200_ERR_UNSUPPORTED, "host drop-in is single-GPU");
    if (!s->factored) return fail(RST_B200_ERR_NOT_FACTORIZED, "solve before factor");
    if ((int64_t)len != s->N * (int64_t)s->nv) return fail(RST_B200_ERR_INVALID, "rhs length mismatch");
    if (s->ls_refine_steps > 0)   // LinearSolver::solve_in_place with iterative_refinement_steps > 0 (linear_solver.rs:84-101): tol 1e-12
        return rst_b200_solve_in_place_with_refinement(s, rhs, len, s->ls_refine_steps, 1e-12, nullptr, nullptr);
    if (s->ls_refine_steps < 0 && s->N >= 4096)   // library default at size: one guarded step
        return rst_b200_solve_in_place_with_refinement(s, rhs, len, 1, 0.0, nullptr, nullptr);
    // rows are interval-major already (row j*nv+i): rhs -> F, solve, gather dY in reduced column order
    CUDA_TRY(cudaMemcpyAsync(s->F, rhs, len * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    TRY(rst_b200_solve(s));
    gather_reduced_kernel<<<s->red_blocks, 256, 0, s->stream>>>(s->layout, s->dY, s->d_tmp);
    s->launches++;
    CUDA_TRY(cudaMemcpyAsync(rhs, s->d_tmp, len * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->h_scal, s->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    unsigned fl;
    std::memcpy(&fl, s->h_scal, sizeof(unsigned));
    if (fl & RST_FLAG_ZERO_PIVOT) return fail(RST_B200_ERR_ZERO_PIVOT, "zero pivot in block elimination");
    return RST_B200_OK;
}

// LinearSolverConfig (solver_policy.rs:24-89).  Every native policy resolves to the pivoted condensation solver -- it is
// the one GPU backend and it accepts banded BVP systems and block-tridiagonal chains alike; the two CPU-only diagnostic
// back-ends have no counterpart here.
extern "C" int rst_b200_set_linear_solver_config(rst_b200_solver *s, int policy, int fallback, int iterative_refinement_steps) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (policy < RST_B200_LS_AUTO || policy > RST_B200_LS_FORCE_FAER_SPARSE) return fail(RST_B200_ERR_INVALID, "unknown LinearSolverPolicy");
    if (fallback != RST_B200_FALLBACK_NEVER && fallback != RST_B200_FALLBACK_TO_FAER_SPARSE) return fail(RST_B200_ERR_INVALID, "unknown FallbackPolicy");
    if (iterative_refinement_steps < -1) return fail(RST_B200_ERR_INVALID, "refinement budget must be >= 0 (or -1: library default)");
    if (policy == RST_B200_LS_FORCE_GENERAL_BANDED_PARTIAL_PIVOT || policy == RST_B200_LS_FORCE_FAER_SPARSE)
        return fail(RST_B200_ERR_UNSUPPORTED, "dense general-pivot and faer sparse LU are CPU diagnostic back-ends of the reference; "
                                              "the B200 library has the pivoted condensation solver only (no CPU fallback)");
    s->ls_policy = policy;
    s->ls_fallback = fallback;
    s->ls_refine_steps = iterative_refinement_steps;
    return RST_B200_OK;
}
extern "C" int rst_b200_get_linear_solver_config(const rst_b200_solver *s, int *policy, int *fallback, int *iterative_refinement_steps) {
    if (!s) return fail(RST_B200_ERR_INVALID, "null handle");
    if (policy) *policy = s->ls_policy;
    if (fallback) *fallback = s->ls_fallback;
    if (iterative_refinement_steps) *iterative_refinement_steps = s->ls_refine_steps;
    return RST_B200_OK;
}

// solve_in_place_with_refinement (lapack_style_banded.rs:1227-1298): solve, then up to `max_iters` guarded refinement
// steps  r = b - J x,  x' = x + J^-1 r,  accepted only while the relative residual max|r| / max(max|b|, 1) decreases;
// stops once it is below `tol`.  The residual uses the Jacobian blocks resident in HBM (the matrix that was factored).
extern "C" int rst_b200_solve_in_place_with_refinement(rst_b200_solver *s, double *rhs, size_t len, int max_iters, double tol,
                                                       int *accepted_steps, double *final_rel_residual) {
    if (!s || !rhs) return fail(RST_B200_ERR_INVALID, "null argument");
    if (s->world != 1) return fail(RST_B200_ERR_UNSUPPORTED, "host drop-in is single-GPU");
    if (!s->factored) return fail(RST_B200_ERR_NOT_FACTORIZED, "solve before factor");
    if ((int64_t)len != s->N * (int64_t)s->nv) return fail(RST_B200_ERR_INVALID, "rhs length mismatch");
    const int nv = s->nv;
    const int64_t rows = s->n * (int64_t)nv, nodes = (s->n + 1) * (int64_t)nv;
    if (!s->ref_b) {
        TRY(dev_alloc(s, &s->ref_b, (size_t)rows));
        TRY(dev_alloc(s, &s->ref_x, (size_t)nodes));
        TRY(dev_alloc(s, &s->ref_xt, (size_t)nodes));
        TRY(dev_alloc(s, &s->ref_norm, (size_t)2));
    }
    cudaStream_t st = s->stream;
    auto residual_into_F = [&](const double *x) {   // F <- b - J x
        if (s->sv.residual) s->sv.residual(s->A, s->B, x, s->ref_b, s->F, s->n, st);
        else block_residual_generic_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(s->A, s->B, x, s->ref_b, s->F, s->n, nv);
        s->launches++;
    };
    auto absmax = [&](const double *v, int64_t total, double *out) -> int {
        CUDA_TRY(cudaMemsetAsync(s->ref_norm, 0, sizeof(unsigned long long), st));
        absmax_kernel<<<s->red_blocks, 256, 0, st>>>(v, total, s->ref_norm);
        s->launches++;
        unsigned long long bits = 0;
        CUDA_TRY(cudaMemcpyAsync(&bits, s->ref_norm, sizeof(bits), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        double d;
        std::memcpy(&d, &bits, sizeof(d));
        *out = (bits == ~0ull) ? std::numeric_limits<double>::quiet_NaN() : d;
        return RST_B200_OK;
    };
    CUDA_TRY(cudaMemcpyAsync(s->ref_b, rhs, len * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->F, s->ref_b, len * sizeof(double), cudaMemcpyDeviceToDevice, st));
    set_solution_buffer(s, s->dY);
    TRY(rst_b200_solve(s));
    CUDA_TRY(cudaMemcpyAsync(s->ref_x, s->dY, nodes * sizeof(double), cudaMemcpyDeviceToDevice, st));
    double b_norm = 0.0;
    TRY(absmax(s->ref_b, rows, &b_norm));
    b_norm = b_norm > 1.0 ? b_norm : 1.0;                       // :1245
    double current_rr = std::numeric_limits<double>::infinity(), last_rr = std::numeric_limits<double>::quiet_NaN();
    int accepted = 0;
    for (int it = 0; it < max_iters; ++it) {
        residual_into_F(s->ref_x);
        double r_norm = 0.0;
        TRY(absmax(s->F, rows, &r_norm));
        const double rr = r_norm / b_norm;
        last_rr = rr;
        if (!std::isfinite(rr) || rr < tol) break;              // :1264-1266
        if (!std::isfinite(current_rr)) current_rr = rr;
        TRY(rst_b200_solve(s));                                  // delta = J^-1 r   (rhs = F, result in dY)
        CUDA_TRY(cudaMemcpyAsync(s->ref_xt, s->ref_x, nodes * sizeof(double), cudaMemcpyDeviceToDevice, st));
        axpy_kernel<<<s->red_blocks, 256, 0, st>>>(s->ref_xt, s->dY, 1.0, nodes);
        s->launches++;
        residual_into_F(s->ref_xt);
        double t_norm = 0.0;
        TRY(absmax(s->F, rows, &t_norm));
        const double trial_rr = t_norm / b_norm;
        if (!std::isfinite(trial_rr) || trial_rr >= std::min(rr, current_rr)) break;   // :1287-1289
        std::swap(s->ref_x, s->ref_xt);
        current_rr = trial_rr;
        last_rr = trial_rr;
        accepted++;
    }
    gather_reduced_kernel<<<s->red_blocks, 256, 0, st>>>(s->layout, s->ref_x, s->d_tmp);
    s->launches++;
    CUDA_TRY(cudaMemcpyAsync(rhs, s->d_tmp, len * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(s->h_scal, s->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    unsigned fl;
    std::memcpy(&fl, s->h_scal, sizeof(unsigned));
    if (fl & RST_FLAG_ZERO_PIVOT) return fail(RST_B200_ERR_ZERO_PIVOT, "zero pivot in block elimination");
    if (accepted_steps) *accepted_steps = accepted;
    if (final_rel_residual) *final_rel_residual = last_rr;
    return RST_B200_OK;
}

extern "C" int rst_b200_solve_multiple_in_place(rst_b200_solver *s, double *rhs, size_t nrhs, size_t ldb) {
    if (!s || !rhs) return fail(RST_B200_ERR_INVALID, "null argument");
    const size_t n = (size_t)(s->N * (int64_t)s->nv);
    if (ldb < n) return fail(RST_B200_ERR_INVALID, "ldb < n");  // BandedError::InvalidRhsLayout
    for (size_t c = 0; c < nrhs; ++c) TRY(rst_b200_solve_in_place(s, rhs + c * ldb, n));
    return RST_B200_OK;
}

// structure of the value stream, sorted by (row, col) (symbolic_functions_BVP.rs:1338-1359)
static void build_structure(rst_b200_solver *s) {
    if (!s->st_rows.empty()) return;
    const int nv = s->nv;
    const unsigned char *pat = s->pb->pattern;
    const bool trap = s->scheme == RST_B200_SCHEME_TRAPEZOID;
    int64_t kl = 0, ku = 0;
    std::vector<int64_t> src;
    for (int64_t j = 0; j < s->N; ++j)
        for (int i = 0; i < nv; ++i) {
            const int64_t row = j * nv + i;
            for (int side = 0; side < 2; ++side)
                for (int c = 0; c < nv; ++c) {
                    const bool present = (side == 0 || trap) ? (c == i || pat[i * nv + c]) : (c == i);
                    if (!present) continue;
                    const int64_t col = reduced_index(s, j + side, c);
                    if (col < 0) continue;
                    s->st_rows.push_back(row);
                    s->st_cols.push_back(col);
                    const int64_t off = (j * nv + i) * (int64_t)nv + c;
                    src.push_back(side == 0 ? off : (trap ? -(off + 2) : -1));
                    if (col > row) ku = std::max(ku, col - row); else kl = std::max(kl, row - col);
                }
        }
    s->st_kl = (int)kl;
    s->st_ku = (int)ku;
    if (cudaMalloc((void **)&s->d_stream_src, src.size() * sizeof(int64_t)) == cudaSuccess) {
        s->owned.push_back(s->d_stream_src);
        cudaMemcpy(s->d_stream_src, src.data(), src.size() * sizeof(int64_t), cudaMemcpyHostToDevice);
    }
    if (cudaMalloc((void **)&s->d_stream_out, src.size() * sizeof(double)) == cudaSuccess) s->owned.push_back(s->d_stream_out);
}

extern "C" int64_t rst_b200_jacobian_structure(rst_b200_solver *s, int64_t *rows, int64_t *cols, int64_t *doff,
                                               int64_t *dpos, int *kl, int *ku) {
    if (!s || s->world != 1) return -1;
    build_structure(s);
    const int64_t nnz = (int64_t)s->st_rows.size();
    for (int64_t k = 0; k < nnz; ++k) {
        if (rows) rows[k] = s->st_rows[k];
        if (cols) cols[k] = s->st_cols[k];
        const int64_t off = s->st_cols[k] - s->st_rows[k];
        if (doff) doff[k] = off;
        if (dpos) dpos[k] = off >= 0 ? s->st_rows[k] : s->st_cols[k];
    }
    if (kl) *kl = s->st_kl;
    if (ku) *ku = s->st_ku;
    return nnz;
}

extern "C" int rst_b200_aot_bind(rst_b200_solver *s) {
    t_bound = s;
    g_bound.store(s, std::memory_order_release);
    return RST_B200_OK;
}
// bind for the calling thread only (the process-wide default is left alone)
extern "C" int rst_b200_aot_bind_thread(rst_b200_solver *s) {
    t_bound = s;
    return RST_B200_OK;
}

// manifest.h of the reference's AOT artefacts (codegen_c_aot_library.rs:250-278) for this handle, byte for byte in the
// reference's format; returns the length (excluding the terminating NUL) or -1 when `cap` is too small.
extern "C" int64_t rst_b200_manifest_h(rst_b200_solver *s, char *buf, size_t cap) {
    if (!s || s->world != 1) return -1;
    build_structure(s);
    const long long n = (long long)(s->N * (int64_t)s->nv);
    char tmp[1024];
    const int len = std::snprintf(tmp, sizeof(tmp),
        "/* AUTO-GENERATED C AOT MANIFEST */\n\n#ifndef MANIFEST_H\n#define MANIFEST_H\n\n#define BACKEND_KIND \"aot\"\n"
        "#define MATRIX_BACKEND \"banded\"\n#define RESIDUAL_LEN %lld\n#define JACOBIAN_ROWS %lld\n#define JACOBIAN_COLS %lld\n"
        "#define JACOBIAN_NNZ %lld\n#define RESIDUAL_FN_NAME \"eval_residual\"\n#define JACOBIAN_FN_NAME \"eval_banded_values\"\n\n"
        "#endif /* MANIFEST_H */\n", n, n, n, (long long)s->st_rows.size());
    if (len < 0 || (size_t)len + 1 > cap || !buf) return -1;
    std::memcpy(buf, tmp, (size_t)len + 1);
    return len;
}

// codegen_c_aot_library.rs:376-389: returns 1 ok / 0 on null pointers or out_len mismatch
extern "C" int rustedscithe_aot_eval_residual(const double *args, size_t args_len, double *out, size_t out_len) {
    rst_b200_solver *s = aot_current();
    if (!s || !args || !out || s->world != 1) return 0;
    std::lock_guard<std::mutex> lock(s->aot_mtx);
    if (cudaSetDevice(s->device) != cudaSuccess) return 0;   // the caller may be a fresh worker thread
    const size_t n = (size_t)(s->N * (int64_t)s->nv);
    if (out_len != n || args_len != n + (size_t)s->np) return 0;
    if (s->np > 0 && cudaMemcpyAsync(s->d_params, args, s->np * sizeof(double), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return 0;
    if (rst_b200_set_state(s, args + s->np, n) != RST_B200_OK) return 0;
    if (rst_b200_eval_residual(s, 0) != RST_B200_OK) return 0;
    if (cudaMemcpyAsync(out, s->F, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return 0;
    return 1;
}

// codegen_c_aot_library.rs:391-409: nnz values in the manifest's entry order
extern "C" int rustedscithe_aot_eval_jacobian_values(const double *args, size_t args_len, double *out, size_t out_len) {
    rst_b200_solver *s = aot_current();
    if (!s || !args || !out || s->world != 1) return 0;
    std::lock_guard<std::mutex> lock(s->aot_mtx);
    if (cudaSetDevice(s->device) != cudaSuccess) return 0;
    const size_t n = (size_t)(s->N * (int64_t)s->nv);
    build_structure(s);
    const size_t nnz = s->st_rows.size();
    if (out_len != nnz || args_len != n + (size_t)s->np || !s->d_stream_src || !s->d_stream_out) return 0;
    if (s->np > 0 && cudaMemcpyAsync(s->d_params, args, s->np * sizeof(double), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return 0;
    if (rst_b200_set_state(s, args + s->np, n) != RST_B200_OK) return 0;
    if (rst_b200_eval_jacobian(s, 0) != RST_B200_OK) return 0;
    gather_values_kernel<<<s->red_blocks, 256, 0, s->stream>>>(s->d_stream_src, s->A, s->B, s->d_stream_out, (int64_t)nnz);
    s->launches++;
    if (cudaMemcpyAsync(out, s->d_stream_out, nnz * sizeof(double), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return 0;
    return 1;
}

// ================================================================================================
// BlockTridiagonal drop-in: BlockTridiagonalLuConsistent::{new, factor_from,
// solve_in_place, solve_multiple_in_place} (block_tridiagonal_lu_consistent.rs:28, :98-218) on the GPU.
// ================================================================================================
struct rst_b200_btd {
    int bs = 0, nv = 0, device = 0;
    int64_t nb = 0, nb_eff = 0, n_int = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    SolverVT sv{};
    std::vector<void *> owned;
    std::vector<LevelBuf> levels;
    double *lower = nullptr, *diag = nullptr, *upper = nullptr;  // device copies, padded to nb_eff blocks
    double *A = nullptr, *B = nullptr, *r = nullptr, *Z = nullptr;
    double *Ba = nullptr, *Bb = nullptr, *ba = nullptr, *bb = nullptr;
    double *top_rows = nullptr, *top_g = nullptr, *Ztop = nullptr;
    double *b_dev = nullptr, *x_dev = nullptr;
    unsigned *d_flags = nullptr;
    bool factored = false;
};

extern "C" int rst_b200_btd_create(int64_t n_blocks, int block_size, int device, void *stream, rst_b200_btd **out) {
    if (!out) return fail(RST_B200_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_blocks <= 0 || block_size <= 0) return fail(RST_B200_ERR_INVALID, "n_blocks and block_size must be positive");  // :29-31
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(RST_B200_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
    rst_b200_btd *s = new rst_b200_btd();
    s->bs = block_size;
    s->nv = 2 * block_size;
    if (!solver_vt_for(s->nv, &s->sv, false)) {
        delete s;
        return fail(RST_B200_ERR_UNSUPPORTED, "block size not compiled in (supported: 1, 2, 3, 4, 6, 8)");
    }
    s->device = device;
    s->nb = n_blocks;
    s->nb_eff = n_blocks < 3 ? 3 : n_blocks;  // padded with decoupled identity blocks
    s->n_int = s->nb_eff - 2;
    CUDA_TRY(cudaSetDevice(device));
    if (stream) s->stream = (cudaStream_t)stream;
    else {
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return fail(RST_B200_ERR_CUDA, "stream"); }
        s->own_stream = true;
    }
    const int bs = s->bs, nv = s->nv;
    auto build = [&]() -> int {
        const size_t bl = (size_t)bs * bs;
        TRY(dev_alloc(s, &s->lower, (size_t)(s->nb_eff - 1) * bl));
        TRY(dev_alloc(s, &s->upper, (size_t)(s->nb_eff - 1) * bl));
        TRY(dev_alloc(s, &s->diag, (size_t)s->nb_eff * bl));
        TRY(dev_alloc(s, &s->A, (size_t)s->n_int * nv * nv));
        TRY(dev_alloc(s, &s->B, (size_t)s->n_int * nv * nv));
        TRY(dev_alloc(s, &s->r, (size_t)s->n_int * nv));
        TRY(dev_alloc(s, &s->Z, (size_t)(s->n_int + 1) * nv));
        TRY(dev_alloc(s, &s->Ba, (size_t)bs * nv));
        TRY(dev_alloc(s, &s->Bb, (size_t)bs * nv));
        TRY(dev_alloc(s, &s->ba, (size_t)bs));
        TRY(dev_alloc(s, &s->bb, (size_t)bs));
        TRY(dev_alloc(s, &s->top_rows, (size_t)2 * nv * nv));
        TRY(dev_alloc(s, &s->top_g, (size_t)nv));
        TRY(dev_alloc(s, &s->Ztop, (size_t)2 * nv));
        TRY(dev_alloc(s, &s->b_dev, (size_t)s->nb_eff * bs));
        TRY(dev_alloc(s, &s->x_dev, (size_t)s->nb_eff * bs));
        TRY(dev_alloc(s, &s->d_flags, (size_t)4));
        int64_t ln = s->n_int;
        const double *lA = s->A, *lB = s->B, *lr = s->r;
        double *lZ = s->Z;
        while (ln > 1) {
            LevelBuf lv{};
            TRY(build_level(s, &lv, ln, default_slab(ln, 0), lA, lB, lr, lZ));
            const int64_t nn = lv.k.nslabs;
            if (nn > 1) {
                TRY(dev_alloc(s, &lv.k.A_next, (size_t)nn * nv * nv));
                TRY(dev_alloc(s, &lv.k.B_next, (size_t)nn * nv * nv));
                TRY(dev_alloc(s, &lv.k.r_next, (size_t)nn * nv));
                double *zn = nullptr;
                TRY(dev_alloc(s, &zn, (size_t)(nn + 1) * nv));
                lv.k.Z_next = zn;
            } else {
                lv.k.A_next = s->top_rows;
                lv.k.B_next = s->top_rows + nv * nv;
                lv.k.r_next = s->top_g;
                lv.k.Z_next = s->Ztop;
            }
            s->levels.push_back(lv);
            lA = lv.k.A_next; lB = lv.k.B_next; lr = lv.k.r_next;
            lZ = const_cast<double *>(lv.k.Z_next);
            ln = nn;
        }
        return RST_B200_OK;
    };
    const int rc = build();
    if (rc != RST_B200_OK) {
        for (void *p : s->owned) cudaFree(p);
        if (s->own_stream) cudaStreamDestroy(s->stream);
        delete s;
        return rc;
    }
    *out = s;
    return RST_B200_OK;
}

extern "C" void rst_b200_btd_destroy(rst_b200_btd *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    for (void *p : s->owned) cudaFree(p);
    if (s->own_stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int64_t rst_b200_btd_n(const rst_b200_btd *s) { return s ? s->nb * s->bs : 0; }

// factor_from(&BlockTridiagonal): lower[k] = block (k+1, k), upper[k] = block (k, k+1), row-major bs x bs blocks
extern "C" int rst_b200_btd_factor_from(rst_b200_btd *s, const double *lower, const double *diag, const double *upper) {
    if (!s || !diag || (s->nb > 1 && (!lower || !upper))) return fail(RST_B200_ERR_INVALID, "null argument");
    const size_t bl = (size_t)s->bs * s->bs;
    const int bs = s->bs;
    s->factored = false;
    CUDA_TRY(cudaMemsetAsync(s->lower, 0, (size_t)(s->nb_eff - 1) * bl * sizeof(double), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->upper, 0, (size_t)(s->nb_eff - 1) * bl * sizeof(double), s->stream));
    if (s->nb > 1) {
        CUDA_TRY(cudaMemcpyAsync(s->lower, lower, (size_t)(s->nb - 1) * bl * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->upper, upper, (size_t)(s->nb - 1) * bl * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(s->diag, diag, (size_t)s->nb * bl * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (s->nb_eff > s->nb) {  // identity padding blocks
        std::vector<double> eye((size_t)(s->nb_eff - s->nb) * bl, 0.0);
        for (int64_t k = 0; k < s->nb_eff - s->nb; ++k)
            for (int i = 0; i < bs; ++i) eye[k * bl + (size_t)i * bs + i] = 1.0;
        CUDA_TRY(cudaMemcpyAsync(s->diag + s->nb * bl, eye.data(), eye.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    CUDA_TRY(cudaMemsetAsync(s->d_flags, 0, sizeof(unsigned), s->stream));
    btd_embed_kernel<<<148 * 4, 128, 0, s->stream>>>(s->lower, s->diag, s->upper, s->A, s->B, s->n_int, bs);
    btd_boundary_kernel<<<1, 128, 0, s->stream>>>(s->lower, s->diag, s->upper, s->Ba, s->Bb, s->nb_eff, bs);
    for (auto &lv : s->levels) s->sv.factor(lv.k, false, s->d_flags, s->stream);
    CUDA_TRY(cudaGetLastError());
    unsigned fl = 0;
    CUDA_TRY(cudaMemcpyAsync(&fl, s->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (fl & RST_FLAG_ZERO_PIVOT) return fail(RST_B200_ERR_ZERO_PIVOT, "zero pivot: the block-tridiagonal matrix is singular");
    s->factored = true;
    return RST_B200_OK;
}

extern "C" int rst_b200_btd_solve_in_place(rst_b200_btd *s, double *rhs, size_t len) {
    if (!s || !rhs) return fail(RST_B200_ERR_INVALID, "null argument");
    if (!s->factored) return fail(RST_B200_ERR_NOT_FACTORIZED, "solve before factor");                 // BandedError::NotFactorized
    if ((int64_t)len != s->nb * s->bs) return fail(RST_B200_ERR_INVALID, "rhs length mismatch");       // DimensionMismatch
    const int nv = s->nv, bs = s->bs;
    CUDA_TRY(cudaMemsetAsync(s->b_dev, 0, (size_t)s->nb_eff * bs * sizeof(double), s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->b_dev, rhs, len * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    btd_rhs_kernel<<<148 * 2, 128, 0, s->stream>>>(s->b_dev, s->r, s->ba, s->bb, s->nb_eff, bs);
    for (auto &lv : s->levels) s->sv.forward(lv.k, s->stream);
    const bool single = s->levels.empty();
    const double *C = single ? s->A : s->top_rows, *D = single ? s->B : s->top_rows + nv * nv, *g = single ? s->r : s->top_g;
    s->sv.top_general(C, D, g, s->Ba, s->ba, bs, s->Bb, s->bb, bs, single ? s->Z : s->Ztop, s->d_flags, s->stream);
    for (int i = (int)s->levels.size() - 1; i >= 0; --i) s->sv.backward(s->levels[i].k, s->stream);
    btd_extract_kernel<<<148 * 2, 128, 0, s->stream>>>(s->Z, s->x_dev, s->nb_eff, bs);
    CUDA_TRY(cudaGetLastError());
    unsigned fl = 0;
    CUDA_TRY(cudaMemcpyAsync(rhs, s->x_dev, len * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaMemcpyAsync(&fl, s->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (fl & RST_FLAG_ZERO_PIVOT) return fail(RST_B200_ERR_ZERO_PIVOT, "zero pivot: the block-tridiagonal matrix is singular");
    return RST_B200_OK;
}

// column-major multi-RHS with leading dimension ldb >= n (block_tridiagonal_lu_consistent.rs:220-247)
extern "C" int rst_b200_btd_solve_multiple_in_place(rst_b200_btd *s, double *rhs, size_t nrhs, size_t ldb) {
    if (!s || !rhs) return fail(RST_B200_ERR_INVALID, "null argument");
    const size_t n = (size_t)(s->nb * s->bs);
    if (ldb < n) return fail(RST_B200_ERR_INVALID, "ldb < n");  // BandedError::InvalidRhsLayout
    for (size_t c = 0; c < nrhs; ++c) TRY(rst_b200_btd_solve_in_place(s, rhs + c * ldb, n));
    return RST_B200_OK;
}
