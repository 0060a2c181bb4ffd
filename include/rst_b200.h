/*
 * rst_b200.h -- C ABI of the B200-native damped-Newton BVP hot path (librst_b200.so).
 *
 * Drop-in boundary for RustedSciThe's residual/Jacobian backend and banded-solver interfaces.
 * Plain pointers and sizes only; every entry point returns an int status (never throws / panics
 * across the ABI).  Paths below are relative to the reference repository root.
 *
 * Two groups of symbols:
 *  (1) the reference's existing AOT artefact ABI, unchanged:
 *        src/symbolic/codegen/c_backend/codegen_c_aot_library.rs:376-409 (exported C),
 *        src/symbolic/codegen/codegen_aot_runtime_link.rs:38, :385-417 (dlopen side,
 *        `unsafe extern "C" fn(*const f64, usize, *mut f64, usize) -> bool`)
 *  (2) a device-resident handle API that the Rust-side `Fun` / `Jac` / `MatrixType` /
 *      `DirectLinearSolver` shims call (src/numerical/BVP_Damp/BVP_traits.rs:459-461, :943-953,
 *      :613-630; src/somelinalg/banded/solver_traits.rs:3-14), plus the whole damped loop
 *      (src/numerical/BVP_Damp/NR_Damp_solver_damped.rs:1782-2156).
 *
 * A handle is thread-compatible (not thread-safe) and owns one CUDA stream's worth of work.
 * There is no CPU fallback: without a CUDA device every call fails with RST_B200_ERR_CUDA.
 */
#ifndef RST_B200_H
#define RST_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RST_B200_OK 0
#define RST_B200_ERR_INVALID (-1)        /* bad argument / dimension mismatch (BandedError::DimensionMismatch, error.rs:4-39) */
#define RST_B200_ERR_CUDA (-2)           /* CUDA runtime failure or no device */
#define RST_B200_ERR_ZERO_PIVOT (-3)     /* BandedError::ZeroPivot */
#define RST_B200_ERR_NAN (-4)            /* NaN in residual or step (panic in the reference, NR_Damp_solver_damped.rs:1821-1844) */
#define RST_B200_ERR_UNSUPPORTED (-5)    /* problem key / block size not compiled in */
#define RST_B200_ERR_NOT_FACTORIZED (-6) /* BandedError::NotFactorized */

#define RST_B200_SCHEME_FORWARD 0
#define RST_B200_SCHEME_TRAPEZOID 1

typedef struct rst_b200_solver rst_b200_solver;

/* Problem descriptor: the arguments of NRBVP::new_with_options (NR_Damp_solver_damped.rs:847) that
 * the hot path needs, plus the mesh-slab partition for multi-GPU runs. */
typedef struct {
    const char *problem_key;   /* key of a compiled-in AOT problem (rst_b200_problem_key) */
    int64_t n_steps;           /* GLOBAL number of intervals N (mesh has N+1 nodes, solver_common.rs:91-94) */
    int scheme;                /* RST_B200_SCHEME_* (eq_step, symbolic_functions_BVP.rs:5221-5257) */
    int trap_time;             /* trapezoid second-term time: 0 = t_j (symbolic route), 1 = t_{j+1} (closure route) */
    double t0, t_end;          /* uniform mesh: h = (t_end - t0)/N, t_j = t0 + j*h (NR_Damp_solver_damped.rs:550-553) */
    const double *mesh;        /* optional explicit GLOBAL mesh [N+1] (host); NULL = uniform */
    int n_bc;                  /* must equal nv (numeric_discretization.rs:60-67) */
    const int *bc_var;         /* variable index of each boundary condition */
    const int *bc_side;        /* 0 = left (node 0), 1 = right (node N) */
    const double *bc_val;
    const double *params;      /* [n_params] (codegen_runtime_api.rs:74-83: args = [params..., unknowns...]) */
    int n_params;
    const double *lo, *hi;     /* per-variable bounds [nv] (looked up by base name, symbolic_functions_BVP.rs:5337-5378) */
    const double *rtol;        /* per-variable relative tolerance [nv] */
    int device;                /* CUDA device ordinal */
    void *stream;              /* cudaStream_t to run on; NULL = the library creates its own */
    int slab;                  /* rows per slab of the condensation (0 = library default) */
    int rank, world;           /* mesh-slab partition: this process owns intervals [j_begin, j_end) of rank `rank` of `world` */
} rst_b200_desc;

typedef struct {
    int max_iterations;        /* default 25 (NR_Damp_solver_damped.rs:444-459) */
    int max_jac;               /* SolverParams.max_jac, default 3 (:142-151) */
    int max_damp_iter;         /* default 5 */
    double damp_factor;        /* default 0.5 */
    double abs_tol;            /* default 1e-6 */
} rst_b200_newton_opts;

typedef struct {
    int status;                /* 1 converged, 0 not converged, -2 no damping coefficient, -3 bound violation, RST_B200_ERR_* */
    int iterations;            /* "number of iterations"                 (:2058-2061) */
    int linear_solves;         /* "number of solving linear systems"     (:1881-1884) */
    int jac_recalcs;           /* "number of jacobians recalculations"   (:1736-1738) */
    int factorizations;        /* 1 per Jacobian here; the reference refactors per solve (BVP_traits.rs:875-899) */
    double last_s0, last_s1, last_conv, last_fbound, last_lambda;
    double ms_assemble, ms_factor, ms_solve, ms_total; /* CUDA-event timings, cumulative */
    int solves_reused;         /* of linear_solves (the reference's count): undamped steps taken over, bit for bit, from the accepted
                                * trial of the previous iteration (same y, same factors) instead of being recomputed */
} rst_b200_newton_stats;

/* scalars of one damping trial, all that crosses PCIe inside the loop */
typedef struct {
    double sumsq_step;         /* ||step||_2^2 over unknowns */
    double tol_sum;            /* sum |y_i| rtol_i over unknowns */
    double fbound;             /* bound_step_Cantera2(y, step) */
    double nan_flag;           /* != 0: NaN in F or step */
    double device_flags;       /* RST_FLAG_* bits from the kernels (zero pivot) */
} rst_b200_scalars;

/* ---- library / problem registry ------------------------------------------------------------- */
int rst_b200_version(void);
int rst_b200_device_count(void);
int rst_b200_problem_count(void);
const char *rst_b200_problem_key(int index);
/* nv, n_params, emitter hash (analogue of problem_key, codegen_aot_registry.rs) of a compiled-in problem */
int rst_b200_problem_info(const char *key, int *nv, int *n_params, const char **hash);
const char *rst_b200_last_error(void);

/* ---- lifecycle ------------------------------------------------------------------------------- */
int rst_b200_create(const rst_b200_desc *desc, rst_b200_solver **out);
void rst_b200_destroy(rst_b200_solver *s);
int64_t rst_b200_n_unknowns(const rst_b200_solver *s);      /* local N_local * nv rows (global N*nv when world == 1) */
int rst_b200_set_params(rst_b200_solver *s, const double *params, int n_params); /* re-bind parameters */
int rst_b200_sync(rst_b200_solver *s);

/* ---- state: reduced ordering (reference's unknown vector) <-> device-resident full state ---- */
/* host y_reduced[N*nv] -> device (boundary slots pinned, numeric_discretization.rs:195-201) */
int rst_b200_set_state(rst_b200_solver *s, const double *y_reduced, size_t len);
int rst_b200_get_state(rst_b200_solver *s, double *y_reduced, size_t len);
/* (N+1) x nv with boundary values re-inserted (construct_full_solution, BVP_utils.rs:534-558) */
int rst_b200_get_full_solution(rst_b200_solver *s, double *full, size_t len);
/* raw device pointers (double*): which = 0 Y, 1 F, 2 A, 3 B, 4 dY, 5 Y_trial, 6 interface rows, 7 interface rhs */
void *rst_b200_device_ptr(rst_b200_solver *s, int which);

/* ---- hot path, device resident ---------------------------------------------------------------- */
/* F(Y) (Fun::call) ; which_state: 0 = current state, 1 = trial state */
int rst_b200_eval_residual(rst_b200_solver *s, int which_state);
/* F(Y) and the Jacobian blocks (Jac::call); invalidates the factorisation */
int rst_b200_eval_jacobian(rst_b200_solver *s, int which_state);
/* factor the current Jacobian blocks (LapackStyleBandedLuFaithful::factor_from replacement) */
int rst_b200_factor(rst_b200_solver *s);
/* dY = J^{-1} F (MatrixType::solve_sys on the cached factors) */
int rst_b200_solve(rst_b200_solver *s);
/* factor the current Jacobian blocks AND solve J dY = F with the residual the handle holds, in one call -- what the
 * reference's solve_sys does every time (factor_from + solve_in_place, BVP_traits.rs:875-899).  Same result, bit for bit,
 * as rst_b200_factor followed by rst_b200_solve; for 12-variable problems the right-hand side rides through the
 * factorisation as a spare panel column (abd4.cuh), so the forward sweeps of that first solve are not run at all.
 * The factors stay cached for later rst_b200_solve calls. */
int rst_b200_factor_solve(rst_b200_solver *s);
/* reductions of the current step against state `which_state` */
int rst_b200_reduce(rst_b200_solver *s, int which_state, rst_b200_scalars *out);
/* Y_trial = Y - lambda dY */
int rst_b200_make_trial(rst_b200_solver *s, double lambda);
/* accept the trial state as the current state */
int rst_b200_accept_trial(rst_b200_solver *s);
/* device-side checkpoint of the state: restore == 0 saves Y, restore != 0 copies it back (no host traffic) */
int rst_b200_checkpoint(rst_b200_solver *s, int restore);

/* ---- adaptive grid refinement hand-off ------------------------------------------- */
/* GridRefinementMethod (src/numerical/BVP_Damp/grid_api.rs:62-115); TwoPoint is not provided */
#define RST_B200_GRID_DOUBLE_POINTS 0   /* refine_all_grid_par      adaptive_grid_basic.rs:187 */
#define RST_B200_GRID_EASY 1            /* easy_grid_refinement_par :257, params = {tolerance} */
#define RST_B200_GRID_PEARSON 2         /* pearson_grid_refinement_par :469, params = {d, C} */
#define RST_B200_GRID_GRCAR_SMOOKE 3    /* grcar_smooke_grid_refinement :565, params = {d, g, C} */
#define RST_B200_GRID_SCI 4             /* scipy_grid_refinement :885, uses abs_tolerance and the residual of the state */
/* new_grid (grid_api.rs:154): marks, refined mesh and linearly interpolated full solution computed on the device from
 * the CURRENT state; results stay in the handle.  n_added = number of inserted nodes (0 = "no new grid is needed"). */
int rst_b200_new_grid(rst_b200_solver *s, int method, const double *params, double abs_tolerance,
                      int64_t *n_points_new, int64_t *n_added);
/* host copies of the last rst_b200_new_grid result (any pointer may be NULL): mesh [n_points_new], full solution
 * [n_points_new * nv] node-major, marks [N + 1] */
int rst_b200_new_grid_fetch(rst_b200_solver *s, double *mesh_out, size_t mesh_len, double *y_full_out, size_t y_len,
                            int32_t *marks_out, size_t marks_len);
/* try_solve_with_new_grid (NR_Damp_solver_damped.rs:2281-2340): a new handle on the refined mesh whose state is the
 * interpolated solution (device to device); the caller continues with rst_b200_newton_solve and destroys the old handle */
int rst_b200_create_refined(rst_b200_solver *s, rst_b200_solver **out);

/* whole damped Newton loop (try_main_loop_damped, NR_Damp_solver_damped.rs:2027-2156) */
int rst_b200_newton_solve(rst_b200_solver *s, const rst_b200_newton_opts *opts, rst_b200_newton_stats *stats);
/* ONE iteration of that loop from the current state (bench step): recalc_jacobian != 0 forces a
 * Jacobian evaluation + factorisation first. Returns the damped_step status (1 / 0 / -2 / -3). */
int rst_b200_newton_iterate(rst_b200_solver *s, const rst_b200_newton_opts *opts, int recalc_jacobian,
                            rst_b200_newton_stats *stats);

/* frozen-Jacobian Newton (NR_Damp_solver_frozen.rs:1035-1148): y_new = y - J^-1 F(y), converged when
 * ||y_new - y||_2 < tolerance, Jacobian recalculated per the reference's strategies (BVP_utils.rs:362-480) */
#define RST_B200_FROZEN_NAIVE 0         /* "Naive": every iteration */
#define RST_B200_FROZEN_FROZEN_NAIVE 1  /* "Frozen_naive": first iteration only */
#define RST_B200_FROZEN_EVERY_M 2       /* "every_m" {m} */
#define RST_B200_FROZEN_AT_HIGH_NORM 3  /* "at_high_morm" {norm} */
#define RST_B200_FROZEN_AT_LOW_SPEED 4  /* "at_low_speed" {rate} */
#define RST_B200_FROZEN_COMPLEX 5       /* "complex" {m, norm, rate} */
typedef struct {
    int strategy;
    double strategy_params[3];
    double tolerance;
    int max_iterations;
} rst_b200_frozen_opts;
int rst_b200_newton_frozen(rst_b200_solver *s, const rst_b200_frozen_opts *opts, rst_b200_newton_stats *stats);

/* ---- host-pointer drop-ins ---------------------------------------------------------------------- */
/* DirectLinearSolver::solve_in_place (solver_traits.rs:3-14): rhs[N*nv] in the reference's row order,
 * overwritten with the solution in the reference's reduced column order. Needs rst_b200_factor. */
int rst_b200_solve_in_place(rst_b200_solver *s, double *rhs, size_t len);
/* LinearSolverConfig (src/somelinalg/banded/solver_policy.rs:10-89): policy / fallback / guarded-refinement budget of the
 * DirectLinearSolver drop-ins.  Numbering = declaration order of the reference's enums. */
enum rst_b200_linear_solver_policy {            /* LinearSolverPolicy, solver_policy.rs:10-17 */
    RST_B200_LS_AUTO = 0,
    RST_B200_LS_FORCE_BLOCK_TRIDIAGONAL = 1,
    RST_B200_LS_FORCE_BLOCK_TRIDIAGONAL_CONSISTENT = 2,
    RST_B200_LS_FORCE_BANDED = 3,
    RST_B200_LS_FORCE_GENERAL_BANDED_PARTIAL_PIVOT = 4,   /* CPU diagnostic back-end: RST_B200_ERR_UNSUPPORTED */
    RST_B200_LS_FORCE_FAER_SPARSE = 5                     /* CPU back-end (faer): RST_B200_ERR_UNSUPPORTED */
};
enum rst_b200_fallback_policy {                 /* FallbackPolicy, solver_policy.rs:19-23 */
    RST_B200_FALLBACK_NEVER = 0,
    RST_B200_FALLBACK_TO_FAER_SPARSE = 1        /* accepted and remembered; a zero pivot is reported, never retried on a CPU */
};
/* Defaults mirror LinearSolverConfig::default(): Auto, ToFaerSparse, 0 refinement steps.  Auto / ForceBanded /
 * ForceBlockTridiagonal* all select the pivoted condensation solver (it handles both matrix classes, including the node
 * blocks the reference's block solver rejects).  iterative_refinement_steps > 0 makes rst_b200_solve_in_place /
 * _solve_multiple_in_place run the guarded refinement below with tol = 1e-12 (linear_solver.rs:84-101); 0 = never (the
 * reference's default).  A fresh handle holds -1 = library default: one guarded step (tol 0) on meshes of >= 4096 intervals,
 * none below -- the condensation has the banded LU's backward error, but its forward error grows with the mesh and passes
 * 1e-10 there.  The Newton drivers never refine. */
int rst_b200_set_linear_solver_config(rst_b200_solver *s, int policy, int fallback, int iterative_refinement_steps);
int rst_b200_get_linear_solver_config(const rst_b200_solver *s, int *policy, int *fallback, int *iterative_refinement_steps);
/* solve_in_place_with_refinement (lapack_style_banded.rs:1227-1298): solve, then at most max_iters guarded steps of
 * iterative refinement x' = x + J^-1 (b - J x); a step is kept only while max|b - J x| / max(max|b|, 1) decreases, the loop
 * ends once that ratio is below tol.  The residual is formed on the device with the blocks that were factored.
 * accepted_steps / final_rel_residual may be NULL. */
int rst_b200_solve_in_place_with_refinement(rst_b200_solver *s, double *rhs, size_t len, int max_iters, double tol,
                                            int *accepted_steps, double *final_rel_residual);
/* solve_multiple_in_place: column-major, leading dimension ldb >= n */
int rst_b200_solve_multiple_in_place(rst_b200_solver *s, double *rhs, size_t nrhs, size_t ldb);
/* sparse/banded structure of the value stream (BandedJacobianStructure, codegen_runtime_api.rs:414-421):
 * returns nnz; arrays may be NULL */
int64_t rst_b200_jacobian_structure(rst_b200_solver *s, int64_t *rows, int64_t *cols, int64_t *diag_offsets,
                                    int64_t *diag_positions, int *kl, int *ku);

/* ---- the reference's AOT artefact ABI (codegen_c_aot_library.rs:376-409), bound to one handle ---- */
/* args = [params..., y_reduced...]; returns 1 ok / 0 on NULL pointer or length mismatch (:382-384) */
int rst_b200_aot_bind(rst_b200_solver *s);
/* Re-entrancy (the reference calls the artefact's symbols from rayon worker threads, symbolic_functions_BVP.rs:2308-2357,
 * and marks the linked library Send + Sync, codegen_aot_runtime_link.rs:21-36): rst_b200_aot_bind binds the calling thread
 * AND the library-wide default that threads without a binding of their own use; rst_b200_aot_bind_thread binds the calling
 * thread only.  Concurrent calls on one handle are serialised by the handle, calls on different handles run concurrently.
 * Every AOT artefact (.so) carries its own binding. */
int rst_b200_aot_bind_thread(rst_b200_solver *s);
/* manifest.h of the reference's C AOT artefacts (codegen_c_aot_library.rs:250-278) for this handle, in the reference's
 * exact text format: BACKEND_KIND "aot", MATRIX_BACKEND "banded", RESIDUAL_LEN, JACOBIAN_ROWS / _COLS, JACOBIAN_NNZ,
 * RESIDUAL_FN_NAME "eval_residual", JACOBIAN_FN_NAME "eval_banded_values" (codegen_runtime_api.rs:1100).  Returns the
 * length without the terminating NUL, -1 if cap is too small. */
int64_t rst_b200_manifest_h(rst_b200_solver *s, char *buf, size_t cap);
int rustedscithe_aot_eval_residual(const double *args, size_t args_len, double *out, size_t out_len);
int rustedscithe_aot_eval_jacobian_values(const double *args, size_t args_len, double *out, size_t out_len);

/* ---- multi-GPU mesh slabs -------------------------------------------------
 * Each rank owns a contiguous slab of intervals and condenses it to ONE interface block row
 * (2 nv^2 + nv doubles); the `world` interface rows form a tiny block-bidiagonal system that every
 * rank solves redundantly.  The only data-path exchange is an all-gather of those rows; the library
 * calls back into the host application for it (torch.distributed / NCCL over NVLink). */
typedef int (*rst_b200_allgather_fn)(void *ctx, const void *d_send, void *d_recv, int64_t doubles_per_rank,
                                     void *stream);
int rst_b200_set_allgather(rst_b200_solver *s, rst_b200_allgather_fn fn, void *ctx);
/* NVLink peer-memory exchange (preferred): every rank exports the CUDA IPC handle of its mailbox (64 bytes), the host
 * application all-gathers the handles once (any transport) and hands the array back; from then on rst_b200_factor /
 * _solve / _reduce run the all-gather as a small kernel on the solver's own stream -- no host callback, no NCCL. */
int rst_b200_p2p_get_handle(rst_b200_solver *s, void *handle_out, size_t handle_bytes);
int rst_b200_p2p_open(rst_b200_solver *s, const void *handles /* [world][handle_stride] */, size_t handle_stride);
/* Wire format: every 8-byte word carries its own 32-bit sequence tag (a double travels as two flagged words), so no
 * store wider than the architecture's single-copy-atomic unit has to arrive untorn.  A rank that waits longer than
 * RST_B200_PEER_TIMEOUT_MS (environment, default 30000) for a peer poisons its handle: the exchange returns NaN words,
 * every later reduction reports RST_B200_ERR_CUDA.  Callers barrier once after rst_b200_p2p_open (all ranks mapped)
 * before the first factorisation.
 * rst_b200_p2p_selftest: `world` emulated ranks (one CTA each, ONE cooperative launch on `device`) run `rounds`
 * back-to-back gathers through the same device function the solver uses; every received word is checked bit for bit.
 * absent_rank >= 0: that rank never arrives -- the others must time out (timeout_ms), flag it and receive NaN. */
int rst_b200_p2p_selftest(int device, int world, int rounds, int pay, int absent_rank, double timeout_ms,
                          int64_t *mismatches, int64_t *nan_words, int *ranks_timed_out);
/* local intervals [begin, end) of this rank */
int rst_b200_dist_range(const rst_b200_solver *s, int64_t *j_begin, int64_t *j_end);
/* phase-split primitives (rst_b200_factor / _solve / _reduce call them around the all-gather callback;
 * exposed so a single process can drive several emulated ranks):
 *   buffers: which = 6 interface rows [2 nv^2] (send), 7 interface rhs [nv] (send), 8 gathered rows
 *   [world][2 nv^2] (recv), 9 gathered rhs [world][nv] (recv), 10 scalars [8] (send), 11 gathered scalars */
int rst_b200_factor_local(rst_b200_solver *s);
int rst_b200_factor_reduced(rst_b200_solver *s);
int rst_b200_solve_local_forward(rst_b200_solver *s);
int rst_b200_solve_reduced_backward(rst_b200_solver *s);
int rst_b200_reduce_local(rst_b200_solver *s, int which_state);
int rst_b200_reduce_finish(rst_b200_solver *s, rst_b200_scalars *out);

/* ---- BlockTridiagonal drop-in (src/somelinalg/banded/block_tridiagonal.rs:3-19,
 *      block_tridiagonal_lu_consistent.rs:28, :98-247): BlockTridiagonalLuConsistent::{new, factor_from,
 *      solve_in_place, solve_multiple_in_place}.  Blocks are row-major bs x bs; lower[k] = block (k+1, k),
 *      upper[k] = block (k, k+1).  Unlike the reference, pivoting runs across neighbouring block rows, so singular
 *      diagonal blocks are accepted when the matrix itself is non-singular.  Block sizes 1, 2, 3, 4, 6, 8. ---- */
typedef struct rst_b200_btd rst_b200_btd;
int rst_b200_btd_create(int64_t n_blocks, int block_size, int device, void *stream, rst_b200_btd **out);
void rst_b200_btd_destroy(rst_b200_btd *s);
int64_t rst_b200_btd_n(const rst_b200_btd *s);
int rst_b200_btd_factor_from(rst_b200_btd *s, const double *lower, const double *diag, const double *upper);
int rst_b200_btd_solve_in_place(rst_b200_btd *s, double *rhs, size_t len);
int rst_b200_btd_solve_multiple_in_place(rst_b200_btd *s, double *rhs, size_t nrhs, size_t ldb);

/* synchronous cudaMemcpy on raw pointers (kind: 1 H2D, 2 D2H, 3 D2D) -- lets a host language without CUDA
 * bindings move data to / from rst_b200_device_ptr buffers */
int rst_b200_memcpy(void *dst, const void *src, size_t bytes, int kind);

/* page-locked host memory for the state vectors a host hands to rst_b200_set_state / rst_b200_get_state: from pinned
 * buffers the two copies run at the PCIe rate and asynchronously to the kernels of OTHER handles, which is what lets a
 * host that runs several independent solves (one handle per worker thread, as a Rust host would with rayon over
 * NRBVP instances, NR_Damp_solver_damped.rs:2130-2156 being one solve) hide the copies of one behind the kernels of
 * another -- rustedscithe_b200/pipeline.py is that pattern.  Pageable buffers work too, only slower. */
int rst_b200_host_alloc(size_t bytes, void **out);
int rst_b200_host_free(void *p);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t rst_b200_launch_count(const rst_b200_solver *s);

#ifdef __cplusplus
}
#endif
#endif
