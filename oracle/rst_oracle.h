/*
 * rst_oracle.h -- CPU restatement ("oracle") of RustedSciThe's damped-Newton BVP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rustedscithe_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * The reference is Rust and cannot be compiled in this image (no cargo/rustc), so this is a
 * port ("kind": "port"), pinned against the reference's own known-answer tests and generated
 * fixtures (tests/test_oracle_*.py) and cross-checked against scipy's LAPACK dgbtrf/dgbtrs.
 *
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 */
#ifndef RST_ORACLE_H
#define RST_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- continuous problems y' = f(t, y, p) (hand written, independent of the CUDA emitter) ---- */
typedef void (*orc_rhs_fn)(double t, const double *y, const double *p, double *f);
typedef void (*orc_jac_fn)(double t, const double *y, const double *p, double *fy /* nv*nv row-major */);
typedef struct {
    const char *name;
    int nv;
    int n_params;
    orc_rhs_fn rhs;
    orc_jac_fn jac;
} orc_problem;
const orc_problem *orc_problem_get(const char *name);
/* structural non-zero pattern of f_y (nv*nv bytes, 1 = symbolically non-zero) */
void orc_problem_pattern(const orc_problem *pb, unsigned char *pattern);

/* ---- unknown layout (numeric_discretization.rs:27-106; symbolic_functions_BVP.rs:5529-5537) ---- */
/* bc_var/bc_side/bc_val: nbc boundary conditions.  Returns 0 ok, -1 wrong count (!= nv),
 * -2 duplicate position, -3 bad side flag.  unknown_pos[n_steps*nv] = full position of unknown k;
 * full_to_unknown[(n_steps+1)*nv] = k or -1; bc_pos_sorted/bc_val_sorted[nv]. */
int orc_layout_build(int nv, int n_steps, int nbc, const int *bc_var, const int *bc_side,
                     const double *bc_val, int64_t *unknown_pos, int64_t *full_to_unknown,
                     int64_t *bc_pos_sorted, double *bc_val_sorted);

typedef struct {
    const orc_problem *pb;
    const double *params;
    int nv;
    int n_steps;
    int scheme;        /* 0 forward, 1 trapezoid */
    int trap_time;     /* trapezoid second term time: 0 = t_j (symbolic route), 1 = t_{j+1} (closure route) */
    const double *mesh; /* n_steps+1 */
    const int64_t *unknown_pos;
    const int64_t *full_to_unknown;
    const int64_t *bc_pos;
    const double *bc_val;
} orc_bvp;

/* full[(n_steps+1)*nv] <- BC values + reduced unknowns (numeric_discretization.rs:195-201) */
void orc_expand_full(const orc_bvp *b, const double *y_reduced, double *full);
/* residual r[n_steps*nv] (numeric_discretization.rs:203-240, symbolic_functions_BVP.rs:5486-5490) */
void orc_residual(const orc_bvp *b, const double *y_reduced, double *r);

/* sparse structure of J sorted by (row, col) (symbolic_functions_BVP.rs:1338-1359):
 * returns nnz; arrays may be NULL to query the count. kl/ku as find_bandwidths (:5639-5656). */
int64_t orc_jacobian_structure(const orc_bvp *b, int64_t *rows, int64_t *cols,
                               int64_t *diag_offset, int64_t *diag_pos, int *kl, int *ku);
/* values in the same order (numeric_discretization.rs:328-342) */
void orc_jacobian_values(const orc_bvp *b, const double *y_reduced, const int64_t *rows,
                         const int64_t *cols, int64_t nnz, double *values);

/* ---- band layouts (banded_assembly.rs:5-168, storage.rs:5-16, lapack_style_banded.rs:246-278) ---- */
/* BandedAssembly flattened: diagonal d=-kl..ku stored at diag_start[d+kl], length n-|d| */
int64_t orc_assembly_size(int64_t n, int kl, int ku, int64_t *diag_start /* kl+ku+1 */);
void orc_assemble_diagonals(int64_t n, int kl, int ku, int64_t nnz, const int64_t *diag_offset,
                            const int64_t *diag_pos, const double *values, double *diags);
/* Banded<f64>: data[(ku+i-j)*n + j] */
void orc_assembly_to_banded(int64_t n, int kl, int ku, const double *diags, double *banded);
/* LAPACK AB: ab[(kv+i-j) + j*ldab], ldab = 2kl+ku+1 */
void orc_banded_to_ab(int64_t n, int kl, int ku, const double *banded, double *ab);

/* ---- LapackStyleBandedLuFaithful (lapack_style_banded.rs:386-475, :1163-1225) ---- */
/* returns 0 ok, j+1 if |pivot| <= eps at column j */
int64_t orc_dgb_factor(int64_t n, int kl, int ku, double *ab, int64_t *ipiv, double eps);
int64_t orc_dgb_solve(int64_t n, int kl, int ku, const double *ab, const int64_t *ipiv,
                      double *rhs, double eps);
/* y = A x for a Banded<f64> (ops.rs:4) */
void orc_banded_matvec(int64_t n, int kl, int ku, const double *banded, const double *x, double *y);

/* ---- BlockTridiagonalLuConsistent (block_tridiagonal_lu_consistent.rs:98-218) ---- */
/* blocks row-major bs*bs; lower[k] = (k+1,k), upper[k] = (k,k+1). In place. returns 0 / block*bs+k+1 */
int64_t orc_btd_factor(int64_t nb, int bs, double *lower, double *diag, double *upper, int64_t *pivots);
int64_t orc_btd_solve(int64_t nb, int bs, const double *lower, const double *diag,
                      const double *upper, const int64_t *pivots, double *rhs);

/* ---- damping helpers (BVP_utils_damped.rs:39-61, :185-222) ---- */
double orc_bound_step_cantera2(int64_t n, const double *x, const double *step, const double *lo,
                               const double *hi);
double orc_convergence_condition(int64_t n, const double *y, double abs_tol, const double *rtol);
double orc_norm2(int64_t n, const double *x);

/* ---- damped Newton driver (NR_Damp_solver_damped.rs:1746-2156) ---- */
typedef struct {
    int max_iterations;   /* :444-459 default 25 */
    int max_jac;          /* SolverParams default 3 */
    int max_damp_iter;    /* 5 */
    double damp_factor;   /* 0.5 */
    double abs_tol;       /* 1e-6 */
    int refactor_per_solve; /* 1 = reference behaviour (BVP_traits.rs:875-899) */
} orc_newton_opts;
typedef struct {
    int status;           /* 1 converged, 0 max iterations / break without convergence, <0 failure code */
    int iterations;
    int linear_solves;
    int jac_recalcs;
    int factorizations;
    double last_s1;
    double last_conv;
    double t_fun, t_jac, t_lin; /* seconds */
} orc_newton_stats;
/* y[n] in: initial reduced guess (used verbatim, NR_Damp_solver_damped.rs:2034-2040); out: result */
int orc_newton_damped(const orc_bvp *b, const orc_newton_opts *o, const double *lo, const double *hi,
                      const double *rtol, double *y, orc_newton_stats *st);
/* frozen-Jacobian Newton (NR_Damp_solver_frozen.rs:1035-1148, BVP_utils.rs:362-480); strategy: 0 Naive,
 * 1 Frozen_naive, 2 every_m{m}, 3 at_high_morm{norm}, 4 at_low_speed{rate}, 5 complex{m,norm,rate} */
int orc_newton_frozen(const orc_bvp *b, int strategy, const double *strategy_params, double tolerance,
                      int max_iterations, int refactor_per_solve, double *y, orc_newton_stats *st);
/* re-insert BC values (BVP_utils.rs:534-558): full[(n_steps+1)*nv] */
void orc_construct_full_solution(const orc_bvp *b, const double *y_reduced, double *full);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
