/*
 * rst_oracle.c -- CPU restatement of the reference's damped-Newton BVP hot path.
 * TEST INFRASTRUCTURE ONLY (see rst_oracle.h).  kind = "port": the reference (Rust) cannot be built
 * here; this restatement is pinned by the reference's known-answer tests (tests/test_oracle_*.py).
 *
 * Paths below are relative to /root/reference/.
 */
#include "rst_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------------
 * Unknown layout.  src/numerical/BVP_Damp/numeric_discretization.rs:27-38 (boundary_position),
 * :40-78 (build_bc_layout: exactly nv conditions, sorted by position, duplicates rejected),
 * :80-106 (build_unknown_layout: node-major, BC positions deleted).
 * ---------------------------------------------------------------------------------------------- */
int orc_layout_build(int nv, int n_steps, int nbc, const int *bc_var, const int *bc_side,
                     const double *bc_val, int64_t *unknown_pos, int64_t *full_to_unknown,
                     int64_t *bc_pos_sorted, double *bc_val_sorted) {
    if (nbc != nv) return -1;
    for (int k = 0; k < nbc; ++k) {
        int64_t pos;
        if (bc_side[k] == 0) pos = bc_var[k];
        else if (bc_side[k] == 1) pos = (int64_t)n_steps * nv + bc_var[k];
        else return -3;
        /* insertion sort by position */
        int i = k;
        while (i > 0 && bc_pos_sorted[i - 1] > pos) {
            bc_pos_sorted[i] = bc_pos_sorted[i - 1];
            bc_val_sorted[i] = bc_val_sorted[i - 1];
            --i;
        }
        bc_pos_sorted[i] = pos;
        bc_val_sorted[i] = bc_val[k];
    }
    for (int k = 1; k < nbc; ++k)
        if (bc_pos_sorted[k] == bc_pos_sorted[k - 1]) return -2;
    const int64_t full_len = (int64_t)nv * (n_steps + 1);
    for (int64_t i = 0; i < full_len; ++i) full_to_unknown[i] = 0;
    for (int k = 0; k < nbc; ++k) full_to_unknown[bc_pos_sorted[k]] = -1;
    int64_t idx = 0;
    for (int64_t i = 0; i < full_len; ++i) {
        if (full_to_unknown[i] == -1) continue;
        unknown_pos[idx] = i;
        full_to_unknown[i] = idx++;
    }
    return 0;
}

/* numeric_discretization.rs:195-201 */
void orc_expand_full(const orc_bvp *b, const double *y, double *full) {
    const int64_t n = (int64_t)b->nv * b->n_steps;
    for (int k = 0; k < b->nv; ++k) full[b->bc_pos[k]] = b->bc_val[k];
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) full[b->unknown_pos[k]] = y[k];
}

/* BVP_utils.rs:534-558 (same result, without the reference's quadratic scan) */
void orc_construct_full_solution(const orc_bvp *b, const double *y, double *full) { orc_expand_full(b, y, full); }

/* ------------------------------------------------------------------------------------------------
 * Residual.  numeric_discretization.rs:203-240; symbolic twin symbolic_functions_BVP.rs:5486-5490
 * with eq_step :5221-5257.  Forward: r = Y[j+1] - Y[j] - h f(t_j, Y[j]).
 * Trapezoid: r = Y[j+1] - Y[j] - 0.5 h (f(t_j,Y[j]) + f(t*,Y[j+1])), t* = t_j on the symbolic route
 * (:5250-5253), t_{j+1} on the closure route (numeric_discretization.rs:223).
 * ---------------------------------------------------------------------------------------------- */
static void residual_full(const orc_bvp *b, const double *full, double *r) {
    const int nv = b->nv;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < b->n_steps; ++j) {
        double f0[64], f1[64];
        const double *y0 = full + j * nv, *y1 = full + (j + 1) * nv;
        const double x0 = b->mesh[j], x1 = b->mesh[j + 1];
        const double h = x1 - x0;
        b->pb->rhs(x0, y0, b->params, f0);
        if (b->scheme == 1) {
            b->pb->rhs(b->trap_time ? x1 : x0, y1, b->params, f1);
            for (int i = 0; i < nv; ++i) r[j * nv + i] = y1[i] - y0[i] - 0.5 * h * (f0[i] + f1[i]);
        } else {
            for (int i = 0; i < nv; ++i) r[j * nv + i] = y1[i] - y0[i] - h * f0[i];
        }
    }
}

void orc_residual(const orc_bvp *b, const double *y, double *r) {
    const int64_t full_len = (int64_t)b->nv * (b->n_steps + 1);
    double *full = (double *)malloc(sizeof(double) * (size_t)full_len);
    orc_expand_full(b, y, full);
    residual_full(b, full, r);
    free(full);
}

/* ------------------------------------------------------------------------------------------------
 * Jacobian structure and values.  numeric_discretization.rs:328-342 (entries), zero pruning is
 * structural as on the symbolic route (View/jacobian.rs:111-128); entry order = sorted (row, col)
 * (symbolic_functions_BVP.rs:1343); diagonal metadata :1338-1357; bandwidth :5639-5656.
 * For row (j,i): columns idx(j,c) for c with (c==i or pattern[i][c]) then idx(j+1,i) (forward) or
 * idx(j+1,c) for pattern/diag (trapezoid).  Since idx is increasing in full position the natural
 * emission order is already (row, col)-sorted.
 * ---------------------------------------------------------------------------------------------- */
int64_t orc_jacobian_structure(const orc_bvp *b, int64_t *rows, int64_t *cols, int64_t *doff,
                               int64_t *dpos, int *kl_out, int *ku_out) {
    const int nv = b->nv;
    unsigned char pat[64 * 64];
    orc_problem_pattern(b->pb, pat);
    int64_t nnz = 0, kl = 0, ku = 0;
    for (int64_t j = 0; j < b->n_steps; ++j) {
        for (int i = 0; i < nv; ++i) {
            const int64_t row = j * nv + i;
            for (int side = 0; side < 2; ++side) {
                for (int c = 0; c < nv; ++c) {
                    int present;
                    if (side == 0) present = (c == i) || pat[i * nv + c];
                    else present = (b->scheme == 1) ? ((c == i) || pat[i * nv + c]) : (c == i);
                    if (!present) continue;
                    const int64_t col = b->full_to_unknown[(j + side) * nv + c];
                    if (col < 0) continue;
                    if (rows) {
                        rows[nnz] = row;
                        cols[nnz] = col;
                        const int64_t off = col - row;
                        doff[nnz] = off;
                        dpos[nnz] = off >= 0 ? row : col;
                    }
                    if (col > row) { if (col - row > ku) ku = col - row; }
                    else if (row - col > kl) kl = row - col;
                    ++nnz;
                }
            }
        }
    }
    if (kl_out) *kl_out = (int)kl;
    if (ku_out) *ku_out = (int)ku;
    return nnz;
}

void orc_jacobian_values(const orc_bvp *b, const double *y, const int64_t *rows, const int64_t *cols,
                         int64_t nnz, double *values) {
    const int nv = b->nv;
    const int64_t full_len = (int64_t)nv * (b->n_steps + 1);
    double *full = (double *)malloc(sizeof(double) * (size_t)full_len);
    orc_expand_full(b, y, full);
    /* per-interval dense blocks, then gather in stream order */
#pragma omp parallel
    {
        double j0[64 * 64], j1[64 * 64];
        int64_t cached = -1;
#pragma omp for schedule(static)
        for (int64_t k = 0; k < nnz; ++k) {
            const int64_t row = rows[k], col = cols[k];
            const int64_t j = row / nv;
            const int i = (int)(row % nv);
            if (j != cached) {
                const double x0 = b->mesh[j], x1 = b->mesh[j + 1];
                b->pb->jac(x0, full + j * nv, b->params, j0);
                if (b->scheme == 1) b->pb->jac(b->trap_time ? x1 : x0, full + (j + 1) * nv, b->params, j1);
                cached = j;
            }
            const double h = b->mesh[j + 1] - b->mesh[j];
            const int64_t fpos = b->unknown_pos[col];
            const int64_t node = fpos / nv;
            const int c = (int)(fpos % nv);
            double v;
            if (node == j) {
                const double id = (i == c) ? -1.0 : 0.0;
                v = (b->scheme == 1) ? id - 0.5 * h * j0[i * nv + c] : id - h * j0[i * nv + c];
            } else {
                const double id = (i == c) ? 1.0 : 0.0;
                v = (b->scheme == 1) ? id - 0.5 * h * j1[i * nv + c] : id;
            }
            values[k] = v;
        }
    }
    free(full);
}

/* ------------------------------------------------------------------------------------------------
 * Band layouts.  BandedAssembly: src/somelinalg/banded/banded_assembly.rs:5-10, :18-23, :100-135;
 * scatter of the value stream: src/symbolic/codegen/codegen_runtime_api.rs:462-489.
 * ---------------------------------------------------------------------------------------------- */
int64_t orc_assembly_size(int64_t n, int kl, int ku, int64_t *diag_start) {
    int64_t total = 0;
    for (int d = -kl; d <= ku; ++d) {
        if (diag_start) diag_start[d + kl] = total;
        const int64_t ad = d < 0 ? -d : d;
        total += ad < n ? n - ad : 0;
    }
    return total;
}

void orc_assemble_diagonals(int64_t n, int kl, int ku, int64_t nnz, const int64_t *doff,
                            const int64_t *dpos, const double *values, double *diags) {
    int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(kl + ku + 1));
    const int64_t total = orc_assembly_size(n, kl, ku, start);
    memset(diags, 0, sizeof(double) * (size_t)total);
    for (int64_t k = 0; k < nnz; ++k) diags[start[doff[k] + kl] + dpos[k]] = values[k];
    free(start);
}

/* banded_assembly.rs:156-168 -> storage.rs:5-16 data[(ku+i-j)*n + j] */
void orc_assembly_to_banded(int64_t n, int kl, int ku, const double *diags, double *banded) {
    int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(kl + ku + 1));
    orc_assembly_size(n, kl, ku, start);
    memset(banded, 0, sizeof(double) * (size_t)(kl + ku + 1) * (size_t)n);
    for (int d = -kl; d <= ku; ++d) {
        const int64_t ad = d < 0 ? -d : d;
        const int64_t len = ad < n ? n - ad : 0;
        for (int64_t pos = 0; pos < len; ++pos) {
            const int64_t i = d >= 0 ? pos : pos + ad;
            const int64_t j = d >= 0 ? pos + ad : pos;
            banded[(ku + i - j) * n + j] = diags[start[d + kl] + pos];
        }
    }
    free(start);
}

/* lapack_style_banded.rs:246-278 load_from_banded (zero fill, then AB(kv+i-j, j) = A(i,j)) */
void orc_banded_to_ab(int64_t n, int kl, int ku, const double *banded, double *ab) {
    const int64_t kv = kl + ku, ldab = 2 * (int64_t)kl + ku + 1;
    memset(ab, 0, sizeof(double) * (size_t)ldab * (size_t)n);
    for (int64_t j = 0; j < n; ++j) {
        const int64_t i0 = j > ku ? j - ku : 0;
        const int64_t i1 = j + kl + 1 < n ? j + kl + 1 : n;
        for (int64_t i = i0; i < i1; ++i) ab[(kv + i - j) + j * ldab] = banded[(ku + i - j) * n + j];
    }
}

/* ops.rs:4 banded_matvec */
void orc_banded_matvec(int64_t n, int kl, int ku, const double *banded, const double *x, double *y) {
    for (int64_t i = 0; i < n; ++i) {
        const int64_t j0 = i > kl ? i - kl : 0;
        const int64_t j1 = i + ku + 1 < n ? i + ku + 1 : n;
        double s = 0.0;
        for (int64_t j = j0; j < j1; ++j) s += banded[(ku + i - j) * n + j] * x[j];
        y[i] = s;
    }
}

/* ------------------------------------------------------------------------------------------------
 * LapackStyleBandedLuFaithful.  lapack_style_banded.rs:386-447 (factor_one_step_unblocked_lapack:
 * zero fill column j+kv, IDAMAX with strict '>' :309-325, ZeroPivot if |pivot| <= 1e-14 :398-403,
 * ju update :404, row swap over columns j..ju :415-424, per-element division :330-351, rank-1
 * update skipping zero u / l :427-445).  The reference takes the blocked DGBTRF route for
 * kl >= 64 (:464-471, :477-947); blocked and unblocked produce the same pivots and the same factors
 * up to summation order, and this port always runs the unblocked DGBTF2 form.
 * ---------------------------------------------------------------------------------------------- */
int64_t orc_dgb_factor(int64_t n, int kl, int ku, double *ab, int64_t *ipiv, double eps) {
    const int64_t kv = (int64_t)kl + ku, ldab = 2 * (int64_t)kl + ku + 1;
    int64_t ju = 0;
    for (int64_t j = 0; j < n; ++j) {
        if (j + kv < n)
            for (int r = 0; r < kl; ++r) ab[r + (j + kv) * ldab] = 0.0;
        const int64_t km = (n - 1 - j) < kl ? (n - 1 - j) : kl;
        double *colj = ab + j * ldab;
        int64_t jp = 0;
        double best = fabs(colj[kv]);
        for (int64_t off = 1; off <= km; ++off) {
            const double v = fabs(colj[kv + off]);
            if (v > best) { best = v; jp = off; }
        }
        const int64_t p = j + jp;
        ipiv[j] = p;
        const double pivot = colj[kv + jp];
        if (fabs(pivot) <= eps) return j + 1;
        {
            const int64_t cand = (j + ku + jp) < (n - 1) ? (j + ku + jp) : (n - 1);
            if (cand > ju) ju = cand;
        }
        if (jp != 0) {
            for (int64_t col = j; col <= ju; ++col) {
                double *c = ab + col * ldab;
                const int64_t rj = kv + j - col, rp = kv + p - col;
                const double t = c[rj]; c[rj] = c[rp]; c[rp] = t;
            }
        }
        for (int64_t off = 1; off <= km; ++off) colj[kv + off] = colj[kv + off] / pivot;
        for (int64_t col = j + 1; col <= ju; ++col) {
            double *c = ab + col * ldab;
            const int64_t u_row = kv + j - col;
            if (u_row < 0) continue;
            const double u = c[u_row];
            if (u == 0.0) continue;
            for (int64_t off = 1; off <= km; ++off) {
                const double l = colj[kv + off];
                if (l == 0.0) continue;
                c[u_row + off] -= l * u;
            }
        }
    }
    return 0;
}

/* lapack_style_banded.rs:1163-1225 (DGBTRS 'N': interleaved swap + axpy forward :1179-1199,
 * back substitution over kv super-diagonals with the 1e-14 guard :1203-1222) */
int64_t orc_dgb_solve(int64_t n, int kl, int ku, const double *ab, const int64_t *ipiv, double *rhs,
                      double eps) {
    const int64_t kv = (int64_t)kl + ku, ldab = 2 * (int64_t)kl + ku + 1;
    if (kl > 0 && n > 1) {
        for (int64_t j = 0; j < n - 1; ++j) {
            const int64_t p = ipiv[j];
            if (p != j) { const double t = rhs[j]; rhs[j] = rhs[p]; rhs[p] = t; }
            const int64_t lm = (n - 1 - j) < kl ? (n - 1 - j) : kl;
            const double bj = rhs[j];
            if (bj == 0.0) continue;
            const double *colj = ab + j * ldab;
            for (int64_t off = 1; off <= lm; ++off) {
                const double l = colj[kv + off];
                if (l != 0.0) rhs[j + off] -= l * bj;
            }
        }
    }
    for (int64_t i = n - 1; i >= 0; --i) {
        const int64_t j_hi = (i + kv) < (n - 1) ? (i + kv) : (n - 1);
        double sum = rhs[i];
        for (int64_t j = i + 1; j <= j_hi; ++j) {
            const double u = ab[(kv + i - j) + j * ldab];
            if (u != 0.0) sum -= u * rhs[j];
        }
        const double uii = ab[kv + i * ldab];
        if (fabs(uii) <= eps) return i + 1;
        rhs[i] = sum / uii;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * BlockTridiagonalLuConsistent.  block_tridiagonal_lu_consistent.rs:98-164 (factor_from),
 * :166-218 (solve_in_place), helpers :587-842, dense LU dense_block_kernels.rs:285-330.
 * Restated bug-for-bug, including that solve applies the diagonal-block pivots to the rhs block
 * BEFORE subtracting lower_fac * prev (:188-194).
 * ---------------------------------------------------------------------------------------------- */
static int dense_lu_pivot(double *a, int bs, int64_t *piv) {
    for (int k = 0; k < bs; ++k) {
        int p = k;
        double mx = fabs(a[k * bs + k]);
        for (int i = k + 1; i < bs; ++i) {
            const double v = fabs(a[i * bs + k]);
            if (v > mx) { mx = v; p = i; }
        }
        if (mx <= 1e-14) return k + 1;
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < bs; ++j) { const double t = a[k * bs + j]; a[k * bs + j] = a[p * bs + j]; a[p * bs + j] = t; }
        const double pivot = a[k * bs + k];
        for (int i = k + 1; i < bs; ++i) {
            const double m = a[i * bs + k] / pivot;
            a[i * bs + k] = m;
            for (int j = k + 1; j < bs; ++j) a[i * bs + j] -= m * a[k * bs + j];
        }
    }
    return 0;
}

int64_t orc_btd_factor(int64_t nb, int bs, double *lower, double *diag, double *upper, int64_t *pivots) {
    const size_t bl = (size_t)bs * bs;
    int rc = dense_lu_pivot(diag, bs, pivots);
    if (rc) return rc;
    double *tmp = (double *)malloc(sizeof(double) * bl);
    for (int64_t k = 0; k + 1 < nb; ++k) {
        double *lu = diag + k * bl, *up = upper + k * bl, *lo = lower + k * bl, *nx = diag + (k + 1) * bl;
        const int64_t *pv = pivots + k * bs;
        /* apply_pivots_to_block_rows :600-615 */
        for (int r = 0; r < bs; ++r)
            if (pv[r] != r)
                for (int c = 0; c < bs; ++c) { const double t = up[r * bs + c]; up[r * bs + c] = up[pv[r] * bs + c]; up[pv[r] * bs + c] = t; }
        /* left_solve_unit_lower_block_in_place :657-678 */
        for (int row = 0; row < bs; ++row)
            for (int pc = 0; pc < row; ++pc) {
                const double l = lu[row * bs + pc];
                if (l == 0.0) continue;
                for (int c = 0; c < bs; ++c) up[row * bs + c] -= l * up[pc * bs + c];
            }
        /* right_solve_upper_block_in_place :708-741: each row x of lower solves x U = row */
        for (int r = 0; r < bs; ++r) {
            for (int i = 0; i < bs; ++i) {
                double sum = lo[r * bs + i];
                for (int j = 0; j < i; ++j) sum -= lu[j * bs + i] * tmp[j];
                const double uii = lu[i * bs + i];
                if (fabs(uii) <= 1e-14) { free(tmp); return k * bs + i + 1; }
                tmp[i] = sum / uii;
            }
            for (int i = 0; i < bs; ++i) lo[r * bs + i] = tmp[i];
        }
        /* dense_matmul_sub_assign_square :821-842 */
        for (int i = 0; i < bs; ++i)
            for (int kk = 0; kk < bs; ++kk) {
                const double a = lo[i * bs + kk];
                if (a == 0.0) continue;
                for (int j = 0; j < bs; ++j) nx[i * bs + j] -= a * up[kk * bs + j];
            }
        rc = dense_lu_pivot(nx, bs, pivots + (k + 1) * bs);
        if (rc) { free(tmp); return (k + 1) * bs + rc; }
    }
    free(tmp);
    return 0;
}

int64_t orc_btd_solve(int64_t nb, int bs, const double *lower, const double *diag, const double *upper,
                      const int64_t *pivots, double *rhs) {
    const size_t bl = (size_t)bs * bs;
    for (int64_t k = 0; k < nb; ++k) {
        double *cur = rhs + k * bs;
        const int64_t *pv = pivots + k * bs;
        for (int r = 0; r < bs; ++r)
            if (pv[r] != r) { const double t = cur[r]; cur[r] = cur[pv[r]]; cur[pv[r]] = t; }
        if (k > 0) {
            const double *lo = lower + (k - 1) * bl, *prev = rhs + (k - 1) * bs;
            for (int i = 0; i < bs; ++i) {
                double corr = 0.0;
                for (int j = 0; j < bs; ++j) corr += lo[i * bs + j] * prev[j];
                cur[i] -= corr;
            }
        }
        const double *lu = diag + k * bl;
        for (int i = 0; i < bs; ++i) {
            double sum = cur[i];
            for (int j = 0; j < i; ++j) sum -= lu[i * bs + j] * cur[j];
            cur[i] = sum;
        }
    }
    for (int64_t k = nb - 1; k >= 0; --k) {
        double *cur = rhs + k * bs;
        if (k + 1 < nb) {
            const double *up = upper + k * bl, *next = rhs + (k + 1) * bs;
            for (int i = 0; i < bs; ++i) {
                double corr = 0.0;
                for (int j = 0; j < bs; ++j) corr += up[i * bs + j] * next[j];
                cur[i] -= corr;
            }
        }
        const double *lu = diag + k * bl;
        for (int i = bs - 1; i >= 0; --i) {
            double sum = cur[i];
            for (int j = i + 1; j < bs; ++j) sum -= lu[i * bs + j] * cur[j];
            const double uii = lu[i * bs + i];
            if (fabs(uii) <= 1e-14) return k * bs + i + 1;
            cur[i] = sum / uii;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Damping helpers.  src/numerical/BVP_Damp/BVP_utils_damped.rs:39-61 (bound_step_Cantera2),
 * :209-222 (convergence_condition); norm = nalgebra Matrix::norm = sqrt(sum x^2) (BVP_traits.rs:283).
 * ---------------------------------------------------------------------------------------------- */
double orc_bound_step_cantera2(int64_t n, const double *x, const double *step, const double *lo,
                               const double *hi) {
    double fbound = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        const double val = x[i], newval = val + step[i];
        if (newval > hi[i]) {
            double c = (hi[i] - val) / (newval - val);
            fbound = fbound < c ? fbound : c;       /* f64::min */
            fbound = fbound > 0.0 ? fbound : 0.0;   /* .max(0.0) */
        } else if (newval < lo[i]) {
            double c = (val - lo[i]) / (val - newval);
            fbound = fbound < c ? fbound : c;
        }
    }
    return fbound;
}

double orc_convergence_condition(int64_t n, const double *y, double abs_tol, const double *rtol) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum += fabs(y[i]) * rtol[i];
    return sum > abs_tol ? sum : abs_tol;
}

double orc_norm2(int64_t n, const double *x) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}

/* ------------------------------------------------------------------------------------------------
 * Damped Newton driver.  src/numerical/BVP_Damp/NR_Damp_solver_damped.rs:1746-1771
 * (recalc_jacobian), :1782-1847 (step), :1862-2004 (damped_step), :2027-2156
 * (try_main_loop_damped), BVP_utils_damped.rs:185-207 (jac_recalc).
 * The linear solve is the faithful banded LU; with refactor_per_solve the band copy + factorisation
 * is redone for every solve exactly as BVP_traits.rs:875-899 / linear_solver.rs:198-210 do.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const orc_bvp *b;
    int64_t n, nnz;
    int kl, ku;
    int64_t *rows, *cols, *doff, *dpos, *ipiv;
    double *values, *diags, *banded, *ab;
    int have_factor;
    orc_newton_stats *st;
    const orc_newton_opts *o;
} nctx;

static int newton_step(nctx *c, const double *y, double *step) {
    double t0 = now_s();
    orc_residual(c->b, y, step); /* F_k */
    c->st->t_fun += now_s() - t0;
    for (int64_t i = 0; i < c->n; ++i)
        if (isnan(step[i])) return -10; /* panic: residual contains NaN (:1821-1826) */
    t0 = now_s();
    if (c->o->refactor_per_solve || !c->have_factor) {
        orc_assembly_to_banded(c->n, c->kl, c->ku, c->diags, c->banded);
        orc_banded_to_ab(c->n, c->kl, c->ku, c->banded, c->ab);
        if (orc_dgb_factor(c->n, c->kl, c->ku, c->ab, c->ipiv, 1e-14) != 0) return -11;
        c->have_factor = 1;
        c->st->factorizations++;
    }
    if (orc_dgb_solve(c->n, c->kl, c->ku, c->ab, c->ipiv, step, 1e-14) != 0) return -11;
    c->st->t_lin += now_s() - t0;
    for (int64_t i = 0; i < c->n; ++i)
        if (isnan(step[i])) return -12; /* panic: Newton update contains NaN (:1839-1844) */
    return 0;
}

int orc_newton_damped(const orc_bvp *b, const orc_newton_opts *o, const double *lo, const double *hi,
                      const double *rtol, double *y, orc_newton_stats *st) {
    nctx c;
    memset(&c, 0, sizeof(c));
    memset(st, 0, sizeof(*st));
    c.b = b; c.st = st; c.o = o;
    c.n = (int64_t)b->nv * b->n_steps;
    c.nnz = orc_jacobian_structure(b, NULL, NULL, NULL, NULL, &c.kl, &c.ku);
    c.rows = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    c.cols = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    c.doff = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    c.dpos = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    orc_jacobian_structure(b, c.rows, c.cols, c.doff, c.dpos, &c.kl, &c.ku);
    c.values = (double *)malloc(sizeof(double) * (size_t)c.nnz);
    c.diags = (double *)malloc(sizeof(double) * (size_t)orc_assembly_size(c.n, c.kl, c.ku, NULL));
    c.banded = (double *)malloc(sizeof(double) * (size_t)(c.kl + c.ku + 1) * (size_t)c.n);
    c.ab = (double *)malloc(sizeof(double) * (size_t)(2 * c.kl + c.ku + 1) * (size_t)c.n);
    c.ipiv = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.n);
    double *step0 = (double *)malloc(sizeof(double) * (size_t)c.n);
    double *step1 = (double *)malloc(sizeof(double) * (size_t)c.n);
    double *y1 = (double *)malloc(sizeof(double) * (size_t)c.n);

    int have_jac = 0, recalc = 1, m = 0, n_jac_reeval = 0, it = 0, status = 0, rc = 0;
    st->status = 0;
    while (it < o->max_iterations) {
        /* jac_recalc: BVP_utils_damped.rs:185-207 */
        recalc = (!have_jac || m > o->max_jac) ? 1 : recalc;
        if (recalc) { /* recalc_jacobian :1746-1771 */
            double t0 = now_s();
            orc_jacobian_values(b, y, c.rows, c.cols, c.nnz, c.values);
            orc_assemble_diagonals(c.n, c.kl, c.ku, c.nnz, c.doff, c.dpos, c.values, c.diags);
            st->t_jac += now_s() - t0;
            have_jac = 1; c.have_factor = 0; m = 0;
            st->jac_recalcs++;
        }
        m += 1; it += 1;
        st->iterations++;
        /* damped_step :1862-2004 */
        rc = newton_step(&c, y, step0);
        if (rc) { status = rc; break; }
        st->linear_solves++;
        const double fbound = orc_bound_step_cantera2(c.n, y, step0, lo, hi);
        if (isnan(fbound) || isinf(fbound)) { status = -13; break; }
        double lambda = 1.0 * fbound;
        if (fbound < 1e-10) status = -3;
        else {
            const double s0 = orc_norm2(c.n, step0);
            double s1 = 0.0, conv = 0.0;
            int k = 0, accepted = 0;
            while (k < o->max_damp_iter) {
                for (int64_t i = 0; i < c.n; ++i) y1[i] = y[i] - step0[i] * lambda; /* mul_float then sub */
                rc = newton_step(&c, y1, step1);
                if (rc) break;
                st->linear_solves++;
                s1 = orc_norm2(c.n, step1);
                conv = orc_convergence_condition(c.n, y1, o->abs_tol, rtol);
                if (s1 < 1.0 || s1 < s0) { accepted = 1; break; }
                lambda = lambda / pow(2.0, (double)k + o->damp_factor);
                k += 1;
            }
            if (rc) { status = rc; break; }
            st->last_s1 = s1; st->last_conv = conv;
            if (accepted) status = (s1 > conv) ? 0 : 1;
            else status = -2;
        }
        if (status == 0) {
            memcpy(y, y1, sizeof(double) * (size_t)c.n);
            recalc = 0;
        } else if (status == 1) {
            memcpy(y, y1, sizeof(double) * (size_t)c.n);
            break;
        } else { /* status < 0 :2119-2136 */
            if (m > 1) {
                recalc = 1;
                if (n_jac_reeval > 3) break;
                n_jac_reeval += 1;
            } else break;
        }
    }
    st->status = status;
    free(c.rows); free(c.cols); free(c.doff); free(c.dpos); free(c.values); free(c.diags);
    free(c.banded); free(c.ab); free(c.ipiv); free(step0); free(step1); free(y1);
    return status;
}

/* ------------------------------------------------------------------------------------------------
 * Frozen-Jacobian Newton.  src/numerical/BVP_Damp/NR_Damp_solver_frozen.rs:1035-1096 (iteration: F(y), Jacobian
 * when jac_recalc else m += 1, delta = J^-1 F, y_new = y - delta), :1105-1148 (try_main_loop: error = ||y_new - y||,
 * converged when error < tolerance), recalculation strategies BVP_utils.rs:362-480 (frozen_jac_recalc).
 * strategy: 0 Naive, 1 Frozen_naive, 2 every_m {m}, 3 at_high_morm {norm}, 4 at_low_speed {rate},
 * 5 complex {m, norm, rate}.  Linear solves refactor every time (BVP_traits.rs:875-899) when refactor_per_solve.
 * ---------------------------------------------------------------------------------------------- */
static int frozen_recalc(int strategy, const double *sp, int have_jac, int m, double error, double error_old) {
    switch (strategy) {
        case 0: return 1;
        case 1: return !have_jac;
        case 2: return (!have_jac || m > (int)sp[0]);
        case 3: return error > sp[0];
        case 4: return sp[0] * error_old < error;
        case 5: return (error > sp[1]) || (m > (int)sp[0]) || (sp[2] * error_old < error);
        default: return 1;
    }
}

int orc_newton_frozen(const orc_bvp *b, int strategy, const double *sp, double tolerance, int max_iterations,
                      int refactor_per_solve, double *y, orc_newton_stats *st) {
    nctx c;
    orc_newton_opts o;
    memset(&c, 0, sizeof(c));
    memset(&o, 0, sizeof(o));
    memset(st, 0, sizeof(*st));
    o.refactor_per_solve = refactor_per_solve;
    c.b = b; c.st = st; c.o = &o;
    c.n = (int64_t)b->nv * b->n_steps;
    c.nnz = orc_jacobian_structure(b, NULL, NULL, NULL, NULL, &c.kl, &c.ku);
    c.rows = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    c.cols = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    c.doff = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    c.dpos = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.nnz);
    orc_jacobian_structure(b, c.rows, c.cols, c.doff, c.dpos, &c.kl, &c.ku);
    c.values = (double *)malloc(sizeof(double) * (size_t)c.nnz);
    c.diags = (double *)malloc(sizeof(double) * (size_t)orc_assembly_size(c.n, c.kl, c.ku, NULL));
    c.banded = (double *)malloc(sizeof(double) * (size_t)(c.kl + c.ku + 1) * (size_t)c.n);
    c.ab = (double *)malloc(sizeof(double) * (size_t)(2 * c.kl + c.ku + 1) * (size_t)c.n);
    c.ipiv = (int64_t *)malloc(sizeof(int64_t) * (size_t)c.n);
    double *delta = (double *)malloc(sizeof(double) * (size_t)c.n);
    int jac_recalc = 1, have_jac = 0, m = 0, status = 0, it = 0;
    double error_old = 0.0;
    while (it < max_iterations) {
        st->iterations++;
        if (jac_recalc) {
            orc_jacobian_values(b, y, c.rows, c.cols, c.nnz, c.values);
            orc_assemble_diagonals(c.n, c.kl, c.ku, c.nnz, c.doff, c.dpos, c.values, c.diags);
            have_jac = 1; c.have_factor = 0; m = 0;
            st->jac_recalcs++;
        } else {
            m += 1;
        }
        const int rc = newton_step(&c, y, delta); /* F(y) then delta = J^-1 F */
        if (rc) { status = rc; break; }
        st->linear_solves++;
        for (int64_t i = 0; i < c.n; ++i) y[i] = y[i] - delta[i];
        const double error = orc_norm2(c.n, delta);
        jac_recalc = frozen_recalc(strategy, sp, have_jac, m, error, error_old);
        error_old = error;
        st->last_s1 = error;
        if (error < tolerance) { status = 1; break; }
        it += 1;
    }
    st->status = status;
    free(c.rows); free(c.cols); free(c.doff); free(c.dpos); free(c.values); free(c.diags);
    free(c.banded); free(c.ab); free(c.ipiv); free(delta);
    return status;
}
