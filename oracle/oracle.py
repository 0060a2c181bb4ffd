"""ctypes binding of the CPU oracle (oracle/rst_oracle.{h,c}, oracle/problems.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package ``rustedscithe_b200``.
kind = "port" (the Rust reference cannot be built in this image); pinned against the reference's
known-answer tests in tests/test_oracle_*.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librst_oracle.so")
_lib = None

c_i64_p = C.POINTER(C.c_int64)
c_f64_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the oracle with the system gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("rst_oracle.c", "problems.c", "rst_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "librst_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Problem(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("nv", C.c_int),
        ("n_params", C.c_int),
        ("rhs", C.c_void_p),
        ("jac", C.c_void_p),
    ]


class _Bvp(C.Structure):
    _fields_ = [
        ("pb", C.POINTER(_Problem)),
        ("params", c_f64_p),
        ("nv", C.c_int),
        ("n_steps", C.c_int),
        ("scheme", C.c_int),
        ("trap_time", C.c_int),
        ("mesh", c_f64_p),
        ("unknown_pos", c_i64_p),
        ("full_to_unknown", c_i64_p),
        ("bc_pos", c_i64_p),
        ("bc_val", c_f64_p),
    ]


class _NewtonOpts(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int),
        ("max_jac", C.c_int),
        ("max_damp_iter", C.c_int),
        ("damp_factor", C.c_double),
        ("abs_tol", C.c_double),
        ("refactor_per_solve", C.c_int),
    ]


class _NewtonStats(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("iterations", C.c_int),
        ("linear_solves", C.c_int),
        ("jac_recalcs", C.c_int),
        ("factorizations", C.c_int),
        ("last_s1", C.c_double),
        ("last_conv", C.c_double),
        ("t_fun", C.c_double),
        ("t_jac", C.c_double),
        ("t_lin", C.c_double),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_problem_get.restype = C.POINTER(_Problem)
        L.orc_problem_get.argtypes = [C.c_char_p]
        L.orc_jacobian_structure.restype = C.c_int64
        L.orc_assembly_size.restype = C.c_int64
        L.orc_dgb_factor.restype = C.c_int64
        L.orc_dgb_solve.restype = C.c_int64
        L.orc_btd_factor.restype = C.c_int64
        L.orc_btd_solve.restype = C.c_int64
        L.orc_bound_step_cantera2.restype = C.c_double
        L.orc_convergence_condition.restype = C.c_double
        L.orc_norm2.restype = C.c_double
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=c_f64_p):
    return a.ctypes.data_as(t)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def problem_rhs(name: str, t: float, y, params=None):
    """Evaluate the hand-written continuous f and df/dy of a built-in oracle problem."""
    L = lib()
    pb = L.orc_problem_get(name.encode())
    if not pb:
        raise KeyError(name)
    nv = pb.contents.nv
    y = _f64(y)
    p = _f64(params if params is not None else np.zeros(max(1, pb.contents.n_params)))
    f = np.zeros(nv)
    fy = np.zeros((nv, nv))
    RHS = C.CFUNCTYPE(None, C.c_double, c_f64_p, c_f64_p, c_f64_p)
    RHS(pb.contents.rhs)(t, _p(y), _p(p), _p(f))
    RHS(pb.contents.jac)(t, _p(y), _p(p), _p(fy))
    return f, fy


def problem_pattern(name: str):
    L = lib()
    pb = L.orc_problem_get(name.encode())
    nv = pb.contents.nv
    pat = np.zeros((nv, nv), dtype=np.uint8)
    L.orc_problem_pattern(pb, pat.ctypes.data_as(C.c_void_p))
    return pat


@dataclass
class NewtonResult:
    status: int
    iterations: int
    linear_solves: int
    jac_recalcs: int
    factorizations: int
    last_s1: float
    last_conv: float
    t_fun: float
    t_jac: float
    t_lin: float
    y: np.ndarray


FROZEN_STRATEGIES = {"Naive": 0, "Frozen_naive": 1, "every_m": 2, "at_high_morm": 3, "at_low_speed": 4, "complex": 5}


class OracleBVP:
    """Discretised BVP on the CPU oracle: layout, residual, Jacobian stream, band layouts, Newton."""

    def __init__(self, problem: str, n_steps: int, bcs, t0=0.0, t_end=1.0, mesh=None, scheme="forward",
                 trap_time=0, params=None):
        L = lib()
        self.L = L
        self.pb = L.orc_problem_get(problem.encode())
        if not self.pb:
            raise KeyError(problem)
        self.nv = nv = self.pb.contents.nv
        self.n_steps = n_steps
        self.n = nv * n_steps
        if mesh is None:
            # NR_Damp_solver_damped.rs:550-553, symbolic_functions_BVP.rs:5300-5308: t0 + i*h
            h = (t_end - t0) / n_steps
            mesh = t0 + np.arange(n_steps + 1, dtype=np.float64) * h
        self.mesh = _f64(mesh)
        assert self.mesh.shape == (n_steps + 1,)
        self.params = _f64(params if params is not None else np.zeros(max(1, self.pb.contents.n_params)))
        bc_var = np.array([b[0] for b in bcs], dtype=np.int32)
        bc_side = np.array([b[1] for b in bcs], dtype=np.int32)
        bc_val = _f64([b[2] for b in bcs])
        self.unknown_pos = np.zeros(self.n, dtype=np.int64)
        self.full_to_unknown = np.zeros(nv * (n_steps + 1), dtype=np.int64)
        self.bc_pos = np.zeros(nv, dtype=np.int64)
        self.bc_val = np.zeros(nv, dtype=np.float64)
        rc = L.orc_layout_build(nv, n_steps, len(bcs), _p(bc_var, c_int_p), _p(bc_side, c_int_p),
                                _p(bc_val), _p(self.unknown_pos, c_i64_p), _p(self.full_to_unknown, c_i64_p),
                                _p(self.bc_pos, c_i64_p), _p(self.bc_val))
        if rc != 0:
            raise ValueError({-1: f"expects exactly {nv} total boundary conditions", -2: "duplicate boundary condition",
                              -3: "boundary side flag must be 0 or 1"}[rc])
        self.bvp = _Bvp(self.pb, _p(self.params), nv, n_steps, 1 if scheme == "trapezoid" else 0, trap_time,
                        _p(self.mesh), _p(self.unknown_pos, c_i64_p), _p(self.full_to_unknown, c_i64_p),
                        _p(self.bc_pos, c_i64_p), _p(self.bc_val))
        self._structure = None

    # -- residual / Jacobian ------------------------------------------------------------------
    def residual(self, y):
        y = _f64(y)
        assert y.shape == (self.n,)
        r = np.zeros(self.n)
        self.L.orc_residual(C.byref(self.bvp), _p(y), _p(r))
        return r

    def structure(self):
        if self._structure is None:
            kl, ku = C.c_int(), C.c_int()
            nnz = self.L.orc_jacobian_structure(C.byref(self.bvp), None, None, None, None, C.byref(kl), C.byref(ku))
            rows = np.zeros(nnz, dtype=np.int64)
            cols = np.zeros(nnz, dtype=np.int64)
            doff = np.zeros(nnz, dtype=np.int64)
            dpos = np.zeros(nnz, dtype=np.int64)
            self.L.orc_jacobian_structure(C.byref(self.bvp), _p(rows, c_i64_p), _p(cols, c_i64_p),
                                          _p(doff, c_i64_p), _p(dpos, c_i64_p), C.byref(kl), C.byref(ku))
            self._structure = (rows, cols, doff, dpos, kl.value, ku.value)
        return self._structure

    def jacobian_values(self, y):
        rows, cols, _, _, _, _ = self.structure()
        y = _f64(y)
        vals = np.zeros(rows.shape[0])
        self.L.orc_jacobian_values(C.byref(self.bvp), _p(y), _p(rows, c_i64_p), _p(cols, c_i64_p),
                                   C.c_int64(rows.shape[0]), _p(vals))
        return vals

    def full_solution(self, y):
        full = np.zeros(self.nv * (self.n_steps + 1))
        self.L.orc_construct_full_solution(C.byref(self.bvp), _p(_f64(y)), _p(full))
        return full.reshape(self.n_steps + 1, self.nv)

    # -- band layouts + faithful LU -----------------------------------------------------------
    def banded(self, values):
        _, _, doff, dpos, kl, ku = self.structure()
        return assemble_banded(self.n, kl, ku, doff, dpos, values)

    def solve_banded(self, values, rhs):
        """J dy = rhs through BandedAssembly -> Banded -> AB -> DGBTF2 -> DGBTRS."""
        _, _, _, _, kl, ku = self.structure()
        banded = self.banded(values)
        ab, ipiv = dgb_factor(self.n, kl, ku, banded_to_ab(self.n, kl, ku, banded))
        return dgb_solve(self.n, kl, ku, ab, ipiv, rhs)

    def solve_banded_with_refinement(self, values, rhs, max_iters=2, tol=1e-14):
        """solve_in_place_with_refinement (lapack_style_banded.rs:1227-1298) with the banded matvec of the same matrix:
        a refinement step is accepted only while max|b - A x| / max(max|b|, 1) decreases.  Returns (x, accepted)."""
        _, _, _, _, kl, ku = self.structure()
        banded = self.banded(values)
        ab, ipiv = dgb_factor(self.n, kl, ku, banded_to_ab(self.n, kl, ku, banded))
        b = _f64(rhs).copy()
        x = dgb_solve(self.n, kl, ku, ab, ipiv, b)
        b_norm = max(float(np.max(np.abs(b))), 1.0)
        current_rr = np.inf
        accepted = 0
        for _ in range(max_iters):
            residual = b - banded_matvec(self.n, kl, ku, banded, x)
            rr = float(np.max(np.abs(residual))) / b_norm
            if not np.isfinite(rr) or rr < tol:
                break
            if not np.isfinite(current_rr):
                current_rr = rr
            x_trial = x + dgb_solve(self.n, kl, ku, ab, ipiv, residual)
            trial_rr = float(np.max(np.abs(b - banded_matvec(self.n, kl, ku, banded, x_trial)))) / b_norm
            if not np.isfinite(trial_rr) or trial_rr >= min(rr, current_rr):
                break
            x, current_rr = x_trial, trial_rr
            accepted += 1
        return x, accepted

    # -- Newton ------------------------------------------------------------------------------
    def newton(self, y0, lo, hi, rtol, abs_tol=1e-6, max_iterations=25, max_jac=3, max_damp_iter=5,
               damp_factor=0.5, refactor_per_solve=True) -> NewtonResult:
        y = _f64(y0).copy()
        lo, hi, rtol = _f64(lo), _f64(hi), _f64(rtol)
        assert y.shape == lo.shape == hi.shape == rtol.shape == (self.n,)
        o = _NewtonOpts(max_iterations, max_jac, max_damp_iter, damp_factor, abs_tol, 1 if refactor_per_solve else 0)
        st = _NewtonStats()
        self.L.orc_newton_damped(C.byref(self.bvp), C.byref(o), _p(lo), _p(hi), _p(rtol), _p(y), C.byref(st))
        return NewtonResult(st.status, st.iterations, st.linear_solves, st.jac_recalcs, st.factorizations,
                            st.last_s1, st.last_conv, st.t_fun, st.t_jac, st.t_lin, y)

    def newton_frozen(self, y0, strategy="Frozen_naive", strategy_params=(), tolerance=1e-6, max_iterations=25,
                      refactor_per_solve=True) -> NewtonResult:
        sid = FROZEN_STRATEGIES[strategy]
        y = _f64(y0).copy()
        sp = _f64(list(strategy_params) + [0.0] * (3 - len(strategy_params)))
        st = _NewtonStats()
        self.L.orc_newton_frozen(C.byref(self.bvp), sid, _p(sp), C.c_double(tolerance), max_iterations,
                                 1 if refactor_per_solve else 0, _p(y), C.byref(st))
        return NewtonResult(st.status, st.iterations, st.linear_solves, st.jac_recalcs, st.factorizations,
                            st.last_s1, st.last_conv, st.t_fun, st.t_jac, st.t_lin, y)

    def per_unknown(self, per_var):
        """bounds_vec / rel_tolerance_vec lookup by base variable (numeric_discretization.rs:108-130)."""
        per_var = np.asarray(per_var)
        return per_var[self.unknown_pos % self.nv]


# ---- free functions over band storage -----------------------------------------------------------
def assemble_banded(n, kl, ku, doff, dpos, values):
    L = lib()
    size = L.orc_assembly_size(C.c_int64(n), kl, ku, None)
    diags = np.zeros(size)
    values = _f64(values)
    L.orc_assemble_diagonals(C.c_int64(n), kl, ku, C.c_int64(values.shape[0]), _p(doff, c_i64_p), _p(dpos, c_i64_p),
                             _p(values), _p(diags))
    banded = np.zeros((kl + ku + 1) * n)
    L.orc_assembly_to_banded(C.c_int64(n), kl, ku, _p(diags), _p(banded))
    return banded


def banded_from_dense(a, kl, ku):
    """Banded<f64> (storage.rs:5-16) from a dense matrix: data[(ku+i-j)*n + j]."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    data = np.zeros((kl + ku + 1) * n)
    for i in range(n):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            data[(ku + i - j) * n + j] = a[i, j]
    return data


def banded_to_ab(n, kl, ku, banded):
    ab = np.zeros((2 * kl + ku + 1) * n)
    lib().orc_banded_to_ab(C.c_int64(n), kl, ku, _p(_f64(banded)), _p(ab))
    return ab


def dgb_factor(n, kl, ku, ab, eps=1e-14):
    ab = _f64(ab).copy()
    ipiv = np.zeros(n, dtype=np.int64)
    info = lib().orc_dgb_factor(C.c_int64(n), kl, ku, _p(ab), _p(ipiv, c_i64_p), C.c_double(eps))
    if info != 0:
        raise ZeroDivisionError(f"ZeroPivot at column {info - 1}")
    return ab, ipiv


def dgb_solve(n, kl, ku, ab, ipiv, rhs, eps=1e-14):
    x = _f64(rhs).copy()
    info = lib().orc_dgb_solve(C.c_int64(n), kl, ku, _p(ab), _p(ipiv, c_i64_p), _p(x), C.c_double(eps))
    if info != 0:
        raise ZeroDivisionError(f"ZeroPivot at row {info - 1}")
    return x


def banded_matvec(n, kl, ku, banded, x):
    y = np.zeros(n)
    lib().orc_banded_matvec(C.c_int64(n), kl, ku, _p(_f64(banded)), _p(_f64(x)), _p(y))
    return y


def btd_factor(lower, diag, upper):
    """BlockTridiagonalLuConsistent::factor_from; arrays (nb-1|nb, bs, bs) row-major."""
    diag = _f64(diag).copy()
    nb, bs = diag.shape[0], diag.shape[1]
    lower = _f64(lower).copy() if nb > 1 else np.zeros((0, bs, bs))
    upper = _f64(upper).copy() if nb > 1 else np.zeros((0, bs, bs))
    piv = np.zeros((nb, bs), dtype=np.int64)
    info = lib().orc_btd_factor(C.c_int64(nb), bs, _p(lower), _p(diag), _p(upper), _p(piv, c_i64_p))
    if info != 0:
        raise ZeroDivisionError(f"ZeroPivot at {info - 1}")
    return lower, diag, upper, piv


def btd_solve(fac, rhs):
    lower, diag, upper, piv = fac
    nb, bs = diag.shape[0], diag.shape[1]
    x = _f64(rhs).copy()
    info = lib().orc_btd_solve(C.c_int64(nb), bs, _p(lower), _p(diag), _p(upper), _p(piv, c_i64_p), _p(x))
    if info != 0:
        raise ZeroDivisionError(f"ZeroPivot at {info - 1}")
    return x


def bound_step_cantera2(x, step, bounds):
    x, step = _f64(x), _f64(step)
    lo = _f64([b[0] for b in bounds])
    hi = _f64([b[1] for b in bounds])
    return float(lib().orc_bound_step_cantera2(C.c_int64(x.shape[0]), _p(x), _p(step), _p(lo), _p(hi)))


def convergence_condition(y, abs_tol, rtol):
    y, rtol = _f64(y), _f64(rtol)
    return float(lib().orc_convergence_condition(C.c_int64(y.shape[0]), _p(y), C.c_double(abs_tol), _p(rtol)))
