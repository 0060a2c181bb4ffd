"""Host-side mirror of the reference's solver interface for the damped BVP hot path, over the C ABI.

Names and argument meaning follow the reference so parity tests read like its own tests:
  NRBVP.new_with_options / try_solve / get_result / get_statistics
      src/numerical/BVP_Damp/NR_Damp_solver_damped.rs:847, :2393, :2507, :1577-1689
  Fun.call / Jac.call (residual / Jacobian callbacks)          BVP_traits.rs:459-461, :943-953
  DirectLinearSolver.solve_in_place / solve_multiple_in_place  somelinalg/banded/solver_traits.rs:3-14
Everything numerical happens in librst_b200.so on the GPU; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib
from . import problems as _problems


GRID_METHODS = {"DoublePoints": 0, "Easy": 1, "Pearson": 2, "GrcarSmooke": 3, "Sci": 4}  # grid_api.rs:62-115


@dataclass
class AdaptiveGridConfig:  # NR_Damp_solver_damped.rs:115-122
    version: int = 1
    max_refinements: int = 3
    grid_method: tuple = ("DoublePoints",)   # (GridRefinementMethod name, *parameters)


@dataclass
class SolverParams:  # NR_Damp_solver_damped.rs:100-151
    max_jac: int = 3
    max_damp_iter: int = 5
    damp_factor: float = 0.5
    adaptive: AdaptiveGridConfig | None = None


@dataclass
class DampedSolverOptions:  # NR_Damp_solver_damped.rs:155-178
    strategy_params: SolverParams = field(default_factory=SolverParams)
    abs_tolerance: float = 1e-6
    rel_tolerance: dict | None = None
    bounds: dict | None = None
    max_iterations: int = 25
    scheme: str = "forward"
    trap_time: int = 0
    slab: int = 0
    device: int = 0
    stream: int | None = None
    rank: int = 0
    world: int = 1

    @classmethod
    def from_spec(cls, spec, **kw):
        return cls(strategy_params=SolverParams(spec.max_jac, spec.max_damp_iter, spec.damp_factor),
                   abs_tolerance=spec.abs_tol, rel_tolerance=dict(spec.rtol), bounds=dict(spec.bounds),
                   max_iterations=spec.max_iterations, **kw)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=_lib.c_f64_p):
    return a.ctypes.data_as(t)


class NRBVP:
    """Damped Newton BVP solver bound to one compiled-in AOT problem (device-resident state)."""

    def __init__(self, problem, initial_guess, n_steps, options: DampedSolverOptions | None = None,
                 t0=None, t_end=None, mesh=None, boundary_conditions=None, params=None,
                 aot_build_policy="build_if_missing"):
        self.spec = _problems.get(problem) if isinstance(problem, str) else problem
        spec = self.spec
        self.options = options or DampedSolverOptions.from_spec(spec)
        o = self.options
        self.L = _lib.load()
        if not _lib.has_problem(self.L, spec.key):
            # AOT lifecycle (generated_solver_handoff.rs:321-331): a problem outside the default library gets its
            # own artefact, built once (BuildIfMissing) or required to exist (RequirePrebuilt)
            from . import build as _build
            self.L = _lib.load(_build.build_artifact(spec, aot_build_policy))
        self.nv = spec.nv
        self.n_steps = int(n_steps)
        self.n = self.nv * self.n_steps
        bcs = boundary_conditions if boundary_conditions is not None else spec.bcs
        bc_list = []
        for i, name in enumerate(spec.names):
            for side, val in bcs.get(name, []):
                bc_list.append((i, side, val))
        bounds = o.bounds or spec.bounds
        rtol = o.rel_tolerance or spec.rtol
        self._keep = dict(
            bc_var=np.array([b[0] for b in bc_list], dtype=np.int32),
            bc_side=np.array([b[1] for b in bc_list], dtype=np.int32),
            bc_val=_f64([b[2] for b in bc_list]),
            lo=_f64([bounds[n][0] for n in spec.names]),
            hi=_f64([bounds[n][1] for n in spec.names]),
            rtol=_f64([rtol[n] for n in spec.names]),
            params=_f64(params if params is not None else (spec.param_values or [0.0])),
            mesh=_f64(mesh) if mesh is not None else None,
            key=spec.key.encode(),
        )
        k = self._keep
        if k["mesh"] is not None and k["mesh"].shape != (self.n_steps + 1,):
            raise ValueError(f"mesh length mismatch: expected {self.n_steps + 1}, got {k['mesh'].shape[0]}")
        d = _lib.Desc()
        d.problem_key = k["key"]
        d.n_steps = self.n_steps
        d.scheme = 1 if o.scheme == "trapezoid" else 0
        d.trap_time = o.trap_time
        d.t0 = spec.t0 if t0 is None else t0
        d.t_end = spec.t_end if t_end is None else t_end
        d.mesh = _p(k["mesh"]) if k["mesh"] is not None else None
        d.n_bc = len(bc_list)
        d.bc_var = _p(k["bc_var"], _lib.c_int_p)
        d.bc_side = _p(k["bc_side"], _lib.c_int_p)
        d.bc_val = _p(k["bc_val"])
        d.params = _p(k["params"])
        d.n_params = len(spec.params)
        d.lo, d.hi, d.rtol = _p(k["lo"]), _p(k["hi"]), _p(k["rtol"])
        d.device = o.device
        d.stream = o.stream
        d.slab = o.slab
        d.rank, d.world = o.rank, o.world
        self.handle = C.c_void_p()
        self._ck(self.L.rst_b200_create(C.byref(d), C.byref(self.handle)))
        self.stats = _lib.NewtonStats()
        self.result = None
        self._allgather_cb = None
        if initial_guess is not None:
            # column-major nv x n_steps guess flattened verbatim (NR_Damp_solver_damped.rs:2034-2040)
            g = np.asarray(initial_guess, dtype=np.float64)
            if g.ndim == 2:
                g = g.reshape(-1, order="F")
            self.set_state(g)

    def _ck(self, rc):
        return _lib.check(rc, self.L)

    @classmethod
    def new_with_options(cls, problem, initial_guess, n_steps, options=None, **kw):
        return cls(problem, initial_guess, n_steps, options, **kw)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.L.rst_b200_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ---------------------------------------------------------------------------------
    def set_state(self, y_reduced):
        y = _f64(y_reduced)
        self._ck(self.L.rst_b200_set_state(self.handle, _p(y), y.shape[0]))

    def get_state(self):
        y = np.zeros(self.n)
        self._ck(self.L.rst_b200_get_state(self.handle, _p(y), y.shape[0]))
        return y

    def set_params(self, params):
        p = _f64(params)
        self._ck(self.L.rst_b200_set_params(self.handle, _p(p), p.shape[0]))

    def device_ptr(self, which):
        return self.L.rst_b200_device_ptr(self.handle, which)

    def _opts(self):
        o = self.options
        return _lib.NewtonOpts(o.max_iterations, o.strategy_params.max_jac, o.strategy_params.max_damp_iter,
                               o.strategy_params.damp_factor, o.abs_tolerance)

    # -- reference solver surface -----------------------------------------------------------
    def try_solve(self):
        """try_solve -> try_main_loop_damped (+ try_solve_with_new_grid when strategy_params.adaptive is set);
        returns the reduced solution or None (as the reference)."""
        adaptive = self.options.strategy_params.adaptive
        new_grid_enabled = adaptive is not None                     # NR_Damp_solver_damped.rs:662-666
        self.grid_refinements, self.nodes_added = 0, []
        self.total_statistics = {"number of iterations": 0, "number of solving linear systems": 0,
                                 "number of jacobians recalculations": 0, "number of grid refinements": 0}
        self.result = None
        while True:
            opts = self._opts()
            self._ck(self.L.rst_b200_newton_solve(self.handle, C.byref(opts), C.byref(self.stats)))
            t = self.total_statistics
            t["number of iterations"] += self.stats.iterations
            t["number of solving linear systems"] += self.stats.linear_solves
            t["number of jacobians recalculations"] += self.stats.jac_recalcs
            if self.stats.status != 1:
                # Deliberate deviation, documented in DESIGN.md: after a FAILED solve the reference calls
                # try_solve_with_new_grid (:2145-2154), whose create_new_grid unwraps self.result -- a panic when no grid has
                # converged yet, and after a refinement a size mismatch between the saved result (previous mesh) and the
                # already replaced x_mesh (DMatrix::from_column_slice panics, :2216-2217).  There is no defined behaviour
                # to mirror, so a failed solve ends the adaptive loop with None and keeps the last converged result.
                self.result = None if self.grid_refinements == 0 else self.result
                return None
            self.result = self.get_state()
            if not (new_grid_enabled and adaptive is not None):      # :2102-2113
                return self.result
            # ---- try_solve_with_new_grid (:2281-2340) ----
            if adaptive.version != 1:
                raise ValueError("unsupported adaptive version")    # we_need_refinement panics (:2182)
            name, *gp = adaptive.grid_method
            _, n_added = self.new_grid(name, gp, fetch=False)
            self.grid_refinements += 1
            self.nodes_added.append(n_added)
            t["number of grid refinements"] += 1
            new_grid_enabled = (n_added != 0) and not (adaptive.max_refinements <= self.grid_refinements)  # :2164-2196
            if n_added == 0:
                return None                                          # "no new grid needed" returns Ok(None) (:2333-2336)
            self._adopt_refined()

    def new_grid(self, method, params=(), fetch=True):
        """new_grid (grid_api.rs:154) on the CURRENT device state.  Returns (n_points_new, n_added) or, with fetch,
        (mesh, full solution (n_points_new x nv), n_added, marks)."""
        m = GRID_METHODS[method] if isinstance(method, str) else int(method)
        p = _f64(list(params) + [0.0] * (3 - len(params)))
        n_new, n_added = C.c_int64(), C.c_int64()
        self._ck(self.L.rst_b200_new_grid(self.handle, m, _p(p), self.options.abs_tolerance, C.byref(n_new), C.byref(n_added)))
        if not fetch:
            return n_new.value, n_added.value
        mesh = np.zeros(n_new.value)
        yfull = np.zeros(n_new.value * self.nv)
        marks = np.zeros(self.n_steps + 1, dtype=np.int32)
        self._ck(self.L.rst_b200_new_grid_fetch(self.handle, _p(mesh), mesh.shape[0], _p(yfull), yfull.shape[0],
                                                 marks.ctypes.data_as(C.POINTER(C.c_int32)), marks.shape[0]))
        return mesh, yfull.reshape(n_new.value, self.nv), n_added.value, marks

    def _adopt_refined(self):
        """Continue on the refined mesh: new handle, state = interpolated solution (stays on the device)."""
        new = C.c_void_p()
        self._ck(self.L.rst_b200_create_refined(self.handle, C.byref(new)))
        n_new, _ = C.c_int64(), None
        self.L.rst_b200_destroy(self.handle)
        self.handle = new
        self.n = int(self.L.rst_b200_n_unknowns(self.handle))
        self.n_steps = self.n // self.nv

    def try_solve_frozen(self, strategy="Frozen_naive", strategy_params=(), tolerance=1e-6, max_iterations=25):
        """Frozen-Jacobian Newton (NR_Damp_solver_frozen.rs: NRBVP::try_solve with FrozenSolverOptions)."""
        o = _lib.FrozenOpts()
        o.strategy = _lib.FROZEN_STRATEGIES[strategy]
        for i, v in enumerate(strategy_params):
            o.strategy_params[i] = float(v)
        o.tolerance, o.max_iterations = tolerance, max_iterations
        self._ck(self.L.rst_b200_newton_frozen(self.handle, C.byref(o), C.byref(self.stats)))
        self.result = self.get_state() if self.stats.status == 1 else None
        return self.result

    def newton_iterate(self, recalc_jacobian=True):
        opts = self._opts()
        self._ck(self.L.rst_b200_newton_iterate(self.handle, C.byref(opts), 1 if recalc_jacobian else 0,
                                                  C.byref(self.stats)))
        return self.stats.status

    def get_result(self):
        """(N+1) x nv matrix with boundary values re-inserted (NR_Damp_solver_damped.rs:2540-2558)."""
        full = np.zeros((self.n_steps + 1) * self.nv)
        self._ck(self.L.rst_b200_get_full_solution(self.handle, _p(full), full.shape[0]))
        return full.reshape(self.n_steps + 1, self.nv)

    def get_statistics(self):
        s = self.stats
        return {"number of iterations": s.iterations, "number of solving linear systems": s.linear_solves,
                "number of jacobians recalculations": s.jac_recalcs, "factorizations": s.factorizations,
                "solves reused": s.solves_reused,
                "number of grid refinements": getattr(self, "grid_refinements", 0),
                "status": s.status, "ms_total": s.ms_total}

    # -- Fun / Jac callbacks through the reference's AOT ABI (host pointers) -----------------
    def fun_call(self, y_reduced, params=None):
        p = _f64(params) if params is not None else (self._keep["params"] if self.spec.params else np.zeros(0))
        args = np.concatenate([p, _f64(y_reduced)])
        out = np.zeros(self.n)
        self.L.rst_b200_aot_bind(self.handle)
        ok = self.L.rustedscithe_aot_eval_residual(_p(args), args.shape[0], _p(out), out.shape[0])
        if ok != 1:
            raise _lib.RstB200Error(ok, "rustedscithe_aot_eval_residual returned 0")
        return out

    def jacobian_structure(self):
        kl, ku = C.c_int(), C.c_int()
        nnz = self.L.rst_b200_jacobian_structure(self.handle, None, None, None, None, C.byref(kl), C.byref(ku))
        rows, cols, doff, dpos = (np.zeros(nnz, dtype=np.int64) for _ in range(4))
        self.L.rst_b200_jacobian_structure(self.handle, _p(rows, _lib.c_i64_p), _p(cols, _lib.c_i64_p),
                                           _p(doff, _lib.c_i64_p), _p(dpos, _lib.c_i64_p), C.byref(kl), C.byref(ku))
        return rows, cols, doff, dpos, kl.value, ku.value

    def jac_call(self, y_reduced, params=None):
        p = _f64(params) if params is not None else (self._keep["params"] if self.spec.params else np.zeros(0))
        args = np.concatenate([p, _f64(y_reduced)])
        nnz = self.L.rst_b200_jacobian_structure(self.handle, None, None, None, None, None, None)
        out = np.zeros(nnz)
        self.L.rst_b200_aot_bind(self.handle)
        ok = self.L.rustedscithe_aot_eval_jacobian_values(_p(args), args.shape[0], _p(out), out.shape[0])
        if ok != 1:
            raise _lib.RstB200Error(ok, "rustedscithe_aot_eval_jacobian_values returned 0")
        return out

    # -- DirectLinearSolver surface ---------------------------------------------------------
    def factor(self):
        self._ck(self.L.rst_b200_factor(self.handle))

    def solve_in_place(self, rhs):
        x = _f64(rhs).copy()
        self._ck(self.L.rst_b200_solve_in_place(self.handle, _p(x), x.shape[0]))
        return x

    def solve_multiple_in_place(self, rhs, nrhs, ldb):
        x = _f64(rhs).copy()
        self._ck(self.L.rst_b200_solve_multiple_in_place(self.handle, _p(x), nrhs, ldb))
        return x

    def set_linear_solver_config(self, policy="Auto", fallback="ToFaerSparse", iterative_refinement_steps=0):
        """LinearSolverConfig { policy, fallback, iterative_refinement_steps } (solver_policy.rs:24-89)."""
        self._ck(self.L.rst_b200_set_linear_solver_config(self.handle, _lib.LINEAR_SOLVER_POLICIES[policy],
                                                            _lib.FALLBACK_POLICIES[fallback], int(iterative_refinement_steps)))

    def get_linear_solver_config(self):
        p, f, r = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.rst_b200_get_linear_solver_config(self.handle, C.byref(p), C.byref(f), C.byref(r)))
        inv_p = {v: k for k, v in _lib.LINEAR_SOLVER_POLICIES.items()}
        inv_f = {v: k for k, v in _lib.FALLBACK_POLICIES.items()}
        return inv_p[p.value], inv_f[f.value], r.value

    def solve_in_place_with_refinement(self, rhs, max_iters=2, tol=1e-14):
        """LapackStyleBandedLuFaithful::solve_in_place_with_refinement (lapack_style_banded.rs:1227-1298); the matvec is
        the block Jacobian resident on the device.  Returns (x, accepted_steps, final relative residual)."""
        x = _f64(rhs).copy()
        acc, rr = C.c_int(0), C.c_double(0.0)
        self._ck(self.L.rst_b200_solve_in_place_with_refinement(self.handle, _p(x), x.shape[0], max_iters, tol,
                                                                  C.byref(acc), C.byref(rr)))
        return x, acc.value, rr.value

    # -- device-resident primitives ------------------------------------------------------------
    def eval_residual(self, which=0):
        self._ck(self.L.rst_b200_eval_residual(self.handle, which))

    def eval_jacobian(self, which=0):
        self._ck(self.L.rst_b200_eval_jacobian(self.handle, which))

    def solve(self):
        self._ck(self.L.rst_b200_solve(self.handle))

    def factor_solve(self):
        """factor_from + solve_in_place in one call (the reference's solve_sys, BVP_traits.rs:875-899); factors stay cached."""
        self._ck(self.L.rst_b200_factor_solve(self.handle))

    def reduce(self, which=0):
        sc = _lib.Scalars()
        self._ck(self.L.rst_b200_reduce(self.handle, which, C.byref(sc)))
        return sc

    def checkpoint(self, restore=False):
        self._ck(self.L.rst_b200_checkpoint(self.handle, 1 if restore else 0))

    def sync(self):
        self._ck(self.L.rst_b200_sync(self.handle))

    def launch_count(self):
        return int(self.L.rst_b200_launch_count(self.handle))

    def p2p_handle(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's NVLink mailbox."""
        buf = C.create_string_buffer(64)
        self._ck(self.L.rst_b200_p2p_get_handle(self.handle, buf, 64))
        return buf.raw

    def p2p_open(self, handles):
        """handles: list of every rank's p2p_handle() in rank order."""
        blob = b"".join(handles)
        self._ck(self.L.rst_b200_p2p_open(self.handle, blob, 64))

    def set_allgather(self, fn):
        """fn(d_send_ptr, d_recv_ptr, doubles_per_rank, stream_ptr) -> 0 on success."""
        def tramp(ctx, send, recv, count, stream):
            try:
                return int(fn(send, recv, count, stream) or 0)
            except Exception:  # never let an exception cross the C ABI
                import traceback
                traceback.print_exc()
                return 1
        self._allgather_cb = _lib.ALLGATHER_FN(tramp)
        self._ck(self.L.rst_b200_set_allgather(self.handle, self._allgather_cb, None))


class BlockTridiagonalLu:
    """GPU drop-in for BlockTridiagonalLuConsistent (src/somelinalg/banded/block_tridiagonal_lu_consistent.rs:28,
    :98-247): ``new(n_blocks, block_size)`` -> ``factor_from(lower, diag, upper)`` -> ``solve_in_place(rhs)``."""

    def __init__(self, n_blocks, block_size, device=0, stream=None):
        self.L = _lib.load()
        self.n_blocks, self.block_size = int(n_blocks), int(block_size)
        self.handle = C.c_void_p()
        _lib.check(self.L.rst_b200_btd_create(self.n_blocks, self.block_size, device, stream, C.byref(self.handle)))

    def n(self):
        return int(self.L.rst_b200_btd_n(self.handle))

    def factor_from(self, lower, diag, upper):
        nb, bs = self.n_blocks, self.block_size
        diag = _f64(diag).reshape(nb, bs, bs)
        lower = _f64(lower).reshape(max(nb - 1, 0), bs, bs) if nb > 1 else np.zeros((1, bs, bs))
        upper = _f64(upper).reshape(max(nb - 1, 0), bs, bs) if nb > 1 else np.zeros((1, bs, bs))
        _lib.check(self.L.rst_b200_btd_factor_from(self.handle, _p(lower), _p(diag), _p(upper)))

    def solve_in_place(self, rhs):
        x = _f64(rhs).copy()
        _lib.check(self.L.rst_b200_btd_solve_in_place(self.handle, _p(x), x.shape[0]))
        return x

    def solve_multiple_in_place(self, rhs, nrhs, ldb):
        x = _f64(rhs).copy()
        _lib.check(self.L.rst_b200_btd_solve_multiple_in_place(self.handle, _p(x), nrhs, ldb))
        return x

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.L.rst_b200_btd_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
