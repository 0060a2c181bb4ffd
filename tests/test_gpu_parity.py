"""GPU parity tests: the CUDA path (through the C ABI, librst_b200.so) against the CPU oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: residual / Jacobian entries 1e-12 relative, each linear-solve
dy 1e-10 relative to the reference banded LU, converged solutions 1e-9 with the same Newton iteration count."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from rustedscithe_b200 import build as rbuild
from rustedscithe_b200 import problems
from rustedscithe_b200.solver import NRBVP, DampedSolverOptions

pytestmark = pytest.mark.gpu

TOL_FJ = 1e-12      # per-entry residual / Jacobian, relative
TOL_DY = 1e-10      # linear solve, relative to the oracle's banded LU
TOL_SOL = 1e-9      # converged solution


@pytest.fixture(scope="module", autouse=True)
def _built():
    rbuild.build()


def _oracle(key, n_steps, scheme="forward", params=None, **kw):
    spec = problems.get(key)
    return orc.OracleBVP(key, n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end, scheme=scheme,
                         params=params if params is not None else (spec.param_values or None), **kw)


def _gpu(key, n_steps, scheme="forward", slab=0, guess=None, **kw):
    spec = problems.get(key)
    opts = DampedSolverOptions.from_spec(spec, scheme=scheme, slab=slab)
    return NRBVP(key, guess, n_steps, opts, **kw)


def _rel(a, b, floor=1.0):
    return np.max(np.abs(a - b) / np.maximum(floor, np.abs(b)))


CASES = [("linear2", 1000), ("damp1", 128), ("damp1p", 77), ("combustion6", 300), ("flame12", 257), ("coupled8", 100),
         ("coupled64", 40)]


@pytest.mark.parametrize("key,n_steps", CASES)
@pytest.mark.parametrize("scheme", ["forward", "trapezoid"])
def test_residual_and_jacobian_values_match_oracle(key, n_steps, scheme):
    """a1/a2 through the reference's own AOT ABI (host pointers, args = [params..., y_reduced...])."""
    rng = np.random.default_rng(11)
    o = _oracle(key, n_steps, scheme)
    g = _gpu(key, n_steps, scheme)
    y = 0.35 + 0.6 * rng.random(o.n)
    r_gpu, r_ref = g.fun_call(y), o.residual(y)
    # residual rows are differences of O(1) terms: relative to max(1, |ref|)
    assert _rel(r_gpu, r_ref) < TOL_FJ
    rows, cols, doff, dpos, kl, ku = g.jacobian_structure()
    orows, ocols, odoff, odpos, okl, oku = o.structure()
    assert np.array_equal(rows, orows) and np.array_equal(cols, ocols)
    assert np.array_equal(doff, odoff) and np.array_equal(dpos, odpos) and (kl, ku) == (okl, oku)
    v_gpu, v_ref = g.jac_call(y), o.jacobian_values(y)
    assert np.max(np.abs(v_gpu - v_ref) / np.maximum(np.abs(v_ref), 1e-300 + 1e-12 * np.max(np.abs(v_ref)))) < 50 * TOL_FJ
    assert _rel(v_gpu, v_ref) < TOL_FJ
    g.close()


def test_aot_abi_error_behaviour():
    """codegen_c_aot_library.rs:382-384: 0 on NULL pointers or length mismatch."""
    import ctypes as C
    g = _gpu("linear2", 16)
    L = g.L
    L.rst_b200_aot_bind(g.handle)
    args, out = np.zeros(32), np.zeros(32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert L.rustedscithe_aot_eval_residual(p(args), 32, p(out), 32) == 1
    assert L.rustedscithe_aot_eval_residual(p(args), 31, p(out), 32) == 0
    assert L.rustedscithe_aot_eval_residual(p(args), 32, p(out), 31) == 0
    assert L.rustedscithe_aot_eval_residual(None, 32, p(out), 32) == 0
    assert L.rustedscithe_aot_eval_jacobian_values(p(args), 32, p(out), 5) == 0
    g.close()


def test_nonuniform_mesh_and_parameters():
    n = 90
    rng = np.random.default_rng(5)
    mesh = np.concatenate([[0.0], np.sort(rng.random(n - 1)), [1.0]])
    spec = problems.get("damp1p")
    o = orc.OracleBVP("damp1p", n, spec.bc_list(), mesh=mesh, params=[1.3, -0.4])
    g = NRBVP("damp1p", None, n, mesh=mesh, params=[1.3, -0.4])
    y = 0.2 + rng.random(o.n)
    assert _rel(g.fun_call(y), o.residual(y)) < TOL_FJ
    assert _rel(g.jac_call(y), o.jacobian_values(y)) < TOL_FJ
    # parameter re-binding through the args prefix (generated_solver_handoff.rs:4017)
    o2 = orc.OracleBVP("damp1p", n, spec.bc_list(), mesh=mesh, params=[0.1, 2.0])
    assert _rel(g.fun_call(y, params=[0.1, 2.0]), o2.residual(y)) < TOL_FJ
    g.close()


SOLVE_CASES = [
    ("linear2", 1000, 0), ("linear2", 37, 4), ("linear2", 2, 0), ("linear2", 3, 2), ("damp1", 128, 8),
    ("combustion6", 1000, 0), ("combustion6", 333, 5), ("flame12", 500, 0), ("flame12", 2000, 32),
    ("flame12", 65, 64), ("coupled8", 200, 3), ("coupled64", 60, 0), ("coupled64", 37, 5),
]


@pytest.mark.parametrize("key,n_steps,slab", SOLVE_CASES)
@pytest.mark.parametrize("scheme", ["forward", "trapezoid"])
def test_linear_solve_matches_faithful_banded_lu(key, n_steps, slab, scheme):
    """a3-a5/a7: DirectLinearSolver::solve_in_place drop-in vs BandedAssembly -> Banded -> AB -> DGBTF2 -> DGBTRS."""
    rng = np.random.default_rng(23)
    spec = problems.get(key)
    o = _oracle(key, n_steps, scheme)
    g = _gpu(key, n_steps, scheme, slab=slab)
    y = spec.guess(n_steps) * (1.0 + 0.01 * rng.standard_normal(o.n))
    vals = o.jacobian_values(y)
    rhs = o.residual(y)
    dy_ref = o.solve_banded(vals, rhs)
    g.set_state(y)
    g.eval_jacobian()
    g.factor()
    dy = g.solve_in_place(rhs)
    err = np.linalg.norm(dy - dy_ref) / np.linalg.norm(dy_ref)
    assert err < TOL_DY, err
    # multi-RHS, column-major with ldb > n (solver_traits.rs:9-14)
    ldb = o.n + 3
    buf = np.zeros(2 * ldb)
    rhs2 = np.cos(np.arange(o.n))
    buf[:o.n], buf[ldb:ldb + o.n] = rhs, rhs2
    out = g.solve_multiple_in_place(buf, 2, ldb)
    assert np.linalg.norm(out[:o.n] - dy_ref) / np.linalg.norm(dy_ref) < TOL_DY
    ref2 = o.solve_banded(vals, rhs2)
    assert np.linalg.norm(out[ldb:ldb + o.n] - ref2) / np.linalg.norm(ref2) < TOL_DY
    g.close()


@pytest.mark.parametrize("key,n_steps,slab", [("flame12", 500, 0), ("coupled8", 200, 3), ("combustion6", 333, 5)])
@pytest.mark.parametrize("scheme", ["forward", "trapezoid"])
def test_block_per_slab_kernels_on_narrow_blocks(key, n_steps, slab, scheme, monkeypatch):
    """The run-time-nv block-per-slab kernels (csrc/big.cuh, used for 16 < nv <= 64) forced onto narrow problems: same
    factor layout, same answers as the banded LU."""
    monkeypatch.setenv("RST_B200_FORCE_BIG", "1")
    rng = np.random.default_rng(29)
    spec = problems.get(key)
    o = _oracle(key, n_steps, scheme)
    g = _gpu(key, n_steps, scheme, slab=slab)
    y = spec.guess(n_steps) * (1.0 + 0.01 * rng.standard_normal(o.n))
    vals, rhs = o.jacobian_values(y), o.residual(y)
    dy_ref = o.solve_banded(vals, rhs)
    g.set_state(y)
    g.eval_jacobian()
    g.factor()
    dy = g.solve_in_place(rhs)
    assert np.linalg.norm(dy - dy_ref) / np.linalg.norm(dy_ref) < TOL_DY
    g.close()


@pytest.mark.parametrize("env", ["RST_B200_BIG_V1", "RST_B200_BIG_W32"])
def test_wide_block_kernel_generations_agree(env, monkeypatch):
    """Generation 1 (shared-memory panel, plain sweeps; also the path of odd nv) and the 32-warp tile variant share the
    factor layout with the default kernels: same answers on the 64-variable case."""
    monkeypatch.setenv(env, "1")
    rng = np.random.default_rng(31)
    key, n_steps = "coupled64", 45
    spec = problems.get(key)
    o = _oracle(key, n_steps)
    g = _gpu(key, n_steps, slab=4)
    y = spec.guess(n_steps) * (1.0 + 0.01 * rng.standard_normal(o.n))
    vals, rhs = o.jacobian_values(y), o.residual(y)
    dy_ref = o.solve_banded(vals, rhs)
    g.set_state(y)
    g.eval_jacobian()
    g.factor()
    dy = g.solve_in_place(rhs)
    assert np.linalg.norm(dy - dy_ref) / np.linalg.norm(dy_ref) < TOL_DY
    g.close()


@pytest.mark.parametrize("key,n_steps", [("flame12", 3000), ("combustion6", 5000), ("coupled64", 300), ("linear2", 20000)])
def test_guarded_refinement_and_linear_solver_config(key, n_steps):
    """solve_in_place_with_refinement (lapack_style_banded.rs:1227-1298) and LinearSolverConfig (solver_policy.rs:24-89):
    refined GPU and refined oracle solutions agree, the relative residual does not grow, policies map as documented."""
    from rustedscithe_b200 import lib as rlib
    rng = np.random.default_rng(41)
    spec = problems.get(key)
    o = _oracle(key, n_steps)
    g = _gpu(key, n_steps)
    y = spec.guess(n_steps) * (1.0 + 0.01 * rng.standard_normal(o.n))
    vals, rhs = o.jacobian_values(y), o.residual(y)
    g.set_state(y)
    g.eval_jacobian()
    g.factor()
    # tol far below rounding: only the guard (the residual must decrease) ends the loop
    x_ref, _ = o.solve_banded_with_refinement(vals, rhs, 3, 1e-30)
    x, accepted, rr = g.solve_in_place_with_refinement(rhs, 3, 1e-30)
    assert 0 <= accepted <= 3 and rr < 1e-12
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < TOL_DY
    # max_iters = 0 is the plain solve
    x0, acc0, _ = g.solve_in_place_with_refinement(rhs, 0, 1e-30)
    # the config routes solve_in_place through the refinement (tol 1e-12, linear_solver.rs:84-101)
    assert g.get_linear_solver_config() == ("Auto", "ToFaerSparse", -1)       # -1: library default (one guarded step at size)
    g.set_linear_solver_config("Auto", "ToFaerSparse", 0)                     # the reference's default: never refine
    assert acc0 == 0 and np.array_equal(x0, g.solve_in_place(rhs))            # 0 steps configured: the plain solve
    g.set_linear_solver_config("ForceBanded", "Never", 2)
    assert g.get_linear_solver_config() == ("ForceBanded", "Never", 2)
    x2 = g.solve_in_place(rhs)
    x2_expected, _, _ = g.solve_in_place_with_refinement(rhs, 2, 1e-12)
    assert np.array_equal(x2, x2_expected)
    for pol in ("ForceBlockTridiagonal", "ForceBlockTridiagonalConsistent"):
        g.set_linear_solver_config(pol, "ToFaerSparse", 0)     # same backend: the condensation accepts both matrix classes
        assert np.array_equal(g.solve_in_place(rhs), x0)
    for pol in ("ForceGeneralBandedPartialPivot", "ForceFaerSparse"):
        with pytest.raises(rlib.RstB200Error) as e:
            g.set_linear_solver_config(pol)
        assert e.value.code == rlib.ERR_UNSUPPORTED
    g.close()


def test_solver_error_behaviour():
    from rustedscithe_b200 import lib as rlib
    g = _gpu("linear2", 64)
    with pytest.raises(rlib.RstB200Error) as e:      # BandedError::NotFactorized
        g.solve_in_place(np.zeros(128))
    assert e.value.code == rlib.ERR_NOT_FACTORIZED
    g.set_state(np.zeros(128))
    g.eval_jacobian()
    g.factor()
    with pytest.raises(rlib.RstB200Error) as e:      # BandedError::DimensionMismatch
        g.solve_in_place(np.zeros(127))
    assert e.value.code == rlib.ERR_INVALID
    g.close()


def test_singular_system_reports_zero_pivot():
    """Two conditions on y and none on z for y' = z, z' = 0 is fine; pinning z twice leaves y undetermined."""
    from rustedscithe_b200 import lib as rlib
    g = NRBVP("linear2", None, 32, boundary_conditions={"z": [(0, 1.0), (1, 1.0)]})
    g.set_state(np.zeros(64))
    g.eval_jacobian()
    g.factor()
    with pytest.raises(rlib.RstB200Error) as e:
        g.solve_in_place(np.ones(64))
    assert e.value.code == rlib.ERR_ZERO_PIVOT
    g.close()


NEWTON_CASES = [("linear2", 80), ("linear2", 1000), ("damp1", 64), ("combustion6", 1000), ("flame12", 400),
                ("coupled8", 150), ("combustion6", 50), ("flame12", 33), ("coupled64", 48)]


@pytest.mark.parametrize("key,n_steps", NEWTON_CASES)
def test_newton_matches_oracle_iteration_counts_and_solution(key, n_steps):
    """a8-a11: same counters (iterations / linear systems / Jacobians) and converged solution within 1e-9."""
    spec = problems.get(key)
    o = _oracle(key, n_steps)
    lo, hi = spec.bounds_per_var()
    ref = o.newton(spec.guess(n_steps), o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var()),
                   abs_tol=spec.abs_tol, max_iterations=spec.max_iterations, max_jac=spec.max_jac,
                   max_damp_iter=spec.max_damp_iter, damp_factor=spec.damp_factor)
    g = _gpu(key, n_steps, guess=spec.guess(n_steps))
    sol = g.try_solve()
    st = g.get_statistics()
    assert st["status"] == ref.status
    assert (st["number of iterations"], st["number of solving linear systems"],
            st["number of jacobians recalculations"]) == (ref.iterations, ref.linear_solves, ref.jac_recalcs)
    assert st["factorizations"] == ref.jac_recalcs          # factors are cached, the reference refactors per solve
    if ref.status != 1:                                      # the reference fails on this start: same failure, no result
        assert sol is None
        g.close()
        return
    assert np.max(np.abs(sol - ref.y) / np.maximum(1.0, np.abs(ref.y))) < TOL_SOL
    full = g.get_result()
    assert np.max(np.abs(full - o.full_solution(ref.y)) / np.maximum(1.0, np.abs(o.full_solution(ref.y)))) < TOL_SOL
    g.close()


def test_wide_block_newton_trapezoid_matches_oracle():
    """64-variable case with the trapezoid scheme (both Jacobian blocks A_j and B_j are assembled and condensed)."""
    key, n_steps = "coupled64", 30
    spec = problems.get(key)
    o = _oracle(key, n_steps, "trapezoid")
    lo, hi = spec.bounds_per_var()
    ref = o.newton(spec.guess(n_steps), o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var()),
                   abs_tol=spec.abs_tol, max_iterations=spec.max_iterations, max_jac=spec.max_jac,
                   max_damp_iter=spec.max_damp_iter, damp_factor=spec.damp_factor)
    g = _gpu(key, n_steps, "trapezoid", guess=spec.guess(n_steps))
    sol = g.try_solve()
    st = g.get_statistics()
    assert st["status"] == ref.status
    assert (st["number of iterations"], st["number of solving linear systems"],
            st["number of jacobians recalculations"]) == (ref.iterations, ref.linear_solves, ref.jac_recalcs)
    if ref.status == 1:
        assert np.max(np.abs(sol - ref.y) / np.maximum(1.0, np.abs(ref.y))) < TOL_SOL
    g.close()


@pytest.mark.parametrize("key,n_steps", [("combustion6", 1000), ("coupled8", 150), ("linear2", 1000)])
def test_graph_replay_and_direct_launches_agree(key, n_steps, monkeypatch):
    """The damped step / Jacobian refresh are replayed as CUDA graphs from their second execution on; the same solve with
    RST_B200_NO_GRAPH (kernel-by-kernel launches, stream-ordered) must give identical counters and a bit-identical state."""
    spec = problems.get(key)
    out = []
    for no_graph in (False, True):
        if no_graph:
            monkeypatch.setenv("RST_B200_NO_GRAPH", "1")
        g = _gpu(key, n_steps, guess=spec.guess(n_steps))
        g.try_solve()
        st = g.get_statistics()
        out.append((st["status"], st["number of iterations"], st["number of solving linear systems"],
                    st["number of jacobians recalculations"], g.get_state()))
        # a second solve on the same handle replays the cached graphs from the first iteration on
        g.set_state(spec.guess(n_steps))
        g.try_solve()
        st2 = g.get_statistics()
        assert (st2["status"], st2["number of iterations"], st2["number of solving linear systems"]) == out[-1][:3]
        assert np.array_equal(g.get_state(), out[-1][4])
        g.close()
    assert out[0][:4] == out[1][:4]
    assert np.array_equal(out[0][4], out[1][4])


def test_config_c1_closed_form():
    """examples/bvp_damped_aot_guide.rs:84-109 at N = 1000: max |y - x| < 5e-5."""
    spec = problems.get("linear2")
    g = _gpu("linear2", 1000, guess=spec.guess(1000))
    assert g.try_solve() is not None
    full = g.get_result()
    x = np.linspace(0.0, 1.0, 1001)
    assert np.max(np.abs(full[:, 0] - x)) < 5e-5 and np.max(np.abs(full[:, 1] - 1.0)) < 5e-5
    g.close()


def test_bound_violation_status_matches_oracle():
    """status -3 path (BVP_utils_damped.rs:39-61, NR_Damp_solver_damped.rs:1898-1905)."""
    spec = problems.get("linear2")
    n = 50
    o = _oracle("linear2", n)
    y0 = spec.guess(n)
    y0[1::2] = 1.2          # y components sit on the upper bound and the first step points outward
    lo, hi = spec.bounds_per_var()
    ref = o.newton(y0, o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var()), abs_tol=spec.abs_tol,
                   max_iterations=spec.max_iterations)
    g = _gpu("linear2", n, guess=y0)
    g.try_solve()
    st = g.get_statistics()
    assert st["status"] == ref.status
    assert st["number of iterations"] == ref.iterations
    g.close()


@pytest.mark.parametrize("key,n_steps,scale", [("coupled8", 300, 0.8), ("damp1", 400, -4.0), ("coupled8", 80, 0.5),
                                               ("damp1", 400, 3.5), ("combustion6", 2000, 0.8)])
def test_rejected_damping_trials_follow_the_reference(key, n_steps, scale):
    """Starts that force REJECTED damping trials (more than two linear solves in an iteration,
    NR_Damp_solver_damped.rs:1927-1976: every trial re-uses the undamped step with a smaller lambda) and failure
    statuses: counters, status and -- when it converges -- the solution must follow the oracle."""
    spec = problems.get(key)
    o = _oracle(key, n_steps)
    y0 = np.full(o.n, scale)
    lo, hi = spec.bounds_per_var()
    ref = o.newton(y0, o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var()), abs_tol=spec.abs_tol,
                   max_iterations=spec.max_iterations, max_jac=spec.max_jac, max_damp_iter=spec.max_damp_iter,
                   damp_factor=spec.damp_factor)
    g = _gpu(key, n_steps, guess=y0)
    sol = g.try_solve()
    st = g.get_statistics()
    assert st["status"] == ref.status
    assert (st["number of iterations"], st["number of solving linear systems"],
            st["number of jacobians recalculations"]) == (ref.iterations, ref.linear_solves, ref.jac_recalcs)
    if ref.status == 1:
        assert np.max(np.abs(sol - ref.y) / np.maximum(1.0, np.abs(ref.y))) < TOL_SOL
    g.close()


# ---- BASELINE.json full sizes: size-independent properties ------------------------------------------
def test_config_c2_million_points_closed_form_and_idempotence():
    """C2: 2-variable BVP at 10^6 points.  The discrete solution of y'=z, z'=0, y(0)=0, z(1)=1 is y = x, z = 1
    on any mesh; a second solve from the converged state must converge in one iteration without moving."""
    n = 1_000_000
    spec = problems.get("linear2")
    g = _gpu("linear2", n, guess=spec.guess(n))
    sol = g.try_solve()
    assert sol is not None
    full = g.get_result()
    x = np.arange(n + 1) / n
    assert np.max(np.abs(full[:, 0] - x)) < 1e-9 and np.max(np.abs(full[:, 1] - 1.0)) < 1e-9
    g2 = _gpu("linear2", n, guess=sol)
    sol2 = g2.try_solve()
    assert g2.get_statistics()["number of iterations"] == 1
    assert np.max(np.abs(sol2 - sol)) < 1e-9
    g.close(); g2.close()


def test_config_c3_million_points_newton_converges_and_residual_vanishes():
    """C3: flame12 at 10^6 points.  Newton must converge in the same number of iterations as the oracle does on a
    coarse mesh (mesh-independent convergence), the converged residual must be at rounding level, and the linear
    solve must satisfy J dy = F (checked through the damped loop: a Newton step from the solution is ~0)."""
    n = 1_000_000
    spec = problems.get("flame12")
    o = _oracle("flame12", 2000)
    lo, hi = spec.bounds_per_var()
    ref = o.newton(spec.guess(2000), o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var()),
                   abs_tol=spec.abs_tol, max_iterations=spec.max_iterations, max_jac=spec.max_jac,
                   max_damp_iter=spec.max_damp_iter, damp_factor=spec.damp_factor)
    g = _gpu("flame12", n, guess=spec.guess(n))
    sol = g.try_solve()
    st = g.get_statistics()
    assert sol is not None and st["status"] == 1
    # mesh-independent convergence: within one iteration of the coarse-mesh oracle count (5 at N = 2000)
    assert abs(st["number of iterations"] - ref.iterations) <= 1
    r = g.fun_call(sol)
    assert np.max(np.abs(r)) < 1e-8   # convergence is on the step norm; per-row residuals are at rounding level
    # coarse-vs-fine agreement of the profiles at shared nodes (first-order scheme: O(h) apart)
    fine = g.get_result()[:: n // 2000]
    coarse = o.full_solution(ref.y)
    assert np.max(np.abs(fine - coarse) / np.maximum(1.0, np.abs(coarse))) < 5e-2
    g.close()


# ---- multi-rank path, emulated in one process through the phase-split ABI --------------------------------
@pytest.mark.parametrize("key,n_steps,world", [("linear2", 101, 2), ("flame12", 300, 3), ("combustion6", 64, 4),
                                                ("coupled64", 40, 2)])
def test_emulated_mesh_slab_ranks_match_single_rank(key, n_steps, world):
    import ctypes as C
    spec = problems.get(key)
    rng = np.random.default_rng(1)
    nv = spec.nv
    y = spec.guess(n_steps) * (1.0 + 0.01 * rng.standard_normal(nv * n_steps))
    single = _gpu(key, n_steps)
    single.set_state(y)
    single.eval_jacobian()
    single.factor()
    single.solve()
    sc1 = single.reduce(0)
    ranks = [NRBVP(key, None, n_steps, DampedSolverOptions.from_spec(spec, rank=r, world=world)) for r in range(world)]
    L = single.L

    def gather(which_send, which_recv, count):
        for dst in ranks:
            for r, src in enumerate(ranks):
                nbytes = count * 8
                a = src.device_ptr(which_send)
                b = dst.device_ptr(which_recv) + r * nbytes
                _memcpy_d2d(b, a, nbytes)

    for rk in ranks:
        rk.set_state(y)
        rk.eval_jacobian()
        assert L.rst_b200_factor_local(rk.handle) == 0
        rk.sync()
    gather(6, 8, 2 * nv * nv)
    for rk in ranks:
        assert L.rst_b200_factor_reduced(rk.handle) == 0
        assert L.rst_b200_solve_local_forward(rk.handle) == 0
        rk.sync()
    gather(7, 9, nv)
    for rk in ranks:
        assert L.rst_b200_solve_reduced_backward(rk.handle) == 0
        assert L.rst_b200_reduce_local(rk.handle, 0) == 0
        rk.sync()
    gather(10, 11, 8)
    from rustedscithe_b200 import lib as rlib
    dy_single = _read(single.device_ptr(4), (n_steps + 1) * nv)
    for rk in ranks:
        b, e = C.c_int64(), C.c_int64()
        L.rst_b200_dist_range(rk.handle, C.byref(b), C.byref(e))
        dy = _read(rk.device_ptr(4), (e.value - b.value + 1) * nv)
        ref = dy_single[b.value * nv:(e.value + 1) * nv]
        assert np.linalg.norm(dy - ref) <= 1e-10 * np.linalg.norm(dy_single)
        sc = rlib.Scalars()
        assert L.rst_b200_reduce_finish(rk.handle, C.byref(sc)) == 0
        assert abs(sc.sumsq_step - sc1.sumsq_step) <= 1e-10 * sc1.sumsq_step
        assert abs(sc.tol_sum - sc1.tol_sum) <= 1e-12 * sc1.tol_sum
        assert sc.fbound == pytest.approx(sc1.fbound, rel=1e-9)
    for rk in ranks:
        rk.close()
    single.close()


def _memcpy_d2d(dst, src, nbytes):
    import ctypes as C
    from rustedscithe_b200 import lib as rlib
    assert rlib.load().rst_b200_memcpy(C.c_void_p(dst), C.c_void_p(src), nbytes, 3) == 0


def _read(ptr, count):
    import ctypes as C
    from rustedscithe_b200 import lib as rlib
    out = np.zeros(count)
    assert rlib.load().rst_b200_memcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), count * 8, 2) == 0
    return out


def test_manifest_and_reentrant_aot_symbols():
    """f3 + boundary: the handle writes the reference's manifest.h; the handle-less AOT symbols can be driven from several
    threads, each bound to its own handle (rst_b200_aot_bind_thread), and give the single-threaded answers."""
    import ctypes as C
    import threading
    from rustedscithe_b200 import build as rbuild
    cases = [("flame12", 700), ("combustion6", 900), ("linear2", 4000)]
    handles, ys, want = [], [], []
    rng = np.random.default_rng(77)
    for key, n in cases:
        spec = problems.get(key)
        g = _gpu(key, n)
        buf = C.create_string_buffer(2048)
        ln = g.L.rst_b200_manifest_h(g.handle, buf, 2048)
        assert ln > 0 and buf.value.decode() == rbuild.manifest_h(spec, n)
        assert g.L.rst_b200_manifest_h(g.handle, buf, 16) == -1
        y = 0.35 + 0.6 * rng.random(spec.nv * n)
        handles.append(g); ys.append(y); want.append((g.fun_call(y), g.jac_call(y)))
    got = [None] * len(cases)

    def worker(i):
        g, y = handles[i], ys[i]
        L = g.L
        L.rst_b200_aot_bind_thread(g.handle)
        n = y.shape[0]
        nnz = want[i][1].shape[0]
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        rs, js = [], []
        for _ in range(8):
            r, j = np.zeros(n), np.zeros(nnz)
            assert L.rustedscithe_aot_eval_residual(p(y), n, p(r), n) == 1
            assert L.rustedscithe_aot_eval_jacobian_values(p(y), n, p(j), nnz) == 1
            rs.append(r); js.append(j)
        got[i] = (rs, js)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(cases))]
    for t in threads: t.start()
    for t in threads: t.join()
    for i in range(len(cases)):
        assert got[i] is not None
        for r, j in zip(*got[i]):
            assert np.array_equal(r, want[i][0]) and np.array_equal(j, want[i][1])
    for g in handles:
        g.close()


@pytest.mark.parametrize("key,n_steps", [("flame12", 600), ("linear2", 3000), ("combustion6", 500), ("coupled64", 40)])
def test_nan_in_the_state_is_reported_not_propagated(key, n_steps):
    """NRBVP::step scans F and the step for NaN and panics (NR_Damp_solver_damped.rs:1821-1826, :1839-1844); the library
    returns RST_B200_ERR_NAN."""
    from rustedscithe_b200 import lib as rlib
    spec = problems.get(key)
    y0 = spec.guess(n_steps).copy()
    y0[y0.shape[0] // 2] = np.nan
    g = _gpu(key, n_steps, guess=y0)
    with pytest.raises(rlib.RstB200Error) as e:
        g.try_solve()
    assert e.value.code == rlib.ERR_NAN
    # the handle stays usable: a clean state solves normally afterwards
    g.set_state(spec.guess(n_steps))
    assert g.try_solve() is not None
    g.close()


@pytest.mark.gpu
def test_solve_pipeline_two_lanes_matches_single_handle():
    """pipeline.SolvePipeline: independent solves spread over two handles / host threads give, bit for bit, what one handle
    gives for the same guesses (different guesses per solve so that a mixed-up buffer would show)."""
    from rustedscithe_b200.pipeline import SolvePipeline
    spec = problems.get("flame12")
    n = 4000
    base = np.asarray(spec.guess(n), dtype=np.float64).reshape(-1)
    rng = np.random.default_rng(5)
    guesses = [base * (1.0 - 1e-4 * rng.random(base.shape)) for _ in range(6)]
    one = NRBVP("flame12", None, n, DampedSolverOptions.from_spec(spec))
    want = []
    for g in guesses:
        one.set_state(g)
        want.append(one.try_solve())
    iters_one = one.get_statistics()["number of iterations"]
    one.close()
    pipe = SolvePipeline("flame12", n, DampedSolverOptions.from_spec(spec), lanes=2)
    sols, stats = pipe.solve_many(guesses)
    pipe.close()
    assert all(st[0] == 1 for st in stats)
    assert stats[-1][1] == iters_one
    for w, s in zip(want, sols):
        assert w is not None and s is not None
        np.testing.assert_array_equal(np.asarray(w).reshape(-1), s)


_CARRY_SCRIPT = r"""
import json, sys
import numpy as np
from rustedscithe_b200 import problems
from rustedscithe_b200.solver import NRBVP, DampedSolverOptions
out = {}
# the last three start from constant states that force REJECTED damping trials (the accepted trial is then a later one)
for key, n, scale in (("flame12", 3000, None), ("linear2", 5000, None), ("combustion6", 2000, None), ("coupled64", 300, None),
                      ("coupled8", 300, 0.8), ("damp1", 400, -4.0), ("combustion6", 2000, 0.8)):
    spec = problems.get(key)
    sv = NRBVP(key, spec.guess(n), n, DampedSolverOptions.from_spec(spec))
    start = np.asarray(spec.guess(n), dtype=np.float64).reshape(-1) if scale is None else np.full(sv.n, scale)
    for rep in range(3):      # direct launches, graph capture, graph replay
        sv.set_state(start)
        sol = sv.try_solve()
    st = sv.get_statistics()
    key = key if scale is None else f"{key}@{scale}"
    out[key] = {"sol": None if sol is None else np.asarray(sol).tobytes().hex(), "it": st["number of iterations"],
                "ls": st["number of solving linear systems"], "jac": st["number of jacobians recalculations"],
                "reused": st["solves reused"]}
    sv.close()
json.dump(out, open(sys.argv[1], "w"))
"""


@pytest.mark.gpu
def test_carried_undamped_step_is_bit_identical_to_recomputing_it(tmp_path):
    """The undamped step taken over from the accepted first trial (newton_trial_carry_kernel) against recomputing it
    (RST_B200_NO_CARRY=1, the reference's sequence): same bits in the converged state, same counters; and the carry is
    actually taken on the multi-iteration problems."""
    import json
    import subprocess
    import sys
    res = {}
    for tag, env_extra in (("carry", {}), ("recompute", {"RST_B200_NO_CARRY": "1"}), ("no_rhs_column", {"RST_B200_NO_RHS_COLUMN": "1"})):
        env = {k: v for k, v in os.environ.items() if k not in ("RST_B200_NO_CARRY", "RST_B200_NO_RHS_COLUMN")}
        env.update(env_extra)
        path = tmp_path / f"{tag}.json"
        subprocess.run([sys.executable, "-c", _CARRY_SCRIPT, str(path)], check=True, env=env,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        res[tag] = json.load(open(path))
    for key in res["carry"]:
        a, b = res["carry"][key], res["recompute"][key]
        assert a["sol"] == b["sol"] and (a["sol"] is not None or "@" in key), key
        assert (a["it"], a["ls"], a["jac"]) == (b["it"], b["ls"], b["jac"]), key
        assert b["reused"] == 0
        print(f"{key}: iterations {a['it']}, linear solves {a['ls']} of which carried {a['reused']}")
    assert res["carry"]["flame12"]["reused"] >= 1
    assert any(res["carry"][k]["reused"] >= 1 and res["carry"][k]["ls"] > 2 * res["carry"][k]["it"] for k in res["carry"] if "@" in k), \
        "no start with rejected trials carried a step"
    # nv = 12: the right-hand side eliminated inside the factorisation (abd_factor4_kernel<.., RHS>) against the forward sweeps
    for key in res["carry"]:
        a, b = res["carry"][key], res["no_rhs_column"][key]
        assert (a["it"], a["ls"], a["jac"]) == (b["it"], b["ls"], b["jac"]), key
        if a["sol"] is None or b["sol"] is None:      # a start that ends in a failure status: both must fail
            assert a["sol"] is None and b["sol"] is None, key
            continue
        ya = np.frombuffer(bytes.fromhex(a["sol"]), dtype=np.float64)
        yb = np.frombuffer(bytes.fromhex(b["sol"]), dtype=np.float64)
        assert np.max(np.abs(ya - yb) / np.maximum(np.abs(yb), 1.0)) < 1e-13, key
        print(f"rhs column vs forward sweeps, {key}: bit-identical = {a['sol'] == b['sol']}")


@pytest.mark.gpu
@pytest.mark.parametrize("key,n", [("flame12", 5000), ("flame12", 37), ("combustion6", 3000), ("linear2", 4000), ("coupled64", 200)])
def test_factor_solve_equals_factor_then_solve(key, n):
    """rst_b200_factor_solve (the reference's solve_sys: factor_from + solve_in_place in one call; nv = 12 carries the right-hand
    side through the factorisation instead of running the forward sweeps) against rst_b200_factor + rst_b200_solve: same bits,
    and the cached factors serve a later plain solve."""
    spec = problems.get(key)
    sv = NRBVP(key, spec.guess(n), n, DampedSolverOptions.from_spec(spec))
    import ctypes as C
    from rustedscithe_b200 import lib as rlib
    m = (n + 1) * spec.nv

    def dY():
        out = np.empty(m)
        rlib.check(sv.L.rst_b200_memcpy(C.c_void_p(out.ctypes.data), C.c_void_p(sv.device_ptr(4)), m * 8, 2))
        return out

    sv.eval_jacobian(0)
    sv.factor()
    sv.solve()
    want = dY()
    sv.eval_jacobian(0)
    sv.factor_solve()
    got = dY()
    np.testing.assert_array_equal(got, want)
    sv.solve()                       # plain solve (forward sweeps and all) on the factors factor_solve left behind, same F
    np.testing.assert_array_equal(dY(), want)
    assert np.all(np.isfinite(want)) and np.abs(want).max() > 0
    sv.close()
