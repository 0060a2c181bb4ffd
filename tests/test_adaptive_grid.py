"""Adaptive grid refinement hand-off.

CPU part: the numpy oracle (oracle/adaptive_grid.py) pinned by the reference's own unit tests
(src/numerical/BVP_Damp/adaptive_grid_basic.rs:992-1241: fixtures, refinement invariants, uniform mid-points, flat identity,
SciPy threshold known answers).  GPU part: rst_b200_new_grid / rst_b200_create_refined against the oracle, bit for bit on
marks, mesh and interpolated guess, and the adaptive solve loop against the oracle loop."""
import numpy as np
import pytest

from oracle import adaptive_grid as ag
from oracle import oracle as orc
from rustedscithe_b200 import problems


# ---- fixtures of the reference's tests (adaptive_grid_basic.rs:996-1019) ----------------------------------
def boundary_layer_fixture(n_points):
    x = np.array([i / (n_points - 1) for i in range(n_points)], dtype=np.float64)
    y = np.stack([1.2 + np.tanh(18.0 * (x - 0.38)), 0.8 + np.sin(7.0 * x) + 0.15 * x,
                  0.2 + np.exp(-70.0 * (x - 0.72) ** 2)])
    return y, x


def flat_fixture(n_points):
    x = np.array([i / (n_points - 1) for i in range(n_points)], dtype=np.float64)
    return np.full((3, n_points), 2.5), x


def assert_refinement_invariants(old_y, old_x, refined):
    """adaptive_grid_basic.rs:1042-1115."""
    new_grid, new_guess, inserted = refined[:3]
    assert old_y.shape[1] == len(old_x)
    assert new_guess.shape[0] == old_y.shape[0]
    assert new_guess.shape[1] == len(new_grid)
    assert inserted == len(new_grid) - len(old_x)
    assert np.all(np.diff(new_grid) > 0), "refined grid must be strictly increasing"
    eps = 1e-12
    assert abs(new_grid[0] - old_x[0]) <= eps and abs(new_grid[-1] - old_x[-1]) <= eps
    new_idx = 0
    for old_idx in range(len(old_x)):
        while new_idx < len(new_grid) and new_grid[new_idx] < old_x[old_idx] - eps:
            new_idx += 1
        assert new_idx < len(new_grid) and abs(new_grid[new_idx] - old_x[old_idx]) <= eps
        assert np.max(np.abs(old_y[:, old_idx] - new_guess[:, new_idx])) <= 1e-12
    old_left = 0
    for new_col, x_new in enumerate(new_grid):
        while old_left + 1 < len(old_x) and x_new > old_x[old_left + 1] + eps:
            old_left += 1
        if abs(x_new - old_x[old_left]) <= eps:
            continue
        if old_left + 1 < len(old_x) and abs(x_new - old_x[old_left + 1]) <= eps:
            continue
        assert old_left + 1 < len(old_x)
        x_l, x_r = old_x[old_left], old_x[old_left + 1]
        assert x_l < x_new < x_r
        theta = (x_new - x_l) / (x_r - x_l)
        expected = old_y[:, old_left] + (old_y[:, old_left + 1] - old_y[:, old_left]) * theta
        assert np.max(np.abs(new_guess[:, new_col] - expected)) <= 1e-12


CASES = [(ag.EASY, (0.08,)), (ag.PEARSON, (0.08, 1.5)), (ag.GRCAR_SMOOKE, (0.08, 0.08, 1.5))]


@pytest.mark.parametrize("method,params", CASES)
def test_oracle_marking_on_coarse_boundary_layer_mesh(method, params):
    """adaptive_grid_parallel_marking_matches_sequential_on_coarse_mesh (:1144-1162): invariants on the 16-point fixture."""
    y, x = boundary_layer_fixture(16)
    refined = ag.new_grid(method, y, x, params=params)
    assert_refinement_invariants(y, x, refined)
    assert refined[2] > 0


def test_oracle_pearson_parallel_merge_equals_sequential_hashmap_variant():
    """The reference asserts its sequential (:339-467) and parallel (:469-540) Pearson variants agree on this fixture; the
    sequential variant interleaves the merge with the scan, restated here independently."""
    y, x = boundary_layer_fixture(16)
    d, C = 0.08, 1.5
    mark = [0] * len(x)
    for j in range(y.shape[0]):
        row = y[j]
        delta = d * (row.max() - row.min())
        for i in range(len(x)):
            if len(x) - 1 > i > 0:
                tau = abs(row[i] - row[i - 1])
                if tau > delta and abs(row[i]) > 1e-4 and abs(row[i + 1]) > 1e-4:
                    N = max(int(tau / delta), 1)
                    if N > mark[i]:
                        mark[i - 1] = N
    ag._buffer_pass(mark, x, C)
    assert ag.pearson_grid_refinement(y, x, d, C)[3].tolist() == mark


def test_oracle_uniform_refinement_preserves_boundaries_and_midpoints():
    """adaptive_grid_uniform_refinement_preserves_boundaries_and_midpoints (:1164-1181)."""
    y, x = boundary_layer_fixture(12)
    refined = ag.refine_all_grid(y, x)
    assert_refinement_invariants(y, x, refined)
    assert refined[2] == len(x) - 1
    for i in range(len(x) - 1):
        assert abs(refined[0][2 * i + 1] - 0.5 * (x[i] + x[i + 1])) <= 1e-14


def test_oracle_no_refinement_is_identity_for_flat_solution():
    """adaptive_grid_no_refinement_is_identity_for_flat_solution (:1183-1203)."""
    y, x = flat_fixture(14)
    residual = np.full(len(x), 1e-8)
    cases = [ag.easy_grid_refinement(y, x, 0.08), ag.pearson_grid_refinement(y, x, 0.08, 1.5),
             ag.grcar_smooke_grid_refinement(y, x, 0.08, 0.08, 1.5), ag.scipy_grid_refinement(y, x, 1e-4, residual)]
    for refined in cases:
        assert refined[2] == 0
        assert_refinement_invariants(y, x, refined)
        assert len(refined[0]) == len(x) and refined[1].shape == y.shape


def test_oracle_scipy_residual_thresholds_insert_expected_nodes():
    """adaptive_grid_scipy_residual_thresholds_insert_expected_nodes (:1205-1241)."""
    y, x = boundary_layer_fixture(5)
    tol = 1e-3
    residual = np.array([2.0 * tol, 100.0 * tol, 0.1 * tol, 0.1 * tol, 0.0])
    refined = ag.scipy_grid_refinement(y, x, tol, residual)
    assert refined[2] == 3
    assert_refinement_invariants(y, x, refined)
    for expected in (0.125, 1.0 / 3.0, 5.0 / 12.0):
        assert np.any(np.abs(refined[0] - expected) <= 1e-14)


@pytest.mark.parametrize("method,params", [(ag.DOUBLE_POINTS, ())] + CASES + [(ag.PEARSON, (1e-3, 1.4)), (ag.GRCAR_SMOOKE, (1e-3, 1e-2, 1.4))])
def test_oracle_invariants_on_a_nonuniform_mesh(method, params):
    """The refinement invariants of the reference's tests (:1042-1115) on a graded mesh (the hand-off after a first
    refinement always sees one), incl. the production parameter sets of the user guide (delta 1e-3, ratio 1.4)."""
    rng = np.random.default_rng(12)
    w = rng.uniform(0.2, 3.0, 47)
    x = np.concatenate([[0.0], np.cumsum(w) / np.sum(w)])
    x[-1] = 1.0
    y = np.stack([1.2 + np.tanh(18.0 * (x - 0.38)), 0.8 + np.sin(7.0 * x) + 0.15 * x, 0.2 + np.exp(-70.0 * (x - 0.72) ** 2)])
    refined = ag.new_grid(method, y, x, params=params)
    assert_refinement_invariants(y, x, refined)
    assert refined[2] == int(np.sum(refined[3][:-1]))          # inserted nodes = sum of the marks that can be honoured


def test_oracle_rust_cast_semantics():
    assert ag._as_i32(3.99) == 3 and ag._as_i32(float("inf")) == 2147483647 and ag._as_i32(float("nan")) == 0


# ---- GPU ----------------------------------------------------------------------------------------------------
_BUILT = []


def _gpu_solver(key, n_steps, mesh=None, adaptive=None):
    from rustedscithe_b200 import build as rbuild
    from rustedscithe_b200.solver import NRBVP, DampedSolverOptions
    if not _BUILT:
        rbuild.build()
        _BUILT.append(True)
    spec = problems.get(key)
    opts = DampedSolverOptions.from_spec(spec)
    opts.strategy_params.adaptive = adaptive
    return NRBVP(key, None, n_steps, opts, mesh=mesh)


def _profile_state(spec, x, seed):
    """A smooth multi-scale profile per variable (boundary layers, a bump, an oscillation) as the full solution."""
    rng = np.random.default_rng(seed)
    cols = []
    for v in range(spec.nv):
        a, c, w = rng.uniform(0.3, 1.5), rng.uniform(0.2, 0.8), rng.uniform(5.0, 40.0)
        kind = v % 3
        if kind == 0:
            cols.append(a + np.tanh(w * (x - c)))
        elif kind == 1:
            cols.append(a + np.sin(0.3 * w * x) + 0.15 * x)
        else:
            cols.append(0.2 * a + np.exp(-3.0 * w * (x - c) ** 2))
    return np.stack(cols, axis=1)     # (n_points, nv)


GPU_CASES = [("linear2", 15, False), ("linear2", 400, True), ("combustion6", 63, False), ("flame12", 257, True),
             ("coupled64", 40, False)]
METHODS = [("DoublePoints", ()), ("Easy", (0.08,)), ("Pearson", (0.08, 1.5)), ("Pearson", (1e-3, 1.4)),
           ("GrcarSmooke", (0.08, 0.08, 1.5)), ("GrcarSmooke", (1e-3, 1e-2, 1.4))]


@pytest.mark.gpu
@pytest.mark.parametrize("key,n_steps,nonuniform", GPU_CASES)
@pytest.mark.parametrize("method,params", METHODS)
def test_gpu_new_grid_is_bit_identical_to_oracle(key, n_steps, nonuniform, method, params):
    spec = problems.get(key)
    rng = np.random.default_rng(5)
    mesh = None
    if nonuniform:
        w = rng.uniform(0.2, 3.0, n_steps)
        mesh = spec.t0 + (spec.t_end - spec.t0) * np.concatenate([[0.0], np.cumsum(w) / np.sum(w)])
        mesh[-1] = spec.t_end
    g = _gpu_solver(key, n_steps, mesh=mesh)
    o = orc.OracleBVP(key, n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end, mesh=mesh)
    full0 = _profile_state(spec, o.mesh, 17)
    g.set_state(full0.reshape(-1)[o.unknown_pos])
    full = g.get_result()                       # boundary slots pinned by the library: the matrix the reference refines
    new_mesh, new_y, n_added, marks = g.new_grid(method, params)
    ref = ag.new_grid(getattr(ag, {"DoublePoints": "DOUBLE_POINTS", "Easy": "EASY", "Pearson": "PEARSON",
                                   "GrcarSmooke": "GRCAR_SMOOKE"}[method]), full.T, o.mesh, params=params)
    assert n_added == ref[2]
    assert np.array_equal(marks, ref[3])
    assert np.array_equal(new_mesh, ref[0])       # bit-exact: same operations in the same order
    assert np.array_equal(new_y, ref[1].T)
    assert_refinement_invariants(full.T, o.mesh, (new_mesh, new_y.T, n_added))
    g.close()


@pytest.mark.gpu
def test_gpu_sci_refinement_uses_the_residual_of_the_state():
    key, n_steps = "linear2", 40
    spec = problems.get(key)
    g = _gpu_solver(key, n_steps)
    o = orc.OracleBVP(key, n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end)
    y = spec.guess(n_steps).copy()
    y[10:14] += 3e-5          # residual rows of a few intervals leave the tolerance band
    y[20] += 1e-2
    g.set_state(y)
    res = o.residual(y)
    tol = g.options.abs_tolerance
    res_head = res.copy()
    # the reference indexes the mesh with the residual ROW: rows >= n_steps must stay below the tolerance or it panics
    if np.any(res_head[n_steps:-1] > tol):
        with pytest.raises(Exception):
            g.new_grid("Sci")
        g.close()
        return
    new_mesh, new_y, n_added, marks = g.new_grid("Sci")
    ref = ag.scipy_grid_refinement(g.get_result().T, o.mesh, tol, res)
    assert n_added == ref[2] and n_added > 0
    assert np.array_equal(marks, ref[3]) and np.array_equal(new_mesh, ref[0]) and np.array_equal(new_y, ref[1].T)
    g.close()


def _oracle_adaptive(key, n_steps, method, params, max_refinements):
    """The reference's adaptive loop (NR_Damp_solver_damped.rs:2102-2113, :2281-2340) on the CPU oracle."""
    spec = problems.get(key)
    lo, hi = spec.bounds_per_var()
    mesh = None
    y = spec.guess(n_steps)
    refinements, nodes_added, counters = 0, [], [0, 0, 0]
    enabled = True
    while True:
        o = orc.OracleBVP(key, n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end, mesh=mesh)
        r = o.newton(y, o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var()), abs_tol=spec.abs_tol,
                     max_iterations=spec.max_iterations, max_jac=spec.max_jac, max_damp_iter=spec.max_damp_iter,
                     damp_factor=spec.damp_factor)
        counters = [counters[0] + r.iterations, counters[1] + r.linear_solves, counters[2] + r.jac_recalcs]
        if r.status != 1:
            return None, o, refinements, nodes_added, counters
        if not enabled:
            return r.y, o, refinements, nodes_added, counters
        full = o.full_solution(r.y).reshape(n_steps + 1, o.nv)
        new_mesh, new_y, added, _ = ag.new_grid(method, full.T, o.mesh, abs_tolerance=spec.abs_tol, params=params)
        refinements += 1
        nodes_added.append(added)
        enabled = added != 0 and not (max_refinements <= refinements)
        if added == 0:
            return r.y, o, refinements, nodes_added, counters
        mesh, n_steps = new_mesh, len(new_mesh) - 1
        o2 = orc.OracleBVP(key, n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end, mesh=mesh)
        y = new_y.T.reshape(-1)[o2.unknown_pos]          # extract_unknown_variables (BVP_utils.rs:560-627)


@pytest.mark.gpu
@pytest.mark.parametrize("key,n_steps,method,max_ref", [
    ("linear2", 20, ("DoublePoints",), 2), ("combustion6", 40, ("Easy", 0.05), 3), ("combustion6", 30, ("Pearson", 0.05, 1.5), 2),
    ("flame12", 24, ("GrcarSmooke", 0.05, 0.05, 1.5), 2), ("damp1", 32, ("Easy", 0.02), 4),
])
def test_gpu_adaptive_solve_matches_the_oracle_loop(key, n_steps, method, max_ref):
    from rustedscithe_b200.solver import AdaptiveGridConfig
    spec = problems.get(key)
    g = _gpu_solver(key, n_steps, adaptive=AdaptiveGridConfig(1, max_ref, method))
    g.set_state(spec.guess(n_steps))
    out = g.try_solve()
    ref_y, o, refinements, nodes_added, counters = _oracle_adaptive(
        key, n_steps, getattr(ag, {"DoublePoints": "DOUBLE_POINTS", "Easy": "EASY", "Pearson": "PEARSON",
                                   "GrcarSmooke": "GRCAR_SMOOKE"}[method[0]]), method[1:], max_ref)
    assert g.grid_refinements == refinements and g.nodes_added == nodes_added
    t = g.total_statistics
    assert [t["number of iterations"], t["number of solving linear systems"],
            t["number of jacobians recalculations"]] == counters
    assert t["number of grid refinements"] == refinements
    assert g.n_steps == o.n_steps
    assert (g.result is None) == (ref_y is None)
    if ref_y is not None:
        assert np.max(np.abs(g.result - ref_y) / np.maximum(1.0, np.abs(ref_y))) < 1e-9
        if nodes_added[-1] != 0:
            assert out is not None            # the loop ended on the refinement cap: the reference returns the result
        else:
            assert out is None                # "no new grid needed" returns Ok(None); the result stays in `result`
    g.close()


@pytest.mark.gpu
def test_gpu_new_grid_error_behaviour():
    """Argument checking of the C ABI: unknown / unsupported method (TwoPoint), missing parameters, fetch before new_grid,
    buffer length mismatches, and the Sci case in which the reference itself panics (residual row index past the mesh)."""
    import ctypes as C
    from rustedscithe_b200 import lib as rlib
    key, n_steps = "combustion6", 20
    spec = problems.get(key)
    g = _gpu_solver(key, n_steps)
    g.set_state(spec.guess(n_steps))
    n_new, n_added = C.c_int64(), C.c_int64()
    for method in (5, -1, 17):                    # 5 would be TwoPoint
        rc = g.L.rst_b200_new_grid(g.handle, method, None, 1e-6, C.byref(n_new), C.byref(n_added))
        assert rc == rlib.ERR_UNSUPPORTED
    assert g.L.rst_b200_new_grid(g.handle, 2, None, 1e-6, C.byref(n_new), C.byref(n_added)) == rlib.ERR_INVALID
    buf = np.zeros(4)
    assert g.L.rst_b200_new_grid_fetch(g.handle, buf.ctypes.data_as(rlib.c_f64_p), 4, None, 0, None, 0) == rlib.ERR_INVALID
    out = C.c_void_p()
    assert g.L.rst_b200_create_refined(g.handle, C.byref(out)) == rlib.ERR_INVALID      # nothing to refine from yet
    g.new_grid("DoublePoints", fetch=False)
    assert g.L.rst_b200_new_grid_fetch(g.handle, buf.ctypes.data_as(rlib.c_f64_p), 4, None, 0, None, 0) == rlib.ERR_INVALID
    # Sci with the flat start: residual rows far beyond the mesh length exceed the tolerance -> the reference panics
    with pytest.raises(rlib.RstB200Error) as e:
        g.new_grid("Sci")
    assert e.value.code == rlib.ERR_INVALID
    g.close()


@pytest.mark.gpu
def test_gpu_refined_handle_is_a_full_solver():
    """After the hand-off the new handle owns the refined mesh: residual / Jacobian / linear solve agree with an oracle
    built on the fetched mesh, and the state is the interpolated solution."""
    key, n_steps = "flame12", 48
    spec = problems.get(key)
    g = _gpu_solver(key, n_steps)
    g.set_state(spec.guess(n_steps))
    assert g.try_solve() is not None
    mesh, y_new, n_added, _ = g.new_grid("GrcarSmooke", (0.05, 0.05, 1.5))
    assert n_added > 0
    g._adopt_refined()
    assert g.n_steps == len(mesh) - 1
    o = orc.OracleBVP(key, g.n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end, mesh=mesh)
    y_red = y_new.reshape(-1)[o.unknown_pos]
    assert np.array_equal(g.get_state(), y_red)
    assert np.array_equal(g.get_result(), y_new)
    r_ref = o.residual(y_red)
    g.eval_jacobian()
    g.factor()
    dy = g.solve_in_place(r_ref)
    dy_ref = o.solve_banded(o.jacobian_values(y_red), r_ref)
    assert np.linalg.norm(dy - dy_ref) / np.linalg.norm(dy_ref) < 1e-10
    g.close()
