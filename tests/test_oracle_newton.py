"""Pins the oracle's damping helpers and damped Newton driver against the reference's unit tests
(src/numerical/BVP_Damp/BVP_utils_damped.rs:98-182), its closed-form integration tests
(examples/bvp_damped_aot_guide.rs:84-109; tests/basic_correctness.rs:22-61) and the published
counters of the combustion-1000 story run (BVP_DAMP_STORY_TESTS.md:721: 5 iterations, 10 linear
systems, 1 Jacobian)."""
import numpy as np
import pytest

from oracle import oracle as orc
from rustedscithe_b200 import problems


def test_cantera2_full_step_inside_bounds():  # :111-118
    assert orc.bound_step_cantera2([0.5, -0.2], [0.1, -0.3], [(0.0, 1.0), (-1.0, 1.0)]) == pytest.approx(1.0, abs=1e-12)


def test_cantera2_clips_upper_and_lower():  # :120-130
    assert orc.bound_step_cantera2([0.8], [0.5], [(0.0, 1.0)]) == pytest.approx(0.4, abs=1e-12)
    assert orc.bound_step_cantera2([0.2], [-0.5], [(0.0, 1.0)]) == pytest.approx(0.4, abs=1e-12)


def test_cantera2_most_restrictive_component():  # :132-139
    assert orc.bound_step_cantera2([0.8, 0.1, 0.5], [0.5, -0.5, 0.1], [(0.0, 1.0)] * 3) == pytest.approx(0.2, abs=1e-12)


def test_cantera2_zero_at_boundary_moving_outward():  # :141-152
    assert orc.bound_step_cantera2([1.0], [0.1], [(0.0, 1.0)]) == pytest.approx(0.0, abs=1e-12)
    assert orc.bound_step_cantera2([0.0], [-0.1], [(0.0, 1.0)]) == pytest.approx(0.0, abs=1e-12)


def test_cantera2_scaled_step_stays_inside():  # :154-169
    x, step = np.array([0.8, 0.2]), np.array([0.5, -0.5])
    fb = orc.bound_step_cantera2(x, step, [(0.0, 1.0)] * 2)
    nxt = x + fb * step
    assert np.all(nxt >= -1e-12) and np.all(nxt <= 1 + 1e-12)


def test_convergence_condition():  # :209-222
    assert orc.convergence_condition([1.0, -2.0], 1e-6, [1e-3, 1e-3]) == pytest.approx(3e-3)
    assert orc.convergence_condition([1.0, -2.0], 1.0, [1e-3, 1e-3]) == 1.0


def _newton(key, n_steps, **kw):
    spec = problems.get(key)
    b = orc.OracleBVP(key, n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end)
    lo, hi = spec.bounds_per_var()
    res = b.newton(spec.guess(n_steps), b.per_unknown(lo), b.per_unknown(hi), b.per_unknown(spec.rtol_per_var()),
                   abs_tol=spec.abs_tol, max_iterations=spec.max_iterations, max_jac=spec.max_jac,
                   max_damp_iter=spec.max_damp_iter, damp_factor=spec.damp_factor, **kw)
    return b, res


@pytest.mark.parametrize("n_steps", [80, 1000])
def test_linear2_converges_to_y_equals_x(n_steps):
    """examples/bvp_damped_aot_guide.rs:84-109: max |y - x| < 5e-5 (config C1 at n_steps = 1000)."""
    b, res = _newton("linear2", n_steps)
    assert res.status == 1
    full = b.full_solution(res.y)
    x = np.linspace(0.0, 1.0, n_steps + 1)
    assert np.max(np.abs(full[:, 0] - x)) < 5e-5
    assert np.max(np.abs(full[:, 1] - 1.0)) < 5e-5
    assert res.iterations <= 4  # the misaligned flat guess (NR_Damp_solver_damped.rs:2034-2040) costs damping steps


def test_combustion_1000_reproduces_published_counters():
    """BVP_DAMP_STORY_TESTS.md:721 (combustion-1000, Banded AOT): iters 5, linsys 10, jac_re 1."""
    b, res = _newton("combustion6", 1000)
    assert res.status == 1
    assert (res.iterations, res.linear_solves, res.jac_recalcs) == (5, 10, 1)
    assert res.factorizations == 10              # refactor on every solve, BVP_traits.rs:875-899
    b2, res2 = _newton("combustion6", 1000, refactor_per_solve=False)
    assert res2.factorizations == 1 and res2.iterations == 5
    assert np.max(np.abs(res.y - res2.y)) == 0.0
    full = b.full_solution(res.y)
    assert full[0, 0] == pytest.approx((1000.0 - 600.0) / 600.0) and full[-1, 1] == 1e-10
    assert np.linalg.norm(b.residual(res.y)) < 1e-3  # convergence is on the step norm, not the residual


def test_flame12_converges():
    b, res = _newton("flame12", 400)
    assert res.status == 1
    assert np.linalg.norm(b.residual(res.y)) < 1e-3  # convergence is on the step norm, not the residual
