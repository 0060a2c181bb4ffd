"""Frozen-Jacobian Newton: oracle restatement of NR_Damp_solver_frozen.rs:1035-1148
pinned by the reference's own frozen linear test (:1377-1475, y = x exactly), and GPU parity against the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc
from rustedscithe_b200 import problems

LEFT_BCS = [(0, 0, 0.0), (1, 0, 1.0)]   # y(0) = 0, z(0) = 1  (NR_Damp_solver_frozen.rs:1397-1400)


def _linear_guess(n):
    g = np.zeros(2 * n)
    g[0::2] = np.arange(n) / n
    g[1::2] = 1.0
    return g


@pytest.mark.parametrize("strategy,sp", [("Naive", ()), ("Frozen_naive", ()), ("every_m", (2,)),
                                         ("at_high_morm", (1e-3,)), ("at_low_speed", (0.5,)), ("complex", (2, 1e-2, 0.5))])
def test_oracle_frozen_linear_bvp_is_exact(strategy, sp):
    n = 64
    o = orc.OracleBVP("linear2", n, LEFT_BCS)
    r = o.newton_frozen(_linear_guess(n), strategy, sp, tolerance=1e-8, max_iterations=20)
    assert r.status == 1
    full = o.full_solution(r.y)
    assert np.max(np.abs(full[:, 0] - np.linspace(0, 1, n + 1))) < 1e-12 and np.max(np.abs(full[:, 1] - 1.0)) < 1e-12
    assert o.structure()[5] == 0            # all-left (IVP-like) boundary conditions: no super-diagonals


@pytest.mark.gpu
@pytest.mark.parametrize("key,n_steps,strategy,sp,tol", [
    ("linear2", 64, "Frozen_naive", (), 1e-8), ("linear2", 1000, "Naive", (), 1e-8),
    ("damp1", 80, "every_m", (1,), 1e-8), ("damp1", 80, "Naive", (), 1e-9),
    ("combustion6", 200, "every_m", (2,), 1e-6), ("flame12", 150, "complex", (2, 1e-2, 0.5), 1e-6),
    ("flame12", 150, "at_low_speed", (0.5,), 1e-6), ("combustion6", 200, "at_high_morm", (1e-3,), 1e-6),
    ("coupled64", 36, "every_m", (1,), 1e-6),          # wide-block kernels under the frozen loop
])
def test_gpu_frozen_matches_oracle(key, n_steps, strategy, sp, tol):
    from rustedscithe_b200.solver import NRBVP
    spec = problems.get(key)
    bcs = LEFT_BCS if key == "linear2" else spec.bc_list()
    o = orc.OracleBVP(key, n_steps, bcs)
    y0 = _linear_guess(n_steps) if key == "linear2" else spec.guess(n_steps) * 0 + 0.6
    if key == "coupled64":
        # from 0.6 the undamped iterates of this case pass through a Jacobian with a pivot below the reference's 1e-14
        # ZeroPivot threshold (lapack_style_banded.rs:398-403): the band LU gives up there, the condensation (other pivots,
        # exact-zero test only) does not -- a documented difference, so the wide case starts from the catalogue guess
        y0 = spec.guess(n_steps)
    ref = o.newton_frozen(y0, strategy, sp, tolerance=tol, max_iterations=40)
    bc_dict = {"y": [(0, 0.0)], "z": [(0, 1.0)]} if key == "linear2" else None
    g = NRBVP(key, y0, n_steps, boundary_conditions=bc_dict)
    sol = g.try_solve_frozen(strategy, sp, tolerance=tol, max_iterations=40)
    st = g.get_statistics()
    assert st["status"] == ref.status
    assert (st["number of iterations"], st["number of solving linear systems"],
            st["number of jacobians recalculations"]) == (ref.iterations, ref.linear_solves, ref.jac_recalcs)
    if ref.status == 1:
        assert np.max(np.abs(sol - ref.y) / np.maximum(1.0, np.abs(ref.y))) < 1e-9
    g.close()
