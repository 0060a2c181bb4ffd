"""Pins the oracle's banded LU (oracle/rst_oracle.c: orc_dgb_factor / orc_dgb_solve) against the
reference's own known-answer tests in src/somelinalg/banded/lapack_style_banded.rs:1466-2123 and
against scipy's LAPACK dgbtrf/dgbtrs (the routine the reference mirrors, :3-17)."""
import math

import numpy as np
import pytest
from scipy.linalg import lapack

from oracle import oracle as orc


def _solve(a, kl, ku, b):
    n = a.shape[0]
    banded = orc.banded_from_dense(a, kl, ku)
    ab, ipiv = orc.dgb_factor(n, kl, ku, orc.banded_to_ab(n, kl, ku, banded))
    return orc.dgb_solve(n, kl, ku, ab, ipiv, b), ab, ipiv


def make_tridiag_5():  # lapack_style_banded.rs:1469-1479
    a = np.zeros((5, 5))
    for i in range(5):
        a[i, i] = 4.0
    for i in range(4):
        a[i, i + 1] = 1.0
        a[i + 1, i] = 2.0
    return a


def make_pivot_required_small():  # :1481-1491
    a = np.zeros((3, 3))
    a[0, 1] = 2.0
    a[1, 0] = 1.0
    a[1, 1] = 3.0
    a[1, 2] = 1.0
    a[2, 1] = 1.0
    a[2, 2] = 2.0
    return a


def make_pivot_heavy_12():  # :1654-1690, :1895-1920
    n, kl, ku = 12, 2, 2
    a = np.zeros((n, n))
    for j in range(n):
        col_sum = 0.0
        for i in range(max(0, j - ku), min(n, j + kl + 1)):
            if i == j:
                continue
            v = 0.1 * math.sin(float(3 * i + 7 * j + 1))
            a[i, j] = v
            col_sum += abs(v)
        a[j, j] = col_sum + 1.0
    a[0, 0], a[4, 4], a[8, 8] = 1e-10, 1e-9, 1e-11
    a[1, 0] = a[5, 4] = a[9, 8] = 1.0
    return a


def make_tail_pathology_8():  # :1747-1786
    a = np.zeros((8, 8))
    for (i, j), v in {
        (0, 0): 2.0, (1, 0): 1.0, (0, 1): -0.5, (1, 1): 2.5, (2, 1): 0.75, (1, 2): -0.3, (2, 2): 2.2, (3, 2): 0.8,
        (2, 3): -0.2, (3, 3): 2.1, (4, 3): 0.7, (4, 4): 0.0, (5, 4): 1.0, (6, 4): -7.0, (4, 5): 1.0, (5, 5): 0.0,
        (6, 5): -1.0, (4, 6): 0.0, (5, 6): 1.0, (6, 6): 1.5, (7, 6): -1.0, (5, 7): 0.0, (6, 7): 1.0, (7, 7): 1.2,
    }.items():
        a[i, j] = v
    return a


def test_workspace_dimensions():  # :1493-1499
    n, kl, ku = 10, 2, 3
    ab = orc.banded_to_ab(n, kl, ku, np.zeros((kl + ku + 1) * n))
    assert ab.size == (2 * kl + ku + 1) * n == 80


def test_load_preserves_band_entries():  # :1501-1512
    a = make_tridiag_5()
    ab = orc.banded_to_ab(5, 1, 1, orc.banded_from_dense(a, 1, 1)).reshape(5, 4).T  # ab[row, col]
    kv = 2
    for i in range(5):
        for j in range(max(0, i - 1), min(5, i + 2)):
            assert ab[kv + i - j, j] == a[i, j]


def test_tridiagonal_known_solution():  # :1541-1560
    a = make_tridiag_5()
    x_true = np.array([1.0, -1.0, 2.0, 0.5, -0.25])
    x, _, _ = _solve(a, 1, 1, a @ x_true)
    assert np.max(np.abs(x - x_true)) < 1e-10


def test_pivot_required_small_known_solution():  # :1562-1587
    a = make_pivot_required_small()
    x_true = np.array([1.0, 2.0, -1.0])
    x, _, ipiv = _solve(a, 1, 1, a @ x_true)
    assert np.max(np.abs(x - x_true)) < 1e-10
    assert ipiv[0] == 1  # the zero diagonal forces the interchange


def test_multiple_rhs():  # :1589-1617
    a = make_tridiag_5()
    n = 5
    ab, ipiv = orc.dgb_factor(n, 1, 1, orc.banded_to_ab(n, 1, 1, orc.banded_from_dense(a, 1, 1)))
    for x_true in (np.array([1.0, -1.0, 2.0, 0.5, -0.25]), np.array([0.1, 0.2, 0.3, 0.4, 0.5])):
        x = orc.dgb_solve(n, 1, 1, ab, ipiv, a @ x_true)
        assert np.max(np.abs(x - x_true)) < 1e-10


def test_pivot_heavy_n12():  # :1654-1744, :1932-1982
    a = make_pivot_heavy_12()
    x_true = 0.2 * np.cos(np.arange(12, dtype=float))
    b = a @ x_true
    x, _, _ = _solve(a, 2, 2, b)
    assert np.max(np.abs(x - x_true)) < 1e-9
    rr = np.max(np.abs(a @ x - b)) / max(np.max(np.abs(b)), 1.0)
    assert rr < 1e-10


def test_tail_pathology():  # :1747-1841
    a = make_tail_pathology_8()
    x_true = np.array([1.0, -0.5, 0.25, 2.0, -1.0, 0.75, -0.2, 1.5])
    b = a @ x_true
    x, _, _ = _solve(a, 2, 2, b)
    assert np.max(np.abs(x - x_true)) < 1e-8
    assert np.max(np.abs(a @ x - b)) / max(np.max(np.abs(b)), 1.0) < 1e-10


def test_multi_rhs_pivot_heavy_6():  # :1843-1893
    n = 6
    a = np.zeros((n, n))
    a[0, 0] = 1e-10
    a[1, 0] = 1.0
    a[0, 1], a[1, 1], a[2, 1] = 2.0, 3.0, 1.0
    for k in range(2, 5):
        a[k - 1, k], a[k, k], a[k + 1, k] = 1.0, 2.0, 1.0
    a[4, 5], a[5, 5] = 1.0, 2.0
    for x_true in (np.array([1.0, 2.0, -1.0, 0.5, 0.25, -0.75]), np.array([0.1, -0.2, 0.3, -0.4, 0.5, -0.6])):
        x, _, _ = _solve(a, 1, 1, a @ x_true)
        assert np.max(np.abs(x - x_true)) < 1e-9


def test_zero_pivot_is_reported():  # threshold 1e-14, :79, :398-403
    a = np.zeros((3, 3))
    a[0, 1] = a[1, 2] = 1.0
    with pytest.raises(ZeroDivisionError):
        _solve(a, 1, 1, np.ones(3))


@pytest.mark.parametrize("seed", [42, 123, 777, 12345, 424242])
@pytest.mark.parametrize("shape", [(200, 6, 2), (97, 17, 6), (64, 1, 1), (50, 3, 0), (120, 70, 20)])
def test_matches_lapack_dgbtrf(seed, shape):
    """Same pivots and (to rounding) same factors/solution as LAPACK on random banded systems,
    including zero diagonals that force interchanges."""
    n, kl, ku = shape
    rng = np.random.default_rng(seed)
    a = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            a[i, j] = rng.uniform(-0.2, 0.2)
        a[i, i] = rng.uniform(1.0, 2.0)
    if kl > 0 and ku > 0:  # zero diagonals with strong off-diagonal neighbours: forces row interchanges
        z = np.arange(0, n - 1, 3)
        a[z, z] = 0.0
        a[z + 1, z] = 1.5
        a[z, z + 1] = 1.5
    b = rng.uniform(-1, 1, n)
    x, ab, ipiv = _solve(a, kl, ku, b)
    ldab = 2 * kl + ku + 1
    ab_l = np.zeros((ldab, n))
    for i in range(n):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            ab_l[kl + ku + i - j, j] = a[i, j]
    lu, piv, info = lapack.dgbtrf(ab_l, kl, ku)
    assert info == 0
    assert np.array_equal(piv, ipiv)
    xl, info = lapack.dgbtrs(lu, kl, ku, b.copy(), piv)
    assert info == 0
    assert np.max(np.abs(x - xl)) <= 1e-9 * max(1.0, np.max(np.abs(xl)))
    mine = ab.reshape(n, ldab).T
    assert np.allclose(mine, lu, rtol=1e-9, atol=1e-12)


def test_guarded_refinement_restatement_never_increases_the_residual():
    """solve_in_place_with_refinement (lapack_style_banded.rs:1227-1298) as restated in oracle.OracleBVP: on a BVP system
    the accepted steps can only lower max|b - A x| / max(max|b|, 1); with tol above the first residual nothing is accepted
    (the reference's early exit, :1264-1266)."""
    from rustedscithe_b200 import problems
    spec = problems.get("combustion6")
    n = 400
    o = orc.OracleBVP("combustion6", n, spec.bc_list(), t0=spec.t0, t_end=spec.t_end)
    rng = np.random.default_rng(3)
    y = spec.guess(n) * (1.0 + 0.01 * rng.standard_normal(o.n))
    vals, rhs = o.jacobian_values(y), o.residual(y)
    banded = o.banded(vals)
    _, _, _, _, kl, ku = o.structure()

    def rr(x):
        return np.max(np.abs(rhs - orc.banded_matvec(o.n, kl, ku, banded, x))) / max(np.max(np.abs(rhs)), 1.0)

    x0 = o.solve_banded(vals, rhs)
    x, accepted = o.solve_banded_with_refinement(vals, rhs, 3, 1e-30)
    assert 0 <= accepted <= 3
    assert rr(x) <= rr(x0)
    x_loose, acc_loose = o.solve_banded_with_refinement(vals, rhs, 3, 1.0)
    assert acc_loose == 0 and np.array_equal(x_loose, x0)
