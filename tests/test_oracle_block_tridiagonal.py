"""Pins the oracle's BlockTridiagonalLuConsistent restatement (oracle/rst_oracle.c: orc_btd_factor / orc_btd_solve)
against the reference's known-answer tests in
src/somelinalg/banded/block_tridiagonal_lu_consistent.rs:1006-1026, :1076-1112, :1191-1286."""
import numpy as np
import pytest

from oracle import oracle as orc


def dense_from_blocks(lower, diag, upper):
    nb, bs = diag.shape[0], diag.shape[1]
    a = np.zeros((nb * bs, nb * bs))
    for k in range(nb):
        a[k * bs:(k + 1) * bs, k * bs:(k + 1) * bs] = diag[k]
        if k + 1 < nb:
            a[(k + 1) * bs:(k + 2) * bs, k * bs:(k + 1) * bs] = lower[k]
            a[k * bs:(k + 1) * bs, (k + 1) * bs:(k + 2) * bs] = upper[k]
    return a


def small_block_system():  # :1006-1026
    diag = np.array([[[4.0, 1.0], [2.0, 3.0]], [[5.0, 1.0], [1.0, 4.0]]])
    upper = np.array([[[0.5, 0.0], [0.0, 0.5]]])
    lower = np.array([[[0.25, 0.0], [0.0, 0.25]]])
    return lower, diag, upper


def pivoted_diag_system():  # :1233-1250
    diag = np.array([[[0.0, 2.0], [1.0, 3.0]], [[4.0, 0.5], [0.5, 5.0]]])
    upper = np.array([[[0.2, 0.0], [0.0, 0.2]]])
    lower = np.array([[[0.1, 0.0], [0.0, 0.1]]])
    return lower, diag, upper


def combustion_like_blocks():  # :1076-1112
    bs = 6
    d = np.zeros((bs, bs))
    d[0, 0], d[0, 3] = -7.1429, 1.0
    d[1, 0], d[1, 4] = -1.0045, 0.9954916
    d[2, 1], d[2, 5] = -1736.111111111111, 1.0
    d[3, 1] = -1.0008
    d[4, 2] = -1736.111111111111
    d[5, 2] = -1.0008
    up = np.zeros((bs, bs))
    up[3, 0], up[4, 1], up[5, 2] = 0.99925, 1.0, 0.99925
    lo = np.zeros((bs, bs))
    lo[0, 3], lo[1, 4], lo[2, 5] = -0.99925, -1.0, -0.99925
    return lo, d, up


def test_small_system_known_solution():  # :1201-1213
    lower, diag, upper = small_block_system()
    x_true = np.array([1.0, 2.0, 3.0, 4.0])
    rhs = dense_from_blocks(lower, diag, upper) @ x_true
    x = orc.btd_solve(orc.btd_factor(lower, diag, upper), rhs)
    assert np.max(np.abs(x - x_true)) < 1e-10


def test_pivoted_diagonal_block():  # :1233-1265
    lower, diag, upper = pivoted_diag_system()
    x_true = np.array([1.0, -1.0, 2.0, 0.5])
    rhs = dense_from_blocks(lower, diag, upper) @ x_true
    fac = orc.btd_factor(lower, diag, upper)
    assert fac[3][0].tolist() == [1, 1]          # the zero (0,0) entry forces the interchange in block 0
    assert np.max(np.abs(orc.btd_solve(fac, rhs) - x_true)) < 1e-10


def test_combustion_like_singular_node_block_is_rejected():  # :1267-1286
    lo, d, up = combustion_like_blocks()
    with pytest.raises(ZeroDivisionError):
        orc.btd_factor(np.array([lo]), np.array([d, d]), np.array([up]))


@pytest.mark.parametrize("seed", [42, 123, 777])
def test_diagonally_dominant_random_blocks(seed):  # generator shape: block_tridiag_bench_helpers.rs:4-48
    rng = np.random.default_rng(seed)
    nb, bs = 40, 5
    lower = rng.uniform(-0.05, 0.05, (nb - 1, bs, bs))
    upper = rng.uniform(-0.05, 0.05, (nb - 1, bs, bs))
    diag = rng.uniform(-0.2, 0.2, (nb, bs, bs))
    for k in range(nb):
        for i in range(bs):
            diag[k, i, i] = np.sum(np.abs(diag[k, i])) + rng.uniform(1.0, 2.0)
    a = dense_from_blocks(lower, diag, upper)
    x_true = rng.uniform(-1, 1, nb * bs)
    x = orc.btd_solve(orc.btd_factor(lower, diag, upper), a @ x_true)
    assert np.max(np.abs(x - x_true)) < 1e-10
