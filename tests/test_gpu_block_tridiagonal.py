"""GPU tests of the BlockTridiagonal drop-in (row a6): rst_b200_btd_* vs the reference's known answers
(block_tridiagonal_lu_consistent.rs:1191-1341), the oracle's restatement and dense LAPACK."""
import numpy as np
import pytest

from oracle import oracle as orc
from rustedscithe_b200 import build as rbuild
from rustedscithe_b200 import lib as rlib
from rustedscithe_b200.solver import BlockTridiagonalLu
from tests.test_oracle_block_tridiagonal import dense_from_blocks, pivoted_diag_system, small_block_system

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    rbuild.build()


def _gpu_solve(lower, diag, upper, rhs):
    nb, bs = diag.shape[0], diag.shape[1]
    lu = BlockTridiagonalLu(nb, bs)
    lu.factor_from(lower, diag, upper)
    x = lu.solve_in_place(rhs)
    lu.close()
    return x


def test_small_system_known_solution():  # :1201-1213
    lower, diag, upper = small_block_system()
    x_true = np.array([1.0, 2.0, 3.0, 4.0])
    x = _gpu_solve(lower, diag, upper, dense_from_blocks(lower, diag, upper) @ x_true)
    assert np.max(np.abs(x - x_true)) < 1e-10


def test_pivoted_diagonal_block():  # :1233-1265
    lower, diag, upper = pivoted_diag_system()
    x_true = np.array([1.0, -1.0, 2.0, 0.5])
    x = _gpu_solve(lower, diag, upper, dense_from_blocks(lower, diag, upper) @ x_true)
    assert np.max(np.abs(x - x_true)) < 1e-10


def test_bvp_node_blocks_that_the_reference_rejects_are_solved():
    """the reduced-ordering Jacobian of a BVP with split boundary conditions, cut into
    nv x nv node blocks (BandedAssembly::to_block_tridiagonal, banded_assembly.rs:170-229), has structurally singular
    diagonal blocks.  The reference's in-block-pivoting solver answers ZeroPivot (:1267-1286, reproduced by the oracle);
    pivoting across block rows solves it.  Checked against the faithful banded LU."""
    from rustedscithe_b200 import problems
    spec = problems.get("combustion6")
    nb, bs = 24, 6
    o = orc.OracleBVP("combustion6", nb, spec.bc_list())
    y = np.full(o.n, 0.99)
    rows, cols = o.structure()[:2]
    vals = o.jacobian_values(y)
    a = np.zeros((o.n, o.n))
    a[rows, cols] = vals
    blk = lambda i, j: a[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
    diag = np.array([blk(k, k) for k in range(nb)])
    lower = np.array([blk(k + 1, k) for k in range(nb - 1)])
    upper = np.array([blk(k, k + 1) for k in range(nb - 1)])
    assert np.array_equal(dense_from_blocks(lower, diag, upper), a)       # J really is block tridiagonal
    assert np.count_nonzero(np.diag(a) == 0.0) > nb                       # ... whose scalar diagonal is mostly zero
    with pytest.raises(ZeroDivisionError):                                # the reference algorithm rejects it
        orc.btd_factor(lower, diag, upper)
    rhs = o.residual(y)
    x = _gpu_solve(lower, diag, upper, rhs)
    x_ref = o.solve_banded(vals, rhs)
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-10


@pytest.mark.parametrize("nb,bs", [(1, 2), (2, 2), (3, 3), (4, 1), (40, 4), (41, 6), (257, 2), (1000, 8), (5000, 3)])
@pytest.mark.parametrize("seed", [42, 777])
def test_random_diagonally_dominant_matches_oracle_and_dense(nb, bs, seed):
    rng = np.random.default_rng(seed)
    lower = rng.uniform(-0.05, 0.05, (max(nb - 1, 0), bs, bs))
    upper = rng.uniform(-0.05, 0.05, (max(nb - 1, 0), bs, bs))
    diag = rng.uniform(-0.2, 0.2, (nb, bs, bs))
    for k in range(nb):
        for i in range(bs):
            diag[k, i, i] = np.sum(np.abs(diag[k, i])) + rng.uniform(1.0, 2.0)
    rhs = rng.uniform(-1, 1, nb * bs)
    x = _gpu_solve(lower, diag, upper, rhs)
    if nb > 1:
        x_ref = orc.btd_solve(orc.btd_factor(lower, diag, upper), rhs)
    else:
        x_ref = np.linalg.solve(diag[0], rhs)
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-10


def test_multiple_rhs_and_errors():  # :220-247, error.rs:4-39
    lower, diag, upper = small_block_system()
    a = dense_from_blocks(lower, diag, upper)
    lu = BlockTridiagonalLu(2, 2)
    with pytest.raises(rlib.RstB200Error) as e:
        lu.solve_in_place(np.zeros(4))
    assert e.value.code == rlib.ERR_NOT_FACTORIZED
    lu.factor_from(lower, diag, upper)
    x0, x1 = np.array([1.0, 2.0, 3.0, 4.0]), np.array([-0.5, 0.25, 0.0, 2.0])
    buf = np.zeros(2 * 6)
    buf[:4], buf[6:10] = a @ x0, a @ x1
    out = lu.solve_multiple_in_place(buf, 2, 6)
    assert np.max(np.abs(out[:4] - x0)) < 1e-10 and np.max(np.abs(out[6:10] - x1)) < 1e-10
    with pytest.raises(rlib.RstB200Error) as e:
        lu.solve_in_place(np.zeros(5))
    assert e.value.code == rlib.ERR_INVALID
    with pytest.raises(rlib.RstB200Error) as e:
        lu.solve_multiple_in_place(buf, 2, 3)
    assert e.value.code == rlib.ERR_INVALID
    lu.close()
    with pytest.raises(rlib.RstB200Error) as e:
        BlockTridiagonalLu(4, 5)
    assert e.value.code == rlib.ERR_UNSUPPORTED
    singular = np.zeros((2, 2, 2))
    lu = BlockTridiagonalLu(2, 2)
    with pytest.raises(rlib.RstB200Error) as e:
        lu.factor_from(np.zeros((1, 2, 2)), singular, np.zeros((1, 2, 2)))
        lu.solve_in_place(np.ones(4))
    assert e.value.code == rlib.ERR_ZERO_PIVOT
    lu.close()
