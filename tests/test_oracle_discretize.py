"""Pins the oracle's discretiser (layout, residual, Jacobian stream, band layouts) against the
reference's unit tests (src/numerical/BVP_Damp/numeric_discretization.rs:373-497,
src/somelinalg/banded/banded_assembly.rs:414-520) and against golden vectors produced from the
reference's checked-in generated fixtures (tests/golden/make_fixture_vectors.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from rustedscithe_b200 import problems

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "bvp_damp1_fixture_vectors.json")


def _bvp(key, n_steps, **kw):
    spec = problems.get(key)
    return orc.OracleBVP(key, n_steps, spec.bc_list(), t0=spec.t0, t_end=spec.t_end, **kw), spec


def test_layout_sizes():  # numeric_discretization.rs:377-401
    b, _ = _bvp("linear2", 4, mesh=[0.0, 0.25, 0.5, 0.75, 1.0])
    assert b.unknown_pos.shape == (8,)
    assert b.bc_pos.tolist() == [0, 9]           # y_0 and z_4
    assert b.unknown_pos.tolist() == [1, 2, 3, 4, 5, 6, 7, 8]
    assert b.per_unknown(np.array([1e-6, 2e-6])).shape == (8,)


def test_two_bcs_on_same_variable_accepted():  # :439-474
    b = orc.OracleBVP("linear2", 4, [(0, 0, 0.0), (0, 1, 1.0)])
    names = {int(p) for p in b.unknown_pos}
    assert 1 in names and 4 * 2 + 1 in names     # z_0 and z_4 stay unknown
    assert b.unknown_pos.shape == (8,)


def test_underconstrained_rejected():  # :476-496
    with pytest.raises(ValueError, match="expects exactly 2 total boundary conditions"):
        orc.OracleBVP("linear2", 4, [(0, 0, 0.0)])


def test_duplicate_bc_rejected():  # :69-76
    with pytest.raises(ValueError, match="duplicate"):
        orc.OracleBVP("linear2", 4, [(0, 0, 0.0), (0, 0, 1.0)])


def test_uniform_mesh_contract():  # solver_common.rs:107-119
    b, _ = _bvp("linear2", 4)
    assert b.mesh.tolist() == [0.0, 0.25, 0.5, 0.75, 1.0]


@pytest.mark.parametrize("n_steps", [8, 32, 128])
def test_residual_and_jacobian_match_reference_generated_fixture(n_steps):
    """BVP_Damp1 fixture: z(0)=1, y(N)=1, h=1 (create_mesh default, symbolic_functions_BVP.rs:5299-5308)."""
    cases = json.load(open(GOLDEN))[str(n_steps)]
    b = orc.OracleBVP("damp1", n_steps, problems.get("damp1").bc_list(), t0=0.0, t_end=float(n_steps))
    assert np.all(np.diff(b.mesh) == 1.0)
    for case in cases:
        args = np.array(case["args"])
        r = b.residual(args)
        ref = np.array(case["residual"])
        assert np.max(np.abs(r - ref) / np.maximum(1.0, np.abs(ref))) < 1e-12
        vals = b.jacobian_values(args)
        rows, cols, _, _, _, _ = b.structure()
        # at h = 1 the reference's simplifier folds  -1 - h*(-1)  (z-row, z-column) to a symbolic 0 and
        # prunes it (View/jacobian.rs:111-128); drop those exact zeros before comparing the streams.
        keep = vals != 0.0
        ref_vals = np.array(case["sparse_values"])
        assert keep.sum() == ref_vals.shape[0]
        assert np.max(np.abs(vals[keep] - ref_vals) / np.maximum(1.0, np.abs(ref_vals))) < 1e-12


@pytest.mark.parametrize("key,n_steps,scheme", [
    ("damp1", 16, "forward"), ("damp1", 16, "trapezoid"), ("combustion6", 12, "forward"),
    ("flame12", 9, "forward"), ("flame12", 9, "trapezoid"), ("coupled8", 7, "forward"), ("damp1p", 10, "forward"),
])
def test_jacobian_is_derivative_of_residual(key, n_steps, scheme):
    spec = problems.get(key)
    b = orc.OracleBVP(key, n_steps, spec.bc_list(), scheme=scheme, params=spec.param_values or None)
    rng = np.random.default_rng(7)
    y = 0.5 + 0.4 * rng.random(b.n)
    rows, cols, doff, dpos, kl, ku = b.structure()
    vals = b.jacobian_values(y)
    dense = np.zeros((b.n, b.n))
    dense[rows, cols] = vals
    fd = np.zeros_like(dense)
    for c in range(b.n):
        e = np.zeros(b.n)
        hh = 1e-6 * max(1.0, abs(y[c]))
        e[c] = hh
        fd[:, c] = (b.residual(y + e) - b.residual(y - e)) / (2 * hh)
    scale = np.maximum(1.0, np.abs(fd))
    assert np.max(np.abs(dense - fd) / scale) < 1e-6
    assert np.array_equal(doff, cols - rows)
    assert np.array_equal(dpos, np.where(doff >= 0, rows, cols))
    assert kl == int(np.max(rows - cols)) and ku == int(np.max(cols - rows))
    # (row, col)-sorted stream, symbolic_functions_BVP.rs:1343
    order = np.lexsort((cols, rows))
    assert np.array_equal(order, np.arange(rows.shape[0]))


def test_bandwidths_match_survey_worked_values():
    """SURVEY.md A.4 worked values: 2-var 1L/1R -> (1,1); combustion 3L/3R -> (8,3); 12-var 6L/6R -> (17,6)."""
    assert _bvp("linear2", 10)[0].structure()[4:] == (1, 1)
    assert _bvp("combustion6", 10)[0].structure()[4:] == (8, 3)
    assert _bvp("flame12", 10)[0].structure()[4:] == (17, 6)
    b, _ = _bvp("combustion6", 10)
    vals = b.jacobian_values(np.full(b.n, 0.99))
    rows, cols = b.structure()[:2]
    diag_present = set(rows[(rows == cols) & (vals != 0.0)].tolist())
    assert len(diag_present) <= 1                # 59/60 structurally zero scalar diagonals


def test_band_layouts_roundtrip():  # banded_assembly.rs:417-520, storage.rs:96-103, lapack_style_banded.rs:258-275
    b, _ = _bvp("combustion6", 6)
    y = np.linspace(0.3, 0.9, b.n)
    rows, cols, doff, dpos, kl, ku = b.structure()
    vals = b.jacobian_values(y)
    dense = np.zeros((b.n, b.n))
    dense[rows, cols] = vals
    banded = b.banded(vals).reshape(kl + ku + 1, b.n)
    ab = orc.banded_to_ab(b.n, kl, ku, banded.reshape(-1)).reshape(b.n, 2 * kl + ku + 1).T
    for i in range(b.n):
        for j in range(max(0, i - kl), min(b.n, i + ku + 1)):
            assert banded[ku + i - j, j] == dense[i, j]
            assert ab[kl + ku + i - j, j] == dense[i, j]
    assert np.all(ab[:kl] == 0.0)                # fill rows start empty
    x = np.cos(np.arange(b.n, dtype=float))
    assert np.allclose(orc.banded_matvec(b.n, kl, ku, banded.reshape(-1), x), dense @ x, rtol=1e-13, atol=1e-13)


def test_hand_written_problems_match_finite_differences():
    for key in ["linear2", "damp1", "damp1p", "combustion6", "flame12", "coupled8", "coupled64"]:
        spec = problems.get(key)
        nv = spec.nv
        y = 0.4 + 0.05 * np.arange(nv) / nv
        p = spec.param_values or None
        f, fy = orc.problem_rhs(key, 0.3, y, p)
        fd = np.zeros((nv, nv))
        for c in range(nv):
            e = np.zeros(nv)
            e[c] = 1e-6
            fd[:, c] = (orc.problem_rhs(key, 0.3, y + e, p)[0] - orc.problem_rhs(key, 0.3, y - e, p)[0]) / 2e-6
        scale = np.maximum(np.abs(fd), 1e-3 * np.max(np.abs(fd)) + 1e-300)
        assert np.max(np.abs(fy - fd) / scale) < 1e-5, key
        assert np.array_equal(orc.problem_pattern(key) != 0, fy != 0.0), key
