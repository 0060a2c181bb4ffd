"""AOT lifecycle for a problem that is not compiled into the default library (the reference's BuildIfMissing /
RequirePrebuilt contract, generated_solver_handoff.rs:321-331, examples/bvp_damped_aot_guide.rs:119-137)."""
import ctypes as C
import os

import numpy as np
import pytest

from rustedscithe_b200 import build as rbuild
from rustedscithe_b200 import lib as rlib
from rustedscithe_b200 import problems


def _chain3_rhs():
    import sympy as sp
    a, b, c = (sp.Symbol(n, real=True) for n in ("a", "b", "c"))
    return [b, c, sp.Float(0.0)]


# a' = b, b' = c, c' = 0 ; a(0) = 0, b(0) = 0, c(1) = 2  ->  c = 2, b = 2x, a = x^2 (continuous)
CHAIN3 = problems.ProblemSpec(
    key="chain3", names=["a", "b", "c"], rhs=_chain3_rhs,
    bcs={"a": [(0, 0.0)], "b": [(0, 0.0)], "c": [(1, 2.0)]},
    bounds={n: (-10.0, 10.0) for n in "abc"}, rtol={n: 1e-8 for n in "abc"},
    guess=lambda n: np.full(3 * n, 0.1), abs_tol=1e-10,
)


@pytest.fixture(scope="module")
def artifact():
    return rbuild.build_artifact(CHAIN3, "build_if_missing")


def test_artifact_is_built_once_and_registered(artifact):
    assert os.path.basename(artifact).startswith("librst_b200_chain3_")
    L = rlib.load(artifact)
    assert rlib.has_problem(L, "chain3") and not rlib.has_problem(rlib.load(), "chain3")
    nv, npar, h = C.c_int(), C.c_int(), C.c_char_p()
    assert L.rst_b200_problem_info(b"chain3", C.byref(nv), C.byref(npar), C.byref(h)) == 0
    assert nv.value == 3 and h.value.decode() in artifact
    mtime = os.path.getmtime(artifact)
    assert rbuild.build_artifact(CHAIN3, "build_if_missing") == artifact and os.path.getmtime(artifact) == mtime
    assert rbuild.build_artifact(CHAIN3, "require_prebuilt") == artifact


def test_require_prebuilt_fails_when_missing():
    other = problems.ProblemSpec(key="chain3_missing", names=CHAIN3.names, rhs=_chain3_rhs, bcs=CHAIN3.bcs,
                                 bounds=CHAIN3.bounds, rtol=CHAIN3.rtol, guess=CHAIN3.guess)
    with pytest.raises(FileNotFoundError):
        rbuild.build_artifact(other, "require_prebuilt")


@pytest.mark.gpu
def test_custom_problem_solves_to_the_discrete_closed_form(artifact):
    """Odd block size (nv = 3) through the generic kernels.  Forward Euler on the linear chain has the closed form
    c_j = 2, b_j = 2 x_j, a_j = h^2 j (j - 1)."""
    from rustedscithe_b200.solver import NRBVP
    n = 400
    g = NRBVP(CHAIN3, CHAIN3.guess(n), n)
    assert g.try_solve() is not None
    full = g.get_result()
    j = np.arange(n + 1)
    h = 1.0 / n
    assert np.max(np.abs(full[:, 2] - 2.0)) < 1e-11
    assert np.max(np.abs(full[:, 1] - 2.0 * j * h)) < 1e-11
    assert np.max(np.abs(full[:, 0] - h * h * j * (j - 1))) < 1e-11
    g.close()


# ---- wide custom problems: block sizes between the compiled-in ones (odd nv -> generation-1 sweeps, padded register tiles) ----
def _chains_spec(key, n_chains, length):
    """`n_chains` independent integrator chains y_0' = y_1, ..., y_{L-1}' = 0 with y_i(0) = 0 (i < L-1), y_{L-1}(1) = c_b:
    forward Euler gives y_{L-1-m}(x_j) = c_b h^m C(j, m) exactly (repeated summation of a constant)."""
    names = [f"y{b}_{i}" for b in range(n_chains) for i in range(length)]

    def rhs():
        import sympy as sp
        s = [sp.Symbol(n, real=True) for n in names]
        out = []
        for b in range(n_chains):
            for i in range(length):
                out.append(s[b * length + i + 1] if i + 1 < length else sp.Float(0.0))
        return out

    bcs = {}
    for b in range(n_chains):
        for i in range(length):
            bcs[names[b * length + i]] = [(0, 0.0)] if i + 1 < length else [(1, 1.0 + 0.5 * b)]
    return problems.ProblemSpec(key=key, names=names, rhs=rhs, bcs=bcs, bounds={n: (-1e6, 1e6) for n in names},
                                rtol={n: 1e-8 for n in names}, guess=lambda n, nv=len(names): np.full(nv * n, 0.1),
                                abs_tol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("key,n_chains,length,n", [("chains21", 3, 7, 90), ("chains40", 5, 8, 70)])
def test_wide_custom_problems_solve_to_the_discrete_closed_form(key, n_chains, length, n):
    from math import comb
    from rustedscithe_b200.solver import NRBVP, DampedSolverOptions
    spec = _chains_spec(key, n_chains, length)
    rbuild.build_artifact(spec, "build_if_missing")
    g = NRBVP(spec, spec.guess(n), n, DampedSolverOptions.from_spec(spec, slab=4))
    assert g.try_solve() is not None
    full = g.get_result()
    h = 1.0 / n
    for b in range(n_chains):
        c = 1.0 + 0.5 * b
        for m in range(length):
            exact = np.array([c * h ** m * comb(j, m) for j in range(n + 1)])
            col = full[:, b * length + (length - 1 - m)]
            assert np.max(np.abs(col - exact)) <= 1e-10 * max(1.0, np.max(np.abs(exact))), (b, m)
    g.close()
