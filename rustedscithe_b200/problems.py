"""Problem catalogue: the BVPs named by BASELINE.json's configs, as data.

Each entry mirrors the arguments of the reference's ``NRBVP::new_with_options(eqs, guess, names, arg,
BCs, t0, t_end, n_steps, DampedSolverOptions)`` (src/numerical/BVP_Damp/NR_Damp_solver_damped.rs:847;
usage examples/bvp_damped_aot_guide.rs:52-81, src/numerical/BVP_Damp/tests/backend_compare.rs:72-182).
The right-hand sides are sympy expressions consumed by ``rustedscithe_b200.emit`` (the CUDA-C emitter).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import numpy as np


@dataclass
class ProblemSpec:
    key: str
    names: list                      # variable names, solver order
    rhs: Callable[[], list]          # () -> list of sympy expressions in sympy symbols named like `names`
    bcs: dict                        # name -> [(side, value)]   side 0 = left, 1 = right
    bounds: dict                     # name -> (lo, hi)
    rtol: dict                       # name -> relative tolerance
    guess: Callable[[int], np.ndarray]  # n_steps -> flat nv*n_steps (column-major nv x n_steps matrix)
    arg: str = "x"
    params: list = field(default_factory=list)
    param_values: list = field(default_factory=list)
    t0: float = 0.0
    t_end: float = 1.0
    abs_tol: float = 1e-6
    max_iterations: int = 25
    max_jac: int = 3
    max_damp_iter: int = 5
    damp_factor: float = 0.5

    @property
    def nv(self) -> int:
        return len(self.names)

    def bc_list(self):
        """[(var_index, side, value)] in variable order (numeric_discretization.rs:40-59)."""
        out = []
        for i, n in enumerate(self.names):
            for side, val in self.bcs.get(n, []):
                out.append((i, side, val))
        return out

    def bounds_per_var(self):
        return (np.array([self.bounds[n][0] for n in self.names]), np.array([self.bounds[n][1] for n in self.names]))

    def rtol_per_var(self):
        return np.array([self.rtol[n] for n in self.names])


def _sym(names):
    import sympy as sp
    return [sp.Symbol(n, real=True) for n in names]


# ---- C1 / C2: examples/bvp_damped_aot_guide.rs:42-81 ------------------------------------------------
def _linear2_rhs():
    import sympy as sp
    y, z = _sym(["y", "z"])
    return [z, sp.Float(0.0)]


def _linear2_guess(n_steps):
    h = 1.0 / n_steps
    g = np.zeros(2 * n_steps)
    g[0::2] = np.arange(n_steps) * h
    g[1::2] = 1.0
    return g


LINEAR2 = ProblemSpec(
    key="linear2", names=["y", "z"], rhs=_linear2_rhs,
    bcs={"y": [(0, 0.0)], "z": [(1, 1.0)]},
    bounds={"y": (-0.2, 1.2), "z": (-2.0, 2.0)}, rtol={"y": 1e-6, "z": 1e-6},
    guess=_linear2_guess, abs_tol=1e-8, max_iterations=40,
)


# ---- BVP_Damp1 fixture problem: codegen_test_support.rs:132-143 as pinned by the generated fixture
# (test_codegen_generated_bvp_fixtures.rs:13-66): z' = y - z, y' = z^3, names [z, y], h = 1 ---------
def _damp1_rhs():
    z, y = _sym(["z", "y"])
    return [y - z, z ** 3]


DAMP1 = ProblemSpec(
    key="damp1", names=["z", "y"], rhs=_damp1_rhs,
    bcs={"z": [(0, 1.0)], "y": [(1, 1.0)]},
    bounds={"z": (-10.0, 10.0), "y": (-10.0, 10.0)}, rtol={"z": 1e-6, "y": 1e-6},
    guess=lambda n: np.full(2 * n, 0.5),
)


def _damp1p_rhs():
    import sympy as sp
    z, y = _sym(["z", "y"])
    p0, p1, x = sp.Symbol("p0", real=True), sp.Symbol("p1", real=True), sp.Symbol("x", real=True)
    return [y - z, p0 * z ** 3 + p1 * x]


DAMP1P = ProblemSpec(
    key="damp1p", names=["z", "y"], rhs=_damp1p_rhs, params=["p0", "p1"], param_values=[0.7, 0.3],
    bcs={"z": [(0, 1.0)], "y": [(1, 1.0)]},
    bounds={"z": (-10.0, 10.0), "y": (-10.0, 10.0)}, rtol={"z": 1e-6, "y": 1e-6},
    guess=lambda n: np.full(2 * n, 0.5),
)


# ---- flame template: backend_compare.rs:72-149 (2 species = combustion6); 5 species = flame12 --------
def _flame_consts():
    q_heat = 3000.0 * 1e3 * 0.034
    dt = 600.0
    t_scale = 600.0
    l = 3e-4
    m0 = 34.2 / 1000.0
    lam = 0.07
    p = 2e6
    tm = 1500.0
    pe_q = 0.0090168
    d_ro = 2.88e-4
    pe_d = 1.50e-3
    ro_m = m0 * p / (8.314 * tm)
    return dict(q_heat=q_heat, dt=dt, t_scale=t_scale, lam=lam, pe_q=pe_q, d_ro=d_ro, pe_d=pe_d, ro_m=ro_m,
                a=1.3e5, e=5000.0 * 4.184, r_g=8.314, qm=l ** 2.0 / t_scale, qs=l ** 2.0, m_reag=0.342)


def flame_names(nsp):
    names = ["Teta", "q"]
    for k in range(nsp):
        names += [f"C{k}", f"J{k}"]
    return names


def flame_weight(nsp, k):
    return 2.0 * (nsp - k) / (nsp * (nsp - 1))


def _flame_rhs(nsp):
    def build():
        import sympy as sp
        c = {k: sp.Float(v, 17) for k, v in _flame_consts().items()}
        s = _sym(flame_names(nsp))
        teta, q = s[0], s[1]
        rate = c["a"] * sp.exp(-c["e"] / (c["r_g"] * (teta * c["t_scale"] + c["dt"]))) * s[2] * (c["ro_m"] / c["m_reag"])
        eqs = [q / c["lam"], q * c["pe_q"] - c["q_heat"] * rate * c["qm"]]
        for k in range(nsp):
            jk = s[3 + 2 * k]
            eqs.append(jk / c["d_ro"])
            if k == 0:
                eqs.append(jk * c["pe_d"] + rate * c["ro_m"] * c["qs"])
            else:
                eqs.append(jk * c["pe_d"] - sp.Float(flame_weight(nsp, k), 17) * rate * c["ro_m"] * c["qs"])
        return eqs
    return build


def _flame_spec(key, nsp):
    names = flame_names(nsp)
    bcs = {"Teta": [(0, (1000.0 - 600.0) / 600.0)], "q": [(1, 1e-10)]}
    bounds = {"Teta": (0.0, 10.0), "q": (-1e20, 1e20)}
    for k in range(nsp):
        bcs[f"C{k}"] = [(0, 1.0 if k == 0 else 1e-3)]
        bcs[f"J{k}"] = [(1, 1e-7)]
        bounds[f"C{k}"] = (0.0, 1.5)
        bounds[f"J{k}"] = (-1e2, 1e2)
    return ProblemSpec(
        key=key, names=names, rhs=_flame_rhs(nsp), bcs=bcs, bounds=bounds, rtol={n: 1e-5 for n in names},
        guess=lambda n, nv=len(names): np.full(nv * n, 0.99),
        abs_tol=1e-6, max_iterations=100, max_jac=6, max_damp_iter=6, damp_factor=0.5,
    )


COMBUSTION6 = _flame_spec("combustion6", 2)
FLAME12 = _flame_spec("flame12", 5)


# ---- coupled template (C5): every rate depends on all C_k --------------------------------------------
def _coupled_rhs(nsp, eps=1.0, e0=0.0):
    def build():
        import sympy as sp
        c = {k: sp.Float(v, 17) for k, v in _flame_consts().items()}
        s = _sym(flame_names(nsp))
        teta, q = s[0], s[1]
        T = teta * c["t_scale"] + c["dt"]
        w = sp.Float(1.0 / nsp, 17)
        rates = []
        for k in range(nsp):
            ek = sp.Float(_flame_consts()["e"] * (1.0 + 0.01 * k), 17)
            ssum = sum(sp.Float(eps, 17) * s[2 + 2 * m] / sp.Float(1.0 + abs(k - m), 17) for m in range(nsp))
            if e0:
                ssum = ssum + sp.Float(e0, 17) * s[2]
            rates.append(c["a"] * sp.exp(-ek / (c["r_g"] * T)) * (c["ro_m"] / c["m_reag"]) * ssum)
        heat = sum(w * r for r in rates)
        eqs = [q / c["lam"], q * c["pe_q"] - c["q_heat"] * c["qm"] * heat]
        for k in range(nsp):
            jk = s[3 + 2 * k]
            sgn = sp.Float(1.0, 17) if k == 0 else -w
            eqs.append(jk / c["d_ro"])
            eqs.append(jk * c["pe_d"] + sgn * rates[k] * c["ro_m"] * c["qs"])
        return eqs
    return build


def _coupled_spec(key, nsp, eps=1.0, e0=0.0, relax_bounds=False):
    spec = _flame_spec(key, nsp)
    spec.rhs = _coupled_rhs(nsp, eps, e0)
    if not relax_bounds:
        return spec
    # with many weakly produced species the first Newton iterates undershoot C_k = 0 slightly; the flame bounds (0, 1.5)
    # would stop the reference's damped loop at the bound (fbound < 1e-10), so the synthetic wide case relaxes them
    for k in range(nsp):
        spec.bounds[f"C{k}"] = (-1.0, 2.0)
    return spec


COUPLED8 = _coupled_spec("coupled8", 3)
COUPLED64 = _coupled_spec("coupled64", 31, eps=0.01, e0=1.0, relax_bounds=True)

CATALOGUE = {p.key: p for p in [LINEAR2, DAMP1, DAMP1P, COMBUSTION6, FLAME12, COUPLED8, COUPLED64]}


def get(key: str) -> ProblemSpec:
    return CATALOGUE[key]
