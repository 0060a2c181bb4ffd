"""Who is closer to the exact solution of J dY = F -- the faithful banded LU or the GPU condensation solver?

Exact reference: iterative refinement of the banded LU with the residual evaluated in 80-bit extended precision
(numpy longdouble) converges to the correctly rounded solution as long as cond(J) eps < 1.  Both solvers get the SAME
matrix (the oracle's Jacobian entries are uploaded into the GPU handle).  Usage: forward_error_probe.py problem N[,N...]
"""
import ctypes as C
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
from rustedscithe_b200 import lib as rlib
from rustedscithe_b200 import problems
from rustedscithe_b200.solver import NRBVP, DampedSolverOptions


def matvec_ld(n, kl, ku, banded, x):
    """Banded<f64> (storage.rs:5-16: data[(ku + i - j) n + j]) times x in extended precision."""
    D = banded.reshape(kl + ku + 1, n)
    xl = x.astype(np.longdouble)
    y = np.zeros(n, dtype=np.longdouble)
    for b in range(kl + ku + 1):
        off = b - ku                      # i = j + off
        j0, j1 = max(0, -off), min(n, n - off)
        if j1 > j0:
            y[j0 + off:j1 + off] += D[b, j0:j1].astype(np.longdouble) * xl[j0:j1]
    return y


def main():
    key = sys.argv[1]
    spec = problems.get(key)
    for n in [int(v) for v in sys.argv[2].split(",")]:
        rng = np.random.default_rng(23)
        o = orc.OracleBVP(key, n, spec.bc_list(), t0=spec.t0, t_end=spec.t_end)
        y = spec.guess(n) * (1.0 + 0.01 * rng.standard_normal(o.n))
        vals, rhs = o.jacobian_values(y), o.residual(y)
        rows, cols, _, _, kl, ku = o.structure()
        banded = o.banded(vals)
        ab, ipiv = orc.dgb_factor(o.n, kl, ku, orc.banded_to_ab(o.n, kl, ku, banded))
        x_lu = orc.dgb_solve(o.n, kl, ku, ab, ipiv, rhs)
        x = x_lu.astype(np.longdouble)
        bl = rhs.astype(np.longdouble)
        for it in range(6):
            r = bl - matvec_ld(o.n, kl, ku, banded, x)
            d = orc.dgb_solve(o.n, kl, ku, ab, ipiv, np.asarray(r, dtype=np.float64))
            x = x + d
            step = float(np.linalg.norm(d) / np.linalg.norm(np.asarray(x, dtype=np.float64)))
            if step < 1e-17:
                break
        x_true = np.asarray(x, dtype=np.float64)
        nv = spec.nv
        g = NRBVP(key, None, n, DampedSolverOptions.from_spec(spec))
        g.set_state(y)
        g.eval_jacobian()
        A = np.zeros(n * nv * nv)
        rlib.check(g.L.rst_b200_memcpy(A.ctypes.data_as(C.c_void_p), C.c_void_p(g.device_ptr(2)), A.nbytes, 2))
        full = o.unknown_pos[cols]
        node, comp = full // nv, full % nv
        jrow, irow = rows // nv, rows % nv
        m = node == jrow
        A[(jrow[m] * nv + irow[m]) * nv + comp[m]] = vals[m]
        rlib.check(g.L.rst_b200_memcpy(C.c_void_p(g.device_ptr(2)), A.ctypes.data_as(C.c_void_p), A.nbytes, 1))
        g.factor()
        x_gpu = g.solve_in_place(rhs)
        nrm = np.linalg.norm(x_true)
        print(f"{key} N={n}: last refinement step {step:.1e} | forward error vs extended-precision solution: banded LU "
              f"{np.linalg.norm(x_lu - x_true) / nrm:.3e}, GPU condensation {np.linalg.norm(x_gpu - x_true) / nrm:.3e}; "
              f"GPU vs LU {np.linalg.norm(x_gpu - x_lu) / nrm:.3e}", flush=True)
        g.close()


if __name__ == "__main__":
    main()
