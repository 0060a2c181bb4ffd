import sys, os, numpy as np, time
sys.path.insert(0, '/root/repo')
from oracle import oracle as orc
from rustedscithe_b200 import problems
from rustedscithe_b200.solver import NRBVP, DampedSolverOptions
key = sys.argv[1] if len(sys.argv) > 1 else 'flame12'
spec = problems.get(key)
for n in [int(x) for x in sys.argv[2].split(',')]:
    rng = np.random.default_rng(23)
    o = orc.OracleBVP(key, n, spec.bc_list(), t0=spec.t0, t_end=spec.t_end)
    y = spec.guess(n) * (1.0 + 0.01 * rng.standard_normal(o.n))
    vals, rhs = o.jacobian_values(y), o.residual(y)
    dy_ref = o.solve_banded(vals, rhs)
    banded = o.banded(vals); _, _, _, _, kl, ku = o.structure()
    rr = np.linalg.norm(orc.banded_matvec(o.n, kl, ku, banded, dy_ref) - rhs) / np.linalg.norm(rhs)
    for slab in [int(s) for s in sys.argv[3].split(',')]:
        g = NRBVP(key, None, n, DampedSolverOptions.from_spec(spec, slab=slab))
        g.set_state(y); g.eval_jacobian(); g.factor()
        dy = g.solve_in_place(rhs)
        err = np.linalg.norm(dy - dy_ref) / np.linalg.norm(dy_ref)
        res = np.linalg.norm(orc.banded_matvec(o.n, kl, ku, banded, dy) - rhs) / np.linalg.norm(rhs)
        # one step of iterative refinement on the GPU factors
        r = rhs - orc.banded_matvec(o.n, kl, ku, banded, dy)
        dy2 = dy + g.solve_in_place(r)
        err2 = np.linalg.norm(dy2 - dy_ref) / np.linalg.norm(dy_ref)
        res2 = np.linalg.norm(orc.banded_matvec(o.n, kl, ku, banded, dy2) - rhs) / np.linalg.norm(rhs)
        print(f"{key} N={n} slab={slab} gen={os.environ.get('RST_B200_KERNELS','v4')}: err {err:.2e} res {res:.2e} (LU res {rr:.1e}) | refined err {err2:.2e} res {res2:.2e}", flush=True)
        g.close()
