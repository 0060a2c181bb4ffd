import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
from rustedscithe_b200 import problems
from rustedscithe_b200.solver import NRBVP, DampedSolverOptions
for key, n in [("linear2", 1_000_000), ("flame12", 1_000_000)]:
    spec = problems.get(key)
    g = NRBVP(key, spec.guess(n), n, DampedSolverOptions.from_spec(spec))
    g.try_solve()
    for m, p in [("DoublePoints", ()), ("Easy", (0.01,)), ("Pearson", (1e-3, 1.4)), ("GrcarSmooke", (1e-3, 1e-2, 1.4))]:
        g.new_grid(m, p, fetch=False)
        g.L.rst_b200_sync(g.handle)
        t = time.perf_counter()
        for _ in range(5):
            r = g.new_grid(m, p, fetch=False)
        g.L.rst_b200_sync(g.handle)
        print(key, m, "ms", (time.perf_counter() - t) / 5 * 1e3, r)
    g.close()
