"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), REAL NVLink mailbox exchange between the ranks.

Each rank also solves the whole problem alone (world = 1 handle on its own device) and compares its slab of
  * the first linear solve dY = J^-1 F           (5e-9 relative: two condensation orders differ by the solver's forward
                                                 error, which grows with the mesh -- DESIGN.md section 4a; 1e-10 holds
                                                 for the small cases and is printed for all)
  * the converged solution and the iteration counters   (1e-9, equal counts)
with the single-rank result.  Exit code 0 = all ranks agree.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from rustedscithe_b200 import lib as rlib
    from rustedscithe_b200 import problems
    from rustedscithe_b200.solver import NRBVP, DampedSolverOptions

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    cases = [("linear2", 40_000), ("linear2", 1_000_003), ("flame12", 30_000), ("combustion6", 20_001), ("coupled64", 1_200)]
    report = []
    ok = True
    for key, n in cases:
        spec = problems.get(key)
        nv = spec.nv
        y0 = spec.guess(n)
        single = NRBVP(key, y0, n, DampedSolverOptions.from_spec(spec, device=local))
        multi = NRBVP(key, y0, n, DampedSolverOptions.from_spec(spec, device=local, rank=rank, world=world))
        handles = [None] * world
        dist.all_gather_object(handles, multi.p2p_handle())
        multi.p2p_open(handles)
        dist.barrier()                      # every rank has mapped its peers before the first exchange
        L = multi.L
        b, e = C.c_int64(), C.c_int64()
        rlib.check(L.rst_b200_dist_range(multi.handle, C.byref(b), C.byref(e)))
        j0, j1 = b.value, e.value

        def first_step(s, count):
            s.eval_jacobian(0)
            s.factor()
            s.eval_residual(0)
            s.solve()
            s.sync()
            out = np.zeros(count)
            rlib.check(L.rst_b200_memcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(s.device_ptr(4)), out.nbytes, 2))
            return out

        d1 = first_step(single, (n + 1) * nv).reshape(n + 1, nv)
        dm = first_step(multi, (j1 - j0 + 1) * nv).reshape(j1 - j0 + 1, nv)
        ref = d1[j0:j1 + 1]
        dy_err = float(np.max(np.abs(dm - ref)) / max(1e-300, np.max(np.abs(d1))))
        single.set_state(y0)
        multi.set_state(y0)
        s1 = single.try_solve()
        sm = multi.try_solve()
        st1, stm = single.get_statistics(), multi.get_statistics()
        full1 = single.get_result()
        fullm = np.zeros((n + 1) * nv)
        rlib.check(L.rst_b200_get_full_solution(multi.handle, fullm.ctypes.data_as(rlib.c_f64_p), fullm.shape[0]))
        fullm = fullm.reshape(n + 1, nv)
        hi = j1 + 1 if rank == world - 1 else j1
        sol_err = float(np.max(np.abs(fullm[j0:hi] - full1[j0:hi]) / np.maximum(1.0, np.abs(full1[j0:hi]))))
        same = all(st1[k] == stm[k] for k in ("number of iterations", "number of solving linear systems",
                                               "number of jacobians recalculations", "status"))
        good = s1 is not None and sm is not None and dy_err < 5e-9 and sol_err < 1e-9 and same
        ok = ok and good
        report.append({"problem": key, "n": n, "rank": rank, "dy_rel_err": dy_err, "sol_rel_err": sol_err,
                       "iterations": [st1["number of iterations"], stm["number of iterations"]], "ok": bool(good)})
        single.close()
        multi.close()
        dist.barrier()
    allrep = [None] * world
    dist.all_gather_object(allrep, report)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print("MGPU_REPORT " + json.dumps(allrep), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
