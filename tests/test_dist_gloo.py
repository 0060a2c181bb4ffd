"""world_size-2/3 `gloo` tests (CPU) of the mesh-slab path's host logic: partition, the single all-gather of interface
rows per factor/solve, the redundant interface solve and the combination of the damping scalars.  The linear algebra
on each rank is the numpy model of the CUDA kernels (tests/abd_model.py); the result must equal the oracle's banded LU.
The real NCCL path runs in bench.py (torchrun) and the emulated-rank GPU test in tests/test_gpu_parity.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from rustedscithe_b200 import problems
from tests import abd_model as model


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _system(key, n_steps):
    spec = problems.get(key)
    o = orc.OracleBVP(key, n_steps, spec.bc_list())
    rng = np.random.default_rng(17)
    y = spec.guess(n_steps) * (1.0 + 0.01 * rng.standard_normal(o.n))
    rows, cols = o.structure()[:2]
    vals = o.jacobian_values(y)
    nv = spec.nv
    # block rows A_j (forward scheme: B_j = I) from the oracle's dense view, in FULL node-major columns
    full = o.full_solution(y)
    A = np.zeros((n_steps, nv, nv))
    for j in range(n_steps):
        _, fy = orc.problem_rhs(key, float(o.mesh[j]), full[j])
        A[j] = -np.eye(nv) - (o.mesh[j + 1] - o.mesh[j]) * fy
    F = o.residual(y).reshape(n_steps, nv)
    left = [v for v, side, _ in spec.bc_list() if side == 0]
    right = [v for v, side, _ in spec.bc_list() if side == 1]
    dy_ref = o.solve_banded(vals, F.reshape(-1))
    ref_full = np.zeros((n_steps + 1) * nv)
    ref_full[o.unknown_pos] = dy_ref
    return A, F, left, right, ref_full.reshape(n_steps + 1, nv), y, o


def _worker(rank, world, port, key, n_steps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, F, left, right, ref_full, y, o = _system(key, n_steps)
        nv = A.shape[1]
        b, e = model.partition(n_steps, rank, world)
        C, D, g, levels = model.local_condense(A[b:e], F[b:e])
        # ONE all-gather of 2 nv^2 + nv doubles per rank (rst_b200_set_allgather callback in the library)
        send = torch.from_numpy(np.concatenate([C.reshape(-1), D.reshape(-1), g]))
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        rows = [(r[:nv * nv].numpy().reshape(nv, nv), r[nv * nv:2 * nv * nv].numpy().reshape(nv, nv),
                 r[2 * nv * nv:].numpy()) for r in recv]
        ends = model.reduced_solve(rows, left, right)           # redundant on every rank
        dY = model.local_expand(levels, ends[rank], ends[rank + 1])
        err = np.linalg.norm(dY - ref_full[b:e + 1]) / np.linalg.norm(ref_full)
        # damping scalars: every rank reduces the nodes it owns, then sum / sum / min are combined
        own_hi = e + 1 if rank == world - 1 else e
        step = dY[: own_hi - b].reshape(-1)
        sc = torch.tensor([float(step @ step), float(np.min(step, initial=1.0))], dtype=torch.float64)
        gath = [torch.empty_like(sc) for _ in range(world)]
        dist.all_gather(gath, sc)
        total = sum(float(t[0]) for t in gath)
        out[rank] = (err, total, float(ref_full.reshape(-1) @ ref_full.reshape(-1)), b, e)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("key,n_steps,world", [("linear2", 101, 2), ("flame12", 90, 2), ("combustion6", 77, 3)])
def test_mesh_slab_ranks_reproduce_the_banded_lu_solution(key, n_steps, world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), key, n_steps, out), nprocs=world, join=True)
    assert len(out) == world
    covered = []
    for rank in range(world):
        err, total, ref_total, b, e = out[rank]
        assert err < 1e-10, (rank, err)
        assert abs(total - ref_total) <= 1e-9 * ref_total   # every unknown counted exactly once across ranks
        covered.append((b, e))
    assert covered[0][0] == 0 and covered[-1][1] == n_steps
    assert all(covered[i][1] == covered[i + 1][0] for i in range(world - 1))


def test_partition_matches_library_contract():
    for n, world in [(10, 3), (1_000_000, 8), (7, 2), (101, 2)]:
        edges = [model.partition(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
        assert all(e - b >= n // world for b, e in edges)
