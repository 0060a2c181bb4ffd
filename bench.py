#!/usr/bin/env python
"""bench.py -- Newton iterations / second (assembly + block solve) of the damped BVP hot path on B200.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line on rank 0.
  * a "step" is ONE complete damped Newton solve (``try_main_loop_damped``,
    src/numerical/BVP_Damp/NR_Damp_solver_damped.rs:2027-2156) from the configured initial guess to convergence:
    Jacobian assembly + factorisation when the age policy asks for it, and per iteration >= 2 x (residual assembly +
    linear solve) + the damping reductions.  The metric is the reference's own: iterations of that loop per second
    (its story tests time whole solves and count iterations, BVP_DAMP_STORY_TESTS.md:711-732).
    The state is restored from a device-side checkpoint before every step (no host traffic in ``value``).
  * default workload: BASELINE.json configs[2] ("c3": 12-variable flame BVP, 12x12 blocks, 10^6 mesh points per GPU) -- the
    largest single-GPU configuration; at N > 1 every rank owns 10^6 points of an N-times larger mesh (weak scaling;
    N = 8 is 8 * 10^6 points, the size class of configs[3]).  ``--workload c1|c2|c4|c5`` time configs[0] (2 variables,
    10^3 points), configs[1] (2 variables, 10^6 points), configs[3] (the flame at 10^7 points in total, split over the
    GPUs: strong scaling) and configs[4] (64-variable coupled case, 2.5*10^5 points per GPU = 2*10^6 on 8 GPUs).
    Compact results of the other configurations ride along under ``secondary`` (``--no-secondary`` skips them).
  * ``value``   : iterations/s, state resident in HBM (CUDA events on the launching stream, max over ranks)
  * ``e2e``     : the same through the host-buffer C-ABI calls a Rust host would make per solve
                  (rst_b200_set_state H2D -> rst_b200_newton_solve -> rst_b200_get_state D2H, pinned host vectors, both copies
                  of every solve inside the timed region).  ``e2e.value`` runs the solves on two handles / host threads per
                  rank (rustedscithe_b200.pipeline.SolvePipeline: one solve's copies under the other's kernels, results checked
                  against the one-handle answer before timing); ``e2e.one_handle_value`` is the plain sequence on one handle
                  (``--no-pipeline`` reports only that).
  * ``config.per_solve``: the reference's counters of one solve; ``linear_solves_executed`` of the counted ``linear_solves``
                  were run, the others are the accepted trial of the previous iteration taken over bit for bit.
  * ``kernels`` : per-phase launch times with the L2 flushed before each (assemble_FJ, assemble_F, factor, solve, and
                  factor_solve = the factorisation + first solve in one call, which is what follows every new Jacobian)
  * ``roofline``: dominant kernel of the solve, algorithmic bytes / measured launch duration
                  vs MEASURED_PEAKS.json hbm_gbs; ``traffic`` = ncu dram bytes of the committed capture (profiles/)
  * ``cpu_baseline`` / ``--impl reference``: the CPU oracle (port of the reference's path with its
                  refactor-per-solve behaviour; the Rust reference cannot be built in this image) on the host cores,
                  plus the refactor-OFF "best CPU" variant SURVEY.md section 8(d) asks for.
  * ``multi_gpu_check`` (N > 1): after the timed region every rank compares its slab of the converged solution with a
                  single-GPU solve of the same global problem (or the closed form for the linear case).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # key -> (problem, intervals per GPU, bounded CPU sample size of --impl reference, description)
    "c1": ("linear2", 1_000, 1_000, "2-variable BVP (examples/bvp_damped_aot_guide.rs), 1e3 mesh points"),
    "c2": ("linear2", 1_000_000, 1_000_000, "2-variable BVP (examples/bvp_damped_aot_guide.rs), 1e6 mesh points per GPU"),
    "c3": ("flame12", 1_000_000, 50_000, "12-variable flame BVP, 12x12 blocks, 1e6 mesh points per GPU"),
    "c4": ("flame12", 10_000_000, 50_000, "12-variable flame BVP, 12x12 blocks, 1e7 mesh points in total split over the GPUs (strong scaling)"),
    "c5": ("coupled64", 250_000, 1_500, "64-variable coupled BVP, 64x64 blocks, 2.5e5 mesh points per GPU (2e6 on 8 GPUs)"),
}
# full-size CPU sample of the cpu_baseline leg (one solve each, refactor ON and OFF)
CPU_FULL = {"c1": 1_000, "c2": 1_000_000, "c3": 1_000_000, "c4": 200_000, "c5": 1_500}
# FP64 peak of the pool's B200, measured with tools/dmma_bench.cu / tools/dfma_bench.cu (profiles/r02_fp64_peaks.md):
# DFMA 36.8 and DMMA m8n8k4 37.1 TFLOP/s -- one pipe, one peak
FP64_PEAK_TFLOPS = 37.1


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the committed `ncu --set full` captures, keyed
    problem -> phase -> bytes per interval (profiles/ncu_traffic.json names the capture each figure comes from)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, timeout=3.0):
        """nvidia-smi needs a few hundred ms before its first line; the timed region must not start before that."""
        t = time.time()
        while self.proc and not self.rows and time.time() - t < timeout:
            time.sleep(0.01)

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples inside [t0, t1] (wall clock); all samples if none fall inside."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if t0 is None or (t0 <= ts <= t1)]
        window = "timed region"
        if not rows:
            rows, window = [r for (_, r) in self.rows], "timed region + the identical steps run right after it"
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


class RawCuda:
    """Zero-copy view of a raw device pointer for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _oracle_solve(key, n, spec, refactor=True):
    """One complete damped Newton solve on the CPU oracle; refactor=True is the reference's behaviour (band copy +
    LU on every linear solve, BVP_traits.rs:875-899), False keeps the factors of a Jacobian ("best CPU")."""
    from oracle import oracle as orc
    o = orc.OracleBVP(key, n, spec.bc_list(), t0=spec.t0, t_end=spec.t_end)
    lo, hi = spec.bounds_per_var()
    lo, hi, rt = o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var())
    y0 = spec.guess(n)

    def run():
        t = time.perf_counter()
        res = o.newton(y0, lo, hi, rt, abs_tol=spec.abs_tol, max_iterations=spec.max_iterations, max_jac=spec.max_jac,
                       max_damp_iter=spec.max_damp_iter, damp_factor=spec.damp_factor, refactor_per_solve=refactor)
        return time.perf_counter() - t, res
    return run


def _cpu_sample_text(res, n, n_local, threads, refactor=True):
    how = "refactor-per-solve (reference behaviour, BVP_traits.rs:875-899)" if refactor else "factors kept per Jacobian (best CPU)"
    scaled = "" if n == n_local else ", time scaled linearly in N"
    return (f"one complete damped Newton solve ({res.iterations} iterations, {res.linear_solves} banded-LU solves, {how}, "
            f"{res.jac_recalcs} Jacobian) at {n} of {n_local} points{scaled}; assembly on {threads} threads, LU on 1 thread "
            f"as in the reference")


def _config(workload, desc, key, nv, n_local, n_global, world, exchange, extra=None):
    """`config` of the JSON line -- built by ONE function for both arms so their key sets are identical."""
    c = {"workload": f"{workload}: {desc}", "problem": key, "nv": nv, "points_per_gpu": n_local, "points_total": n_global,
         "parallelism": f"mesh-slab x{world}", "interface_exchange": exchange,
         "l2_policy": ("whole-solve steps: working set per step (state + blocks + factors = "
                       f"{n_local * (6 * nv + 7 * nv * nv) * 8 / 1e9:.2f} GB per GPU) exceeds the 126 MB L2 for every workload but c1; "
                       "per-kernel phase timings: L2 flushed (256 MB fill) before every timed launch"),
         "step": "one complete damped Newton solve from the initial guess (device checkpoint restore); value = iterations / second"}
    if extra:
        c.update(extra)
    return c


def _exchange_name(world, nccl=False):
    if world <= 1:
        return "none"
    return "nccl all_gather via host callback" if nccl else "NVLink peer-memory mailbox kernel (CUDA IPC)"


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as orc
    from rustedscithe_b200 import problems
    key, n_local, n_cpu, desc = WORKLOADS[args.workload]
    strong = args.workload == "c4"
    gpus = max(1, args.gpus)
    if strong:
        n_local //= gpus
    if args.points:
        n_local = args.points
    spec = problems.get(key)
    n = min(n_local, args.cpu_points or n_cpu)
    run = _oracle_solve(key, n, spec)
    if args.warmup >= 1:   # the first TWO oracle solves are slow (OpenMP pool start-up, first touch)
        if run()[0] < 3.0:
            run()
    times, res = [], None
    for _ in range(args.steps):
        dt, res = run()
        times.append(dt)
    sec = float(np.mean(times)) * (n_local / n)           # scaled to the full per-GPU mesh
    # weak scaling: the CPU does the world-times larger mesh in world-times the time -> the same unit-iterations/s
    # (linear-in-N assumption for the CPU; checked at full size by the cpu_baseline leg of the b200 arm)
    value = res.iterations / sec
    per_solve = {"iterations": res.iterations, "linear_solves": res.linear_solves, "jac_recalcs": res.jac_recalcs,
                 "factorizations": res.linear_solves}
    line = {
        "impl": "reference", "metric": "newton_iters_per_sec", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args.workload, desc, key, spec.nv, n_local, n_local * gpus, gpus, _exchange_name(gpus),
                          {"per_solve": per_solve}),
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": orc.num_threads(), "kind": "port",
                         "sample": f"{args.steps} x " + _cpu_sample_text(res, n, n_local, orc.num_threads())},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle = port of the reference's algorithms (Rust toolchain absent); value at N > 1 assumes the CPU's "
                "time grows linearly with the mesh (per-GPU mesh share per unit iteration)",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(key, spec, n_local, n_full):
    """cpu_baseline of the b200 arm (rank 0, N = 1): ONE full-size solve with the reference's refactor-per-solve behaviour
    and ONE with the factors kept ("best CPU"), after an untimed small warm-up."""
    from oracle import oracle as orc
    n_c = min(n_local, n_full)
    warm = _oracle_solve(key, min(n_c, 20_000), spec)
    warm()
    warm()
    t_on, res = _oracle_solve(key, n_c, spec, refactor=True)()
    t_off, res_off = _oracle_solve(key, n_c, spec, refactor=False)()
    scale = n_local / n_c
    thr = orc.num_threads()
    return {"value": res.iterations / (t_on * scale), "unit": "iter/s", "cores": thr, "kind": "port",
            "sample": "1 x " + _cpu_sample_text(res, n_c, n_local, thr, True),
            "best_cpu": {"value": res_off.iterations / (t_off * scale), "unit": "iter/s", "cores": thr,
                         "sample": "1 x " + _cpu_sample_text(res_off, n_c, n_local, thr, False)},
            "seconds": {"refactor_per_solve": t_on, "factors_kept": t_off}}


def multi_gpu_check(solver, key, spec, n_global, nv, rank, world, local_rank, per_solve, args):
    """Every rank's slab of the converged multi-GPU solution against (a) the closed form y = x, z = 1 of the linear
    problem, or (b) a single-GPU solve of the same global problem run on rank 0 and broadcast.  Also ||F||_inf of the
    converged state and the iteration counters."""
    import torch
    import torch.distributed as dist
    from rustedscithe_b200 import lib as rlib
    from rustedscithe_b200.solver import NRBVP, DampedSolverOptions
    Lc = solver.L
    b, e = C.c_int64(), C.c_int64()
    rlib.check(Lc.rst_b200_dist_range(solver.handle, C.byref(b), C.byref(e)))
    j0, j1 = b.value, e.value
    own = (j1 - j0 + 1) if rank == world - 1 else (j1 - j0)
    Y = torch.as_tensor(RawCuda(solver.device_ptr(0), (j1 - j0 + 1) * nv), device="cuda").view(-1, nv)[:own]
    solver.eval_residual(0)
    solver.sync()
    F = torch.as_tensor(RawCuda(solver.device_ptr(1), (j1 - j0) * nv), device="cuda")
    res_inf = F.abs().max().reshape(1).clone()
    info = {"method": None, "max_rel_diff": None, "iterations_multi": per_solve["iterations"], "iterations_single": None}
    if key == "linear2":
        x = torch.arange(j0, j0 + own, device="cuda", dtype=torch.float64) / n_global
        err = torch.maximum((Y[:, 0] - x).abs().max(), (Y[:, 1] - 1.0).abs().max()).reshape(1).clone()
        info["method"] = "closed form y = x, z = 1 (forward Euler is exact for it)"
    else:
        need_gb = n_global * (6 * nv + 7 * nv * nv) * 8 / 1e9
        have_gb = (n_global // world) * (6 * nv + 7 * nv * nv) * 8 / 1e9      # what rank 0 already holds
        if need_gb + have_gb > 150.0:
            err = torch.full((1,), float("nan"), device="cuda", dtype=torch.float64)
            info["method"] = (f"solution check skipped: the single-GPU rerun of {n_global} points needs {need_gb:.0f} GB next to the "
                              f"{have_gb:.0f} GB of rank 0's own slab (residual_inf and the multi-GPU parity tests cover this case)")
        else:
            full = torch.empty((n_global + 1) * nv, device="cuda", dtype=torch.float64)
            it1 = torch.zeros(1, device="cuda", dtype=torch.float64)
            if rank == 0:
                one = NRBVP(key, spec.guess(n_global), n_global, DampedSolverOptions.from_spec(spec, device=local_rank, slab=args.slab))
                sol = one.try_solve()
                if sol is None:
                    raise SystemExit("single-GPU rerun did not converge")
                it1[0] = one.get_statistics()["number of iterations"]
                rlib.check(Lc.rst_b200_memcpy(C.c_void_p(full.data_ptr()), C.c_void_p(one.device_ptr(0)), full.numel() * 8, 3))
                one.close()
            dist.broadcast(full, 0)
            dist.broadcast(it1, 0)
            ref = full.view(-1, nv)[j0:j0 + own]
            err = ((Y - ref).abs() / torch.clamp(ref.abs(), min=1.0)).max().reshape(1).clone()
            info["method"] = "single-GPU solve of the same global problem on rank 0, broadcast"
            info["iterations_single"] = int(it1.item())
            del full
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    dist.all_reduce(res_inf, op=dist.ReduceOp.MAX)
    info["max_rel_diff"] = float(err.item())
    info["residual_inf"] = float(res_inf.item())
    skipped = info["max_rel_diff"] != info["max_rel_diff"]
    ok = (skipped or info["max_rel_diff"] < 1e-9) and (info["iterations_single"] in (None, info["iterations_multi"]))
    info["ok"] = bool(ok)
    if not ok and rank == 0:
        print(f"multi-GPU check FAILED: {info}", file=sys.stderr, flush=True)
    return info


def run_workload(args, workload, rank, world, local_rank, primary):
    """Everything measured for one workload on the GPU(s).  Returns the dict rank 0 prints (primary) or nests (secondary)."""
    import torch
    import torch.distributed as dist
    from rustedscithe_b200 import lib as rlib
    from rustedscithe_b200 import problems
    from rustedscithe_b200.solver import NRBVP, DampedSolverOptions

    key, n_local, n_cpu, desc = WORKLOADS[workload]
    strong = workload == "c4"      # fixed total mesh, split over the ranks
    if strong:
        n_local //= world
    if args.points and primary:
        n_local = args.points
    steps = args.steps if primary else max(3, min(args.steps, 5))
    warmup = args.warmup if primary else 3
    spec = problems.get(key)
    nv = spec.nv
    n_global = n_local * world
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        opts = DampedSolverOptions.from_spec(spec, device=local_rank, stream=stream.cuda_stream, rank=rank,
                                             world=world, slab=args.slab)
        y0 = spec.guess(n_global)
        solver = NRBVP(key, y0, n_global, opts)
        if world > 1 and not args.nccl_exchange:
            # NVLink peer-memory mailboxes: exchange the IPC handles once, then the library gathers on its own stream
            handles = [None] * world
            dist.all_gather_object(handles, solver.p2p_handle())
            solver.p2p_open(handles)
            dist.barrier()               # every rank has mapped its peers before the first exchange
        elif world > 1:
            views = {}   # the library's exchange buffers have fixed addresses: wrap each of them once

            def allgather(send, recv, count, _stream):
                key_ = (send, recv, count)
                if key_ not in views:
                    views[key_] = (torch.as_tensor(RawCuda(send, count), device="cuda"),
                                   torch.as_tensor(RawCuda(recv, count * world), device="cuda"))
                s, r = views[key_]
                dist.all_gather_into_tensor(r, s)
                return 0
            solver.set_allgather(allgather)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        solver.checkpoint(restore=False)   # device-side copy of the initial state
        Lc = solver.L
        nopts = solver._opts()

        def step():
            """One complete damped Newton solve from the initial guess (state restored device-to-device)."""
            solver.checkpoint(restore=True)
            rlib.check(Lc.rst_b200_newton_solve(solver.handle, C.byref(nopts), C.byref(solver.stats)))
            if solver.stats.status != 1:
                raise SystemExit(f"Newton did not converge: status {solver.stats.status}")
            return solver.stats.iterations

        # ---- device-resident throughput -------------------------------------------------------------
        for _ in range(warmup):
            step()
        launches0 = solver.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0 and primary:
            sampler.start()
            sampler.wait_ready()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record(stream)
        iters = 0
        for _ in range(steps):
            iters += step()
        e1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        launches = solver.launch_count() - launches0
        st = solver.stats
        # linear_solves is the reference's count; solves_reused of them are the accepted first trial of the previous iteration
        # taken over bit for bit (same y, same factors) instead of being recomputed
        per_solve = {"iterations": st.iterations, "linear_solves": st.linear_solves, "jac_recalcs": st.jac_recalcs,
                     "factorizations": st.factorizations, "linear_solves_executed": st.linear_solves - st.solves_reused}
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        clocks = None
        if primary:
            if ms < 250.0:
                # the timed region is shorter than a few nvidia-smi periods: every rank keeps the SAME steps running (untimed,
                # the same count everywhere: the steps contain collective exchanges) so the sampler sees this workload's clocks
                for _ in range(min(5000, int(250.0 / (ms / steps)) + 1)):
                    step()
                barrier()
            clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        ms_per_step = ms / steps
        # whole-job aggregate: Newton iterations/s normalised to the per-GPU mesh (weak scaling: an iteration of the
        # world-times larger system counts as `world` unit iterations); equals plain iterations/s at N = 1
        mult = 1 if strong else world
        value = mult * iters / (ms * 1e-3)

        # ---- N > 1: is the multi-GPU answer the single-GPU answer? (after the timed region, untimed) ----------------
        check = None
        if world > 1:
            check = multi_gpu_check(solver, key, spec, n_global, nv, rank, world, local_rank, per_solve, args)

        # ---- secondary: every iteration with a Jacobian refresh + refactorisation (worst case of the age policy) ----
        def refresh_step():
            solver.checkpoint(restore=True)
            if solver.newton_iterate(recalc_jacobian=True) < 0:
                raise SystemExit("damped step failed")
        for _ in range(3):
            refresh_step()
        barrier()
        e0.record(stream)
        for _ in range(steps):
            refresh_step()
        e1.record(stream)
        barrier()
        refresh_ms = e0.elapsed_time(e1) / steps

        # ---- end to end through host buffers (what a host that owns the unknown vector does per solve) ------------
        host_y = torch.from_numpy(y0.copy()).pin_memory()
        host_np = host_y.numpy()
        ptr = host_np.ctypes.data_as(C.POINTER(C.c_double))
        host_out = torch.zeros_like(host_y).pin_memory()
        out_np = host_out.numpy()
        optr = out_np.ctypes.data_as(C.POINTER(C.c_double))

        def e2e_step():
            rlib.check(Lc.rst_b200_set_state(solver.handle, ptr, host_np.shape[0]))       # H2D initial guess
            rlib.check(Lc.rst_b200_newton_solve(solver.handle, C.byref(nopts), C.byref(solver.stats)))
            rlib.check(Lc.rst_b200_get_state(solver.handle, optr, out_np.shape[0]))       # D2H converged unknowns
            return solver.stats.iterations

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        e2e_iters = 0
        for _ in range(steps):
            e2e_iters += e2e_step()
        e1.record(stream)
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        bytes_local = n_local * nv * 8
        serial_e2e = mult * e2e_iters / (e2e_ms * 1e-3)
        e2e_steps, pipeline = steps, "off: one handle, copy -> solve -> copy in sequence"

        # ---- the same calls from TWO host threads, each with its own solver handle and stream (N = 1): the PCIe copies of
        # one solve overlap the kernels of the other.  Every solve still pays its own H2D + D2H inside the timed region.
        footprint_gb = n_local * (6 * nv + 7 * nv * nv) * 8 / 1e9
        if not args.no_pipeline and footprint_gb < 40.0 and (world == 1 or not args.nccl_exchange):
            from rustedscithe_b200.pipeline import SolvePipeline
            pipe = SolvePipeline(key, n_global, DampedSolverOptions.from_spec(spec, device=local_rank, rank=rank, world=world,
                                                                              slab=args.slab), lanes=2)
            if world > 1:       # every lane has its own NVLink mailboxes: wire lane k to the peers' lane k
                for sv in pipe.solvers:
                    handles = [None] * world
                    dist.all_gather_object(handles, sv.p2p_handle())
                    sv.p2p_open(handles)
                dist.barrier()
            guess = pipe.pinned(y0)
            # untimed: three solves per lane (direct launches, CUDA-graph capture, replay) and the check against the one-handle answer
            sols, _ = pipe.solve_many([guess] * 6)
            if not all(s_ is not None and np.array_equal(s_, out_np) for s_ in sols):
                raise SystemExit("pipelined e2e: the lanes did not reproduce the one-handle solution")
            del sols
            n_pipe = 2 * max(1, (steps + 1) // 2)
            barrier()
            t0 = time.perf_counter()
            done, st_pipe = pipe.solve_many([guess] * n_pipe, sink=lambda i, view: None)
            barrier()
            pipe_ms = (time.perf_counter() - t0) * 1e3
            if not all(done):
                raise SystemExit("pipelined e2e: a solve did not converge")
            t = torch.tensor([pipe_ms], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_iters, e2e_ms, e2e_steps = sum(s_[1] for s_ in st_pipe), float(t.item()), n_pipe
            pipeline = ("rustedscithe_b200.pipeline.SolvePipeline, 2 solver handles on 2 host threads per rank: set_state (H2D) -> "
                        "newton_solve -> get_state (D2H) per solve from / into pinned host vectors; the solves take turns on the GPU "
                        "(fixed order) while the other handle's copies run; wall clock over all solves, barrier + synchronise on both sides")
            pipe.close()
        # PCIe rate of the two copies alone (same pinned buffers), so the host-side limiter has a name
        barrier()
        tc = time.perf_counter()
        for _ in range(3):
            rlib.check(Lc.rst_b200_set_state(solver.handle, ptr, host_np.shape[0]))
            rlib.check(Lc.rst_b200_get_state(solver.handle, optr, out_np.shape[0]))
        copy_ms = (time.perf_counter() - tc) * 1e3 / 3
        barrier()

        # ---- per-phase kernel timings for the roofline (same stream, CUDA events, L2 flushed before every launch) ------
        phases = {}
        regrid_ms = None
        if world == 1 or not args.nccl_exchange:   # N > 1: every rank runs the same phases (the exchanges are inside factor / solve)
            flush_buf = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

            def time_phase(fn, reps):
                fn()
                torch.cuda.synchronize()
                tot = 0.0
                for _ in range(reps):
                    flush_buf.fill_(1.0)          # 256 MB > 126 MB L2: the phase starts from HBM
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    fn()
                    b.record(stream)
                    torch.cuda.synchronize()
                    tot += a.elapsed_time(b)
                return tot / reps

            reps = max(5, min(20, steps))
            phases["assemble_FJ"] = time_phase(lambda: solver.eval_jacobian(0), reps)
            phases["assemble_F"] = time_phase(lambda: solver.eval_residual(0), reps)
            solver.eval_jacobian(0)
            phases["factor"] = time_phase(lambda: solver.factor(), reps)
            phases["solve"] = time_phase(lambda: solver.solve(), reps)
            # what the Newton loop runs after a new Jacobian: factorisation + first solve in one call (nv = 12: the right-hand
            # side rides through the factorisation, no forward sweeps)
            phases["factor_solve"] = time_phase(lambda: solver.factor_solve(), reps)
            if primary and world == 1:
                # adaptive-grid hand-off on the current state (marks + scan + refined mesh + interpolated state, all on the device)
                regrid_ms = time_phase(lambda: solver.new_grid("Pearson", (1e-3, 1.4), fetch=False), 3)
            del flush_buf
        peak, peak_src = _peaks()
        alg = {  # algorithmic bytes per interval, SURVEY.md section 8(d)
            "assemble_FJ": 8 * (2 * nv + nv * nv), "assemble_F": 8 * 2 * nv,
            "factor": 8 * 3 * nv * nv, "solve": 8 * (2 * nv * nv + 2 * nv),
            "factor_solve": 8 * 3 * nv * nv + 8 * (2 * nv * nv + 2 * nv),
        }
        roofline, kernels = None, {}
        if phases:
            traffic_tab = _ncu_traffic().get(key, {})
            for k, v in phases.items():
                gbs = alg[k] * n_local / (v * 1e-3) / 1e9
                tr = traffic_tab.get(k) if k != "factor_solve" else None
                kernels[k] = {"ms": v, "alg_bytes_per_interval": alg[k], "achieved_gbs": gbs, "frac": gbs / peak,
                              "traffic_bytes_per_interval": tr["bytes_per_interval"] if tr else None}
            # per Newton solve: every factorisation runs as factor_solve (= the factorisation + the first linear solve after it);
            # the part of it beyond the plain factorisation is booked under "solve"
            nf, nx = per_solve["factorizations"], per_solve["linear_solves_executed"]
            share = {"assemble_FJ": per_solve["jac_recalcs"] * phases["assemble_FJ"],
                     "factor": nf * phases["factor"],
                     "assemble_F": max(nx - nf, 0) * phases["assemble_F"],
                     "solve": max(nx - nf, 0) * phases["solve"] + nf * max(phases["factor_solve"] - phases["factor"], 0.0)}
            dom = max(share, key=share.get)
            tr = traffic_tab.get(dom)
            traffic = tr["bytes_per_interval"] * n_local if tr else None
            if nv > 16:
                # wide blocks: the factorisation is FP64-pipe bound (>= 16 flop per byte of factor traffic, ridge ~6).  Useful
                # flops per block row = min(own algorithm 7.7 nv^3, DGBTRF 2 nv kl (kl + ku)) with (kl, ku) = (3nv/2 - 1, nv/2)
                kl, ku = 3 * nv // 2 - 1, nv // 2
                flops_row = min(7.7 * nv ** 3, 2.0 * nv * kl * (kl + ku))
                tfl = flops_row * n_local / (phases["factor"] * 1e-3) / 1e12
                kernels["factor"].update({"useful_flops_per_interval": flops_row, "achieved_tflops": tfl,
                                          "fp64_frac": tfl / FP64_PEAK_TFLOPS})
            if nv > 16 and dom == "factor":
                roofline = {"kernel": "factor", "bound": "fp64", "achieved": kernels["factor"]["achieved_tflops"],
                            "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": kernels["factor"]["fp64_frac"], "traffic": traffic,
                            "traffic_source": tr["source"] if tr else None,
                            "peak_source": "measured (tools/dmma_bench.cu: DFMA and DMMA m8n8k4 share one pipe, 37.1 TFLOP/s)",
                            "share_of_step": share[dom] / ms_per_step,
                            "note": "useful flops = min(own algorithm, DGBTRF 2 nv kl (kl+ku)) per SURVEY.md section 8(d)"}
            else:
                roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                            "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic,
                            "traffic_source": tr["source"] if tr else None, "peak_source": peak_src,
                            "share_of_step": share[dom] / ms_per_step,
                            "note": "solve = forward + backward sweeps over all condensation levels"}

        e2e = {"value": mult * e2e_iters / (e2e_ms * 1e-3), "unit": "iter/s", "h2d_bytes_per_step": bytes_local,
               "d2h_bytes_per_step": bytes_local + 40 * per_solve["linear_solves"],
               "steps": e2e_steps, "pipeline": pipeline, "one_handle_value": serial_e2e,
               "pcie_copy_ms_per_step": copy_ms, "pcie_gbs_per_rank": 2 * bytes_local / (copy_ms * 1e-3) / 1e9}
        out = {
            "value": value, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
            "scaling": "strong" if strong else "weak",
            "config": _config(workload, desc, key, nv, n_local, n_global, world, _exchange_name(world, args.nccl_exchange),
                              {"per_solve": per_solve}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
            "multi_gpu_check": check,
            "regrid": None if regrid_ms is None else {"method": "Pearson(1e-3, 1.4)", "ms": regrid_ms,
                                                       "note": "rst_b200_new_grid on the device-resident state"},
            "jacobian_refresh_every_iteration": {"value": mult * 1000.0 / refresh_ms, "unit": "iter/s",
                                                 "ms_per_iteration": refresh_ms},
        }
        solver.close()
        del host_y, host_out
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override intervals per GPU")
    ap.add_argument("--cpu-points", type=int, default=0, help="mesh size of the CPU sample (default per workload)")
    ap.add_argument("--slab", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the compact results of the other configurations")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e with one solver handle only (no second handle / host thread)")
    ap.add_argument("--nccl-exchange", action="store_true",
                    help="move the interface rows with torch.distributed/NCCL through the host callback instead of the "
                         "NVLink peer-memory kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from rustedscithe_b200 import build as rbuild
    from rustedscithe_b200 import lib as rlib
    from rustedscithe_b200 import problems

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    if rank == 0 and not os.path.exists(rlib.LIB_PATH):
        rbuild.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()

    res = run_workload(args, args.workload, rank, world, local_rank, primary=True)
    key = WORKLOADS[args.workload][0]
    spec = problems.get(key)

    # ---- compact results of the other configurations (same process, after the primary's buffers are released) --------
    secondary = {}
    if not args.no_secondary and not args.points:
        if world == 1:
            others = [w for w in ("c2", "c5") if w != args.workload]
        else:
            others = [w for w in ("c4", "c5", "c2") if w != args.workload]   # c5 weak: 2.5e5 points per GPU = 2e6 on 8 GPUs
        for w in others:
            try:
                r = run_workload(args, w, rank, world, local_rank, primary=False)
                rl = r["roofline"]
                secondary[w] = {
                    "workload": r["config"]["workload"], "value": r["value"], "unit": "iter/s", "scaling": r["scaling"],
                    "ms_per_step": r["ms_per_step"], "steps": r["steps"], "e2e": r["e2e"]["value"],
                    "per_solve": r["config"]["per_solve"], "multi_gpu_check": r["multi_gpu_check"],
                    "roofline": None if rl is None else {k: rl[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic")},
                    "kernels": {k: {"ms": v["ms"], "frac": v["frac"], **({"fp64_frac": v["fp64_frac"]} if "fp64_frac" in v else {})}
                                for k, v in r["kernels"].items()},
                }
            except SystemExit as ex:   # a secondary must never take the primary line down
                secondary[w] = {"error": str(ex)}
                if world > 1:
                    raise
            except Exception as ex:
                secondary[w] = {"error": repr(ex)[:300]}
                if world > 1:
                    raise

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the box's host cores -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_run = res["config"]["points_per_gpu"]
        cpu = cpu_baseline_leg(key, spec, n_run, args.cpu_points or CPU_FULL[args.workload])

    if rank == 0:
        line = {
            "metric": "newton_iters_per_sec", "value": res["value"], "unit": "iter/s", "n_gpus": world, "steps": args.steps,
            "unit_note": "damped Newton iterations per second per per-GPU mesh share, summed over GPUs (weak scaling); the "
                         "reference arm assumes CPU time linear in the mesh size",
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": res["config"], "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
            "roofline": res["roofline"], "kernels": res["kernels"], "cpu_baseline": cpu,
            "fp64_peaks": {"dfma_tflops": 36.8, "dmma_m8n8k4_tflops": 37.1,
                           "source": "tools/dfma_bench.cu, tools/dmma_bench.cu (profiles/r02_fp64_peaks.md): one pipe, one peak"},
            "multi_gpu_check": res["multi_gpu_check"], "regrid": res["regrid"],
            "jacobian_refresh_every_iteration": res["jacobian_refresh_every_iteration"],
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
