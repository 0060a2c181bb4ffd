#!/usr/bin/env python
"""bench.py -- Newton iterations / second (assembly + block solve) of the damped BVP hot path on B200.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line on rank 0.
  * a "step" is ONE complete damped Newton solve (``try_main_loop_damped``,
    src/numerical/BVP_Damp/NR_Damp_solver_damped.rs:2027-2156) from the configured initial guess to convergence:
    Jacobian assembly + factorisation when the age policy asks for it, and per iteration >= 2 x (residual assembly +
    linear solve) + the damping reductions.  The metric is the reference's own: iterations of that loop per second
    (its story tests time whole solves and count iterations, BVP_DAMP_STORY_TESTS.md:711-732).
    The state is restored from a device-side checkpoint before every step (no host traffic in ``value``).
  * N = 1 workload: BASELINE.json configs[1] ("c2": 2-variable BVP, 10^6 mesh points).  ``--workload c1|c3|c4|c5`` time
    configs[0] (10^3 points), configs[2] (12-variable flame, 10^6 points), configs[3] (the flame at 10^7 points in total,
    split over the GPUs: strong scaling) and configs[4] (64-variable coupled case, 2.5*10^5 points per GPU).  N > 1: weak
    scaling except c4, every rank owns the per-GPU share of a world-times larger mesh.
  * ``value``   : iterations/s, state resident in HBM (CUDA events on the launching stream, max over ranks)
  * ``e2e``     : the same through the host-buffer C-ABI calls a Rust host would make per solve
                  (rst_b200_set_state H2D -> rst_b200_newton_solve -> rst_b200_get_state D2H)
  * ``roofline``: dominant kernel of the solve, algorithmic bytes / measured launch duration
                  vs MEASURED_PEAKS.json hbm_gbs
  * ``cpu_baseline`` / ``--impl reference``: the CPU oracle (port of the reference's path with its
                  refactor-per-solve behaviour; the Rust reference cannot be built in this image) on the host cores.
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
    # key -> (problem, intervals per GPU, default CPU sample size, description)
    "c1": ("linear2", 1_000, 1_000, "2-variable BVP (examples/bvp_damped_aot_guide.rs), 1e3 mesh points"),
    "c2": ("linear2", 1_000_000, 1_000_000, "2-variable BVP (examples/bvp_damped_aot_guide.rs), 1e6 mesh points per GPU"),
    "c3": ("flame12", 1_000_000, 50_000, "12-variable flame BVP, 12x12 blocks, 1e6 mesh points per GPU"),
    "c4": ("flame12", 10_000_000, 50_000, "12-variable flame BVP, 12x12 blocks, 1e7 mesh points in total split over the GPUs (strong scaling)"),
    "c5": ("coupled64", 250_000, 1_500, "64-variable coupled BVP, 64x64 blocks, 2.5e5 mesh points per GPU (2e6 on 8 GPUs)"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def _oracle_solve(key, n, spec):
    """One complete damped Newton solve on the CPU oracle, reference behaviour (refactor on every linear solve)."""
    from oracle import oracle as orc
    o = orc.OracleBVP(key, n, spec.bc_list(), t0=spec.t0, t_end=spec.t_end)
    lo, hi = spec.bounds_per_var()
    lo, hi, rt = o.per_unknown(lo), o.per_unknown(hi), o.per_unknown(spec.rtol_per_var())
    y0 = spec.guess(n)

    def run():
        t = time.perf_counter()
        res = o.newton(y0, lo, hi, rt, abs_tol=spec.abs_tol, max_iterations=spec.max_iterations, max_jac=spec.max_jac,
                       max_damp_iter=spec.max_damp_iter, damp_factor=spec.damp_factor, refactor_per_solve=True)
        return time.perf_counter() - t, res
    return run


def _cpu_sample_text(res, n, n_local, threads):
    return (f"one complete damped Newton solve ({res.iterations} iterations, {res.linear_solves} refactor-per-solve "
            f"banded-LU solves, {res.jac_recalcs} Jacobian) at {n} of {n_local} points, time scaled linearly in N; "
            f"assembly on {threads} threads, LU serial as in the reference")


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as orc
    from rustedscithe_b200 import problems
    key, n_local, n_cpu, desc = WORKLOADS[args.workload]
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
    value = res.iterations / sec
    line = {
        "impl": "reference", "metric": "newton_iters_per_sec", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "problem": key, "points_per_gpu": n_local, "nv": spec.nv,
                   "step": "one complete damped Newton solve; value = iterations / second"},
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": orc.num_threads(), "kind": "port",
                         "sample": _cpu_sample_text(res, n, n_local, orc.num_threads())},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    from rustedscithe_b200.solver import NRBVP, DampedSolverOptions

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    if rank == 0 and not os.path.exists(rlib.LIB_PATH):
        rbuild.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()

    key, n_local, n_cpu, desc = WORKLOADS[args.workload]
    strong = args.workload == "c4"      # fixed total mesh, split over the ranks
    if strong:
        n_local //= world
    if args.points:
        n_local = args.points
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
        for _ in range(args.warmup):
            step()
        launches0 = solver.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            sampler.wait_ready()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record(stream)
        iters = 0
        for _ in range(args.steps):
            iters += step()
        e1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        launches = solver.launch_count() - launches0
        st = solver.stats
        per_solve = {"iterations": st.iterations, "linear_solves": st.linear_solves, "jac_recalcs": st.jac_recalcs,
                     "factorizations": st.factorizations}
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if ms < 250.0:
            # the timed region is shorter than a few nvidia-smi periods: every rank keeps the SAME steps running (untimed,
            # the same count everywhere: the steps contain collective exchanges) so the sampler sees this workload's clocks
            for _ in range(min(5000, int(250.0 / (ms / args.steps)) + 1)):
                step()
            barrier()
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        ms_per_step = ms / args.steps
        # whole-job aggregate: Newton iterations/s normalised to the per-GPU mesh (weak scaling: an iteration of the
        # world-times larger system counts as `world` unit iterations); equals plain iterations/s at N = 1
        value = (1 if strong else world) * iters / (ms * 1e-3)

        # ---- secondary: every iteration with a Jacobian refresh + refactorisation (worst case of the age policy) ----
        def refresh_step():
            solver.checkpoint(restore=True)
            if solver.newton_iterate(recalc_jacobian=True) < 0:
                raise SystemExit("damped step failed")
        for _ in range(3):
            refresh_step()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            refresh_step()
        e1.record(stream)
        barrier()
        refresh_ms = e0.elapsed_time(e1) / args.steps

        # ---- end to end through host buffers (what a host that owns the unknown vector does per solve) ------------
        host_y = torch.from_numpy(y0.copy()).pin_memory()
        host_np = host_y.numpy()
        ptr = host_np.ctypes.data_as(C.POINTER(C.c_double))
        host_out = torch.empty_like(host_y).pin_memory()
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
        for _ in range(args.steps):
            e2e_iters += e2e_step()
        e1.record(stream)
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        bytes_local = n_local * nv * 8

        # ---- per-phase kernel timings for the roofline (same stream, CUDA events, after warm-up) ----------
        def time_phase(fn, reps):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        reps = max(5, min(50, args.steps))
        phases = {}
        regrid_ms = None
        if world == 1:
            phases["assemble_FJ"] = time_phase(lambda: solver.eval_jacobian(0), reps)
            phases["assemble_F"] = time_phase(lambda: solver.eval_residual(0), reps)
            solver.eval_jacobian(0)
            phases["factor"] = time_phase(lambda: solver.factor(), reps)
            phases["solve"] = time_phase(lambda: solver.solve(), reps)
            # adaptive-grid hand-off on the current state (marks + scan + refined mesh + interpolated state, all on the device)
            regrid_ms = time_phase(lambda: solver.new_grid("Pearson", (1e-3, 1.4), fetch=False), 3)
        peak, peak_src = _peaks()
        alg = {  # algorithmic bytes per interval, SURVEY.md section 8(d)
            "assemble_FJ": 8 * (2 * nv + nv * nv), "assemble_F": 8 * 2 * nv,
            "factor": 8 * 3 * nv * nv, "solve": 8 * (2 * nv * nv + 2 * nv),
        }
        roofline, kernels = None, {}
        if phases:
            for k, v in phases.items():
                gbs = alg[k] * n_local / (v * 1e-3) / 1e9
                kernels[k] = {"ms": v, "alg_bytes_per_interval": alg[k], "achieved_gbs": gbs, "frac": gbs / peak}
            share = {"assemble_FJ": per_solve["jac_recalcs"] * phases["assemble_FJ"],
                     "factor": per_solve["factorizations"] * phases["factor"],
                     "assemble_F": per_solve["linear_solves"] * phases["assemble_F"],
                     "solve": per_solve["linear_solves"] * phases["solve"]}
            dom = max(share, key=share.get)
            # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel(s) from the committed `ncu --set full`
            # captures (profiles/r01_c_summary.md), per launch at 1e6 intervals (coupled64: measured at 2.5e5, scaled); None where no capture exists
            ncu_traffic = {("flame12", "solve"): (1.773e9 + 0.062e9) + (3.060e9 + 0.087e9),     # forward + backward, level 0
                           ("flame12", "factor"): 1.157e9 + 4.507e9,
                           ("coupled64", "factor"): (8.22e9 + 40.20e9) * 4.0, ("coupled64", "solve"): (16.03e9 + 0.13e9 + 23.94e9 + 0.13e9) * 4.0}
            traffic = ncu_traffic.get((key, dom))
            if traffic is not None:
                traffic *= n_local / 1e6
            if key == "coupled64" and dom == "factor":
                # wide blocks: 5 nv^3 FMA per block row against 160 KB of factor traffic (16 flop/B) -> FP64-pipe bound;
                # peak = DFMA rate measured with tools/dfma_bench.cu on this pool's B200 (63.5 DFMA/clk/SM)
                tflops = 2.0 * 5.0 * nv ** 3 * n_local / (phases["factor"] * 1e-3) / 1e12
                roofline = {"kernel": "factor", "bound": "fp64", "achieved": tflops, "peak": 36.9, "unit": "TFLOP/s",
                            "frac": tflops / 36.9, "traffic": traffic, "peak_source": "measured (tools/dfma_bench.cu)",
                            "share_of_step": share[dom] / ms_per_step}
            else:
                roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                            "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                            "share_of_step": share[dom] / ms_per_step,
                            "note": "solve = forward + backward sweeps over all condensation levels"}

        # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the box's host cores -------------------
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle import oracle as orc
            n_c = min(n_local, args.cpu_points or n_cpu)
            run = _oracle_solve(key, n_c, spec)
            if run()[0] < 3.0:   # untimed: OpenMP pool start-up and first touch make the first two solves slower;
                run()            # one is enough when a solve takes seconds
            tt, reps_cpu, res = time.perf_counter(), 0, None
            while reps_cpu < 3 and time.perf_counter() - tt < 20.0:
                _, res = run()
                reps_cpu += 1
            sec = (time.perf_counter() - tt) / reps_cpu * (n_local / n_c)
            cpu = {"value": res.iterations / sec, "unit": "iter/s", "cores": orc.num_threads(), "kind": "port",
                   "sample": f"{reps_cpu} x " + _cpu_sample_text(res, n_c, n_local, orc.num_threads())}

        if rank == 0:
            line = {
                "metric": "newton_iters_per_sec", "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps,
                "unit_note": "damped Newton iterations per second per 1e6-point mesh slab, summed over GPUs (weak scaling)",
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {desc}", "problem": key, "nv": nv, "points_per_gpu": n_local,
                           "points_total": n_global, "parallelism": f"mesh-slab x{world}",
                           "interface_exchange": ("none" if world == 1 else ("nccl all_gather via host callback" if args.nccl_exchange
                                                                             else "NVLink peer-memory mailbox kernel (CUDA IPC)")),
                           "l2_policy": "working set per step (state + blocks + factors) exceeds the 126 MB L2"
                                        if n_local * nv * nv * 8 * 6 > 126e6 else "working set smaller than L2 (C1)",
                           "step": "one complete damped Newton solve from the initial guess (device checkpoint restore); "
                                   "value = iterations / second", "per_solve": per_solve},
                "clocks": clocks,
                "e2e": {"value": (1 if strong else world) * e2e_iters / (e2e_ms * 1e-3), "unit": "iter/s", "h2d_bytes_per_step": bytes_local,
                        "d2h_bytes_per_step": bytes_local + 40 * per_solve["linear_solves"]},
                "gpu_launches": int(launches),
                "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
                "regrid": None if regrid_ms is None else {"method": "Pearson(1e-3, 1.4)", "ms": regrid_ms,
                                                           "note": "rst_b200_new_grid on the device-resident state"},
                "jacobian_refresh_every_iteration": {"value": (1 if strong else world) * 1000.0 / refresh_ms, "unit": "iter/s",
                                                     "ms_per_iteration": refresh_ms},
            }
            print(json.dumps(line), flush=True)
        solver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
