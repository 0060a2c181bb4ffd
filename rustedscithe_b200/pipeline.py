"""Many independent solves of one discretised problem, host buffers in and out, copies hidden behind kernels.

One NRBVP solve from a host that owns the unknown vector is  set_state (H2D) -> newton_solve -> get_state (D2H)
(NR_Damp_solver_damped.rs:2130-2156 is that solve in the reference; its host owns the vectors as faer / nalgebra columns).
On one handle the three calls run back to back and the GPU idles during the copies (flame12 at 1e6 points: 3.5 ms of PCIe
next to 10.9 ms of kernels).  A host with MANY solves to do -- a sweep over parameters or initial guesses, the outer loop of
a continuation -- runs them on several handles, one worker thread each (the reference side of that is a thread pool over
NRBVP instances); the library's handles are independent (own stream, own buffers, no shared mutable state), so the copies of
one handle overlap the kernels of another.  This module is that pattern, nothing more: `lanes` handles, one thread per lane,
and a turn order (solve i computes after solve i - 1) so the lanes fall into copy / compute alternation instead of
contending for the SMs.  Every solve still pays its own H2D and D2H.
"""
from __future__ import annotations

import ctypes as C
import threading
import time

import numpy as np

from . import lib as _lib
from .solver import NRBVP, DampedSolverOptions


class PinnedVector:
    """A float64 numpy vector in page-locked host memory (rst_b200_host_alloc)."""

    def __init__(self, L, n):
        self._L = L
        self._p = C.c_void_p()
        _lib.check(L.rst_b200_host_alloc(int(n) * 8, C.byref(self._p)), L)
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_double)), shape=(int(n),))
        self.ptr = C.cast(self._p, _lib.c_f64_p)
        self.array[:] = 0.0      # world > 1: get_state writes only the nodes this rank owns

    def free(self):
        if self._p:
            self.array = None
            self._L.rst_b200_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SolvePipeline:
    """`lanes` solver handles for the same problem / mesh; `solve_many` distributes independent solves over them."""

    def __init__(self, problem, n_steps, options: DampedSolverOptions | None = None, lanes: int = 2, initial_guess=None, **kw):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.solvers = [NRBVP(problem, initial_guess, n_steps, options, **kw) for _ in range(lanes)]
        s0 = self.solvers[0]
        self.L, self.n = s0.L, s0.n
        self._in = [PinnedVector(self.L, self.n) for _ in range(lanes)]
        self._out = [PinnedVector(self.L, self.n) for _ in range(lanes)]

    def pinned(self, values=None) -> PinnedVector:
        """A page-locked vector of the state length (optionally filled): guesses kept in these skip the staging copy."""
        v = PinnedVector(self.L, self.n)
        if values is not None:
            v.array[:] = np.asarray(values, dtype=np.float64).reshape(-1)
        return v

    def close(self):
        for s in self.solvers:
            s.close()
        for b in self._in + self._out:
            b.free()
        self.solvers = []

    def solve_many(self, guesses, params=None, sink=None):
        """Solve one problem instance per entry of `guesses` (reduced state vectors, length n, or PinnedVector).  `params[i]`, if given, is
        set on the handle before solve i.  Returns (solutions, stats): solutions[i] is the converged reduced state or None
        (as NRBVP.try_solve), stats[i] = (status, iterations, linear_solves, jac_recalcs).  With `sink(i, view)` the
        converged state is handed over as a view of the lane's pinned buffer instead of being copied into a new array."""
        m = len(guesses)
        lanes = len(self.solvers)
        sols, stats, errors = [None] * m, [None] * m, []
        # Solve i runs on lane i % lanes and the solves take the GPU in the order 0, 1, 2, ...: deterministic, and with
        # world > 1 every rank issues the exchanges of its lanes in the same order (a free-for-all could leave two ranks
        # waiting for each other's OTHER lane).
        state = {"next": 0, "first_copies": 0}

        def lane(k):
            sv, hin, hout = self.solvers[k], self._in[k], self._out[k]
            L = self.L
            try:
                _lib.check(L.rst_b200_sync(sv.handle), L)          # binds the worker thread to the handle's device
                for i in range(k, m, lanes):
                    g = guesses[i]
                    if isinstance(g, PinnedVector):     # the host already keeps this guess in page-locked memory
                        src = g.ptr
                    else:
                        hin.array[:] = np.asarray(g, dtype=np.float64).reshape(-1)
                        src = hin.ptr
                    if params is not None:
                        sv.set_params(params[i])
                    # the first copies of a batch go one after the other (lane 0 first): sharing PCIe between them would only
                    # delay the first solve
                    while i < lanes and state["first_copies"] != i and not errors:
                        time.sleep(0)
                    try:
                        _lib.check(L.rst_b200_set_state(sv.handle, src, self.n), L)
                    finally:
                        if i < lanes:
                            state["first_copies"] = i + 1
                    opts = sv._opts()
                    # wait for the turn.  A condition variable costs ~0.2 ms per hand-over (waiter wake-up + GIL), during which
                    # the GPU idles; the other lane sits inside a ctypes call without the GIL, so yielding in a loop is cheap
                    # and hands over in microseconds
                    while state["next"] != i and not errors:
                        time.sleep(0)
                    if errors:
                        return
                    try:
                        _lib.check(L.rst_b200_newton_solve(sv.handle, C.byref(opts), C.byref(sv.stats)), L)
                    finally:
                        state["next"] = i + 1
                    st = sv.stats
                    stats[i] = (st.status, st.iterations, st.linear_solves, st.jac_recalcs)
                    if st.status == 1:
                        _lib.check(L.rst_b200_get_state(sv.handle, hout.ptr, self.n), L)
                        if sink is not None:
                            sink(i, hout.array)
                            sols[i] = True
                        else:
                            sols[i] = hout.array.copy()
            except Exception as exc:        # re-raised on the calling thread
                errors.append(exc)

        threads = [threading.Thread(target=lane, args=(k,)) for k in range(lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return sols, stats
