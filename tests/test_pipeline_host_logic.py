"""Host logic of rustedscithe_b200.pipeline.SolvePipeline without a GPU: the library calls are replaced by fakes that record
what the lanes do.  Checked: every solve runs on lane i % lanes, the solves take the (fake) GPU strictly in the order
0, 1, 2, ... (what keeps the exchanges of several ranks aligned), the copies of a lane overlap the solve of the other, results
and statistics land at the index of their guess, a failing solve surfaces as an exception and does not hang the other lane."""
import ctypes as C
import threading
import time

import numpy as np
import pytest

from rustedscithe_b200 import pipeline


class _Stats(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("linear_solves", C.c_int), ("jac_recalcs", C.c_int)]


class _FakeLib:
    def __init__(self):
        self.log, self.lock, self.bufs = [], threading.Lock(), {}
        self.solving = 0
        self.max_solving = 0
        self.copy_during_solve = 0

    def rst_b200_host_alloc(self, nbytes, out):
        buf = (C.c_double * (nbytes // 8))()
        self.bufs[C.addressof(buf)] = buf
        out._obj.value = C.addressof(buf)
        return 0

    def rst_b200_host_free(self, p):
        return 0

    def rst_b200_sync(self, h):
        return 0

    def rst_b200_set_state(self, h, ptr, n):
        time.sleep(0.001)            # like the real call, the copy runs without the GIL: the other lane gets to start its solve
        with self.lock:
            self.copy_during_solve += self.solving
        h.state = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
        time.sleep(0.001)
        return 0

    def rst_b200_newton_solve(self, h, opts, stats):
        with self.lock:
            self.solving += 1
            self.max_solving = max(self.max_solving, self.solving)
            self.log.append((h.lane, float(h.state[0])))
        time.sleep(0.004)
        if h.state[0] < 0:
            with self.lock:
                self.solving -= 1
            return -4
        h.state = h.state * 2.0
        with self.lock:
            self.solving -= 1
        return 0

    def rst_b200_get_state(self, h, ptr, n):
        np.ctypeslib.as_array(ptr, shape=(n,))[:] = h.state
        return 0


class _Handle:
    pass


def _install(monkeypatch, fake):
    lanes = []

    class FakeNRBVP:
        def __init__(self, problem, guess, n_steps, options=None, **kw):
            self.L, self.n = fake, 8
            self.handle = _Handle()
            self.handle.lane = len(lanes)
            lanes.append(self)
            self.stats = _Stats(1, 3, 6, 1)

        def _opts(self):
            return C.c_int(0)

        def set_params(self, p):
            self.handle.params = p

        def close(self):
            pass

    def check(rc, L=None):
        if rc != 0:
            raise RuntimeError(f"rc {rc}")
        return rc

    monkeypatch.setattr(pipeline, "NRBVP", FakeNRBVP)
    monkeypatch.setattr(pipeline._lib, "check", check)
    return lanes


@pytest.mark.parametrize("lanes", [1, 2, 3])
def test_solves_take_turns_in_order_and_results_land_at_their_index(monkeypatch, lanes):
    fake = _FakeLib()
    _install(monkeypatch, fake)
    pipe = pipeline.SolvePipeline("any", 4, None, lanes=lanes)
    guesses = [np.full(8, float(i + 1)) for i in range(7)]
    sols, stats = pipe.solve_many(guesses)
    assert [round(v) for _, v in fake.log] == [1, 2, 3, 4, 5, 6, 7]          # strict global order
    assert [lane for lane, _ in fake.log] == [i % lanes for i in range(7)]    # static lane assignment
    assert fake.max_solving == 1                                              # one solve computes at a time
    for i, s in enumerate(sols):
        np.testing.assert_array_equal(s, np.full(8, 2.0 * (i + 1)))
    assert all(st == (1, 3, 6, 1) for st in stats)
    if lanes > 1:
        assert fake.copy_during_solve > 0                                     # a lane's copies ran under another lane's solve
    pipe.close()


def test_pinned_guesses_skip_the_staging_copy_and_sink_gets_views(monkeypatch):
    fake = _FakeLib()
    _install(monkeypatch, fake)
    pipe = pipeline.SolvePipeline("any", 4, None, lanes=2)
    g = pipe.pinned(np.arange(8.0) + 1.0)
    seen = {}
    done, _ = pipe.solve_many([g, g, g], sink=lambda i, view: seen.__setitem__(i, view.copy()))
    assert done == [True, True, True] and sorted(seen) == [0, 1, 2]
    for v in seen.values():
        np.testing.assert_array_equal(v, 2.0 * (np.arange(8.0) + 1.0))
    pipe.close()


def test_a_failing_solve_raises_and_does_not_hang_the_other_lane(monkeypatch):
    fake = _FakeLib()
    _install(monkeypatch, fake)
    pipe = pipeline.SolvePipeline("any", 4, None, lanes=2)
    guesses = [np.full(8, 1.0), np.full(8, -1.0), np.full(8, 3.0), np.full(8, 4.0), np.full(8, 5.0)]
    t0 = time.time()
    with pytest.raises(RuntimeError):
        pipe.solve_many(guesses)
    assert time.time() - t0 < 5.0
    pipe.close()
