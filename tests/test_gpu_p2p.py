"""NVLink mailbox exchange (rustedscithe_b200/csrc/p2p.cuh): protocol stress test on ONE device.

`world` emulated ranks run as the CTAs of one cooperative launch and call the same `p2p_allgather_block` the solver
fuses into its tail / reduction kernels (nothing in the reference to compare with -- the contract is
"every rank receives every rank's words of THIS exchange, bit for bit").  Spin-waiting kernels must not be separate
launches on one GPU, hence the single launch; the real multi-process path is covered by tests/test_multi_gpu.py.
"""
import ctypes as C

import pytest

from rustedscithe_b200 import lib as rlib

pytestmark = pytest.mark.gpu


def _selftest(world, rounds, pay, absent=-1, timeout_ms=20000.0):
    L = rlib.load()
    mism, nanw, tmo = C.c_int64(-1), C.c_int64(-1), C.c_int(-1)
    rlib.check(L.rst_b200_p2p_selftest(0, world, rounds, pay, absent, timeout_ms, C.byref(mism), C.byref(nanw), C.byref(tmo)), L)
    return mism.value, nanw.value, tmo.value


@pytest.mark.parametrize("world,pay", [(2, 8), (2, 288), (4, 288), (8, 288), (16, 32), (8, 2 * 64 * 64)])
def test_every_word_of_every_round_arrives_exactly(world, pay):
    # sizes: 8 damping scalars, the nv = 12 interface row (2 nv^2 = 288), the nv = 64 interface row (8192)
    rounds = 20000 if pay <= 288 else 300
    mism, nanw, tmo = _selftest(world, rounds, pay)
    assert (mism, nanw, tmo) == (0, 0, 0)


def test_missing_peer_times_out_and_poisons_instead_of_returning_stale_words():
    # one warm round trip is impossible with an absent rank: every present rank must flag the timeout and see NaN
    world = 4
    mism, nanw, tmo = _selftest(world, 5, 32, absent=2, timeout_ms=50.0)
    assert mism == 0
    assert tmo == world - 1                      # every rank that showed up reports RST_FLAG_PEER_TIMEOUT
    assert nanw == (world - 1) * 8               # first gather (8 words per rank): exactly the absent rank's words are NaN
