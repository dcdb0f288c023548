"""Collections.shuffle on java.util.Random as the emit step performs it before a shuffled multi-send
(wittgenstein_b200/csrc/wtg_cappos.cuh: javaShuffleAt), including nextInt's rejection loop, against the oracle's
JavaRandom / javaShuffle.  wtg_java_shuffle runs that code on the host: no GPU needed."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib
from wittgenstein_b200 import _lib

A, CADD, MASK = 0x5DEECE66D, 0xB, (1 << 48) - 1
A_INV = pow(A, -1, 1 << 48)


def _state_before(next_state):
    """LCG state whose successor is next_state"""
    return ((next_state - CADD) * A_INV) & MASK


def _shuffle_both(state, n):
    api = _lib.api()
    o = oracle_lib.load()
    a = np.arange(n, dtype=np.int32)
    used = api.check(api.java_shuffle(C.c_ulonglong(state), n, a.ctypes.data_as(C.POINTER(C.c_int))))
    b = np.arange(n, dtype=np.int32)
    o.wo_shuffle(C.c_int64(state ^ A), n, b.ctypes.data_as(C.POINTER(C.c_int32)))  # new Random(seed): state = seed ^ multiplier
    return a, b, used


@pytest.mark.parametrize("n", [1, 2, 3, 7, 50, 51, 64])
def test_shuffle_matches_the_jdk_restatement(n):
    for state in (0, 1, 0x5DEECE66D, 123456789012345, MASK):
        a, b, used = _shuffle_both(state, n)
        assert (a == b).all()
        assert used >= max(0, n - 1)


@pytest.mark.parametrize("n", [3, 7, 50, 51])
def test_shuffle_with_a_rejected_draw(n):
    """next(31) == 2^31 - 1 is rejected by nextInt(bound) for these bounds (bits - val + (bound - 1) overflows): the first draw
    of the shuffle loops, so one more value than n - 1 is consumed and every later swap uses the shifted stream."""
    assert ((1 << 31) - 1) - (((1 << 31) - 1) % n) + (n - 1) >= (1 << 31)
    s1 = (0x7FFFFFFF << 17) | 0x1ABCD  # top 31 bits set
    a, b, used = _shuffle_both(_state_before(s1), n)
    assert (a == b).all()
    assert used == n  # n - 1 swaps + the rejected value
