"""CasperIMD generality on the host build of the device logic (tests/emu; TEST INFRASTRUCTURE), bit for bit against the oracle:
* `randomOnTies` (CasperIMD.java:250-253): `network.rd.nextBoolean()` drawn inside `best()` when two branches have the same
  number of attestations — the draw's index in the network's one Random depends on every event processed before it in the
  millisecond (casperResolveTies);
* several blocks created in the same millisecond: `Block.id` follows the order of the creating events (casperRenumber).
The same cases run on the device in tests/test_gpu_zz_casper_generality.py."""
import pytest

from tests import emu_lib
from tests.oracle_lib import OracleCasper
from tests.parity import compare_casper
from wittgenstein_b200 import CasperIMD, CasperParemeters

TIE_CASES = [  # (cycle, producers, attesters per round), Byzantine delay, latency: every one of them draws on ties (checked below)
    ((2, 3, 6), 9000, None), ((2, 2, 9), 7000, None), ((3, 4, 5), 9000, "NetworkFixedLatency(100)"),
    ((2, 5, 4), 7000, "NetworkNoLatency"),  # every node receives the fork's block in the same millisecond: all tie in one pass
    ((2, 3, 6), 7000, "NetworkNoLatency"),
]


def lockstep(args, delay, kind, steps, step, api, expect_ties=False):
    p = CasperIMD(CasperParemeters(*args), _api=api)
    p.network().set_tunable("casper_votes", 60)
    o = OracleCasper(*args)
    p.init(delay, kind); o.init(delay, kind)
    plain = None
    if expect_ties:
        a2 = list(args); a2[1] = False
        plain = OracleCasper(*a2)
        plain.init(delay, kind)
    drew = False
    for _ in range(steps):
        assert p.network().run_ms(step) == o.run_ms(step)
        bad = compare_casper(p, o, f"t={o.time}")
        assert not bad, bad
        if plain is not None:
            plain.run_ms(step)
            drew = drew or plain.rng_state() != o.rng_state()
    assert not compare_casper(p, o, "end", atts=True)
    assert drew == expect_ties


@pytest.mark.parametrize("shape,delay,latency", TIE_CASES)
def test_casper_random_on_ties(shape, delay, latency):
    cyc, bp, apr = shape
    lockstep((cyc, True, bp, apr, 1000, 1, None, latency), delay, "WF", 50, 3000, emu_lib.api(), expect_ties=True)


@pytest.mark.parametrize("kind,delay", [("WF", 8000), ("plain", 16000), ("NS", 24000), ("SF", 8000)])
def test_casper_blocks_created_in_the_same_millisecond(kind, delay):
    # a Byzantine delay of whole slots puts the Byzantine producer's block in the millisecond of another producer's
    lockstep((4, False, 3, 8, 1000, 1, None, None), delay, kind, 50, 4000, emu_lib.api())
    lockstep((2, True, 3, 6, 1000, 1, None, None), delay, kind, 50, 4000, emu_lib.api())
