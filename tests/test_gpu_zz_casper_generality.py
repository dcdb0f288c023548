"""Device run of tests/test_casper_ties_emu.py: CasperIMD with `randomOnTies` (the tie's rd.nextBoolean() inside best(),
CasperIMD.java:250-253) and with several blocks created in the same millisecond (Block.id order), bit for bit against the oracle
through the C ABI.  (Named zz: added after the last bench session of round 2; runs after the parity tests proper.)"""
import pytest

from tests.oracle_lib import OracleCasper
from tests.parity import compare_casper

pytestmark = pytest.mark.gpu


def lockstep(args, delay, kind, steps, step, expect_ties=False):
    from wittgenstein_b200 import CasperIMD, CasperParemeters

    p = CasperIMD(CasperParemeters(*args))
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


@pytest.mark.parametrize("shape,delay,latency", [((2, 3, 6), 9000, None), ((2, 2, 9), 7000, None), ((3, 4, 5), 9000, "NetworkFixedLatency(100)"),
                                                 ((2, 5, 4), 7000, "NetworkNoLatency"), ((2, 3, 6), 7000, "NetworkNoLatency")])
def test_casper_random_on_ties(shape, delay, latency):
    cyc, bp, apr = shape
    lockstep((cyc, True, bp, apr, 1000, 1, None, latency), delay, "WF", 50, 3000, expect_ties=True)


@pytest.mark.parametrize("kind,delay", [("WF", 8000), ("plain", 16000), ("NS", 24000), ("SF", 8000)])
def test_casper_blocks_created_in_the_same_millisecond(kind, delay):
    lockstep((4, False, 3, 8, 1000, 1, None, None), delay, kind, 50, 4000)
    lockstep((2, True, 3, 6, 1000, 1, None, None), delay, kind, 50, 4000)


def test_casper_default_parameters_with_ties_enabled_larger():
    """CasperParemeters() (randomOnTies = true) scaled up: 1 + 4 + 32 x 8 = 261 nodes, a late Byzantine producer forks the chain"""
    lockstep((8, True, 4, 32, 1000, 1, "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter"), 9000, "WF", 40, 4000)
