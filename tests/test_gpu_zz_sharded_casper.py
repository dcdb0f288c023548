"""Node-sharded CasperIMD on the device (BASELINE config #4): G engines, one contiguous range of node ids each,
replicated block / attestation tables, sendAll records built on every shard, far-future calendar with ordering keys, the
per-pass "next event" minimum of the fast-forward — bit-exact against the oracle.  With one GPU the shards share it (separate
streams); with several GPUs in the box each shard gets its own.
(Named zz: written after the last GPU session of round 2 — the logic is checked on the host build by
tests/test_sharded_casper_emu.py; these tests run after the parity tests proper.)"""
import pytest

from tests.oracle_lib import OracleCasper
from tests.parity import compare_casper

pytestmark = pytest.mark.gpu

RANDOM_NB, DIST_NL = "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter"


def devices_for(world):
    import torch

    n = max(1, torch.cuda.device_count())
    return [r % n for r in range(world)]


def run_pair(args, world, steps, step, byz_delay=0, byz_kind="WF", votes=12, atts_end=True):
    from wittgenstein_b200 import CasperParemeters
    from wittgenstein_b200.sharded import ShardedCasperIMD

    p = ShardedCasperIMD(CasperParemeters(*args), world, devices=devices_for(world), tunables={"casper_votes": votes})
    o = OracleCasper(*args)
    p.init(byz_delay, byz_kind); o.init(byz_delay, byz_kind)
    for _ in range(steps):
        assert p.network().run_ms(step) == o.run_ms(step), f"runMs return at t={o.time}"
        bad = compare_casper(p, o, f"t={o.time}")
        assert not bad, bad
    assert not compare_casper(p, o, "end", atts=atts_end)
    p.close()


@pytest.mark.parametrize("world", [2, 4])
def test_casper_sharded_small(world):
    run_pair((2, False, 3, 6, 1000, 1, None, None), world, 60, 2000, byz_delay=9000)


def test_casper_sharded_uneven_ranges_aws_tor():
    run_pair((3, False, 3, 7, 1000, 1, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"), 2, 40, 3000, byz_delay=-3000)
    run_pair((4, False, 2, 5, 1000, 1, RANDOM_NB, DIST_NL), 4, 40, 4000)


@pytest.mark.parametrize("kind", ["plain", "SF", "NS"])
def test_casper_sharded_byzantine_producers(kind):
    run_pair((2, False, 3, 6, 1000, 1, None, None), 2, 40, 2000, byz_delay=0 if kind != "plain" else 3000, byz_kind=kind)


def test_casper_sharded_1030_nodes_two_cycles():
    """config #4's shape at 1 + 5 + 64 x 16 = 1 030 nodes on 4 shards, two cycles of 16 slots"""
    run_pair((16, False, 5, 64, 1000, 1, RANDOM_NB, DIST_NL), 4, 32, 8000, votes=4, atts_end=False)


def test_casper_16390_config4_sharded_4_equals_unsharded_engine():
    """BASELINE config #4 at full size (16 390 nodes) on 4 shards through 5 slots: same node state, blocks, counters and rd
    state as the unsharded engine (which test_casper_16390_config4_prefix_vs_oracle pins on the oracle)"""
    from wittgenstein_b200 import CasperIMD, CasperParemeters
    from wittgenstein_b200.sharded import ShardedCasperIMD

    prm = CasperParemeters(64, False, 5, 256, 1000, 1, RANDOM_NB, DIST_NL)
    a = CasperIMD(prm)
    b = ShardedCasperIMD(prm, 4, devices=devices_for(4))
    a.init(0); b.init(0)
    for _ in range(10):
        assert a.network().run_ms(4000) == b.network().run_ms(4000)
        sa, sb = a.node_state(), b.node_state()
        for k in sa:
            assert (sa[k] == sb[k]).all(), (k, a.network().time)
        assert (a.network().counters() == b.network().counters()).all()
        assert a.network().rng_state() == b.network().rng_state() and a.network().msgs_size() == b.network().msgs_size()
        ba, bb = a.blocks(), b.blocks()
        assert all((ba[k] == bb[k]).all() for k in ba)
    b.close()
