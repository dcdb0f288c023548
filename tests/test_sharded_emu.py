"""Node-sharded simulation, host logic: the same state-transition bodies and exchange protocol as the CUDA
engine, compiled for the host (tests/emu), G shards driven by G threads of this process, checked bit for bit against the
oracle (and hence against the unsharded engine).  The CUDA kernels of the exchange run under -m gpu."""
import numpy as np
import pytest

from tests import emu_lib, parity
from tests.oracle_lib import OracleGSF
from wittgenstein_b200 import GSFSignatureParameters
from wittgenstein_b200.sharded import ShardedGSFSignature

AWS_NB, AWS_NL = "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"


def run_pair(n, world, until, step, seed=None, nb=AWS_NB, nl=AWS_NL, accel=10, dead=0.1, full_every=1, hook=None):
    prm = GSFSignatureParameters(n, 0.8, 4, 50, 20, accel, dead, nb, nl)
    p = ShardedGSFSignature(prm, world, _api=emu_lib.api())
    o = OracleGSF(n, prm.threshold, 4, 50, 20, accel, prm.nodes_down, nb, nl, seed=seed)
    if seed is not None:
        p.network().set_seed(seed)
    p.init()
    o.init()
    assert not parity.compare_init(p, o)
    i = 0
    while o.time < until:
        if hook:
            hook(p, o)
        r1, r2 = p.network().run_ms(step), o.run_ms(step)
        assert r1 == r2
        i += 1
        bad = parity.compare_gsf(p, o, f"t={o.time}", full=(i % full_every == 0))
        assert not bad, bad
    p.close()


@pytest.mark.parametrize("n,world,until,step,seed", [(64, 2, 300, 1, None), (256, 2, 600, 10, None), (256, 4, 600, 7, 3),
                                                     (512, 8, 400, 10, 1), (1024, 4, 700, 10, None), (1024, 2, 200, 1, 4), (4096, 2, 150, 10, None)])
def test_gsf_sharded_vs_oracle(n, world, until, step, seed):
    run_pair(n, world, until, step, seed, full_every=3)


def test_gsf_sharded_random_positions_no_tor_no_accel():
    run_pair(512, 4, 500, 10, None, "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter", accel=0, dead=0.0, full_every=5)
    run_pair(256, 2, 400, 10, 7, "RANDOM_SPEED=CONSTANT_TOR=0.00", None, accel=3, dead=0.2, full_every=5)


def test_gsf_sharded_stop_start_partition():
    def hook(p, o):
        t = o.time
        live = np.flatnonzero(o.attrs()["down"] == 0) if t == 0 else None
        if t == 0:
            hook.a, hook.b = int(live[3]), int(live[-2])  # nodes of different shards, alive after init()
        if t == 100:
            p.network().stop_node(hook.a); o.stop_node(hook.a)
            p.network().stop_node(hook.b); o.stop_node(hook.b)
        if t == 200:
            p.network().partition(0.4); o.partition(0.4)
        if t == 300:
            p.network().end_partition(); o.end_partition()
            p.network().start_node(hook.a); o.start_node(hook.a)

    run_pair(256, 4, 500, 10, 2, hook=hook, full_every=5)


def test_gsf_wide_peer_ids_host_build():
    """the 32-bit absolute peer-id layout (used when N / 2 > 65 536, i.e. from 262 144 nodes on: BASELINE config #5) forced at a
    size the oracle reaches, unsharded and on 2 shards"""
    from wittgenstein_b200 import GSFSignature

    prm = GSFSignatureParameters(512, 0.8, 4, 50, 20, 10, 0.1, AWS_NB, AWS_NL)
    for world in (1, 2):
        if world == 1:
            p = GSFSignature(prm, _api=emu_lib.api(), tunables={"peer_bits_32": 1})
        else:
            p = ShardedGSFSignature(prm, world, _api=emu_lib.api(), tunables={"peer_bits_32": 1})
        o = OracleGSF(512, prm.threshold, 4, 50, 20, 10, prm.nodes_down, AWS_NB, AWS_NL)
        p.init(); o.init()
        assert p.network().stats()["peer_bits"] == 32
        assert not parity.compare_init(p, o)
        for i in range(40):
            p.network().run_ms(10); o.run_ms(10)
            bad = parity.compare_gsf(p, o, f"t={o.time}", full=(i % 8 == 0))
            assert not bad, bad


def test_gsf_config5_shape_8_shards_wide_peer_ids():
    """BASELINE config #5's shape (8 shards, 32-bit absolute peer ids as from 262 144 nodes on) at 2 048 nodes: 256 nodes per shard,
    the top three levels cross shards"""
    prm = GSFSignatureParameters(2048, 0.8, 4, 50, 20, 10, 0.1, AWS_NB, AWS_NL)
    p = ShardedGSFSignature(prm, 8, _api=emu_lib.api(), tunables={"peer_bits_32": 1})
    o = OracleGSF(2048, prm.threshold, 4, 50, 20, 10, prm.nodes_down, AWS_NB, AWS_NL)
    p.init(); o.init()
    assert p.network().stats()["peer_bits"] == 32
    assert not parity.compare_init(p, o)
    for i in range(45):
        assert p.network().run_ms(10) == o.run_ms(10)
        bad = parity.compare_gsf(p, o, f"t={o.time}", full=(i % 9 == 8))
        assert not bad, bad
    p.close()
