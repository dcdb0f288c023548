"""Node-sharded GSFSignature on the device: G engines, one shard of the node ids each, exchanging items and
envelopes through peer stores — bit-exact against the oracle (= the unsharded engine).  With one GPU the shards share
it (separate streams); with several GPUs in the box each shard gets its own."""
import numpy as np
import pytest

from tests import parity
from tests.oracle_lib import OracleGSF

pytestmark = pytest.mark.gpu

AWS_NB, AWS_NL = "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"


def devices_for(world):
    import torch

    n = torch.cuda.device_count()
    return [r % n for r in range(world)]


def run_pair(n, world, until, step, seed=None, full_every=1, devices=None, to_completion=False):
    from wittgenstein_b200 import GSFSignatureParameters
    from wittgenstein_b200.sharded import ShardedGSFSignature

    prm = GSFSignatureParameters(n, 0.85, 4, 50, 20, 10, 0.1, AWS_NB, AWS_NL)
    p = ShardedGSFSignature(prm, world, devices=devices or devices_for(world))
    o = OracleGSF(n, prm.threshold, 4, 50, 20, 10, prm.nodes_down, AWS_NB, AWS_NL, seed=seed)
    if seed is not None:
        p.network().set_seed(seed)
    p.init()
    o.init()
    assert not parity.compare_init(p, o)
    i = 0
    while o.time < until and (not to_completion or o.continue_if()):
        assert p.network().run_ms(step) == o.run_ms(step)
        i += 1
        bad = parity.compare_gsf(p, o, f"t={o.time}", full=(i % full_every == 0))
        assert not bad, bad
    done = not p.continue_if()
    bad = parity.compare_gsf(p, o, "end", full=True)
    assert not bad, bad
    p.close()
    return done


@pytest.mark.parametrize("n,world,until,step,seed", [(256, 2, 400, 10, None), (1024, 4, 600, 10, 1), (512, 8, 300, 5, 2)])
def test_gsf_sharded_small(n, world, until, step, seed):
    run_pair(n, world, until, step, seed, full_every=4)


@pytest.mark.parametrize("world", [2, 4])
def test_gsf_4096_sharded_to_completion(world):
    """config #2 (GSFSignature 4 096 nodes) on 2 and 4 shards, to completion, every 10 ms against the oracle"""
    assert run_pair(4096, world, 4000, 10, None, full_every=10, to_completion=True)


def test_gsf_sharded_on_one_gpu_equals_unsharded_engine():
    from wittgenstein_b200 import GSFSignature, GSFSignatureParameters
    from wittgenstein_b200.sharded import ShardedGSFSignature

    prm = GSFSignatureParameters(2048, 0.85, 4, 50, 20, 10, 0.1, AWS_NB, AWS_NL)
    a = GSFSignature(prm)
    b = ShardedGSFSignature(prm, 4, devices=[0, 0, 0, 0])
    a.init(); b.init()
    for _ in range(60):
        a.network().run_ms(10); b.network().run_ms(10)
    assert (a.verified() == b.verified()).all()
    assert (a.network().counters() == b.network().counters()).all()
    assert a.network().rng_state() == b.network().rng_state()
    assert a.network().msgs_size() == b.network().msgs_size()
    b.close()


def test_gsf_wide_peer_ids():
    """32-bit absolute peer ids (the layout of runs with more than 131 072 nodes, config #5) forced at 1 024 nodes, 2 shards"""
    from wittgenstein_b200 import GSFSignatureParameters
    from wittgenstein_b200.sharded import ShardedGSFSignature

    prm = GSFSignatureParameters(1024, 0.85, 4, 50, 20, 10, 0.1, AWS_NB, AWS_NL)
    p = ShardedGSFSignature(prm, 2, devices=devices_for(2), tunables={"peer_bits_32": 1})
    o = OracleGSF(1024, prm.threshold, 4, 50, 20, 10, prm.nodes_down, AWS_NB, AWS_NL)
    p.init(); o.init()
    assert p.network().stats()["peer_bits"] == 32
    for i in range(50):
        assert p.network().run_ms(10) == o.run_ms(10)
        bad = parity.compare_gsf(p, o, f"t={o.time}", full=(i % 10 == 0))
        assert not bad, bad
    p.close()
