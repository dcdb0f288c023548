"""Latency models with multi-second arrivals (EthScanNetworkLatency up to 12 000 ms, MeasuredNetworkLatency, Fixed / Uniform
(8000)): the time ring stays at 4 096 buckets and arrivals >= 2 048 ms ahead go through the far-future calendar (host
build of the device logic against the oracle; the CUDA kernels run the same under -m gpu, tests/test_gpu_parity.py)."""
import pytest

from tests import emu_lib, parity
from tests.oracle_lib import OracleGSF, OraclePingPong
from wittgenstein_b200 import GSFSignature, GSFSignatureParameters, PingPong, PingPongParameters

MEASURED = ([20, 30, 25, 15, 10], [40, 300, 1500, 3000, 6000])  # a distribution with a multi-second tail


def pingpong_pair(latency, measured, api):
    p = PingPong(PingPongParameters(300, None, latency), _api=api)
    o = OraclePingPong(300, None, latency)
    if measured:
        p.network().set_network_latency_measured(*MEASURED)
        o.set_network_latency_measured(*MEASURED)
    p.init(); o.init()
    return p, o


def check_pingpong(p, o, steps, step_ms):
    for _ in range(steps):
        assert p.network().run_ms(step_ms) == o.run_ms(step_ms)
        assert (p.pongs() == o.pongs()).all()
        assert (p.network().counters() == o.counters()).all()
        assert p.network().msgs_size() == o.msgs_size()
        assert p.network().rng_state() == o.rng_state()


@pytest.mark.parametrize("latency,measured", [("EthScanNetworkLatency", False), (None, True), ("NetworkFixedLatency(8000)", False),
                                              ("NetworkUniformLatency(8000)", False)])
def test_pingpong_far_latencies_host_build(latency, measured):
    p, o = pingpong_pair(latency, measured, emu_lib.api())
    assert p.network().stats()["ring"] == 4096
    check_pingpong(p, o, 30, 1000)
    assert o.pongs()[0] > 0


@pytest.mark.parametrize("n", [128, 256])
def test_gsf_ethscan_host_build(n):
    args = (n, int(0.8 * n), 3, 20, 10, 10, n // 16, "RANDOM_SPEED=CONSTANT_TOR=0.00", "EthScanNetworkLatency")
    p = GSFSignature(GSFSignatureParameters(*args), _api=emu_lib.api())
    o = OracleGSF(*args)
    p.init(); o.init()
    for i in range(60):
        assert p.network().run_ms(100) == o.run_ms(100)
        bad = parity.compare_gsf(p, o, f"t={o.time}", full=(i % 6 == 0))
        assert not bad, bad
