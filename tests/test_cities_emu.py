"""NetworkLatencyByCity / ByCityWJitter and the CITIES node builder (core/NetworkLatency.java:159-233,
core/NodeBuilder.java:98-148, core/geoinfo/GeoAllCities.java, tools/CSVLatencyReader.java) — the reference's own structural tests
restated on the oracle, then the host build of the engine against the oracle (the CUDA path: tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from tests import emu_lib
from tests.oracle_lib import OracleHandel, OraclePingPong
from wittgenstein_b200 import Handel, HandelParameters, PingPong, PingPongParameters

CITIES_NB = "CITIES_SPEED=CONSTANT_TOR=0.00"


def test_city_tables_properties_of_the_reference_tests():
    """CT/CityPopulationTest.testCumulativeProbability (< 1.00001), CT/NetworkLatencyTest.testCitiesLatency (latency > 0 for every
    pair, 1 for the node itself), CT/CSVLatencyReaderTest.testLoad (some cities) — on the oracle's tables."""
    o = OraclePingPong(400, CITIES_NB, "NetworkLatencyByCity")
    o.init()
    a = o.attrs()
    assert (a["city"] >= 100).all() and len(set(a["city"].tolist())) > 50  # many distinct cities
    assert (a["x"] > 0).all() and (a["x"] <= 2000).all() and (a["y"] > 0).all() and (a["y"] <= 1112).all()
    for _ in range(20):
        o.run_ms(100)
    assert o.pongs()[0] == 400 and o.msgs_size() == 0  # every latency was positive and finite


@pytest.mark.parametrize("nb,nl", [(CITIES_NB, "NetworkLatencyByCityWJitter"), ("CITIES_SPEED=GAUSSIAN_TOR=0.33", "NetworkLatencyByCity")])
def test_pingpong_cities_host_build(nb, nl):
    p = PingPong(PingPongParameters(500, nb, nl), _api=emu_lib.api())
    o = OraclePingPong(500, nb, nl)
    p.init(); o.init()
    a, b = p.network().attrs(), o.attrs()
    for k in ("x", "y", "extra", "city", "down"):
        assert (a[k] == b[k]).all(), k
    for _ in range(25):
        assert p.network().run_ms(100) == o.run_ms(100)
        assert (p.pongs() == o.pongs()).all()
        assert (p.network().counters() == o.counters()).all()
        assert p.network().msgs_size() == o.msgs_size()
        assert p.network().rng_state() == o.rng_state()


def test_by_city_latency_needs_city_nodes():
    from wittgenstein_b200 import WtgError

    p = PingPong(PingPongParameters(10, None, "NetworkLatencyByCity"), _api=emu_lib.api())
    with pytest.raises(WtgError):
        p.init()  # "Can't use NetworkLatencyByCity model with default city location"


def handel_default_params(nodes):
    """HandelScenarios.defaultParams (protocols/HandelScenarios.java:65-120): CITIES builder (uniform speed, no Tor),
    NetworkLatencyByCityWJitter, 10 % dead, threshold 99 % of the live nodes, pairing 4, level wait 50, period 20, fast path 10."""
    dead = int(nodes * 0.10)
    thr = min(int(nodes * ((1.0 - 0.10) * .99)), nodes - dead)
    return (nodes, max(2, thr), 4, 50, 10, 20, 10, dead, "CITIES_SPEED=GAUSSIAN_TOR=0.00", "NetworkLatencyByCityWJitter", 0, False)


def test_handel_default_scenario_host_build():
    args = handel_default_params(128)
    p = Handel(HandelParameters(*args), _api=emu_lib.api())
    o = OracleHandel(*args)
    p.init(); o.init()
    for _ in range(100):
        assert p.network().run_ms(10) == o.run_ms(10)
    a, b = p.scalars(), o.scalars()
    assert all((a[x] == b[x]).all() for x in a)
    for w in range(6):
        assert (p.rows(w) == o.rows(w)).all()
    assert (p.network().counters() == o.counters()).all() and p.network().rng_state() == o.rng_state()
