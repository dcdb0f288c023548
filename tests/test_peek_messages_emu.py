"""network.msgs.peekMessages() (core/Network.java:279-286; the REST façade's GET /w/network/messages) through the C ABI
(wtg_peek_messages) against the oracle's restatement of MessageStorage.peekMessages / Envelope.infos: every pending arrival
— each remaining destination of a multi-destination envelope — with from / to / sentAt / arrivingAt and the Task flag.
Runs on the host build of the engine (tests/emu; TEST INFRASTRUCTURE); the same comparison runs on the device in
tests/test_gpu_zz_rest_server.py."""
import numpy as np
import pytest

from tests import emu_lib
from tests.oracle_lib import OracleGSF, OraclePingPong


@pytest.fixture(scope="module")
def api():
    return emu_lib.api()


def same_rows(net, o, what):
    tg, g = net.peek_messages()
    to, r = o.peek_messages()
    assert tg == to == net.msgs_size_rows(), (what, tg, to)
    for k in ("from", "to", "sent_at", "arriving_at"):
        assert (g[k] == r[k]).all(), (what, k, np.nonzero(g[k] != r[k])[0][:5])
    assert ((g["kind"] >= 2) == (r["is_task"] != 0)).all(), what


def test_pingpong_rows(api):
    from wittgenstein_b200 import PingPong, PingPongParameters

    p = PingPong(PingPongParameters(200, None, None), _api=api)
    o = OraclePingPong(200, None, None)
    p.init(); o.init()
    p.network().msgs_size_rows = lambda: p.network().peek_messages(cap=0)[0]
    same_rows(p.network(), o, "after init")      # one 200-destination envelope: 200 rows
    for step in (3, 20, 40, 100):
        p.network().run_ms(step); o.run_ms(step)
        same_rows(p.network(), o, f"t={o.time}")  # the rest of the Ping envelope + single-destination Pongs
    p.network().send(1, 3, [5, 6, 7], send_time=p.network().time + 4, delay_between=5)
    o.send(1, 3, [5, 6, 7], send_time=o.time + 4, delay_between=5)
    same_rows(p.network(), o, "delayed multi-send")


def test_gsf_rows(api):
    from wittgenstein_b200 import GSFSignature, GSFSignatureParameters

    args = (64, 52, 3, 20, 10, 10, 6, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency")
    p = GSFSignature(GSFSignatureParameters(*args), _api=api)
    o = OracleGSF(*args)
    p.init(); o.init()
    p.network().msgs_size_rows = lambda: p.network().peek_messages(cap=0)[0]
    same_rows(p.network(), o, "after init")  # periodic tasks registered by init(): sentAt 0
    for _ in range(12):
        p.network().run_ms(7); o.run_ms(7)
        same_rows(p.network(), o, f"t={o.time}")  # messages, accelerated multi-sends, update tasks, periodic re-arms
    t, rows = p.network().peek_messages(cap=5)
    assert t > 5 and len(rows["from"]) == 5


def test_rows_of_node_sharded_networks(api):
    """every pending arrival of a node-sharded network is reported by exactly one shard: GSFSignature (records copied to the
    shards of the next group) and CasperIMD (replicated sendAll records, far-future calendar)"""
    from tests.oracle_lib import OracleCasper
    from wittgenstein_b200 import CasperParemeters, GSFSignatureParameters
    from wittgenstein_b200.sharded import ShardedCasperIMD, ShardedGSFSignature

    prm = GSFSignatureParameters(128, 0.8, 4, 50, 20, 10, 0.1, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency")
    p = ShardedGSFSignature(prm, 4, _api=api)
    o = OracleGSF(128, prm.threshold, 4, 50, 20, 10, prm.nodes_down, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency")
    p.init(); o.init()
    p.network().msgs_size_rows = lambda: p.network().peek_messages(cap=0)[0]
    for _ in range(25):
        p.network().run_ms(9); o.run_ms(9)
        same_rows(p.network(), o, f"gsf t={o.time}")
    p.close()
    args = (2, False, 3, 6, 1000, 1, None, None)
    c = ShardedCasperIMD(CasperParemeters(*args), 4, _api=api, tunables={"casper_votes": 12})
    oc = OracleCasper(*args)
    c.init(9000); oc.init(9000)
    c.network().msgs_size_rows = lambda: c.network().peek_messages(cap=0)[0]
    same_rows(c.network(), oc, "casper after init")
    for k in range(60):
        step = 4000 if k % 3 else 37  # stop in the middle of the arrivals of a vote / a block as well
        c.network().run_ms(step); oc.run_ms(step)
        same_rows(c.network(), oc, f"casper t={oc.time}")
    c.close()
