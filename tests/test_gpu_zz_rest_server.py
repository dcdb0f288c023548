"""Device run of the peekMessages read-back (wtg_peek_messages) and of the REST façade: the same comparisons as
tests/test_peek_messages_emu.py / tests/test_rest_server_emu.py, on the CUDA engine through the C ABI.
(Named zz: added after the last GPU session of round 2, so it runs after the parity tests proper.)"""
import numpy as np
import pytest

from tests.oracle_lib import OracleGSF, OraclePingPong

pytestmark = pytest.mark.gpu


def same_rows(net, o, what):
    tg, g = net.peek_messages()
    to, r = o.peek_messages()
    assert tg == to, (what, tg, to)
    for k in ("from", "to", "sent_at", "arriving_at"):
        assert (g[k] == r[k]).all(), (what, k, np.nonzero(g[k] != r[k])[0][:5])
    assert ((g["kind"] >= 2) == (r["is_task"] != 0)).all(), what


def test_peek_messages_pingpong_and_gsf():
    from wittgenstein_b200 import GSFSignature, GSFSignatureParameters, PingPong, PingPongParameters

    p = PingPong(PingPongParameters(1000, None, None))
    o = OraclePingPong(1000, None, None)
    p.init(); o.init()
    same_rows(p.network(), o, "pingpong after init")
    for step in (3, 20, 40, 100):
        p.network().run_ms(step); o.run_ms(step)
        same_rows(p.network(), o, f"pingpong t={o.time}")
    args = (256, 0.8, 4, 50, 20, 10, 0.1, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency")
    prm = GSFSignatureParameters(*args)
    g = GSFSignature(prm)
    og = OracleGSF(256, prm.threshold, 4, 50, 20, 10, prm.nodes_down, args[7], args[8])
    g.init(); og.init()
    same_rows(g.network(), og, "gsf after init")
    for _ in range(20):
        g.network().run_ms(7); og.run_ms(7)
        same_rows(g.network(), og, f"gsf t={og.time}")


def test_rest_workflow_on_device():
    from starlette.testclient import TestClient

    from wittgenstein_b200.server import PKG, create_app

    client = TestClient(create_app())
    assert client.post("/w/network/init/" + PKG + "PingPong", json={"nodeCt": 123}).status_code == 200
    o = OraclePingPong(123, None, None)
    o.init()
    assert len(client.get("/w/network/nodes").json()) == 123
    mis = client.get("/w/network/messages").json()
    tot, rows = o.peek_messages()
    assert len(mis) == tot and [m["arrivingAt"] for m in mis] == rows["arriving_at"].tolist()
    assert client.post("/w/network/runMs/10000").status_code == 200
    o.run_ms(10000)
    assert client.get("/w/network/time").text == "10000" and client.get("/w/network/messages").json() == []
    nodes = client.get("/w/network/nodes").json()
    assert [n["pong"] for n in nodes] == o.pongs().tolist() and [n["msgReceived"] for n in nodes] == o.counters()[0].tolist()
    for p in client.get("/w/protocols").json():
        prm = client.get("/w/protocols/" + p).json()
        prm.update({k: 64 for k in ("nodeCount", "nodeCt") if k in prm})
        if "threshold" in prm:
            prm["threshold"] = 60
        assert client.post("/w/network/init/" + p, json=prm).status_code == 200, p
        assert client.post("/w/network/runMs/50").status_code == 200, p
        assert client.get("/w/network/messages").status_code == 200, p


def test_pingpong_caller_sends_to_many_destinations():
    """network.send(msg, from, dests) with 300 / 40 destinations (more than any handler uses), also with delaysBetweenMessage"""
    from wittgenstein_b200 import PingPong, PingPongParameters

    p = PingPong(PingPongParameters(400, None, None))
    o = OraclePingPong(400, None, None)
    p.init(); o.init()
    p.network().run_ms(300); o.run_ms(300)
    for i in (5, 77, 399):
        p.network().stop_node(i); o.stop_node(i)
    d1 = [(7 * k + 3) % 400 for k in range(300)]
    p.network().send(1, 2, d1); o.send(1, 2, d1)
    d2 = list(range(399, 359, -1))
    p.network().send(1, 9, d2, send_time=p.network().time + 10, delay_between=3)
    o.send(1, 9, d2, send_time=o.time + 10, delay_between=3)
    same_rows(p.network(), o, "after the caller's sends")
    for _ in range(12):
        assert p.network().run_ms(50) == o.run_ms(50)
        assert (p.pongs() == o.pongs()).all() and (p.network().counters() == o.counters()).all()
    assert p.network().rng_state() == o.rng_state() and p.network().msgs_size() == o.msgs_size() == 0

