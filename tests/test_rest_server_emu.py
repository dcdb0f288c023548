"""The REST façade (wittgenstein_b200/server.py) against the reference's own server test
(wserver/src/test/java/net/consensys/wittgenstein/server/ws/WServerTest.java: testGetProtocols :52-64, testBasicAllProtocols
:66-127, testGetProtocolParameters :129-141, testInitProtocol :211-228, testFullWorkflowOnDummyProtocol :230-274 — the dummy
protocol's two bare nodes become a 2-node PingPong whose initial Ping has been delivered) and against the oracle for the JSON
contents.  Host build of the engine (tests/emu; TEST INFRASTRUCTURE); the device run is tests/test_gpu_zz_rest_server.py."""
import pytest
from starlette.testclient import TestClient

from tests import emu_lib
from tests.oracle_lib import OraclePingPong
from wittgenstein_b200.server import PKG, Server, create_app


@pytest.fixture()
def client():
    return TestClient(create_app(Server(_api=emu_lib.api())))


def test_get_protocols(client):
    ps = client.get("/w/protocols").json()
    assert PKG + "PingPong" in ps and PKG + "GSFSignature" in ps and len(ps) == 6


def test_get_protocol_parameters(client):
    r = client.get("/w/protocols/" + PKG + "PingPong")
    assert r.status_code == 200 and r.json() == {"type": "PingPongParameters", "nodeCt": 1000, "nodeBuilderName": None, "networkLatencyName": None}
    g = client.get("/w/protocols/" + PKG + "GSFSignature").json()
    assert g["nodeCount"] == 1024 and g["threshold"] == 1013 and g["acceleratedCallsCount"] == 10  # GSFSignature.java:42-52
    assert client.get("/w/protocols/" + PKG + "Paxos").status_code == 404


def test_init_protocol(client):
    r = client.post("/w/network/init/" + PKG + "PingPong", json={"type": "PingPongParameters", "nodeCt": 123})
    assert r.status_code == 200
    nodes = client.get("/w/network/nodes").json()
    assert len(nodes) == 123 and [n["nodeId"] for n in nodes] == list(range(123))
    o = OraclePingPong(123, None, None)
    o.init()
    a = o.attrs()
    assert [n["x"] for n in nodes] == a["x"].tolist() and [n["y"] for n in nodes] == a["y"].tolist()
    assert client.get("/w/network/nodes/7").json() == nodes[7]
    assert client.get("/w/network/nodes/123").status_code == 404


def test_basic_all_protocols(client):
    small = {"nodeCount": 64, "threshold": 60, "nodeCt": 64}
    for p in client.get("/w/protocols").json():
        prm = client.get("/w/protocols/" + p).json()
        prm.update({k: v for k, v in small.items() if k in prm})
        assert client.post("/w/network/init/" + p, json=prm).status_code == 200, p
        assert len(client.get("/w/network/nodes").json()) != 0, p
        assert client.get("/w/network/messages").status_code == 200, p
        assert client.post("/w/network/runMs/20").status_code == 200, p
        assert client.get("/w/network/time").text == "20"


def test_full_workflow(client):
    assert client.get("/w/network/time").status_code == 400  # nothing initialised yet
    assert client.post("/w/network/init/" + PKG + "PingPong", json={"nodeCt": 2}).status_code == 200
    o = OraclePingPong(2, None, None)
    o.init()
    mis = client.get("/w/network/messages").json()
    tot, rows = o.peek_messages()
    assert len(mis) == tot == 2  # the initial Ping to both nodes
    assert [(m["from"], m["to"], m["sentAt"], m["arrivingAt"]) for m in mis] == list(zip(*(rows[k].tolist() for k in ("from", "to", "sent_at", "arriving_at"))))
    assert all(m["msg"]["type"] == "Ping" for m in mis)
    assert client.post("/w/network/runMs/1000").status_code == 200
    o.run_ms(1000)
    assert client.get("/w/network/messages").json() == []
    sm = {"from": 0, "to": [1], "sendTime": 1001, "delayBetweenSend": 0, "message": {"type": "Ping"}}
    assert client.post("/w/network/send/", json=sm).status_code == 200
    o.send(1, 0, 1, send_time=1001)
    mis = client.get("/w/network/messages").json()
    tot, rows = o.peek_messages()
    assert len(mis) == tot == 1 and mis[0]["arrivingAt"] == int(rows["arriving_at"][0]) and mis[0]["sentAt"] == 1001
    assert client.post("/w/network/runMs/9000").status_code == 200
    o.run_ms(9000)
    assert client.get("/w/network/messages").json() == []
    assert client.get("/w/network/time").text == "10000"
    nodes = client.get("/w/network/nodes").json()
    c = o.counters()
    assert [n["msgReceived"] for n in nodes] == c[0].tolist() and [n["msgSent"] for n in nodes] == c[1].tolist()
    assert [n["pong"] for n in nodes] == o.pongs().tolist()
    # stop / start, error mapping
    assert client.post("/w/network/nodes/1/stop").status_code == 200
    assert client.get("/w/network/nodes/1").json()["down"] is True
    assert client.post("/w/nodes/1/start").status_code == 200  # the reference's route (WServer.java:74)
    assert client.get("/w/network/nodes/1").json()["down"] is False
    assert client.post("/w/network/runMs/0").status_code == 400  # Network.java:319-321
    assert client.post("/w/network/send", json={"from": 0, "to": [1], "sendTime": 5, "message": {"type": "Ping"}}).status_code == 400  # sendTime <= time
    assert client.post("/w/network/send", json={"from": 0, "to": [1], "sendTime": 20000, "message": {"type": "Foo"}}).status_code == 400
    assert client.post("/w/network/nodes/1/external", json="http://localhost:1").status_code == 501
