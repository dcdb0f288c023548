"""world_size-2 test of the node-sharded engine's multi-process form on CPU (gloo): one shard per process, exchange-region
handles through torch.distributed, every shard's state equal to the oracle's rows for its id range (GSFSignature to
completion, CasperIMD through 80 s)."""
import hashlib
import json
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def dg(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_two_shards_over_gloo():
    from tests.oracle_lib import OracleCasper, OracleGSF

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    ranks = {}
    dec = json.JSONDecoder()
    pos = 0
    while True:
        i = out.stdout.find("RANKLINE ", pos)
        if i < 0:
            break
        r, _ = dec.raw_decode(out.stdout[i + 9:])
        ranks[r["rank"]] = r
        pos = i + 9
    assert sorted(ranks) == [0, 1]
    # GSFSignature: both shards stopped at the same time with the same rd state; rows of each id range against the oracle
    g0, g1 = ranks[0]["gsf"], ranks[1]["gsf"]
    assert g0["time"] == g1["time"] and g0["rng"] == g1["rng"] and g0["range"] == [0, 128] and g1["range"] == [128, 128]
    o = OracleGSF(256, 204, 4, 50, 20, 10, 25, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency")
    o.init()
    while o.time < g0["time"]:
        o.run_ms(10)
    assert not o.continue_if() and o.rng_state() == g0["rng"]
    assert o.msgs_live() == g0["msgs"] + g1["msgs"]
    for g in (g0, g1):
        lo, n = g["range"]
        assert g["digest"] == dg(o.verified()[lo:lo + n], o.counters()[:, lo:lo + n], o.scalars()["sig_checked"][lo:lo + n])
    # CasperIMD: 16 nodes on the ranges [0, 8) and [8, 16)
    c0, c1 = ranks[0]["casper"], ranks[1]["casper"]
    oc = OracleCasper(2, False, 3, 6, 1000, 1, None, None)
    oc.init(9000)
    for _ in range(20):
        oc.run_ms(4000)
    st = oc.node_state()
    assert c0["time"] == c1["time"] == oc.time and c0["rng"] == c1["rng"] == oc.rng_state()
    assert c0["heads"] == c1["heads"] == st["head"].tolist() and c0["blocks"] == c1["blocks"] == len(oc.blocks()["height"])
    assert oc.msgs_live() == c0["msgs"] + c1["msgs"]
    for c in (c0, c1):
        lo, n = c["range"]
        assert c["digest"] == dg(st["head"][lo:lo + n], st["atts_received"][lo:lo + n], st["att_hash"][lo:lo + n], oc.counters()[:, lo:lo + n])
