"""world_size-2 test of the replica mode on CPU (gloo): rank r runs seed r, the job's value is the units of all ranks over
the slowest rank's time, and each replica's state equals a single-process run of the same seed."""
import hashlib
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_replicas_over_gloo():
    from tests.oracle_lib import OracleGSF

    env = dict(os.environ, WTG_TEST_EMU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "replica_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    ranks = {}
    job = None
    dec = json.JSONDecoder()
    text = out.stdout
    pos = 0
    while True:  # the two ranks share stdout: their lines may interleave
        i = text.find("RANKLINE ", pos)
        if i < 0:
            break
        r, _ = dec.raw_decode(text[i + 9:])
        ranks[r["rank"]] = r
        pos = i + 9
    i = text.find("JOBLINE ")
    if i >= 0:
        job, _ = dec.raw_decode(text[i + 8:])
    assert sorted(ranks) == [0, 1] and job is not None
    assert ranks[0]["seed"] == 0 and ranks[1]["seed"] == 1
    assert ranks[0]["digest"] != ranks[1]["digest"]  # different seeds, different runs
    assert job["n_gpus"] == 2
    assert job["total_units"] == ranks[0]["units"] + ranks[1]["units"]
    assert abs(job["slowest"] - max(ranks[0]["secs"], ranks[1]["secs"])) < 1e-9
    assert abs(job["value"] - job["total_units"] / job["slowest"]) < 1e-6 * job["value"]
    assert job["slowest"] >= 0.3  # rank 1 slept: the job is as slow as its slowest replica
    # every replica is the bit-exact run of its seed
    for r in (0, 1):
        o = OracleGSF(64, 54, 3, 20, 10, 10, 6, "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter", seed=r)
        o.init()
        while o.time < ranks[r]["units"]:
            o.run_ms(10)
        d = hashlib.sha256(o.verified().tobytes() + o.counters().tobytes()).hexdigest()
        assert d == ranks[r]["digest"]
