"""Generates tests/golden/*.json: end-state digests of seeded runs, produced by the CPU oracle.

The reference is Java and cannot run in this environment (no JVM), so these vectors are NOT outputs of the
reference itself: they pin the oracle (regression) and give the GPU engine a fixed target that does not depend on
the oracle being rebuilt.  Regenerate with:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.oracle_lib import OracleCappos, OracleCasper, OracleGSF, OracleHandel, OraclePingPong, OracleSanFermin  # noqa: E402

AWS_NB, AWS_NL = "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"
NB, NL = "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter"


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


CASES = {
    "pingpong_1000": dict(kind="pingpong", args=[1000, None, None], steps=[100] * 10),
    "gsf_32": dict(kind="gsf", args=[32, 32, 3, 20, 10, 10, 0, NB, NL], steps=[1] * 300),
    "gsf_256_aws_tor": dict(kind="gsf", args=[256, 204, 4, 50, 20, 10, 25, AWS_NB, AWS_NL], steps=[7] * 300),
    "gsf_1024_aws_tor": dict(kind="gsf", args=[1024, 870, 4, 50, 20, 10, 102, AWS_NB, AWS_NL], steps=[10] * 200),
    "sanfermin_1024": dict(kind="sanfermin", args=[1024, 1024, 2, 48, 300, 1, None, None], steps=[10] * 400),
    "handel_64_desync": dict(kind="handel", args=[64, 60, 6, 10, 5, 5, 10, 2, NB, NL, 100, False], steps=[1] * 1200),
    "handel_256_byz": dict(kind="handel", args=[256, 180, 4, 50, 10, 20, 10, 64, "AWS_SPEED=GAUSSIAN_TOR=0.00", AWS_NL, 0, True], steps=[10] * 300),
    "handel_256_hidden": dict(kind="handel", args=[256, 180, 4, 50, 10, 20, 10, 64, "AWS_SPEED=GAUSSIAN_TOR=0.00", AWS_NL, 0, False], hidden=True, steps=[10] * 300),
    "cappos_512_k50": dict(kind="cappos", args=[512, 256, 2, 48, 150, 50, None, None], steps=[10] * 400),
    # casper args: cycleLength, randomOnTies, producers, attestersPerRound, blockTime, attestationTime, builder, latency ; byzDelay
    "casper_3x20_forks": dict(kind="casper", args=[3, False, 3, 20, 1000, 1, None, None], delay=9000, steps=[500] * 400),
    "casper_4x16_aws_late": dict(kind="casper", args=[4, False, 2, 16, 1000, 1, "AWS_SPEED=GAUSSIAN_TOR=0.33", AWS_NL], delay=-7000, steps=[1000] * 200),
}


def run_case(c, make):
    p = make(c["kind"], c["args"], hidden=c.get("hidden", False))
    if c["kind"] == "casper":
        p.init(c["delay"])
    else:
        p.init()
    for s in c["steps"]:
        p.run_ms(s)
    return p


def state_digest(kind, p, net=None):
    """p exposes the same read-back names for oracle and engine wrappers (see tests/test_golden.py)."""
    counters = p.counters() if net is None else net.counters()
    if kind == "pingpong":
        return digest(counters, p.pongs())
    if kind == "gsf":
        sc = p.scalars()
        return digest(counters, p.verified(), sc["sig_checked"], sc["sig_queue_size"], sc["to_verify"])
    if kind == "sanfermin":
        sc = p.scalars()
        return digest(counters, sc["agg"], sc["cpl"], sc["done"], sc["sent_requests"], sc["received_requests"], sc["threshold_at"])
    if kind == "handel":
        sc = p.scalars()
        return digest(counters, p.rows(0), p.rows(1), p.rows(2), p.rows(5), sc["sigs_checked"], sc["sig_queue_size"], sc["msg_filtered"], sc["window"])
    if kind == "cappos":
        sc = p.scalars()
        return digest(counters, sc["cpl"], sc["sigs"], sc["done"], sc["threshold_done"], sc["swapping"], sc["cache_mask"], sc["threshold_at"])
    if kind == "casper":
        st, b = p.node_state(), p.blocks()
        atts = [np.array(p.block_attestations(i), np.int32).reshape(-1, 2) for i in range(1, len(b["height"]))]
        return digest(counters, st["head"], st["atts_received"], st["heads_with_atts"], st["blocks_received"], st["to_reevaluate"],
                      st["att_hash"], b["height"], b["parent"], b["producer"], b["proposal_time"], b["included"], *atts)
    raise ValueError(kind)


def make_oracle(kind, args, hidden=False):
    if hidden:
        return OracleHandel(*args, hidden_byzantine=True)
    return {"pingpong": OraclePingPong, "gsf": OracleGSF, "sanfermin": OracleSanFermin, "handel": OracleHandel,
            "casper": OracleCasper, "cappos": OracleCappos}[kind](*args)


if __name__ == "__main__":
    out = {}
    for name, c in CASES.items():
        p = run_case(c, make_oracle)
        out[name] = {"kind": c["kind"], "args": c["args"], "steps": c["steps"], "time": p.time, "digest": state_digest(c["kind"], p)}
        if "delay" in c:
            out[name]["delay"] = c["delay"]
        if c.get("hidden"):
            out[name]["hidden"] = True
        print(name, out[name]["time"], out[name]["digest"][:16])
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "end_states.json"), "w"), indent=1)
