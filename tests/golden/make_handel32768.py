"""Generates tests/golden/handel32768_config3.json: digests of the ORACLE's state for BASELINE config #3 (Handel 32 768 nodes,
8 192 suicide-Byzantine nodes, AwsRegionNetworkLatency) at t = 100 … CHECK_UNTIL ms with runMs(100) slicing.  The oracle needs
minutes of CPU per simulated second at this size once the Byzantine phase starts, so the vectors are produced offline (this script,
committed) and the -m gpu test replays the same run on the device and compares the digests.  Oracle-generated: regression value
plus parity GPU == oracle past the Byzantine phase; it does not pin the oracle on the reference."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.oracle_lib import OracleHandel  # noqa: E402
from tests.parity import handel_digests  # noqa: E402

N = 32768
ARGS = (N, int(N * 0.7425), 4, 50, 10, 20, 10, N // 4)
NB, NL = "AWS_SPEED=GAUSSIAN_TOR=0.00", "AwsRegionNetworkLatency"
CHECK_UNTIL = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "handel32768_config3.json")

o = OracleHandel(*ARGS, NB, NL, 0, True)
t0 = time.time()
o.init()
print("oracle init %.0f s" % (time.time() - t0), flush=True)
res = {"params": list(ARGS) + [NB, NL, 0, True, False], "slicing": "runMs(100)", "checkpoints": {}}
while o.time < CHECK_UNTIL:
    t1 = time.time()
    o.run_ms(100)
    res["checkpoints"][str(o.time)] = handel_digests(o, True)
    print("t=%d (+%.0f s)" % (o.time, time.time() - t1), flush=True)
    json.dump(res, open(OUT, "w"), indent=1)
