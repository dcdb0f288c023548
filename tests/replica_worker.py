"""Worker of tests/test_replicas_gloo.py (TEST INFRASTRUCTURE): one replica per process over gloo.  The simulation itself
runs on the host debugging build of the device logic (tests/emu) because the product engine needs a GPU; what is under
test is the replica plumbing: seed per rank, barrier, max-over-ranks timing, whole-job aggregation."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import emu_lib  # noqa: E402
from wittgenstein_b200 import GSFSignature, GSFSignatureParameters  # noqa: E402
from wittgenstein_b200.replicas import Replicas  # noqa: E402

rep = Replicas("gloo")
p = GSFSignature(GSFSignatureParameters(64, 54, 3, 20, 10, 10, 6, "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter"),
                 _api=emu_lib.api())
p.network().set_seed(rep.seed)
p.init()
rep.barrier()
t0 = time.perf_counter()
steps = 0
while p.continue_if() and steps < 400:
    p.network().run_ms(10)
    steps += 1
if rep.rank == 1:
    time.sleep(0.3)  # the slower rank sets the job's time
secs = time.perf_counter() - t0
rep.barrier()
units = p.network().time
digest = hashlib.sha256(p.verified().tobytes() + p.network().counters().tobytes()).hexdigest()
value = rep.throughput(units, secs)
total_units = rep.sum_over_ranks(units)
slowest = rep.max_over_ranks(secs)
print("RANKLINE " + json.dumps({"rank": rep.rank, "seed": rep.seed, "units": units, "secs": secs, "digest": digest}), flush=True)
if rep.rank == 0:
    print("JOBLINE " + json.dumps({"n_gpus": rep.world, "value": value, "total_units": total_units, "slowest": slowest,
                                   "parallelism": rep.parallelism()}), flush=True)
rep.finalize()
