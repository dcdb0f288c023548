# RunMultipleTimes on the device engine: sequential vs concurrent seeds (small networks are launch-bound)
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wittgenstein_b200 import (GSFSignature, GSFSignatureParameters, RunMultipleTimes, DoneAtStatGetter, MsgReceivedStatGetter)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
p = GSFSignature(GSFSignatureParameters(n, int(.85 * n), 4, 50, 20, 10, int(.1 * n), "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"))
for conc in (1, 4, 8, 16):
    r = RunMultipleTimes(p, runs, 0, [DoneAtStatGetter(), MsgReceivedStatGetter()])
    t0 = time.time()
    res = r.run(lambda c: c.continue_if(), concurrency=conc)
    dt = time.time() - t0
    print("n=%d runs=%d concurrency=%2d: %.2f s wall, %.0f simulated-ms/s aggregate; doneAt %s msgReceived %s" % (
        n, runs, conc, dt, sum(r.end_times) / dt, res[0], res[1]), flush=True)
