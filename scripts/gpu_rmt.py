# RunMultipleTimes on the device engine: sequential vs concurrent seeds (small networks are launch-bound)
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wittgenstein_b200 import (GSFSignature, GSFSignatureParameters, RunMultipleTimes, DoneAtStatGetter, MsgReceivedStatGetter)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
p = GSFSignature(GSFSignatureParameters(n, int(.85 * n), 4, 50, 20, 10, int(.1 * n), "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"))
import torch
ndev = torch.cuda.device_count()
for conc in (1, 8, 16):
    r = RunMultipleTimes(p, runs, 0, [DoneAtStatGetter(), MsgReceivedStatGetter()])
    t0 = time.time()
    devs = list(range(ndev)) if (ndev > 1 and conc > 1) else None
    res = r.run(lambda c: c.continue_if(), concurrency=conc, devices=devs)
    dt = time.time() - t0
    print("gpus=%d n=%d runs=%d concurrency=%2d: %.2f s wall, %.0f simulated-ms/s aggregate; doneAt %s msgReceived %s" % (
        (ndev if devs else 1), n, runs, conc, dt, sum(r.end_times) / dt, res[0], res[1]), flush=True)
