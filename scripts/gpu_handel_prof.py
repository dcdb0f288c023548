import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wittgenstein_b200 import Handel, HandelParameters
n = int(sys.argv[1]); prof = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = Handel(HandelParameters(n, int(n * 0.7425), 4, 50, 10, 20, 10, n // 4, "AWS_SPEED=GAUSSIAN_TOR=0.00", "AwsRegionNetworkLatency", 0, True, False))
t0 = time.time(); p.init(); print("init_s", time.time() - t0, flush=True)
net = p.network()
if prof: net.profile_enable(True)
dev = 0.0
while p.continue_if() and net.time < 10000:
    net.timer_start(); net.run_ms(100); ms = net.timer_stop_ms(); dev += ms
    print(net.time, round(ms, 1), flush=True)
print("sim", net.time, "device ms", dev, "sim-ms/s", net.time / (dev / 1000))
if prof: print(json.dumps({k: [round(v[0], 1), v[1]] for k, v in net.profile_read().items()}))
print(json.dumps(net.stats()))
