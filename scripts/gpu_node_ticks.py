"""runs GSF-131072 for 1700 ms (under ncu: per-launch durations of k_node_tasks around t = 1600)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wittgenstein_b200 import GSFSignature, GSFSignatureParameters
NB, NL = "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"
n = 131072
p = GSFSignature(GSFSignatureParameters(n, 0.85, 4, 50, 20, 10, 0.10, NB, NL)); p.init()
for _ in range(int(sys.argv[1]) // 10): p.network().run_ms(10)
print("done", p.network().time)
