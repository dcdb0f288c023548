"""persistent window kernel vs kernel-per-stage graph at several sizes (device-timed, same run)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wittgenstein_b200 import GSFSignature, GSFSignatureParameters
NB, NL = "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"
for n in [int(x) for x in sys.argv[1:]]:
    p = GSFSignature(GSFSignatureParameters(n, 0.85, 4, 50, 20, 10, 0.10, NB, NL)); p.init()
    net = p.network(); net.run_ms(50); 
    net.timer_start()
    for _ in range(10): net.run_ms(100)
    ms = net.timer_stop_ms()
    print(f"n={n} persist={'0' if os.environ.get('WTG_NO_PERSIST')=='1' else '1'} bps={os.environ.get('WTG_RUN_BLOCKS_PER_SM','max')}: {1000/(ms/1000):.0f} sim-ms/s ({ms/1000:.3f} ms/tick)", flush=True)
