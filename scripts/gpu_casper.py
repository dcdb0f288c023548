# CasperIMD config #4: 64-slot cycles, 5 producers, 256 attesters per slot -> 16 390 nodes.
# usage: gpu_casper.py [total_ms] [oracle_ms]   (oracle_ms > 0: compare with the CPU oracle over the first oracle_ms)
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wittgenstein_b200 import CasperIMD, CasperParemeters
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1024000
oms = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cyc, bpc, apr = (int(x) for x in os.environ.get("CASPER_CFG", "64,5,256").split(","))
NB, NL = "RANDOM_SPEED=CONSTANT_TOR=0.00", "NetworkLatencyByDistanceWJitter"
t0 = time.time()
p = CasperIMD(CasperParemeters(cyc, False, bpc, apr, 1000, 1, NB, NL))
p.network().set_tunable("casper_votes", total // (8000 * cyc) + 3)
p.init(0)
net = p.network()
print("init %.2fs nodes %d" % (time.time() - t0, p.node_count()), flush=True)
if oms:
    from tests.oracle_lib import OracleCasper
    from tests.parity import compare_casper
    o = OracleCasper(cyc, False, bpc, apr, 1000, 1, NB, NL)
    o.init(0)
    tc = 0.0
    while o.time < oms:
        net.run_ms(4000)
        t1 = time.time(); o.run_ms(4000); tc += time.time() - t1
        bad = compare_casper(p, o, "t=%d" % o.time)
        if bad:
            print("MISMATCH", bad); sys.exit(1)
    print("oracle parity OK to t=%d; oracle %.1f s wall, %d deliveries (%.0f sim-ms/s)" % (o.time, tc, o.deliveries(), oms / tc), flush=True)
if os.environ.get("PROFILE"):
    net.profile_enable(True)
    for _ in range(5):
        net.run_ms(8000)
    prof = net.profile_read()
    net.profile_enable(False)
    tot = sum(v[0] for v in prof.values())
    print("per-kernel ms over 5 slots (total %.1f):" % tot, {k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1]}, flush=True)
    oms = net.time
net.timer_start()
t1 = time.time()
while net.time < total:
    net.run_ms(8000)
dev_ms = net.timer_stop_ms()
wall = time.time() - t1
done = total - oms
st = net.stats()
b = p.blocks()
ns = p.node_state()
print("ran [%d,%d] ms: device %.1f ms, wall %.2f s -> %.0f simulated-ms/s; deliveries %d (%.1f M/s); launches %d" % (
    oms, total, dev_ms, wall, done / (dev_ms / 1000.0), st["deliveries"], st["deliveries"] / dev_ms / 1e3, st["kernel_launches"]))
print("blocks %d, observer head height %d, heads distinct %d, atts received by observer %d, byz %s" % (
    len(b["height"]), b["height"][ns["head"][0]], len(set(ns["head"].tolist())), ns["atts_received"][0], p.byz()))
