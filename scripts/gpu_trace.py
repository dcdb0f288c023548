"""Per-step trace of a GSF run on the GPU: device ms per runMs(100), event counts, capacities."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wittgenstein_b200 import GSFSignature, GSFSignatureParameters
n = int(sys.argv[1]); tmax = int(sys.argv[2]); step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
prof_from = int(sys.argv[4]) if len(sys.argv) > 4 else -1
p = GSFSignature(GSFSignatureParameters(n, int(.85*n), 4, 50, 20, 10, int(.1*n), "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"))
t0 = time.time(); p.init(); p.network().msgs_size(); print("init_s", time.time()-t0, flush=True)
net = p.network(); prev = net.stats()
import torch
print("mem GiB used", (torch.cuda.mem_get_info()[1]-torch.cuda.mem_get_info()[0])/2**30, flush=True)
tot = 0.0
while net.time < tmax:
    if net.time == prof_from: net.profile_enable(True)
    net.timer_start(); net.run_ms(step); ms = net.timer_stop_ms(); tot += ms
    st = net.stats(); d = {k: st[k]-prev[k] for k in ("deliveries","tasks","cond_runs","eval_entries","eval_words","updates","cycles","sends","send_words","events")}; prev = st
    sc = p.scalars()
    print(json.dumps({"t": net.time, "ms": round(ms,2), "us_per_tick": round(1000*ms/step,1), "evalE_per_tick": d["eval_entries"]//step, "evalW_per_tick": d["eval_words"]//step, "ev_per_tick": d["events"]//step, "maxq": st["max_queue"], "meanq": float(sc["to_verify"].mean()), "maxbucket": st["max_bucket"], "min_pool_free": st["min_pool_free"], "live": net.msgs_size(), "mean_card": float(sc["card"].mean()), "cont": p.continue_if()}), flush=True)
    if not p.continue_if(): break
print("total device ms", tot, "sim ms", net.time, "sim-ms/s", net.time/(tot/1000))
if prof_from >= 0: print(json.dumps(net.profile_read()))
print(json.dumps(net.stats()))
