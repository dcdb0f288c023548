"""Worker of tests/test_sharded_gloo.py (TEST INFRASTRUCTURE): one node-id shard per process over gloo — the multi-process
form of the sharded engine that `bench.py --gpus N` uses under torchrun (DistributedGSFSignature / DistributedCasperIMD: the
handles of the exchange regions travel once through torch.distributed, the data path is stores into the peers' regions).
The simulation runs on the host build of the device logic (tests/emu), whose exchange regions are POSIX shared memory where
the CUDA backend uses CUDA IPC."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

from tests import emu_lib  # noqa: E402
from wittgenstein_b200 import CasperParemeters, GSFSignatureParameters  # noqa: E402
from wittgenstein_b200.sharded import DistributedCasperIMD, DistributedGSFSignature  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
api = emu_lib.api()


def dg(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(a.tobytes())
    return h.hexdigest()


# GSFSignature, 256 nodes: this rank owns ids [rank * 128, rank * 128 + 128)
prm = GSFSignatureParameters(256, 0.8, 4, 50, 20, 10, 0.1, "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency")
p = DistributedGSFSignature(prm, dist, rank, world, None, _api=api)
p.init()
steps = 0
while p.continue_if() and steps < 200:
    p.network().run_ms(10)
    steps += 1
net = p.network()
gsf = {"time": net.time, "rng": net.rng_state(), "range": list(net.shard_range()),
       "digest": dg(p.local.verified(), net.counters(), p.local.scalars()["sig_checked"]), "msgs": net.msgs_size()}
del p
# CasperIMD, 1 + 3 + 12 nodes on uneven ranges
c = DistributedCasperIMD(CasperParemeters(2, False, 3, 6, 1000, 1, None, None), dist, rank, world, None, tunables={"casper_votes": 12}, _api=api)
c.init(9000)
for _ in range(20):
    c.network().run_ms(4000)
net = c.network()
st = c.local.node_state()
casper = {"time": net.time, "rng": net.rng_state(), "range": list(net.shard_range()), "heads": c.all_heads().tolist(),
          "digest": dg(st["head"], st["atts_received"], st["att_hash"], net.counters()), "blocks": len(c.blocks()["height"]), "msgs": net.msgs_size()}
print("RANKLINE " + json.dumps({"rank": rank, "gsf": gsf, "casper": casper}), flush=True)
dist.barrier()
dist.destroy_process_group()
