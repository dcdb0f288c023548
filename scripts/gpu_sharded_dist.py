"""torchrun check of the multi-process node-sharded engine: one rank per GPU, exchange regions mapped through CUDA IPC.
Every rank compares its shard with its slice of an oracle run (test infrastructure; small sizes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tests.oracle_lib import OracleGSF
    from wittgenstein_b200 import GSFSignatureParameters
    from wittgenstein_b200.sharded import DistributedGSFSignature

    nb, nl = "AWS_SPEED=GAUSSIAN_TOR=0.33", "AwsRegionNetworkLatency"
    for n, until in ((1024, 500), (4096, 1300)):
        prm = GSFSignatureParameters(n, 0.85, 4, 50, 20, 10, 0.1, nb, nl)
        p = DistributedGSFSignature(prm, dist, rank, world, local)
        p.init()
        o = OracleGSF(n, prm.threshold, 4, 50, 20, 10, prm.nodes_down, nb, nl)
        o.init()
        n0, nl_ = p.network().shard_range()
        sl = slice(n0, n0 + nl_)
        t = 0
        while t < until:
            p.network().run_ms(10); o.run_ms(10)
            t += 10
            if t % 100 == 0:
                assert p.network().rng_state() == o.rng_state(), f"rank {rank}: rd state t={t}"
                assert (p.network().counters() == o.counters()[:, sl]).all(), f"rank {rank}: counters t={t}"
                assert (p.local.verified() == o.verified()[sl]).all(), f"rank {rank}: verified t={t}"
                for w in (1, 2):
                    assert (p.local.rows(w) == o.level_rows(w)[sl]).all(), f"rank {rank}: rows {w} t={t}"
                s1, s2 = p.local.scalars(), o.scalars()
                for k in s1:
                    assert (s1[k] == s2[k][sl]).all(), f"rank {rank}: {k} t={t}"
        msgs = torch.tensor([p.network().msgs_size()], device=f"cuda:{local}")
        dist.all_reduce(msgs)
        assert int(msgs.item()) == o.msgs_live(), f"msgs.size {int(msgs.item())} vs {o.msgs_live()}"
        if rank == 0:
            print(f"sharded over {world} processes: GSF-{n} bit-exact vs oracle through {until} ms", flush=True)
        del p
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
