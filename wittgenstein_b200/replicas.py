"""Multi-GPU plumbing of the replica mode: one process per GPU, rank r simulates seed r — what
RunMultipleTimes (core/RunMultipleTimes.java:44-48) does one seed after the other — with no data-path collective.
torch.distributed is used only for the barrier around the timed region and for the max / sum over ranks of the
per-rank timings and counts (NCCL on GPUs; gloo lets the same logic be tested on CPU with world_size 2)."""
import os


class Replicas:
    def __init__(self, backend="nccl"):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend
        self.dist = None
        self.device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist

            if backend == "nccl":
                torch.cuda.set_device(self.local)
                self.device = f"cuda:{self.local}"
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            else:
                dist.init_process_group(backend)
            self.dist = dist
            dist.barrier()
        os.environ["WTG_DEVICE"] = str(self.local)  # the engine of this process lives on its own GPU

    @property
    def seed(self):
        """network.rd.setSeed(i) of replica i (RunMultipleTimes.java:47)"""
        return self.rank

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x, op):
        if self.dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist is not None else x

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist is not None else x

    def throughput(self, units_this_rank, seconds_this_rank):
        """whole-job value: the units all ranks processed / the slowest rank's time"""
        return self.sum_over_ranks(units_this_rank) / self.max_over_ranks(seconds_this_rank)

    def parallelism(self):
        return "1 GPU" if self.world == 1 else f"{self.world} independent seeded replicas (no data-path collective)"

    def finalize(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None
