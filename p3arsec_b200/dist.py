"""Multi-rank plumbing for bench.py: one process per GPU under torchrun, NO data-path collective.

The blackscholes Map shards into independent contiguous index ranges (the reference's static partition,
parsec-ff/pkgs/libs/fastflow/ff/parallel_for_internals.hpp:498-518), so ranks never exchange option data.
torch.distributed is used only for the barrier around the timed region and for the max-over-ranks of the
device times (NCCL on GPUs, gloo in the CPU tests).
"""
import os


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n, world, rank):
    """Contiguous shard of [0, n) for `rank`: the first n % world ranks take one extra element --
    the same rule bs_gpu_init applies across the devices of one context."""
    q, r = divmod(n, world)
    first = rank * q + min(rank, r)
    return first, q + (1 if rank < r else 0)


class Ranks:
    """Barrier + reductions over the launched ranks (a no-op group when world_size == 1)."""

    def __init__(self, backend=None, device=None):
        self.rank, self.local_rank, self.world = env_world()
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend=backend or "gloo", **kw)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, value, op_name):
        if self.dist is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op_name))
        return float(t.item())

    def max(self, value):
        return self._reduce(value, "MAX")

    def sum(self, value):
        return self._reduce(value, "SUM")

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
            self.dist = None
