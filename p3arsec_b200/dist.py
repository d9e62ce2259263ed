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


def bind_to_gpu_locality(gpu_index):
    """Pin this process to the CPUs that share a NUMA node with GPU `gpu_index` (NVML's ideal CPU affinity), so that the
    staging buffers it first-touches -- and pins -- live in the memory closest to that GPU's PCIe root.  With eight
    ranks each copying 280 MB per step, remote-node staging memory is what limits the end-to-end rate.  Returns the
    CPU list, or None when NVML or the topology is unavailable (then nothing is changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        pynvml.nvmlShutdown()
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


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
