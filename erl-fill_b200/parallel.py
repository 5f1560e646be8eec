"""One process per GPU: the little host logic the multi-GPU path needs.

Rollout: envs are sharded contiguously by rank (global env id = rank * n_per_rank + i, which keys the Philox streams),
each rank keeps its own replay shard, and there is NO data-path collective.  Update: data-parallel — every rank computes
gradients on its local minibatch and the flat gradient buckets are averaged with one all-reduce each (critic bucket,
then actor bucket) between the *_GRAD and *_APPLY phases of the update kernels; optimiser state and target networks
are replicated and stay bit-identical because every rank applies the same reduced gradient.

Works with any torch.distributed backend: "nccl" on the GPUs (NVLink/NVSwitch), "gloo" in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def env_info():
    """(rank, world_size, local_rank) from the torchrun environment (1 process: 0, 1, 0)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init(backend="nccl", device=None):
    """Initialise the default process group when launched under torchrun; returns (rank, world_size, local_rank)."""
    rank, world, local = env_info()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def env_shard(n_total, rank, world):
    """Contiguous shard of n_total envs: (env_id_base, n_local).  The remainder goes to the lowest ranks."""
    base, rem = divmod(int(n_total), int(world))
    n_local = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, n_local


class GradSync:
    """In-place gradient averaging callback for DDPGLearner.update / VecTD3.train_step / VecSAC.update."""

    graph_safe = True      # only stream-ordered collectives: may be captured into the update's CUDA graph

    def __init__(self, world=None, group=None):
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.calls = 0
        self.elements = 0

    def __call__(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.mul_(1.0 / self.world)
        self.calls += 1
        self.elements += t.numel()
        return t


def reduce_scalar(x, op="max", device="cpu"):
    """max / sum of a host scalar over the ranks (timing: max over ranks; throughput: sum of units)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


def broadcast_params(tensors, src=0):
    """Make replicas identical at start-up (checkpoints loaded on rank 0 only)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
