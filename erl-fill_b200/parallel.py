"""One process per GPU: the little host logic the multi-GPU path needs.

Rollout: envs are sharded contiguously by rank (global env id = rank * n_per_rank + i, which keys the Philox streams),
each rank keeps its own replay shard, and there is NO data-path collective.  Update: data-parallel — every rank computes
gradients on its local minibatch and the flat gradient buckets are averaged with one all-reduce each (critic bucket,
then actor bucket) between the *_GRAD and *_APPLY phases of the update kernels; optimiser state and target networks
are replicated and stay bit-identical because every rank applies the same reduced gradient.

Works with any torch.distributed backend: "nccl" on the GPUs (NVLink/NVSwitch), "gloo" in the CPU tests.

PeerExchange replaces the two NCCL all-reduces of the ER-DDPG update by the one-shot peer-memory exchange of csrc/peer.cu
(every rank reads every rank's bucket over NVLink in rank order, fused with the clip-norm partials); torch.distributed only
carries the 64-byte IPC handles at start-up.
"""
import ctypes as C
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


class _DevMem:
    """Raw device memory as a torch tensor (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerExchange:
    """Gradient exchange of DDPGLearner.update over NVLink peer memory (csrc/peer.cu) for the ranks of ONE node.

    Construction is collective: every rank allocates its update workspace (+ control block) with erlfill_peer_alloc, the IPC
    handles travel through torch.distributed.all_gather_object, and every rank maps the others' blocks.  Afterwards
    `learner.update(..., allreduce=exchange)` runs  CRITIC_GRAD | exchange(0) | CRITIC_APPLY+REDUCED | ACTOR_GRAD | exchange(1) |
    ACTOR_APPLY+REDUCED  with no host synchronisation and no NCCL call; the whole sequence is CUDA-graph capturable."""

    graph_safe = True
    peer = True

    def __init__(self, learner, batch, group=None):
        from . import _lib as L
        self.L, self.group = L, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > L.MAX_PEERS:
            raise L.ErlfillError("PeerExchange: at most %d ranks (one node)" % L.MAX_PEERS)
        self.batch = int(batch)
        dev = learner.device
        ws_bytes = (L.lib().erlfill_ddpg_workspace_bytes(C.c_int(self.batch)) + 255) // 256 * 256
        total = ws_bytes + L.lib().erlfill_peer_ctrl_bytes()
        handle = (C.c_ubyte * 64)()
        base = C.c_void_p()
        with torch.cuda.device(dev):
            L.check(L.lib().erlfill_peer_alloc(C.c_size_t(total), C.byref(base), handle), "erlfill_peer_alloc")
        self._base, self._mapped = base.value, []
        infos = [None] * self.world
        dist.all_gather_object(infos, (dev.index if dev.index is not None else torch.cuda.current_device(), bytes(handle)), group=group)
        g = L.PeerGroup()
        g.world, g.rank = self.world, self.rank
        for p, (pdev, h) in enumerate(infos):
            if p == self.rank:
                ptr = self._base
            else:
                m = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(h)
                with torch.cuda.device(dev):
                    L.check(L.lib().erlfill_peer_open(C.c_int(pdev), hb, C.byref(m)), "erlfill_peer_open")
                ptr = m.value
                self._mapped.append(ptr)
            g.ws[p] = ptr
            g.ctrl[p] = ptr + ws_bytes
        self.g = g
        # the learner's workspace IS the shared block (peers read its gradient buckets in place)
        self._holder = _DevMem(self._base, ws_bytes)
        learner._ws = torch.as_tensor(self._holder, device=dev)
        learner._ws_batch = self.batch
        learner._ws_shared = True
        self._learner = learner
        learner._peer_exchange = self
        learner._graphs.clear()
        learner._warm.clear()
        learner._images_ok = False       # the fast path's weight images live in the (new) workspace
        dist.barrier(group=group)        # every rank has mapped every block before the first exchange

    def exchange(self, which):
        L = self.L
        if getattr(self._learner, "precision", "exact") == "fast":      # the X_GRAD phase has signalled already (hyper.peer)
            L.check(L.lib().erlfill_ddpg_peer_allreduce_fast(C.byref(self.g), C.c_int(self.batch), C.c_int(which), L.stream_ptr()),
                    "erlfill_ddpg_peer_allreduce_fast")
        else:
            L.check(L.lib().erlfill_ddpg_peer_allreduce(C.byref(self.g), C.c_int(self.batch), C.c_int(which), L.stream_ptr()),
                    "erlfill_ddpg_peer_allreduce")

    def error(self):
        e = C.c_int()
        self.L.check(self.L.lib().erlfill_peer_error(C.byref(self.g), C.byref(e)), "erlfill_peer_error")
        return e.value

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for m in self._mapped:
            self.L.lib().erlfill_peer_close(C.c_void_p(m))
        self._mapped = []
        dist.barrier(group=self.group)
        if self._base:
            ln = getattr(self, "_learner", None)
            if ln is not None:          # the learner must not keep pointing at the freed block
                ln._ws, ln._ws_batch, ln._ws_shared = None, 0, False
                ln._peer_exchange = None
                ln._graphs.clear(); ln._warm.clear()
            self.L.lib().erlfill_peer_free(C.c_void_p(self._base))
            self._base = None
