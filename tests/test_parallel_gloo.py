"""Multi-process host logic on CPU: world_size 2 over gloo (no GPU, no CUDA library calls).

Covers what the N > 1 path adds on the host: contiguous env sharding (the global env ids that key the Philox streams are
disjoint and cover the whole box), the gradient-bucket averaging callback (the exchange step of the data-parallel
update), max/sum reductions used by bench.py's timing, and replica consistency after identical reduced updates."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import erlfill_b200  # noqa: F401
from erlfill_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port)})
    r, w, _ = P.init(backend="gloo")
    assert (r, w) == (rank, world)
    # 1. env sharding
    base, n_local = P.env_shard(1000003, rank, world)
    ids = torch.tensor([base, base + n_local], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, ids)
    # 2. gradient averaging == gradient of the concatenated batch for a mean loss (a linear model stands in for the nets)
    g = torch.Generator().manual_seed(100)
    X, y = torch.randn(64, 7, generator=g, dtype=torch.float64), torch.randn(64, generator=g, dtype=torch.float64)
    wgt = torch.randn(7, generator=g, dtype=torch.float64)
    lo, hi = rank * 32, (rank + 1) * 32
    local_grad = 2 * X[lo:hi].T @ (X[lo:hi] @ wgt - y[lo:hi]) / 32
    full_grad = 2 * X.T @ (X @ wgt - y) / 64
    sync = P.GradSync()
    bucket = local_grad.clone()
    sync(bucket)
    assert sync.calls == 1 and sync.elements == 7
    # 3. replicas stay identical: apply the reduced gradient on every rank, compare parameters bit for bit
    param = wgt - 0.1 * bucket
    params = [torch.zeros_like(param) for _ in range(world)]
    dist.all_gather(params, param)
    tmax = P.reduce_scalar(1.0 + rank, "max")
    tsum = P.reduce_scalar(10.0 * (rank + 1), "sum")
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), shards=torch.stack(gathered).numpy(), avg=bucket.numpy(), full=full_grad.numpy(),
             same=np.array(all(torch.equal(params[0], p) for p in params)), tmax=tmax, tsum=tsum)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        d = np.load(tmp_path / f"r{rank}.npz")
        shards = d["shards"]
        assert shards[0, 0] == 0 and shards[-1, 1] == 1000003
        assert all(shards[i, 1] == shards[i + 1, 0] for i in range(world - 1))          # contiguous, disjoint
        np.testing.assert_allclose(d["avg"], d["full"], rtol=1e-12, atol=1e-12)
        assert bool(d["same"]) and d["tmax"] == 2.0 and d["tsum"] == 30.0


def test_env_shard_properties():
    for n, w in ((65536, 8), (1000003, 8), (5, 8), (4096, 1)):
        parts = [P.env_shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(p[1] for p in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1


def test_gradsync_single_process_is_identity():
    t = torch.arange(5, dtype=torch.float32)
    s = P.GradSync(world=1)
    assert torch.equal(s(t.clone()), t) and s.calls == 1
