#!/usr/bin/env python
"""bench.py — throughput of the ERL-Fill hot path on B200 (contract: DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sim|rollout|update] [--n-env 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU port of the reference path on the host cores

One JSON line on stdout (rank 0).  A "step" is one lock-step pass of the workload over the batch of every rank:
  sim      configs[1] of BASELINE.json: batched WeightEnv rollout, PID-effective constant action (default, 4096 envs/GPU)
  rollout  configs[2]: MultiConditionWeightEnv, EmotionModule transformer -> ER-DDPG actor + OU -> env step -> emotion update
  update   configs[3]: ER-DDPG update, minibatch 4096 from a 1M-transition replay (sample + gather + U1-U5)
  td3|sac  configs[4]: full TD3 / SAC loop (act -> env step -> remember -> update of B=4096), 131,072 envs/GPU
  esac     EmotionSAC (ERL_Fill_agent.py) full loop incl. the emotion transformer, 65,536 envs/GPU
  erddpg   the whole ER-DDPG loop: rollout (with replay insert / replay-branch resets) + one update of B=4096 per lock-step step
  replay   rows R1/R2 alone against the HBM roofline; n1: configs[0], the N = 1 drop-in loop
  all      (default) the sim line, with the rollout and update results attached under "other_workloads"
`value` is device-timed with inputs resident in HBM; `e2e` goes through the public Python API with pinned HOST
buffers (H2D of the step's inputs, D2H of its results inside the timed region).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENV_STEP_BYTES = 192            # algorithmic bytes per env-step
TRANSFORMER_FLOP = 21.83e6      # per env per call at S=20
# what emotion_forward_tc_kernel executes: layers 0-2 in full (3 x 5,457,920), layer 3 q|k|v + attention for 20 tokens
# (174,080) and out_proj + FFN for the last token only (264,192)
TRANSFORMER_FLOP_EXECUTED = 3 * 5457920 + 174080 + 264192
UPDATE_FLOP_PER_SAMPLE = 1.43e6
L2_FLUSH_BYTES = 256 << 20      # > 126 MB L2
# warp-instructions one env-step of the `sim` workload issues (ncu smsp__inst_executed.sum / n_env): the fused
# env_sim_warp_kernel<true> when the batch is resident at once (profiles/r01d_env_sim_warp_4096.txt), else
# env_sim_warp_kernel<false> + env_finish_kernel (profiles/r01d_env_sim_warp_65536.txt)
SIM_WARP_INSTR_PER_ENV_STEP = {("pid", True): 3620.3, ("pid", False): 3074.6 + 20.4}
SIM_FUSED_MAX_ENVS = 148 * 32   # csrc/env_step.cu: resident_warp_slots()
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of exactly these configurations
# (cold L2, one launch): {(workload, envs per GPU): (bytes, summary file)}; any other configuration reports traffic null
NCU_TRAFFIC = {("sim", 4096): (280064, "profiles/r01d_env_sim_warp_4096.txt"),
               ("sim", 65536): (2662400 + 4749824, "profiles/r01d_env_sim_warp_65536.txt (simulator + finish kernel)"),
               ("rollout", 65536): (17171968, "profiles/r01b_emotion_tc.txt")}


def _traffic(workload, n):
    t = NCU_TRAFFIC.get((workload, n))
    return (t[0], t[1]) if t else (None, None)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# timing helpers
# ------------------------------------------------------------------------------------------------
class Timer:
    """K steps, each bracketed by its own CUDA-event pair on the launching stream with an L2 flush in between
    (outside the pair); barrier + synchronize on both sides; max over ranks."""

    def __init__(self, device, world):
        import torch
        self.torch, self.device, self.world = torch, device, world
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=device)
        self.stream = torch.cuda.current_stream()

    def sync(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return float(x)
        import torch.distributed as dist
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def maxr(self, x):
        import torch.distributed as dist
        return self._red(x, dist.ReduceOp.MAX)

    def sumr(self, x):
        import torch.distributed as dist
        return self._red(x, dist.ReduceOp.SUM)

    def run(self, fn, steps, flush=True):
        torch = self.torch
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        self.sync()
        w0 = time.perf_counter()
        for k in range(steps):
            if flush:
                self.flush.zero_()
            starts[k].record(self.stream)
            fn()
            ends[k].record(self.stream)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        self.sync()
        total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
        return self.maxr(total_ms), w1 - w0

    def run_block(self, fn, steps):
        """K back-to-back calls inside one event pair (host work of fn included): the e2e region."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        torch.cuda.synchronize()
        self.sync()
        return self.maxr(e0.elapsed_time(e1))


def _kernel_ms(timer, fn, reps=5):
    """Average device time of one call of fn (an L2 flush before every call, outside the event pair)."""
    ms, _ = timer.run(fn, reps)
    return ms / reps


# ------------------------------------------------------------------------------------------------
# workloads (each returns the pieces of the JSON line)
# ------------------------------------------------------------------------------------------------
def wl_sim(args, rank, world, device, timer):
    import torch
    import erlfill_b200 as E
    K = E.conditions
    n = args.n_env
    env = E.VecFillEnv([K.SINGLE_25KG], n, kind="single", max_steps=100, seed=0, device=device,
                       env_id_base=rank * n, sim_mode=args.sim_mode)
    if args.actions == "pid":
        a_host = torch.tensor(K.PID_EFFECTIVE_ACTION, dtype=torch.float32).repeat(n, 1)
    else:   # uniform in bounds (divergent episode lengths)
        g = torch.Generator().manual_seed(1234 + rank)
        lo = torch.tensor([K.SINGLE_25KG["bounds"][k][0] for k in E._lib.ACTION_ORDER], dtype=torch.float32)
        hi = torch.tensor([K.SINGLE_25KG["bounds"][k][1] for k in E._lib.ACTION_ORDER], dtype=torch.float32)
        a_host = lo + torch.rand(n, 10, generator=g) * (hi - lo)
    a_host = a_host.contiguous().pin_memory()
    a_dev = a_host.to(device)
    env.reset()
    for _ in range(args.warmup):
        env.step(a_dev)
    torch.cuda.synchronize()
    env.tick_sum.zero_()
    total_ms, wall = timer.run(lambda: env.step(a_dev), args.steps)
    ticks_all = timer.sumr(int(env.tick_sum.item()))

    def e2e_step():      # the host-facing call: pinned host actions in, pinned host (next_state, reward, done) out
        env.step_host(a_host)

    for _ in range(3):
        e2e_step()
    e2e_ms = timer.run_block(e2e_step, args.steps)

    ms_per_step = total_ms / args.steps
    units = n * world * args.steps
    fused = args.sim_mode != "thread" and n <= SIM_FUSED_MAX_ENVS
    launches = 1 if (fused or args.sim_mode == "thread") else 2
    peaks = _peaks()
    achieved = ENV_STEP_BYTES * n / (ms_per_step * 1e-3) / 1e9
    kname = ("env_sim_warp_kernel<fused epilogue>" if fused else "env_sim_warp_kernel + env_finish_kernel") \
        if args.sim_mode != "thread" else "env_step_kernel<false>"
    out = {
        "metric": "env_steps_per_sec", "value": units / (total_ms * 1e-3), "unit": "env-steps/s", "dtype": "f64",
        "ms_per_step": ms_per_step, "wall_s_timed_region": wall,
        "config": {"workload": "configs[1]: batched WeightEnv rollout, %d envs/GPU, %s action (sim-kernel throughput), "
                               "Philox noise" % (n, "PID-effective constant" if args.actions == "pid" else "uniform-random"),
                   "n_env_per_gpu": n, "actions": args.actions, "sim_mode": args.sim_mode,
                   "parallelism": "env-sharded x%d, no data-path collective" % world,
                   "l2": "flushed between timed steps (256 MiB memset outside the per-step event pair)",
                   "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks"},
        "ticks_per_sec": ticks_all / (total_ms * 1e-3), "mean_ticks_per_env_step": ticks_all / units,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": a_host.numel() * 4,
                "d2h_bytes_per_step": 13 * n,
                "path": "VecFillEnv.step_host (CUDA graph replay + stream sync per step): " +
                        ("the fused simulator kernel reads the pinned host actions and writes next_state/reward/done to pinned "
                         "host memory directly (zero-copy over PCIe, no copy nodes)" if fused and
                         os.environ.get("ERLFILL_STEP_HOST_MODE", "2") == "2" else
                         "H2D copy of the actions, step kernels, D2H copy of the packed outputs")},
        "gpu_launches": launches * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"],
                     "traffic": _traffic("sim", n)[0] if (args.actions == "pid" and args.sim_mode != "thread") else None,
                     "traffic_source": _traffic("sim", n)[1] if (args.actions == "pid" and args.sim_mode != "thread") else None,
                     "peak_source": peaks["source"], "kernel": kname,
                     "algorithmic_bytes_per_env_step": ENV_STEP_BYTES,
                     "note": "the simulator is instruction-issue bound (about %d serial ticks of integer/Philox work per "
                             "env-step against 192 B of HBM traffic), not HBM bound: see issue_roofline"
                             % round(ticks_all / units)},
    }
    ipe = SIM_WARP_INSTR_PER_ENV_STEP.get((args.actions, fused)) if args.sim_mode != "thread" else None
    if ipe:
        out["_issue"] = (n * args.steps / (total_ms * 1e-3), ipe)
    return out


def _make_agent(args, rank, device, n, batch, memory):
    import erlfill_b200 as E
    K = E.conditions
    conds = list(K.EXPERIMENT_3.values())
    env = E.VecFillEnv(conds, n, kind="multi", max_steps=100, seed=0, device=device, env_id_base=rank * n,
                       sim_mode=args.sim_mode)
    agent = E.VecERDDPG(env, seed=0, batch_size=batch, memory_size=memory, rank=rank,
                        emotion_precision=args.emotion_precision)
    return env, agent


def wl_rollout(args, rank, world, device, timer):
    import torch
    n = args.n_env
    env, agent = _make_agent(args, rank, device, n, 128, 1024)
    for _ in range(args.warmup):
        agent.rollout_step(store=False, replay_reset=False)
    torch.cuda.synchronize()
    env.tick_sum.zero_()
    total_ms, wall = timer.run(lambda: agent.rollout_step(store=False, replay_reset=False), args.steps)
    ticks_all = timer.sumr(int(env.tick_sum.item()))
    # e2e: the observations come from the host and the step's results go back to it
    obs_host = env.obs.cpu().pin_memory()
    res_host = [torch.empty(n, 2).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()]

    def e2e_step():
        env.obs.copy_(obs_host, non_blocking=True)
        ns, r, d = agent.rollout_step(store=False, replay_reset=False)
        res_host[0].copy_(ns, non_blocking=True)
        res_host[1].copy_(r, non_blocking=True)
        res_host[2].copy_(d, non_blocking=True)
        timer.stream.synchronize()

    for _ in range(2):
        e2e_step()
    e2e_ms = timer.run_block(e2e_step, args.steps)
    # the dominant kernel (emotion transformer) timed alone
    tf_ms = _kernel_ms(timer, agent._emotion_forward)
    peaks = _peaks()
    achieved = TRANSFORMER_FLOP * n / (tf_ms * 1e-3) / 1e12
    units = n * world * args.steps
    return {
        "metric": "env_steps_per_sec", "value": units / (total_ms * 1e-3), "unit": "env-steps/s",
        "dtype": "f64 simulator, %s transformer, f32 actor" % agent.emotion_precision,
        "ms_per_step": total_ms / args.steps, "wall_s_timed_region": wall,
        "config": {"workload": "configs[2]: MultiConditionWeightEnv (3 conditions round-robin), %d envs/GPU, EmotionModule "
                               "transformer -> ER-DDPG actor + OU -> env step -> emotion update, dropout off" % n,
                   "n_env_per_gpu": n, "sim_mode": args.sim_mode, "parallelism": "env-sharded x%d, no data-path collective" % world,
                   "l2": "flushed between timed steps", "timing": "sum of per-step CUDA-event pairs, max over ranks"},
        "ticks_per_sec": ticks_all / (total_ms * 1e-3), "mean_ticks_per_env_step": ticks_all / units,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": obs_host.numel() * 4,
                "d2h_bytes_per_step": n * 13},
        "gpu_launches": 8 * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"],
                     "traffic": _traffic("rollout", n)[0] if agent.emotion_precision == "bf16" else None,
                     "traffic_source": _traffic("rollout", n)[1] if agent.emotion_precision == "bf16" else None,
                     "peak_source": peaks["source"] + " (sustained)",
                     "kernel": "emotion_forward_tc_kernel" if agent.emotion_precision == "bf16" else "emotion_forward_kernel",
                     "kernel_ms": tf_ms,
                     "share_of_step": tf_ms / (total_ms / args.steps), "algorithmic_flop_per_env": TRANSFORMER_FLOP,
                     "executed_flop_per_env": TRANSFORMER_FLOP_EXECUTED,
                     "note": "achieved = the reference's algorithmic FLOPs / time; the kernel executes 16.81 of the 21.83 MFLOP "
                             "(in the last layer only the last token's out_proj/FFN feed fc_out; the other 19 rows are skipped), "
                             "so the tensor pipe itself runs at achieved x %.2f" % (TRANSFORMER_FLOP_EXECUTED / TRANSFORMER_FLOP)},
    }


def wl_update(args, rank, world, device, timer):
    import torch
    import torch.distributed as dist
    B, C = args.batch, args.replay
    env, agent = _make_agent(args, rank, device, 256, B, C)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    # synthetic replay
    chunk = 1 << 18
    for o in range(0, C, chunk):
        m = min(chunk, C - o)
        rows = torch.zeros(m, 20)
        rows[:, 0:2] = torch.rand(m, 2, generator=g) * 5
        rows[:, 2:12] = torch.rand(m, 10, generator=g) * 2 - 1
        rows[:, 12] = torch.rand(m, generator=g) * 9000 - 3000
        rows[:, 13:15] = torch.rand(m, 2, generator=g) * 5
        rows[:, 15] = (torch.rand(m, generator=g) < 0.01).float()
        rows[:, 16:19] = torch.rand(m, 3, generator=g)
        agent.memory.buf[o:o + m].copy_(rows.to(device))
    agent.memory.size = agent.memory.inserted = C
    from erlfill_b200 import parallel
    allreduce = None
    if world > 1:      # gradient exchange: one-shot NVLink peer-memory kernel (default) or the NCCL all-reduce
        if args.exchange == "peer":
            allreduce = parallel.PeerExchange(agent.learner, B)
        else:
            allreduce = parallel.GradSync(world)

    def step():
        agent.learn(allreduce=allreduce)

    for _ in range(args.warmup):
        step()
    total_ms, wall = timer.run(step, args.steps)
    rep_host = torch.empty(4).pin_memory()
    idx_host = torch.randint(0, C, (B,), generator=g)
    batch_host = agent.memory.buf[idx_host.to(device)].cpu().pin_memory()

    def e2e_step():   # minibatch supplied by the host (the reference's learn() assembles it on the host), losses read back
        agent.batch.copy_(batch_host, non_blocking=True)
        rep = agent.learner.update(agent.batch, agent.next_emotion, agent.train_step, allreduce=allreduce)
        rep_host.copy_(rep, non_blocking=True)
        timer.stream.synchronize()

    for _ in range(2):
        e2e_step()
    e2e_ms = timer.run_block(e2e_step, args.steps)
    peaks = _peaks()
    ms = total_ms / args.steps
    achieved = UPDATE_FLOP_PER_SAMPLE * B / (ms * 1e-3) / 1e12
    return {
        # weak scaling: every rank pushes its own minibatch of B through the update each step, so the units processed per step
        # are `world` minibatch updates (one synchronous optimiser step on the global batch of world x B)
        "metric": "er_ddpg_updates_per_sec", "value": world * args.steps / (total_ms * 1e-3), "unit": "updates/s", "dtype": "f32",
        "ms_per_step": ms, "wall_s_timed_region": wall, "samples_per_sec": B * world * args.steps / (total_ms * 1e-3),
        "sync_steps_per_sec": args.steps / (total_ms * 1e-3),
        "config": {"workload": "configs[3]: ER-DDPG update, minibatch %d per GPU, %d-transition replay, data-parallel x%d "
                               "(%s); value counts per-GPU minibatch updates, sync_steps_per_sec the optimiser steps on the "
                               "global batch" % (B, C, world, "no exchange" if world == 1 else
                                                 ("one-shot NVLink peer-memory gradient exchange fused with the clip-norm partials"
                                                  if args.exchange == "peer" else "NCCL flat-bucket all-reduce")),
                   "minibatch_per_gpu": B, "replay": C, "parallelism": "dp%d" % world,
                   "l2": "flushed between timed steps", "timing": "sum of per-step CUDA-event pairs, max over ranks"},
        "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": B * 80, "d2h_bytes_per_step": 16},
        # sample + gather + 24 launches of erlfill_ddpg_update (+ 2 x (signal, reduce) - 2 sumsq with the peer exchange)
        "gpu_launches": (26 if world == 1 else 28) * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + " (sustained)", "kernel": "whole update (sample + gather + U1-U5)",
                     "algorithmic_flop_per_sample": UPDATE_FLOP_PER_SAMPLE,
                     "note": "%.2f GFLOP per update: launch/latency bound at this size" % (UPDATE_FLOP_PER_SAMPLE * B / 1e9)},
    }


def wl_replay(args, rank, world, device, timer):
    """Rows R1/R2 on their own: replay insert of one lock-step batch of transitions (configs[4] size: 131,072 envs per GPU)
    into a 1M-row ring, and the minibatch path sample + gather.  HBM-bound kernels: achieved GB/s on the algorithmic bytes
    (insert: 76 B read + 80 B written per transition; gather: 8 B index + 80 B read + 80 B written per sample)."""
    import torch
    from erlfill_b200.ddpg import ReplayBuffer
    n, C, B = args.n_env, args.replay, args.batch
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    mem = ReplayBuffer(C, device=device)
    state = torch.rand(n, 2, generator=g).to(device)
    action = (torch.rand(n, 10, generator=g) * 100).to(device)
    reward = torch.rand(n, generator=g).to(device)
    next_state = torch.rand(n, 2, generator=g).to(device)
    done = (torch.rand(n, generator=g) < 0.01).to(torch.uint8).to(device)
    emotion = torch.rand(n, 3, generator=g).to(device)
    scale = (torch.rand(10, generator=g) + 1).to(device)
    offset = torch.rand(10, generator=g).to(device)

    def insert():
        mem.insert(state, action, reward, next_state, done, emotion, scale=scale, offset=offset)

    for _ in range(max(args.warmup, C // n + 1)):     # fill the ring once so sampling covers all of it
        insert()
    total_ms, wall = timer.run(insert, args.steps)
    ms = total_ms / args.steps
    idx = torch.empty(B, dtype=torch.int64, device=device)
    out = torch.empty(B, 20, dtype=torch.float32, device=device)
    cnt = [0]

    def sample_gather():
        cnt[0] += 1
        mem.sample_indices(B, 1234, cnt[0], out=idx)
        mem.gather(idx, out=out)

    sg_ms = _kernel_ms(timer, sample_gather, reps=10)
    big = 1 << 20
    idx_big = torch.randint(0, C, (big,), generator=g).to(device)
    out_big = torch.empty(big, 20, dtype=torch.float32, device=device)
    gb_ms = _kernel_ms(timer, lambda: mem.gather(idx_big, out=out_big), reps=10)
    nb = min(C, 1 << 20)
    bigin = [torch.rand(nb, 2, generator=g).to(device), (torch.rand(nb, 10, generator=g) * 100).to(device), torch.rand(nb, generator=g).to(device),
             torch.rand(nb, 2, generator=g).to(device), (torch.rand(nb, generator=g) < 0.01).to(torch.uint8).to(device),
             torch.rand(nb, 3, generator=g).to(device)]
    ib_ms = _kernel_ms(timer, lambda: mem.insert(*bigin, scale=scale, offset=offset), reps=10)
    host = [t.cpu().pin_memory() for t in (state, action, reward, next_state, done, emotion)]
    dev = [state, action, reward, next_state, done, emotion]
    out_host = torch.empty(B, 20).pin_memory()

    def e2e_step():    # transitions handed over by a host producer, one minibatch read back
        for d, h in zip(dev, host):
            d.copy_(h, non_blocking=True)
        insert()
        sample_gather()
        out_host.copy_(out, non_blocking=True)
        timer.stream.synchronize()

    for _ in range(2):
        e2e_step()
    e2e_ms = timer.run_block(e2e_step, args.steps)
    peaks = _peaks()
    ins_bytes = (76 + 80) * n
    achieved = ins_bytes / (ms * 1e-3) / 1e9
    return {
        "metric": "replay_insert_transitions_per_sec", "value": n * world * args.steps / (total_ms * 1e-3), "unit": "transitions/s",
        "dtype": "f32", "ms_per_step": ms, "wall_s_timed_region": wall,
        "config": {"workload": "rows R1/R2: replay insert of %d transitions per GPU per step into a %d-row ring (FIFO); "
                               "sample + gather of B = %d timed beside it" % (n, C, B),
                   "n_env_per_gpu": n, "replay": C, "minibatch_per_gpu": B, "parallelism": "per-rank replay shard x%d" % world,
                   "l2": "flushed between timed steps", "timing": "sum of per-step CUDA-event pairs, max over ranks"},
        "e2e": {"value": n * world * args.steps / (e2e_ms * 1e-3), "unit": "transitions/s",
                "h2d_bytes_per_step": sum(h.numel() * h.element_size() for h in host), "d2h_bytes_per_step": B * 80},
        "gpu_launches": 1 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": None, "peak_source": peaks["source"], "kernel": "replay_insert_kernel",
                     "algorithmic_bytes_per_transition": 156},
        "sample_gather": {"batch": B, "ms": sg_ms, "note": "two launches, latency bound at this size"},
        "insert_1M": {"transitions": nb, "ms": ib_ms, "achieved_gbs": 156 * nb / (ib_ms * 1e-3) / 1e9,
                      "frac": 156 * nb / (ib_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "note": "the same kernel on a batch large enough to amortise launch and ramp"},
        "gather_1M": {"samples": big, "ms": gb_ms, "achieved_gbs": (8 + 160) * big / (gb_ms * 1e-3) / 1e9,
                      "frac": (8 + 160) * big / (gb_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "note": "random 80 B rows: 96 of every 3 x 32 B sectors fetched are 80 useful bytes"},
    }


def wl_n1(args, rank, world, device, timer):
    """configs[0]: the reference's own case — ONE env, ER-DDPG (`DDPGAgent(use_td_error=True)` + `EmotionModule`), B = 128, replay
    100,000 — through the N = 1 drop-ins and the loop body of train_ddpg (`loops.train_ddpg`: reset, act, step, emotion.update,
    remember, learn, get_emotion per step, numpy in / numpy out like the reference).  Host-driven by nature: value == e2e.
    The reference's own figure for this loop is 27 env-steps/s on 8 vCPU (survey-measured)."""
    import random
    import numpy as np
    import torch
    import erlfill_b200 as E
    from erlfill_b200 import loops
    K = E.conditions
    cfg = K.EXPERIMENT_3["variant_25kg"]
    torch.manual_seed(0); random.seed(0); np.random.seed(0)
    env = E.MultiConditionWeightEnv({k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time", "bounds")})
    agent = E.DDPGAgent(env, device=device, use_td_error=True)
    agent.emotion = E.EmotionModule(device=device)
    env.attach_agent(agent)
    steps_per_ep = 50
    n_steps = [0]
    loops.train_ddpg(env, agent, episodes=4, max_steps=steps_per_ep)        # warm-up: past the first learn() (128 transitions)
    timer.sync()
    t0 = time.perf_counter()
    loops.train_ddpg(env, agent, episodes=max(2, args.steps // 2), max_steps=steps_per_ep,
                     on_step=lambda *a: n_steps.__setitem__(0, n_steps[0] + 1))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    v = n_steps[0] / dt
    peaks = _peaks()
    return {
        "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "dtype": "f64 simulator, f32 networks",
        "ms_per_step": dt / n_steps[0] * 1e3, "wall_s_timed_region": dt,
        "config": {"workload": "configs[0]: ER-DDPG on one MultiConditionWeightEnv (25 kg), N = 1 drop-ins, train_ddpg loop body "
                               "(act -> step -> emotion.update -> remember -> learn B=128 -> get_emotion), %d env-steps timed by "
                               "wall clock (host-driven loop: every call returns numpy like the reference)" % n_steps[0],
                   "n_env_per_gpu": 1, "minibatch": 128, "parallelism": "single env", "l2": "not flushed (host-paced launches)",
                   "timing": "wall clock around the loop, device synchronised on both sides"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * (2 + 10 + 19), "d2h_bytes_per_step": 4 * (10 + 2 + 3 + 3 + 2)},
        "gpu_launches": None,
        "roofline": {"bound": "hbm", "achieved": ENV_STEP_BYTES * v / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ENV_STEP_BYTES * v / 1e9 / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                     "kernel": "whole N = 1 step", "note": "latency bound: about 40 dependent launches and 8 host syncs per step"},
        "reference_cpu": {"value": 27.0, "unit": "env-steps/s", "source": "SURVEY.md section 6: the reference's train_ddpg body, 8 vCPU, survey-measured"},
    }


def wl_erddpg(args, rank, world, device, timer):
    """The whole north-star loop for ER-DDPG: per lock-step step, transformer -> actor + OU -> env step -> replay-branch reset ->
    emotion update -> replay insert of every env's transition -> ONE minibatch update (B = 4096, data-parallel at N > 1)."""
    import torch
    from erlfill_b200 import parallel
    n, B = args.n_env, args.batch
    env, agent = _make_agent(args, rank, device, n, B, max(4 * n, 1 << 20))
    allreduce = None
    if world > 1:
        allreduce = parallel.PeerExchange(agent.learner, B) if args.exchange == "peer" else parallel.GradSync(world)

    def step():
        agent.rollout_step()
        agent.learn(allreduce=allreduce)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    env.tick_sum.zero_()
    total_ms, wall = timer.run(step, args.steps)
    ticks_all = timer.sumr(int(env.tick_sum.item()))
    obs_host = env.obs.cpu().pin_memory()
    res_host = [torch.empty(n, 2).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
                torch.empty(4).pin_memory()]

    def e2e_step():
        env.obs.copy_(obs_host, non_blocking=True)
        ns, r, d = agent.rollout_step()
        rep = agent.learn(allreduce=allreduce)
        res_host[0].copy_(ns, non_blocking=True); res_host[1].copy_(r, non_blocking=True); res_host[2].copy_(d, non_blocking=True)
        res_host[3].copy_(rep, non_blocking=True)
        timer.stream.synchronize()

    for _ in range(2):
        e2e_step()
    e2e_ms = timer.run_block(e2e_step, args.steps)
    tf_ms = _kernel_ms(timer, agent._emotion_forward)
    peaks = _peaks()
    achieved = TRANSFORMER_FLOP * n / (tf_ms * 1e-3) / 1e12
    units = n * world * args.steps
    return {
        "metric": "env_steps_per_sec", "value": units / (total_ms * 1e-3), "unit": "env-steps/s",
        "dtype": "f64 simulator, %s transformer, f32 actor, f32 (3xTF32) update" % agent.emotion_precision,
        "ms_per_step": total_ms / args.steps, "wall_s_timed_region": wall, "updates_per_sec": world * args.steps / (total_ms * 1e-3),
        "config": {"workload": "ER-DDPG full loop on MultiConditionWeightEnv (3 conditions round-robin), %d envs/GPU: transformer -> actor "
                               "+ OU -> env step -> replay-branch reset -> emotion update -> replay insert -> one update of minibatch %d, "
                               "dropout off" % (n, B),
                   "n_env_per_gpu": n, "minibatch_per_gpu": B,
                   "parallelism": "env-sharded x%d, data-parallel update (%s)" % (world, "none" if world == 1 else args.exchange),
                   "l2": "flushed between timed steps", "timing": "sum of per-step CUDA-event pairs, max over ranks"},
        "ticks_per_sec": ticks_all / (total_ms * 1e-3), "mean_ticks_per_env_step": ticks_all / units,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": obs_host.numel() * 4,
                "d2h_bytes_per_step": n * 13 + 16},
        "gpu_launches": (10 + 26) * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": _traffic("rollout", n)[0],
                     "traffic_source": _traffic("rollout", n)[1], "peak_source": peaks["source"] + " (sustained)",
                     "kernel": "emotion_forward_tc_kernel", "kernel_ms": tf_ms, "share_of_step": tf_ms / (total_ms / args.steps),
                     "algorithmic_flop_per_env": TRANSFORMER_FLOP, "executed_flop_per_env": TRANSFORMER_FLOP_EXECUTED},
    }


def wl_loop(args, rank, world, device, timer):
    """configs[4]: TD3 / SAC on MultiConditionWeightEnv, full lock-step loop: act -> env step -> remember -> one update of
    B = 4096 from the rank's replay shard (data-parallel gradient all-reduce when world > 1)."""
    import torch
    import erlfill_b200 as E
    from erlfill_b200 import parallel
    K = E.conditions
    n, B = args.n_env, args.batch
    kind = args.workload
    env = E.VecFillEnv(list(K.EXPERIMENT_3.values())[:1], n, kind="multi", max_steps=100, seed=0, device=device, env_id_base=rank * n,
                       sim_mode=args.sim_mode)
    if kind == "esac":      # EmotionSAC (row N3): the emotion transformer runs every step on the tensor-core kernel
        agent = E.VecEmotionSAC(env, seed=0, batch_size=B, memory_size=max(4 * n, 1 << 20), rank=rank, emotion_precision="bf16")
        agent.reset()
    else:
        agent = (E.VecTD3 if kind == "td3" else E.VecSAC)(env, seed=0, batch_size=B, memory_size=max(4 * n, 1 << 20), rank=rank)
        env.reset()
    sync = parallel.GradSync(world) if world > 1 else None

    def step():
        agent.rollout_step()
        if kind == "td3":
            agent.train_step(allreduce=sync)
        else:
            agent.update(allreduce=sync)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    env.tick_sum.zero_()
    total_ms, wall = timer.run(step, args.steps)
    ticks_all = timer.sumr(int(env.tick_sum.item()))
    obs_host = env.obs.cpu().pin_memory()
    res_host = [torch.empty(n, 2).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
                torch.empty(4 if kind == "esac" else 3).pin_memory()]

    def e2e_step():
        env.obs.copy_(obs_host, non_blocking=True)
        ns, r, d = agent.rollout_step()
        rep = agent.train_step(allreduce=sync)[0] if kind == "td3" else agent.update(allreduce=sync)
        res_host[0].copy_(ns, non_blocking=True); res_host[1].copy_(r, non_blocking=True); res_host[2].copy_(d, non_blocking=True)
        res_host[3].copy_(rep, non_blocking=True)
        timer.stream.synchronize()

    for _ in range(2):
        e2e_step()
    e2e_ms = timer.run_block(e2e_step, args.steps)
    sim_ms = _kernel_ms(timer, lambda: env.step(agent.action))
    peaks = _peaks()
    units = n * world * args.steps
    achieved = ENV_STEP_BYTES * n / (sim_ms * 1e-3) / 1e9
    return {
        "metric": "env_steps_per_sec", "value": units / (total_ms * 1e-3), "unit": "env-steps/s",
        "dtype": "f64 simulator, f32 (3xTF32 tensor-core GEMM) networks", "ms_per_step": total_ms / args.steps, "wall_s_timed_region": wall,
        "updates_per_sec": args.steps / (total_ms * 1e-3),
        "config": {"workload": "configs[4]: %s on MultiConditionWeightEnv, %d envs/GPU, full loop: %sact -> env step -> remember -> "
                               "one update of minibatch %d" % ("EmotionSAC" if kind == "esac" else kind.upper(), n,
                                                               "emotion transformer -> " if kind == "esac" else "", B),
                   "n_env_per_gpu": n, "minibatch_per_gpu": B,
                   "parallelism": "env-sharded x%d, data-parallel update (flat-bucket gradient all-reduce)" % world,
                   "l2": "flushed between timed steps", "timing": "sum of per-step CUDA-event pairs, max over ranks"},
        "ticks_per_sec": ticks_all / (total_ms * 1e-3), "mean_ticks_per_env_step": ticks_all / units,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": obs_host.numel() * 4,
                "d2h_bytes_per_step": n * 13 + 12},
        "gpu_launches": 130 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": None, "peak_source": peaks["source"], "kernel": "env_sim_warp_kernel + env_finish_kernel",
                     "kernel_ms": sim_ms, "share_of_step": sim_ms / (total_ms / args.steps),
                     "algorithmic_bytes_per_env_step": ENV_STEP_BYTES,
                     "note": "the simulator dominates this loop and is instruction-issue bound, not HBM bound (DESIGN.md section 4.1)"},
    }


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import erlfill_b200 as E
    E._lib.require_device()
    timer = Timer(device, world)
    sampler = ClockSampler(local)
    sampler.start()
    wl = {"sim": wl_sim, "rollout": wl_rollout, "update": wl_update, "td3": wl_loop, "sac": wl_loop, "esac": wl_loop,
          "replay": wl_replay, "n1": wl_n1, "erddpg": wl_erddpg}
    head = "sim" if args.workload == "all" else args.workload
    if args.workload == "all":
        args.n_env = args.n_env or 4096
    out = wl[head](args, rank, world, device, timer)
    clocks = sampler.stop()
    others = []
    if args.workload == "all":
        # the two other metrics of BASELINE.json on their own configs, compact (same timing rules)
        import copy
        for name, n_env, steps in (("rollout", 65536, 10), ("update", 256, 20)):
            a2 = copy.copy(args)
            a2.n_env, a2.steps = n_env, steps
            o = wl[name](a2, rank, world, device, timer)
            others.append({k: o[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "e2e", "gpu_launches", "roofline")}
                          | {"workload": o["config"]["workload"], "steps": steps})
    if rank == 0:
        issue = out.pop("_issue", None)
        line = {"metric": out.pop("metric"), "value": out.pop("value"), "unit": out.pop("unit"), "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": out.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": out.pop("dtype"), "data": "synthetic"}
        line.update(out)
        line["clocks"] = clocks
        if issue:
            mhz = clocks.get("sm_mhz") or 1965.0
            ceiling = 148 * 4 * mhz * 1e6 / issue[1]
            line["issue_roofline"] = {"achieved_env_steps_per_s_per_gpu": issue[0], "ceiling": ceiling, "frac": issue[0] / ceiling,
                                      "warp_instr_per_env_step": issue[1], "sm_mhz_used": mhz,
                                      "model": "148 SMs x 4 schedulers x 1 warp-instruction/clk / measured warp-instructions per env-step"}
        if head == "sim":
            line["cpu_baseline"] = cpu_baseline_sim(args, budget_s=args.cpu_budget)
        if others:
            line["other_workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        # no NCCL / interpreter teardown: captured graphs that hold collectives and peer mappings are still alive
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py may execute oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_sim_once(args, n_env, n_threads, steps):
    import numpy as np
    from oracle import oracle_lib as O
    import erlfill_b200 as E
    K = E.conditions
    c = K.SINGLE_25KG
    cond = [O.make_condition(c["target_weight"], c["target_weight_err"], c["target_time"], c["bounds"])]
    env = O.BatchEnv(O.ENV_SINGLE, cond, n_env, max_steps=100, seed=0)
    if args.actions == "pid":
        a = np.tile(np.array(K.PID_EFFECTIVE_ACTION, np.float32), (n_env, 1))
    else:
        rng = np.random.RandomState(1234)
        lo = np.array([c["bounds"][k][0] for k in O.ACTION_ORDER], np.float32)
        hi = np.array([c["bounds"][k][1] for k in O.ACTION_ORDER], np.float32)
        a = (lo + rng.uniform(0, 1, (n_env, 10)) * (hi - lo)).astype(np.float32)
    t0 = time.perf_counter()
    ticks = 0
    for s in range(steps):
        ticks += env.step(a, s, n_threads=n_threads)
    dt = time.perf_counter() - t0
    return n_env * steps / dt, ticks / dt, dt


def cpu_baseline_sim(args, budget_s=12.0):
    """C port of the reference path (oracle/erlfill_oracle.c), all host cores, bounded sample."""
    import random
    from oracle import oracle_lib as O, py_port
    cores = O.lib().orc_max_threads()
    rate, _, _ = _cpu_sim_once(args, 64 * cores, cores, 1)          # calibrate
    n_env = max(64 * cores, int(rate * budget_s / 4) // 64 * 64)
    v, tps, dt = _cpu_sim_once(args, n_env, cores, 4)
    # the reference's own language: pure-Python restatement, one core, ~2 s
    p = dict(zip(["target_weight", "fast_weight", "medium_weight", "slow_weight", "fast_delay", "medium_delay",
                  "slow_delay", "fast_opening", "medium_opening", "slow_opening", "unload_delay"],
                 [25000, 18000, 0, 24900, 100, 100, 100, 35, 3, 3, 300]))
    rng = random.Random(0)
    t0, runs = time.perf_counter(), 0
    while time.perf_counter() - t0 < min(2.0, budget_s):
        py_port.simulate_run(p, rng)
        runs += 1
    py_rate = runs / (time.perf_counter() - t0)
    return {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d envs x 4 steps of the same workload, C oracle (pthreads over envs), %.1f s" % (n_env, dt),
            "ticks_per_sec": tps,
            "python_port_one_core": {"value": py_rate, "unit": "env-steps/s", "cores": 1,
                                     "note": "pure-Python restatement of simulate_run (the reference is a CPython loop)"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.workload not in ("sim", "all"):
        print(json.dumps({"impl": "reference", "unavailable": "the CPU port is timed on the sim workload (configs[1]) only"}),
              flush=True)
        return
    from oracle import oracle_lib as O
    cores = O.lib().orc_max_threads()
    n = args.n_env
    # bounded sample: every "step" is one pass of the C port over n_ref envs of the same workload
    rate, _, _ = _cpu_sim_once(args, 64 * cores, cores, 1)
    per_step_budget = max(0.5, min(10.0, 120.0 / max(1, args.steps + args.warmup)))
    n_ref = int(min(n, max(64 * cores, rate * per_step_budget)))
    for _ in range(args.warmup):
        _cpu_sim_once(args, n_ref, cores, 1)
    t0 = time.perf_counter()
    v, tps, dt = _cpu_sim_once(args, n_ref, cores, args.steps)
    total = time.perf_counter() - t0
    sample = "%d of %d envs per step (bounded sample), C port of the reference path, %d threads" % (n_ref, n, cores)
    out = {"impl": "reference", "metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "configs[1]: batched WeightEnv rollout, %d envs/GPU, %s action (sim-kernel "
                                  "throughput)" % (n, "PID-effective constant" if args.actions == "pid" else "uniform-random"),
                      "n_env_per_gpu": n, "actions": args.actions},
           "ticks_per_sec": tps,
           "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "sim", "rollout", "update", "td3", "sac", "esac", "replay", "n1", "erddpg"],
                    help="all = the sim headline line plus compact rollout and update results under other_workloads")
    ap.add_argument("--n-env", type=int, default=None)
    ap.add_argument("--actions", default="pid", choices=["pid", "random"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--replay", type=int, default=1000000)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--sim-mode", default="auto", choices=["auto", "warp", "thread"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="gradient exchange of the data-parallel update at N > 1 (csrc/peer.cu or ncclAllReduce)")
    ap.add_argument("--emotion-precision", default="bf16", choices=["bf16", "fp32"],
                    help="EmotionTransformer kernel of the rollout: tcgen05 bf16 (default) or the exact fp32 SIMT twin")
    args = ap.parse_args()
    if args.n_env is None:
        args.n_env = {"all": 4096, "sim": 4096, "rollout": 65536, "update": 256, "td3": 131072, "sac": 131072, "esac": 65536,
                      "replay": 131072, "n1": 1, "erddpg": 65536}[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
