#!/usr/bin/env python
"""bench.py — throughput of the ERL-Fill hot path on B200 (contract: DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sim|rollout|update] [--n-env 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU port of the reference path on the host cores

One JSON line on stdout (rank 0).  A "step" is one lock-step pass of the workload over the batch of every rank:
  rollout  configs[2] of BASELINE.json (the configuration the north-star target is quoted on): MultiConditionWeightEnv,
           65,536 envs/GPU, EmotionModule transformer -> ER-DDPG actor + OU -> env step -> emotion update
  sim      configs[1]: batched WeightEnv rollout, PID-effective constant action (4096 envs/GPU)
  update   configs[3]: ER-DDPG update, minibatch 4096 from a 1M-transition replay (sample + gather + U1-U5)
  td3|sac  configs[4]: full TD3 / SAC loop (act -> env step -> remember -> update of B=4096), 131,072 envs/GPU
  esac     EmotionSAC (ERL_Fill_agent.py) full loop incl. the emotion transformer, 65,536 envs/GPU
  erddpg   the whole ER-DDPG loop: rollout (with replay insert / replay-branch resets) + one update of B=4096 per lock-step step
  replay   rows R1/R2 alone against the HBM roofline; n1: configs[0], the N = 1 drop-in loop
  all      (default) the rollout line, with the sim and update results attached under "other_workloads"
`value` is device-timed with inputs resident in HBM; `e2e` goes through the public Python API with pinned HOST
buffers (H2D of the step's inputs, D2H of its results inside the timed region).

`--impl reference` times the UNMODIFIED reference (vendored byte for byte into the git-ignored baseline/_ref/ by
oracle/vendor_ref.py, imported through oracle/ref_shim.py) on the host cores for the SAME workload and prints the same line
with an identical `config` (oracle/ref_bench.py).  At N = 1 our arm runs that leg itself
on a bounded sample first (before CUDA is touched) and reports it as `cpu_baseline` (kind "reference").
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
SIM_FUSED_MAX_ENVS = 148 * 32   # csrc/env_step.cu: resident_warp_slots()


def _ncu_constants():
    """Numbers that only an ncu capture can give (warp-instructions per env-step, DRAM bytes per launch) live in
    profiles/ncu_constants.json next to the sha256 of the kernel source they were captured from; when the source has
    changed since, the entry is reported with "stale": true instead of silently going out of date."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    if not os.path.exists(path):
        return {}
    d = json.load(open(path))
    for e in d.values():
        src = os.path.join(ROOT, "erl-fill_b200", "csrc", e.get("source", ""))
        sha = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16] if os.path.isfile(src) else None
        e["stale"] = sha != e.get("source_sha16")
    return d


def _traffic(key):
    e = _ncu_constants().get(key)
    if not e or "dram_bytes_per_launch" not in e:
        return None, None
    return e["dram_bytes_per_launch"], e["file"] + (" (STALE: kernel source changed since the capture)" if e["stale"] else "")


def _sim_instr(actions, fused):
    e = _ncu_constants().get("sim_%s_%s" % (actions, "fused" if fused else "two_kernel"))
    return (e["warp_instr_per_env_step"], e["stale"], e["file"]) if e and "warp_instr_per_env_step" in e else None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


DEFAULT_N_ENV = {"all": 65536, "sim": 4096, "rollout": 65536, "update": 256, "td3": 131072, "sac": 131072, "esac": 65536,
                 "replay": 131072, "n1": 1, "erddpg": 65536}
DEFAULT_REPLAY = {"replay": 4 << 20}        # rows R1/R2 are measured on a ring larger than the 126 MB L2 (4 Mi rows x 80 B = 336 MB)


def workload_config(name, args, world):
    """The `config` object of the JSON line.  BOTH arms (ours and `--impl reference`) print exactly this dict for a given
    command line: it names the workload, nothing about how an arm runs it (that goes under `timing` / `cpu_baseline`)."""
    n, B, C = args.n_env, args.batch, args.replay
    act = "PID-effective constant" if args.actions == "pid" else "uniform-random"
    if name == "sim":
        return {"workload": "configs[1]: batched WeightEnv rollout, %d envs/GPU, %s action (sim-kernel throughput)" % (n, act),
                "n_env_per_gpu": n, "actions": args.actions, "parallelism": "env-sharded x%d, no data-path collective" % world}
    if name == "rollout":
        return {"workload": "configs[2]: MultiConditionWeightEnv (3 conditions round-robin), %d envs/GPU, per env-step: EmotionModule "
                            "transformer -> ER-DDPG actor + OU -> env step -> emotion update" % n,
                "n_env_per_gpu": n, "parallelism": "env-sharded x%d, no data-path collective" % world}
    if name == "update":
        return {"workload": "configs[3]: ER-DDPG update (DDPGAgent.learn), minibatch %d per GPU, %d-transition replay" % (B, C),
                "minibatch_per_gpu": B, "replay": C, "parallelism": "dp%d" % world}
    if name == "erddpg":
        return {"workload": "ER-DDPG full loop on MultiConditionWeightEnv (3 conditions round-robin), %d envs/GPU: transformer -> "
                            "actor + OU -> env step -> reset -> emotion update -> remember -> one update of minibatch %d" % (n, B),
                "n_env_per_gpu": n, "minibatch_per_gpu": B, "parallelism": "env-sharded x%d, data-parallel update" % world}
    if name in ("td3", "sac", "esac"):
        return {"workload": "configs[4]: %s on MultiConditionWeightEnv, %d envs/GPU, full loop: %sact -> env step -> remember -> one "
                            "update of minibatch %d" % ("EmotionSAC" if name == "esac" else name.upper(), n,
                                                        "emotion transformer -> " if name == "esac" else "", B),
                "n_env_per_gpu": n, "minibatch_per_gpu": B, "parallelism": "env-sharded x%d, data-parallel update" % world}
    if name == "replay":
        return {"workload": "rows R1/R2: replay insert of %d transitions per GPU per step into a %d-row ring (FIFO); sample + gather "
                            "of B = %d timed beside it" % (n, C, B),
                "n_env_per_gpu": n, "replay": C, "minibatch_per_gpu": B, "parallelism": "per-rank replay shard x%d" % world}
    if name == "n1":
        return {"workload": "configs[0]: ER-DDPG on one MultiConditionWeightEnv (25 kg), train_ddpg loop body (act -> step -> "
                            "emotion.update -> remember -> learn B=128 -> get_emotion)",
                "n_env_per_gpu": 1, "minibatch": 128, "parallelism": "single env"}
    raise KeyError(name)


TIMING = {"l2": "flushed between timed steps (256 MiB memset outside the per-step event pair)",
          "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks"}


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
    tkey = None
    if args.actions == "pid" and args.sim_mode != "thread":
        tkey = "sim_pid_fused" if (fused and n == 4096) else ("sim_pid_two_kernel" if (not fused and n == 65536) else None)
    kname = ("env_sim_warp_kernel<fused epilogue>" if fused else "env_sim_warp_kernel + env_finish_kernel") \
        if args.sim_mode != "thread" else "env_step_kernel<false>"
    out = {
        "metric": "env_steps_per_sec", "value": units / (total_ms * 1e-3), "unit": "env-steps/s", "dtype": "f64",
        "ms_per_step": ms_per_step, "wall_s_timed_region": wall,
        "config": workload_config("sim", args, world), "timing": dict(TIMING, sim_mode=args.sim_mode, noise="Philox-4x32-10"),
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
                     "traffic": _traffic(tkey)[0], "traffic_source": _traffic(tkey)[1],
                     "peak_source": peaks["source"], "kernel": kname,
                     "algorithmic_bytes_per_env_step": ENV_STEP_BYTES,
                     "note": "the simulator is instruction-issue bound (about %d serial ticks of integer/Philox work per "
                             "env-step against 192 B of HBM traffic), not HBM bound: see issue_roofline"
                             % round(ticks_all / units)},
    }
    ipe = _sim_instr(args.actions, fused) if args.sim_mode != "thread" else None
    if ipe:
        out["_issue"] = (n * args.steps / (total_ms * 1e-3), ipe)
    return out


def _dp_check(agent, allreduce, world, device):
    """After a data-parallel timed region: (1) the peer exchange never timed out waiting for a rank (a timeout would have
    summed stale buckets), (2) every rank holds bit-identical actor / critic parameters (integer checksums of the raw bits)."""
    if world == 1:
        return None
    import torch
    import torch.distributed as dist
    err = allreduce.error() if hasattr(allreduce, "error") else 0
    L = agent.learner
    bits = torch.cat([L.actor.flat, L.critic.flat, L.actor_target.flat, L.critic_target.flat]).view(torch.int32).to(torch.int64)
    chk = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=device)).sum(),
                       torch.tensor(err, dtype=torch.int64, device=device)])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    errs = [int(c[2]) for c in allc]
    same = all(bool((c[:2] == allc[0][:2]).all()) for c in allc)
    if any(errs) or not same:
        raise SystemExit("data-parallel check FAILED: peer-exchange errors per rank %s, replicas bit-identical: %s" % (errs, same))
    return {"peer_exchange_errors": errs, "replicas_bit_identical": same, "checksum": int(allc[0][0])}


def _make_agent(args, rank, device, n, batch, memory):
    import erlfill_b200 as E
    K = E.conditions
    conds = list(K.EXPERIMENT_3.values())
    env = E.VecFillEnv(conds, n, kind="multi", max_steps=100, seed=0, device=device, env_id_base=rank * n,
                       sim_mode=args.sim_mode)
    agent = E.VecERDDPG(env, seed=0, batch_size=batch, memory_size=memory, rank=rank,
                        emotion_precision=args.emotion_precision, dropout=args.dropout, update_precision=args.update_precision)
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
    units = n * world * args.steps
    # the other dropout mode of configs[2] ("report both .eval() and Philox-mask mode"), same envs, fewer steps
    other = "philox" if args.dropout == "off" else "off"
    del agent
    torch.cuda.empty_cache()
    a2 = type(args)(**vars(args)); a2.dropout = other
    env2, agent2 = _make_agent(a2, rank, device, n, 128, 1024)
    for _ in range(3):
        agent2.rollout_step(store=False, replay_reset=False)
    k2 = max(3, args.steps // 4)
    ms2, _ = timer.run(lambda: agent2.rollout_step(store=False, replay_reset=False), k2)
    tf2 = _kernel_ms(timer, agent2._emotion_forward, reps=3)
    other_mode = {"dropout": other, "value": n * world * k2 / (ms2 * 1e-3), "unit": "env-steps/s", "ms_per_step": ms2 / k2, "steps": k2,
                  "transformer_kernel_ms": tf2,
                  "note": "philox = train() mode, the mode the reference runs (it never calls .eval() on this path): four dropout layers per "
                          "encoder layer + two in the actor, keep-masks from Philox (16 random bits per mask bit; 4x32-7 in the transformer, csrc/tf_dropout.cuh, 4x32-10 in the actor); "
                          "the transformer kernel is then bound by the integer pipe that draws 163,840 mask bits per env and call, not by the tensor pipe"}
    env, agent = env2, agent2       # (the roofline object below describes the headline mode: its kernel time was taken above)
    agent.emotion_precision = args.emotion_precision
    return {
        "metric": "env_steps_per_sec", "value": units / (total_ms * 1e-3), "unit": "env-steps/s",
        "dtype": "f64 simulator, %s transformer, f32 actor" % agent.emotion_precision,
        "other_dropout_mode": other_mode,
        "ms_per_step": total_ms / args.steps, "wall_s_timed_region": wall,
        "config": workload_config("rollout", args, world),
        "timing": dict(TIMING, sim_mode=args.sim_mode, dropout=args.dropout,
                       dropout_note="off = module.eval(); philox = train() mode as the reference ships it (it never calls .eval() on "
                                    "this path), keep-masks from counter-based generators"),
        "ticks_per_sec": ticks_all / (total_ms * 1e-3), "mean_ticks_per_env_step": ticks_all / units,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": obs_host.numel() * 4,
                "d2h_bytes_per_step": n * 13},
        "gpu_launches": agent.launches_per_rollout_step(store=False, replay_reset=False) * args.steps,
        "roofline": _transformer_roofline(agent, n, tf_ms, total_ms / args.steps),
    }


def _transformer_roofline(agent, n, tf_ms, step_ms):
    """Tensor roofline of the emotion-transformer kernel, timed ALONE (an L2 flush before each of its launches): the peak that
    applies is the burst bf16 figure of MEASURED_PEAKS.json.  `frac` counts the reference's algorithmic FLOPs; `frac_executed`
    the FLOPs the kernel really issues (it skips the 19 dead rows of the last layer), i.e. what the tensor pipe does."""
    peaks = _peaks()
    achieved = TRANSFORMER_FLOP * n / (tf_ms * 1e-3) / 1e12
    executed = TRANSFORMER_FLOP_EXECUTED * n / (tf_ms * 1e-3) / 1e12
    tc = agent.emotion_precision == "bf16"
    tr, trs = _traffic("emotion_tc_65536") if (tc and n == 65536) else (None, None)
    return {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"], "traffic": tr, "traffic_source": trs,
            "peak_source": peaks["source"] + " (burst: the kernel is timed alone)",
            "kernel": "emotion_forward_tc_kernel" if tc else "emotion_forward_kernel", "kernel_ms": tf_ms,
            "share_of_step": tf_ms / step_ms, "algorithmic_flop_per_env": TRANSFORMER_FLOP,
            "executed_flop_per_env": TRANSFORMER_FLOP_EXECUTED, "achieved_executed": executed,
            "frac_executed": executed / peaks["bf16_tflops"], "frac_vs_sustained_peak": achieved / peaks["bf16_tflops_sustained"]}


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
    dp = _dp_check(agent, allreduce, world, device)
    peaks = _peaks()
    ms = total_ms / args.steps
    achieved = UPDATE_FLOP_PER_SAMPLE * B / (ms * 1e-3) / 1e12
    return {
        # weak scaling: every rank pushes its own minibatch of B through the update each step, so the units processed per step
        # are `world` minibatch updates (one synchronous optimiser step on the global batch of world x B)
        "metric": "er_ddpg_updates_per_sec", "value": world * args.steps / (total_ms * 1e-3), "unit": "updates/s",
        "dtype": agent.learner.precision_name(),
        "ms_per_step": ms, "wall_s_timed_region": wall, "samples_per_sec": B * world * args.steps / (total_ms * 1e-3),
        "sync_steps_per_sec": args.steps / (total_ms * 1e-3),
        "config": workload_config("update", args, world),
        "timing": dict(TIMING, exchange="none" if world == 1 else args.exchange, dropout=args.dropout, precision=args.update_precision,
                       note="value counts per-GPU minibatch updates; sync_steps_per_sec the optimiser steps on the global batch"),
        "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": B * 80, "d2h_bytes_per_step": 16},
        "gpu_launches": (1 + agent.learner.launches_per_update(world > 1)) * args.steps,      # sample_gather + the update
        "dp_check": dp,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": None,
                     "peak_source": peaks["source"] + " (sustained: timed inside the whole step)",
                     "kernel": "whole update (sample + gather + U1-U5)",
                     "algorithmic_flop_per_sample": UPDATE_FLOP_PER_SAMPLE,
                     "floor_note": "%.2f GFLOP and about 5.5 MB per update: 4.3 us at the tensor peak, 0.8 us at the HBM peak; a chain of "
                                   "about 30 dependent layer passes bounds it by latency (DESIGN.md section 4.5)"
                                   % (UPDATE_FLOP_PER_SAMPLE * B / 1e9)},
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
    big = min(C, 4 << 20)      # 4 Mi random rows: 336 MB read + 336 MB written, far beyond the 126 MB L2
    idx_big = torch.randint(0, C, (big,), generator=g).to(device)
    out_big = torch.empty(big, 20, dtype=torch.float32, device=device)
    gb_ms = _kernel_ms(timer, lambda: mem.gather(idx_big, out=out_big), reps=10)
    nb = min(C, 4 << 20)
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
        "config": workload_config("replay", args, world), "timing": dict(TIMING),
        "e2e": {"value": n * world * args.steps / (e2e_ms * 1e-3), "unit": "transitions/s",
                "h2d_bytes_per_step": sum(h.numel() * h.element_size() for h in host), "d2h_bytes_per_step": B * 80},
        "gpu_launches": 1 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": None, "peak_source": peaks["source"], "kernel": "replay_insert_kernel",
                     "algorithmic_bytes_per_transition": 156},
        "sample_gather": {"batch": B, "ms": sg_ms, "note": "two launches, latency bound at this size"},
        "insert_big": {"transitions": nb, "ms": ib_ms, "achieved_gbs": 156 * nb / (ib_ms * 1e-3) / 1e9,
                       "frac": 156 * nb / (ib_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                       "note": "the same kernel on %d transitions: %.0f MB read + %.0f MB written, beyond what the 126 MB L2 can hold "
                               "back, so the device time is DRAM time" % (nb, 76e-6 * nb, 80e-6 * nb)},
        "gather_big": {"samples": big, "ms": gb_ms, "achieved_gbs": (8 + 160) * big / (gb_ms * 1e-3) / 1e9,
                       "frac": (8 + 160) * big / (gb_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                       "note": "random 80 B rows out of a %.0f MB ring" % (80e-6 * C)},
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
        "config": workload_config("n1", args, world),
        "timing": {"l2": "not flushed (host-paced launches)", "timing": "wall clock around the loop, device synchronised on both sides; "
                   "%d env-steps timed (host-driven loop: every call returns numpy like the reference)" % n_steps[0]},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * (2 + 10 + 19), "d2h_bytes_per_step": 4 * (10 + 2 + 3 + 3 + 2)},
        "gpu_launches": None,
        "roofline": {"bound": "hbm", "achieved": ENV_STEP_BYTES * v / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ENV_STEP_BYTES * v / 1e9 / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                     "kernel": "whole N = 1 step", "note": "latency bound: about 40 dependent launches and 8 host syncs per step"},
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
    dp = _dp_check(agent, allreduce, world, device)
    units = n * world * args.steps
    return {
        "metric": "env_steps_per_sec", "value": units / (total_ms * 1e-3), "unit": "env-steps/s",
        "dtype": "f64 simulator, %s transformer, f32 actor, %s update" % (agent.emotion_precision, agent.learner.precision_name()),
        "ms_per_step": total_ms / args.steps, "wall_s_timed_region": wall, "updates_per_sec": world * args.steps / (total_ms * 1e-3),
        "config": workload_config("erddpg", args, world),
        "timing": dict(TIMING, exchange="none" if world == 1 else args.exchange, dropout=args.dropout, precision=args.update_precision),
        "dp_check": dp,
        "ticks_per_sec": ticks_all / (total_ms * 1e-3), "mean_ticks_per_env_step": ticks_all / units,
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": obs_host.numel() * 4,
                "d2h_bytes_per_step": n * 13 + 16},
        "gpu_launches": (agent.launches_per_rollout_step() + 1 + agent.learner.launches_per_update(world > 1)) * args.steps,
        "roofline": _transformer_roofline(agent, n, tf_ms, total_ms / args.steps),
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
        "config": workload_config(kind, args, world), "timing": dict(TIMING, exchange="flat-bucket gradient all-reduce (NCCL)"),
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
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    head = "rollout" if args.workload == "all" else args.workload
    # CPU leg first (N = 1 only), in a child process, BEFORE this process touches CUDA or joins a process group: no GPU waits on it
    cpu = None
    if world == 1 and args.cpu_budget > 0:
        cpu = cpu_baseline(args, head)
    import torch
    import torch.distributed as dist
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
    out = wl[head](args, rank, world, device, timer)
    clocks = sampler.stop()
    others = []
    if args.workload == "all":
        # the other metrics of BASELINE.json on their own configs, compact (same timing rules)
        import copy
        for name, n_env, steps in (("sim", 4096, 20), ("update", 256, 20)):
            a2 = copy.copy(args)
            a2.n_env, a2.steps = n_env, steps
            o = wl[name](a2, rank, world, device, timer)
            o.pop("_issue", None)
            others.append({k: o[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "e2e", "gpu_launches", "roofline", "config", "timing",
                                             "dp_check", "sync_steps_per_sec", "issue_roofline")
                           if k in o} | {"steps": steps})
    if rank == 0:
        issue = out.pop("_issue", None)
        line = {"metric": out.pop("metric"), "value": out.pop("value"), "unit": out.pop("unit"), "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": out.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": out.pop("dtype"), "data": "synthetic"}
        line.update(out)
        line["clocks"] = clocks
        if issue:
            mhz = clocks.get("sm_mhz") or 1965.0
            ceiling = 148 * 4 * mhz * 1e6 / issue[1][0]
            line["issue_roofline"] = {"achieved_env_steps_per_s_per_gpu": issue[0], "ceiling": ceiling, "frac": issue[0] / ceiling,
                                      "warp_instr_per_env_step": issue[1][0], "warp_instr_source": issue[1][2],
                                      "warp_instr_stale": issue[1][1], "sm_mhz_used": mhz,
                                      "model": "148 SMs x 4 schedulers x 1 warp-instruction/clk / measured warp-instructions per env-step"}
        if cpu is not None:
            line["cpu_baseline"] = cpu
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
def cpu_baseline(args, head):
    """The reference arm of this very file on a bounded sample (about 10-30 s), run as a child process so that the forked
    per-core workers never see a CUDA context; returns the `cpu_baseline` object of the JSON line."""
    import subprocess
    steps, warm = 4, 1
    slice_s = max(0.5, min(4.0, args.cpu_budget / (steps + warm)))
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", head, "--steps", str(steps), "--warmup", str(warm),
           "--ref-slice", "%.3f" % slice_s, "--n-env", str(args.n_env), "--batch", str(args.batch), "--replay", str(args.replay),
           "--actions", args.actions]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        d = json.loads(lines[-1])
    except Exception as e:      # the line still prints; the failure is visible
        return {"value": None, "unit": None, "cores": None, "kind": "reference", "sample": "FAILED: %r" % (e,)}
    if "unavailable" in d:
        return {"value": None, "unit": None, "cores": None, "kind": "reference", "sample": "unavailable: " + d["unavailable"]}
    cb = d["cpu_baseline"]
    for k in ("ticks_per_sec", "c_port"):
        if k in d:
            cb[k] = d[k]
    return cb


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


def c_port_sim(args, budget_s=4.0):
    """Extra, beside the reference figure: the C port of the simulator path (oracle/erlfill_oracle.c), all host cores."""
    from oracle import oracle_lib as O
    cores = O.lib().orc_max_threads()
    rate, _, _ = _cpu_sim_once(args, 64 * cores, cores, 1)          # calibrate
    n_env = max(64 * cores, int(rate * budget_s / 4) // 64 * 64)
    v, tps, dt = _cpu_sim_once(args, n_env, cores, 4)
    return {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "ticks_per_sec": tps,
            "sample": "%d envs x 4 steps, C restatement of the simulator (pthreads over envs), %.1f s" % (n_env, dt)}


REF_ROWS = {"sim": "C1", "rollout": "C2", "update": "C3", "n1": "C4", "erddpg": "C4", "td3": "C4", "sac": "C4", "esac": "C4"}


def _ref_measure(args, name):
    """One reference measurement -> (value, unit, cores, sample text, extras).  BASELINE.md section 3."""
    from oracle import ref_bench as R
    T = args.ref_slice
    if name in ("sim", "rollout"):
        r = R.run_per_core(name, args.steps, args.warmup, T, actions=args.actions)
        what = ("controller.reset(); load_params(action); simulate_run()" if name == "sim" else
                "agent.act(state) -> env.step(action) -> agent.emotion.update(reward, next_state - state), networks in train() mode as shipped")
        sample = ("row %s: the unmodified reference's `%s` on %d single-env worker processes (one per host core, torch single-threaded); "
                  "a step = a %.2f s slice on every worker, %d env-steps in %d timed slices"
                  % (REF_ROWS[name], what, r["cores"], T, r["units"], args.steps))
        extras = {"ticks_per_sec": r["ticks_per_sec"], "mean_ticks_per_env_step": r["mean_ticks"]} if "ticks_per_sec" in r else {}
        return r["value"], "env-steps/s", r["cores"], sample, extras, T * 1e3
    if name == "update":
        r = R.run_update(args.batch, args.replay, args.steps, args.warmup)
        sample = ("row C3: the unmodified reference's DDPGAgent.learn() with batch_size = %d on a %d-transition PrioritizedReplayBuffer, one "
                  "process, %d torch threads; a step = one learn() call, %d timed" % (args.batch, args.replay, r["cores"], args.steps))
        return r["value"], "updates/s", r["cores"], sample, {}, r["seconds"] / args.steps * 1e3
    r = R.run_loop(name, max(args.steps, 20), args.warmup)
    sample = ("row C4: per-step body of the unmodified reference's train loop (%s) on ONE env with its own minibatch of 128, one process, "
              "%d torch threads; a step = one env-step + one update, %d timed (the batched loop has no single-process analogue "
              "in the reference)" % ({"n1": "train_ddpg", "erddpg": "train_ddpg", "td3": "train_td3", "sac": "train_sac",
                                      "esac": "train_erl_fill"}[name], r["cores"], r["units"]))
    return r["value"], "env-steps/s", r["cores"], sample, {}, r["seconds"] / r["units"] * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    head = "rollout" if args.workload == "all" else args.workload
    if head == "replay":
        print(json.dumps({"impl": "reference", "unavailable": "rows R1/R2 have no separate reference entry point (list append / np.random.choice "
                                                              "inside remember() and learn()): timed as part of --workload update / n1"}), flush=True)
        return
    from oracle import ref_bench as R
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "no reference tree: neither /root/reference nor baseline/_ref/erlfill_reference "
                                                              "(run `python oracle/vendor_ref.py` where /root/reference exists)"}), flush=True)
        return
    v, unit, cores, sample, extras, ms = _ref_measure(args, head)
    metric = "er_ddpg_updates_per_sec" if head == "update" else "env_steps_per_sec"
    out = {"impl": "reference", "metric": metric, "value": v, "unit": unit,
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 simulator, f32 networks (torch CPU)",
           "data": "synthetic", "config": workload_config(head, args, max(world, args.gpus)),
           "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "host_cpu_count": os.cpu_count()}
    out.update(extras)
    if head == "sim":
        out["c_port"] = c_port_sim(args)
    if args.workload == "all":
        import copy
        others = []
        for name, n_env, steps in (("sim", 4096, 4), ("update", 256, 5)):
            a2 = copy.copy(args)
            a2.n_env, a2.steps, a2.warmup = n_env, steps, 1
            v2, u2, c2, s2, e2, ms2 = _ref_measure(a2, name)
            others.append({"metric": "er_ddpg_updates_per_sec" if name == "update" else "env_steps_per_sec", "value": v2, "unit": u2,
                           "ms_per_step": ms2, "config": workload_config(name, a2, max(world, args.gpus)),
                           "cpu_baseline": {"value": v2, "unit": u2, "cores": c2, "kind": "reference", "sample": s2}, "steps": steps} | e2)
        out["other_workloads"] = others
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "sim", "rollout", "update", "td3", "sac", "esac", "replay", "n1", "erddpg"],
                    help="all = the rollout headline line (configs[2]) plus compact sim and update results under other_workloads")
    ap.add_argument("--n-env", type=int, default=None)
    ap.add_argument("--actions", default="pid", choices=["pid", "random"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--replay", type=int, default=None)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of reference CPU work for the cpu_baseline leg (0: skip)")
    ap.add_argument("--ref-slice", type=float, default=None, help="reference arm: seconds per step slice of the per-core workloads")
    ap.add_argument("--sim-mode", default="auto", choices=["auto", "warp", "thread"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="gradient exchange of the data-parallel update at N > 1 (csrc/peer.cu or ncclAllReduce)")
    ap.add_argument("--emotion-precision", default="bf16", choices=["bf16", "fp32"],
                    help="EmotionTransformer kernel of the rollout: tcgen05 bf16 (default) or the exact fp32 SIMT twin")
    ap.add_argument("--dropout", default="off", choices=["off", "philox"],
                    help="off = module.eval(); philox = train() mode like the reference, counter-based keep-masks")
    ap.add_argument("--update-precision", default="fast", choices=["fast", "exact"],
                    help="ER-DDPG update GEMMs: fast = tcgen05 bf16 operands / fp32 accumulate (stated tolerance); exact = 3xTF32 (1e-4 parity)")
    args = ap.parse_args()
    if args.n_env is None:
        args.n_env = DEFAULT_N_ENV[args.workload]
    if args.replay is None:
        args.replay = DEFAULT_REPLAY.get(args.workload, 1000000)
    if args.ref_slice is None:      # the whole reference run stays within about two minutes
        args.ref_slice = max(0.5, min(3.0, 100.0 / max(1, args.steps + args.warmup)))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
