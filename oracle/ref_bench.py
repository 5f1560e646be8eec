"""CPU legs of bench.py: the UNMODIFIED reference (through oracle/ref_shim.py) timed on the host cores.

Test infrastructure: only bench.py's `--impl reference` arm and its `cpu_baseline` leg import this file (BASELINE.md
section 3 is the plan it follows; rows C1-C4).  The reference tree comes from /root/reference in the build container and
from the byte-for-byte copy under baseline/_ref/ (oracle/vendor_ref.py) on the GPU box.

A "step" of a per-core workload (C1 sim, C2 rollout) is a fixed wall-clock slice T on every worker process: each worker
counts the iterations of the reference loop it FINISHES inside slice k, so units / (K * T) is exact up to one iteration per
worker per slice.  Workers are forked before any torch op and run torch single-threaded.  The update
(C3) and the full loops (C4) are one process with all host threads; a step is one learn() / one loop iteration.
"""
import contextlib
import io
import multiprocessing as mp
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def available():
    from oracle import ref_shim
    return ref_shim.available()


def _quiet():
    """The reference env prints twice per step (MutiConditionEnv.py:455-458): send stdout to /dev/null."""
    devnull = open(os.devnull, "w")
    sys.stdout = devnull
    return devnull


def _conditions():
    # experiment_runner.py:417-478 builds these dicts inside run_experiment_3 (not importable); conditions.py restates them
    import importlib.util
    spec = importlib.util.spec_from_file_location("erlfill_conditions", os.path.join(ROOT, "erl-fill_b200", "conditions.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _env_cfg(c):
    return {k: c[k] for k in ("target_weight", "target_weight_err", "target_time", "bounds")}


# ------------------------------------------------------------------------------------------------
# per-core workers (C1, C2)
# ------------------------------------------------------------------------------------------------
def _slice_loop(body, t0, T, n_slices, counts):
    """Run body() back to back; counts[k] = iterations finished inside slice k = [t0 + k T, t0 + (k + 1) T)."""
    end = t0 + n_slices * T
    while time.perf_counter() < t0:
        pass
    while True:
        body()
        now = time.perf_counter()
        if now >= end:
            k = int((now - t0) / T)
            if k < n_slices:
                counts[k] += 1
            break
        counts[int((now - t0) / T)] += 1


def _sim_worker(widx, t0, T, n_slices, q, actions):
    """C1: controller.reset(); controller.load_params(action); controller.simulate_run()  (VirtualWeightController.py:16-25,64-83,127-244)."""
    _quiet()
    from oracle import ref_shim
    V = ref_shim.install()
    K = _conditions()
    c = K.SINGLE_25KG
    ctl = V.VirtualWeightController()
    order = ("fast_weight", "medium_weight", "slow_weight", "fast_opening", "medium_opening", "slow_opening", "fast_delay",
             "medium_delay", "slow_delay", "unload_delay")
    rng = np.random.RandomState(1234 + widx)
    random.seed(widx)
    ticks = [0]

    def body():
        if actions == "pid":
            a = K.PID_EFFECTIVE_ACTION
        else:
            a = [c["bounds"][k][0] + rng.uniform() * (c["bounds"][k][1] - c["bounds"][k][0]) for k in order]
        p = dict(zip(order, a))
        p["target_weight"] = c["target_weight"]
        ctl.reset()
        ctl.load_params(p)
        out = ctl.simulate_run()
        ticks[0] += len(out[2])

    counts = [0] * n_slices
    _slice_loop(body, t0, T, n_slices, counts)
    q.put((widx, counts, ticks[0]))


def _rollout_worker(widx, t0, T, n_slices, q, max_steps):
    """C2: per env-step `agent.act(state)` -> `env.step(action)` -> `agent.emotion.update(reward, next_state - state)` in
    train_ddpg's episode structure (ddpg_agent.py:628-642, 676-678), no remember / learn, networks in train() mode as shipped."""
    _quiet()
    import torch
    torch.set_num_threads(1)
    from oracle import ref_shim
    ref_shim.install()
    import MutiConditionEnv as ME
    import ddpg_agent as DA
    from EmotionModule import EmotionModule
    import warnings
    warnings.filterwarnings("ignore")
    K = _conditions()
    conds = list(K.EXPERIMENT_3.values())
    random.seed(widx); np.random.seed(widx); torch.manual_seed(0)
    env = ME.MultiConditionWeightEnv(_env_cfg(conds[widx % len(conds)]))
    env.max_steps = max_steps
    agent = DA.DDPGAgent(env, device="cpu", use_td_error=True)
    agent.emotion = EmotionModule(device="cpu")
    env.attach_agent(agent)
    st = {"state": env.reset(), "step": 0}
    agent.ou_noise.reset(); agent.ou_noise.anneal()

    def body():
        state = st["state"]
        action = np.transpose(agent.act(state))
        next_state, reward, done, info = env.step(action)
        agent.emotion.update(reward, next_state - state)
        st["state"] = next_state
        st["step"] += 1
        if done or st["step"] >= max_steps:
            st["state"] = env.reset()
            st["step"] = 0
            agent.ou_noise.reset(); agent.ou_noise.anneal()

    body()       # first call outside the slices (lazy initialisation)
    counts = [0] * n_slices
    _slice_loop(body, t0, T, n_slices, counts)
    q.put((widx, counts, 0))


def run_per_core(kind, steps, warmup, slice_s, cores=None, actions="pid", max_steps=100):
    """Returns dict(value = units/s over the timed slices, cores, units, seconds, ticks_per_sec?)."""
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    if kind == "rollout":
        import torch           # imported (no op executed) before the fork so that the workers share the loaded pages
        torch.set_num_threads(1)
    q = ctx.Queue()
    n_slices = warmup + steps
    t0 = time.perf_counter() + (12.0 if kind == "rollout" else 1.0)      # workers construct env + agent before the first slice
    procs = []
    for w in range(cores):
        args = (w, t0, slice_s, n_slices, q, actions) if kind == "sim" else (w, t0, slice_s, n_slices, q, max_steps)
        p = ctx.Process(target=_sim_worker if kind == "sim" else _rollout_worker, args=args, daemon=True)
        p.start()
        procs.append(p)
    res = [q.get(timeout=t0 - time.perf_counter() + n_slices * slice_s + 120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    units = sum(sum(c[warmup:]) for _, c, _ in res)
    total_units = sum(sum(c) for _, c, _ in res)
    ticks = sum(t for _, _, t in res)
    out = {"value": units / (steps * slice_s), "cores": cores, "units": units, "seconds": steps * slice_s}
    if kind == "sim" and total_units:
        out["ticks_per_sec"] = out["value"] * ticks / total_units
        out["mean_ticks"] = ticks / total_units
    return out


# ------------------------------------------------------------------------------------------------
# one process, all host threads (C3, C4)
# ------------------------------------------------------------------------------------------------
def _make_ddpg(K, cond_idx=0):
    import torch
    from oracle import ref_shim
    ref_shim.install()
    import MutiConditionEnv as ME
    import ddpg_agent as DA
    from EmotionModule import EmotionModule
    conds = list(K.EXPERIMENT_3.values())
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    env = ME.MultiConditionWeightEnv(_env_cfg(conds[cond_idx]))
    agent = DA.DDPGAgent(env, device="cpu", use_td_error=True)
    agent.emotion = EmotionModule(device="cpu")
    env.attach_agent(agent)
    return env, agent


def run_update(batch, capacity, steps, warmup):
    """C3: DDPGAgent.learn() (ddpg_agent.py:422-512) with agent.batch_size = batch on a buffer pre-filled with `capacity`
    synthetic transitions (the distribution of bench.py's configs[3] replay; priorities all epsilon as the reference's own
    remember() leaves them)."""
    import warnings
    import torch
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    K = _conditions()
    with contextlib.redirect_stdout(io.StringIO()):
        env, agent = _make_ddpg(K)
    agent.batch_size = batch
    rng = np.random.RandomState(100)
    chunk = 1 << 16
    buf, pri = agent.memory.buffer, agent.memory.priorities
    eps = agent.memory.epsilon
    for o in range(0, capacity, chunk):
        m = min(chunk, capacity - o)
        s = (rng.rand(m, 2) * 5).astype(np.float32); a = (rng.rand(m, 10) * 2 - 1)
        r = rng.rand(m) * 9000 - 3000; s2 = (rng.rand(m, 2) * 5).astype(np.float32)
        d = rng.rand(m) < 0.01; e = rng.rand(m, 3).astype(np.float32)
        for i in range(m):
            buf.append((s[i], a[i], float(r[i]), s2[i], bool(d[i]), e[i], float("nan")))
        pri.extend([eps] * m)
    for _ in range(20):        # a history for learn()'s get_emotion() (ddpg_agent.py:466)
        agent.emotion.update(float(rng.rand() * 100), rng.rand(2) * 0.1)
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(warmup):
            agent.learn()
        t0 = time.perf_counter()
        for _ in range(steps):
            agent.learn()
        dt = time.perf_counter() - t0
    return {"value": steps / dt, "cores": torch.get_num_threads(), "units": steps, "seconds": dt}


def run_loop(kind, steps, warmup, max_steps=100):
    """C4: the per-step body of train_ddpg (ddpg_agent.py:636-678) / train_td3 (td3_agent.py:424-460) / train_sac
    (sac_agent.py:362-409) on one MultiConditionWeightEnv, reference default minibatch 128, without the TensorBoard calls."""
    import warnings
    import torch
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    K = _conditions()
    from oracle import ref_shim
    ref_shim.install()
    import MutiConditionEnv as ME
    conds = list(K.EXPERIMENT_3.values())
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        if kind in ("n1", "erddpg"):
            env, agent = _make_ddpg(K)
        else:
            env = ME.MultiConditionWeightEnv(_env_cfg(conds[0]))
            if kind == "td3":
                import td3_agent as TA
                agent = TA.TD3Agent(env, device="cpu")
            elif kind == "sac":
                import sac_agent as SA
                agent = SA.SACAgent(env, device="cpu")
            else:
                import ERL_Fill_agent as EA
                from EmotionModule import EmotionModule
                agent = EA.EmotionSACAgent(env, device="cpu")
                agent.emotion = EmotionModule(device="cpu")
        env.max_steps = max_steps
        st = {"state": env.reset(), "step": 0}

        def body():
            state = st["state"]
            if kind in ("n1", "erddpg"):
                action = np.transpose(agent.act(state))
                next_state, reward, done, info = env.step(action)
                agent.emotion.update(reward, next_state - state)
                agent.remember(state, action, reward, next_state, done)
                agent.learn()
                agent.emotion.get_emotion()
            elif kind == "td3":
                action = agent.select_action(state)
                next_state, reward, done, info = env.step(action)
                agent.remember(state, action, reward, next_state, float(done))
                agent.train_step()
            elif kind == "sac":
                action = agent.act(state)
                next_state, reward, done, info = env.step(action)
                agent.remember(state, action, reward, next_state, float(done))
                agent.update()
            else:
                action = agent.act(state)
                next_state, reward, done, info = env.step(action)
                agent.emotion.update(reward, action)
                agent.remember(state, action, reward, next_state, float(done))
                agent.update()
            st["state"] = next_state
            st["step"] += 1
            if done or st["step"] >= max_steps:
                st["state"] = env.reset()
                st["step"] = 0
                if kind in ("n1", "erddpg"):
                    agent.ou_noise.reset(); agent.ou_noise.anneal()
                    agent.train_step += 1

        for _ in range(max(warmup, 140)):       # past the first minibatch (128 transitions)
            body()
        t0 = time.perf_counter()
        for _ in range(steps):
            body()
        dt = time.perf_counter() - t0
    return {"value": steps / dt, "cores": torch.get_num_threads(), "units": steps, "seconds": dt}
