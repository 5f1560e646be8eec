"""Generate tests/golden/*.npz by running the UNMODIFIED reference (container only).

Run from the repo root:   python oracle/gen_golden.py [sim] [env] [emotion] [nets] [learn] [td3sac] [loop] [esac]

The reference (/root/reference) is imported through oracle/ref_shim.py (pymodbus stub +
frozen clock).  All randomness is CPython `random` seeded per case,
so a test can regenerate the identical Mersenne-Twister stream with `random.Random(seed)`
and does not need the reference at run time; the fixtures hold only inputs, seeds and the
reference's outputs.  Fixture provenance is recorded inside each file (`meta`).
"""
import collections
import contextlib
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

PARAM_ORDER = ["target_weight", "fast_weight", "medium_weight", "slow_weight", "fast_delay",
               "medium_delay", "slow_delay", "fast_opening", "medium_opening", "slow_opening",
               "unload_delay"]
ACTION_ORDER = ["fast_weight", "medium_weight", "slow_weight", "fast_opening", "medium_opening",
                "slow_opening", "fast_delay", "medium_delay", "slow_delay", "unload_delay"]

# bounds tables (reference: WeightEnv.py:32-46, MutiConditionEnv.py:47-61, experiment_runner.py:417-478)
BOUNDS = {
    "single_25kg": dict(target_weight=25000, target_weight_err=25, target_time=2.5, bounds={
        "fast_weight": (18000, 23000), "medium_weight": (0, 1), "slow_weight": (24900, 25025),
        "fast_opening": (35, 60), "medium_opening": (3, 5), "slow_opening": (3, 7),
        "fast_delay": (100, 300), "medium_delay": (100, 200), "slow_delay": (100, 200),
        "unload_delay": (300, 500)}),
    "multi_default": dict(target_weight=25000, target_weight_err=25, target_time=2.5, bounds={
        "fast_weight": (7000, 18000), "medium_weight": (0, 1), "slow_weight": (24900, 25000),
        "fast_opening": (35, 75), "medium_opening": (3, 5), "slow_opening": (5, 20),
        "fast_delay": (100, 300), "medium_delay": (100, 200), "slow_delay": (100, 200),
        "unload_delay": (300, 500)}),
    "variant_25kg": dict(target_weight=25000, target_weight_err=25, target_time=2.5, bounds={
        "fast_weight": (7000, 18000), "medium_weight": (0, 1), "slow_weight": (24900, 25000),
        "fast_opening": (35, 75), "medium_opening": (3, 5), "slow_opening": (5, 20),
        "fast_delay": (100, 300), "medium_delay": (100, 200), "slow_delay": (100, 200),
        "unload_delay": (300, 500)}),
    "variant_20kg": dict(target_weight=20000, target_weight_err=20, target_time=2.0, bounds={
        "fast_weight": (6000, 15000), "medium_weight": (0, 1), "slow_weight": (19800, 20000),
        "fast_opening": (30, 70), "medium_opening": (3, 5), "slow_opening": (5, 18),
        "fast_delay": (80, 250), "medium_delay": (80, 180), "slow_delay": (80, 180),
        "unload_delay": (280, 480)}),
    "variant_15kg": dict(target_weight=15000, target_weight_err=15, target_time=1.5, bounds={
        "fast_weight": (5000, 12000), "medium_weight": (0, 1), "slow_weight": (14800, 15000),
        "fast_opening": (25, 65), "medium_opening": (2, 4), "slow_opening": (4, 15),
        "fast_delay": (60, 200), "medium_delay": (60, 150), "slow_delay": (60, 150),
        "unload_delay": (250, 450)}),
}


def _meta(what):
    import platform
    import numpy
    return json.dumps({"what": what, "generator": "oracle/gen_golden.py",
                       "reference": "Harrison-Ma/ERL-Fill @ /root/reference (unmodified, pymodbus stubbed, frozen clock)",
                       "python": platform.python_version(), "numpy": numpy.__version__})


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ----------------------------------------------------------------------------------------
def gen_sim():
    """E1 known-answer vectors: (11 int params, seed) -> (iterations, total_time, final_weight)."""
    V = ref_shim.install()
    rng = np.random.RandomState(1234)
    cases = []
    # the SURVEY section-4 KATs first
    kat = dict(target_weight=25000, fast_weight=12000, medium_weight=0, slow_weight=24950,
               fast_delay=200, medium_delay=150, slow_delay=150, fast_opening=55, medium_opening=4,
               slow_opening=12, unload_delay=400)
    for s in (0, 1, 2):
        cases.append((dict(kat), s))
    # random integer params inside every shipped bounds table
    for name, cfg in BOUNDS.items():
        for _ in range(60):
            p = {"target_weight": cfg["target_weight"]}
            for k, (lo, hi) in cfg["bounds"].items():
                p[k] = int(rng.randint(lo, hi + 1))
            cases.append((p, int(rng.randint(0, 2**31 - 1))))
    # config-2 constant action
    const = dict(target_weight=25000, fast_weight=18000, medium_weight=0, slow_weight=24900,
                 fast_opening=35, medium_opening=3, slow_opening=3, fast_delay=100, medium_delay=100,
                 slow_delay=100, unload_delay=300)
    for s in range(5):
        cases.append((dict(const), 100 + s))
    # edge cases outside the shipped bounds: reachable medium stage, reversing motor, first-tick
    # delay taken in the medium/slow stage, stagnation stop, iteration-cap stop, tiny targets
    edge = [
        dict(target_weight=25000, fast_weight=8000, medium_weight=16000, slow_weight=24900, fast_delay=120,
             medium_delay=130, slow_delay=140, fast_opening=70, medium_opening=30, slow_opening=6, unload_delay=333),
        dict(target_weight=25000, fast_weight=8000, medium_weight=16000, slow_weight=24900, fast_delay=120,
             medium_delay=130, slow_delay=140, fast_opening=20, medium_opening=60, slow_opening=90, unload_delay=333),
        dict(target_weight=1000, fast_weight=0, medium_weight=500, slow_weight=900, fast_delay=11,
             medium_delay=22, slow_delay=33, fast_opening=40, medium_opening=10, slow_opening=3, unload_delay=5),
        dict(target_weight=1000, fast_weight=0, medium_weight=0, slow_weight=900, fast_delay=11,
             medium_delay=22, slow_delay=33, fast_opening=40, medium_opening=10, slow_opening=3, unload_delay=5),
        dict(target_weight=25000, fast_weight=12000, medium_weight=0, slow_weight=24950, fast_delay=200,
             medium_delay=150, slow_delay=150, fast_opening=0, medium_opening=4, slow_opening=12, unload_delay=400),
        dict(target_weight=25000, fast_weight=12000, medium_weight=0, slow_weight=24950, fast_delay=200,
             medium_delay=150, slow_delay=150, fast_opening=1, medium_opening=4, slow_opening=1, unload_delay=400),
        dict(target_weight=25000, fast_weight=2000, medium_weight=0, slow_weight=24950, fast_delay=200,
             medium_delay=150, slow_delay=150, fast_opening=50, medium_opening=4, slow_opening=2, unload_delay=400),
        dict(target_weight=90000, fast_weight=60000, medium_weight=0, slow_weight=89000, fast_delay=0,
             medium_delay=0, slow_delay=0, fast_opening=100, medium_opening=4, slow_opening=3, unload_delay=0),
        dict(target_weight=500, fast_weight=0, medium_weight=0, slow_weight=0, fast_delay=1,
             medium_delay=2, slow_delay=3, fast_opening=50, medium_opening=4, slow_opening=3, unload_delay=7),
        dict(target_weight=25000, fast_weight=12000, medium_weight=0, slow_weight=24950, fast_delay=200,
             medium_delay=150, slow_delay=150, fast_opening=2, medium_opening=4, slow_opening=99, unload_delay=400),
    ]
    for i, p in enumerate(edge):
        for s in (7, 8):
            cases.append((dict(p), 1000 + 10 * i + s))

    params = np.zeros((len(cases), 11), np.int32)
    seeds = np.zeros(len(cases), np.int64)
    iters = np.zeros(len(cases), np.int32)
    total_time = np.zeros(len(cases), np.float64)
    final_weight = np.zeros(len(cases), np.float64)
    peak_open = np.zeros(len(cases), np.float64)
    peak_speed = np.zeros(len(cases), np.float64)
    c = V.VirtualWeightController()
    for i, (p, s) in enumerate(cases):
        random.seed(s)
        c.reset()
        c.load_params(p)
        with _quiet():
            speeds, openings, weights, tt, fw = c.simulate_run()
        params[i] = [p[k] for k in PARAM_ORDER]
        seeds[i] = s
        iters[i] = len(weights)
        total_time[i] = tt
        final_weight[i] = fw
        peak_open[i] = max(openings)
        peak_speed[i] = max(speeds)
    np.savez_compressed(os.path.join(GOLD, "sim_golden.npz"), params=params, seeds=seeds, iters=iters,
                        total_time=total_time, final_weight=final_weight, peak_open=peak_open,
                        peak_speed=peak_speed, meta=_meta("E1 simulate_run KATs"))
    print("sim:", len(cases), "cases; iters min/mean/max", iters.min(), iters.mean(), iters.max(),
          "peak speed", peak_speed.max())


def gen_onboard():
    """Row N1 known-answer vectors.  The on-board per-tick models are methods of the env classes (WeightEnv.py:176-229,
    MutiConditionEnv.py:184-226); in the reference their stage logic runs in an external PLC polled over Modbus, so the
    scripted stand-in is the reference's OWN offline loop: VirtualWeightController.simulate_run is run unmodified on a
    subclass whose two per-tick methods and three state attributes forward to a real env instance."""
    V = ref_shim.install()
    import WeightEnv as W
    import MutiConditionEnv as M

    def controller_on(env):
        class Onboard(V.VirtualWeightController):
            curr_weight = property(lambda self: env.current_weight, lambda self, v: setattr(env, "current_weight", v))
            current_opening = property(lambda self: env.current_opening, lambda self, v: setattr(env, "current_opening", v))
            current_speed = property(lambda self: env.current_speed, lambda self, v: setattr(env, "current_speed", v))

            def motor_simulation(self, target_pos):
                env.motor_simulation(target_pos)

            def weight_simulation(self, opening):
                env.weight_simulation(opening, env.material_coefficient)
        return Onboard()

    rng = np.random.RandomState(4321)
    cases = []
    with _quiet():
        envs = {"single": W.WeightEnv(), "multi": M.MultiConditionWeightEnv(None)}
    for name, cfg in BOUNDS.items():
        which = "single" if name == "single_25kg" else "multi"
        for _ in range(40):
            p = {"target_weight": cfg["target_weight"]}
            for k, (lo, hi) in cfg["bounds"].items():
                p[k] = int(rng.randint(lo, hi + 1))
            cases.append((which, p, int(rng.randint(0, 2**31 - 1))))
    # outside the shipped bounds: long moves where the 10000 speed cap binds, a reversing motor, tiny / zero openings
    # (stagnation stop), an unreachable threshold (iteration cap), a reachable medium stage
    edge = [
        dict(target_weight=90000, fast_weight=60000, medium_weight=0, slow_weight=89000, fast_delay=0, medium_delay=0, slow_delay=0,
             fast_opening=100, medium_opening=4, slow_opening=3, unload_delay=0),
        dict(target_weight=25000, fast_weight=8000, medium_weight=16000, slow_weight=24900, fast_delay=120, medium_delay=130,
             slow_delay=140, fast_opening=250, medium_opening=30, slow_opening=6, unload_delay=333),
        dict(target_weight=25000, fast_weight=8000, medium_weight=16000, slow_weight=24900, fast_delay=120, medium_delay=130,
             slow_delay=140, fast_opening=20, medium_opening=60, slow_opening=90, unload_delay=333),
        dict(target_weight=25000, fast_weight=12000, medium_weight=0, slow_weight=24950, fast_delay=200, medium_delay=150,
             slow_delay=150, fast_opening=0, medium_opening=4, slow_opening=12, unload_delay=400),
        dict(target_weight=25000, fast_weight=12000, medium_weight=0, slow_weight=24950, fast_delay=200, medium_delay=150,
             slow_delay=150, fast_opening=1, medium_opening=4, slow_opening=1, unload_delay=400),
        dict(target_weight=25000, fast_weight=2000, medium_weight=0, slow_weight=24950, fast_delay=200, medium_delay=150,
             slow_delay=150, fast_opening=50, medium_opening=4, slow_opening=2, unload_delay=400),
        dict(target_weight=500, fast_weight=0, medium_weight=0, slow_weight=0, fast_delay=1, medium_delay=2, slow_delay=3,
             fast_opening=50, medium_opening=4, slow_opening=3, unload_delay=7),
    ]
    for i, p in enumerate(edge):
        for s in (7, 8):
            cases.append(("single" if i % 2 else "multi", dict(p), 2000 + 10 * i + s))
    n = len(cases)
    out = dict(params=np.zeros((n, 11), np.int32), seeds=np.zeros(n, np.int64), iters=np.zeros(n, np.int32),
               total_time=np.zeros(n), final_weight=np.zeros(n), peak_speed=np.zeros(n), min_speed=np.zeros(n),
               env=np.array([c[0] for c in cases]))
    for i, (which, p, s) in enumerate(cases):
        env = envs[which]
        env.current_weight = 0; env.current_speed = 0; env.current_opening = 0       # WeightEnv.reset (:141-143)
        c = controller_on(env)
        random.seed(s)
        c.reset()
        c.load_params(p)
        with _quiet():
            speeds, openings, weights, tt, fw = c.simulate_run()
        out["params"][i] = [p[k] for k in PARAM_ORDER]
        out["seeds"][i], out["iters"][i], out["total_time"][i], out["final_weight"][i] = s, len(weights), tt, fw
        out["peak_speed"][i], out["min_speed"][i] = max(speeds), min(speeds)
    out["meta"] = _meta("N1 on-board motor_simulation / weight_simulation of the env classes under simulate_run's stage logic")
    np.savez_compressed(os.path.join(GOLD, "onboard_golden.npz"), **out)
    print("onboard:", n, "cases; iters min/mean/max", out["iters"].min(), out["iters"].mean(), out["iters"].max(),
          "speed range", out["min_speed"].min(), out["peak_speed"].max())


# ----------------------------------------------------------------------------------------
def _make_env(kind, name):
    ref_shim.install()
    if kind == "single":
        import WeightEnv as W
        env = W.WeightEnv()
    else:
        import MutiConditionEnv as M
        cfg = BOUNDS[name]
        if name == "multi_default":
            env = M.MultiConditionWeightEnv({k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time")})
        else:
            env = M.MultiConditionWeightEnv({**{k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time")},
                                             "bounds": dict(cfg["bounds"])})
    return env


def gen_env():
    """E2/E3/E4 vectors: whole episodes of reset()+step() with seeded noise and recorded actions."""
    out = {}
    runs = []
    rng = np.random.RandomState(4321)
    plan = [("single", "single_25kg", 12, 40), ("multi", "multi_default", 12, 40),
            ("multi", "variant_25kg", 8, 30), ("multi", "variant_20kg", 8, 30),
            ("multi", "variant_15kg", 8, 30),
            # long good-action runs that end by the mean(recent_errors) < 25 rule
            ("multi", "variant_25kg", 60, 70), ("single", "single_25kg", 60, 70)]
    for ri, (kind, name, max_steps, n_steps) in enumerate(plan):
        env = _make_env(kind, name)
        env.max_steps = max_steps
        cfg = BOUNDS[name]
        lo = np.array([cfg["bounds"][k][0] for k in ACTION_ORDER], np.float64)
        hi = np.array([cfg["bounds"][k][1] for k in ACTION_ORDER], np.float64)
        seed = 9000 + ri
        random.seed(seed)
        good = max_steps >= 60
        rec = dict(actions=[], next_state=[], reward=[], reward_is_f32=[], done=[], final_weight=[],
                   total_time=[], weight_error=[], reset_state=[], reset_before=[])
        state = env.reset()
        rec["reset_state"].append(state)
        need_reset = False
        for t in range(n_steps):
            if need_reset:
                state = env.reset()
                rec["reset_state"].append(state)
            rec["reset_before"].append(need_reset or t == 0)
            if good:
                # near-optimal action found by scanning the reference: small final error
                a = (lo + hi) / 2
                a[2] = cfg["target_weight"] - 12.0   # slow_weight
                a[5] = lo[5] + 0.2 * (hi[5] - lo[5])
                a = a + rng.uniform(-0.4, 0.4, 10)
            else:
                span = hi - lo
                a = lo - 0.1 * span + rng.uniform(0, 1, 10) * 1.2 * span   # 10% outside on both sides
                if t % 7 == 3:
                    a[rng.randint(10)] = lo[rng.randint(10)]               # exactly-on-bound values
            a32 = a.astype(np.float32)
            with _quiet():
                ns, r, d, info = env.step(a32.copy())
            rec["actions"].append(a32)
            rec["next_state"].append(ns)
            rec["reward"].append(float(r))
            rec["reward_is_f32"].append(isinstance(r, np.float32))
            rec["done"].append(bool(d))
            rec["final_weight"].append(info["final_weight"])
            rec["total_time"].append(info["total_time"])
            rec["weight_error"].append(info["weight_error"])
            need_reset = bool(d)
        key = f"run{ri}"
        out[key + "_kind"] = kind
        out[key + "_name"] = name
        out[key + "_seed"] = seed
        out[key + "_max_steps"] = max_steps
        out[key + "_actions"] = np.array(rec["actions"], np.float32)
        out[key + "_next_state"] = np.array(rec["next_state"], np.float32)
        out[key + "_reward"] = np.array(rec["reward"], np.float64)
        out[key + "_reward_is_f32"] = np.array(rec["reward_is_f32"], np.bool_)
        out[key + "_done"] = np.array(rec["done"], np.bool_)
        out[key + "_final_weight"] = np.array(rec["final_weight"], np.float64)
        out[key + "_total_time"] = np.array(rec["total_time"], np.float64)
        out[key + "_weight_error"] = np.array(rec["weight_error"], np.float64)
        out[key + "_reset_state"] = np.array(rec["reset_state"], np.float32)
        out[key + "_reset_before"] = np.array(rec["reset_before"], np.bool_)
        runs.append(key)
        print(key, kind, name, "steps", n_steps, "dones", int(np.sum(rec["done"])),
              "f32 rewards", int(np.sum(rec["reward_is_f32"])),
              "mean err", float(np.mean(rec["weight_error"])))
    out["runs"] = np.array(runs)
    out["bounds_json"] = json.dumps(BOUNDS)
    out["meta"] = _meta("E2/E3/E4 env.reset/env.step episodes")
    np.savez_compressed(os.path.join(GOLD, "env_golden.npz"), **out)


# ----------------------------------------------------------------------------------------
def gen_emotion():
    """M1 vectors: EmotionModule.update over a reward/movement sequence (history + current_emotion)."""
    ref_shim.install()
    import torch
    import EmotionModule as E
    torch.manual_seed(0)
    em = E.EmotionModule(device="cpu")
    rng = np.random.RandomState(77)
    rewards, moves, cur, returned = [], [], [], []
    for t in range(64):
        r = float(rng.choice([rng.uniform(-3000, 6000), rng.uniform(-2, 2), rng.uniform(-0.01, 0.01)]))
        if t % 2 == 0:
            r = np.float32(r)    # MutiConditionEnv returns np.float32 rewards on the [f32] path
        mv = (rng.uniform(-1, 1, 2) * rng.choice([0.01, 0.3, 3.0])).astype(np.float32)
        out = em.update(r, mv)
        rewards.append(float(r)); moves.append(mv); cur.append(np.array(em.current_emotion, np.float64))
        returned.append(np.array(out, np.float64))
    hist = np.stack([h.numpy() for h in em.transformer.history])
    np.savez_compressed(os.path.join(GOLD, "emotion_update_golden.npz"), rewards=np.array(rewards, np.float64),
                        reward_was_f32=np.array([t % 2 == 0 for t in range(64)]),
                        moves=np.array(moves, np.float32), current=np.array(cur), history_last20=hist,
                        meta=_meta("M1 EmotionModule.update sequence"))
    print("emotion: 64 updates, final", em.current_emotion)


# ----------------------------------------------------------------------------------------
def _build_agent(seed=0):
    """Reference construction order (experiment_runner.py:481-484,508-510): env, DDPGAgent, EmotionModule."""
    ref_shim.install()
    import torch
    import MutiConditionEnv as M
    import ddpg_agent as D
    import EmotionModule as E
    torch.manual_seed(seed)
    cfg = BOUNDS["variant_25kg"]
    env = M.MultiConditionWeightEnv({**{k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time")},
                                     "bounds": dict(cfg["bounds"])})
    agent = D.DDPGAgent(env, device="cpu", use_td_error=True)
    agent.emotion = E.EmotionModule(device="cpu")
    env.attach_agent(agent)
    return env, agent


def _sd(module, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def _replay_masks(seed, shapes_p):
    """Re-draw, with torch's CPU generator, the keep-masks at::dropout draws for the given shapes in
    the given order: empty_like(x).bernoulli_(1 - p)."""
    import torch
    torch.manual_seed(seed)
    return [torch.empty(shape).bernoulli_(1 - p) for shape, p in shapes_p]


def gen_nets():
    """A1 / U1 / M2 vectors from the real nn.Modules, eval() and train() (dropout masks replayed)."""
    import torch
    env, agent = _build_agent(0)
    out = {}
    out.update(_sd(agent.actor, "actor."))
    out.update(_sd(agent.critic, "critic."))
    out.update(_sd(agent.emotion.transformer, "trans."))
    g = torch.Generator().manual_seed(5)
    B = 48
    state = torch.rand(B, 2, generator=g) * 5
    emotion = torch.rand(B, 3, generator=g)
    emotion[0] = torch.tensor([0.5, 0.5, 0.0])
    action = torch.rand(B, 10, generator=g) * 2 - 1
    seqs = torch.rand(12, 20, 3, generator=g)
    seqs[0, :15] = seqs[0, 15]                      # front-padded history
    seq1 = torch.tensor([[[0.5, 0.5, 0.0]]])        # empty history: S = 1
    seq7 = torch.rand(3, 7, 3, generator=g)         # shorter sequences are legal for the module
    out.update(state=state.numpy(), emotion=emotion.numpy(), action=action.numpy(), seqs=seqs.numpy(),
               seq1=seq1.numpy(), seq7=seq7.numpy())
    tr = agent.emotion.transformer
    with torch.no_grad():
        for m in (agent.actor, agent.critic, tr):
            m.eval()
        out["actor_eval"] = agent.actor(state, emotion).numpy()
        out["critic_eval"] = agent.critic(state, action, emotion).numpy()
        # the fused "fast path" of nn.TransformerEncoder only exists in eval(); the reference never
        # calls eval(), so the eval golden is taken through the same (slow) code path
        torch.backends.mha.set_fastpath_enabled(False)
        out["trans_eval"] = tr(seqs).numpy()
        out["trans_eval_s1"] = tr(seq1).numpy()
        out["trans_eval_s7"] = tr(seq7).numpy()
        for m in (agent.actor, agent.critic, tr):
            m.train()
        torch.manual_seed(101)
        out["actor_train"] = agent.actor(state, emotion).numpy()
        am = _replay_masks(101, [((B, 128), 0.2), ((B, 128), 0.2)])
        out["actor_m1"], out["actor_m2"] = am[0].numpy(), am[1].numpy()
        torch.manual_seed(102)
        out["critic_train"] = agent.critic(state, action, emotion).numpy()
        cm = _replay_masks(102, [((B, 256), 0.2), ((B, 256), 0.2)])
        out["critic_m1"], out["critic_m2"] = cm[0].numpy(), cm[1].numpy()
        torch.manual_seed(103)
        nb = seqs.shape[0]
        out["trans_train"] = tr(seqs).numpy()
        shapes = []
        for _ in range(4):
            # module layout is [S, B, .] (batch_first=False); attention probs are [B, heads, S, S]
            shapes += [((nb, 4, 20, 20), 0.3), ((20, nb, 32), 0.3), ((20, nb, 2048), 0.3), ((20, nb, 32), 0.3)]
        tm = _replay_masks(103, shapes)
        for l in range(4):
            out[f"trans_attn{l}"] = tm[4 * l].numpy().astype(np.uint8)
            out[f"trans_d1_{l}"] = tm[4 * l + 1].permute(1, 0, 2).numpy().astype(np.uint8)
            out[f"trans_ff{l}"] = np.packbits(tm[4 * l + 2].permute(1, 0, 2).numpy().astype(np.uint8), axis=-1)
            out[f"trans_d2_{l}"] = tm[4 * l + 3].permute(1, 0, 2).numpy().astype(np.uint8)
    out["meta"] = _meta("A1/U1/M2 forwards of the reference nn.Modules (torch.manual_seed(0) construction)")
    np.savez_compressed(os.path.join(GOLD, "nets_golden.npz"), **out)
    print("nets: actor_eval[0]", out["actor_eval"][0][:4], "trans_eval[0]", out["trans_eval"][0],
          "s1", out["trans_eval_s1"])


def gen_learn():
    """U2-U5 vectors: DDPGAgent.learn() three times on an injected replay buffer, eval-mode networks
    (dropout off) so the update is deterministic; weights, Adam state, losses and LRs recorded."""
    import torch
    env, agent = _build_agent(0)
    for m in (agent.actor, agent.actor_target, agent.critic, agent.critic_target, agent.emotion.transformer):
        m.eval()
    torch.backends.mha.set_fastpath_enabled(False)
    g = np.random.RandomState(9)
    B = 128
    agent.batch_size = B
    n_buf = 600
    buf = dict(states=(g.rand(n_buf, 2) * 5).astype(np.float32), actions=(g.rand(n_buf, 10) * 2 - 1).astype(np.float32),
               rewards=g.uniform(-3000, 6000, n_buf), next_states=(g.rand(n_buf, 2) * 5).astype(np.float32),
               dones=(g.rand(n_buf) < 0.05), emotions=g.rand(n_buf, 3).astype(np.float32))
    for i in range(n_buf):
        agent.memory.buffer.append((buf["states"][i], buf["actions"][i], float(buf["rewards"][i]),
                                    buf["next_states"][i], bool(buf["dones"][i]), buf["emotions"][i], float("nan")))
        agent.memory.priorities.append(1e-4)
    # emotion history so that get_emotion() inside learn() (:466) is a real S=20 call
    for t in range(25):
        agent.emotion.update(float(g.uniform(-2000, 4000)), (g.rand(2) - 0.5).astype(np.float32))
    out = {}
    out.update(_sd(agent.actor, "a0."))
    out.update(_sd(agent.critic, "c0."))
    out.update(_sd(agent.emotion.transformer, "trans."))
    out["history"] = np.stack([h.numpy() for h in agent.emotion.transformer.history])
    for k, v in buf.items():
        out["buf_" + k] = np.asarray(v)
    agent.train_step = 7
    np.random.seed(2024)
    idx_all, closs, aloss, lrs, next_emotions = [], [], [], [], []
    orig_choice = np.random.choice
    orig_get = agent.emotion.get_emotion

    def tap_choice(n, size, p=None):
        r = orig_choice(n, size, p=p)
        idx_all.append(np.array(r))
        return r

    def tap_get():
        e = orig_get()
        next_emotions.append(np.array(e))
        return e
    np.random.choice = tap_choice
    agent.emotion.get_emotion = tap_get
    try:
        for it in range(3):
            loss = agent.learn()
            closs.append(loss["critic_loss"]); aloss.append(loss["actor_loss"])
            lrs.append((agent.critic_optim.param_groups[0]["lr"], agent.actor_optim.param_groups[0]["lr"]))
            out.update(_sd(agent.actor, f"a{it + 1}."))
            out.update(_sd(agent.critic, f"c{it + 1}."))
            out.update(_sd(agent.actor_target, f"at{it + 1}."))
            out.update(_sd(agent.critic_target, f"ct{it + 1}."))
    finally:
        np.random.choice = orig_choice
    out["idx"] = np.array(idx_all)
    out["critic_loss"] = np.array(closs)
    out["actor_loss"] = np.array(aloss)
    out["lrs"] = np.array(lrs)
    out["next_emotion"] = np.array(next_emotions)
    out["train_step"] = 7
    out["meta"] = _meta("U2-U5 DDPGAgent.learn x3, eval-mode nets, B=128, injected buffer")
    np.savez_compressed(os.path.join(GOLD, "learn_golden.npz"), **out)
    print("learn: critic_loss", closs, "actor_loss", aloss, "lrs", lrs)


def gen_learn_train():
    """U2-U5 in the mode the reference actually runs: DDPGAgent.learn() x3 with actor / critic / both targets in train()
    (Dropout(0.2) active in all five forwards, ddpg_agent.py:40,43,150,155).  torch is re-seeded before every learn() and the ten
    keep-masks it drew are re-drawn from the same seed in at::dropout's order (empty_like(x).bernoulli_(0.8)): actor_target (2),
    critic_target (2), critic (2), actor (2), critic (2).  The EmotionTransformer is put in eval() so that get_emotion() inside
    learn() (:466) draws nothing from torch's generator; its output is recorded and injected on the other side."""
    import torch
    env, agent = _build_agent(0)
    agent.emotion.transformer.eval()
    torch.backends.mha.set_fastpath_enabled(False)
    for m in (agent.actor, agent.actor_target, agent.critic, agent.critic_target):
        m.train()
    g = np.random.RandomState(19)
    B = 128
    agent.batch_size = B
    n_buf = 600
    buf = dict(states=(g.rand(n_buf, 2) * 5).astype(np.float32), actions=(g.rand(n_buf, 10) * 2 - 1).astype(np.float32),
               rewards=g.uniform(-3000, 6000, n_buf), next_states=(g.rand(n_buf, 2) * 5).astype(np.float32),
               dones=(g.rand(n_buf) < 0.05), emotions=g.rand(n_buf, 3).astype(np.float32))
    for i in range(n_buf):
        agent.memory.buffer.append((buf["states"][i], buf["actions"][i], float(buf["rewards"][i]),
                                    buf["next_states"][i], bool(buf["dones"][i]), buf["emotions"][i], float("nan")))
        agent.memory.priorities.append(1e-4)
    for t in range(25):
        agent.emotion.update(float(g.uniform(-2000, 4000)), (g.rand(2) - 0.5).astype(np.float32))
    out = {}
    out.update(_sd(agent.actor, "a0."))
    out.update(_sd(agent.critic, "c0."))
    for k, v in buf.items():
        out["buf_" + k] = np.asarray(v)
    agent.train_step = 3
    np.random.seed(77)
    idx_all, closs, aloss, lrs, next_emotions = [], [], [], [], []
    orig_choice = np.random.choice
    orig_get = agent.emotion.get_emotion

    def tap_choice(n, size, p=None):
        r = orig_choice(n, size, p=p)
        idx_all.append(np.array(r))
        return r

    def tap_get():
        e = orig_get()
        next_emotions.append(np.array(e))
        return e
    np.random.choice = tap_choice
    agent.emotion.get_emotion = tap_get
    shapes = [((B, 128), 0.2)] * 2 + [((B, 256), 0.2)] * 4 + [((B, 128), 0.2)] * 2 + [((B, 256), 0.2)] * 2
    try:
        for it in range(3):
            torch.manual_seed(500 + it)
            loss = agent.learn()
            closs.append(loss["critic_loss"]); aloss.append(loss["actor_loss"])
            lrs.append((agent.critic_optim.param_groups[0]["lr"], agent.actor_optim.param_groups[0]["lr"]))
            masks = _replay_masks(500 + it, shapes)
            for s, m in enumerate(masks):
                out[f"mask{it}_{s}"] = np.packbits(m.numpy().astype(np.uint8), axis=-1)
            out.update(_sd(agent.actor, f"a{it + 1}."))
            out.update(_sd(agent.critic, f"c{it + 1}."))
            out.update(_sd(agent.actor_target, f"at{it + 1}."))
            out.update(_sd(agent.critic_target, f"ct{it + 1}."))
    finally:
        np.random.choice = orig_choice
    out["idx"] = np.array(idx_all)
    out["critic_loss"] = np.array(closs)
    out["actor_loss"] = np.array(aloss)
    out["lrs"] = np.array(lrs)
    out["next_emotion"] = np.array(next_emotions)
    out["train_step"] = 3
    out["meta"] = _meta("U2-U5 DDPGAgent.learn x3, train()-mode nets (dropout masks replayed), B=128, injected buffer")
    np.savez_compressed(os.path.join(GOLD, "learn_train_golden.npz"), **out)
    print("learn_train: critic_loss", closs, "actor_loss", aloss, "lrs", lrs)


# ----------------------------------------------------------------------------------------
def gen_td3sac():
    """T1 vectors: TD3Agent.train_step() x4 and SACAgent.update() x3 of the unmodified reference on an injected
    replay buffer (B = 128).  The sampled indices are tapped from random.sample; the torch noise of each call
    (randn_like, td3_agent.py:268; two rsample() draws, sac_agent.py:69) is re-drawn from the same seed."""
    ref_shim.install()
    import torch
    import MutiConditionEnv as M
    import td3_agent as T3
    import sac_agent as SA
    cfg = BOUNDS["variant_25kg"]
    env = M.MultiConditionWeightEnv({**{k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time")},
                                     "bounds": dict(cfg["bounds"])})
    g = np.random.RandomState(11)
    B, n_buf = 128, 700
    buf = dict(states=(g.rand(n_buf, 2) * 5).astype(np.float32), actions=(g.rand(n_buf, 10) * 2 - 1).astype(np.float32),
               rewards=g.uniform(-3000, 6000, n_buf).astype(np.float32),
               next_states=(g.rand(n_buf, 2) * 5).astype(np.float32), dones=(g.rand(n_buf) < 0.05))
    out = {"buf_" + k: np.asarray(v) for k, v in buf.items()}

    def fill(append):
        for i in range(n_buf):
            append((buf["states"][i], buf["actions"][i].astype(np.float64), float(buf["rewards"][i]),
                    buf["next_states"][i], float(buf["dones"][i])))

    def tap_sample(mod, store):
        orig = mod.random.sample

        def tapped(pop, k):
            st = mod.random.getstate()
            res = orig(pop, k)
            mod.random.setstate(st)
            store.append(np.array(orig(range(len(pop)), k)))
            return res
        return orig, tapped

    # ---- TD3 ----
    torch.manual_seed(3)
    agent = T3.TD3Agent(env, device="cpu")
    fill(lambda t: agent.memory.append(*t))
    out.update(_sd(agent.actor, "td3.a0."))
    out.update(_sd(agent.critic, "td3.c0."))
    random.seed(77)
    idx, noises, res = [], [], []
    orig, tapped = tap_sample(T3, idx)
    T3.random.sample = tapped
    try:
        for it in range(4):
            torch.manual_seed(500 + it)
            noises.append(torch.randn(B, 10).numpy())
            torch.manual_seed(500 + it)
            al, cl, qv = agent.train_step()
            res.append((np.nan if al is None else al, cl, qv))
        # weights after the last step only (fixture size); the per-step losses pin the steps in between
        out.update(_sd(agent.actor, "td3.a4."))
        out.update(_sd(agent.critic, "td3.c4."))
        out.update(_sd(agent.actor_target, "td3.at4."))
        out.update(_sd(agent.critic_target, "td3.ct4."))
    finally:
        T3.random.sample = orig
    out["td3.idx"], out["td3.noise"], out["td3.res"] = np.array(idx), np.array(noises), np.array(res, np.float64)
    # select_action (td3_agent.py:221-236) on a few states, injected numpy normals
    st = (g.rand(6, 2) * 5).astype(np.float32)
    acts, rn = [], []
    for i in range(6):
        np.random.seed(40 + i)
        rn.append(np.random.randn(10))
        np.random.seed(40 + i)
        acts.append(agent.select_action(st[i], noise_scale=0.1))
    out["td3.act_states"], out["td3.act_randn"], out["td3.act_out"] = st, np.array(rn), np.array(acts)

    # ---- SAC ----
    torch.manual_seed(4)
    sac = SA.SACAgent(env, device="cpu")
    fill(sac.memory.buffer.append)
    out.update(_sd(sac.policy, "sac.p0."))
    out.update(_sd(sac.q1, "sac.q1_0."))
    out.update(_sd(sac.q2, "sac.q2_0."))
    random.seed(78)
    idx, e1, e2, res = [], [], [], []
    orig, tapped = tap_sample(SA, idx)
    SA.random.sample = tapped
    try:
        for it in range(3):
            torch.manual_seed(600 + it)
            e1.append(torch.randn(B, 10).numpy()); e2.append(torch.randn(B, 10).numpy())
            torch.manual_seed(600 + it)
            r = sac.update()
            res.append((r["q1_loss"], r["q2_loss"], r["policy_loss"]))
        out.update(_sd(sac.policy, "sac.p3."))
        out.update(_sd(sac.q1, "sac.q1_3."))
        out.update(_sd(sac.q2, "sac.q2_3."))
        out.update(_sd(sac.q1_target, "sac.q1t_3."))
        out.update(_sd(sac.q2_target, "sac.q2t_3."))
    finally:
        SA.random.sample = orig
    out["sac.idx"], out["sac.eps_next"], out["sac.eps_new"] = np.array(idx), np.array(e1), np.array(e2)
    out["sac.res"] = np.array(res, np.float64)
    acts, eps = [], []
    for i in range(6):
        torch.manual_seed(50 + i)
        eps.append(torch.randn(1, 10).numpy()[0])
        torch.manual_seed(50 + i)
        acts.append(sac.act(st[i], deterministic=False))
    out["sac.act_states"], out["sac.act_eps"], out["sac.act_out"] = st, np.array(eps), np.array(acts)
    out["sac.act_det"] = np.array([sac.act(st[i], deterministic=True) for i in range(6)])
    out["meta"] = _meta("T1 TD3Agent.train_step x4 / SACAgent.update x3, B=128, injected buffer, tapped indices and noise")
    np.savez_compressed(os.path.join(GOLD, "td3sac_golden.npz"), **out)
    print("td3:", res if False else out["td3.res"].tolist(), "sac:", out["sac.res"].tolist())

# ----------------------------------------------------------------------------------------
def gen_loop():
    """Config-1 anchor: the reference's OWN train_ddpg (ddpg_agent.py:579-732) driving its own env and agent for
    3 short episodes, eval-mode networks, with every call across the env/agent boundary tapped: what the caller
    passed and what came back.  tests/test_dropin_gpu.py replays the same call sequence through the drop-ins."""
    import tempfile
    import logging
    import torch
    random.seed(0); np.random.seed(0)
    env, agent = _build_agent(0)
    import ddpg_agent as D
    for m in (agent.actor, agent.actor_target, agent.critic, agent.critic_target, agent.emotion.transformer):
        m.eval()
    torch.backends.mha.set_fastpath_enabled(False)
    agent.batch_size = 16
    agent.memory.capacity = 30          # small, so PrioritizedReplayBuffer.add's overwrite branch (:222-226) runs too
    out = {}                            # initial weights = nets_golden.npz's actor./critic./trans. (same _build_agent(0))
    calls = []                          # (kind, payload) in call order
    rec = collections.defaultdict(list)

    def tap(obj, name, kind, pack):
        orig = getattr(obj, name)

        def wrapped(*a, **k):
            r = orig(*a, **k)
            calls.append(kind)
            for key, val in pack(a, k, r).items():
                rec[kind + "." + key].append(np.asarray(val))
            return r
        setattr(obj, name, wrapped)

    idx_log = []
    orig_choice = np.random.choice

    def tap_choice(n, size, p=None):
        r = orig_choice(n, size, p=p)
        idx_log.append(np.array(r))
        return r

    tap(env, "reset", "reset", lambda a, k, r: {"state": r})
    tap(agent, "act", "act", lambda a, k, r: {"state": a[0], "action": r})
    tap(env, "step", "step", lambda a, k, r: {"action": a[0], "next_state": r[0], "reward": np.float64(r[1]), "done": r[2],
                                               "final_weight": np.float64(r[3]["final_weight"]), "total_time": np.float64(r[3]["total_time"])})
    tap(agent.emotion, "update", "update", lambda a, k, r: {"reward": np.float64(a[0]), "movement": a[1], "emotion": r})
    def _slot(a):                       # the slot add() just wrote (the tuple holding this call's state object)
        return [j for j, tr in enumerate(agent.memory.buffer) if tr[0] is a[0]][-1]

    # reward_f32: under numpy 2 some reward branches of MutiConditionEnv.step return np.float32 (weak scalar promotion); the
    # type flows into the NaN td_error and its priority, and float32(1e-4) < float64(1e-4) decides np.argmin when full
    tap(agent, "remember", "remember", lambda a, k, r: {"state": a[0], "action": a[1], "reward": np.float64(a[2]), "next_state": a[3],
                                                        "reward_f32": isinstance(a[2], np.float32),
                                                        "done": a[4], "slot": _slot(a), "stored_emotion": agent.memory.buffer[_slot(a)][5],
                                                        "stored_norm_action": agent.memory.buffer[_slot(a)][1],
                                                        "len": len(agent.memory)})
    tap(agent, "learn", "learn", lambda a, k, r: {"ran": r is not None,
                                                  "critic_loss": np.float64(r["critic_loss"]) if r else np.nan,
                                                  "actor_loss": np.float64(r["actor_loss"]) if r else np.nan,
                                                  "critic_lr": agent.critic_optim.param_groups[0]["lr"],
                                                  "actor_lr": agent.actor_optim.param_groups[0]["lr"]})
    np.random.choice = tap_choice
    cwd = os.getcwd()
    try:
        with tempfile.TemporaryDirectory() as td, _quiet():
            os.chdir(td)
            lg = logging.getLogger("gen_loop"); lg.addHandler(logging.NullHandler()); lg.propagate = False
            import tqdm as _tq
            D.tqdm = lambda it, **k: it
            rewards = D.train_ddpg(env, agent, episodes=3, max_steps=12, log_prefix="gen_loop", logger=lg)
    finally:
        os.chdir(cwd)
        np.random.choice = orig_choice
    out["calls"] = np.array(calls)
    for k, v in rec.items():
        out[k] = np.stack(v)
    out["learn.idx"] = np.stack(idx_log)
    out["episode_rewards"] = np.array(rewards, np.float64)
    out.update(_sd(agent.actor, "actor_end."))
    out.update(_sd(agent.critic, "critic_end."))
    out.update(_sd(agent.actor_target, "actor_target_end."))
    out.update(_sd(agent.critic_target, "critic_target_end."))
    out["train_step_end"] = agent.train_step
    # the checkpoint the reference writes for this agent (N2): keys and tensors, flattened
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ref.pth")
        agent.save(path)
        ck = torch.load(path, weights_only=False)
    out["ckpt_keys"] = np.array(sorted(ck.keys()))
    for name in ("actor", "critic", "actor_target", "critic_target", "emotion_transformer"):
        out["ckpt_order." + name] = np.array(list(ck[name].keys()))
    for name in ("actor_optim", "critic_optim"):
        st = ck[name]["state"]
        out[f"ckpt.{name}.n"] = len(st)
        out[f"ckpt.{name}.step"] = np.array([float(st[i]["step"]) for i in sorted(st)])
        for i in sorted(st):
            out[f"ckpt.{name}.exp_avg.{i}"] = st[i]["exp_avg"].numpy()
            out[f"ckpt.{name}.exp_avg_sq.{i}"] = st[i]["exp_avg_sq"].numpy()
        out[f"ckpt.{name}.lr"] = ck[name]["param_groups"][0]["lr"]
    out["meta"] = _meta("config-1 anchor: reference train_ddpg x3 episodes x12 steps, eval-mode nets, B=16, capacity 30, "
                        "seeds random/np/torch = 0; plus the reference's save() checkpoint layout")
    np.savez_compressed(os.path.join(GOLD, "loop_golden.npz"), **out)
    print("loop: calls", len(calls), "learns", int(np.sum(out["learn.ran"])), "rewards", rewards)

# ----------------------------------------------------------------------------------------
def gen_esac():
    """N3 vectors: EmotionSACAgent.update() x3 and act() of the reference (ERL_Fill_agent.py) on an injected replay buffer
    (B = 128), eval-mode EmotionTransformer; sampled indices tapped from random.sample, the two rsample() draws per update
    re-drawn from the same torch seed."""
    ref_shim.install()
    import torch
    import MutiConditionEnv as M
    import ERL_Fill_agent as EF
    cfg = BOUNDS["variant_25kg"]
    env = M.MultiConditionWeightEnv({**{k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time")},
                                     "bounds": dict(cfg["bounds"])})
    g = np.random.RandomState(21)
    B, n_buf = 128, 600
    buf = dict(states=(g.rand(n_buf, 2) * 5).astype(np.float32), actions=(g.rand(n_buf, 10) * 2 - 1).astype(np.float32),
               rewards=g.uniform(-3000, 6000, n_buf).astype(np.float32), next_states=(g.rand(n_buf, 2) * 5).astype(np.float32),
               dones=(g.rand(n_buf) < 0.05), emotions=g.rand(n_buf, 3).astype(np.float32))
    out = {"buf_" + k: np.asarray(v) for k, v in buf.items()}
    torch.manual_seed(5)
    with _quiet():
        agent = EF.EmotionSACAgent(env, device="cpu")
    agent.emotion.transformer.eval()
    torch.backends.mha.set_fastpath_enabled(False)
    for i in range(n_buf):
        agent.memory.add((buf["states"][i], buf["actions"][i].astype(np.float64), float(buf["rewards"][i]), buf["next_states"][i],
                          float(buf["dones"][i]), buf["emotions"][i]))
    out.update(_sd(agent.policy, "p0."))
    out.update(_sd(agent.q1, "q1_0."))
    out.update(_sd(agent.q2, "q2_0."))
    for t in range(6):      # an emotion history, so get_emotion() is a real S = 20 call and e_now != e_prev between updates
        agent.emotion.update(float(g.uniform(-2000, 4000)), (g.rand(10) * 100).astype(np.float32))
    random.seed(79)
    idx, e1, e2, res, enow = [], [], [], [], []
    orig = EF.random.sample

    def tapped(pop, k):
        st = EF.random.getstate()
        r = orig(pop, k)
        EF.random.setstate(st)
        idx.append(np.array(orig(range(len(pop)), k)))
        return r
    EF.random.sample = tapped
    upd_rewards, upd_moves = [], []
    try:
        with _quiet():
            for it in range(3):
                torch.manual_seed(700 + it)
                e1.append(torch.randn(B, 10).numpy()); e2.append(torch.randn(B, 10).numpy())
                torch.manual_seed(700 + it)
                enow.append(np.array(agent.emotion.get_emotion()))
                r = agent.update()
                res.append((r["q1_loss"], r["q2_loss"], r["policy_loss"], r["emo_reg"]))
                rw, mv = float(g.uniform(-2000, 4000)), (g.rand(10) * 100).astype(np.float32)
                upd_rewards.append(rw); upd_moves.append(mv)
                agent.emotion.update(rw, mv)                       # the emotion moves between updates (train_erl_fill :538)
    finally:
        EF.random.sample = orig
    # end-of-run weights: every 9th element of every tensor (fixture size); the per-update losses pin the steps in between
    for name, mod in (("p3.", agent.policy), ("q1_3.", agent.q1), ("q2_3.", agent.q2), ("q1t_3.", agent.q1_target), ("q2t_3.", agent.q2_target)):
        out.update({k: v.reshape(-1)[::9].copy() for k, v in _sd(mod, name).items()})
    out["idx"], out["eps_next"], out["eps_new"] = np.array(idx), np.array(e1), np.array(e2)
    out["res"], out["e_now"] = np.array(res, np.float64), np.array(enow)
    out["upd_rewards"], out["upd_moves"] = np.array(upd_rewards), np.array(upd_moves)
    # act (:322-348): stochastic (injected eps), deterministic
    st = (g.rand(6, 2) * 5).astype(np.float32)
    acts, eps = [], []
    out["act_emotion"] = np.array(agent.emotion.get_emotion())
    for i in range(6):
        torch.manual_seed(60 + i)
        eps.append(torch.randn(1, 10).numpy()[0])
        torch.manual_seed(60 + i)
        acts.append(agent.act(st[i]))
    out["act_states"], out["act_eps"], out["act_out"] = st, np.array(eps), np.array(acts)
    out["act_det"] = np.array([agent.act(st[i], deterministic=True) for i in range(6)])
    out["meta"] = _meta("N3 EmotionSACAgent.update x3 + act, B=128, injected buffer, tapped indices and noise, eval-mode transformer")
    np.savez_compressed(os.path.join(GOLD, "esac_golden.npz"), **out)
    print("esac:", out["res"].tolist())


def gen_exp3():
    """Row N4: the reference's OWN evaluation loop and result files.  `experiment_runner.run_experiment_3(train=False, algo='er_ddpg')`
    (experiment_runner.py:397-611) is run unmodified on a checkpoint written by the reference's own `train_ddpg` + `agent.save`: ten
    test episodes per condition with `act(state, add_noise=False)`, `analysis_outputs/exp3/test_results_er_ddpg_<cond>.csv` and
    `logs/exp3_summary.log`.  Two taps only: `DDPGAgent.load` also switches the loaded modules to eval() (the reference leaves dropout
    on in its test loop, which no second implementation can reproduce draw for draw), and `DDPGAgent.act` records (state, action)."""
    import shutil
    import tempfile
    import torch
    ref_shim.install()
    import experiment_runner as ER
    work = tempfile.mkdtemp(prefix="erlfill_exp3_")
    os.chdir(work)
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    # a short reference training run, saved with the reference's own save()
    cfg = BOUNDS["variant_25kg"]
    with _quiet():
        env = ER.MultiConditionWeightEnv({**{k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time")},
                                          "bounds": dict(cfg["bounds"])})
        agent = ER.DDPGAgent(env, "cpu", use_td_error=True)
        agent.emotion = ER.EmotionModule("cpu")
        env.attach_agent(agent)
        agent.batch_size = 16
        ER.train_ddpg(env, agent, episodes=2, max_steps=12, log_prefix="exp3_gold")
        os.makedirs("saved_models/exp3_er_ddpg", exist_ok=True)
        for key in ("variant_25kg", "variant_20kg", "variant_15kg"):
            agent.save(f"saved_models/exp3_er_ddpg/{key}_final.pth")
    ck = torch.load("saved_models/exp3_er_ddpg/variant_25kg_final.pth", map_location="cpu", weights_only=False)
    out = {}
    for part in ("actor", "critic", "emotion_transformer"):
        for k, v in ck[part].items():
            out[f"ck.{part}.{k}"] = v.numpy()
    taps = []
    orig_load, orig_act = ER.DDPGAgent.load, ER.DDPGAgent.act

    def load_eval(self, filename):
        orig_load(self, filename)
        for m in (self.actor, self.critic, self.emotion.transformer):
            m.eval()

    def act_tap(self, state, add_noise=True):
        a = orig_act(self, state, add_noise)
        taps.append((np.array(state, np.float32), np.array(a, np.float32)))
        return a
    ER.DDPGAgent.load, ER.DDPGAgent.act = load_eval, act_tap
    try:
        random.seed(2024); np.random.seed(2024); torch.manual_seed(2024)
        with _quiet():
            ER.run_experiment_3(train=False, max_steps=8, device="cpu", algo="er_ddpg")
    finally:
        ER.DDPGAgent.load, ER.DDPGAgent.act = orig_load, orig_act
    out["states"] = np.stack([t[0] for t in taps]); out["actions"] = np.stack([t[1] for t in taps])
    for key in ("variant_25kg", "variant_20kg", "variant_15kg"):
        out["csv_" + key] = np.frombuffer(open(f"analysis_outputs/exp3/test_results_er_ddpg_{key}.csv", "rb").read(), np.uint8)
    out["summary"] = np.frombuffer(open("logs/exp3_summary.log", "rb").read(), np.uint8)
    out["seed"], out["max_steps"] = 2024, 8
    out["meta"] = _meta("N4 run_experiment_3(train=False, er_ddpg): test loop, per-condition CSVs, summary log")
    np.savez_compressed(os.path.join(GOLD, "exp3_golden.npz"), **out)
    print("exp3:", len(taps), "test steps;", open("logs/exp3_summary.log").read())
    os.chdir("/tmp")
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["sim", "env", "emotion"]
    os.chdir("/tmp")
    for w in which:
        {"sim": gen_sim, "env": gen_env, "emotion": gen_emotion, "nets": gen_nets, "learn": gen_learn, "td3sac": gen_td3sac, "loop": gen_loop, "esac": gen_esac,
         "onboard": gen_onboard, "learn_train": gen_learn_train,
         "exp3": gen_exp3}[w]()
