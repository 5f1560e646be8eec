"""GPU parity of SURVEY.md section 8f row N1: the env classes' "on-board" per-tick motor / weight models (speed cap,
int()-truncated weight; WeightEnv.py:176-229, MutiConditionEnv.py:184-226) as a switch of the tick kernel
(`VecFillEnv(onboard=True)` -> ERLFILL_ENV_ONBOARD), against (1) fixtures made by the reference's own simulate_run driving the
reference env's own methods (oracle/gen_golden.py onboard) and (2) the C oracle on the same Philox noise."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle_lib as O  # noqa: E402  (checker only)

WIDE = {k: (-1e9, 1e9) for k in O.ACTION_ORDER}


def test_onboard_golden_bit_exact(golden_dir):
    import erlfill_b200 as E
    g = np.load(os.path.join(golden_dir, "onboard_golden.npz"))
    n = len(g["seeds"])
    rows = int(g["iters"].max()) + 2
    u = np.zeros((rows, n), np.float64)
    for i in range(n):
        rng = random.Random(int(g["seeds"][i]))
        u[0, i] = 15 + 35 * rng.random()
        k = int(g["iters"][i]) + 1
        u[1:1 + k, i] = [-1.0 + 2.0 * rng.random() for _ in range(k)]
    conds = [{"target_weight": int(g["params"][i, 0]), "target_weight_err": 25, "target_time": 2.5, "bounds": WIDE} for i in range(n)]
    env = E.VecFillEnv(conds, n, kind="multi", cond_id=np.arange(n), auto_reset=False, onboard=True)
    p = g["params"]
    act = np.stack([p[:, 1], p[:, 2], p[:, 3], p[:, 7], p[:, 8], p[:, 9], p[:, 4], p[:, 5], p[:, 6], p[:, 10]], 1)
    env.step(torch.tensor(act, dtype=torch.float32, device="cuda"), noise_u=torch.tensor(u, device="cuda"))
    torch.cuda.synchronize()
    assert np.array_equal(env.ticks.cpu().numpy(), g["iters"])
    assert np.array_equal(env.total_time.cpu().numpy(), g["total_time"])
    assert np.array_equal(env.final_weight.cpu().numpy(), g["final_weight"])
    assert g["peak_speed"].max() == 10000.0 and (g["iters"] == 5001).any()       # the cap and the iteration stop are in the fixture


@pytest.mark.parametrize("kind,n_env,max_steps", [("multi", 3001, 6), ("single", 777, 4)])
def test_onboard_philox_batch_matches_oracle(kind, n_env, max_steps):
    """Philox noise, ragged N, three conditions round-robin, actions 10 % outside the bounds, auto-reset: every output of the
    env step (ticks, final weight, time, state, reward, done, reset observation) equals the C oracle's on-board variant, and
    differs from the offline controller's on the same noise."""
    import erlfill_b200 as E
    K = E.conditions
    cond_dicts = list(K.EXPERIMENT_3.values()) if kind == "multi" else [K.SINGLE_25KG]
    seed, base = 0xBEEF12345, 4242
    env = E.VecFillEnv(cond_dicts, n_env, kind=kind, max_steps=max_steps, seed=seed, env_id_base=base, onboard=True)
    off = E.VecFillEnv(cond_dicts, n_env, kind=kind, max_steps=max_steps, seed=seed, env_id_base=base)
    oc = [O.make_condition(c["target_weight"], c["target_weight_err"], c["target_time"], c["bounds"]) for c in cond_dicts]
    ref = O.BatchEnv((O.ENV_MULTI if kind == "multi" else O.ENV_SINGLE) | O.ENV_ONBOARD, oc, n_env, max_steps=max_steps, seed=seed,
                     env_id_base=base)
    rng = np.random.RandomState(5)
    lo = np.array([[c["bounds"][k][0] for k in O.ACTION_ORDER] for c in cond_dicts])
    hi = np.array([[c["bounds"][k][1] for k in O.ACTION_ORDER] for c in cond_dicts])
    cid = (base + np.arange(n_env)) % len(cond_dicts)
    total = 0
    for step in range(2 * max_steps + 2):
        span = hi[cid] - lo[cid]
        a = (lo[cid] - 0.1 * span + rng.uniform(0, 1, (n_env, 10)) * 1.2 * span).astype(np.float32)
        at = torch.tensor(a, device="cuda")
        ns, r, d, info = env.step(at)
        total += ref.step(a, step)
        torch.cuda.synchronize()
        assert np.array_equal(info["ticks"].cpu().numpy(), ref.ticks), step
        assert np.array_equal(info["final_weight"].cpu().numpy(), ref.final_weight), step
        assert np.array_equal(info["total_time"].cpu().numpy(), ref.total_time), step
        assert np.array_equal(ns.cpu().numpy(), ref.next_state), step
        assert np.array_equal(d.cpu().numpy(), ref.done), step
        assert np.array_equal(env.obs.cpu().numpy(), ref.obs), step
        rr = r.cpu().numpy()
        assert np.all(np.abs(rr - ref.reward) <= np.spacing(np.abs(ref.reward))), step
        if step == 0:
            off.step(at)
            torch.cuda.synchronize()
            assert (off.final_weight.cpu().numpy() != ref.final_weight).mean() > 0.9       # a different trajectory, not a relabelling
            assert (ref.final_weight == np.trunc(ref.final_weight)).mean() > 0.9          # integer weights (the first-tick delay aside)
    assert int(env.tick_sum.item()) == total
