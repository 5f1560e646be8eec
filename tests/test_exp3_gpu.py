"""GPU: SURVEY.md section 8f row N4 against files the REFERENCE's own evaluation loop produced.

tests/golden/exp3_golden.npz (oracle/gen_golden.py exp3) holds what `experiment_runner.run_experiment_3(train=False, algo='er_ddpg')`
wrote for a checkpoint saved by the reference's own train_ddpg: the three per-condition CSVs, exp3_summary.log, and every
(state, action) of its 240 test steps (modules switched to eval() at load).  Here the drop-in env + agent load the same checkpoint and
`loops.test_agent` / `write_results_csv` / `write_summary` walk the same ten episodes per condition:
  * the environment is bit-exact, so with the reference's actions replayed every state the loop sees equals the reference's, and
    Reward / WeightErr / Time / SlowWeight of every CSV row are equal (rewards to 1 ulp of binary32 per step);
  * the drop-in agent's own `act(state, add_noise=False)` on those states matches the reference's action to 2e-5 of the action range,
    its emotion columns to 2e-5;
  * the files: same header bytes (BOM included), same columns and row order, the summary log text identical."""
import csv
import io
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = {"variant_25kg": "Condition Variant - 25kg±25g", "variant_20kg": "Condition Variant - 20kg±20g",
         "variant_15kg": "Condition Variant - 15kg±15g"}


class _Replayed:
    """The drop-in agent, but `act` hands back the reference's recorded action after checking that the loop is at the recorded state."""

    def __init__(self, agent, states, actions, cursor, errs, scale):
        self.agent, self.states, self.actions, self.cursor, self.errs, self.scale = agent, states, actions, cursor, errs, scale
        self.emotion = agent.emotion

    def act(self, state, add_noise=False):
        i = self.cursor[0]
        assert not add_noise
        assert np.array_equal(np.asarray(state, np.float32), self.states[i]), ("state of test step %d differs from the reference's" % i)
        own = self.agent.act(state, add_noise=False)
        self.errs.append(float(np.max(np.abs(own - self.actions[i]) / self.scale)))
        self.cursor[0] += 1
        return self.actions[i].copy()


def test_reference_evaluation_loop_and_result_files(tmp_path):
    import erlfill_b200 as E
    from erlfill_b200 import loops
    K = E.conditions
    g = np.load(os.path.join(GOLD, "exp3_golden.npz"))
    ck = {part: {k[len("ck." + part) + 1:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("ck." + part + ".")}
          for part in ("actor", "critic", "emotion_transformer")}
    path = str(tmp_path / "variant_final.pth")
    torch.save(ck, path)
    random.seed(int(g["seed"]))
    cursor, results = [0], []
    for key, name in NAMES.items():
        cfg = K.EXPERIMENT_3[key]
        env = E.MultiConditionWeightEnv({k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time", "bounds")})
        agent = E.DDPGAgent(env, device="cuda", use_td_error=True)
        agent.emotion = E.EmotionModule(device="cuda")
        env.attach_agent(agent)
        agent.load(path)
        errs = []
        scale = np.array([(hi - lo) for lo, hi in cfg["bounds"].values()], np.float32)
        rep = _Replayed(agent, g["states"], g["actions"], cursor, errs, scale)
        (avg_r, avg_e, avg_t), detail = loops.test_agent(env, rep, "er_ddpg", episodes=10, max_steps=int(g["max_steps"]), condition=name)
        results.append((name, avg_r, avg_e, avg_t))
        assert max(errs) <= 2e-5, (key, max(errs))                                   # the drop-in actor on the reference's states
        ref_text = bytes(g["csv_" + key]).decode("utf-8-sig")
        ref_rows = list(csv.DictReader(io.StringIO(ref_text)))
        assert len(ref_rows) == len(detail) == 10
        for ours, ref in zip(detail, ref_rows):
            assert ours["Condition"] == ref["Condition"] and ours["Episode"] == int(ref["Episode"])
            for col in ("WeightErr", "Time"):
                assert float(ours[col]) == float(ref[col]), (key, ours["Episode"], col)
            assert isinstance(ours["SlowWeight"], np.float32) and ours["SlowWeight"] == np.float32(ref["SlowWeight"])   # a binary32 column
            assert abs(float(ours["Reward"]) - float(ref["Reward"])) <= 2e-6 * abs(float(ref["Reward"])) + 1e-3
            for col in ("Curiosity", "Conservativeness", "Anxiety"):
                assert abs(float(ours[col]) - float(ref[col])) <= 2e-5, (key, col)
        out = str(tmp_path / ("test_results_er_ddpg_%s.csv" % key))
        loops.write_results_csv(out, detail)
        mine = open(out, "rb").read()
        ref_bytes = bytes(g["csv_" + key])
        assert mine.split(b"\n")[0] == ref_bytes.split(b"\n")[0].rstrip(b"\r")        # BOM + header
        assert mine.count(b"\n") == ref_bytes.count(b"\n")
        mine_rows = list(csv.DictReader(io.StringIO(mine.decode("utf-8-sig"))))
        for a, b in zip(mine_rows, ref_rows):
            assert a["Condition"] == b["Condition"] and a["Episode"] == b["Episode"]
            assert a["WeightErr"] == b["WeightErr"] and a["Time"] == b["Time"] and a["SlowWeight"] == b["SlowWeight"]   # same text
    assert cursor[0] == len(g["states"])                                             # every reference step was walked
    out = str(tmp_path / "exp3_summary.log")
    loops.write_summary(out, results)
    assert open(out, "rb").read() == bytes(g["summary"])


def test_evaluate_batched_equals_the_scalar_loop_on_the_oracle():
    """`loops.evaluate_batched` (one deterministic episode per env on the device) against the C oracle stepped with the same actions:
    per-condition mean reward / weight error / time from the oracle's own outputs."""
    import erlfill_b200 as E
    from erlfill_b200 import loops
    from oracle import oracle_lib as O
    K = E.conditions
    conds = list(K.EXPERIMENT_3.values())
    n, max_steps, seed = 96, 6, 11
    env = E.VecFillEnv(conds, n, kind="multi", max_steps=max_steps, seed=seed)
    agent = E.VecERDDPG(env, seed=seed, batch_size=32, memory_size=1024)
    out, per_env = loops.evaluate_batched(agent, max_steps=max_steps, condition_names=list(K.EXPERIMENT_3))
    # the same episode on the oracle: the agent's actions are a function of (obs, emotion), both reproduced step by step
    env2 = E.VecFillEnv(conds, n, kind="multi", max_steps=max_steps, seed=seed, auto_reset=False)
    agent2 = E.VecERDDPG(env2, seed=seed, batch_size=32, memory_size=1024)
    oc = [O.make_condition(c["target_weight"], c["target_weight_err"], c["target_time"], c["bounds"]) for c in conds]
    ref = O.BatchEnv(O.ENV_MULTI, oc, n, max_steps=max_steps, seed=seed)
    alive = np.ones(n, bool)
    ep_r, werr, ftime = np.zeros(n), np.zeros(n), np.zeros(n)
    tw = np.array([conds[i % 3]["target_weight"] for i in range(n)], np.float64)
    for step in range(max_steps):
        a = agent2.act(add_noise=False)
        env2.step(a)
        ref.step(a.cpu().numpy(), step, auto_reset=False)
        ep_r += np.where(alive, ref.reward.astype(np.float64), 0.0)
        werr = np.where(alive, np.abs(ref.final_weight - tw), werr)
        ftime = np.where(alive, ref.total_time, ftime)
        alive &= ~ref.done.astype(bool)
    cid = np.arange(n) % 3
    for c, name in enumerate(K.EXPERIMENT_3):
        m = cid == c
        got = out[name]
        np.testing.assert_allclose(got[0], ep_r[m].mean(), rtol=1e-6)
        np.testing.assert_allclose(got[1], werr[m].mean(), rtol=1e-12)
        np.testing.assert_allclose(got[2], ftime[m].mean(), rtol=1e-12)
        assert got[3] == int(m.sum())
