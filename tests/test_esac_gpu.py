"""GPU: EmotionSAC (csrc/esac.cu, erlfill_b200.esac) against the torch restatement (oracle/nets_ref.py esac_*) and the fixtures
recorded from the reference agent (tests/golden/esac_golden.npz).  Bars: losses 1e-4 relative, updated parameters 1e-4 relative
(2e-6 absolute), actions 1e-5 of the action range."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nets_ref as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "esac_golden.npz")


def _sd(G, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(G[k])) for k in G.files if k.startswith(prefix)}


def _agent(E, G, batch=128):
    cfg = E.conditions.EXPERIMENT_3["variant_25kg"]
    env = E.VecFillEnv([cfg], 8, kind="multi", max_steps=10, seed=1)
    q = {"q1." + k: v for k, v in _sd(G, "q1_0.").items()}
    q.update({"q2." + k: v for k, v in _sd(G, "q2_0.").items()})
    agent = E.VecEmotionSAC(env, batch_size=batch, memory_size=1024, state_dicts=(None, _sd(G, "p0."), q))
    n = G["buf_states"].shape[0]
    dev = env.device
    t = lambda k: torch.from_numpy(np.asarray(G["buf_" + k]).astype(np.float32)).to(dev)
    agent.memory.insert(t("states"), t("actions"), t("rewards"), t("next_states"), torch.from_numpy(G["buf_dones"].astype(np.uint8)).to(dev),
                        t("emotions"))
    assert len(agent.memory) == n
    return env, agent


def test_esac_update_matches_reference_and_oracle():
    import erlfill_b200 as E
    G = np.load(GOLD)
    env, agent = _agent(E, G)
    dev = env.device
    # oracle side, same inputs
    pol, q1, q2 = _sd(G, "p0."), _sd(G, "q1_0."), _sd(G, "q2_0.")
    q1t, q2t = {k: v.clone() for k, v in q1.items()}, {k: v.clone() for k, v in q2.items()}
    z = lambda w, order: [torch.zeros_like(w[k]) for k in order]
    opt = {"pm": z(pol, R.ESAC_POLICY_ORDER), "pv": z(pol, R.ESAC_POLICY_ORDER), "q1m": z(q1, R.ESAC_Q_ORDER),
           "q1v": z(q1, R.ESAC_Q_ORDER), "q2m": z(q2, R.ESAC_Q_ORDER), "q2v": z(q2, R.ESAC_Q_ORDER), "step": 0}
    e_prev = G["e_now"][0]
    for it in range(3):
        idx = G["idx"][it]
        e_now = torch.from_numpy(G["e_now"][it]).to(dev)
        rep = agent.update(idx=torch.from_numpy(idx.astype(np.int64)).to(dev), eps_next=torch.from_numpy(G["eps_next"][it]).to(dev),
                           eps_new=torch.from_numpy(G["eps_new"][it]).to(dev), e_now=e_now).cpu().numpy()
        b = {k: torch.from_numpy(np.asarray(G["buf_" + k])[idx].astype(np.float32)) for k in ("states", "actions", "rewards", "next_states", "dones", "emotions")}
        ref = R.esac_update(pol, q1, q2, q1t, q2t, opt, b, torch.from_numpy(G["eps_next"][it]), torch.from_numpy(G["eps_new"][it]),
                            G["e_now"][it], e_prev)
        e_prev = G["e_now"][it]
        np.testing.assert_allclose(rep[:3], ref[:3], rtol=1e-4)                      # vs the oracle
        np.testing.assert_allclose(rep[:3], G["res"][it][:3], rtol=1e-4)             # vs the reference's own numbers
        assert abs(rep[3] - G["res"][it][3]) <= 1e-6
    ours_p, ours_q, ours_qt = agent.policy.state_dict(), agent.q.state_dict(), agent.q_target.state_dict()
    for k in R.ESAC_POLICY_ORDER:
        np.testing.assert_allclose(ours_p[k].cpu().numpy(), pol[k].numpy(), rtol=1e-4, atol=2e-6, err_msg=k)
        np.testing.assert_allclose(ours_p[k].cpu().numpy().reshape(-1)[::9], G["p3." + k], rtol=2e-4, atol=2e-6, err_msg=k)
    for name, w, wt in (("q1", q1, q1t), ("q2", q2, q2t)):
        for k in R.ESAC_Q_ORDER:
            np.testing.assert_allclose(ours_q[f"{name}.{k}"].cpu().numpy(), w[k].numpy(), rtol=1e-4, atol=2e-6, err_msg=name + k)
            np.testing.assert_allclose(ours_qt[f"{name}.{k}"].cpu().numpy(), wt[k].numpy(), rtol=1e-4, atol=2e-6, err_msg=name + "t" + k)
            np.testing.assert_allclose(ours_q[f"{name}.{k}"].cpu().numpy().reshape(-1)[::9], G[f"{name}_3.{k}"], rtol=2e-4, atol=2e-6)


def test_esac_act_matches_reference():
    """act() of the reference was recorded with its end-of-run policy: replay the three updates, then compare."""
    import erlfill_b200 as E
    G = np.load(GOLD)
    env, agent = _agent(E, G)
    dev = env.device
    for it in range(3):
        agent.update(idx=torch.from_numpy(G["idx"][it].astype(np.int64)).to(dev), eps_next=torch.from_numpy(G["eps_next"][it]).to(dev),
                     eps_new=torch.from_numpy(G["eps_new"][it]).to(dev), e_now=torch.from_numpy(G["e_now"][it]).to(dev))
    st = torch.from_numpy(G["act_states"]).to(dev)
    em = torch.from_numpy(G["act_emotion"]).float().to(dev)
    scale = ((agent.hi - agent.lo) / 2).cpu().numpy()
    a = agent.act(st, emotion=em, eps=torch.from_numpy(G["act_eps"]).to(dev)).cpu().numpy()
    assert np.all(np.abs(a - G["act_out"]) <= 2e-5 * 2 * scale + 1e-6 * np.abs(G["act_out"])), np.abs(a - G["act_out"]).max(0)
    ad = agent.act(st, emotion=em, deterministic=True).cpu().numpy()
    assert np.all(np.abs(ad - G["act_det"]) <= 2e-5 * 2 * scale + 1e-6 * np.abs(G["act_det"]))
    # per-row emotions equal the broadcast call
    a2 = agent.act(st, emotion=em.unsqueeze(0).repeat(6, 1).contiguous(), deterministic=True).cpu().numpy()
    assert np.array_equal(ad, a2)


def test_esac_batched_loop_runs_and_learns():
    """The batched train_erl_fill body: rollout steps fill the replay with (s, a_norm, r, s', d, e), updates run on it with Philox
    noise; losses finite, parameters move, emotions in [0, 1]."""
    import erlfill_b200 as E
    cfgs = list(E.conditions.EXPERIMENT_3.values())[:1]
    env = E.VecFillEnv(cfgs, 64, kind="multi", max_steps=6, seed=2)
    agent = E.VecEmotionSAC(env, seed=2, batch_size=128, memory_size=4096)
    agent.reset()
    p0 = agent.policy.flat.clone()
    reps = []
    for t in range(6):
        agent.rollout_step()
        r = agent.update()
        if r is not None:
            reps.append(r.cpu().numpy().copy())
    assert len(agent.memory) == 6 * 64 and len(reps) >= 4
    assert np.all(np.isfinite(np.array(reps)))
    assert float((agent.policy.flat - p0).abs().max()) > 0
    e = agent.memory.buf[:len(agent.memory), 16:19]
    assert float(e.min()) >= 0.0 and float(e.max()) <= 1.0
    na = agent.memory.buf[:len(agent.memory), 2:12]
    assert float(na.abs().max()) <= 1.0


def test_esac_scalar_dropin_matches_reference_update():
    """EmotionSACAgent (N = 1 drop-in): the reference's recorded update inputs through the scalar API (random.sample indices from
    the same seed, injected rsample() normals) reproduce the reference's losses; act / remember / save / load surface."""
    import random
    import erlfill_b200 as E
    G = np.load(GOLD)
    cfg = E.conditions.EXPERIMENT_3["variant_25kg"]
    env = E.MultiConditionWeightEnv({k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time", "bounds")})
    q = {"q1." + k: v for k, v in _sd(G, "q1_0.").items()}
    q.update({"q2." + k: v for k, v in _sd(G, "q2_0.").items()})
    torch.manual_seed(0)
    agent = E.EmotionSACAgent(env, state_dicts=(E.init._transformer_sd(), _sd(G, "p0."), q))
    n = G["buf_states"].shape[0]
    for i in range(n):       # remember() takes physical actions and normalises them (:350-363)
        a_phys = agent.denormalize_action(G["buf_actions"][i].astype(np.float64))
        # the stored emotion comes from the agent's EmotionModule; parity needs the fixture's, so write the tuple as the reference had it
        agent.memory.add((G["buf_states"][i], G["buf_actions"][i].astype(np.float64), float(G["buf_rewards"][i]), G["buf_next_states"][i],
                          float(G["buf_dones"][i]), G["buf_emotions"][i]))
        assert np.allclose(agent.normalize_action(a_phys), G["buf_actions"][i], atol=1e-12)
    dev = agent.device
    t = lambda k: torch.from_numpy(np.asarray(G["buf_" + k]).astype(np.float32)).to(dev)
    agent._vec.memory.insert(t("states"), t("actions"), t("rewards"), t("next_states"), torch.from_numpy(G["buf_dones"].astype(np.uint8)).to(dev),
                             t("emotions"))
    random.seed(79)
    out = agent.update(eps_next=G["eps_next"][0], eps_new=G["eps_new"][0])
    assert np.array_equal(agent.last_indices, G["idx"][0])                                 # random.sample indices: bit-exact
    for k, ref in zip(("q1_loss", "q2_loss", "policy_loss"), G["res"][0][:3]):
        assert abs(out[k] - ref) <= 1e-4 * abs(ref), (k, out[k], ref)
    assert out["emo_reg"] == 0.0                                                           # first update: prev_emotion = e_now
    s = np.array([0.5, 1.0], np.float32)
    a = agent.act(s)
    lo = np.array([b[0] for b in env.bounds.values()]); hi = np.array([b[1] for b in env.bounds.values()])
    assert a.shape == (10,) and a.dtype == np.float64 and np.all(a >= lo) and np.all(a <= hi)
    assert np.array_equal(agent.act(s, deterministic=True), agent.act(s, add_noise=False))
    e = agent.emotion.update(123.0, a)                                                     # 10-vector movement (train_erl_fill :538)
    assert e.shape == (3,)
    agent.remember(s, a, 1.0, s, 0.0)
    assert len(agent.memory) == n + 1 and np.allclose(agent.memory.buffer[-1][5], agent.emotion.get_emotion())
    assert list(agent.policy.state_dict().keys()) == R.ESAC_POLICY_ORDER and list(agent.q1.state_dict().keys()) == R.ESAC_Q_ORDER
