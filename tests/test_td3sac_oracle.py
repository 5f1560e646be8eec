"""The torch restatement of TD3 / SAC (oracle/nets_ref.py, T1) against fixtures produced by the unmodified reference
agents (oracle/gen_golden.py td3sac): per-step losses and the weights after the last step."""
import os

import numpy as np
import torch

from oracle import nets_ref as R


def _w(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith(prefix)}


def _batch(g, idx):
    return {"states": torch.from_numpy(g["buf_states"][idx]), "actions": torch.from_numpy(g["buf_actions"][idx]),
            "rewards": torch.from_numpy(g["buf_rewards"][idx]).float(), "next_states": torch.from_numpy(g["buf_next_states"][idx]),
            "dones": torch.from_numpy(g["buf_dones"][idx].astype(np.float32))}


def test_td3_restatement_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "td3sac_golden.npz"))
    actor, critic = _w(g, "td3.a0."), _w(g, "td3.c0.")
    actor_t, critic_t = {k: v.clone() for k, v in actor.items()}, {k: v.clone() for k, v in critic.items()}
    opt = {"am": [torch.zeros_like(actor[k]) for k in R.TD3_ACTOR_ORDER], "av": [torch.zeros_like(actor[k]) for k in R.TD3_ACTOR_ORDER],
           "cm": [torch.zeros_like(critic[k]) for k in R.TD3_CRITIC_ORDER], "cv": [torch.zeros_like(critic[k]) for k in R.TD3_CRITIC_ORDER],
           "cstep": 0, "astep": 0}
    for it in range(4):
        al, cl, qv = R.td3_train_step(actor, actor_t, critic, critic_t, opt, _batch(g, g["td3.idx"][it]),
                                      torch.from_numpy(g["td3.noise"][it]), it + 1)
        ral, rcl, rqv = g["td3.res"][it]
        assert (al is None) == bool(np.isnan(ral))
        assert abs(cl - rcl) <= 1e-5 * abs(rcl) and abs(qv - rqv) <= 1e-4 * max(1.0, abs(rqv))
        if al is not None:
            assert abs(al - ral) <= 1e-4 * max(1.0, abs(ral))
    for net, pre in ((actor, "td3.a4."), (critic, "td3.c4."), (actor_t, "td3.at4."), (critic_t, "td3.ct4.")):
        for k, v in net.items():
            d = np.abs(v.numpy() - g[pre + k])
            assert (d > 1e-6 + 1e-4 * np.abs(g[pre + k])).mean() <= 1e-3 and d.max() <= 2e-4, (pre, k, d.max())


def test_sac_restatement_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "td3sac_golden.npz"))
    pol, q1, q2 = _w(g, "sac.p0."), _w(g, "sac.q1_0."), _w(g, "sac.q2_0.")
    q1t, q2t = {k: v.clone() for k, v in q1.items()}, {k: v.clone() for k, v in q2.items()}
    z = lambda net, order: [torch.zeros_like(net[k]) for k in order]
    opt = {"pm": z(pol, R.SAC_POLICY_ORDER), "pv": z(pol, R.SAC_POLICY_ORDER), "q1m": z(q1, R.SAC_Q_ORDER), "q1v": z(q1, R.SAC_Q_ORDER),
           "q2m": z(q2, R.SAC_Q_ORDER), "q2v": z(q2, R.SAC_Q_ORDER), "step": 0}
    for it in range(3):
        l1, l2, pl = R.sac_update(pol, q1, q2, q1t, q2t, opt, _batch(g, g["sac.idx"][it]), torch.from_numpy(g["sac.eps_next"][it]),
                                  torch.from_numpy(g["sac.eps_new"][it]))
        r1, r2, rp = g["sac.res"][it]
        assert abs(l1 - r1) <= 1e-5 * abs(r1) and abs(l2 - r2) <= 1e-5 * abs(r2) and abs(pl - rp) <= 1e-4 * max(1.0, abs(rp))
    for net, pre in ((pol, "sac.p3."), (q1, "sac.q1_3."), (q2, "sac.q2_3."), (q1t, "sac.q1t_3."), (q2t, "sac.q2t_3.")):
        for k, v in net.items():
            d = np.abs(v.numpy() - g[pre + k])
            assert (d > 1e-6 + 1e-4 * np.abs(g[pre + k])).mean() <= 1e-3 and d.max() <= 6e-4, (pre, k, d.max())


def test_action_selection_restatement(golden_dir):
    """select_action (td3_agent.py:221-236) and SAC act (sac_agent.py:186-205) with injected normals."""
    g = np.load(os.path.join(golden_dir, "td3sac_golden.npz"))
    from erlfill_b200 import conditions as K
    b = K.EXPERIMENT_3["variant_25kg"]["bounds"] if isinstance(K.EXPERIMENT_3, dict) and "variant_25kg" in K.EXPERIMENT_3 \
        else list(K.EXPERIMENT_3.values())[0]["bounds"]
    lo = np.array([v[0] for v in b.values()], np.float64)
    hi = np.array([v[1] for v in b.values()], np.float64)
    actor = _w(g, "td3.a4.")
    for i in range(6):
        na = R.td3_actor_forward(actor, torch.from_numpy(g["td3.act_states"][i:i + 1])).numpy()[0].astype(np.float64)
        na = np.clip(na + 0.1 * g["td3.act_randn"][i], -1, 1)
        np.testing.assert_allclose(0.5 * (na + 1) * (hi - lo) + lo, g["td3.act_out"][i], rtol=1e-5, atol=1e-3)
    pol = _w(g, "sac.p3.")
    for i in range(6):
        a, _ = R.sac_policy_sample(pol, torch.from_numpy(g["sac.act_states"][i:i + 1]), torch.from_numpy(g["sac.act_eps"][i:i + 1]))
        np.testing.assert_allclose(a.numpy()[0] * (hi - lo) / 2 + (hi + lo) / 2, g["sac.act_out"][i], rtol=1e-5, atol=1e-2)
