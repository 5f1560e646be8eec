"""CPU: pins oracle/nets_ref.py (plain torch fp32 restatement of Actor / Critic / EmotionTransformer /
DDPGAgent.learn) to fixtures produced from the unmodified reference nn.Modules."""
import os

import numpy as np
import pytest
import torch

from oracle import nets_ref as R


def load_weights(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith(prefix)}


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "nets_golden.npz"))


def T(x):
    return torch.from_numpy(np.array(x))


def test_param_counts(G):
    a, c, t = load_weights(G, "actor."), load_weights(G, "critic."), load_weights(G, "trans.")
    assert sum(a[k].numel() for k in R.ACTOR_PARAM_ORDER) == 26216
    assert sum(c[k].numel() for k in R.CRITIC_PARAM_ORDER) == 103937
    assert sum(v.numel() for v in t.values()) == 550883


def test_actor_eval_and_train(G):
    W = load_weights(G, "actor.")
    s, e = T(G["state"]), T(G["emotion"])
    torch.testing.assert_close(R.actor_forward(W, s, e), T(G["actor_eval"]), rtol=1e-6, atol=1e-3)
    out = R.actor_forward(W, s, e, (T(G["actor_m1"]), T(G["actor_m2"])))
    torch.testing.assert_close(out, T(G["actor_train"]), rtol=1e-6, atol=1e-3)
    assert not np.allclose(G["actor_eval"], G["actor_train"])


def test_critic_eval_and_train(G):
    W = load_weights(G, "critic.")
    s, a, e = T(G["state"]), T(G["action"]), T(G["emotion"])
    torch.testing.assert_close(R.critic_forward(W, s, a, e), T(G["critic_eval"]), rtol=1e-5, atol=1e-6)
    out = R.critic_forward(W, s, a, e, (T(G["critic_m1"]), T(G["critic_m2"])))
    torch.testing.assert_close(out, T(G["critic_train"]), rtol=1e-5, atol=1e-6)


def test_transformer_eval(G):
    W = load_weights(G, "trans.")
    torch.testing.assert_close(R.transformer_forward(W, T(G["seqs"])), T(G["trans_eval"]), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(R.transformer_forward(W, T(G["seq1"])), T(G["trans_eval_s1"]), rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(R.transformer_forward(W, T(G["seq7"])), T(G["trans_eval_s7"]), rtol=1e-5, atol=2e-6)


def test_transformer_train_with_replayed_masks(G):
    W = load_weights(G, "trans.")
    masks = []
    for l in range(4):
        ff = np.unpackbits(G[f"trans_ff{l}"], axis=-1)[..., :2048]
        masks.append({"attn": T(G[f"trans_attn{l}"]).float(), "d1": T(G[f"trans_d1_{l}"]).float(),
                      "ff": torch.from_numpy(ff).float(), "d2": T(G[f"trans_d2_{l}"]).float()})
    out = R.transformer_forward(W, T(G["seqs"]), masks)
    torch.testing.assert_close(out, T(G["trans_train"]), rtol=1e-5, atol=2e-6)
    assert not np.allclose(G["trans_eval"], G["trans_train"], atol=1e-3)


def test_pad_history():
    assert R.pad_history([]).shape == (1, 1, 3)
    h = [np.array([i, i, i], np.float32) for i in range(3)]
    p = R.pad_history(h)
    assert p.shape == (1, 20, 3) and torch.all(p[0, :17] == 0) and p[0, 19, 0] == 2
    h = [np.array([i, i, i], np.float32) for i in range(25)]
    assert R.pad_history(h)[0, 0, 0] == 5


def test_learn_three_updates(golden_dir):
    g = np.load(os.path.join(golden_dir, "learn_golden.npz"))
    actor, critic = load_weights(g, "a0."), load_weights(g, "c0.")
    actor_t = {k: v.clone() for k, v in actor.items()}
    critic_t = {k: v.clone() for k, v in critic.items()}
    tw = load_weights(g, "trans.")
    opt = {"am": [torch.zeros_like(actor[k]) for k in R.ACTOR_PARAM_ORDER],
           "av": [torch.zeros_like(actor[k]) for k in R.ACTOR_PARAM_ORDER],
           "cm": [torch.zeros_like(critic[k]) for k in R.CRITIC_PARAM_ORDER],
           "cv": [torch.zeros_like(critic[k]) for k in R.CRITIC_PARAM_ORDER], "step": 0}
    # the fresh get_emotion() inside learn (ddpg_agent.py:466), eval mode: one M2 call on the history
    ne = R.transformer_forward(tw, R.pad_history(list(g["history"])))[0]
    for it in range(3):
        torch.testing.assert_close(ne, T(g["next_emotion"][it]), rtol=1e-5, atol=2e-6)
        idx = g["idx"][it]
        batch = {"states": T(g["buf_states"][idx]), "actions": T(g["buf_actions"][idx]),
                 "rewards": T(g["buf_rewards"][idx]).float(), "next_states": T(g["buf_next_states"][idx]),
                 "dones": T(g["buf_dones"][idx]).float(), "emotions": T(g["buf_emotions"][idx])}
        cl, al, (c_lr, a_lr) = R.ddpg_learn(actor, actor_t, critic, critic_t, opt, batch, T(g["next_emotion"][it]),
                                            int(g["train_step"]))
        assert abs(cl - g["critic_loss"][it]) <= 2e-5 * abs(g["critic_loss"][it])
        assert abs(al - g["actor_loss"][it]) <= 2e-5 * abs(g["actor_loss"][it])
        assert abs(c_lr - g["lrs"][it][0]) < 1e-12 and abs(a_lr - g["lrs"][it][1]) < 1e-12
        for k in R.ACTOR_PARAM_ORDER:
            torch.testing.assert_close(actor[k], T(g[f"a{it + 1}.{k}"]), rtol=2e-4, atol=2e-6)
            torch.testing.assert_close(actor_t[k], T(g[f"at{it + 1}.{k}"]), rtol=2e-4, atol=2e-6)
        for k in R.CRITIC_PARAM_ORDER:
            torch.testing.assert_close(critic[k], T(g[f"c{it + 1}.{k}"]), rtol=2e-4, atol=5e-6)
            torch.testing.assert_close(critic_t[k], T(g[f"ct{it + 1}.{k}"]), rtol=2e-4, atol=5e-6)


def golden_masks(g, it, B=128):
    """The ten keep-masks of update `it` of tests/golden/learn_train_golden.npz as the oracle's dict and as a flat list."""
    widths = [128, 128, 256, 256, 256, 256, 128, 128, 256, 256]
    ms = [torch.from_numpy(np.unpackbits(g[f"mask{it}_{s}"], axis=-1)[:, :w].astype(np.float32)) for s, w in enumerate(widths)]
    return {"at": (ms[0], ms[1]), "ct": (ms[2], ms[3]), "c": (ms[4], ms[5]), "a": (ms[6], ms[7]), "c2": (ms[8], ms[9])}, ms


def test_learn_train_mode_three_updates(golden_dir):
    """The mode the reference really runs: all four networks in train() — Dropout(0.2) active in the five forwards of learn().
    The oracle with the replayed masks reproduces the reference's losses, learning rates and parameters."""
    g = np.load(os.path.join(golden_dir, "learn_train_golden.npz"))
    actor, critic = load_weights(g, "a0."), load_weights(g, "c0.")
    actor_t = {k: v.clone() for k, v in actor.items()}
    critic_t = {k: v.clone() for k, v in critic.items()}
    opt = {"am": [torch.zeros_like(actor[k]) for k in R.ACTOR_PARAM_ORDER],
           "av": [torch.zeros_like(actor[k]) for k in R.ACTOR_PARAM_ORDER],
           "cm": [torch.zeros_like(critic[k]) for k in R.CRITIC_PARAM_ORDER],
           "cv": [torch.zeros_like(critic[k]) for k in R.CRITIC_PARAM_ORDER], "step": 0}
    for it in range(3):
        idx = g["idx"][it]
        batch = {"states": T(g["buf_states"][idx]), "actions": T(g["buf_actions"][idx]),
                 "rewards": T(g["buf_rewards"][idx]).float(), "next_states": T(g["buf_next_states"][idx]),
                 "dones": T(g["buf_dones"][idx]).float(), "emotions": T(g["buf_emotions"][idx])}
        masks, _ = golden_masks(g, it)
        cl, al, (c_lr, a_lr) = R.ddpg_learn(actor, actor_t, critic, critic_t, opt, batch, T(g["next_emotion"][it]),
                                            int(g["train_step"]), masks=masks)
        assert abs(cl - g["critic_loss"][it]) <= 2e-5 * abs(g["critic_loss"][it])
        assert abs(al - g["actor_loss"][it]) <= 2e-5 * abs(g["actor_loss"][it])
        assert abs(c_lr - g["lrs"][it][0]) < 1e-12 and abs(a_lr - g["lrs"][it][1]) < 1e-12
        for k in R.ACTOR_PARAM_ORDER:
            torch.testing.assert_close(actor[k], T(g[f"a{it + 1}.{k}"]), rtol=2e-4, atol=2e-6)
            torch.testing.assert_close(actor_t[k], T(g[f"at{it + 1}.{k}"]), rtol=2e-4, atol=2e-6)
        for k in R.CRITIC_PARAM_ORDER:
            torch.testing.assert_close(critic[k], T(g[f"c{it + 1}.{k}"]), rtol=2e-4, atol=5e-6)
            torch.testing.assert_close(critic_t[k], T(g[f"ct{it + 1}.{k}"]), rtol=2e-4, atol=5e-6)
    e = np.load(os.path.join(golden_dir, "learn_golden.npz"))
    assert abs(g["critic_loss"][0] - e["critic_loss"][0]) > 1e-3          # a different algorithm from the eval()-mode one


def test_init_reproduces_reference_construction(G):
    """erl-fill_b200/init.py consumes torch's RNG in the reference's order: same seed, same weights."""
    import erlfill_b200
    from erlfill_b200 import init, conditions as K
    a, c, t = init.er_ddpg_state_dicts(K.EXPERIMENT_3["variant_25kg"]["bounds"], seed=0)
    for prefix, sd in (("actor.", a), ("critic.", c), ("trans.", t)):
        ref = load_weights(G, prefix)
        assert set(ref) == set(sd), (prefix, set(ref) ^ set(sd))
        for k in ref:
            assert torch.equal(ref[k], sd[k]), prefix + k
