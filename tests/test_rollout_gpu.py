"""GPU: the batched ER-DDPG loop (agents.VecERDDPG) against a step-by-step composition of the oracles."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nets_ref as R  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402


def test_rollout_matches_oracle_composition():
    import erlfill_b200 as E
    K = E.conditions
    conds = list(K.EXPERIMENT_3.values())[:1]          # one condition: the actor's scaling matches its bounds
    n, seed = 96, 5
    env = E.VecFillEnv(conds, n, kind="multi", max_steps=6, seed=seed, env_id_base=10)
    agent = E.VecERDDPG(env, seed=seed, batch_size=64, memory_size=1024)
    a_sd, c_sd, t_sd = E.init.er_ddpg_state_dicts(conds[0]["bounds"], seed=seed)
    c = conds[0]
    oc = [O.make_condition(c["target_weight"], c["target_weight_err"], c["target_time"], c["bounds"])]
    ref_env = O.BatchEnv(O.ENV_MULTI, oc, n, max_steps=6, seed=seed, env_id_base=10)
    scale, offset = a_sd["scale_params"].numpy(), a_sd["offset_params"].numpy()
    obs = env.obs.cpu().numpy().copy()
    ou = np.zeros((n, 10), np.float32)
    sigma = np.full(n, max(0.1, 0.6 * 0.9995))
    hist = [[] for _ in range(n)]
    cur = np.tile(np.array([0.5, 0.5, 0.0]), (n, 1))
    emotion = R.transformer_forward(t_sd, R.pad_history([])).numpy().repeat(n, 0)
    np.testing.assert_allclose(agent.emotion.cpu().numpy(), emotion, rtol=1e-5, atol=3e-6)
    for t in range(9):
        ns, r, d = agent.rollout_step(replay_reset=False)
        torch.cuda.synchronize()
        # actor + OU on the oracle side; the device's scaled action is used as the OU input so that the
        # comparison of the discontinuous int() truncation downstream stays bit-exact
        ref_scaled = R.actor_forward(a_sd, torch.from_numpy(obs), torch.from_numpy(emotion)).numpy()
        dev_scaled = agent.scaled.cpu().numpy()
        assert np.all(np.abs(dev_scaled - ref_scaled) <= 1e-5 * np.abs(ref_scaled) + 3e-5 * scale[None, :])
        act = np.zeros((n, 10), np.float32)
        for i in range(n):
            ou[i], act[i] = O.ou_act(ou[i], np.float32(sigma[i]), O.philox_randn10(seed, 10 + i, t), dev_scaled[i], scale, offset)
        dev_act = agent.action.cpu().numpy()
        assert np.abs(dev_act - act).max() <= 1e-3     # libm vs device log/sincos: <= 1 ulp of a 25 kg action
        ref_env.step(dev_act, t)
        assert np.array_equal(ns.cpu().numpy(), ref_env.next_state) and np.array_equal(d.cpu().numpy(), ref_env.done)
        assert np.array_equal(env.obs.cpu().numpy(), ref_env.obs)
        rr = r.cpu().numpy()
        assert np.all(np.abs(rr - ref_env.reward) <= np.spacing(np.abs(ref_env.reward)))
        for i in range(n):
            cur[i], pushed = O.emotion_update(cur[i], float(rr[i]), ref_env.next_state[i] - obs[i])
            hist[i].append(pushed)
        emotion = torch.cat([R.transformer_forward(t_sd, R.pad_history(hist[i])) for i in range(n)]).numpy()
        np.testing.assert_allclose(agent.emotion.cpu().numpy(), emotion, rtol=2e-5, atol=5e-6)
        # replay rows of this step
        rows = agent.memory.buf[t * n:(t + 1) * n].cpu().numpy()
        assert np.array_equal(rows[:, 0:2], obs) and np.array_equal(rows[:, 13:15], ref_env.next_state)
        assert np.array_equal(rows[:, 12], rr) and np.array_equal(rows[:, 15], ref_env.done.astype(np.float32))
        np.testing.assert_allclose(rows[:, 16:19], agent.emotion.cpu().numpy(), rtol=0, atol=0)
        for i in range(0, n, 9):
            assert np.array_equal(rows[i, 2:12], O.normalize_action(dev_act[i], scale, offset))
        for i in range(n):
            if ref_env.done[i]:
                ou[i] = 0
                sigma[i] = max(0.1, sigma[i] * 0.9995)
        obs = ref_env.obs.copy()
        ou = agent.ou_state.cpu().numpy().copy()        # keep the two sides in lock-step past 1-ulp OU differences
    assert int(agent.episodes.item()) == int(env.episode_counter.sum().item()) - n
    rep = agent.learn()
    assert rep is not None and torch.isfinite(rep).all()


def test_replay_branch_reset_and_learning_loop():
    import erlfill_b200 as E
    K = E.conditions
    conds = list(K.EXPERIMENT_3.values())
    env = E.VecFillEnv(conds, 300, kind="multi", max_steps=4, seed=1)
    agent = E.VecERDDPG(env, seed=1, batch_size=128, memory_size=2000)
    for t in range(12):
        agent.rollout_step()
        rep = agent.learn()
    torch.cuda.synchronize()
    assert len(agent.memory) == 2000 and agent.updates == 12
    assert torch.isfinite(rep).all() and torch.isfinite(agent.learner.actor.flat).all()
    # after the first reset the reference starts episodes from the doubly-normalised replay mean: tiny values
    ep = env.episode_counter.cpu().numpy()
    assert ep.min() >= 2
    assert agent.sync_counters() == 1 + int(agent.episodes.item()) // 300
