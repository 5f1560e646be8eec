"""GPU: the reference's UNMODIFIED `train_ddpg` (DifferentModules/ddpg_agent.py:579-732) driving the drop-in env and agent.

SURVEY.md section 2 row 12: the callers "must keep running unchanged against the drop-in".  The reference trainer is imported through
oracle/ref_shim.py from /root/reference (build container) or from the byte-for-byte copy under baseline/_ref/ (GPU box; skipped when
neither exists) and touches the whole duck-typed surface: `env.max_steps = ...`, `env.reset()`, `agent.ou_noise.reset()/anneal()`,
`agent.act`, `env.step`, `agent.emotion.update / get_emotion`, `agent.remember`, `agent.learn()` (a dict of two losses),
`agent.actor_optim.param_groups[0]['lr']`, `info[...]` incl. the `action` dict, `agent.train_step += 1`, TensorBoard scalars, its logger."""
import os
import random
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _shim():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference tree (/root/reference or baseline/_ref/erlfill_reference)")
    ref_shim.install()
    return ref_shim


def test_reference_train_ddpg_runs_unmodified_on_the_dropins(tmp_path, capsys):
    _shim()
    import erlfill_b200 as E
    K = E.conditions
    cwd = os.getcwd()
    os.chdir(tmp_path)                      # the trainer creates runs/ and logs/ relative to the working directory
    try:
        import ddpg_agent as RD             # the reference module, unmodified
        cfg = K.EXPERIMENT_3["variant_20kg"]
        random.seed(3); np.random.seed(3); torch.manual_seed(3)
        env = E.MultiConditionWeightEnv({k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time", "bounds")})
        agent = E.DDPGAgent(env, device="cuda", use_td_error=True)
        agent.emotion = E.EmotionModule(device="cuda")
        env.attach_agent(agent)
        agent.batch_size = 8                # learn() from the 8th transition on
        # max_steps >= 10: the trainer evaluates `step % (max_steps // 10)` (:668)
        rewards = RD.train_ddpg(env, agent, episodes=3, max_steps=10, log_prefix="dropin_under_reference_trainer")
    finally:
        os.chdir(cwd)
    assert len(rewards) == 3 and all(np.isfinite(r) for r in rewards)
    assert agent.train_step == 1 + 3                                   # `agent.train_step += 1` per episode (:729)
    assert 10 <= len(agent.memory) <= 30 and env.max_steps == 10
    assert agent.learner.adam_step >= len(agent.memory) - 8            # one update per step once the minibatch is there
    # the trainer read the emotion-scaled learning rates through param_groups (:655-656) and wrote its TensorBoard run
    assert 1e-5 <= agent.actor_optim.param_groups[0]["lr"] <= 3e-4 and 1e-4 <= agent.critic_optim.param_groups[0]["lr"] <= 3e-3
    assert os.path.isdir(tmp_path / "runs" / "dropin_under_reference_trainer")
    for t in (agent.actor.state_dict(), agent.critic.state_dict()):
        assert all(torch.isfinite(v).all() for v in t.values())


def test_reference_env_under_the_dropin_agent_and_dropin_env_under_the_reference_agent(tmp_path):
    """The boundary is duck typing in both directions: a reference env stepped by the drop-in agent, and the drop-in env stepped by
    the reference's own DDPGAgent (CPU torch) — the same loop body, no adapter."""
    _shim()
    import erlfill_b200 as E
    from erlfill_b200 import loops
    import MutiConditionEnv as ME
    import ddpg_agent as RD
    from EmotionModule import EmotionModule as RefEmotion
    K = E.conditions
    cfg = {k: K.EXPERIMENT_3["variant_25kg"][k] for k in ("target_weight", "target_weight_err", "target_time", "bounds")}
    random.seed(1); np.random.seed(1); torch.manual_seed(1)
    devnull = open(os.devnull, "w")
    old = sys.stdout
    try:
        sys.stdout = devnull
        ref_env = ME.MultiConditionWeightEnv(dict(cfg))
        a = E.DDPGAgent(ref_env, device="cuda", use_td_error=True)
        a.emotion = E.EmotionModule(device="cuda")
        ref_env.attach_agent(a)
        a.batch_size = 8
        r1 = loops.train_ddpg(ref_env, a, episodes=2, max_steps=7)
        env = E.MultiConditionWeightEnv(dict(cfg))
        b = RD.DDPGAgent(env, device="cpu", use_td_error=True)
        b.emotion = RefEmotion(device="cpu")
        env.attach_agent(b)
        b.batch_size = 8
        r2 = loops.train_ddpg(env, b, episodes=2, max_steps=7)
    finally:
        sys.stdout = old
    assert len(r1) == 2 and len(r2) == 2 and all(np.isfinite(r1)) and all(np.isfinite(r2))
