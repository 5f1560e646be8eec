"""GPU: per-condition action scaling of a mixed-condition batch.

The reference builds one agent per operating condition (experiment_runner.py:504-510), each with scale_params / offset_params
from its own env.bounds (ddpg_agent.py:56-71).  A batched agent that serves three conditions at once must therefore map the
network output of every env with the row of ITS condition: a 3-condition batch has to equal three single-condition runs with
the same weights, for the actor (+OU), for what remember() stores, and for the update's actor forward."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _agents(n=96, seed=3):
    import erlfill_b200 as E
    K = E.conditions
    conds = list(K.EXPERIMENT_3.values())
    env3 = E.VecFillEnv(conds, n, kind="multi", max_steps=100, seed=seed)
    mixed = E.VecERDDPG(env3, seed=seed, batch_size=64, memory_size=4096)
    singles = []
    for c in range(3):
        # the envs of condition c in the mixed batch are i = c, c + 3, ...: same global ids => same Philox streams
        ids = torch.arange(c, n, 3)
        env1 = E.VecFillEnv(conds, len(ids), kind="multi", max_steps=100, seed=seed, cond_id=torch.full((len(ids),), c, dtype=torch.int32))
        singles.append((ids, env1))
    return E, conds, env3, mixed, singles


def test_mixed_batch_actor_equals_single_condition_runs():
    E, conds, env3, mixed, singles = _agents()
    L = E._lib
    n = env3.n_env
    g = torch.Generator().manual_seed(1)
    obs = (torch.rand(n, 2, generator=g) * 4).cuda()
    emo = torch.rand(n, 3, generator=g).cuda()
    env3.obs.copy_(obs); mixed.emotion.copy_(emo)
    a_mixed = mixed.act(add_noise=False).clone()
    from erlfill_b200 import nets
    w = mixed.learner.actor.state_dict()
    for c, (ids, _) in enumerate(singles):
        b = conds[c]["bounds"]
        sd = dict(w)
        sd["scale_params"] = torch.tensor([(b[k][1] - b[k][0]) / 2.0 for k in L.ACTION_ORDER], dtype=torch.float32)
        sd["offset_params"] = torch.tensor([(b[k][1] + b[k][0]) / 2.0 for k in L.ACTION_ORDER], dtype=torch.float32)
        p = nets.ActorParams(sd)
        a_single = nets.actor_forward(p, obs[ids.cuda()].contiguous(), emo[ids.cuda()].contiguous())
        assert torch.equal(a_single, a_mixed[ids.cuda()]), "condition %d" % c       # the same kernel, the same arithmetic: bit-exact
    # and the actions stay inside every env's OWN bounds (they were clipped to condition 0's before)
    lo = torch.tensor([[c["bounds"][k][0] for k in L.ACTION_ORDER] for c in conds]).cuda()[env3.cond_id.long()]
    hi = torch.tensor([[c["bounds"][k][1] for k in L.ACTION_ORDER] for c in conds]).cuda()[env3.cond_id.long()]
    a_noise = mixed.act(add_noise=True)
    assert bool(((a_noise >= lo) & (a_noise <= hi)).all())


def test_mixed_batch_remember_and_update_use_each_rows_condition():
    E, conds, env3, mixed, singles = _agents(n=192)
    L = E._lib
    from oracle import nets_ref as R
    for _ in range(3):
        mixed.rollout_step()
    torch.cuda.synchronize()
    n = env3.n_env
    rows = mixed.memory.buf[:3 * n].cpu()
    cid = rows[:, 19].long()
    assert torch.equal(cid, (torch.arange(3 * n) % n) % 3)                           # column 19 carries the condition index
    assert bool((rows[:, 2:12].abs() <= 1).all())
    # normalisation used the row's own condition: the last step's action, mapped back, reproduces the stored one
    sc, of = mixed.learner.scale_tab.cpu(), mixed.learner.offset_tab.cpu()
    last = rows[2 * n:3 * n]
    want = torch.clamp((mixed.action.cpu() - of[cid[:n]]) / sc[cid[:n]], -1, 1)
    assert torch.allclose(last[:, 2:12], want, atol=1e-6)
    # one update on these rows against the torch restatement with per-row scale / offset
    B = 64
    idx = torch.arange(0, 3 * n, 9)[:B].cuda()
    batch = mixed.memory.gather(idx).clone()
    ne = torch.tensor([0.4, 0.5, 0.2], device="cuda")
    a_sd, c_sd = mixed.learner.actor.state_dict(), mixed.learner.critic.state_dict()
    cpu = lambda d: {k: v.detach().cpu().clone() for k, v in d.items()}
    actor, actor_t, critic, critic_t = cpu(a_sd), cpu(a_sd), cpu(c_sd), cpu(c_sd)
    bc = batch.cpu()
    rc = bc[:, 19].long()
    for d in (actor, actor_t):
        d["scale_params"], d["offset_params"] = sc[rc], of[rc]                       # broadcasting [B,10] rows in raw * scale + offset
    opt = {"am": [torch.zeros_like(actor[k]) for k in R.ACTOR_PARAM_ORDER], "av": [torch.zeros_like(actor[k]) for k in R.ACTOR_PARAM_ORDER],
           "cm": [torch.zeros_like(critic[k]) for k in R.CRITIC_PARAM_ORDER], "cv": [torch.zeros_like(critic[k]) for k in R.CRITIC_PARAM_ORDER],
           "step": 0}
    bd = dict(states=bc[:, 0:2], actions=bc[:, 2:12], rewards=bc[:, 12], next_states=bc[:, 13:15], dones=bc[:, 15], emotions=bc[:, 16:19])
    cl, al, _ = R.ddpg_learn(actor, actor_t, critic, critic_t, opt, bd, ne.cpu(), 1, exact_target_emotion=True)
    rep = mixed.learner.update(batch, ne, 1, graph=False).cpu().numpy()
    np.testing.assert_allclose(rep[0], cl, rtol=1e-4)
    np.testing.assert_allclose(rep[1], al, rtol=1e-4)
    got = mixed.learner.actor.state_dict()
    for k in R.ACTOR_PARAM_ORDER:
        np.testing.assert_allclose(got[k].cpu().numpy(), actor[k].numpy(), rtol=2e-4, atol=2e-6, err_msg=k)
