"""CPU: host-side logic of the N = 1 drop-ins that needs no device — fixture sanity (the call sequence the GPU test
replays is the reference's train_ddpg order), the PrioritizedReplayBuffer slot rule including the scalar-type
effect on np.argmin, and the reference-compatible Adam state_dict layout."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_loop_fixture_is_train_ddpg_call_order():
    G = np.load(os.path.join(GOLD, "loop_golden.npz"))
    calls = [str(c) for c in G["calls"]]
    assert calls[0] == "reset" and calls.count("reset") == 3
    body = [c for c in calls if c != "reset"]
    assert body == ["act", "step", "update", "remember", "learn"] * (len(body) // 5)      # ddpg_agent.py:638-649
    assert G["learn.idx"].shape == (int(G["learn.ran"].sum()), 16)
    assert G["remember.len"].max() == 30 and list(G["remember.slot"][:30]) == list(range(30))
    # once full, the slot is np.argmin(priorities): the first float32-typed priority if any (numpy 2 scalar promotion), else 0
    pri = []
    for i, f32 in enumerate(G["remember.reward_f32"]):
        nan = float("nan")
        r = np.float32(1.0) if f32 else 1.0
        td = abs((r + 0.99 * nan) - nan)
        p = np.nan_to_num((abs(td) + 1e-4) ** 0.6, nan=1e-4)
        if len(pri) < 30:
            slot = len(pri); pri.append(p)
        else:
            slot = int(np.argmin(pri)); pri[slot] = p
        assert slot == int(G["remember.slot"][i]), i


def test_adam_facade_state_dict_is_torch_compatible():
    import erlfill_b200  # noqa: F401
    from erlfill_b200 import ddpg, dropin
    net = ddpg.FlatNet(ddpg.ACTOR_LAYOUT, sum(int(np.prod(s)) for _, s in ddpg.ACTOR_LAYOUT), "cpu")
    m, v = torch.rand_like(net.flat), torch.rand_like(net.flat)
    opt = dropin.AdamFacade(net, m, v, 1e-4, lambda: 7)
    sd = opt.state_dict()
    ps = [torch.nn.Parameter(torch.zeros(*s)) for _, s in ddpg.ACTOR_LAYOUT]
    ref = torch.optim.Adam(ps, lr=1e-4)
    ref.load_state_dict(sd)                                        # torch accepts the layout
    back = ref.state_dict()
    assert all(float(back["state"][i]["step"]) == 7.0 for i in back["state"])
    m2, v2 = torch.zeros_like(m), torch.zeros_like(v)
    opt2 = dropin.AdamFacade(net, m2, v2, 1e-4, lambda: 0)
    assert opt2.state_dict()["state"] == {}                        # an optimiser that never stepped has no state
    assert opt2.load_state_dict(back) == 7
    assert torch.equal(m, m2) and torch.equal(v, v2)


def test_net_facade_strict_keys_and_order():
    import erlfill_b200  # noqa: F401
    from erlfill_b200 import ddpg, dropin
    net = ddpg.FlatNet(ddpg.CRITIC_LAYOUT, sum(int(np.prod(s)) for _, s in ddpg.CRITIC_LAYOUT), "cpu")
    f = dropin.NetFacade("Critic", net, dropin.CRITIC_SD_ORDER)
    G = np.load(os.path.join(GOLD, "nets_golden.npz"))
    ref_keys = [k[len("critic."):] for k in G.files if k.startswith("critic.")]
    assert list(f.state_dict().keys()) == ref_keys
    f.load_state_dict({k: torch.from_numpy(np.array(G["critic." + k])) for k in ref_keys})
    assert torch.equal(net.views["net.4.weight"], torch.from_numpy(np.array(G["critic.net.4.weight"])))
    try:
        f.load_state_dict({k: torch.zeros(1) for k in ref_keys[:-1]})
    except RuntimeError as e:
        assert "missing keys" in str(e)
    else:
        raise AssertionError("strict load must raise on a missing key")
    a_keys = [k[len("actor."):] for k in G.files if k.startswith("actor.")]
    assert dropin.ACTOR_SD_ORDER == a_keys


def test_reference_noise_stream_is_cpythons_mersenne_twister():
    """envs._ScalarEnvBase.step (noise="reference") draws the run's uniforms vectorised from numpy's RandomState loaded with
    CPython `random`'s state, then advances `random` by exactly the draws consumed.  Both claims, on the CPU: the values equal
    random.uniform(15, 50) / random.uniform(-1, 1) bit for bit (VirtualWeightController.py:137,122), and the state written back
    equals the state after that many random.random() calls (getrandbits(64 k) == k calls)."""
    import random

    import numpy as np
    random.seed(20240607)
    for trial in range(4):
        st = random.getstate()
        key = np.fromiter(st[1], dtype=np.uint32, count=625)
        rs = np.random.RandomState(0)
        rs.set_state(("MT19937", key[:624], int(st[1][-1])))
        u = rs.random_sample(5001)
        ours = np.concatenate([[15.0 + 35.0 * u[0]], -1.0 + 2.0 * u[1:]])
        ref = np.array([random.uniform(15, 50)] + [random.uniform(-1, 1) for _ in range(5000)])
        assert np.array_equal(ours, ref)
        random.setstate(st)
        n_draws = 700 + 611 * trial
        random.getrandbits(64 * (1 + n_draws))
        ours_state = random.getstate()
        random.setstate(st)
        for _ in range(1 + n_draws):
            random.random()
        assert random.getstate() == ours_state
