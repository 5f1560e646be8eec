"""GPU parity of the policy-side kernels (actor.cu, emotion_tf.cu) through the C ABI against the torch fp32
restatement (oracle/nets_ref.py) and the fixtures from the reference nn.Modules.

Tolerances (fp32 path): actor physical-unit actions rtol 1e-5 of the action range; critic/transformer
outputs rtol 1e-5 / atol 2e-6."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nets_ref as R  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402


def load_weights(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith(prefix)}


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "nets_golden.npz"))


def T(x, dev="cuda"):
    return torch.from_numpy(np.array(x)).to(dev)


def test_actor_forward_golden(G):
    from erlfill_b200 import nets, _lib as L
    W = load_weights(G, "actor.")
    P = nets.ActorParams(W)
    s, e = T(G["state"]), T(G["emotion"])
    scale = W["scale_params"].numpy()
    out, base = nets.actor_forward(P, s, e, want_base=True)
    np.testing.assert_allclose(out.cpu().numpy(), G["actor_eval"], rtol=0, atol=1e-5 * scale.max())
    assert np.all(np.abs(out.cpu().numpy() - G["actor_eval"]) <= 2e-5 * scale[None, :] + 1e-6)
    ref_out, ref_base = R.actor_forward(W, T(G["state"], "cpu"), T(G["emotion"], "cpu"), return_base=True)
    np.testing.assert_allclose(base.cpu().numpy(), ref_base.numpy(), rtol=1e-5, atol=2e-6)
    # train() mode with the reference's own dropout masks injected
    m1, m2 = T(G["actor_m1"]).to(torch.uint8).contiguous(), T(G["actor_m2"]).to(torch.uint8).contiguous()
    out_t = nets.actor_forward(P, s, e, dropout=L.DROPOUT_MASK, masks=(m1, m2))
    assert np.all(np.abs(out_t.cpu().numpy() - G["actor_train"]) <= 2e-5 * scale[None, :] + 1e-6)
    # broadcast emotion row (learn()'s next_emotion.expand, ddpg_agent.py:467)
    out_b = nets.actor_forward(P, s, e[3].contiguous())
    ref_b = R.actor_forward(W, T(G["state"], "cpu"), T(G["emotion"], "cpu")[3])
    assert np.all(np.abs(out_b.cpu().numpy() - ref_b.numpy()) <= 2e-5 * scale[None, :] + 1e-6)


def test_actor_ragged_large_batch_and_philox_dropout(G):
    from erlfill_b200 import nets, _lib as L
    W = load_weights(G, "actor.")
    P = nets.ActorParams(W)
    g = torch.Generator().manual_seed(3)
    n = 64 * 150 + 37
    s, e = torch.rand(n, 2, generator=g) * 5, torch.rand(n, 3, generator=g)
    out = nets.actor_forward(P, s.cuda(), e.cuda())
    ref, ref_base = R.actor_forward(W, s, e, return_base=True)
    scale = W["scale_params"].numpy()
    # contract: 1e-5 relative on the physical value (one binary32 ulp of 24950 is 3.9e-5 of its range)
    # plus 2e-5 of the action range
    err = np.abs(out.cpu().numpy() - ref.numpy()) - 1e-5 * np.abs(ref.numpy())
    # -sign(base) (ddpg_agent.py:113) is discontinuous at base == 0: leave out the rows that sit on it
    ok = np.abs(ref_base.numpy()) > 1e-4
    assert ok.mean() > 0.999 and (err / scale[None, :])[ok].max() <= 2e-5, (ok.mean(), (err / scale[None, :])[ok].max())
    # Philox dropout: deterministic in (seed, call_idx, row id), independent of batch split, keep rate 0.8
    a = nets.actor_forward(P, s.cuda(), e.cuda(), dropout=L.DROPOUT_PHILOX, seed=11, call_idx=5)
    b = nets.actor_forward(P, s.cuda(), e.cuda(), dropout=L.DROPOUT_PHILOX, seed=11, call_idx=5)
    c = nets.actor_forward(P, s[1000:].contiguous().cuda(), e[1000:].contiguous().cuda(), dropout=L.DROPOUT_PHILOX,
                           seed=11, call_idx=5, row_id_base=1000)
    d = nets.actor_forward(P, s.cuda(), e.cuda(), dropout=L.DROPOUT_PHILOX, seed=11, call_idx=6)
    assert torch.equal(a, b) and torch.equal(a[1000:], c) and not torch.equal(a, d)
    assert not torch.equal(a, out)


def test_ou_act_matches_oracle(G):
    from erlfill_b200 import nets
    W = load_weights(G, "actor.")
    P = nets.ActorParams(W)
    n = 500
    g = torch.Generator().manual_seed(4)
    s, e = (torch.rand(n, 2, generator=g) * 5).cuda(), torch.rand(n, 3, generator=g).cuda()
    ou_state = torch.zeros(n, 10, device="cuda")
    sigma = torch.full((n,), 0.6, device="cuda")
    final = torch.empty(n, 10, device="cuda")
    scale, offset = W["scale_params"].numpy(), W["offset_params"].numpy()
    ref_state = np.zeros((n, 10), np.float32)
    for step in range(3):
        # injected normals (numpy stream like OUNoise.sample) on odd steps, Philox Box-Muller on even steps
        randn = torch.randn(n, 10, generator=g).cuda() if step % 2 else None
        scaled = nets.actor_forward(P, s, e, ou=dict(state=ou_state, sigma=sigma, randn=randn, final=final, seed=99,
                                                     env_id_base=40, step_idx=step))
        torch.cuda.synchronize()
        sc, fin, st = scaled.cpu().numpy(), final.cpu().numpy(), ou_state.cpu().numpy()
        for i in range(0, n, 7):
            rn = randn[i].cpu().numpy() if randn is not None else O.philox_randn10(99, 40 + i, step)
            rs, rf = O.ou_act(ref_state[i], 0.6, rn, sc[i], scale, offset)
            ref_state[i] = rs
            if randn is not None:
                assert np.array_equal(st[i], rs) and np.array_equal(fin[i], rf), (step, i)
            else:   # device log/sin/cos vs libm: identical after rounding to binary32 except for rare 1-ulp cases
                np.testing.assert_allclose(st[i], rs, rtol=0, atol=2e-7 * max(1.0, np.abs(rs).max()))
                np.testing.assert_allclose(fin[i], rf, rtol=2e-7, atol=1e-3)
        ref_state[:] = st   # keep in lock-step for rows not checked


def test_emotion_update_golden(golden_dir):
    from erlfill_b200 import nets
    g = np.load(os.path.join(golden_dir, "emotion_update_golden.npz"))
    n = 5
    emo = nets.EmotionStateVec(n)
    zeros = torch.zeros(n, 2, device="cuda")
    for t in range(len(g["rewards"])):
        r = torch.full((n,), float(g["rewards"][t]), dtype=torch.float32, device="cuda")
        ns = T(np.tile(g["moves"][t], (n, 1)))
        nets.emotion_update(emo, r, zeros, ns.contiguous())
        cur = emo.current.cpu().numpy()
        # rewards above 2^24 ulp are binary32-rounded on the device input; contract 1e-5
        np.testing.assert_allclose(cur[:, 2], g["current"][t], rtol=1e-5, atol=3e-7)
    assert int(emo.hist_len[0].item()) == 20
    head = int(emo.hist_head[0].item())
    hist = emo.history.cpu().numpy()[:, :, 1]
    ordered = np.stack([hist[(head + j) % 20] for j in range(20)])
    np.testing.assert_allclose(ordered, g["history_last20"], rtol=1e-5, atol=3e-7)


def _masks_from_golden(G, n):
    attn = np.stack([G[f"trans_attn{l}"] for l in range(4)])
    d1 = np.stack([G[f"trans_d1_{l}"] for l in range(4)])
    ff = np.stack([np.unpackbits(G[f"trans_ff{l}"], axis=-1)[..., :2048] for l in range(4)])
    d2 = np.stack([G[f"trans_d2_{l}"] for l in range(4)])
    return {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.uint8))).cuda() for k, v in
            dict(attn=attn, d1=d1, ff=ff, d2=d2).items()}


def test_transformer_forward_golden(G):
    from erlfill_b200 import nets, _lib as L
    W = load_weights(G, "trans.")
    P = nets.TransformerParams(W)
    for key_in, key_out in (("seqs", "trans_eval"), ("seq1", "trans_eval_s1"), ("seq7", "trans_eval_s7")):
        out = nets.emotion_forward(P, seq=T(G[key_in]).contiguous())
        np.testing.assert_allclose(out.cpu().numpy(), G[key_out], rtol=1e-5, atol=3e-6)
    out = nets.emotion_forward(P, seq=T(G["seqs"]).contiguous(), dropout=L.DROPOUT_MASK,
                               masks=_masks_from_golden(G, 12))
    np.testing.assert_allclose(out.cpu().numpy(), G["trans_train"], rtol=1e-5, atol=3e-6)
    # state_dict round trip keeps the reference's keys
    sd = P.state_dict()
    assert set(sd) == set(W) and all(torch.equal(sd[k].cpu(), W[k]) for k in W)


def test_transformer_from_history_ring(G):
    """get_state() semantics: empty history -> neutral token; partial history front-padded with the oldest."""
    from erlfill_b200 import nets
    W = load_weights(G, "trans.")
    P = nets.TransformerParams(W)
    n = 9
    emo = nets.EmotionStateVec(n)
    out = nets.emotion_forward(P, emo=emo)
    np.testing.assert_allclose(out.cpu().numpy(), np.tile(G["trans_eval_s1"], (n, 1)), rtol=1e-5, atol=3e-6)
    rng = np.random.RandomState(2)
    hist = [[] for _ in range(n)]
    zeros = torch.zeros(n, 2, device="cuda")
    for t in range(27):
        r = torch.tensor(rng.uniform(-3, 3, n), dtype=torch.float32, device="cuda")
        ns = torch.tensor(rng.uniform(-1, 1, (n, 2)), dtype=torch.float32, device="cuda")
        nets.emotion_update(emo, r, zeros, ns)
        cur = emo.current.cpu().numpy()
        for i in range(n):
            hist[i].append(cur[:, i].astype(np.float32))
        if t in (0, 4, 18, 19, 20, 26):
            out = nets.emotion_forward(P, emo=emo).cpu().numpy()
            ref = torch.cat([R.transformer_forward(W, R.pad_history(hist[i])) for i in range(n)]).numpy()
            np.testing.assert_allclose(out, ref, rtol=1e-5, atol=3e-6)
