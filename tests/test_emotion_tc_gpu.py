"""GPU parity of the tensor-core EmotionTransformer (emotion_tf_tc.cu: tcgen05.mma, bf16 operands, fp32
accumulation) through the C ABI.

Checks with dropout off (module.eval()), and at the end of the file train() mode (injected masks, Philox masks):
  1. against a torch emulation of the kernel's own arithmetic (same bf16 rounding points: activations entering
     each projection, q/k/v and the softmax probabilities, the ReLU'd hidden activation, every weight; biases as
     hi+lo bf16 pairs): atol 2e-4 -- this
     pins the kernel's indexing/pipeline, not just its rough magnitude;
  2. against the reference fp32 outputs (fixtures from the reference nn.Module, and oracle/nets_ref.py):
     stated bf16 tolerance 1e-2 absolute on the [0,1] outputs (typical 1e-3 .. 4e-3).
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nets_ref as R  # noqa: E402

BF16_ATOL = 1e-2
EMUL_ATOL = 2e-4


def load_weights(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith(prefix)}


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "nets_golden.npz"))


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def emulate_tc(W, seq):
    """float64 torch model of emotion_tf_tc.cu for seq [n,S,3] (bf16 rounding where the kernel rounds)."""
    W = {k: v.double() for k, v in W.items()}
    n, S, _ = seq.shape
    x = seq.double() @ W["embedding.weight"].T + W["embedding.bias"] + W["positional_encoding"][:S]
    qs = math.log2(math.e) / math.sqrt(8.0)

    def lin(a_bf, w, b, row_scale=None):
        # [a | 1 | 1] . [bf16(w) | bf16(b) | bf16(b - bf16(b))]^T with fp32 scaling applied before the rounding
        w, b = w.float(), b.float()
        if row_scale is not None:
            w, b = w * row_scale[:, None], b * row_scale
        b_hi = bf(b)
        b_lo = bf((b.double() - b_hi).float())
        return a_bf @ bf(w).T + b_hi + b_lo

    q_scale = torch.ones(96, dtype=torch.float32)
    q_scale[:32] = float(np.float32(qs))
    for l in range(4):
        p = f"transformer_encoder.layers.{l}."
        qkv = lin(bf(x.float()), W[p + "self_attn.in_proj_weight"], W[p + "self_attn.in_proj_bias"], q_scale)
        q, k, v = qkv[..., :32], qkv[..., 32:64], qkv[..., 64:]
        q = q.view(n, S, 4, 8).transpose(1, 2)
        k = k.view(n, S, 4, 8).transpose(1, 2)
        v = v.view(n, S, 4, 8).transpose(1, 2)
        sc = bf(q.float()) @ bf(k.float()).transpose(-1, -2)
        pr = torch.exp2(sc - sc.max(-1, keepdim=True).values).float()
        o = (bf(pr) @ bf(v.float())) / pr.double().sum(-1, keepdim=True)
        o = o.transpose(1, 2).reshape(n, S, 32)
        att = lin(bf(o.float()), W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
        x = torch.nn.functional.layer_norm((x + att).float(), (32,), W[p + "norm1.weight"].float(), W[p + "norm1.bias"].float(),
                                           1e-5).double()
        h = lin(bf(x.float()), W[p + "linear1.weight"], W[p + "linear1.bias"])
        hb = bf(torch.relu(h).float())
        y = hb @ bf(W[p + "linear2.weight"].float()).T + W[p + "linear2.bias"]
        x = torch.nn.functional.layer_norm((x + y).float(), (32,), W[p + "norm2.weight"].float(), W[p + "norm2.bias"].float(),
                                           1e-5).double()
    out = x[:, -1] @ W["fc_out.weight"].T + W["fc_out.bias"]
    return (0.5 * torch.tanh(out) + 0.5).float()


def T(x, dev="cuda"):
    return torch.from_numpy(np.array(x)).to(dev)


def test_tc_matches_reference_fixtures(G):
    from erlfill_b200 import nets
    W = load_weights(G, "trans.")
    P = nets.TransformerParams(W)
    for key_in, key_out in (("seqs", "trans_eval"), ("seq1", "trans_eval_s1"), ("seq7", "trans_eval_s7")):
        seq = T(G[key_in]).contiguous()
        out = nets.emotion_forward(P, seq=seq, precision="bf16").cpu()
        err_ref = np.abs(out.numpy() - G[key_out]).max()
        d_emul = (out - emulate_tc(W, seq.cpu())).abs().flatten()
        err_emul = d_emul.max().item()
        print(key_in, "max |tc - reference fp32| = %.2e, |tc - bf16 emulation| median %.2e max %.2e" %
              (err_ref, d_emul.median().item(), err_emul))
        # a bf16 rounding that falls the other way (fp32 vs fp64 accumulation order) moves an output by up to ~1e-3
        assert d_emul.median().item() <= EMUL_ATOL and err_emul <= 2e-3, (key_in, err_emul)
        assert err_ref <= BF16_ATOL, (key_in, err_ref)


def test_tc_large_ragged_batch_vs_fp32_kernel(G):
    """Many groups per CTA (persistent loop, weight ring wrap-around, barrier phases) and a ragged tail."""
    from erlfill_b200 import nets
    W = load_weights(G, "trans.")
    P = nets.TransformerParams(W)
    g = torch.Generator().manual_seed(5)
    for n in (1, 11, 12, 13, 148 * 12 * 3 + 5):
        seq = torch.rand(n, 20, 3, generator=g).cuda()
        a = nets.emotion_forward(P, seq=seq, precision="bf16")
        b = nets.emotion_forward(P, seq=seq)
        err = (a - b).abs().max().item()
        assert torch.isfinite(a).all() and err <= BF16_ATOL, (n, err)
        a2 = nets.emotion_forward(P, seq=seq, precision="bf16")
        assert torch.equal(a, a2)                      # deterministic
    # against the emulation on many random sequences: a bf16 rounding that falls the other way (fp32 vs fp64
    # accumulation order) moves an output by up to ~1e-3, so bound the bulk tightly and the tail loosely
    err_emul = (a[:500].cpu() - emulate_tc(W, seq[:500].cpu())).abs().flatten()
    assert err_emul.quantile(0.9).item() <= EMUL_ATOL and err_emul.max().item() <= 2e-3, (err_emul.quantile(0.9), err_emul.max())


def test_tc_history_ring_sources(G):
    """get_state() semantics through the tc kernel: empty history -> neutral token; partial history front-padded."""
    from erlfill_b200 import nets
    W = load_weights(G, "trans.")
    P = nets.TransformerParams(W)
    n = 29
    emo = nets.EmotionStateVec(n)
    out = nets.emotion_forward(P, emo=emo, precision="bf16")
    assert np.abs(out.cpu().numpy() - np.tile(G["trans_eval_s1"], (n, 1))).max() <= BF16_ATOL
    rng = np.random.RandomState(2)
    zeros = torch.zeros(n, 2, device="cuda")
    for t in range(27):
        r = torch.tensor(rng.uniform(-3, 3, n), dtype=torch.float32, device="cuda")
        ns = torch.tensor(rng.uniform(-1, 1, (n, 2)), dtype=torch.float32, device="cuda")
        nets.emotion_update(emo, r, zeros, ns)
        if t in (0, 4, 18, 19, 20, 26):
            a = nets.emotion_forward(P, emo=emo, precision="bf16")
            b = nets.emotion_forward(P, emo=emo)
            assert (a - b).abs().max().item() <= BF16_ATOL, t
    # mixed: some envs with an empty history (S = 1) next to full ones in the same token tile
    emo.hist_len[::3] = 0
    a = nets.emotion_forward(P, emo=emo, precision="bf16")
    b = nets.emotion_forward(P, emo=emo)
    assert (a - b).abs().max().item() <= BF16_ATOL


# ---- train() mode: the four dropout layers of every encoder layer (the reference never calls .eval(), EmotionModule.py:9,35) -------------
def _masks_from_golden(G):
    attn = np.stack([G[f"trans_attn{l}"] for l in range(4)])
    d1 = np.stack([G[f"trans_d1_{l}"] for l in range(4)])
    ff = np.stack([np.unpackbits(G[f"trans_ff{l}"], axis=-1)[..., :2048] for l in range(4)])
    d2 = np.stack([G[f"trans_d2_{l}"] for l in range(4)])
    return {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.uint8))).cuda() for k, v in dict(attn=attn, d1=d1, ff=ff, d2=d2).items()}


def test_tc_train_mode_with_the_references_own_masks(G):
    """The reference module's train()-mode output with its dropout masks replayed (oracle/gen_golden.py nets): the tensor-core
    kernel with the same masks injected, at the stated bf16 tolerance; the fp32 kernel beside it."""
    from erlfill_b200 import nets, _lib as L
    W = load_weights(G, "trans.")
    P = nets.TransformerParams(W)
    seq = T(G["seqs"]).contiguous()
    masks = _masks_from_golden(G)
    out = nets.emotion_forward(P, seq=seq, dropout=L.DROPOUT_MASK, masks=masks, precision="bf16").cpu().numpy()
    f32 = nets.emotion_forward(P, seq=seq, dropout=L.DROPOUT_MASK, masks=masks).cpu().numpy()
    err = np.abs(out - G["trans_train"]).max()
    print("train mode: max |tc - reference| = %.2e, |tc - fp32 kernel| = %.2e" % (err, np.abs(out - f32).max()))
    assert err <= BF16_ATOL
    assert np.abs(G["trans_train"] - G["trans_eval"]).max() > 5 * BF16_ATOL      # the masks matter far more than the tolerance


def test_tc_philox_dropout_draws_the_fp32_kernels_masks(G):
    """ERLFILL_DROPOUT_PHILOX: both kernels draw the masks of csrc/tf_dropout.cuh, so with the same (seed, call index, row ids) the
    tensor-core output equals the exact fp32 kernel's within the bf16 tolerance — across groups, ragged tails and partial histories."""
    from erlfill_b200 import nets, _lib as L
    W = load_weights(G, "trans.")
    P = nets.TransformerParams(W)
    g = torch.Generator().manual_seed(9)
    for n in (1, 13, 148 * 12 + 7):
        seq = torch.rand(n, 20, 3, generator=g).cuda()
        kw = dict(dropout=L.DROPOUT_PHILOX, seed=0xABCDEF12345, call_idx=77, row_id_base=1000)
        a = nets.emotion_forward(P, seq=seq, precision="bf16", **kw)
        b = nets.emotion_forward(P, seq=seq, **kw)
        err = (a - b).abs().max().item()
        assert torch.isfinite(a).all() and err <= BF16_ATOL, (n, err)
        assert torch.equal(a, nets.emotion_forward(P, seq=seq, precision="bf16", **kw))          # deterministic
        ev = nets.emotion_forward(P, seq=seq, precision="bf16")
        kw2 = dict(kw, call_idx=78)
        assert (a - ev).abs().max().item() > 5 * BF16_ATOL or n == 1
        assert (a - nets.emotion_forward(P, seq=seq, precision="bf16", **kw2)).abs().max().item() > 1e-3
    # history-ring source with mixed sequence lengths (S = 1 rows draw no masks they should not)
    n = 29
    emo = nets.EmotionStateVec(n)
    rng = np.random.RandomState(2)
    zeros = torch.zeros(n, 2, device="cuda")
    for t in range(23):
        nets.emotion_update(emo, torch.tensor(rng.uniform(-3, 3, n), dtype=torch.float32, device="cuda"), zeros,
                            torch.tensor(rng.uniform(-1, 1, (n, 2)), dtype=torch.float32, device="cuda"))
    emo.hist_len[::3] = 0
    kw = dict(dropout=L.DROPOUT_PHILOX, seed=5, call_idx=3, row_id_base=64)
    a = nets.emotion_forward(P, emo=emo, precision="bf16", **kw)
    b = nets.emotion_forward(P, emo=emo, **kw)
    assert (a - b).abs().max().item() <= BF16_ATOL
