"""Host-side parameter containers and thin wrappers over the policy-side C ABI (actor.cu, emotion_tf.cu).

Parameters are kept as torch tensors keyed exactly like the reference's state_dict (so reference
checkpoints interchange, ddpg_agent.py:526-560); the transformer's per-layer tensors are additionally
stacked on a leading [4] axis for the kernels.
"""
import ctypes as C

import torch

from . import _lib as L

ACTOR_KEYS = ["input_layer.weight", "input_layer.bias", "hidden.2.weight", "hidden.2.bias", "hidden.5.weight",
              "hidden.5.bias", "hidden.7.weight", "hidden.7.bias", "emotion_weights", "scale_params", "offset_params"]
_ACTOR_FIELDS = ["w_in", "b_in", "w_h2", "b_h2", "w_h5", "b_h5", "w_h7", "b_h7", "emotion_weights", "scale", "offset"]

_TF_LAYER = [("in_proj_w", "self_attn.in_proj_weight"), ("in_proj_b", "self_attn.in_proj_bias"),
             ("out_proj_w", "self_attn.out_proj.weight"), ("out_proj_b", "self_attn.out_proj.bias"),
             ("lin1_w", "linear1.weight"), ("lin1_b", "linear1.bias"), ("lin2_w", "linear2.weight"),
             ("lin2_b", "linear2.bias"), ("norm1_w", "norm1.weight"), ("norm1_b", "norm1.bias"),
             ("norm2_w", "norm2.weight"), ("norm2_b", "norm2.bias")]
_TF_GLOBAL = [("emb_w", "embedding.weight"), ("emb_b", "embedding.bias"), ("pos", "positional_encoding"),
              ("fc_w", "fc_out.weight"), ("fc_b", "fc_out.bias")]
N_LAYERS = 4


def _dev_f32(x, device):
    return torch.as_tensor(x).to(device=device, dtype=torch.float32).contiguous()


class ActorParams:
    """Actor parameters + scaling buffers (ddpg_agent.py:31-72)."""

    def __init__(self, state_dict, device="cuda"):
        self.device = torch.device(device)
        self.t = {k: _dev_f32(state_dict[k], self.device) for k in ACTOR_KEYS}

    def struct(self):
        return L.ActorWeights(*[L.ptr(self.t[k]) for k in ACTOR_KEYS])

    def state_dict(self):
        return {k: v.clone() for k, v in self.t.items()}


class TransformerParams:
    """EmotionTransformer parameters (EmotionModule.py:27-40), stacked per layer."""

    def __init__(self, state_dict, device="cuda"):
        self.device = torch.device(device)
        self.t = {}
        for f, k in _TF_GLOBAL:
            self.t[f] = _dev_f32(state_dict[k], self.device)
        for f, k in _TF_LAYER:
            self.t[f] = torch.stack([_dev_f32(state_dict[f"transformer_encoder.layers.{l}.{k}"], self.device)
                                     for l in range(N_LAYERS)]).contiguous()

        self._tc_image = None

    def struct(self):
        order = [n for n, _ in L.TransformerWeights._fields_]
        return L.TransformerWeights(*[L.ptr(self.t[n]) for n in order])

    def tc_image(self, repack=False):
        """bf16 operand images of the parameters for the tensor-core kernel (emotion_tf_tc.cu); packed on first use
        (call with repack=True after changing self.t in place)."""
        if self._tc_image is None or repack:
            if self._tc_image is None:
                self._tc_image = torch.empty(L.lib().erlfill_emotion_tc_image_bytes(), dtype=torch.uint8, device=self.device)
            w = self.struct()
            with torch.cuda.device(self.device):
                L.check(L.lib().erlfill_emotion_tc_pack(C.byref(w), L.ptr(self._tc_image), L.stream_ptr()),
                        "erlfill_emotion_tc_pack")
        return self._tc_image

    def state_dict(self):
        sd = {k: self.t[f].clone() for f, k in _TF_GLOBAL}
        for f, k in _TF_LAYER:
            for l in range(N_LAYERS):
                sd[f"transformer_encoder.layers.{l}.{k}"] = self.t[f][l].clone()
        return sd


class EmotionStateVec:
    """Per-env EmotionModule state: current_emotion (binary64) + 20-deep history ring (EmotionModule.py:105-120)."""

    def __init__(self, n, device="cuda"):
        dev = torch.device(device)
        self.n = n
        self.current = torch.zeros(3, n, dtype=torch.float64, device=dev)
        self.current[0].fill_(0.5)
        self.current[1].fill_(0.5)
        self.history = torch.zeros(L.HIST, 3, n, dtype=torch.float32, device=dev)
        self.hist_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.hist_head = torch.zeros(n, dtype=torch.int32, device=dev)

    def struct(self):
        return L.EmotionState(L.ptr(self.current), L.ptr(self.history), L.ptr(self.hist_len), L.ptr(self.hist_head))


def actor_forward(params, state, emotion, dropout=L.DROPOUT_OFF, masks=None, seed=0, call_idx=0, row_id_base=0,
                  ou=None, want_base=False):
    """A1 (+A2 with `ou`): returns scaled_action [n,10] (and base_action if want_base).
    ou: dict(state=[n,10] f32, sigma=[n] f32, randn=None|[n,10], final=[n,10] out, seed, env_id_base, step_idx)."""
    n = state.shape[0]
    assert state.dtype == torch.float32 and state.is_contiguous() and emotion.is_contiguous()
    bc = 1 if emotion.dim() == 1 or emotion.shape[0] != n else 0
    out = torch.empty(n, 10, dtype=torch.float32, device=state.device)
    base = torch.empty(n, 10, dtype=torch.float32, device=state.device) if want_base else None
    w = params.struct()
    m1, m2 = masks if masks is not None else (None, None)
    oup = None
    if ou is not None:
        ous = L.OU(L.ptr(ou["state"]), L.ptr(ou["sigma"]), L.ptr(ou.get("randn")), L.ptr(ou["final"]),
                   int(ou.get("seed", 0)), int(ou.get("env_id_base", 0)), int(ou.get("step_idx", 0)) & 0xFFFFFFFF)
        oup = C.byref(ous)
    with torch.cuda.device(state.device):
        L.check(L.lib().erlfill_actor_forward(C.byref(w), C.c_int(n), L.ptr(state), L.ptr(emotion), C.c_int(bc),
                                              C.c_int(dropout), C.c_uint64(seed), C.c_uint32(call_idx),
                                              C.c_uint32(row_id_base), L.ptr(m1), L.ptr(m2), L.ptr(out), L.ptr(base),
                                              oup, L.stream_ptr()), "erlfill_actor_forward")
    return (out, base) if want_base else out


def emotion_update(emo, reward, state, next_state):
    with torch.cuda.device(reward.device):
        es = emo.struct()
        L.check(L.lib().erlfill_emotion_update(C.c_int(emo.n), L.ptr(reward), L.ptr(state), L.ptr(next_state),
                                               C.byref(es), L.stream_ptr()), "erlfill_emotion_update")


def emotion_forward(params, seq=None, emo=None, dropout=L.DROPOUT_OFF, masks=None, seed=0, call_idx=0,
                    row_id_base=0, out=None, precision="fp32"):
    """M2: [n,3] emotion vectors from explicit sequences seq[n,S,3] or from the history rings in `emo`."""
    if seq is not None:
        n, S = seq.shape[0], seq.shape[1]
        assert seq.dtype == torch.float32 and seq.is_contiguous()
        dev = seq.device
    else:
        n, S, dev = emo.n, L.HIST, emo.history.device
    if out is None:
        out = torch.empty(n, 3, dtype=torch.float32, device=dev)
    es = emo.struct() if emo is not None else None
    if precision == "bf16":
        # tensor-core path (tcgen05, bf16 operands / fp32 accumulate); train()-mode dropout draws the same masks as the fp32 kernel
        img = params.tc_image()
        ms = None
        if masks is not None:
            ms = L.TransformerMasks(L.ptr(masks["attn"]), L.ptr(masks["d1"]), L.ptr(masks["ff"]), L.ptr(masks["d2"]))
        with torch.cuda.device(dev):
            L.check(L.lib().erlfill_emotion_forward_tc_train(L.ptr(img), C.c_int(n), L.ptr(seq), C.c_int(S),
                                                             C.byref(es) if es is not None else None, C.c_int(dropout),
                                                             C.c_uint64(seed), C.c_uint32(call_idx), C.c_uint32(row_id_base),
                                                             C.byref(ms) if ms is not None else None, L.ptr(out), L.stream_ptr()),
                    "erlfill_emotion_forward_tc_train")
        return out
    if precision != "fp32":
        raise L.ErlfillError(f"emotion_forward: unknown precision {precision!r}")
    w = params.struct()
    ms = None
    if masks is not None:
        ms = L.TransformerMasks(L.ptr(masks["attn"]), L.ptr(masks["d1"]), L.ptr(masks["ff"]), L.ptr(masks["d2"]))
    with torch.cuda.device(dev):
        L.check(L.lib().erlfill_emotion_forward(C.byref(w), C.c_int(n), L.ptr(seq), C.c_int(S),
                                                C.byref(es) if es is not None else None, C.c_int(dropout),
                                                C.c_uint64(seed), C.c_uint32(call_idx), C.c_uint32(row_id_base),
                                                C.byref(ms) if ms is not None else None, L.ptr(out), L.stream_ptr()),
                "erlfill_emotion_forward")
    return out


def episode_end(done, ou_state, sigma32, sigma64, episode_count=None):
    with torch.cuda.device(done.device):
        L.check(L.lib().erlfill_episode_end(C.c_int(done.shape[0]), L.ptr(done), L.ptr(ou_state), L.ptr(sigma32),
                                            L.ptr(sigma64), L.ptr(episode_count), L.stream_ptr()), "erlfill_episode_end")
