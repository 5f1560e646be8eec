"""Parameter initialisation that reproduces the reference's construction order, so the same
torch.manual_seed gives bit-identical initial weights (host-side, once; no compute path).

Reference order (ddpg_agent.py:295-302, then experiment_runner.py:484): Actor, Actor (target), Critic,
Critic (target), then EmotionTransformer.  Inside Actor: input_layer, the three hidden Linear layers,
emotion_weights = randn(3,10)*0.2 (ddpg_agent.py:31-56); Critic: ddpg_agent.py:140-161; EmotionTransformer:
embedding, positional randn(20,32), ONE TransformerEncoderLayer deep-copied 4 times (so the four layers start
identical), fc_out (EmotionModule.py:27-40).
"""
import torch
import torch.nn as nn


def _actor_sd(bounds):
    input_layer = nn.Linear(5, 128)
    hidden = nn.Sequential(nn.LeakyReLU(0.1), nn.Dropout(0.2), nn.Linear(128, 128), nn.ReLU(), nn.Dropout(0.2),
                           nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 10), nn.Softsign())
    emotion_weights = torch.randn(3, 10) * 0.2
    sd = {"emotion_weights": emotion_weights}
    sd.update({"input_layer." + k: v.detach().clone() for k, v in input_layer.state_dict().items()})
    sd.update({"hidden." + k: v.detach().clone() for k, v in hidden.state_dict().items()})
    scale = torch.tensor([(hi - lo) / 2.0 for lo, hi in bounds.values()], dtype=torch.float32)
    offset = torch.tensor([(hi + lo) / 2.0 for lo, hi in bounds.values()], dtype=torch.float32)
    sd["scale_params"], sd["offset_params"] = scale, offset
    return sd


def _critic_sd():
    net = nn.Sequential(nn.Linear(15, 256), nn.LayerNorm(256), nn.LeakyReLU(0.1), nn.Dropout(0.2),
                        nn.Linear(256, 256), nn.LayerNorm(256), nn.ReLU(), nn.Dropout(0.2),
                        nn.Linear(256, 128), nn.ReLU(), nn.Linear(128, 1))
    return {"net." + k: v.detach().clone() for k, v in net.state_dict().items()}


def _transformer_sd():
    import warnings
    embedding = nn.Linear(3, 32)
    pos = torch.randn(20, 32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        layer = nn.TransformerEncoderLayer(d_model=32, nhead=4, dropout=0.3, batch_first=False)
        enc = nn.TransformerEncoder(layer, num_layers=4)
    fc_out = nn.Linear(32, 3)
    sd = {"positional_encoding": pos}
    sd.update({"embedding." + k: v.detach().clone() for k, v in embedding.state_dict().items()})
    sd.update({"transformer_encoder." + k: v.detach().clone() for k, v in enc.state_dict().items()})
    sd.update({"fc_out." + k: v.detach().clone() for k, v in fc_out.state_dict().items()})
    return sd


def er_ddpg_state_dicts(bounds, seed=None, with_transformer=True):
    """(actor_sd, critic_sd, transformer_sd) in the reference's RNG consumption order."""
    if seed is not None:
        torch.manual_seed(seed)
    actor = _actor_sd(bounds)
    _actor_sd(bounds)            # actor_target is constructed (and consumes RNG) before load_state_dict
    critic = _critic_sd()
    _critic_sd()
    trans = _transformer_sd() if with_transformer else None
    return actor, critic, trans


def transformer_keys():
    """EmotionTransformer.state_dict() keys in torch's order (own parameter, then children in registration order)."""
    per_layer = ["self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                 "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias", "norm1.weight", "norm1.bias",
                 "norm2.weight", "norm2.bias"]
    keys = ["positional_encoding", "embedding.weight", "embedding.bias"]
    keys += [f"transformer_encoder.layers.{l}.{k}" for l in range(4) for k in per_layer]
    return keys + ["fc_out.weight", "fc_out.bias"]
