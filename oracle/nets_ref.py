"""Plain PyTorch fp32 restatement of the reference networks — TEST INFRASTRUCTURE ONLY.

Functional forms (weights passed as a dict keyed like the reference's state_dict, dropout masks
passed explicitly) of:
  A1  Actor.forward                (DifferentModules/ddpg_agent.py:73-122)
  U1  Critic.forward               (DifferentModules/ddpg_agent.py:163-183)
  M2  EmotionTransformer.forward   (EmotionModule.py:48-71) = embedding + learned positions +
      4 x nn.TransformerEncoderLayer(d=32, heads=4, ff=2048, post-norm, relu) + last token + fc_out
  U2-U5 DDPGAgent.learn            (DifferentModules/ddpg_agent.py:422-512) as one function on
      explicit tensors, using autograd on the functional forms above.

Parity status: pinned against tests/golden/nets_golden.npz, produced by oracle/gen_golden.py from
the unmodified reference modules both in eval() and in train() mode (dropout masks replayed from
torch's CPU generator); see tests/test_nets_oracle.py.

Dropout convention: a mask is the 0/1 keep tensor; the kept activations are scaled by 1/(1-p)
exactly like at::dropout (input * (mask / (1 - p))).  masks=None means eval mode.
"""
import math

import torch
import torch.nn.functional as F

ACTOR_P = 0.2      # ddpg_agent.py:30,40,43
CRITIC_P = 0.2     # ddpg_agent.py:139,150,155
TRANS_P = 0.3      # EmotionModule.py:9,35
N_LAYERS, D_MODEL, N_HEAD, D_FF, MAX_LEN = 4, 32, 4, 2048, 20


def _drop(x, mask, p):
    if mask is None:
        return x
    return x * (mask / (1.0 - p))


# ---------------------------------------------------------------------------------------------
# Emulation of the rounding points of the tcgen05 ("fast") update path (csrc/update_tc.cu): every GEMM operand is rounded to bf16,
# products are accumulated in fp32; everything else (LayerNorm, activations, losses, bias / LayerNorm gradients, optimiser) is fp32.
# It pins the kernel's LOGIC tightly; the distance of this emulation from the plain fp32 functions is the price of bf16 operands.
def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _Bf16Linear(torch.autograd.Function):
    """y = bf(x) bf(W)^T + b;  dX = bf(dY) bf(W);  dW = bf(dY)^T bf(x);  db = sum dY (fp32)."""

    @staticmethod
    def forward(ctx, x, W, b):
        xb, Wb = _bf(x), _bf(W)
        ctx.save_for_backward(xb, Wb)
        return xb @ Wb.t() + b

    @staticmethod
    def backward(ctx, dy):
        xb, Wb = ctx.saved_tensors
        dyb = _bf(dy)
        return dyb @ Wb, dyb.t() @ xb, dy.sum(0)


class _SplitLinear(torch.autograd.Function):
    """First layers: the forward is the hi/lo-split product (fp32-accurate to ~2^-16); the gradients use the bf16 hi parts only."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(_bf(x), _bf(W))
        return x @ W.t() + b

    @staticmethod
    def backward(ctx, dy):
        xb, Wb = ctx.saved_tensors
        dyb = _bf(dy)
        return dyb @ Wb, dyb.t() @ xb, dy.sum(0)


def _lin(x, W, b, emu, first=False):
    if not emu:
        return F.linear(x, W, b)
    return (_SplitLinear if first else _Bf16Linear).apply(x, W, b)


# ---------------------------------------------------------------------------------------------
def actor_forward(W, state, emotion, masks=None, return_base=False, emu=False):
    """masks: None or (m1[B,128], m2[B,128]) keep-masks of the two Dropout(0.2) layers.  emu: bf16-operand emulation (see _lin)."""
    if emotion.dim() == 1:
        emotion = emotion.unsqueeze(0)
    if state.dim() == 1:
        state = state.unsqueeze(0)
    if emotion.size(0) != state.size(0):
        emotion = emotion.expand(state.size(0), -1)
    m1, m2 = masks if masks is not None else (None, None)
    x = torch.cat([state, emotion], dim=1)
    x = _lin(x, W["input_layer.weight"], W["input_layer.bias"], emu, first=True)
    x = F.leaky_relu(x, 0.1)
    x = _drop(x, m1, ACTOR_P)
    x = F.relu(_lin(x, W["hidden.2.weight"], W["hidden.2.bias"], emu))
    x = _drop(x, m2, ACTOR_P)
    x = F.relu(_lin(x, W["hidden.5.weight"], W["hidden.5.bias"], emu))
    x = _lin(x, W["hidden.7.weight"], W["hidden.7.bias"], emu)
    base = F.softsign(x)
    curiosity, conserv, anxiety = emotion[:, 0:1], emotion[:, 1:2], emotion[:, 2:3]
    raw_effect = torch.tanh((emotion @ W["emotion_weights"]) * 1.5)
    alpha = 0.1 + 0.9 * anxiety
    alpha = alpha * (1.0 - 0.5 * conserv)
    alpha = alpha * (1.0 + 0.5 * curiosity)
    alpha = torch.clamp(alpha, min=0.01, max=1)
    modulator = (1 - torch.abs(base)) ** 2
    sign_flip = -torch.sign(base)
    raw_action = base + alpha * raw_effect * sign_flip * modulator
    out = raw_action * W["scale_params"] + W["offset_params"]
    return (out, base) if return_base else out


def critic_forward(W, state, action, emotion, masks=None, pre_normalized=None, emu=False):
    """masks: None or (m1[B,256], m2[B,256]).  Emotion is standardised with BATCH statistics
    (unbiased std), ddpg_agent.py:176 — B = 1 gives NaN; all-equal rows give (e-mean)/1e-6, i.e. torch's
    summation rounding noise amplified by 1e6 (0 in exact arithmetic).  pre_normalized: use this tensor as
    the standardised emotion instead (the CUDA path defines the all-equal case as exactly 0)."""
    m1, m2 = masks if masks is not None else (None, None)
    if pre_normalized is not None:
        emotion = pre_normalized
    else:
        emotion = (emotion - emotion.mean(dim=0, keepdim=True)) / (emotion.std(dim=0, keepdim=True) + 1e-6)
    x = torch.cat([state, action, emotion], dim=1)
    x = _lin(x, W["net.0.weight"], W["net.0.bias"], emu, first=True)
    x = F.layer_norm(x, (256,), W["net.1.weight"], W["net.1.bias"], 1e-5)
    x = _drop(F.leaky_relu(x, 0.1), m1, CRITIC_P)
    x = _lin(x, W["net.4.weight"], W["net.4.bias"], emu)
    x = F.layer_norm(x, (256,), W["net.5.weight"], W["net.5.bias"], 1e-5)
    x = _drop(F.relu(x), m2, CRITIC_P)
    x = F.relu(_lin(x, W["net.8.weight"], W["net.8.bias"], emu))
    q = F.linear(x, W["net.10.weight"], W["net.10.bias"])
    return torch.clamp(q, -1e6, 1e6)


def transformer_forward(W, seq, masks=None):
    """seq: [B, S, 3] (S <= 20).  masks: None or a list of 4 dicts with keep-masks
    {"attn": [B,4,S,S], "d1": [B,S,32], "ff": [B,S,2048], "d2": [B,S,32]}.  Returns [B, 3]."""
    B, S, _ = seq.shape
    x = F.linear(seq, W["embedding.weight"], W["embedding.bias"])            # [B,S,32]
    x = x + W["positional_encoding"][:S].unsqueeze(0)
    hd = D_MODEL // N_HEAD
    for l in range(N_LAYERS):
        pre = f"transformer_encoder.layers.{l}."
        m = masks[l] if masks is not None else {}
        qkv = F.linear(x, W[pre + "self_attn.in_proj_weight"], W[pre + "self_attn.in_proj_bias"])
        q, k, v = qkv.split(D_MODEL, dim=-1)
        q = q.view(B, S, N_HEAD, hd).transpose(1, 2)
        k = k.view(B, S, N_HEAD, hd).transpose(1, 2)
        v = v.view(B, S, N_HEAD, hd).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        att = _drop(att, m.get("attn"), TRANS_P)
        a = (att @ v).transpose(1, 2).reshape(B, S, D_MODEL)
        a = F.linear(a, W[pre + "self_attn.out_proj.weight"], W[pre + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + _drop(a, m.get("d1"), TRANS_P), (D_MODEL,), W[pre + "norm1.weight"],
                         W[pre + "norm1.bias"], 1e-5)
        h = F.relu(F.linear(x, W[pre + "linear1.weight"], W[pre + "linear1.bias"]))
        h = _drop(h, m.get("ff"), TRANS_P)
        h = F.linear(h, W[pre + "linear2.weight"], W[pre + "linear2.bias"])
        x = F.layer_norm(x + _drop(h, m.get("d2"), TRANS_P), (D_MODEL,), W[pre + "norm2.weight"],
                         W[pre + "norm2.bias"], 1e-5)
    out = F.linear(x[:, -1], W["fc_out.weight"], W["fc_out.bias"])
    return torch.tanh(out) * 0.5 + 0.5


def pad_history(history, max_len=MAX_LEN):
    """EmotionTransformer.get_state (EmotionModule.py:83-102): front-pad with the OLDEST entry; an
    empty history gives the single neutral token [0.5, 0.5, 0.0] (sequence length 1)."""
    if len(history) == 0:
        return torch.tensor([[[0.5, 0.5, 0.0]]], dtype=torch.float32)
    h = [torch.as_tensor(x, dtype=torch.float32) for x in history][-max_len:]
    padded = [h[0]] * (max_len - len(h)) + h
    return torch.stack(padded).unsqueeze(0)


# ---------------------------------------------------------------------------------------------
ACTOR_PARAM_ORDER = ["emotion_weights", "input_layer.weight", "input_layer.bias", "hidden.2.weight",
                     "hidden.2.bias", "hidden.5.weight", "hidden.5.bias", "hidden.7.weight", "hidden.7.bias"]
CRITIC_PARAM_ORDER = ["net.0.weight", "net.0.bias", "net.1.weight", "net.1.bias", "net.4.weight", "net.4.bias",
                      "net.5.weight", "net.5.bias", "net.8.weight", "net.8.bias", "net.10.weight", "net.10.bias"]


def _clip_grad_norm_(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (norm of per-tensor norms, coef clamped to 1, eps 1e-6)."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


def adam_step(params, grads, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad), single-tensor formulation."""
    out = []
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    for p, g, mi, vi in zip(params, grads, m, v):
        mi.mul_(b1).add_(g, alpha=1 - b1)
        vi.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (vi.sqrt() / math.sqrt(bc2)).add_(eps)
        out.append(p - (lr / bc1) * (mi / denom))
    return out


def ddpg_learn(actor, actor_t, critic, critic_t, opt, batch, next_emotion, train_step,
               masks=None, gamma=0.99, tau=0.01, actor_lr=1e-4, critic_lr=1e-3, lr_decay=0.999,
               exact_target_emotion=False, emu=False, debug=None):
    """One DDPGAgent.learn() (ddpg_agent.py:436-512) on explicit tensors.

    actor/actor_t/critic/critic_t: weight dicts (updated in place); opt: {"am","av","cm","cv","step"}
    Adam state; batch: dict(states[B,2], actions[B,10] normalised, rewards[B], next_states[B,2],
    dones[B], emotions[B,3]); next_emotion[3] = the fresh get_emotion() of :466.
    masks: None (eval) or dict with keys at_/ct_/c_/a_/c2_ -> mask pairs for the five forwards.
    emu: emulate the bf16 operand rounding of the tcgen05 update path (see _lin).
    Returns (critic_loss, actor_loss, lrs)."""
    mk = masks or {}
    s, a, r = batch["states"], batch["actions"], batch["rewards"].unsqueeze(1) * 0.01
    s2, d, e = batch["next_states"], batch["dones"].unsqueeze(1), batch["emotions"]
    anxiety, conserv, curiosity = e[:, 2].mean().item(), e[:, 1].mean().item(), e[:, 0].mean().item()
    actor_factor = 1.0 + 0.8 * curiosity - 0.4 * conserv + 0.2 * anxiety
    critic_factor = 1.0 - 0.3 * anxiety + 0.6 * conserv
    decay = lr_decay ** train_step
    c_lr = float(min(max(critic_lr * critic_factor * decay, critic_lr * 0.1), critic_lr * 3))
    a_lr = float(min(max(actor_lr * actor_factor * decay, actor_lr * 0.1), actor_lr * 3))
    with torch.no_grad():
        ne = next_emotion.view(1, 3).expand(s2.size(0), -1)
        a2 = actor_forward(actor_t, s2, ne, mk.get("at"), emu=emu)
        tq = critic_forward(critic_t, s2, a2, ne, mk.get("ct"),
                            pre_normalized=torch.zeros_like(ne) if exact_target_emotion else None, emu=emu)
        y = r + (1 - d) * gamma * tq
    cp = [critic[k].clone().requires_grad_(True) for k in CRITIC_PARAM_ORDER]
    cw = dict(zip(CRITIC_PARAM_ORDER, cp))
    q = critic_forward(cw, s, a, e, mk.get("c"), emu=emu)
    critic_loss = F.smooth_l1_loss(q, y)
    cg = torch.autograd.grad(critic_loss, cp)
    if debug is not None:
        debug["critic_grad"] = torch.cat([g.reshape(-1) for g in cg]).clone()
        debug["q"], debug["y"] = q.detach().clone(), y.clone()
    cg, _ = _clip_grad_norm_(list(cg), 5.0)
    opt["step"] += 1
    new_c = adam_step([p.detach() for p in cp], cg, opt["cm"], opt["cv"], opt["step"], c_lr)
    for k, p in zip(CRITIC_PARAM_ORDER, new_c):
        critic[k] = p
    ap = [actor[k].clone().requires_grad_(True) for k in ACTOR_PARAM_ORDER]
    aw = dict(zip(ACTOR_PARAM_ORDER, ap))
    aw["scale_params"], aw["offset_params"] = actor["scale_params"], actor["offset_params"]
    act_out = actor_forward(aw, s, e, mk.get("a"), emu=emu)
    actor_loss = -critic_forward(critic, s, act_out, e, mk.get("c2"), emu=emu).mean()
    actor_loss = actor_loss - 0.03 * torch.std(act_out, dim=1).mean()
    l2 = torch.tensor(0.0)
    for p in ap:
        l2 = l2 + torch.norm(p, p=2)
    actor_loss = actor_loss + 1e-4 * l2
    ag = torch.autograd.grad(actor_loss, ap)
    if debug is not None:
        debug["actor_grad"] = torch.cat([g.reshape(-1) for g in ag]).clone()
        debug["action_out"] = act_out.detach().clone()
    ag, _ = _clip_grad_norm_(list(ag), 5.0)
    new_a = adam_step([p.detach() for p in ap], ag, opt["am"], opt["av"], opt["step"], a_lr)
    for k, p in zip(ACTOR_PARAM_ORDER, new_a):
        actor[k] = p
    for k in ACTOR_PARAM_ORDER:
        actor_t[k] = tau * actor[k] + (1 - tau) * actor_t[k]
    for k in CRITIC_PARAM_ORDER:
        critic_t[k] = tau * critic[k] + (1 - tau) * critic_t[k]
    return critic_loss.item(), actor_loss.item(), (c_lr, a_lr)


# ---------------------------------------------------------------------------------------------
# T1 — TD3 (DifferentModules/td3_agent.py:51-303) and fixed-alpha SAC (DifferentModules/sac_agent.py:14-283)
# ---------------------------------------------------------------------------------------------
TD3_ACTOR_ORDER = ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "net.4.weight", "net.4.bias"]
TD3_CRITIC_ORDER = [f"{q}.{i}.{w}" for q in ("q1", "q2") for i in (0, 2, 4) for w in ("weight", "bias")]
SAC_POLICY_ORDER = ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "mean.weight", "mean.bias",
                    "log_std.weight", "log_std.bias"]
SAC_Q_ORDER = ["q.0.weight", "q.0.bias", "q.2.weight", "q.2.bias", "q.4.weight", "q.4.bias"]


def _mlp3(w, prefix, x, idx=(0, 2, 4)):
    h = F.relu(F.linear(x, w[f"{prefix}{idx[0]}.weight"], w[f"{prefix}{idx[0]}.bias"]))
    h = F.relu(F.linear(h, w[f"{prefix}{idx[1]}.weight"], w[f"{prefix}{idx[1]}.bias"]))
    return F.linear(h, w[f"{prefix}{idx[2]}.weight"], w[f"{prefix}{idx[2]}.bias"])


def td3_actor_forward(w, s):
    """TD3Actor.forward (td3_agent.py:51-85): tanh(MLP 2-256-256-10)."""
    return torch.tanh(_mlp3(w, "net.", s))


def td3_critic_forward(w, s, a):
    """TD3Critic.forward (td3_agent.py:88-133): twin Q on cat([s, a])."""
    sa = torch.cat([s, a], dim=1)
    return _mlp3(w, "q1.", sa), _mlp3(w, "q2.", sa)


def td3_train_step(actor, actor_t, critic, critic_t, opt, batch, noise, total_it, gamma=0.99, tau=0.005,
                   policy_noise=0.2, noise_clip=0.5, policy_delay=2, actor_lr=1e-4, critic_lr=1e-3):
    """One TD3Agent.train_step() (td3_agent.py:238-303) on explicit tensors.  noise: the N(0,1) draws of
    torch.randn_like(actions) (:268).  opt: {"am","av","cm","cv","cstep","astep"}.  total_it: value AFTER the
    increment of :254.  Returns (actor_loss | None, critic_loss, q1.mean())."""
    s, a, r = batch["states"], batch["actions"], batch["rewards"].unsqueeze(1)
    s2, d = batch["next_states"], batch["dones"].unsqueeze(1)
    with torch.no_grad():
        n = (noise * policy_noise).clamp(-noise_clip, noise_clip)
        a2 = (td3_actor_forward(actor_t, s2) + n).clamp(-1, 1)
        q1t, q2t = td3_critic_forward(critic_t, s2, a2)
        y = r + gamma * (1 - d) * torch.min(q1t, q2t)
    cp = [critic[k].clone().requires_grad_(True) for k in TD3_CRITIC_ORDER]
    cw = dict(zip(TD3_CRITIC_ORDER, cp))
    q1, q2 = td3_critic_forward(cw, s, a)
    critic_loss = F.mse_loss(q1, y) + F.mse_loss(q2, y)
    cg = torch.autograd.grad(critic_loss, cp)
    opt["cstep"] += 1
    for k, p in zip(TD3_CRITIC_ORDER, adam_step([p.detach() for p in cp], list(cg), opt["cm"], opt["cv"], opt["cstep"], critic_lr)):
        critic[k] = p
    actor_loss = None
    if total_it % policy_delay == 0:
        ap = [actor[k].clone().requires_grad_(True) for k in TD3_ACTOR_ORDER]
        aw = dict(zip(TD3_ACTOR_ORDER, ap))
        al = -_mlp3(critic, "q1.", torch.cat([s, td3_actor_forward(aw, s)], dim=1)).mean()
        ag = torch.autograd.grad(al, ap)
        opt["astep"] += 1
        for k, p in zip(TD3_ACTOR_ORDER, adam_step([p.detach() for p in ap], list(ag), opt["am"], opt["av"], opt["astep"], actor_lr)):
            actor[k] = p
        for k in TD3_CRITIC_ORDER:
            critic_t[k] = tau * critic[k] + (1 - tau) * critic_t[k]
        for k in TD3_ACTOR_ORDER:
            actor_t[k] = tau * actor[k] + (1 - tau) * actor_t[k]
        actor_loss = al.item()
    return actor_loss, critic_loss.item(), q1.mean().item()


def sac_policy_sample(w, s, eps):
    """GaussianPolicy.sample (sac_agent.py:45-84) with the N(0,1) draw of rsample() given explicitly."""
    h = F.relu(F.linear(s, w["net.0.weight"], w["net.0.bias"]))
    h = F.relu(F.linear(h, w["net.2.weight"], w["net.2.bias"]))
    mean = F.linear(h, w["mean.weight"], w["mean.bias"])
    std = F.linear(h, w["log_std.weight"], w["log_std.bias"]).clamp(-20, 2).exp()
    z = mean + std * eps
    action = torch.tanh(z)
    log_prob = (-((z - mean) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
                - torch.log(1 - action.pow(2) + 1e-7)).sum(dim=-1, keepdim=True)
    return action, log_prob


def sac_q_forward(w, s, a):
    return _mlp3(w, "q.", torch.cat([s, a], dim=1))


def sac_update(policy, q1, q2, q1_t, q2_t, opt, batch, eps_next, eps_new, gamma=0.99, tau=0.005, alpha=0.3, lr=3e-4):
    """One SACAgent.update() (sac_agent.py:221-283).  opt: {"pm","pv","q1m","q1v","q2m","q2v","step"}.
    Returns (q1_loss, q2_loss, policy_loss)."""
    s, a, r = batch["states"], batch["actions"], batch["rewards"].unsqueeze(1)
    s2, d = batch["next_states"], batch["dones"].unsqueeze(1)
    with torch.no_grad():
        a2, lp2 = sac_policy_sample(policy, s2, eps_next)
        y = r + (1 - d) * gamma * (torch.min(sac_q_forward(q1_t, s2, a2), sac_q_forward(q2_t, s2, a2)) - alpha * lp2)
    opt["step"] += 1
    losses = []
    for net, m, v in ((q1, opt["q1m"], opt["q1v"]), (q2, opt["q2m"], opt["q2v"])):
        ps = [net[k].clone().requires_grad_(True) for k in SAC_Q_ORDER]
        loss = F.mse_loss(sac_q_forward(dict(zip(SAC_Q_ORDER, ps)), s, a), y)
        gs = torch.autograd.grad(loss, ps)
        for k, p in zip(SAC_Q_ORDER, adam_step([p.detach() for p in ps], list(gs), m, v, opt["step"], lr)):
            net[k] = p
        losses.append(loss.item())
    pp = [policy[k].clone().requires_grad_(True) for k in SAC_POLICY_ORDER]
    an, lp = sac_policy_sample(dict(zip(SAC_POLICY_ORDER, pp)), s, eps_new)
    pl = (alpha * lp - sac_q_forward(q1, s, an)).mean()
    pg = torch.autograd.grad(pl, pp)
    for k, p in zip(SAC_POLICY_ORDER, adam_step([p.detach() for p in pp], list(pg), opt["pm"], opt["pv"], opt["step"], lr)):
        policy[k] = p
    for k in SAC_Q_ORDER:
        q1_t[k] = tau * q1[k] + (1 - tau) * q1_t[k]
        q2_t[k] = tau * q2[k] + (1 - tau) * q2_t[k]
    return losses[0], losses[1], pl.item()


# ---------------------------------------------------------------------------------------------
# EmotionSAC: DifferentModules/ERL_Fill_agent.py
# ---------------------------------------------------------------------------------------------
ESAC_POLICY_ORDER = ["state_encoder.0.weight", "state_encoder.0.bias", "state_encoder.1.weight", "state_encoder.1.bias",
                     "emotion_encoder.0.weight", "emotion_encoder.0.bias", "emotion_encoder.1.weight", "emotion_encoder.1.bias",
                     "policy_net.0.weight", "policy_net.0.bias", "policy_net.2.weight", "policy_net.2.bias",
                     "mean_layer.weight", "mean_layer.bias", "log_std_layer.weight", "log_std_layer.bias",
                     "emotion_adapter_mean.0.weight", "emotion_adapter_mean.0.bias",
                     "emotion_adapter_std.0.weight", "emotion_adapter_std.0.bias"]
ESAC_Q_ORDER = ["state_encoder.0.weight", "state_encoder.0.bias", "state_encoder.1.weight", "state_encoder.1.bias",
                "action_encoder.0.weight", "action_encoder.0.bias", "action_encoder.1.weight", "action_encoder.1.bias",
                "emotion_encoder.0.weight", "emotion_encoder.0.bias", "emotion_encoder.1.weight", "emotion_encoder.1.bias",
                "q_net.0.weight", "q_net.0.bias", "q_net.2.weight", "q_net.2.bias", "q_net.4.weight", "q_net.4.bias"]


def _enc(w, name, x):
    """Linear -> LayerNorm -> ReLU encoder (ERL_Fill_agent.py:78-89, 196-213)."""
    z = F.linear(x, w[name + ".0.weight"], w[name + ".0.bias"])
    return F.relu(F.layer_norm(z, (z.shape[-1],), w[name + ".1.weight"], w[name + ".1.bias"], 1e-5))


def esac_policy_forward(w, s, e):
    """EmotionGaussianPolicy.forward (ERL_Fill_agent.py:117-144): (mean, log_std clamped to [-20, 2])."""
    x = torch.cat([_enc(w, "state_encoder", s), _enc(w, "emotion_encoder", e)], dim=-1)
    x = F.relu(F.linear(x, w["policy_net.0.weight"], w["policy_net.0.bias"]))
    x = F.relu(F.linear(x, w["policy_net.2.weight"], w["policy_net.2.bias"]))
    mean = F.linear(x, w["mean_layer.weight"], w["mean_layer.bias"]) + \
        torch.tanh(F.linear(e, w["emotion_adapter_mean.0.weight"], w["emotion_adapter_mean.0.bias"]))
    log_std = F.linear(x, w["log_std_layer.weight"], w["log_std_layer.bias"]) + \
        torch.tanh(F.linear(e, w["emotion_adapter_std.0.weight"], w["emotion_adapter_std.0.bias"]))
    return mean, torch.clamp(log_std, -20, 2)


def esac_policy_sample(w, s, e, eps):
    """EmotionGaussianPolicy.sample (:146-167) with the rsample() draw injected: z = mean + std * eps."""
    mean, log_std = esac_policy_forward(w, s, e)
    std = torch.exp(log_std)
    z = mean + std * eps
    a = torch.tanh(z)
    log_prob = -((z - mean) ** 2) / (2 * std ** 2) - log_std - math.log(math.sqrt(2 * math.pi)) - torch.log(1 - a.pow(2) + 1e-6)
    return a, log_prob.sum(dim=-1, keepdim=True)


def esac_q_forward(w, s, a, e):
    """EmotionAwareQNetwork.forward (:222-239)."""
    x = torch.cat([_enc(w, "state_encoder", s), _enc(w, "action_encoder", a), _enc(w, "emotion_encoder", e)], dim=-1)
    x = F.relu(F.linear(x, w["q_net.0.weight"], w["q_net.0.bias"]))
    x = F.relu(F.linear(x, w["q_net.2.weight"], w["q_net.2.bias"]))
    return F.linear(x, w["q_net.4.weight"], w["q_net.4.bias"])


def esac_update(policy, q1, q2, q1_t, q2_t, opt, batch, eps_next, eps_new, e_now, e_prev, gamma=0.99, tau=0.005, alpha=0.3,
                lr=3e-4):
    """One EmotionSACAgent.update() (ERL_Fill_agent.py:365-448).  opt as in sac_update; batch additionally holds "emotions".
    The emotion regulariser ||e_now - e_prev||_2 (:417-430) is a constant of the parameters: reported, no gradient.
    Returns (q1_loss, q2_loss, policy_loss, emo_reg)."""
    s, a, r = batch["states"], batch["actions"], batch["rewards"].unsqueeze(1)
    s2, d, e = batch["next_states"], batch["dones"].unsqueeze(1), batch["emotions"]
    with torch.no_grad():
        a2, lp2 = esac_policy_sample(policy, s2, e, eps_next)
        y = r + (1 - d) * gamma * (torch.min(esac_q_forward(q1_t, s2, a2, e), esac_q_forward(q2_t, s2, a2, e)) - alpha * lp2)
    opt["step"] += 1
    losses = []
    for net, m, v in ((q1, opt["q1m"], opt["q1v"]), (q2, opt["q2m"], opt["q2v"])):
        ps = [net[k].clone().requires_grad_(True) for k in ESAC_Q_ORDER]
        loss = F.mse_loss(esac_q_forward(dict(zip(ESAC_Q_ORDER, ps)), s, a, e), y)
        gs = torch.autograd.grad(loss, ps)
        for k, p in zip(ESAC_Q_ORDER, adam_step([p.detach() for p in ps], list(gs), m, v, opt["step"], lr)):
            net[k] = p
        losses.append(loss.item())
    pp = [policy[k].clone().requires_grad_(True) for k in ESAC_POLICY_ORDER]
    an, lp = esac_policy_sample(dict(zip(ESAC_POLICY_ORDER, pp)), s, e, eps_new)
    pl = (alpha * lp - esac_q_forward(q1, s, an, e)).mean()
    pg = torch.autograd.grad(pl, pp)
    for k, p in zip(ESAC_POLICY_ORDER, adam_step([p.detach() for p in pp], list(pg), opt["pm"], opt["pv"], opt["step"], lr)):
        policy[k] = p
    for k in ESAC_Q_ORDER:
        q1_t[k] = tau * q1[k] + (1 - tau) * q1_t[k]
        q2_t[k] = tau * q2[k] + (1 - tau) * q2_t[k]
    emo_reg = float(torch.norm(torch.as_tensor(e_now, dtype=torch.float32) - torch.as_tensor(e_prev, dtype=torch.float32), p=2))
    return losses[0], losses[1], pl.item(), emo_reg
