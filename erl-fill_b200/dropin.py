"""N = 1 drop-ins for the reference's agent objects ("agent surface to preserve").

`DDPGAgent`, `EmotionModule`, `TD3Agent` and `SACAgent` present the attribute and method surface that
`train_ddpg` / `train_td3` / `train_sac` and `experiment_runner.py` touch (ddpg_agent.py:598-729,
td3_agent.py:371-470, sac_agent.py:308-400, experiment_runner.py:96,160-180,484) — numpy rows in, numpy rows /
Python floats out — while every piece of arithmetic on the path (transformer, actor, OU + clip, emotion update,
replay insert / gather, the whole `learn()`) runs in the sm_100a kernels behind include/erlfill.h at batch 1 / B.
Nothing here computes on the host: the host only moves the row, draws the reference's own host-side random
numbers (np.random.randn for OU / TD3 noise, np.random.choice / random.sample for minibatch indices — so a seeded
run consumes the global streams exactly like the reference and indices are bit-exact), and keeps the tuple list
`memory.buffer` that `env.reset()` samples from (MutiConditionEnv.py:145-154).

Declared differences:
  * dropout.  The reference never calls `.eval()`, so Dropout is live in `act`, in `remember`'s TD-error forwards
    and in `learn`.  `dropout="off"` (default) is the deterministic `.eval()` behaviour the parity fixtures pin;
    `dropout="philox"` turns it on for the act-time networks (transformer + actor) with counter-based masks.
    `learn()` always runs dropout off (erlfill_ddpg_update).
  * `use_td_error=True` evaluates the critic at batch 1, whose batch standardisation divides by an unbiased std of
    one row: the TD error is NaN in the reference and every priority collapses to 1e-4.
    The drop-in stores that NaN (with the reference's scalar type, which decides the overwrite slot of a full
    buffer) without running the three wasted forwards.
"""
import collections
import random

import numpy as np
import torch

from . import _lib as L
from . import ddpg, esac, init, nets, td3sac


# ---------------------------------------------------------------------------------------------------------
# state_dict facades
# ---------------------------------------------------------------------------------------------------------
def _check_keys(got, want, what):
    missing = [k for k in want if k not in got]
    unexpected = [k for k in got if k not in want]
    if missing or unexpected:      # torch's strict load_state_dict raises RuntimeError with both lists
        raise RuntimeError(f"Error(s) in loading state_dict for {what}: missing keys {missing}, unexpected keys {unexpected}")


class NetFacade:
    """What callers see as `agent.actor` / `agent.critic` / ...: a state_dict keyed and ordered like the reference
    module's, backed by views into the flat device buffer the kernels read."""

    def __init__(self, name, flatnet, order, buffers=None, on_load=None):
        self._name, self._net, self._order = name, flatnet, order
        self._buffers = buffers or {}
        self._on_load = on_load

    def state_dict(self):
        sd = collections.OrderedDict()
        for k in self._order:
            sd[k] = (self._buffers[k] if k in self._buffers else self._net.views[k]).detach().clone()
        return sd

    def load_state_dict(self, sd, strict=True):
        if strict:
            _check_keys(sd, self._order, self._name)
        for k in self._order:
            if k in sd:
                dst = self._buffers[k] if k in self._buffers else self._net.views[k]
                src = torch.as_tensor(sd[k])
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(src.shape)}, model {tuple(dst.shape)}")
                dst.copy_(src.to(dst.device, torch.float32))
        if self._on_load is not None:
            self._on_load()

    def parameters(self):
        return [self._net.views[k] for k, _ in self._net.layout]

    def eval(self):
        return self

    def train(self, mode=True):
        return self


ACTOR_SD_ORDER = ["emotion_weights", "scale_params", "offset_params", "input_layer.weight", "input_layer.bias",
                  "hidden.2.weight", "hidden.2.bias", "hidden.5.weight", "hidden.5.bias", "hidden.7.weight", "hidden.7.bias"]
CRITIC_SD_ORDER = [k for k, _ in ddpg.CRITIC_LAYOUT]


class ActorFacade(NetFacade):
    """Adds the Actor's registered buffers and `_init_param_scaling` (ddpg_agent.py:58-71; reset_actor_scaling :563-576)."""

    def __init__(self, name, flatnet, scale, offset):
        super().__init__(name, flatnet, ACTOR_SD_ORDER, buffers={"scale_params": scale, "offset_params": offset})
        self.scale_params, self.offset_params = scale, offset

    def _init_param_scaling(self, bounds):
        self.scale_params.copy_(torch.tensor([(hi - lo) / 2.0 for lo, hi in bounds.values()], dtype=torch.float32))
        self.offset_params.copy_(torch.tensor([(hi + lo) / 2.0 for lo, hi in bounds.values()], dtype=torch.float32))


class AdamFacade:
    """`agent.actor_optim` / `critic_optim`: `param_groups[0]['lr']` (read by train_ddpg, :655-656) and a
    torch.optim.Adam-compatible state_dict over the flat first/second-moment buffers."""

    def __init__(self, flatnet, m, v, lr, step_of):
        self._net, self._m, self._v, self._step_of = flatnet, m, v, step_of
        self.param_groups = [{"lr": lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                              "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                              "fused": None, "decoupled_weight_decay": False,
                              "params": list(range(len(flatnet.layout)))}]

    def state_dict(self):
        state, o, step = {}, 0, self._step_of()
        if step > 0:
            for i, (_, shape) in enumerate(self._net.layout):
                n = int(np.prod(shape))
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": self._m[o:o + n].view(*shape).clone(),
                            "exp_avg_sq": self._v[o:o + n].view(*shape).clone()}
                o += n
        return {"state": state, "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        """Returns the step count found in the checkpoint (0 if the optimiser had not stepped)."""
        o, step = 0, 0
        for i, (_, shape) in enumerate(self._net.layout):
            n = int(np.prod(shape))
            st = sd["state"].get(i)
            if st is None:
                self._m[o:o + n].zero_(); self._v[o:o + n].zero_()
            else:
                self._m[o:o + n].copy_(torch.as_tensor(st["exp_avg"]).reshape(-1).to(self._m.device, torch.float32))
                self._v[o:o + n].copy_(torch.as_tensor(st["exp_avg_sq"]).reshape(-1).to(self._v.device, torch.float32))
                step = max(step, int(float(st["step"])))
            o += n
        if sd.get("param_groups"):
            self.param_groups[0]["lr"] = sd["param_groups"][0].get("lr", self.param_groups[0]["lr"])
        return step


# ---------------------------------------------------------------------------------------------------------
# EmotionModule
# ---------------------------------------------------------------------------------------------------------
class TransformerFacade:
    """`emotion.transformer`: state_dict round trip (ddpg_agent.py:538-539,555-557) + the history length."""

    def __init__(self, owner):
        self._o = owner

    def state_dict(self):
        sd = self._o.params.state_dict()
        keys = init.transformer_keys()
        return collections.OrderedDict((k, sd[k]) for k in keys)

    def load_state_dict(self, sd, strict=True):
        keys = init.transformer_keys()
        if strict:
            _check_keys(sd, keys, "EmotionTransformer")
        merged = self._o.params.state_dict()
        merged.update({k: torch.as_tensor(v) for k, v in sd.items() if k in merged})
        self._o.params = nets.TransformerParams(merged, device=self._o.device)
        self._o._cache = None

    def eval(self):
        self._o.dropout = L.DROPOUT_OFF
        return self

    def train(self, mode=True):
        self._o.dropout = L.DROPOUT_PHILOX if mode else L.DROPOUT_OFF
        return self

    @property
    def history(self):
        """The emotion vectors currently in the 20-deep window, oldest first (EmotionModule.py:42,72-79)."""
        e = self._o.emo
        n, head = int(e.hist_len[0].item()), int(e.hist_head[0].item())
        h = e.history[:, :, 0].cpu()
        return [h[(head - n + i) % L.HIST].clone() for i in range(n)]


class EmotionModule:
    """EmotionModule.EmotionModule (EmotionModule.py:105-199) at N = 1 on the device kernels."""

    def __init__(self, device="cuda", state_dict=None, dropout="off", precision="fp32", seed=0):
        L.require_device()
        self.device = torch.device("cuda" if str(device) == "cpu" else device)   # reference callers pass their torch device
        if state_dict is None:
            state_dict = init._transformer_sd()                                    # consumes torch's global RNG like the reference ctor
        self.params = nets.TransformerParams(state_dict, device=self.device)
        self.emo = nets.EmotionStateVec(1, device=self.device)
        self.dropout = {"off": L.DROPOUT_OFF, "philox": L.DROPOUT_PHILOX}[dropout]
        self.precision = precision
        self.seed = int(seed)
        self.transformer = TransformerFacade(self)
        self._calls = 0
        self._cache = None
        self._cache_host = None
        self._out = torch.zeros(1, 3, dtype=torch.float32, device=self.device)
        self._r = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._zero2 = torch.zeros(1, 2, dtype=torch.float32, device=self.device)
        self._mv = torch.zeros(1, 2, dtype=torch.float32, device=self.device)

    @property
    def current_emotion(self):
        return self.emo.current[:, 0].cpu().numpy()

    @current_emotion.setter
    def current_emotion(self, value):       # DDPGAgent.load writes it (ddpg_agent.py:549)
        self.emo.current[:, 0] = torch.as_tensor(np.asarray(value, dtype=np.float64)).to(self.device)

    def update(self, reward, movement):
        """M1 (EmotionModule.py:121-162).  `movement` is the 2-vector s' - s."""
        self._r.fill_(float(reward))
        mv = np.asarray(movement, dtype=np.float32).reshape(-1)
        if mv.shape[0] != 2:
            # only ||movement|| enters (EmotionModule.py:139): a k-vector (train_erl_fill passes the 10-vector action, :538) goes
            # to the kernel as the 2-vector (||movement||, 0), capped at 2 where the clip to [0, 1] has long saturated
            mv = np.array([min(float(np.linalg.norm(np.asarray(movement, dtype=np.float64))), 2.0), 0.0], dtype=np.float32)
        self._mv.copy_(torch.from_numpy(mv).view(1, 2))
        nets.emotion_update(self.emo, self._r, self._zero2, self._mv)              # (0 + movement) - 0: exact
        self._cache = None
        self._cache_host = None
        return self.current_emotion.copy()

    def get_emotion_device(self):
        """[1,3] device tensor: one transformer call on the current history (cached while nothing changed and
        dropout is off — the reference's repeated calls differ only in their dropout draws)."""
        if self._cache is None or self.dropout != L.DROPOUT_OFF:
            if self.dropout != L.DROPOUT_OFF and self.precision != "fp32":
                raise L.ErlfillError("EmotionModule: dropout needs precision='fp32'")
            nets.emotion_forward(self.params, emo=self.emo, dropout=self.dropout, seed=self.seed, call_idx=self._calls,
                                 out=self._out, precision=self.precision)
            self._calls += 1
            self._cache = self._out
            self._cache_host = None
        return self._cache

    def get_emotion(self):
        dev = self.get_emotion_device()
        if self._cache_host is None or self.dropout != L.DROPOUT_OFF:
            self._cache_host = dev[0].cpu().numpy()          # one D2H per transformer call; repeated calls return copies
        return self._cache_host.copy()

    def save(self, path):
        torch.save(self.transformer.state_dict(), path)

    def load(self, path):
        self.transformer.load_state_dict(torch.load(path, map_location=self.device))


class EmotionModuleNone:
    """EmotionModule.EmotionModuleNone (EmotionModule.py:368-400): the fixed neutral emotion; DDPGAgent's default."""

    def __init__(self, device="cuda"):
        self.device = torch.device("cuda" if str(device) == "cpu" else device)
        self.current_emotion = np.array([0.5, 0.5, 0.0], dtype=np.float32)
        self._dev = torch.tensor([[0.5, 0.5, 0.0]], dtype=torch.float32, device=self.device)

    def update(self, reward, movement):
        return self.current_emotion

    def get_emotion(self):
        return self.current_emotion

    def get_emotion_device(self):
        return self._dev


# ---------------------------------------------------------------------------------------------------------
# ER-DDPG
# ---------------------------------------------------------------------------------------------------------
class OUNoise:
    """OUNoise surface (ddpg_agent.py:735-785).  The state lives on the device and is advanced inside the actor
    kernel (A2) from a host-drawn `np.random.randn(10).astype(float32)` — the reference's draw, same stream."""

    def __init__(self, size, device, mu=0.0, theta=0.15, sigma=0.6, min_sigma=0.1, decay=0.9995):
        self.size, self.theta, self.sigma, self.min_sigma, self.decay = size, theta, sigma, min_sigma, decay
        self.mu = mu * np.ones(size, dtype=np.float32)
        self._state = torch.zeros(1, size, dtype=torch.float32, device=device)
        self._sigma = torch.full((1,), sigma, dtype=torch.float32, device=device)

    @property
    def state(self):
        return self._state[0].cpu().numpy()

    def reset(self):
        self._state.zero_()

    def anneal(self):
        self.sigma = max(self.min_sigma, self.sigma * self.decay)
        self._sigma.fill_(self.sigma)


class PrioritizedReplay:
    """PrioritizedReplayBuffer surface (ddpg_agent.py:186-262): `buffer` (tuple list, element 0 = state),
    `priorities`, `add`, `sample`, `len`; every record is mirrored into the device ring the kernels gather from."""

    def __init__(self, capacity, device, scale, offset, alpha=0.6, epsilon=1e-4):
        self.capacity, self.alpha, self.epsilon = capacity, alpha, epsilon
        self.buffer, self.priorities = [], []
        self.device_ring = ddpg.ReplayBuffer(capacity, device=device)
        self._scale, self._offset, self._dev = scale, offset, device
        self._stage = None

    def __len__(self):
        return len(self.buffer)

    def _slot(self, p):
        if len(self.buffer) < self.capacity:
            return len(self.buffer)
        return int(np.argmin(self.priorities))

    def add(self, transition, priority, raw_action=None):
        """transition = (state, norm_action, reward, next_state, done, emotion, td_error)."""
        p = (abs(priority) + self.epsilon) ** self.alpha
        p = np.nan_to_num(p, nan=self.epsilon)
        slot = self._slot(p)
        if slot == len(self.buffer):
            self.buffer.append(transition); self.priorities.append(p)
        else:
            self.buffer[slot] = transition; self.priorities[slot] = p
        s, a, r, s2, d, e = transition[:6]
        dev = self._dev
        ring = self.device_ring
        # One pinned staging record (96 B: state | action | reward | next_state | emotion | done | ring position), ONE H2D
        # copy, the insert kernel reads the device copy field by field.  The kernel normalises a raw action with
        # (a - offset) / scale, clip +-1 (ddpg_agent.py:374-376); an already normalised one is written as is.
        if self._stage is None:
            self._stage_h = torch.zeros(96, dtype=torch.uint8).pin_memory()
            self._stage = torch.zeros(96, dtype=torch.uint8, device=dev)
            self._stage_ev = torch.cuda.Event()
            self._stage_busy = False
        if self._stage_busy:
            self._stage_ev.synchronize()            # the previous record's copy has left the pinned buffer (normally long ago)
        hf = self._stage_h[:80].view(torch.float32).numpy()
        hf[0:2] = np.asarray(s, dtype=np.float32).reshape(2)
        hf[2:12] = np.asarray(raw_action if raw_action is not None else a, dtype=np.float32).reshape(10)
        hf[12] = float(r)
        hf[14:16] = np.asarray(s2, dtype=np.float32).reshape(2)
        hf[16:19] = np.asarray(e, dtype=np.float32).reshape(3)
        self._stage_h[80] = 1 if d else 0
        self._stage_h[88:96].view(torch.int64)[0] = slot
        import ctypes as C
        with torch.cuda.device(dev):
            self._stage.copy_(self._stage_h, non_blocking=True)
            self._stage_ev.record()
            self._stage_busy = True
            base = self._stage.data_ptr()
            P = lambda off: C.c_void_p(base + off)
            sc, of = (self._scale, self._offset) if raw_action is not None else (None, None)
            L.check(L.lib().erlfill_replay_insert(C.c_int(1), L.ptr(ring.buf), C.c_int64(ring.capacity), C.c_int64(0), P(88),
                                                  P(0), P(8), P(48), P(56), P(80), P(64),
                                                  L.ptr(sc), L.ptr(of), L.stream_ptr()), "erlfill_replay_insert")
        ring.size = len(self.buffer)
        ring.inserted += 1

    def sample_indices(self, batch_size):
        """np.random.choice(n, B, p) with the reference's probabilities (ddpg_agent.py:238-251) — the same call on
        the same global stream, so the indices are the reference's."""
        pr = np.array(self.priorities, dtype=np.float32)
        total = pr.sum()
        if total <= 0 or np.isnan(total):
            pr += self.epsilon
            total = pr.sum()
        probs = np.nan_to_num(pr / total, nan=1.0 / len(pr))
        return np.random.choice(len(self.buffer), batch_size, p=probs)

    def sample(self, batch_size):
        return [self.buffer[i] for i in self.sample_indices(batch_size)]


class _LiveActor:
    """nets.actor_forward's parameter argument, pointing at the learner's live flat buffer."""

    def __init__(self, learner):
        self._l = learner

    def struct(self):
        return self._l.actor_struct()


class DDPGAgent:
    """DDPGAgent (ddpg_agent.py:265-560) at N = 1.  Construct after `torch.manual_seed(s)` for the reference's
    initial weights (actor, actor_target, critic, critic_target consume the global RNG in that order, :295-302)."""

    def __init__(self, env, device="cuda", use_td_error=False, dropout="off", state_dicts=None):
        L.require_device()
        self.env = env
        self.device = torch.device("cuda" if str(device) == "cpu" else device)
        self.use_td_error = use_td_error
        if state_dicts is None:
            a_sd, c_sd, _ = init.er_ddpg_state_dicts(env.bounds, seed=None, with_transformer=False)
        else:
            a_sd, c_sd = state_dicts[0], state_dicts[1]
        self.gamma, self.tau, self.batch_size, self.memory_size = 0.99, 0.01, 128, 100000
        self.actor_lr, self.critic_lr, self.train_step, self.lr_decay_rate = 1e-4, 1e-3, 1, 0.999
        self.target_update_freq, self.learn_step, self.actor_update_freq = 1, 0, 1
        self.learner = lr = ddpg.DDPGLearner(a_sd, c_sd, device=self.device, batch_size=self.batch_size, gamma=self.gamma,
                                             tau=self.tau, actor_lr=self.actor_lr, critic_lr=self.critic_lr,
                                             lr_decay_rate=self.lr_decay_rate)
        lr.use_graph = False
        self.actor = ActorFacade("Actor", lr.actor, lr.scale, lr.offset)
        self.actor_target = ActorFacade("Actor", lr.actor_target, lr.scale, lr.offset)
        self.critic = NetFacade("Critic", lr.critic, CRITIC_SD_ORDER)
        self.critic_target = NetFacade("Critic", lr.critic_target, CRITIC_SD_ORDER)
        for f in (self.actor, self.actor_target, self.critic, self.critic_target):      # fast path: rebuild the bf16 weight images after a load
            prev = getattr(f, "_on_load", None)
            f._on_load = (lambda prev=prev: (prev() if prev is not None else None, lr.params_changed())[1])
        self.actor_optim = AdamFacade(lr.actor, lr.actor_m, lr.actor_v, self.actor_lr, lambda: lr.adam_step)
        self.critic_optim = AdamFacade(lr.critic, lr.critic_m, lr.critic_v, self.critic_lr, lambda: lr.adam_step)
        self.emotion = EmotionModuleNone(self.device)
        self.memory = PrioritizedReplay(self.memory_size, self.device, lr.scale, lr.offset)
        self.ou_noise = OUNoise(env.action_dim, self.device)
        self.dropout = {"off": L.DROPOUT_OFF, "philox": L.DROPOUT_PHILOX}[dropout]
        self.seed = 0
        self._acts = 0
        self._randn = torch.zeros(1, 10, dtype=torch.float32, device=self.device)
        self._final = torch.zeros(1, 10, dtype=torch.float32, device=self.device)
        self._batch = None
        self._actor_params = _LiveActor(lr)

    # -- acting ---------------------------------------------------------------------------------------------
    def act(self, state, add_noise=True):
        """ddpg_agent.py:327-357: physical-unit action float32[10]."""
        emotion = self.emotion.get_emotion_device()
        s = torch.as_tensor(np.asarray(state, dtype=np.float32).reshape(1, 2)).to(self.device)
        ou = None
        if add_noise:
            self._randn.copy_(torch.from_numpy(np.random.randn(self.ou_noise.size).astype(np.float32)).view(1, -1))
            ou = dict(state=self.ou_noise._state, sigma=self.ou_noise._sigma, randn=self._randn, final=self._final)
        scaled = nets.actor_forward(self._actor_params, s, emotion, dropout=self.dropout, seed=self.seed,
                                    call_idx=self._acts, ou=ou)
        self._acts += 1
        return (self._final if add_noise else scaled)[0].cpu().numpy()

    def remember(self, state, action, reward, next_state, done):
        """ddpg_agent.py:359-395."""
        emotion_state = self.emotion.get_emotion()
        a32 = np.asarray(action, dtype=np.float32)
        ver = (self.learner.scale._version, self.learner.offset._version)
        if getattr(self, "_so_ver", None) != ver:                 # host copies of the scaling buffers, refreshed when they change
            self._so_host = (self.learner.scale.cpu().numpy(), self.learner.offset.cpu().numpy())
            self._so_ver = ver
        scale, offset = self._so_host
        norm_action = np.clip((np.array(action) - offset) / scale, -1.0, 1.0)      # what the tuple list shows; the device row is
        td_error = 0                                                              # normalised by the insert kernel from a32
        if self.use_td_error:
            # :378-390 with the batch-1 critic outputs replaced by the NaN they always are.  The host arithmetic is kept so
            # the NaN has the scalar TYPE the reference's has (np.float32 when the env handed out an np.float32 reward,
            # else float64): the type survives into the priority, and float32(1e-4) < float64(1e-4) is what np.argmin
            # sees once the buffer is full.
            nan = float("nan")
            td_error = abs((reward + (1 - int(done)) * self.gamma * nan) - nan)
        self.memory.add((state, norm_action, reward, next_state, done, emotion_state, td_error), priority=td_error, raw_action=a32)

    # -- learning -------------------------------------------------------------------------------------------
    def learn(self):
        """ddpg_agent.py:422-512: {"critic_loss", "actor_loss"} or None while the buffer is short."""
        B = self.batch_size
        if len(self.memory) < B:
            return None
        if self.use_td_error:
            idx = self.memory.sample_indices(B)
        else:
            idx = np.asarray(random.sample(range(len(self.memory)), B))           # random.sample(buffer, B) picks these positions
        if self._batch is None or self._batch.shape[0] != B:
            self._batch = torch.zeros(B, L.REC, dtype=torch.float32, device=self.device)
        self.last_indices = idx
        d_idx = torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(self.device)
        self.memory.device_ring.gather(d_idx, out=self._batch)
        next_emotion = self.emotion.get_emotion_device()[0].contiguous()
        lr = self.learner
        lr.actor_lr, lr.critic_lr, lr.gamma, lr.tau, lr.lr_decay_rate = self.actor_lr, self.critic_lr, self.gamma, self.tau, self.lr_decay_rate
        rep = lr.update(self._batch, next_emotion, self.train_step).cpu().numpy()
        self.learn_step += 1
        self.critic_optim.param_groups[0]["lr"] = float(rep[2])
        self.actor_optim.param_groups[0]["lr"] = float(rep[3])
        return {"critic_loss": float(rep[0]), "actor_loss": float(rep[1])}

    # -- checkpoints (N2) -----------------------------------------------------------------------------------
    def get_weights(self):
        return {"actor": self.actor.state_dict(), "critic": self.critic.state_dict()}

    def load_weights(self, weights):
        self.actor.load_state_dict(weights["actor"])
        self.critic.load_state_dict(weights["critic"])

    def save(self, path):
        """Same keys as the reference checkpoint (ddpg_agent.py:526-542)."""
        d = {"actor": self.actor.state_dict(), "critic": self.critic.state_dict(),
             "actor_target": self.actor_target.state_dict(), "critic_target": self.critic_target.state_dict(),
             "actor_optim": self.actor_optim.state_dict(), "critic_optim": self.critic_optim.state_dict(),
             "train_step": self.train_step}
        if hasattr(self.emotion, "transformer"):
            d["emotion_transformer"] = self.emotion.transformer.state_dict()
        torch.save(d, path)

    def load(self, filename, full=False):
        """ddpg_agent.py:544-560 restores actor, critic, the optional `emotion` vector and the transformer.
        full=True additionally restores the targets, Adam moments and train_step the file also holds (resume)."""
        ck = torch.load(filename, map_location=self.device, weights_only=False)
        self.actor.load_state_dict(ck["actor"])
        self.critic.load_state_dict(ck["critic"])
        if "emotion" in ck:
            self.emotion.current_emotion = ck["emotion"]
        if "emotion_transformer" in ck and hasattr(self.emotion, "transformer"):
            self.emotion.transformer.load_state_dict(ck["emotion_transformer"])
        if full:
            self.actor_target.load_state_dict(ck["actor_target"])
            self.critic_target.load_state_dict(ck["critic_target"])
            s1 = self.actor_optim.load_state_dict(ck["actor_optim"])
            s2 = self.critic_optim.load_state_dict(ck["critic_optim"])
            self.learner.adam_step = max(s1, s2)
            self.train_step = int(ck.get("train_step", self.train_step))


def reset_actor_scaling(agent, new_bounds):
    """ddpg_agent.py:563-576."""
    agent.actor._init_param_scaling(new_bounds)
    agent.actor_target._init_param_scaling(new_bounds)


# ---------------------------------------------------------------------------------------------------------
# TD3 / SAC
# ---------------------------------------------------------------------------------------------------------
class _EnvView:
    """What the batched agents need from an env, taken from a scalar drop-in env."""

    def __init__(self, env, device):
        self.conditions = [{"bounds": env.bounds}]
        self.device, self.n_env, self.env_id_base = device, 1, 0


class _DequeMemory:
    def __init__(self, ring, maxlen=100000):
        self.buffer = collections.deque(maxlen=maxlen)
        self._ring = ring

    def __len__(self):
        return len(self.buffer)

    def append(self, *args):
        self.buffer.append(args)


class _T1Scalar:
    def _setup(self, env, device):
        L.require_device()
        self.env = env
        self.device = torch.device("cuda" if str(device) == "cpu" else device)
        self.state_dim, self.action_dim, self.bounds = env.state_dim, env.action_dim, env.bounds
        self.batch_size = 128
        self._view = _EnvView(env, self.device)

    def _dev(self, x, n, dtype=torch.float32):
        return torch.as_tensor(np.asarray(x, dtype=np.float32).reshape(1, n)).to(self.device, dtype)

    def _store(self, state, norm_action, reward, next_state, done):
        """Appends to the reference-shaped deque and mirrors the row into the device ring (FIFO like deque(maxlen))."""
        self.memory.buffer.append((state, norm_action, reward, next_state, done))
        v = self._vec
        v.memory.insert(self._dev(state, 2), self._dev(norm_action, 10), torch.tensor([float(reward)], dtype=torch.float32, device=self.device),
                        self._dev(next_state, 2), torch.tensor([1 if done else 0], dtype=torch.uint8, device=self.device), v._zero3[:1])

    def _sync_batch(self):
        """`agent.batch_size` is a plain attribute in the reference; follow a caller that changes it."""
        v = self._vec
        if v.batch_size != self.batch_size:
            import ctypes as C
            v.batch_size = int(self.batch_size)
            v.idx = torch.zeros(v.batch_size, dtype=torch.int64, device=self.device)
            v.batch = torch.zeros(v.batch_size, L.REC, dtype=torch.float32, device=self.device)
            v._ws = torch.zeros((L.lib().erlfill_t1_workspace_bytes(C.c_int(v.batch_size)) + 3) // 4, dtype=torch.float32,
                                device=self.device)

    def _sample(self):
        """random.sample(buffer, B) positions -> device ring slots (the ring wraps where the deque drops its head)."""
        n = len(self.memory.buffer)
        pos = np.asarray(random.sample(range(n), self.batch_size), dtype=np.int64)
        ring = self._vec.memory
        first = (ring.inserted - n) % ring.capacity
        self.last_indices = pos
        return torch.as_tensor((first + pos) % ring.capacity).to(self.device)


class TD3Agent(_T1Scalar):
    """TD3Agent (td3_agent.py:137-347) at N = 1."""

    def __init__(self, env, device="cuda", state_dicts=None):
        self._setup(env, device)
        self._vec = td3sac.VecTD3(self._view, batch_size=self.batch_size, state_dicts=state_dicts or td3sac.td3_state_dicts(None))
        v = self._vec
        self.actor = NetFacade("TD3Actor", v.actor, [k for k, _ in td3sac.TD3_ACTOR_LAYOUT])
        self.actor_target = NetFacade("TD3Actor", v.actor_target, [k for k, _ in td3sac.TD3_ACTOR_LAYOUT])
        self.critic = NetFacade("TD3Critic", v.critic, [k for k, _ in td3sac.TD3_CRITIC_LAYOUT])
        self.critic_target = NetFacade("TD3Critic", v.critic_target, [k for k, _ in td3sac.TD3_CRITIC_LAYOUT])
        self.actor_optim = AdamFacade(v.actor, v.actor_m, v.actor_v, v.actor_lr, lambda: v.actor_steps)
        self.critic_optim = AdamFacade(v.critic, v.critic_m, v.critic_v, v.critic_lr, lambda: v.total_it)
        self.memory = _DequeMemory(v.memory)
        self.gamma, self.tau, self.policy_noise, self.noise_clip, self.policy_delay = v.gamma, v.tau, v.policy_noise, v.noise_clip, v.policy_delay

    @property
    def total_it(self):
        return self._vec.total_it

    def normalize_action(self, action):
        return np.array([2 * (action[i] - lo) / (hi - lo) - 1 for i, (lo, hi) in enumerate(self.bounds.values())])

    def denormalize_action(self, norm_action):
        return np.array([0.5 * (norm_action[i] + 1) * (hi - lo) + lo for i, (lo, hi) in enumerate(self.bounds.values())])

    def remember(self, state, action, reward, next_state, done):
        self._store(state, self.normalize_action(action), reward, next_state, done)

    def select_action(self, state, noise_scale=0.1):
        """td3_agent.py:221-236; the Gaussian is the reference's np.random.randn(10) draw."""
        randn = self._dev(np.random.randn(self.action_dim), 10)
        return self._vec.select_action(self._dev(state, 2), noise_scale, randn=randn)[0].cpu().numpy().astype(np.float64)

    def act(self, state, add_noise=False):
        return self.select_action(state, noise_scale=0.1 if add_noise else 0.0)

    def train_step(self, noise=None):
        """td3_agent.py:238-303 -> (actor_loss | None, critic_loss, mean q1); target noise from torch.randn (the
        reference's torch.randn_like on a CPU agent) unless given."""
        if len(self.memory) < self.batch_size:
            return None, None, None
        v = self._vec
        v.gamma, v.tau, v.policy_noise, v.noise_clip, v.policy_delay = self.gamma, self.tau, self.policy_noise, self.noise_clip, self.policy_delay
        self._sync_batch()
        idx = self._sample()
        if noise is None:
            noise = torch.randn(self.batch_size, self.action_dim)
        rep, upd = v.train_step(idx=idx, noise=torch.as_tensor(noise, dtype=torch.float32).to(self.device).contiguous())
        r = rep.cpu().numpy()
        return (float(r[2]) if upd else None), float(r[0]), float(r[1])

    def get_weights(self):
        return {"actor": self.actor.state_dict(), "critic": self.critic.state_dict()}

    def load_weights(self, weights):
        self.actor.load_state_dict(weights["actor"])
        self.critic.load_state_dict(weights["critic"])

    def save(self, path):
        torch.save(self.get_weights(), path)

    def load(self, path):
        self.load_weights(torch.load(path, map_location=self.device, weights_only=False))


class _SACQFacade:
    """`agent.q1` / `q2` / `q1_target` / `q2_target`: the reference holds four QNetwork modules whose keys are `q.N.*`;
    here they are the `q1.` / `q2.` halves of one flat buffer."""

    def __init__(self, flatnet, prefix):
        self._net, self._p = flatnet, prefix
        self._keys = [k[len(prefix):] for k, _ in flatnet.layout if k.startswith(prefix)]

    def state_dict(self):
        return collections.OrderedDict((k, self._net.views[self._p + k].detach().clone()) for k in self._keys)

    def load_state_dict(self, sd, strict=True):
        if strict:
            _check_keys(sd, self._keys, "QNetwork")
        for k in self._keys:
            if k in sd:
                self._net.views[self._p + k].copy_(torch.as_tensor(sd[k]).to(self._net.flat.device, torch.float32))


class SACAgent(_T1Scalar):
    """SACAgent (sac_agent.py:115-305) at N = 1 (fixed alpha = 0.3)."""

    def __init__(self, env, device="cuda", state_dicts=None):
        self._setup(env, device)
        self._vec = td3sac.VecSAC(self._view, batch_size=self.batch_size, state_dicts=state_dicts or td3sac.sac_state_dicts(None))
        v = self._vec
        self.policy = NetFacade("GaussianPolicy", v.policy, [k for k, _ in td3sac.SAC_POLICY_LAYOUT])
        self.q1, self.q2 = _SACQFacade(v.q, "q1."), _SACQFacade(v.q, "q2.")
        self.q1_target, self.q2_target = _SACQFacade(v.q_target, "q1."), _SACQFacade(v.q_target, "q2.")
        self.gamma, self.tau, self.alpha = v.gamma, v.tau, v.alpha
        self.offsets = np.array([(hi + lo) / 2 for lo, hi in env.bounds.values()])
        self.scales = np.array([(hi - lo) / 2 for lo, hi in env.bounds.values()])
        self.memory = _DequeMemory(v.memory)
        self.emotion_module = EmotionModuleNone(self.device)

    def denormalize_action(self, norm_action):
        return norm_action * self.scales + self.offsets

    def normalize_action(self, real_action):
        return (real_action - self.offsets) / self.scales

    def act(self, state, deterministic=False, eps=None):
        """sac_agent.py:186-205; the reparameterisation draw is torch.randn(1, 10) (Normal.rsample on a CPU agent)."""
        if eps is None and not deterministic:
            eps = torch.randn(1, self.action_dim)
        d_eps = None if deterministic else torch.as_tensor(eps, dtype=torch.float32).reshape(1, -1).to(self.device).contiguous()
        return self._vec.act(self._dev(state, 2), deterministic, eps=d_eps)[0].cpu().numpy().astype(np.float64)

    def remember(self, s, a, r, s_, d):
        self._store(s, self.normalize_action(np.asarray(a)), r, s_, d)

    def update(self, eps_next=None, eps_new=None):
        """sac_agent.py:221-283 -> {"q1_loss", "q2_loss", "policy_loss"} or None."""
        if len(self.memory.buffer) < self.batch_size:
            return None
        v = self._vec
        v.gamma, v.tau, v.alpha = self.gamma, self.tau, self.alpha
        self._sync_batch()
        idx = self._sample()
        B, A = self.batch_size, self.action_dim
        e1 = torch.randn(B, A) if eps_next is None else torch.as_tensor(eps_next, dtype=torch.float32)
        e2 = torch.randn(B, A) if eps_new is None else torch.as_tensor(eps_new, dtype=torch.float32)
        r = v.update(idx=idx, eps_next=e1.to(self.device).contiguous(), eps_new=e2.to(self.device).contiguous()).cpu().numpy()
        return {"q1_loss": float(r[0]), "q2_loss": float(r[1]), "policy_loss": float(r[2])}

    def save(self, path):
        torch.save({"policy": self.policy.state_dict()}, path)

    def load(self, path):
        self.policy.load_state_dict(torch.load(path, map_location=self.device, weights_only=False)["policy"])


# ---------------------------------------------------------------------------------------------------------
# EmotionSAC
# ---------------------------------------------------------------------------------------------------------
class _EmotionReplay:
    """EmotionReplayBuffer surface (ERL_Fill_agent.py:17-60): deque(maxlen) of (s, a_norm, r, s', d, e) + add / sample / len."""

    def __init__(self, capacity=100000):
        self.buffer = collections.deque(maxlen=capacity)

    def add(self, transition):
        self.buffer.append(transition)

    def sample(self, batch_size):
        return random.sample(self.buffer, batch_size)

    def __len__(self):
        return len(self.buffer)


class EmotionSACAgent(_T1Scalar):
    """EmotionSACAgent (ERL_Fill_agent.py:242-489) at N = 1.  Construct after `torch.manual_seed(s)` for the reference's initial
    weights (EmotionModule, policy, q1, q2 and the two targets consume the global RNG in that order, :270-281)."""

    def __init__(self, env, lambda_emo=0.05, device="cuda", state_dicts=None):
        self._setup(env, device)
        self.lambda_emo = lambda_emo
        self.train_step = 1
        if state_dicts is None:
            self.emotion = EmotionModule(device=self.device)
            _, pol, q = esac.esac_state_dicts(None, with_transformer=False)
        else:
            trans, pol, q = state_dicts
            self.emotion = EmotionModule(device=self.device, state_dict=trans)
        self._vec = esac.VecEmotionSAC(self._view, batch_size=self.batch_size, state_dicts=(None, pol, q), lambda_emo=lambda_emo)
        v = self._vec
        self.policy = NetFacade("EmotionGaussianPolicy", v.policy, [k for k, _ in esac.POLICY_LAYOUT])
        self.q1, self.q2 = _SACQFacade(v.q, "q1."), _SACQFacade(v.q, "q2.")
        self.q1_target, self.q2_target = _SACQFacade(v.q_target, "q1."), _SACQFacade(v.q_target, "q2.")
        self.gamma, self.tau, self.alpha = v.gamma, v.tau, v.alpha
        self.offsets = np.array([(hi + lo) / 2 for lo, hi in env.bounds.values()])
        self.scales = np.array([(hi - lo) / 2 for lo, hi in env.bounds.values()])
        self.memory = _EmotionReplay()

    def normalize_action(self, action):
        return (action - self.offsets) / self.scales

    def denormalize_action(self, norm_action):
        return norm_action * self.scales + self.offsets

    def act(self, state, deterministic=False, add_noise=True, eps=None):
        """:322-348.  The reparameterisation draw is torch.randn(1, 10) (Normal.rsample on a CPU agent) unless given."""
        det = deterministic or not add_noise
        if eps is None and not det:
            eps = torch.randn(1, self.action_dim)
        d_eps = None if det else torch.as_tensor(eps, dtype=torch.float32).reshape(1, -1).to(self.device).contiguous()
        self._vec.act(self._dev(state, 2), emotion=self.emotion.get_emotion_device(), deterministic=det, eps=d_eps)
        return self.denormalize_action(self._vec.norm_action[0].cpu().numpy())

    def remember(self, s, a, r, s_, d):
        """:350-363: the stored emotion is get_emotion() at remember time."""
        norm_a = self.normalize_action(np.asarray(a))
        e = self.emotion.get_emotion()
        self.memory.add((s, norm_a, r, s_, d, e))
        v = self._vec
        v.memory.insert(self._dev(s, 2), self._dev(norm_a, 10), torch.tensor([float(r)], dtype=torch.float32, device=self.device),
                        self._dev(s_, 2), torch.tensor([1 if d else 0], dtype=torch.uint8, device=self.device), self._dev(e, 3))

    def update(self, eps_next=None, eps_new=None):
        """:365-448 -> {"q1_loss", "q2_loss", "policy_loss", "emo_reg"} or None."""
        if len(self.memory) < self.batch_size:
            return None
        v = self._vec
        v.gamma, v.tau, v.alpha = self.gamma, self.tau, self.alpha
        if v.batch_size != self.batch_size:
            import ctypes as C
            v.batch_size = int(self.batch_size)
            v.batch = torch.zeros(v.batch_size, L.REC, dtype=torch.float32, device=self.device)
            v._ws = torch.zeros((L.lib().erlfill_esac_workspace_bytes(C.c_int(v.batch_size)) + 3) // 4, dtype=torch.float32, device=self.device)
        idx = self._sample()
        B, A = self.batch_size, self.action_dim
        e1 = torch.randn(B, A) if eps_next is None else torch.as_tensor(eps_next, dtype=torch.float32)
        e2 = torch.randn(B, A) if eps_new is None else torch.as_tensor(eps_new, dtype=torch.float32)
        e_now = self.emotion.get_emotion_device()[0].contiguous().clone()
        r = v.update(idx=idx, eps_next=e1.to(self.device).contiguous(), eps_new=e2.to(self.device).contiguous(), e_now=e_now).cpu().numpy()
        return {"q1_loss": float(r[0]), "q2_loss": float(r[1]), "policy_loss": float(r[2]), "emo_reg": float(r[3])}

    def save(self, path):
        """:450-465: the policy, and the emotion transformer next to it as *_emo.pth."""
        import os
        if os.path.dirname(path):
            os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"policy": self.policy.state_dict()}, path)
        if hasattr(self.emotion, "transformer"):
            self.emotion.save(path.replace(".pth", "_emo.pth"))

    def load(self, path):
        import os
        self.policy.load_state_dict(torch.load(path, map_location=self.device, weights_only=False)["policy"])
        emo_path = path.replace(".pth", "_emo.pth")
        if hasattr(self.emotion, "transformer") and os.path.exists(emo_path):
            self.emotion.load(emo_path)
