"""Host side of EmotionSAC (csrc/esac.cu), the emotion-aware SAC agent of DifferentModules/ERL_Fill_agent.py (SURVEY.md section 8f
row N3): flat parameter buffers keyed like the reference state_dicts, the act / update calls and the batched loop body of
train_erl_fill (:526-541): act -> env.step -> emotion.update -> remember -> update.

Batched definitions (as for VecERDDPG): every env has its own EmotionModule state; with dropout off one
transformer call per step serves the reference's get_emotion() call sites of that step; the emotion regulariser compares the
mean emotion over the rank's envs at consecutive updates (N = 1: the reference's e_now / e_prev).  train_erl_fill feeds
`emotion.update(reward, action)` — the 10-vector ACTION as "movement" (:538), whose norm always clips to 1.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L
from . import init, nets
from .ddpg import FlatNet, ReplayBuffer

POLICY_LAYOUT = [("state_encoder.0.weight", (256, 2)), ("state_encoder.0.bias", (256,)), ("state_encoder.1.weight", (256,)),
                 ("state_encoder.1.bias", (256,)), ("emotion_encoder.0.weight", (32, 3)), ("emotion_encoder.0.bias", (32,)),
                 ("emotion_encoder.1.weight", (32,)), ("emotion_encoder.1.bias", (32,)), ("policy_net.0.weight", (256, 288)),
                 ("policy_net.0.bias", (256,)), ("policy_net.2.weight", (256, 256)), ("policy_net.2.bias", (256,)),
                 ("mean_layer.weight", (10, 256)), ("mean_layer.bias", (10,)), ("log_std_layer.weight", (10, 256)),
                 ("log_std_layer.bias", (10,)), ("emotion_adapter_mean.0.weight", (10, 3)), ("emotion_adapter_mean.0.bias", (10,)),
                 ("emotion_adapter_std.0.weight", (10, 3)), ("emotion_adapter_std.0.bias", (10,))]
_Q_ONE = [("state_encoder.0.weight", (256, 2)), ("state_encoder.0.bias", (256,)), ("state_encoder.1.weight", (256,)),
          ("state_encoder.1.bias", (256,)), ("action_encoder.0.weight", (256, 10)), ("action_encoder.0.bias", (256,)),
          ("action_encoder.1.weight", (256,)), ("action_encoder.1.bias", (256,)), ("emotion_encoder.0.weight", (32, 3)),
          ("emotion_encoder.0.bias", (32,)), ("emotion_encoder.1.weight", (32,)), ("emotion_encoder.1.bias", (32,)),
          ("q_net.0.weight", (256, 544)), ("q_net.0.bias", (256,)), ("q_net.2.weight", (256, 256)), ("q_net.2.bias", (256,)),
          ("q_net.4.weight", (1, 256)), ("q_net.4.bias", (1,))]
Q_KEYS = [k for k, _ in _Q_ONE]
Q_LAYOUT = [("q1." + k, s) for k, s in _Q_ONE] + [("q2." + k, s) for k, s in _Q_ONE]


def _encoder(n_in, n_out):
    return nn.Sequential(nn.Linear(n_in, n_out), nn.LayerNorm(n_out), nn.ReLU())


def _policy_sd():
    """EmotionGaussianPolicy.__init__'s construction order (ERL_Fill_agent.py:76-114)."""
    mods = {"state_encoder": _encoder(2, 256), "emotion_encoder": _encoder(3, 32),
            "policy_net": nn.Sequential(nn.Linear(288, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()),
            "mean_layer": nn.Linear(256, 10), "log_std_layer": nn.Linear(256, 10),
            "emotion_adapter_mean": nn.Sequential(nn.Linear(3, 10), nn.Tanh()), "emotion_adapter_std": nn.Sequential(nn.Linear(3, 10), nn.Tanh())}
    return {f"{name}.{k}": v.detach().clone() for name, m in mods.items() for k, v in m.state_dict().items()}


def _q_sd():
    """EmotionAwareQNetwork.__init__ (:200-220)."""
    mods = {"state_encoder": _encoder(2, 256), "action_encoder": _encoder(10, 256), "emotion_encoder": _encoder(3, 32),
            "q_net": nn.Sequential(nn.Linear(544, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 1))}
    return {f"{name}.{k}": v.detach().clone() for name, m in mods.items() for k, v in m.state_dict().items()}


def esac_state_dicts(seed=None, with_transformer=True):
    """(transformer_sd, policy_sd, q_sd) consuming torch's RNG in EmotionSACAgent.__init__'s order (:270-281): EmotionModule first,
    then policy, q1, q2, q1_target, q2_target.  q_sd holds q1.* and q2.*."""
    if seed is not None:
        torch.manual_seed(seed)
    trans = init._transformer_sd() if with_transformer else None
    pol = _policy_sd()
    q1, q2 = _q_sd(), _q_sd()
    _q_sd(); _q_sd()                                     # the target networks are constructed before load_state_dict
    q = {"q1." + k: v for k, v in q1.items()}
    q.update({"q2." + k: v for k, v in q2.items()})
    return trans, pol, q


class VecEmotionSAC:
    """EmotionSACAgent for N envs on one GPU."""

    def __init__(self, env, seed=0, batch_size=128, memory_size=100000, state_dicts=None, rank=0, gamma=0.99, tau=0.005, alpha=0.3,
                 lr=3e-4, lambda_emo=0.05, emotion_precision="fp32"):
        L.require_device()
        self.env, self.device, self.n = env, env.device, env.n_env
        self.seed, self.rank, self.batch_size = int(seed), int(rank), int(batch_size)
        trans_sd, pol_sd, q_sd = state_dicts if state_dicts is not None else esac_state_dicts(seed)
        d = self.device
        self.policy = FlatNet(POLICY_LAYOUT, L.ESAC_POLICY_PARAMS, d)
        self.q, self.q_target = FlatNet(Q_LAYOUT, 2 * L.ESAC_Q_PARAMS, d), FlatNet(Q_LAYOUT, 2 * L.ESAC_Q_PARAMS, d)
        self.policy.load(pol_sd); self.q.load(q_sd); self.q_target.load(q_sd)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=d)
        self.policy_m, self.policy_v = z(L.ESAC_POLICY_PARAMS), z(L.ESAC_POLICY_PARAMS)
        self.q_m, self.q_v = z(2 * L.ESAC_Q_PARAMS), z(2 * L.ESAC_Q_PARAMS)
        self.gamma, self.tau, self.alpha, self.lr, self.lambda_emo = gamma, tau, alpha, lr, lambda_emo
        self.adam_step = 0
        self.transformer = nets.TransformerParams(trans_sd, device=d) if trans_sd is not None else None
        self.emotion_precision = emotion_precision
        self.emo = nets.EmotionStateVec(self.n, device=d)
        self.emotion = torch.zeros(self.n, 3, dtype=torch.float32, device=d)
        self.memory = ReplayBuffer(memory_size, device=d)
        b = env.conditions[0]["bounds"]
        self.lo = torch.tensor([float(v[0]) for v in b.values()], dtype=torch.float32, device=d)
        self.hi = torch.tensor([float(v[1]) for v in b.values()], dtype=torch.float32, device=d)
        self.action = torch.zeros(self.n, 10, dtype=torch.float32, device=d)
        self.norm_action = torch.zeros(self.n, 10, dtype=torch.float32, device=d)
        self.state = torch.zeros(self.n, 2, dtype=torch.float32, device=d)
        self.idx = torch.zeros(batch_size, dtype=torch.int64, device=d)
        self.batch = torch.zeros(batch_size, L.REC, dtype=torch.float32, device=d)
        self.report = torch.zeros(4, dtype=torch.float32, device=d)
        self.e_now = torch.zeros(3, dtype=torch.float32, device=d)
        self.e_prev = None
        self._ws = torch.zeros((L.lib().erlfill_esac_workspace_bytes(C.c_int(batch_size)) + 3) // 4, dtype=torch.float32, device=d)
        self._act_ws = torch.zeros((L.lib().erlfill_esac_act_workspace_bytes() + 3) // 4, dtype=torch.float32, device=d)
        self._unit_move = torch.zeros(self.n, 2, dtype=torch.float32, device=d)
        self._unit_move[:, 0] = 1.0                       # ||action|| >= 1 always: movement factor clips to 1 (:538, EmotionModule.py:139-140)
        self._zero2 = torch.zeros(self.n, 2, dtype=torch.float32, device=d)
        self.t = 0
        self.updates = 0

    # -- acting -------------------------------------------------------------------------------------------------
    def _emotion_forward(self):
        if self.transformer is not None:
            nets.emotion_forward(self.transformer, emo=self.emo, out=self.emotion, precision=self.emotion_precision)

    def act(self, state, emotion=None, deterministic=False, eps=None):
        """Batched act: state [n,2], emotion [n,3] (default: the agent's per-env emotion) -> physical actions [n,10]."""
        n = state.shape[0]
        em = self.emotion if emotion is None else emotion
        bc = 1 if em.dim() == 1 or em.shape[0] != n else 0
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_esac_act(L.ptr(self.policy.flat), C.c_int(n), L.ptr(state), L.ptr(em), C.c_int(bc), L.ptr(eps),
                                             C.c_int(int(deterministic)), L.ptr(self.lo), L.ptr(self.hi), C.c_uint64(self.seed),
                                             C.c_uint32(self.env.env_id_base), C.c_uint32(self.t & 0xFFFFFFFF), L.ptr(self.action),
                                             L.ptr(self.norm_action), L.ptr(self._act_ws), L.stream_ptr()), "erlfill_esac_act")
        return self.action[:n]

    def act_batch(self, state, deterministic=False):
        return self.act(state, deterministic=deterministic)

    def reset(self):
        self.env.reset()
        self._emotion_forward()

    def rollout_step(self, store=True):
        """act -> env.step -> emotion.update(reward, action) -> remember for all N envs (train_erl_fill :535-540)."""
        env = self.env
        self.state.copy_(env.obs)
        a = self.act(env.obs)
        next_state, reward, done, _ = env.step(a)
        nets.emotion_update(self.emo, reward, self._zero2, self._unit_move)
        self.t += 1
        self._emotion_forward()
        if store:       # remember() stores get_emotion() AFTER the update (:362)
            self.memory.insert(self.state, self.norm_action, reward, next_state, done, self.emotion)
        return next_state, reward, done

    # -- learning -----------------------------------------------------------------------------------------------
    def grad_buckets(self):
        qg, pg = C.c_void_p(), C.c_void_p()
        L.check(L.lib().erlfill_esac_grad_buckets(L.ptr(self._ws), C.c_int(self.batch_size), C.byref(qg), C.byref(pg)), "erlfill_esac_grad_buckets")
        qo, po = (qg.value - self._ws.data_ptr()) // 4, (pg.value - self._ws.data_ptr()) // 4
        return self._ws[qo:qo + 2 * L.ESAC_Q_PARAMS], self._ws[po:po + L.ESAC_POLICY_PARAMS]

    def update(self, idx=None, eps_next=None, eps_new=None, e_now=None, allreduce=None):
        """One EmotionSACAgent.update(); returns the device report [q1_loss, q2_loss, policy_loss, emo_reg] (None until the buffer
        holds a batch).  idx / eps_* / e_now inject the reference's sampled indices, rsample() normals and emotion for parity runs."""
        if len(self.memory) < self.batch_size:
            return None
        self.adam_step += 1
        if idx is None:
            self.memory.sample_indices(self.batch_size, self.seed, self.updates, stream_id=self.rank, out=self.idx)
            idx = self.idx
        if idx.shape[0] > self.batch_size:      # the update workspace is sized for batch_size rows
            raise L.ErlfillError("minibatch of %d indices exceeds batch_size %d" % (idx.shape[0], self.batch_size))
        batch = self.memory.gather(idx, out=self.batch if idx.shape[0] == self.batch_size else None)
        if e_now is None:
            torch.mean(self.emotion, dim=0, out=self.e_now)
            e_now = self.e_now
        if self.e_prev is None:                            # first update: prev_emotion = e_now (:419-420)
            self.e_prev = e_now.clone()
        nets_ = L.SACNets(L.ptr(self.policy.flat), L.ptr(self.q.flat), L.ptr(self.q_target.flat), L.ptr(self.policy_m), L.ptr(self.policy_v),
                          L.ptr(self.q_m), L.ptr(self.q_v))
        hyper = L.SACHyper(batch.shape[0], self.adam_step, self.gamma, self.tau, self.alpha, 0.0, self.lr, self.seed,
                           self.updates & 0xFFFFFFFF, self.rank)

        def run(ph):
            with torch.cuda.device(self.device):
                L.check(L.lib().erlfill_esac_update(C.byref(nets_), C.byref(hyper), L.ptr(batch), L.ptr(eps_next), L.ptr(eps_new), L.ptr(e_now),
                                                    L.ptr(self.e_prev), L.ptr(self._ws), C.c_int(ph), L.ptr(self.report), L.stream_ptr()),
                        "erlfill_esac_update")
        if allreduce is None:
            run(L.PHASE_ALL)
        else:
            qb, pb = self.grad_buckets()
            run(L.PHASE_CRITIC_GRAD); allreduce(qb); run(L.PHASE_CRITIC_APPLY)
            run(L.PHASE_ACTOR_GRAD); allreduce(pb); run(L.PHASE_ACTOR_APPLY)
        self.e_prev.copy_(e_now)
        self.updates += 1
        return self.report
