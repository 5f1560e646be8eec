"""Host side of the TD3 and SAC agents (csrc/td3_sac.cu): flat parameter buffers keyed like the reference's
state_dicts, the update calls, and the batched rollout loops of train_td3 / train_sac.

Reference: DifferentModules/td3_agent.py (TD3Agent :137-347, train_td3 :371-470) and DifferentModules/sac_agent.py
(SACAgent :115-305, train_sac :308-400).  Method names follow the reference: TD3 `select_action / act / remember /
train_step`, SAC `act / remember / update`; they take and return batched device tensors ([N, ...]) instead of one
numpy row.  Replay is the FIFO ring both agents use (deque(maxlen=100000): td3_agent.py:29, sac_agent.py:155) in the
80-byte record layout of replay.cu; sampling is uniform — the reference's `random.sample` draws WITHOUT replacement,
the batched path draws with replacement from the rank's shard (collision probability B/n), and
`train_step(idx=...)` / `update(idx=...)` accept the reference's own index list for parity runs.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L
from .ddpg import FlatNet, ReplayBuffer


def _mlp_layout(prefix, idx, in_dim, out_dim):
    a, b, c = idx
    return [(f"{prefix}{a}.weight", (256, in_dim)), (f"{prefix}{a}.bias", (256,)), (f"{prefix}{b}.weight", (256, 256)),
            (f"{prefix}{b}.bias", (256,)), (f"{prefix}{c}.weight", (out_dim, 256)), (f"{prefix}{c}.bias", (out_dim,))]


TD3_ACTOR_LAYOUT = _mlp_layout("net.", (0, 2, 4), 2, 10)
TD3_CRITIC_LAYOUT = _mlp_layout("q1.", (0, 2, 4), 12, 1) + _mlp_layout("q2.", (0, 2, 4), 12, 1)
SAC_POLICY_LAYOUT = [("net.0.weight", (256, 2)), ("net.0.bias", (256,)), ("net.2.weight", (256, 256)), ("net.2.bias", (256,)),
                     ("mean.weight", (10, 256)), ("mean.bias", (10,)), ("log_std.weight", (10, 256)), ("log_std.bias", (10,))]
SAC_Q_LAYOUT = _mlp_layout("q1.q.", (0, 2, 4), 12, 1) + _mlp_layout("q2.q.", (0, 2, 4), 12, 1)


def _seq(in_dim, out_dim):
    return nn.Sequential(nn.Linear(in_dim, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, out_dim))


def td3_state_dicts(seed=None):
    """(actor_sd, critic_sd) consuming torch's RNG in TD3Agent.__init__'s order (td3_agent.py:156-164)."""
    if seed is not None:
        torch.manual_seed(seed)
    actor = {"net." + k: v.detach().clone() for k, v in _seq(2, 10).state_dict().items()}
    _seq(2, 10)                                          # actor_target
    q1, q2 = _seq(12, 1), _seq(12, 1)
    critic = {"q1." + k: v.detach().clone() for k, v in q1.state_dict().items()}
    critic.update({"q2." + k: v.detach().clone() for k, v in q2.state_dict().items()})
    _seq(12, 1); _seq(12, 1)                             # critic_target
    return actor, critic


def sac_state_dicts(seed=None):
    """(policy_sd, q_sd) in SACAgent.__init__'s order (sac_agent.py:134-141); q_sd holds q1.* and q2.*."""
    if seed is not None:
        torch.manual_seed(seed)
    net = nn.Sequential(nn.Linear(2, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU())
    mean, log_std = nn.Linear(256, 10), nn.Linear(256, 10)
    pol = {"net." + k: v.detach().clone() for k, v in net.state_dict().items()}
    pol.update({"mean." + k: v.detach().clone() for k, v in mean.state_dict().items()})
    pol.update({"log_std." + k: v.detach().clone() for k, v in log_std.state_dict().items()})
    q = {}
    for name in ("q1", "q2"):
        q.update({f"{name}.q." + k: v.detach().clone() for k, v in _seq(12, 1).state_dict().items()})
    return pol, q


def _bounds_tensors(bounds, device):
    lo = torch.tensor([float(v[0]) for v in bounds.values()], dtype=torch.float32, device=device)
    hi = torch.tensor([float(v[1]) for v in bounds.values()], dtype=torch.float32, device=device)
    return lo, hi


class _T1Base:
    """Shared plumbing: workspace, replay ring, sampling, data-parallel bucket views."""

    def __init__(self, env, seed, batch_size, memory_size, rank):
        L.require_device()
        self.env, self.device, self.n = env, env.device, env.n_env
        self.seed, self.rank, self.batch_size = int(seed), int(rank), int(batch_size)
        self.memory = ReplayBuffer(memory_size, device=self.device)
        self.lo, self.hi = _bounds_tensors(env.conditions[0]["bounds"], self.device)
        self.action = torch.zeros(self.n, 10, dtype=torch.float32, device=self.device)
        self.norm_action = torch.zeros(self.n, 10, dtype=torch.float32, device=self.device)
        self.state = torch.zeros(self.n, 2, dtype=torch.float32, device=self.device)
        self.idx = torch.zeros(batch_size, dtype=torch.int64, device=self.device)
        self.batch = torch.zeros(batch_size, L.REC, dtype=torch.float32, device=self.device)
        self.report = torch.zeros(3, dtype=torch.float32, device=self.device)
        self._zero3 = torch.zeros(self.n, 3, dtype=torch.float32, device=self.device)
        self._ws = torch.zeros((L.lib().erlfill_t1_workspace_bytes(C.c_int(batch_size)) + 3) // 4, dtype=torch.float32,
                               device=self.device)
        self._act_ws = torch.zeros((L.lib().erlfill_t1_act_workspace_bytes() + 3) // 4, dtype=torch.float32, device=self.device)
        self.t = 0
        self.updates = 0

    def grad_buckets(self, critic_len, actor_len):
        cg, ag = C.c_void_p(), C.c_void_p()
        L.check(L.lib().erlfill_t1_grad_buckets(L.ptr(self._ws), C.c_int(self.batch_size), C.byref(cg), C.byref(ag)),
                "erlfill_t1_grad_buckets")
        co, ao = (cg.value - self._ws.data_ptr()) // 4, (ag.value - self._ws.data_ptr()) // 4
        return self._ws[co:co + critic_len], self._ws[ao:ao + actor_len]

    def remember(self, state, norm_action, reward, next_state, done):
        """Stores the already normalised action (what both agents' remember() computes from the env action)."""
        self.memory.insert(state, norm_action, reward, next_state, done, self._zero3[:state.shape[0]])

    def _gather(self, idx):
        if idx is None:
            self.memory.sample_indices(self.batch_size, self.seed, self.updates, stream_id=self.rank, out=self.idx)
            idx = self.idx
        if idx.shape[0] > self.batch_size:      # the update workspace is sized for batch_size rows
            raise L.ErlfillError("minibatch of %d indices exceeds batch_size %d" % (idx.shape[0], self.batch_size))
        return self.memory.gather(idx, out=self.batch if idx.shape[0] == self.batch_size else None)

    def rollout_step(self, store=True, **act_kw):
        """act -> env.step -> remember for all N envs (train_td3 / train_sac loop bodies without the update)."""
        env = self.env
        self.state.copy_(env.obs)
        a = self.act_batch(env.obs, **act_kw)
        next_state, reward, done, _ = env.step(a)
        if store:
            self.remember(self.state, self.norm_action, reward, next_state, done)
        self.t += 1
        return next_state, reward, done


class VecTD3(_T1Base):
    """TD3Agent for N envs on one GPU."""

    def __init__(self, env, seed=0, batch_size=128, memory_size=100000, state_dicts=None, rank=0, gamma=0.99, tau=0.005,
                 policy_noise=0.2, noise_clip=0.5, policy_delay=2, actor_lr=1e-4, critic_lr=1e-3):
        super().__init__(env, seed, batch_size, memory_size, rank)
        actor_sd, critic_sd = state_dicts if state_dicts is not None else td3_state_dicts(seed)
        d = self.device
        self.actor, self.actor_target = FlatNet(TD3_ACTOR_LAYOUT, L.TD3_ACTOR_PARAMS, d), FlatNet(TD3_ACTOR_LAYOUT, L.TD3_ACTOR_PARAMS, d)
        self.critic, self.critic_target = FlatNet(TD3_CRITIC_LAYOUT, L.TD3_CRITIC_PARAMS, d), FlatNet(TD3_CRITIC_LAYOUT, L.TD3_CRITIC_PARAMS, d)
        self.actor.load(actor_sd); self.actor_target.load(actor_sd)
        self.critic.load(critic_sd); self.critic_target.load(critic_sd)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=d)
        self.actor_m, self.actor_v, self.critic_m, self.critic_v = z(L.TD3_ACTOR_PARAMS), z(L.TD3_ACTOR_PARAMS), z(L.TD3_CRITIC_PARAMS), z(L.TD3_CRITIC_PARAMS)
        self.gamma, self.tau, self.policy_noise, self.noise_clip, self.policy_delay = gamma, tau, policy_noise, noise_clip, policy_delay
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.total_it = 0
        self.actor_steps = 0

    def select_action(self, state, noise_scale=0.1, randn=None):
        n = state.shape[0]
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_td3_act(L.ptr(self.actor.flat), C.c_int(n), L.ptr(state), L.ptr(randn), C.c_float(noise_scale),
                                            L.ptr(self.lo), L.ptr(self.hi), C.c_uint64(self.seed), C.c_uint32(self.env.env_id_base),
                                            C.c_uint32(self.t & 0xFFFFFFFF), L.ptr(self.action), L.ptr(self.norm_action),
                                            L.ptr(self._act_ws), L.stream_ptr()), "erlfill_td3_act")
        return self.action[:n]

    def act(self, state, add_noise=False):
        return self.select_action(state, noise_scale=0.1 if add_noise else 0.0)

    def act_batch(self, state, noise_scale=0.1):
        return self.select_action(state, noise_scale)

    def train_step(self, idx=None, noise=None, allreduce=None):
        """One TD3Agent.train_step(); returns the device report [critic_loss, mean q1, actor_loss] and whether the
        delayed actor update ran (the reference returns actor_loss None otherwise)."""
        if len(self.memory) < self.batch_size:
            return None, False
        self.total_it += 1
        upd = self.total_it % self.policy_delay == 0
        if upd:
            self.actor_steps += 1
        batch = self._gather(idx)
        nets = L.TD3Nets(L.ptr(self.actor.flat), L.ptr(self.actor_target.flat), L.ptr(self.critic.flat), L.ptr(self.critic_target.flat),
                         L.ptr(self.actor_m), L.ptr(self.actor_v), L.ptr(self.critic_m), L.ptr(self.critic_v))
        hyper = L.TD3Hyper(batch.shape[0], self.total_it, max(self.actor_steps, 1), int(upd), self.gamma, self.tau, self.policy_noise,
                           self.noise_clip, self.critic_lr, self.actor_lr, self.seed, self.updates & 0xFFFFFFFF, self.rank)

        def run(ph):
            with torch.cuda.device(self.device):
                L.check(L.lib().erlfill_td3_update(C.byref(nets), C.byref(hyper), L.ptr(batch), L.ptr(noise), L.ptr(self._ws), C.c_int(ph),
                                                   L.ptr(self.report), L.stream_ptr()), "erlfill_td3_update")
        if allreduce is None:
            run(L.PHASE_ALL)
        else:
            cb, ab = self.grad_buckets(L.TD3_CRITIC_PARAMS, L.TD3_ACTOR_PARAMS)
            run(L.PHASE_CRITIC_GRAD); allreduce(cb); run(L.PHASE_CRITIC_APPLY)
            run(L.PHASE_ACTOR_GRAD)
            if upd:
                allreduce(ab)
            run(L.PHASE_ACTOR_APPLY)
        self.updates += 1
        return self.report, upd


class VecSAC(_T1Base):
    """SACAgent (fixed alpha) for N envs on one GPU."""

    def __init__(self, env, seed=0, batch_size=128, memory_size=100000, state_dicts=None, rank=0, gamma=0.99, tau=0.005, alpha=0.3,
                 lr=3e-4):
        super().__init__(env, seed, batch_size, memory_size, rank)
        pol_sd, q_sd = state_dicts if state_dicts is not None else sac_state_dicts(seed)
        d = self.device
        self.policy = FlatNet(SAC_POLICY_LAYOUT, L.SAC_POLICY_PARAMS, d)
        self.q, self.q_target = FlatNet(SAC_Q_LAYOUT, L.SAC_Q_PARAMS, d), FlatNet(SAC_Q_LAYOUT, L.SAC_Q_PARAMS, d)
        self.policy.load(pol_sd); self.q.load(q_sd); self.q_target.load(q_sd)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=d)
        self.policy_m, self.policy_v, self.q_m, self.q_v = z(L.SAC_POLICY_PARAMS), z(L.SAC_POLICY_PARAMS), z(L.SAC_Q_PARAMS), z(L.SAC_Q_PARAMS)
        self.gamma, self.tau, self.alpha, self.lr = gamma, tau, alpha, lr
        self.adam_step = 0

    def act(self, state, deterministic=False, eps=None):
        n = state.shape[0]
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_sac_act(L.ptr(self.policy.flat), C.c_int(n), L.ptr(state), L.ptr(eps), C.c_int(int(deterministic)),
                                            L.ptr(self.lo), L.ptr(self.hi), C.c_uint64(self.seed), C.c_uint32(self.env.env_id_base),
                                            C.c_uint32(self.t & 0xFFFFFFFF), L.ptr(self.action), L.ptr(self.norm_action),
                                            L.ptr(self._act_ws), L.stream_ptr()), "erlfill_sac_act")
        return self.action[:n]

    def act_batch(self, state, deterministic=False):
        return self.act(state, deterministic)

    def update(self, idx=None, eps_next=None, eps_new=None, allreduce=None):
        """One SACAgent.update(); returns the device report [q1_loss, q2_loss, policy_loss] (None until the buffer
        holds a batch)."""
        if len(self.memory) < self.batch_size:
            return None
        self.adam_step += 1
        batch = self._gather(idx)
        nets = L.SACNets(L.ptr(self.policy.flat), L.ptr(self.q.flat), L.ptr(self.q_target.flat), L.ptr(self.policy_m), L.ptr(self.policy_v),
                         L.ptr(self.q_m), L.ptr(self.q_v))
        hyper = L.SACHyper(batch.shape[0], self.adam_step, self.gamma, self.tau, self.alpha, 0.0, self.lr, self.seed,
                           self.updates & 0xFFFFFFFF, self.rank)

        def run(ph):
            with torch.cuda.device(self.device):
                L.check(L.lib().erlfill_sac_update(C.byref(nets), C.byref(hyper), L.ptr(batch), L.ptr(eps_next), L.ptr(eps_new),
                                                   L.ptr(self._ws), C.c_int(ph), L.ptr(self.report), L.stream_ptr()), "erlfill_sac_update")
        if allreduce is None:
            run(L.PHASE_ALL)
        else:
            cb, ab = self.grad_buckets(L.SAC_Q_PARAMS, L.SAC_POLICY_PARAMS)
            run(L.PHASE_CRITIC_GRAD); allreduce(cb); run(L.PHASE_CRITIC_APPLY)
            run(L.PHASE_ACTOR_GRAD); allreduce(ab); run(L.PHASE_ACTOR_APPLY)
        self.updates += 1
        return self.report
