"""Batched ER-DDPG on one GPU (one process per GPU): the rollout-and-update loop of train_ddpg
(DifferentModules/ddpg_agent.py:628-729) for N envs in lock-step.

Per lock-step step (reference order):
  act      a_t = clip(actor(s_t, e_t) + OU, bounds)                         A1 + A2   (:638)
  step     s'_t, r_t, done_t = env.step(a_t)                                E1-E4     (:639)
  emotion  EmotionModule.update(r_t, s'_t - s_t)                            M1        (:641-642)
           e_{t+1} = get_emotion()                                          M2        (:370 in remember)
  remember store (s_t, norm(a_t), r_t, s'_t, done_t, e_{t+1})               R1        (:644)
  learn    one minibatch update                                             R2, U1-U5 (:649)

Batched definitions the reference leaves open: every env has its own OU state/sigma and
its own emotion history; with dropout off one transformer call per step serves all of the reference's
get_emotion() call sites (they differ only by their dropout draws); learn()'s fresh emotion (:466) is the mean of
e_{t+1} over this rank's envs (N = 1: the reference's value); train_step = 1 + completed episodes // N.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ddpg, init, nets


class VecERDDPG:
    def __init__(self, env, seed=0, batch_size=128, memory_size=100000, dropout="off", state_dicts=None,
                 emotion_precision="fp32", rank=0, world_size=1, update_precision="exact"):
        self.env = env
        self.device = env.device
        self.n = env.n_env
        self.seed = int(seed)
        self.rank, self.world = rank, world_size
        bounds = env.conditions[0]["bounds"]
        if state_dicts is None:
            state_dicts = init.er_ddpg_state_dicts(bounds, seed=seed)
        actor_sd, critic_sd, trans_sd = state_dicts
        self.learner = ddpg.DDPGLearner(actor_sd, critic_sd, device=self.device, batch_size=batch_size, precision=update_precision,
                                        dropout=dropout, seed=seed, rank=rank)
        self.transformer = nets.TransformerParams(trans_sd, device=self.device)
        self.emotion_precision = emotion_precision
        self.memory = ddpg.ReplayBuffer(memory_size, device=self.device)
        self.batch_size = batch_size
        self.dropout = {"off": L.DROPOUT_OFF, "philox": L.DROPOUT_PHILOX}[dropout]
        d, n = self.device, self.n
        self.emo = nets.EmotionStateVec(n, device=d)
        self.emotion = torch.zeros(n, 3, dtype=torch.float32, device=d)       # e_t
        self.ou_state = torch.zeros(n, 10, dtype=torch.float32, device=d)
        self.sigma64 = torch.full((n,), 0.6, dtype=torch.float64, device=d)
        self.sigma = self.sigma64.float()
        self.action = torch.zeros(n, 10, dtype=torch.float32, device=d)       # a_t (physical units, clipped)
        self.scaled = torch.zeros(n, 10, dtype=torch.float32, device=d)
        self.state = torch.zeros(n, 2, dtype=torch.float32, device=d)         # s_t
        self.next_emotion = torch.zeros(3, dtype=torch.float32, device=d)
        self.idx = torch.zeros(batch_size, dtype=torch.int64, device=d)
        self.batch = torch.zeros(batch_size, L.REC, dtype=torch.float32, device=d)
        self._stats_partials = torch.zeros(L.lib().erlfill_replay_sample_gather_blocks(C.c_int(batch_size)), 8, dtype=torch.float32, device=d)
        self.episodes = torch.zeros(1, dtype=torch.int64, device=d)
        self.train_step = 1
        self.t = 0
        self.updates = 0
        self._actor_struct = self.learner.actor_struct()
        # one agent per condition in the reference (experiment_runner.py:504-510): each env's action is scaled with the
        # scale_params / offset_params of ITS condition's bounds (ddpg_agent.py:56-71)
        self._cond = None
        if len(env.conditions) > 1:
            sc = [[(c["bounds"][k][1] - c["bounds"][k][0]) / 2.0 for k in L.ACTION_ORDER] for c in env.conditions]
            of = [[(c["bounds"][k][1] + c["bounds"][k][0]) / 2.0 for k in L.ACTION_ORDER] for c in env.conditions]
            self.learner.set_condition_scaling(torch.tensor(sc, dtype=torch.float32), torch.tensor(of, dtype=torch.float32))
            self._cond_id = env._d_cond_id          # explicit table or None (round-robin on the global env id, like the env)
            self._cond = L.CondScaling(L.ptr(self.learner.scale_tab), L.ptr(self.learner.offset_tab), L.ptr(self._cond_id),
                                       len(env.conditions), env.env_id_base)
        self.reset()

    # -- pieces -------------------------------------------------------------------------------
    def reset(self):
        self.env.reset()
        # anneal once at the start of the first episode like train_ddpg (:633-634)
        nets.episode_end(torch.ones(self.n, dtype=torch.uint8, device=self.device), self.ou_state, self.sigma, self.sigma64)
        self._emotion_forward()

    def _emotion_forward(self):
        nets.emotion_forward(self.transformer, emo=self.emo, dropout=self.dropout, seed=self.seed, call_idx=self.t,
                             row_id_base=self.env.env_id_base, out=self.emotion, precision=self.emotion_precision)

    def act(self, add_noise=True):
        """A1+A2 for all envs on the current observations; returns the actions handed to env.step."""
        import ctypes as C
        w = self._actor_struct
        ou = L.OU(L.ptr(self.ou_state), L.ptr(self.sigma), None, L.ptr(self.action), self.seed, self.env.env_id_base,
                  self.t & 0xFFFFFFFF)
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_actor_forward_cond(C.byref(w), C.c_int(self.n), L.ptr(self.env.obs), L.ptr(self.emotion),
                                                       C.c_int(0), C.c_int(self.dropout), C.c_uint64(self.seed),
                                                       C.c_uint32(self.t & 0xFFFFFFFF), C.c_uint32(self.env.env_id_base), None, None,
                                                       L.ptr(self.scaled), None, C.byref(ou) if add_noise else None,
                                                       C.byref(self._cond) if self._cond is not None else None, L.stream_ptr()),
                    "erlfill_actor_forward_cond")
        return self.action if add_noise else self.scaled

    def rollout_step(self, store=True, replay_reset=True):
        """One lock-step env step for all N envs (everything stays on the device, no host sync)."""
        env = self.env
        self.state.copy_(env.obs)
        a = self.act()
        next_state, reward, done, _ = env.step(a)
        if replay_reset and len(self.memory) >= 10:
            # reset()'s replay branch for the envs that just finished (MutiConditionEnv.py:145-154)
            env.reset_from_replay(self.memory.buf, mask=done, n_valid=len(self.memory), update_counters=False)
        nets.emotion_update(self.emo, reward, self.state, next_state)
        self.t += 1
        self._emotion_forward()
        if store:
            if self._cond is not None:
                self.memory.insert(self.state, a, reward, next_state, done, self.emotion, cond=self._cond)
            else:
                self.memory.insert(self.state, a, reward, next_state, done, self.emotion, self.learner.scale, self.learner.offset)
        nets.episode_end(done, self.ou_state, self.sigma, self.sigma64, episode_count=self.episodes)
        return next_state, reward, done

    def launches_per_rollout_step(self, store=True, replay_reset=True):
        """Kernels of liberlfill.so one rollout_step launches: actor(+OU), simulator (one fused kernel for a resident batch, else
        simulator + finish), [replay-branch reset], emotion update, transformer, [replay insert], episode end."""
        with torch.cuda.device(self.device):
            fused = self.env.sim_mode != L.SIM_THREAD_PER_ENV and self.n <= L.lib().erlfill_env_step_fused_max_envs()
        sim = 1 if (fused or self.env.sim_mode == L.SIM_THREAD_PER_ENV) else 2
        return 1 + sim + (1 if (replay_reset and len(self.memory) >= 10) else 0) + 1 + 1 + (1 if store else 0) + 1

    def rollout_eval_step(self):
        """One evaluation step (experiment_runner.py:545-556: `act(state, add_noise=False)` -> `env.step`).  The test
        loop never calls `emotion.update`, so every env keeps the emotion its history gave at the end of training."""
        return self.env.step(self.act(add_noise=False))

    def learn(self, allreduce=None):
        """R2 + U1-U5 on a fresh minibatch; returns the device report [critic_loss, actor_loss, critic_lr, actor_lr]."""
        if len(self.memory) < self.batch_size:
            return None
        # one launch: index draw + assembly (+ the emotion sums Critic.forward's batch standardisation needs)
        self.memory.sample_gather(self.batch_size, self.seed, self.updates, stream_id=self.rank, out=self.batch, idx_out=self.idx,
                                  stats_out=self._stats_partials)
        torch.mean(self.emotion, dim=0, out=self.next_emotion)
        rep = self.learner.update(self.batch, self.next_emotion, self.train_step, allreduce=allreduce,
                                  stats_partials=self._stats_partials if self.learner.precision == "fast" else None)
        self.updates += 1
        return rep

    def sync_counters(self):
        """Host copy of the episode count (one D2H sync): train_step = 1 + episodes // N (:455, :729).  Data-parallel: episodes
        and env counts are summed over the ranks first, so every replica decays its learning rate identically (ranks finish
        episodes at different times; a rank-local count would let the replicas drift apart)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.stack([self.episodes.reshape(()).to(torch.int64), torch.tensor(self.n, dtype=torch.int64, device=self.device)])
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            self.train_step = 1 + int(t[0].item()) // int(t[1].item())
        else:
            self.train_step = 1 + int(self.episodes.item()) // self.n
        return self.train_step
