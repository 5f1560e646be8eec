"""Batched filling-line environments on the GPU, plus N=1 drop-ins with the reference's API.

  VecFillEnv                 N independent envs on one GPU, torch tensors in/out, one fused CUDA
                             kernel per lock-step step (csrc/env_step.cu through the C ABI).
  WeightEnv()                drop-in for /root/reference/WeightEnv.py:21-588 (offline-sim branch)
  MultiConditionWeightEnv()  drop-in for /root/reference/MutiConditionEnv.py:470-486

The drop-ins keep the reference's duck-typed surface: `bounds`, `state_dim`,
`action_dim`, `max_steps`, `target_weight`, `scale_centers/scale_ranges`, `attach_agent`, `reset()`,
`step(action) -> (np.float32[2], float, bool, info)`, `normalize_action`.
"""
import ctypes as C
import os
import random

import numpy as np
import torch

from . import _lib as L
from . import conditions as K

MAX_DRAWS = 5002   # 1 system_delay + at most 5001 tick draws (VirtualWeightController.py:158)


class VecFillEnv:
    """N envs stepped in lock-step by one kernel launch.

    conditions : list of dicts {target_weight, target_weight_err, target_time, bounds}
    cond_id    : optional int32[N]; default env (env_id_base + i) uses condition (env_id_base + i) % len(conditions)
    kind       : "multi" (MutiConditionEnv reward) or "single" (WeightEnv reward v2)
    sim_mode   : "auto" | "warp" (one warp per env, integer hold phase) | "thread" (one thread per env)
    onboard    : per tick the env classes' "on-board" motor / weight models (WeightEnv.py:176-229) instead of the offline controller's
    """

    SIM_MODES = {"auto": L.SIM_AUTO, "thread": L.SIM_THREAD_PER_ENV, "warp": L.SIM_WARP_PER_ENV,
                 "warp_windows": L.SIM_WARP_WINDOWS, "warp_reject_runs": L.SIM_WARP_REJECT_RUNS}   # last two: test hooks

    def __init__(self, conditions, n_env, kind="multi", max_steps=100, seed=0, device="cuda",
                 env_id_base=0, cond_id=None, auto_reset=True, sim_mode="auto", onboard=False):
        L.require_device()
        self.device = torch.device(device)
        self.n_env = int(n_env)
        # onboard: the env classes' own per-tick motor / weight models (speed cap, int()-truncated weight)
        self.onboard = bool(onboard)
        self.kind = {"multi": L.ENV_MULTI, "single": L.ENV_SINGLE}[kind] | (L.ENV_ONBOARD if onboard else 0)
        self.max_steps = int(max_steps)
        self.seed = int(seed)
        self.env_id_base = int(env_id_base)
        self.auto_reset = bool(auto_reset)
        self.sim_mode = self.SIM_MODES[sim_mode]      # which simulator kernel (bit-identical results; see erlfill.h)
        self.conditions = list(conditions)
        conds = (L.Condition * len(conditions))(*[
            L.make_condition(c["target_weight"], c["target_weight_err"], c["target_time"], c["bounds"])
            for c in conditions])
        raw = np.frombuffer(bytes(conds), dtype=np.uint8).copy()
        dev = self.device
        self._d_cond = torch.from_numpy(raw).to(dev)
        self._d_cond_id = None if cond_id is None else torch.as_tensor(cond_id, dtype=torch.int32, device=dev).contiguous()
        n = self.n_env
        self.step_counter = torch.zeros(n, dtype=torch.int32, device=dev)
        self.episode_counter = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ring_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ring_head = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ring = torch.zeros(L.RING, n, dtype=torch.float64, device=dev)
        self.obs = torch.zeros(n, 2, dtype=torch.float32, device=dev)
        # step outputs (reused every step).  next_state | reward | done are views into ONE allocation so that a host
        # consumer fetches them with a single device-to-host copy (step_host)
        self._out_pack = torch.zeros(13 * n, dtype=torch.uint8, device=dev)
        self.next_state = self._out_pack[:8 * n].view(torch.float32).view(n, 2)
        self.reward = self._out_pack[8 * n:12 * n].view(torch.float32)
        self.done = self._out_pack[12 * n:]
        self._host = None
        self.final_weight = torch.zeros(n, dtype=torch.float64, device=dev)
        self.total_time = torch.zeros(n, dtype=torch.float64, device=dev)
        self.ticks = torch.zeros(n, dtype=torch.int32, device=dev)
        self.tick_sum = torch.zeros(1, dtype=torch.int64, device=dev)
        self.step_idx = 0
        self._out = L.StepOut(L.ptr(self.next_state), L.ptr(self.reward), L.ptr(self.done),
                              L.ptr(self.final_weight), L.ptr(self.total_time), L.ptr(self.ticks),
                              L.ptr(self.tick_sum))

    def use_host_outputs(self):
        """Point every step output at pinned, device-mapped HOST memory (zero-copy): after a stream synchronise the caller
        reads next_state / reward / done / final_weight / total_time / ticks as numpy views, with no device-to-host copies.
        Only for batches that run as one kernel (include/erlfill.h: erlfill_env_step_fused_max_envs) — the scalar drop-ins."""
        n = self.n_env
        with torch.cuda.device(self.device):
            if n > L.lib().erlfill_env_step_fused_max_envs():
                raise ValueError("use_host_outputs: batch too large for the single-kernel step")
        if self.onboard and n > 256:
            raise ValueError("use_host_outputs: the on-board variant runs one kernel of one CTA per 32..128 envs; keep host outputs to small batches")
        buf = torch.zeros(40 * n, dtype=torch.uint8).pin_memory()
        self._host_buf = buf
        self.final_weight = buf[:8 * n].view(torch.float64)
        self.total_time = buf[8 * n:16 * n].view(torch.float64)
        self.next_state = buf[16 * n:24 * n].view(torch.float32).view(n, 2)
        self.reward = buf[24 * n:28 * n].view(torch.float32)
        self.ticks = buf[28 * n:32 * n].view(torch.int32)
        self.done = buf[32 * n:33 * n]
        self._out = L.StepOut(L.ptr(self.next_state), L.ptr(self.reward), L.ptr(self.done),
                              L.ptr(self.final_weight), L.ptr(self.total_time), L.ptr(self.ticks),
                              L.ptr(self.tick_sum))

    # -- C-ABI descriptor (rebuilt per call: max_steps is assigned by trainers at any time) ----
    def _batch(self):
        return L.EnvBatch(self.n_env, self.kind, int(self.max_steps), len(self.conditions), self.env_id_base, 0,
                          self.seed, L.ptr(self._d_cond), L.ptr(self._d_cond_id), L.ptr(self.step_counter),
                          L.ptr(self.episode_counter), L.ptr(self.ring_len), L.ptr(self.ring_head),
                          L.ptr(self.ring), L.ptr(self.obs))

    @property
    def cond_id(self):
        """int32[N]: the condition each env runs (explicit table, or round-robin on the global env id)."""
        if self._d_cond_id is not None:
            return self._d_cond_id
        return ((torch.arange(self.n_env, device=self.device, dtype=torch.int64) + self.env_id_base) % len(self.conditions)).to(torch.int32)

    @property
    def target_weight_per_env(self):
        tw = torch.tensor([float(c["target_weight"]) for c in self.conditions], dtype=torch.float64, device=self.device)
        return tw[self.cond_id.long()]

    @staticmethod
    def _noise(u=None, reset_u=None):
        if u is None and reset_u is None:
            return None, None
        nz = L.Noise(L.ptr(u), 0 if u is None else int(u.shape[0]), 0, L.ptr(reset_u))
        return nz, C.byref(nz)

    def reset(self, mask=None, reset_u=None):
        """Cold-start reset (MutiConditionEnv.py:135-164) of the masked envs (all if mask is None)."""
        b = self._batch()
        nz, nzp = self._noise(None, reset_u)
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_env_reset(C.byref(b), L.ptr(mask), C.c_uint32(0xFFFFFFFF - (self.step_idx & 0xFFFF)),
                                              nzp, L.stream_ptr()), "erlfill_env_reset")
        return self.obs

    def reset_from_replay(self, replay_rows, idx=None, mask=None, n_valid=0, update_counters=True):
        """Replay branch of reset (MutiConditionEnv.py:145-154).  replay_rows: [C, stride] float32 whose first
        two columns are stored states; idx int64[N,10] row indices or None (Philox draws over n_valid rows)."""
        b = self._batch()
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_env_reset_from_replay(
                C.byref(b), L.ptr(mask), L.ptr(replay_rows), C.c_int(replay_rows.shape[1]), C.c_int64(int(n_valid)),
                L.ptr(idx), C.c_uint32(self.step_idx & 0xFFFFFFFF), C.c_int(int(update_counters)), L.stream_ptr()),
                "erlfill_env_reset_from_replay")
        return self.obs

    def step(self, actions, noise_u=None, reset_u=None, auto_reset=None):
        """actions float32[N,10] (physical units) on the device.  Returns (next_state, reward, done, info);
        `self.obs` is the observation to act on next (reset state where done and auto_reset)."""
        if actions.dtype != torch.float32 or not actions.is_contiguous() or tuple(actions.shape) != (self.n_env, 10):
            raise ValueError("actions must be a contiguous float32 tensor of shape [n_env, 10]")
        if actions.device != self.obs.device:
            raise ValueError("actions must live on the env's device")
        b = self._batch()
        nz, nzp = self._noise(noise_u, reset_u)
        ar = self.auto_reset if auto_reset is None else bool(auto_reset)
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_env_step_set_mode(C.c_int(self.sim_mode)), "erlfill_env_step_set_mode")
            L.check(L.lib().erlfill_env_step(C.byref(b), L.ptr(actions), C.c_uint32(self.step_idx & 0xFFFFFFFF),
                                             C.c_int(int(ar)), nzp, C.byref(self._out), L.stream_ptr()),
                    "erlfill_env_step")
        self.step_idx += 1
        info = {"final_weight": self.final_weight, "total_time": self.total_time, "ticks": self.ticks}
        return self.next_state, self.reward, self.done, info

    def step_host(self, actions_host):
        """Host-facing step: float32[N,10] actions in (pinned) host memory in, (next_state, reward, done) as numpy-viewable
        pinned host tensors out.  Batches that run as one fused kernel (<= erlfill_env_step_fused_max_envs() envs) read the
        actions from and write the results to the pinned host buffers directly; larger ones use one H2D copy, the step
        kernels and one D2H copy of the packed outputs.  Either way the sequence is captured ONCE in a CUDA
        graph (the lock-step index lives in device memory: erlfill_env_step_counted) and replayed per call, then one stream
        sync.  `actions_host` must be the same pinned tensor on every call (the graph holds its address); max_steps, the
        simulator kernel choice and auto_reset are frozen at the first call."""
        n = self.n_env
        if self._host is None:
            if not actions_host.is_pinned():
                raise ValueError("step_host: actions_host must be a pinned host tensor")
            pack = torch.empty(13 * n, dtype=torch.uint8).pin_memory()
            a_dev = torch.empty(n, 10, dtype=torch.float32, device=self.device)
            v = self.step_idx & 0xFFFFFFFF
            d_idx = torch.tensor([v - (1 << 32) if v >= (1 << 31) else v, 0], dtype=torch.int32, device=self.device)   # uint32 {index, ticket}
            with torch.cuda.device(self.device):
                L.check(L.lib().erlfill_env_step_set_mode(C.c_int(self.sim_mode)), "erlfill_env_step_set_mode")
            b = self._batch()

            # How the bytes cross PCIe.  0: copy nodes (H2D actions, D2H packed outputs).  2: the simulator kernel reads the
            # pinned host actions in place and writes next_state/reward/done straight into the pinned pack (zero-copy; only
            # when the batch runs as ONE fused kernel, whose outputs are written exactly once: include/erlfill.h).  1: reads
            # in place, outputs by copy.  Default: 2 when eligible, else 0; ERLFILL_STEP_HOST_MODE overrides (measurements).
            with torch.cuda.device(self.device):
                fused = (not self.onboard) and self.sim_mode != L.SIM_THREAD_PER_ENV and n <= L.lib().erlfill_env_step_fused_max_envs()
            zc = int(os.environ.get("ERLFILL_STEP_HOST_MODE", "2" if fused else "0"))
            if not fused:
                zc = 0
            out = self._out
            if zc == 2:
                hp = pack.data_ptr()
                out = L.StepOut(hp, hp + 8 * n, hp + 12 * n, L.ptr(self.final_weight), L.ptr(self.total_time), L.ptr(self.ticks),
                                L.ptr(self.tick_sum))
            self._host_out = out

            def body():
                if zc == 0:
                    a_dev.copy_(actions_host, non_blocking=True)
                L.check(L.lib().erlfill_env_step_counted(C.byref(b), L.ptr(a_dev) if zc == 0 else C.c_void_p(actions_host.data_ptr()), L.ptr(d_idx),
                                                         C.c_int(int(self.auto_reset)), None, C.byref(out), L.stream_ptr()),
                        "erlfill_env_step_counted")
                if zc != 2:
                    pack.copy_(self._out_pack, non_blocking=True)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):          # one eager pass first (module load, motor tables) outside the capture
                body()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self.step_idx += 1
            self._host_counter = self.step_idx
            self._host = (g, actions_host, pack, pack[:8 * n].view(torch.float32).view(n, 2), pack[8 * n:12 * n].view(torch.float32),
                          pack[12 * n:], d_idx, b)
            return self._host[3], self._host[4], self._host[5]
        g, a_bound, pack, ns_h, r_h, d_h, d_idx = self._host[:7]
        if actions_host.data_ptr() != a_bound.data_ptr():
            raise ValueError("step_host: pass the same pinned actions tensor on every call (its address is captured)")
        if self._host_counter != self.step_idx:      # step() was called in between: bring the device counter up to date
            v = self.step_idx & 0xFFFFFFFF
            d_idx[0] = v - (1 << 32) if v >= (1 << 31) else v
        g.replay()
        self.step_idx += 1
        self._host_counter = self.step_idx
        torch.cuda.current_stream(self.device).synchronize()
        return ns_h, r_h, d_h


# ----------------------------------------------------------------------------------------------
# N = 1 drop-ins
# ----------------------------------------------------------------------------------------------
class _ScalarEnvBase:
    """Shared implementation of the two reference env classes at N = 1.

    noise="reference": the simulator consumes CPython's global `random` stream exactly like the
    reference (one uniform(15,50), then one uniform(-1,1) per tick; VirtualWeightController.py:122,137),
    so `random.seed(s)` reproduces the reference trajectory bit for bit.  noise="philox": in-kernel
    counter-based noise (fast; statistically equivalent).
    """
    _kind = "multi"

    def __init__(self, cfg, noise="reference", device="cuda", seed=0, onboard=False):
        self._onboard = bool(onboard)
        self.target_weight = cfg["target_weight"]
        self.target_weight_err = cfg["target_weight_err"]
        self.target_time = cfg["target_time"]
        self.bounds = dict(cfg["bounds"])
        self.max_steps = 100
        self.state_dim = 2
        self.action_dim = len(self.bounds)
        self.state = None
        self.agent = None
        self.use_offline_sim = 1
        self.episode_counter = 0
        self.step_counter = 0
        self._noise_mode = noise
        self._device = device
        self._seed = seed
        self._vec = None
        self._update_scaling_params()

    def _update_scaling_params(self):
        self.scale_centers, self.scale_ranges = {}, {}
        for key, (low, high) in self.bounds.items():
            self.scale_centers[key] = (high + low) / 2
            self.scale_ranges[key] = (high - low) / 2

    def _ensure(self):
        if self._vec is None:
            cfg = {"target_weight": self.target_weight, "target_weight_err": self.target_weight_err,
                   "target_time": self.target_time, "bounds": self.bounds}
            self._vec = VecFillEnv([cfg], 1, kind=self._kind, max_steps=self.max_steps, seed=self._seed,
                                   device=self._device, auto_reset=False, onboard=self._onboard)
            self._vec.use_host_outputs()
            self._rs = np.random.RandomState(0)            # scratch generator: its state is overwritten from `random` every step
            self._a_host = torch.zeros(1, 10, dtype=torch.float32).pin_memory()
            self._a_dev = torch.zeros(1, 10, dtype=torch.float32, device=self._vec.device)
            self._u = torch.zeros(MAX_DRAWS, 1, dtype=torch.float64, device=self._vec.device)
            self._u_host = torch.zeros(MAX_DRAWS, 1, dtype=torch.float64).pin_memory()
            self._ru = torch.zeros(2, 1, dtype=torch.float64, device=self._vec.device)
        return self._vec

    def normalize_action(self, action):
        return {key: action[idx] * self.scale_ranges[key] + self.scale_centers[key]
                for idx, key in enumerate(self.bounds)}

    def attach_agent(self, agent):
        self.agent = agent

    def reset(self):
        v = self._ensure()
        if self.episode_counter == 0 or self.agent is None or len(self.agent.memory) < 10:
            u0 = random.random()   # random.uniform(0, x) == 0 + x*random()
            u1 = random.random()
            self._ru.copy_(torch.tensor([[u0], [u1]], dtype=torch.float64))
            v.reset(reset_u=self._ru)
        else:
            samples = random.sample(self.agent.memory.buffer, 10)
            states = torch.tensor(np.stack([np.asarray(s[0], np.float32) for s in samples]), device=v.device)
            idx = torch.arange(10, dtype=torch.int64, device=v.device).view(1, 10)
            v.reset_from_replay(states, idx)
        self.step_counter = 0
        self.episode_counter += 1
        return v.obs[0].cpu().numpy()

    def step(self, action):
        v = self._ensure()
        v.max_steps = int(self.max_steps)
        if isinstance(action, dict):
            action = np.array([action[k] for k in L.ACTION_ORDER])
        self._a_host.numpy()[0, :] = np.asarray(action, dtype=np.float32).reshape(10)
        a = self._a_dev
        a.copy_(self._a_host, non_blocking=True)
        if self._noise_mode == "reference":
            # The reference draws uniform(15, 50) once and uniform(-1, 1) per tick from CPython's global Mersenne Twister
            # (VirtualWeightController.py:122,137).  numpy's RandomState is the same generator with the same 53-bit double
            # construction, so the draws are produced vectorised from a copy of `random`'s state (a + (b - a) * random(),
            # the same two roundings as random.uniform), and `random` is then advanced by exactly the draws the run consumed
            # (getrandbits(64 k) takes 2 k outputs of the generator, like k calls of random.random()).
            st = random.getstate()
            key = np.fromiter(st[1], dtype=np.uint32, count=625)          # 624 state words + the position
            self._rs.set_state(("MT19937", key[:624], int(st[1][-1])))
            u = self._rs.random_sample(MAX_DRAWS)
            host = self._u_host.numpy()
            host[0, 0] = 15.0 + 35.0 * u[0]
            host[1:, 0] = -1.0 + 2.0 * u[1:]
            self._u.copy_(self._u_host, non_blocking=True)
            ns, r, d, info = v.step(a, noise_u=self._u)
            torch.cuda.current_stream(v.device).synchronize()      # outputs live in pinned host memory (use_host_outputs)
            n_draws = int(info["ticks"][0])
            random.getrandbits(64 * (1 + n_draws))     # == 1 + n_draws calls of random.random(): two 32-bit outputs each
        else:
            ns, r, d, info = v.step(a)
            torch.cuda.current_stream(v.device).synchronize()
        clipped = np.clip(np.asarray(action, dtype=np.float32), [self.bounds[k][0] for k in L.ACTION_ORDER],
                          [self.bounds[k][1] for k in L.ACTION_ORDER]).astype(np.float32)
        act = {k: clipped[i] for i, k in enumerate(L.ACTION_ORDER)}
        act["target_weight"] = self.target_weight
        fw = float(info["final_weight"][0])
        tt = float(info["total_time"][0])
        self.step_counter += 1
        return (ns[0].numpy().copy(), float(r[0]), bool(d[0]),
                {"final_weight": fw, "weight_error": abs(fw - self.target_weight), "total_time": tt, "action": act})


class WeightEnv(_ScalarEnvBase):
    """Drop-in for WeightEnv.WeightEnv (WeightEnv.py:21-588), offline-simulator branch."""
    _kind = "single"

    def __init__(self, noise="reference", device="cuda", seed=0, onboard=False):
        super().__init__(K.SINGLE_25KG, noise=noise, device=device, seed=seed, onboard=onboard)


class MultiConditionWeightEnv(_ScalarEnvBase):
    """Drop-in for MutiConditionEnv.MultiConditionWeightEnv (MutiConditionEnv.py:470-486)."""
    _kind = "multi"

    def __init__(self, config=None, noise="reference", device="cuda", seed=0, onboard=False):
        cfg = dict(K.MULTI_CLASS_DEFAULT)
        if config is not None:
            for k in ("target_weight", "target_weight_err", "target_time"):
                cfg[k] = config.get(k, cfg[k])
            if "bounds" in config:
                cfg["bounds"] = config["bounds"]
        super().__init__(cfg, noise=noise, device=device, seed=seed, onboard=onboard)
