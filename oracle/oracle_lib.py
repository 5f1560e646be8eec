"""ctypes binding of oracle/_build/liberlfill_oracle.so (TEST INFRASTRUCTURE ONLY).

Parity status: pinned against tests/golden/*.npz (generated from the unmodified reference by
oracle/gen_golden.py); see tests/test_oracle_golden.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liberlfill_oracle.so")

RING = 20
ENV_MULTI, ENV_SINGLE = 0, 1
ENV_ONBOARD = 2      # OR into the kind: the on-board per-tick motor / weight models (row N1) instead of the offline controller's
ACTION_ORDER = ["fast_weight", "medium_weight", "slow_weight", "fast_opening", "medium_opening",
                "slow_opening", "fast_delay", "medium_delay", "slow_delay", "unload_delay"]


class Condition(C.Structure):
    _fields_ = [("target_weight", C.c_int), ("target_weight_err", C.c_double),
                ("target_time", C.c_double), ("lo", C.c_float * 10), ("hi", C.c_float * 10)]


class EnvState(C.Structure):
    _fields_ = [("step_counter", C.c_int), ("episode_counter", C.c_int), ("ring_len", C.c_int),
                ("ring_head", C.c_int), ("ring", C.c_double * RING)]


class StepOut(C.Structure):
    _fields_ = [("next_state", C.c_float * 2), ("reward", C.c_float), ("reward64", C.c_double),
                ("done", C.c_int), ("final_weight", C.c_double), ("total_time", C.c_double),
                ("weight_error", C.c_double), ("ticks", C.c_int), ("iparams", C.c_int * 11),
                ("clipped", C.c_float * 10)]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "erlfill_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        assert L.orc_sizeof_condition() == C.sizeof(Condition)
        assert L.orc_sizeof_env_state() == C.sizeof(EnvState)
        assert L.orc_sizeof_step_out() == C.sizeof(StepOut)
        L.orc_philox_sysdelay.restype = C.c_double
        L.orc_env_step_batch_philox.restype = C.c_longlong
        _lib = L
    return _lib


def make_condition(target_weight, target_weight_err, target_time, bounds):
    c = Condition()
    c.target_weight = int(target_weight)
    c.target_weight_err = float(target_weight_err)
    c.target_time = float(target_time)
    for i, k in enumerate(ACTION_ORDER):
        c.lo[i], c.hi[i] = float(bounds[k][0]), float(bounds[k][1])
    return c


def philox(seed, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32(C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32), C.c_uint32(c0),
                         C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), out)
    return list(out)


def simulate_run_noise(params, sysdelay, u):
    p = (C.c_int * 11)(*[int(x) for x in params])
    u = np.ascontiguousarray(u, np.float64)
    tt, fw, tick = C.c_double(), C.c_double(), C.c_int()
    n = lib().orc_simulate_run_noise(p, C.c_double(sysdelay), u.ctypes.data_as(C.POINTER(C.c_double)),
                                     C.c_int(len(u)), C.byref(tt), C.byref(fw), C.byref(tick))
    return n, tt.value, fw.value, tick.value


def simulate_onboard_noise(params, sysdelay, u):
    """Row N1: the env classes' on-board motor_simulation / weight_simulation under simulate_run's stage logic."""
    p = (C.c_int * 11)(*[int(x) for x in params])
    u = np.ascontiguousarray(u, np.float64)
    tt, fw, tick = C.c_double(), C.c_double(), C.c_int()
    n = lib().orc_simulate_onboard_noise(p, C.c_double(sysdelay), u.ctypes.data_as(C.POINTER(C.c_double)),
                                         C.c_int(len(u)), C.byref(tt), C.byref(fw), C.byref(tick))
    return n, tt.value, fw.value, tick.value


def simulate_run_philox(params, seed, env_id, step_idx):
    p = (C.c_int * 11)(*[int(x) for x in params])
    tt, fw, tick = C.c_double(), C.c_double(), C.c_int()
    n = lib().orc_simulate_run_philox(p, C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32),
                                      C.c_uint32(env_id), C.c_uint32(step_idx), C.byref(tt), C.byref(fw),
                                      C.byref(tick))
    return n, tt.value, fw.value, tick.value


def env_step_noise(kind, cond, state, max_steps, action, sysdelay, u):
    a = (C.c_float * 10)(*[float(x) for x in np.asarray(action, np.float32)])
    u = np.ascontiguousarray(u, np.float64)
    o = StepOut()
    lib().orc_env_step_noise(C.c_int(kind), C.byref(cond), C.byref(state), C.c_int(max_steps), a,
                             C.c_double(sysdelay), u.ctypes.data_as(C.POINTER(C.c_double)),
                             C.c_int(len(u)), C.byref(o))
    return o


def env_reset_cold(cond, state, u0, u1):
    s = (C.c_float * 2)()
    lib().orc_env_reset_cold(C.byref(cond), C.byref(state), C.c_double(u0), C.c_double(u1), s)
    return np.array(s, np.float32)


def env_reset_replay(cond, state, states10):
    s = (C.c_float * 2)()
    a = np.ascontiguousarray(states10, np.float32)
    lib().orc_env_reset_replay(C.byref(cond), C.byref(state), a.ctypes.data_as(C.POINTER(C.c_float)), s)
    return np.array(s, np.float32)


def emotion_update(cur, reward, movement):
    c = (C.c_double * 3)(*cur)
    m = (C.c_float * 2)(*[float(x) for x in movement])
    p = (C.c_float * 3)()
    lib().orc_emotion_update(c, C.c_double(reward), m, p)
    return np.array(c, np.float64), np.array(p, np.float32)


def ou_act(ou_state, sigma, randn, scaled_action, scale, offset):
    f10 = C.c_float * 10
    st = f10(*ou_state)
    out = f10()
    lib().orc_ou_act(st, C.c_float(sigma), f10(*randn), f10(*scaled_action), f10(*scale), f10(*offset), out)
    return np.array(st, np.float32), np.array(out, np.float32)


def normalize_action(a, scale, offset):
    f10 = C.c_float * 10
    out = f10()
    lib().orc_normalize_action(f10(*a), f10(*scale), f10(*offset), out)
    return np.array(out, np.float32)


def philox_randn10(seed, env_id, step_idx):
    out = (C.c_float * 10)()
    lib().orc_philox_randn10(C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32), C.c_uint32(env_id),
                             C.c_uint32(step_idx), out)
    return np.array(out, np.float32)


def sample_indices_uniform(u, n):
    u = np.ascontiguousarray(u, np.float64)
    idx = np.zeros(len(u), np.int64)
    lib().orc_sample_indices_uniform(u.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(len(u)),
                                     C.c_int64(n), idx.ctypes.data_as(C.POINTER(C.c_int64)))
    return idx


def sample_indices_philox(seed, update_idx, B, n):
    idx = np.zeros(B, np.int64)
    lib().orc_sample_indices_philox(C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(seed >> 32),
                                    C.c_uint32(update_idx), C.c_int(B), C.c_int64(n),
                                    idx.ctypes.data_as(C.POINTER(C.c_int64)))
    return idx


class BatchEnv:
    """N independent envs stepped by the C oracle (Philox noise) — the checker for the CUDA kernel
    and the `cpu_baseline` leg of bench.py."""

    def __init__(self, kind, conds, n_env, max_steps=100, seed=0, env_id_base=0, cond_id=None):
        self.kind, self.n, self.max_steps, self.seed, self.base = kind, n_env, max_steps, seed, env_id_base
        self.conds = (Condition * len(conds))(*conds)
        self.n_cond = len(conds)
        self.cond_id = None if cond_id is None else np.ascontiguousarray(cond_id, np.int32)
        self.states = (EnvState * n_env)()
        self.next_state = np.zeros((n_env, 2), np.float32)
        self.obs = np.zeros((n_env, 2), np.float32)
        self.reward = np.zeros(n_env, np.float32)
        self.done = np.zeros(n_env, np.uint8)
        self.final_weight = np.zeros(n_env, np.float64)
        self.total_time = np.zeros(n_env, np.float64)
        self.ticks = np.zeros(n_env, np.int32)

    def step(self, actions, step_idx, auto_reset=True, n_threads=0):
        a = np.ascontiguousarray(actions, np.float32)
        assert a.shape == (self.n, 10)
        P = lambda arr, t: arr.ctypes.data_as(C.POINTER(t))
        cid = None if self.cond_id is None else P(self.cond_id, C.c_int)
        total = lib().orc_env_step_batch_philox(
            C.c_int(self.kind), C.c_int(self.n), self.conds, C.c_int(self.n_cond), cid, self.states,
            C.c_int(self.max_steps), P(a, C.c_float), C.c_uint32(self.seed & 0xFFFFFFFF),
            C.c_uint32(self.seed >> 32), C.c_uint32(self.base), C.c_uint32(step_idx), C.c_int(int(auto_reset)),
            P(self.next_state, C.c_float), P(self.obs, C.c_float), P(self.reward, C.c_float),
            P(self.done, C.c_ubyte), P(self.final_weight, C.c_double), P(self.total_time, C.c_double),
            P(self.ticks, C.c_int), C.c_int(n_threads))
        return int(total)
