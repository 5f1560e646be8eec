"""ctypes binding of erl-fill_b200/lib/liberlfill.so (the C ABI declared in include/erlfill.h).

There is no CPU fallback: if the library is missing or the current device is not sm_100,
`lib()` / `require_device()` raise.  Build with `python erl-fill_b200/build.py` (or
`__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "liberlfill.so")

ACTION_DIM, STATE_DIM, EMOTION_DIM, RING, HIST, REC = 10, 2, 3, 20, 20, 20
ENV_MULTI, ENV_SINGLE, ENV_ONBOARD = 0, 1, 2
ACTION_ORDER = ("fast_weight", "medium_weight", "slow_weight", "fast_opening", "medium_opening",
                "slow_opening", "fast_delay", "medium_delay", "slow_delay", "unload_delay")


class ErlfillError(RuntimeError):
    pass


class Condition(C.Structure):
    _fields_ = [("target_weight", C.c_int32), ("_pad", C.c_int32), ("target_weight_err", C.c_double),
                ("target_time", C.c_double), ("lo", C.c_float * ACTION_DIM), ("hi", C.c_float * ACTION_DIM)]


class EnvBatch(C.Structure):
    _fields_ = [("n_env", C.c_int32), ("kind", C.c_int32), ("max_steps", C.c_int32), ("n_cond", C.c_int32),
                ("env_id_base", C.c_uint32), ("_pad", C.c_uint32), ("seed", C.c_uint64),
                ("d_cond", C.c_void_p), ("d_cond_id", C.c_void_p), ("d_step_counter", C.c_void_p),
                ("d_episode_counter", C.c_void_p), ("d_ring_len", C.c_void_p), ("d_ring_head", C.c_void_p),
                ("d_ring", C.c_void_p), ("d_obs", C.c_void_p)]


class StepOut(C.Structure):
    _fields_ = [("d_next_state", C.c_void_p), ("d_reward", C.c_void_p), ("d_done", C.c_void_p),
                ("d_final_weight", C.c_void_p), ("d_total_time", C.c_void_p), ("d_ticks", C.c_void_p),
                ("d_tick_sum", C.c_void_p)]


class Noise(C.Structure):
    _fields_ = [("d_u", C.c_void_p), ("rows", C.c_int32), ("_pad", C.c_int32), ("d_reset_u", C.c_void_p)]


class ActorWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_in", "b_in", "w_h2", "b_h2", "w_h5", "b_h5", "w_h7", "b_h7",
                                           "emotion_weights", "scale", "offset")]


class OU(C.Structure):
    _fields_ = [("d_state", C.c_void_p), ("d_sigma", C.c_void_p), ("d_randn", C.c_void_p),
                ("d_final_action", C.c_void_p), ("seed", C.c_uint64), ("env_id_base", C.c_uint32),
                ("step_idx", C.c_uint32)]


class CondScaling(C.Structure):
    _fields_ = [("d_scale", C.c_void_p), ("d_offset", C.c_void_p), ("d_cond_id", C.c_void_p), ("n_cond", C.c_int32),
                ("row_id_base", C.c_uint32)]


class EmotionState(C.Structure):
    _fields_ = [("d_current", C.c_void_p), ("d_history", C.c_void_p), ("d_hist_len", C.c_void_p),
                ("d_hist_head", C.c_void_p)]


class TransformerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("emb_w", "emb_b", "pos", "in_proj_w", "in_proj_b", "out_proj_w",
                                           "out_proj_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "norm1_w",
                                           "norm1_b", "norm2_w", "norm2_b", "fc_w", "fc_b")]


class TransformerMasks(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("attn", "d1", "ff", "d2")]


class DDPGNets(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("actor", "actor_target", "critic", "critic_target", "actor_m", "actor_v",
                                           "critic_m", "critic_v", "scale", "offset")]


class DDPGMasks(C.Structure):
    _fields_ = [("keep", C.c_void_p * 10)]


PRECISION_EXACT, PRECISION_FAST = 0, 1


class DDPGHyper(C.Structure):
    _fields_ = [("batch", C.c_int32), ("adam_step", C.c_int32), ("gamma", C.c_float), ("tau", C.c_float),
                ("max_grad_norm", C.c_float), ("n_cond", C.c_int32), ("critic_lr", C.c_double),
                ("actor_lr", C.c_double), ("lr_decay", C.c_double), ("precision", C.c_int32), ("dropout_mode", C.c_int32),
                ("seed", C.c_uint64), ("update_idx", C.c_uint32), ("rank", C.c_uint32), ("masks", C.c_void_p),
                ("d_stats_partials", C.c_void_p), ("n_stats_partials", C.c_int32), ("_pad2", C.c_int32), ("peer", C.c_void_p)]


class TD3Nets(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("actor", "actor_target", "critic", "critic_target", "actor_m", "actor_v",
                                           "critic_m", "critic_v")]


class TD3Hyper(C.Structure):
    _fields_ = [("batch", C.c_int32), ("critic_step", C.c_int32), ("actor_step", C.c_int32), ("update_actor", C.c_int32),
                ("gamma", C.c_float), ("tau", C.c_float), ("policy_noise", C.c_float), ("noise_clip", C.c_float),
                ("critic_lr", C.c_double), ("actor_lr", C.c_double), ("seed", C.c_uint64), ("update_idx", C.c_uint32),
                ("rank", C.c_uint32)]


class SACNets(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("policy", "q", "q_target", "policy_m", "policy_v", "q_m", "q_v")]


class SACHyper(C.Structure):
    _fields_ = [("batch", C.c_int32), ("adam_step", C.c_int32), ("gamma", C.c_float), ("tau", C.c_float),
                ("alpha", C.c_float), ("_pad", C.c_float), ("lr", C.c_double), ("seed", C.c_uint64),
                ("update_idx", C.c_uint32), ("rank", C.c_uint32)]


TD3_ACTOR_PARAMS, TD3_CRITIC_PARAMS, SAC_POLICY_PARAMS, SAC_Q_PARAMS = 69130, 138754, 71700, 138754
ESAC_POLICY_PARAMS, ESAC_Q_PARAMS = 146468, 210369
PHASE_CRITIC_GRAD, PHASE_CRITIC_APPLY, PHASE_ACTOR_GRAD, PHASE_ACTOR_APPLY, PHASE_ALL, PHASE_REDUCED = 1, 2, 4, 8, 15, 16
MAX_PEERS = 8


class PeerGroup(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("ws", C.c_void_p * MAX_PEERS), ("ctrl", C.c_void_p * MAX_PEERS)]

ACTOR_PARAMS, CRITIC_PARAMS = 26216, 103937

DROPOUT_OFF, DROPOUT_PHILOX, DROPOUT_MASK = 0, 1, 2
SIM_AUTO, SIM_THREAD_PER_ENV, SIM_WARP_PER_ENV, SIM_WARP_WINDOWS, SIM_WARP_REJECT_RUNS = 0, 1, 2, 3, 4

# every symbol include/erlfill.h declares (tests/test_abi.py checks the library exports them all)
EXPORTS = [
    "erlfill_version", "erlfill_last_cuda_error", "erlfill_last_cuda_error_string", "erlfill_device_ok",
    "erlfill_env_reset", "erlfill_env_reset_from_replay", "erlfill_env_step", "erlfill_env_step_counted", "erlfill_env_step_set_mode",
    "erlfill_env_step_fused_max_envs",
    "erlfill_actor_forward", "erlfill_actor_forward_cond", "erlfill_replay_insert_cond", "erlfill_emotion_update", "erlfill_episode_end", "erlfill_emotion_forward",
    "erlfill_emotion_tc_image_bytes", "erlfill_emotion_tc_pack", "erlfill_emotion_forward_tc", "erlfill_emotion_forward_tc_train",
    "erlfill_replay_insert", "erlfill_replay_sample", "erlfill_replay_gather", "erlfill_replay_sample_gather",
    "erlfill_replay_sample_gather_blocks",
    "erlfill_ddpg_workspace_bytes", "erlfill_ddpg_grad_buckets", "erlfill_ddpg_set_step", "erlfill_ddpg_update", "erlfill_ddpg_pack_images",
    "erlfill_ddpg_exchange_buffers", "erlfill_ddpg_update_launches", "erlfill_ddpg_peer_launches", "erlfill_peer_ctrl_bytes", "erlfill_peer_alloc", "erlfill_peer_open", "erlfill_peer_close",
    "erlfill_peer_free", "erlfill_ddpg_peer_allreduce", "erlfill_ddpg_peer_allreduce_fast", "erlfill_peer_error",
    "erlfill_t1_workspace_bytes", "erlfill_t1_act_workspace_bytes", "erlfill_t1_grad_buckets", "erlfill_td3_act",
    "erlfill_td3_update", "erlfill_sac_act", "erlfill_sac_update",
    "erlfill_esac_workspace_bytes", "erlfill_esac_act_workspace_bytes", "erlfill_esac_grad_buckets", "erlfill_esac_act",
    "erlfill_esac_update",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ErlfillError(
                f"{SO_PATH} is missing: build it with `python erl-fill_b200/build.py` "
                "(there is no CPU fallback for the ERL-Fill hot path)")
        L = C.CDLL(SO_PATH)
        L.erlfill_last_cuda_error_string.restype = C.c_char_p
        L.erlfill_ddpg_workspace_bytes.restype = C.c_size_t
        L.erlfill_emotion_tc_image_bytes.restype = C.c_size_t
        L.erlfill_t1_workspace_bytes.restype = C.c_size_t
        L.erlfill_t1_act_workspace_bytes.restype = C.c_size_t
        L.erlfill_peer_ctrl_bytes.restype = C.c_size_t
        L.erlfill_esac_workspace_bytes.restype = C.c_size_t
        L.erlfill_esac_act_workspace_bytes.restype = C.c_size_t
        for name in EXPORTS:
            getattr(L, name)
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        L = lib()
        detail = ""
        if rc == -2:
            detail = f" (CUDA error {L.erlfill_last_cuda_error()}: {L.erlfill_last_cuda_error_string().decode()})"
        raise ErlfillError(f"{what} failed with status {rc}{detail}")


def require_device():
    """Raise unless the current CUDA device can run the sm_100a kernels."""
    import torch
    if not torch.cuda.is_available():
        raise ErlfillError("no CUDA device: the ERL-Fill B200 path has no CPU fallback")
    if not lib().erlfill_device_ok():
        raise ErlfillError("current CUDA device is not sm_100 (B200); liberlfill.so is built for sm_100a only")


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_condition(target_weight, target_weight_err, target_time, bounds):
    c = Condition()
    c.target_weight = int(target_weight)
    c.target_weight_err = float(target_weight_err)
    c.target_time = float(target_time)
    for i, k in enumerate(ACTION_ORDER):
        lo, hi = bounds[k]
        c.lo[i], c.hi[i] = float(lo), float(hi)
    return c
