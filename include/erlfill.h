/*
 * erlfill.h — C ABI of liberlfill.so, the B200 (sm_100a) implementation of the ERL-Fill
 * rollout-and-update hot path.
 *
 * The reference (Harrison-Ma/ERL-Fill) has no FFI: its boundary is Python duck typing between
 * train_ddpg()/run_experiment_*() and the env/agent objects.  Every entry
 * point below therefore cites the reference *Python function* whose arithmetic it replaces; the
 * Python drop-in classes in erl-fill_b200/ bind these with ctypes and present the reference's
 * own method names (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  Pointers named d_* are DEVICE pointers on the
 *     current CUDA device; everything else is host memory read during the call.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing
 *     and returns 0 on success or a negative erlfill_status.  Errors never throw.
 *   - SoA layout: per-env arrays are indexed [field][env]; "row" arrays are [env][k].
 *   - all kernels are compiled for sm_100a only; on a machine without such a device every
 *     compute entry returns ERLFILL_ERR_CUDA (there is no CPU fallback).
 */
#ifndef ERLFILL_H_
#define ERLFILL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERLFILL_VERSION 100          /* 0.1.0 */
#define ERLFILL_ACTION_DIM 10
#define ERLFILL_STATE_DIM 2
#define ERLFILL_EMOTION_DIM 3
#define ERLFILL_RING 20              /* deque(maxlen=int(100/5)): MutiConditionEnv.py:43-45 */
#define ERLFILL_HIST 20              /* EmotionTransformer max_len: EmotionModule.py:9,43 */
#define ERLFILL_MAX_ITER 5000        /* VirtualWeightController.py:158 */
#define ERLFILL_NOCHANGE_LIMIT 500   /* VirtualWeightController.py:159 */

typedef enum {
    ERLFILL_OK = 0,
    ERLFILL_ERR_ARG = -1,            /* null pointer / bad size */
    ERLFILL_ERR_CUDA = -2,           /* CUDA runtime error (erlfill_last_cuda_error() has the code) */
    ERLFILL_ERR_UNSUPPORTED = -3
} erlfill_status;

typedef enum { ERLFILL_ENV_MULTI = 0,   /* MutiConditionEnv.WeightEnv.step reward (MutiConditionEnv.py:376-436) */
               ERLFILL_ENV_SINGLE = 1   /* WeightEnv.step reward v2 (WeightEnv.py:480-556) */
} erlfill_env_kind;

/* One operating condition (experiment_runner.py:417-478).  bounds are in ACTION order:
 * fast_weight medium_weight slow_weight fast_opening medium_opening slow_opening fast_delay
 * medium_delay slow_delay unload_delay (MutiConditionEnv.py:350-362). */
typedef struct {
    int32_t target_weight;
    int32_t _pad;
    double target_weight_err;
    double target_time;
    float lo[ERLFILL_ACTION_DIM];
    float hi[ERLFILL_ACTION_DIM];
} erlfill_condition;

/* Per-env persistent simulator/env state, SoA, all device pointers (caller-owned). */
typedef struct {
    int32_t n_env;
    int32_t kind;                    /* erlfill_env_kind */
    int32_t max_steps;               /* env.max_steps (assigned by trainers, ddpg_agent.py:598) */
    int32_t n_cond;
    uint32_t env_id_base;            /* global id of env 0 of this shard (rank * n_env) */
    uint32_t _pad;
    uint64_t seed;                   /* Philox key */
    const erlfill_condition *d_cond; /* [n_cond] */
    const int32_t *d_cond_id;        /* [n_env] or NULL => (env_id_base + i) % n_cond */
    int32_t *d_step_counter;         /* [n_env] */
    int32_t *d_episode_counter;      /* [n_env] */
    int32_t *d_ring_len;             /* [n_env] */
    int32_t *d_ring_head;            /* [n_env] index of the oldest entry */
    double *d_ring;                  /* [ERLFILL_RING][n_env] recent weight errors */
    float *d_obs;                    /* [n_env][2] state the agent acts on (in/out) */
} erlfill_env_batch;

/* Outputs of one env step, device pointers; any pointer except d_next_state/d_reward/d_done may be NULL. */
typedef struct {
    float *d_next_state;             /* [n_env][2] terminal observation of this step (for replay) */
    float *d_reward;                 /* [n_env] */
    uint8_t *d_done;                 /* [n_env] */
    double *d_final_weight;          /* [n_env] info["final_weight"] */
    double *d_total_time;            /* [n_env] info["total_time"] */
    int32_t *d_ticks;                /* [n_env] simulator loop iterations (noise draws consumed) */
    unsigned long long *d_tick_sum;  /* [1] += sum of ticks over the batch (atomic), or NULL */
} erlfill_step_out;

/* Optional injected noise (parity mode against the reference's Mersenne-Twister stream).
 * d_u:[rows][n_env] binary64: row 0 = system_delay ~ U(15,50) (VirtualWeightController.py:137),
 * row 1+t = uniform(-1,1) of tick t (:122).  Rows beyond `rows` read as 0. */
typedef struct {
    const double *d_u;
    int32_t rows;
    int32_t _pad;
    const double *d_reset_u;         /* [2][n_env] uniforms in [0,1) for the cold-start reset, or NULL */
} erlfill_noise;

int erlfill_version(void);
int erlfill_last_cuda_error(void);
const char *erlfill_last_cuda_error_string(void);
/* sm_100a device present and usable on the current device: 1/0 */
int erlfill_device_ok(void);

/* E4 — WeightEnv.reset cold-start branch (MutiConditionEnv.py:135-164, WeightEnv.py:113-148) for
 * every env whose d_mask[i] != 0 (d_mask == NULL: all).  Philox lanes: (env, step_idx, 0, MISC)[1..2]. */
int erlfill_env_reset(const erlfill_env_batch *env, const uint8_t *d_mask, uint32_t step_idx,
                      const erlfill_noise *noise, void *stream);

/* E4 — replay branch: state = mean of 10 stored states, normalised again (MutiConditionEnv.py:145-154).
 * d_replay_states:[capacity][2]; d_idx:[n_env][10] int64 indices into it. */
int erlfill_env_reset_from_replay(const erlfill_env_batch *env, const uint8_t *d_mask,
                                  const float *d_replay_states, const int64_t *d_idx, void *stream);

/* E1+E2/E3 (+E4 when auto_reset) — one filling cycle per env:
 * clip (MutiConditionEnv.py:365-367) -> int() truncation (VirtualWeightController.py:71-83) ->
 * simulate_run (:127-244) -> reward/state/done (MutiConditionEnv.py:376-452 | WeightEnv.py:480-572).
 * d_actions:[n_env][10] binary32 physical units.  On return env->d_obs holds the next observation
 * (the reset state where done && auto_reset). */
int erlfill_env_step(const erlfill_env_batch *env, const float *d_actions, uint32_t step_idx,
                     int auto_reset, const erlfill_noise *noise, const erlfill_step_out *out,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ERLFILL_H_ */
