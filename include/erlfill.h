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

#include <stddef.h>
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
#define ERLFILL_REC 20               /* floats per replay record (19 used + 1 pad = 80 B) */

typedef enum {
    ERLFILL_OK = 0,
    ERLFILL_ERR_ARG = -1,            /* null pointer / bad size */
    ERLFILL_ERR_CUDA = -2,           /* CUDA runtime error (erlfill_last_cuda_error() has the code) */
    ERLFILL_ERR_UNSUPPORTED = -3
} erlfill_status;

typedef enum { ERLFILL_ENV_MULTI = 0,   /* MutiConditionEnv.WeightEnv.step reward (MutiConditionEnv.py:376-436) */
               ERLFILL_ENV_SINGLE = 1,  /* WeightEnv.step reward v2 (WeightEnv.py:480-556) */
               /* OR into the kind: per tick the env classes' "on-board" models instead of the offline
                * controller's — motor_simulation with the speed cap `if current_speed < 10000` and weight_simulation with
                * `int(current_weight + opening * (1 + u))` (WeightEnv.py:176-229, MutiConditionEnv.py:184-226) — under the same
                * 3-stage threshold logic, first-tick delay and stop rules as E1.  One thread per env (no integer hold phase:
                * the weight is truncated every tick); the fused host-buffer path does not apply. */
               ERLFILL_ENV_ONBOARD = 2
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
 * d_replay: rows of `row_stride` floats whose first two are the stored state (row_stride = 2 for a plain
 * [n][2] array, ERLFILL_REC for the replay buffer).  d_idx:[n_env][10] int64 row indices (the N=1 drop-in
 * passes random.sample's), or NULL: 10 Philox draws floor(U*n_valid) with replacement per env.
 * update_counters = 0 when the env-step kernel's auto-reset already reset the episode counters. */
int erlfill_env_reset_from_replay(const erlfill_env_batch *env, const uint8_t *d_mask, const float *d_replay,
                                  int row_stride, int64_t n_valid, const int64_t *d_idx, uint32_t step_idx,
                                  int update_counters, void *stream);

/* E1+E2/E3 (+E4 when auto_reset) — one filling cycle per env:
 * clip (MutiConditionEnv.py:365-367) -> int() truncation (VirtualWeightController.py:71-83) ->
 * simulate_run (:127-244) -> reward/state/done (MutiConditionEnv.py:376-452 | WeightEnv.py:480-572).
 * d_actions:[n_env][10] binary32 physical units.  On return env->d_obs holds the next observation
 * (the reset state where done && auto_reset). */
int erlfill_env_step(const erlfill_env_batch *env, const float *d_actions, uint32_t step_idx,
                     int auto_reset, const erlfill_noise *noise, const erlfill_step_out *out,
                     void *stream);

/* The same step with the lock-step index read from device memory (d_step_idx[0], incremented by one after the step): the call
 * sequence then holds no per-step constant, so [H2D actions, step, D2H results] can be captured once in a CUDA graph and
 * replayed for every step.  d_step_idx points at TWO words: {index, ticket}; the ticket word must be zero before the first
 * call and belongs to the library (the fused simulator kernel lets its last env advance the index, no extra launch). */
int erlfill_env_step_counted(const erlfill_env_batch *env, const float *d_actions, uint32_t *d_step_idx,
                             int auto_reset, const erlfill_noise *noise, const erlfill_step_out *out,
                             void *stream);

/* Which simulator kernel erlfill_env_step launches for Philox noise (injected noise always uses the
 * thread-per-env binary64 kernel).  All produce bit-identical results; AUTO = warp per env.  The last two are
 * the warp kernel with its multi-window hold runs switched off / computed and then rejected: they exist so the
 * tests can pin the window-by-window path and the redo path of a rejected run against the oracle. */
typedef enum { ERLFILL_SIM_AUTO = 0, ERLFILL_SIM_THREAD_PER_ENV = 1, ERLFILL_SIM_WARP_PER_ENV = 2,
               ERLFILL_SIM_WARP_WINDOWS = 3, ERLFILL_SIM_WARP_REJECT_RUNS = 4 } erlfill_sim_mode;
int erlfill_env_step_set_mode(int mode);

/* Largest batch the warp-per-env simulator runs as ONE kernel (every env resident at once, the per-env epilogue fused
 * into it): its outputs are then written exactly once, so a host caller may point d_actions / d_next_state / d_reward /
 * d_done at pinned, device-mapped host memory and skip the copies (VecFillEnv.step_host does).  Larger batches use two
 * kernels that pass intermediate values through the output slots: keep those in device memory. */
int erlfill_env_step_fused_max_envs(void);

/* ------------------------------------------------------------------------------------------ */
/* Policy side of the rollout                                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef enum { ERLFILL_DROPOUT_OFF = 0,     /* module.eval() */
               ERLFILL_DROPOUT_PHILOX = 1,  /* train() mode, keep-masks from Philox (row, call_idx, col/4, DROPOUT+layer) */
               ERLFILL_DROPOUT_MASK = 2     /* train() mode, keep-masks injected by the caller (parity tests) */
} erlfill_dropout_mode;

/* Actor parameters, device pointers, each in the reference state_dict layout [out][in]
 * (ddpg_agent.py:31-56): input_layer, hidden.2, hidden.5, hidden.7, emotion_weights[3][10],
 * buffers scale_params[10] / offset_params[10]. */
typedef struct {
    const float *w_in, *b_in;      /* [128][5], [128] */
    const float *w_h2, *b_h2;      /* [128][128], [128] */
    const float *w_h5, *b_h5;      /* [64][128], [64] */
    const float *w_h7, *b_h7;      /* [10][64], [10] */
    const float *emotion_weights;  /* [3][10] */
    const float *scale, *offset;   /* [10], [10] */
} erlfill_actor_weights;

/* OU exploration state, one process per env (ddpg_agent.py:735-785). */
typedef struct {
    float *d_state;                /* [n][10] in/out */
    const float *d_sigma;          /* [n] current sigma as binary32 (sigma * randn is a binary32 product) */
    const float *d_randn;          /* [n][10] injected N(0,1) draws, or NULL => Philox Box-Muller (STREAM_OU) */
    float *d_final_action;         /* [n][10] clip(scaled + noise*scale, offset-scale, offset+scale) */
    uint64_t seed;
    uint32_t env_id_base, step_idx;
} erlfill_ou;

/* Per-env EmotionModule state (EmotionModule.py:105-162): current_emotion (binary64) and the 20-deep
 * history of binary32 vectors the transformer reads. */
typedef struct {
    double *d_current;             /* [3][n] */
    float *d_history;              /* [20][3][n] ring */
    int32_t *d_hist_len;           /* [n] */
    int32_t *d_hist_head;          /* [n] slot of the oldest entry */
} erlfill_emotion_state;

/* Per-condition action scaling for a batch that mixes operating conditions.  The reference builds ONE AGENT PER CONDITION
 * (experiment_runner.py:504-510), each with scale_params / offset_params from its own env.bounds (ddpg_agent.py:56-71), so the
 * network output of an env under condition c is mapped with row c of these tables; a batch of one condition equals the
 * plain entry points.  A replay record written through the _cond insert carries its condition in column 19, which
 * erlfill_ddpg_update reads back when hyper->n_cond > 1 (nets->scale / nets->offset are then the [n_cond][10] tables). */
typedef struct {
    const float *d_scale;      /* [n_cond][10] */
    const float *d_offset;     /* [n_cond][10] */
    const int32_t *d_cond_id;  /* [n_rows], or NULL => (row_id_base + row) % n_cond (erlfill_env_batch's round-robin rule) */
    int32_t n_cond;
    uint32_t row_id_base;
} erlfill_cond_scaling;

/* A1 (+A2 when ou != NULL) — Actor.forward (ddpg_agent.py:73-122); with ou also OUNoise.sample and
 * DDPGAgent.act's clip (:348-356, :768-778).  d_emotion is [n][3], or [3] when emotion_broadcast. */
int erlfill_actor_forward(const erlfill_actor_weights *w, int n_rows, const float *d_state, const float *d_emotion,
                          int emotion_broadcast, int dropout_mode, uint64_t seed, uint32_t call_idx,
                          uint32_t row_id_base, const uint8_t *d_mask1, const uint8_t *d_mask2,
                          float *d_scaled_action, float *d_base_action, const erlfill_ou *ou, void *stream);

/* The same with per-condition scaling (cs == NULL: identical to erlfill_actor_forward). */
int erlfill_actor_forward_cond(const erlfill_actor_weights *w, int n_rows, const float *d_state, const float *d_emotion,
                               int emotion_broadcast, int dropout_mode, uint64_t seed, uint32_t call_idx,
                               uint32_t row_id_base, const uint8_t *d_mask1, const uint8_t *d_mask2,
                               float *d_scaled_action, float *d_base_action, const erlfill_ou *ou,
                               const erlfill_cond_scaling *cs, void *stream);

/* M1 — EmotionModule.update(reward, next_state - state) (EmotionModule.py:121-162) for every env. */
int erlfill_emotion_update(int n_env, const float *d_reward, const float *d_state, const float *d_next_state,
                           const erlfill_emotion_state *emo, void *stream);

/* train_ddpg's per-episode `ou_noise.reset(); ou_noise.anneal()` (ddpg_agent.py:633-634) for envs with done.
 * d_episode_count (int64[1], may be NULL) += number of envs with done: the batched stand-in for the trainer's
 * `agent.train_step += 1` per episode (ddpg_agent.py:729). */
int erlfill_episode_end(int n_env, const uint8_t *d_done, float *d_ou_state, float *d_sigma, double *d_sigma64,
                        long long *d_episode_count, void *stream);

/* EmotionTransformer parameters (EmotionModule.py:27-40), device pointers, per-layer tensors stacked on a
 * leading [4] axis, each slice in the reference state_dict layout. */
typedef struct {
    const float *emb_w, *emb_b;            /* embedding [32][3], [32] */
    const float *pos;                      /* positional_encoding [20][32] */
    const float *in_proj_w, *in_proj_b;    /* [4][96][32], [4][96] */
    const float *out_proj_w, *out_proj_b;  /* [4][32][32], [4][32] */
    const float *lin1_w, *lin1_b;          /* [4][2048][32], [4][2048] */
    const float *lin2_w, *lin2_b;          /* [4][32][2048], [4][32] */
    const float *norm1_w, *norm1_b;        /* [4][32] */
    const float *norm2_w, *norm2_b;        /* [4][32] */
    const float *fc_w, *fc_b;              /* fc_out [3][32], [3] */
} erlfill_transformer_weights;

/* Injected keep-masks for ERLFILL_DROPOUT_MASK (uint8 0/1), in at::dropout's draw order per layer:
 * attention probabilities, dropout1, FFN hidden dropout, dropout2. */
typedef struct {
    const uint8_t *attn;                   /* [4][n][4][S][S] */
    const uint8_t *d1;                     /* [4][n][S][32] */
    const uint8_t *ff;                     /* [4][n][S][2048] */
    const uint8_t *d2;                     /* [4][n][S][32] */
} erlfill_transformer_masks;

/* M2 — EmotionModule.get_emotion = EmotionTransformer.forward(get_state()) (EmotionModule.py:48-71, 83-102,
 * 164-175), fp32.  Sequence source: d_seq [n][seq_len][3] if non-NULL, else the history ring in `emo`
 * (front-padded with the oldest entry to 20; the neutral token [0.5,0.5,0] alone when empty).  d_out [n][3]. */
int erlfill_emotion_forward(const erlfill_transformer_weights *w, int n_env, const float *d_seq, int seq_len,
                            const erlfill_emotion_state *emo, int dropout_mode, uint64_t seed, uint32_t call_idx,
                            uint32_t row_id_base, const erlfill_transformer_masks *masks, float *d_out, void *stream);

/* M2 on the tensor cores (tcgen05.mma, bf16 operands, fp32 accumulation in tensor memory; attention, LayerNorm and
 * the residual stream stay fp32), dropout off (module.eval()): the batched rollout's get_emotion().  Same
 * sequence sources as erlfill_emotion_forward.  The parameters are packed once per weight change into a device
 * image of erlfill_emotion_tc_image_bytes() bytes (16-byte aligned, caller-owned).  Stated tolerance against the
 * fp32 path: 1e-2 absolute on the [0,1] outputs (tests/test_emotion_tc_gpu.py; measured ~2e-3). */
size_t erlfill_emotion_tc_image_bytes(void);
int erlfill_emotion_tc_pack(const erlfill_transformer_weights *w, void *d_image, void *stream);
int erlfill_emotion_forward_tc(const void *d_image, int n_env, const float *d_seq, int seq_len,
                               const erlfill_emotion_state *emo, float *d_out, void *stream);
/* The same in train() mode — what the reference actually runs: it never calls .eval() on the transformer (EmotionModule.py:9,35), so
 * the four dropout layers of every encoder layer (attention probabilities, dropout1, FFN hidden, dropout2; p = 0.3) are active in
 * every get_emotion().  dropout_mode / seed / call_idx / row_id_base / masks as in erlfill_emotion_forward: PHILOX draws the SAME
 * masks as the fp32 kernel (csrc/tf_dropout.cuh), MASK takes the caller's (d_seq required).  n_env < 2^24. */
int erlfill_emotion_forward_tc_train(const void *d_image, int n_env, const float *d_seq, int seq_len,
                                     const erlfill_emotion_state *emo, int dropout_mode, uint64_t seed, uint32_t call_idx,
                                     uint32_t row_id_base, const erlfill_transformer_masks *masks, float *d_out, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* Replay buffer: [capacity][ERLFILL_REC] binary32 rows                                        */
/*   [0:2] state [2:12] normalised action [12] reward [13:15] next_state [15] done [16:19] emotion [19] condition index (0.0f unless
 *   written by erlfill_replay_insert_cond) */
/* ------------------------------------------------------------------------------------------ */

/* R1 — DDPGAgent.remember + PrioritizedReplayBuffer.add (ddpg_agent.py:359-395, 206-226) for n transitions.
 * Row of transition i: d_pos[i] if d_pos != NULL, else (start + i) % capacity.  When d_scale/d_offset are
 * given the physical action is normalised clip((a-offset)/scale, -1, 1) (ddpg_agent.py:373-375); pass NULL
 * for both to store d_action as is.  (The TD-error priority of the reference is NaN by construction and
 * every priority is 1e-4, SURVEY.md section 0.7, so no priority array exists.) */
int erlfill_replay_insert(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                          const float *d_state, const float *d_action, const float *d_reward,
                          const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                          const float *d_scale, const float *d_offset, void *stream);

/* R1 for a mixed-condition batch: transition i is normalised with the scale / offset row of its own condition, and the
 * condition index is stored (as a float) in column 19 of the record. */
int erlfill_replay_insert_cond(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                               const float *d_state, const float *d_action, const float *d_reward,
                               const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                               const erlfill_cond_scaling *cs, void *stream);

/* R2 — PrioritizedReplayBuffer.sample's index draw (ddpg_agent.py:228-254) with equal priorities:
 * idx[i] = floor(U_i * n_valid), U from Philox (i/4, update_idx, stream_id, SAMPLE)[i%4]. */
int erlfill_replay_sample(int batch, int64_t n_valid, uint64_t seed, uint32_t update_idx, uint32_t stream_id,
                          int64_t *d_idx, void *stream);

/* R2 — minibatch assembly (ddpg_agent.py:434-442): d_out[b][:] = d_buf[d_idx[b]][:]. */
int erlfill_replay_gather(int batch, const float *d_buf, const int64_t *d_idx, float *d_out, void *stream);

/* R2 in one launch: the index draw of erlfill_replay_sample followed by the assembly of erlfill_replay_gather (same indices, same rows).
 * d_idx (optional) receives the indices.  d_stats_partials (optional): [erlfill_replay_sample_gather_blocks(batch)][8] floats, per block
 * of 64 samples the sums  sum(e_c - 0.5) | sum((e_c - 0.5)^2)  of the three emotion columns — the by-product that lets
 * erlfill_ddpg_update skip its own pass over the minibatch for the batch statistics of Critic.forward (ddpg_agent.py:176). */
int erlfill_replay_sample_gather_blocks(int batch);
int erlfill_replay_sample_gather(int batch, int64_t n_valid, uint64_t seed, uint32_t update_idx, uint32_t stream_id, const float *d_buf,
                                 int64_t *d_idx, float *d_out, float *d_stats_partials, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* ER-DDPG update                                                                              */
/* ------------------------------------------------------------------------------------------ */
/* Flat fp32 parameter buffers in the reference's parameters() order.
 * actor (26,216): emotion_weights[3][10] | input_layer.weight[128][5] .bias[128] | hidden.2.weight[128][128] .bias |
 *                 hidden.5.weight[64][128] .bias | hidden.7.weight[10][64] .bias
 * critic (103,937): net.0.weight[256][15] .bias | net.1 (LayerNorm) weight, bias | net.4.weight[256][256] .bias |
 *                 net.5 weight, bias | net.8.weight[128][256] .bias | net.10.weight[1][128] .bias */
#define ERLFILL_ACTOR_PARAMS 26216
#define ERLFILL_CRITIC_PARAMS 103937
typedef struct {
    float *actor, *actor_target, *critic, *critic_target;
    float *actor_m, *actor_v, *critic_m, *critic_v;    /* Adam moments */
    const float *scale, *offset;                       /* actor scale_params / offset_params [10] */
} erlfill_ddpg_nets;

/* Injected keep-masks (uint8 0/1) of the ten Dropout(0.2) layers one learn() evaluates, in at::dropout's draw order:
 * keep[0,1] actor_target [B][128], keep[2,3] critic_target [B][256], keep[4,5] critic(s, a) [B][256], keep[6,7] actor [B][128],
 * keep[8,9] critic(s, actor(s)) [B][256]   (ddpg_agent.py:40,43,150,155 through :468-483). */
typedef struct { const uint8_t *keep[10]; } erlfill_ddpg_masks;

typedef enum { ERLFILL_PRECISION_EXACT = 0,  /* 3xTF32 tensor-core GEMMs (mma.sync), fp32-level results: the 1e-4 parity path */
               ERLFILL_PRECISION_FAST = 1    /* tcgen05 bf16 operands, fp32 accumulation in tensor memory, fp32 master weights / LayerNorm /
                                              * optimiser; first layers hi/lo split (physical-unit inputs).  Six launches per update.  Needs
                                              * erlfill_ddpg_pack_images after every external change of the parameters. */
} erlfill_precision;

typedef struct {
    int32_t batch;
    int32_t adam_step;        /* 1-based optimiser step of this update; 0 = scalars already uploaded by erlfill_ddpg_set_step */
    float gamma, tau, max_grad_norm;
    int32_t n_cond;           /* <= 1: nets->scale / offset are [10]; > 1: [n_cond][10] tables indexed by column 19 of each record */
    double critic_lr, actor_lr;   /* base rates (1e-3, 1e-4) */
    double lr_decay;              /* lr_decay_rate ** train_step (ddpg_agent.py:455) */
    int32_t precision;            /* erlfill_precision */
    int32_t dropout_mode;         /* erlfill_dropout_mode: OFF = module.eval(); PHILOX / MASK = train() mode as the reference runs it */
    uint64_t seed;                /* Philox key of the dropout masks */
    uint32_t update_idx, rank;    /* Philox counter words: update index, data-parallel rank (row id = rank << 20 | minibatch row) */
    const erlfill_ddpg_masks *masks;   /* ERLFILL_DROPOUT_MASK only */
    /* optional (ERLFILL_PRECISION_FAST): the emotion sums erlfill_replay_sample_gather left for THIS minibatch, n_stats_partials rows of 8 */
    const float *d_stats_partials;
    int32_t n_stats_partials, _pad2;
    /* optional (ERLFILL_PRECISION_FAST, data-parallel): the peer group (erlfill_peer_group, below).  The X_GRAD phases then raise this
     * rank's "bucket ready" flags themselves and the exchange is erlfill_ddpg_peer_allreduce_fast. */
    const void *peer;
} erlfill_ddpg_hyper;

typedef enum { ERLFILL_PHASE_CRITIC_GRAD = 1, ERLFILL_PHASE_CRITIC_APPLY = 2, ERLFILL_PHASE_ACTOR_GRAD = 4,
               ERLFILL_PHASE_ACTOR_APPLY = 8, ERLFILL_PHASE_ALL = 15,
               /* with *_APPLY: take the gradient (and its sum-of-squares partials) that erlfill_ddpg_peer_allreduce left in
                * the reduced buffers instead of the local bucket */
               ERLFILL_PHASE_REDUCED = 16 } erlfill_ddpg_phase;

/* bytes of the caller-provided device workspace for a given minibatch size */
size_t erlfill_ddpg_workspace_bytes(int batch);
/* device addresses/lengths of the two flat gradient buckets inside the workspace (data-parallel: all-reduce
 * (average) the critic bucket between CRITIC_GRAD and CRITIC_APPLY and the actor bucket between ACTOR_GRAD and
 * ACTOR_APPLY; the critic bucket carries the 3 batch emotion means that set the learning rates in its tail). */
int erlfill_ddpg_grad_buckets(void *d_workspace, int batch, float **critic_grad, int *critic_len,
                              float **actor_grad, int *actor_len);
/* Uploads the per-update scalars (Adam bias corrections of optimiser step `adam_step` >= 1, lr_decay) into the workspace.
 * erlfill_ddpg_update does this itself when hyper->adam_step >= 1; call it separately and pass hyper->adam_step = 0 to
 * replay one captured CUDA graph of the update for every step (the graph then contains no per-step constants). */
int erlfill_ddpg_set_step(void *d_workspace, int batch, int adam_step, double lr_decay, void *stream);
/* (Re)builds the bf16 weight images of the four networks inside the workspace from the fp32 parameters (ERLFILL_PRECISION_FAST).
 * Call once after construction and after every external change of nets->* (checkpoint load, load_weights); the update keeps
 * the images current by itself. */
int erlfill_ddpg_pack_images(const erlfill_ddpg_nets *nets, void *d_workspace, int batch, void *stream);
/* U1-U5 — DDPGAgent.learn (ddpg_agent.py:422-512).  d_batch [B][ERLFILL_REC] gathered records,
 * d_next_emotion [3] the agent's fresh emotion (:466).  d_report (optional) [4] = critic_loss, actor_loss,
 * critic_lr, actor_lr. */
int erlfill_ddpg_update(const erlfill_ddpg_nets *nets, const erlfill_ddpg_hyper *hyper, const float *d_batch,
                        const float *d_next_emotion, void *d_workspace, int phases, float *d_report, void *stream);

/* kernels one erlfill_ddpg_update(ERLFILL_PHASE_ALL) launches (fast: the tcgen05 bf16 path; else the 3xTF32 path) */
int erlfill_ddpg_update_launches(int fast);

/* Exchange buffers inside the workspace: the two gradient buckets (as erlfill_ddpg_grad_buckets), the two reduced
 * buffers of the same lengths and the [2][64] sum-of-squares partials. */
int erlfill_ddpg_exchange_buffers(void *d_workspace, int batch, float **critic_grad, int *critic_len, float **actor_grad,
                                  int *actor_len, float **critic_reduced, float **actor_reduced, float **sumsq);

/* ------------------------------------------------------------------------------------------ */
/* Data-parallel gradient exchange over NVLink peer memory            */
/* ------------------------------------------------------------------------------------------ */
/* The reference has no multi-GPU path (nothing to mirror).  One process per GPU; each rank puts its update workspace
 * (+ erlfill_peer_ctrl_bytes() of control block behind it) in a block from erlfill_peer_alloc, ships the 64-byte IPC handle
 * to the other ranks of the node (any transport: torch.distributed.all_gather_object), and maps theirs with
 * erlfill_peer_open.  These four are the only entry points that allocate or map memory. */
#define ERLFILL_MAX_PEERS 8
typedef struct {
    int32_t world, rank;
    void *ws[ERLFILL_MAX_PEERS];     /* update workspace of every rank as mapped in THIS process (own entry: the local block) */
    void *ctrl[ERLFILL_MAX_PEERS];   /* control block of every rank (= ws + workspace bytes rounded up to 256) */
} erlfill_peer_group;
size_t erlfill_peer_ctrl_bytes(void);
int erlfill_peer_alloc(size_t bytes, void **dptr, unsigned char handle[64]);          /* cudaMalloc + zero + IPC handle */
int erlfill_peer_open(int peer_device, const unsigned char handle[64], void **mapped); /* peer access + cudaIpcOpenMemHandle */
int erlfill_peer_close(void *mapped);
int erlfill_peer_free(void *dptr);
/* Averages bucket `which` (0 critic incl. the 3 emotion means, 1 actor) over the ranks in ONE pass over peer memory, in rank
 * order (bit-identical on every rank), into the reduced buffer, and writes the sum-of-squares partials of the reduced
 * gradient.  Call between X_GRAD and (X_APPLY | ERLFILL_PHASE_REDUCED); stream-ordered, graph-capturable, no host sync.
 * Every rank must make the same sequence of calls. */
int erlfill_ddpg_peer_allreduce(const erlfill_peer_group *g, int batch, int which, void *stream);
/* The exchange of the tcgen05 update (hyper->precision == ERLFILL_PRECISION_FAST with hyper->peer set): the X_GRAD phase has already
 * signalled, so this is one launch — wait for the peers' flags, one element per thread, one norm partial per 256 elements (the
 * geometry the fast path's X_APPLY | ERLFILL_PHASE_REDUCED expects). */
int erlfill_ddpg_peer_allreduce_fast(const erlfill_peer_group *g, int batch, int which, void *stream);
/* kernels the two exchanges of one update add (bench.py's gpu_launches claim): 4 (fast = 0: signal + reduce, twice) or 2 */
int erlfill_ddpg_peer_launches(int fast);
/* 1 in *error_out if a wait for a peer timed out (ERLFILL_PEER_TIMEOUT_S seconds, default 30) since the allocation; synchronises
 * the device.  A timed-out exchange also writes NaN into the reduced bucket and its norm partials, so the following APPLY phase
 * poisons this rank's parameters and update report instead of applying a stale sum. */
int erlfill_peer_error(const erlfill_peer_group *g, int *error_out);

/* ------------------------------------------------------------------------------------------ */
/* TD3 and SAC (BASELINE.json configs[4])                         */
/* ------------------------------------------------------------------------------------------ */
/* Flat fp32 parameter buffers in the reference's parameters() order; every MLP is  W0[256][in] b0 | W1[256][256] b1 | W2[out][256] b2.
 *   TD3 actor  (69,130): TD3Actor.net.{0,2,4}            2-256-256-10               (td3_agent.py:51-85)
 *   TD3 critic (138,754): TD3Critic.q1.{0,2,4} | q2.{0,2,4}   12-256-256-1 each     (td3_agent.py:88-133)
 *   SAC policy (71,700): GaussianPolicy.net.{0,2} | mean | log_std                  (sac_agent.py:14-84)
 *   SAC q      (138,754): q1.q.{0,2,4} | q2.q.{0,2,4}                               (sac_agent.py:87-112) */
#define ERLFILL_TD3_ACTOR_PARAMS 69130
#define ERLFILL_TD3_CRITIC_PARAMS 138754
#define ERLFILL_SAC_POLICY_PARAMS 71700
#define ERLFILL_SAC_Q_PARAMS 138754
typedef struct {
    float *actor, *actor_target, *critic, *critic_target;
    float *actor_m, *actor_v, *critic_m, *critic_v;
} erlfill_td3_nets;
typedef struct {
    int32_t batch;
    int32_t critic_step, actor_step;   /* 1-based Adam steps of the two optimisers for this call */
    int32_t update_actor;              /* total_it % policy_delay == 0 (td3_agent.py:287) */
    float gamma, tau, policy_noise, noise_clip;
    double critic_lr, actor_lr;
    uint64_t seed;                     /* Philox key when d_noise == NULL */
    uint32_t update_idx, rank;         /* data-parallel rank: mixed into the Philox counter so ranks draw different update noise */
} erlfill_td3_hyper;
typedef struct {
    float *policy, *q, *q_target;
    float *policy_m, *policy_v, *q_m, *q_v;
} erlfill_sac_nets;
typedef struct {
    int32_t batch, adam_step;
    float gamma, tau, alpha, _pad;
    double lr;
    uint64_t seed;
    uint32_t update_idx, rank;
} erlfill_sac_hyper;

/* bytes of the caller-provided device workspace of erlfill_td3_update / erlfill_sac_update, and of the *_act entries */
size_t erlfill_t1_workspace_bytes(int batch);
size_t erlfill_t1_act_workspace_bytes(void);
/* flat gradient buckets inside the workspace for the data-parallel all-reduce between *_GRAD and *_APPLY
 * (critic: 138,754 floats; actor: 69,130 (TD3) / 71,700 (SAC)) */
int erlfill_t1_grad_buckets(void *d_workspace, int batch, float **critic_grad, float **actor_grad);

/* TD3Agent.select_action (td3_agent.py:221-236): clip(actor(s) + noise_scale * N(0,1), -1, 1), denormalised with the bounds
 * lo/hi (:198-219).  d_randn [n][10] injected draws or NULL (Philox).  d_norm_action (optional) receives the clipped
 * normalised action, i.e. what remember() stores. */
int erlfill_td3_act(const float *d_actor, int n, const float *d_state, const float *d_randn, float noise_scale, const float *d_lo,
                    const float *d_hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx, float *d_action, float *d_norm_action,
                    void *d_workspace, void *stream);
/* TD3Agent.train_step (td3_agent.py:238-303) on gathered replay rows d_batch [B][ERLFILL_REC] (emotion columns unused).
 * d_noise [B][10]: the N(0,1) draws of randn_like (:268) or NULL.  d_report (optional) [3] = critic_loss, mean q1, actor_loss. */
int erlfill_td3_update(const erlfill_td3_nets *nets, const erlfill_td3_hyper *hyper, const float *d_batch, const float *d_noise,
                       void *d_workspace, int phases, float *d_report, void *stream);
/* SACAgent.act (sac_agent.py:186-205): tanh(mean + sigma * eps) (tanh(mean) when deterministic), denormalised (:163-184). */
int erlfill_sac_act(const float *d_policy, int n, const float *d_state, const float *d_eps, int deterministic, const float *d_lo,
                    const float *d_hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx, float *d_action, float *d_norm_action,
                    void *d_workspace, void *stream);
/* SACAgent.update (sac_agent.py:221-283).  d_eps_next / d_eps_new [B][10]: the two rsample() draws (:69) or NULL.
 * d_report (optional) [3] = q1_loss, q2_loss, policy_loss. */
int erlfill_sac_update(const erlfill_sac_nets *nets, const erlfill_sac_hyper *hyper, const float *d_batch, const float *d_eps_next,
                       const float *d_eps_new, void *d_workspace, int phases, float *d_report, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* EmotionSAC: DifferentModules/ERL_Fill_agent.py                */
/* ------------------------------------------------------------------------------------------ */
/* Flat fp32 parameter buffers in the reference's parameters() order.
 *   policy (146,468): EmotionGaussianPolicy — state_encoder.0 [256][2] .bias | state_encoder.1 (LayerNorm) weight, bias |
 *       emotion_encoder.0 [32][3] .bias | emotion_encoder.1 weight, bias | policy_net.0 [256][288] .bias | policy_net.2 [256][256] .bias |
 *       mean_layer [10][256] .bias | log_std_layer [10][256] .bias | emotion_adapter_mean.0 [10][3] .bias | emotion_adapter_std.0 [10][3] .bias
 *   q (2 x 210,369): EmotionAwareQNetwork q1 then q2 — state_encoder.0 [256][2] .bias | .1 | action_encoder.0 [256][10] .bias | .1 |
 *       emotion_encoder.0 [32][3] .bias | .1 | q_net.0 [256][544] .bias | q_net.2 [256][256] .bias | q_net.4 [1][256] .bias      (:63-239) */
#define ERLFILL_ESAC_POLICY_PARAMS 146468
#define ERLFILL_ESAC_Q_PARAMS 210369
typedef erlfill_sac_nets erlfill_esac_nets;      /* policy | q = [q1 | q2] | q_target | Adam moments */
typedef erlfill_sac_hyper erlfill_esac_hyper;    /* gamma 0.99, tau 0.005, alpha 0.3, lr 3e-4 (:262-290) */
size_t erlfill_esac_workspace_bytes(int batch);
size_t erlfill_esac_act_workspace_bytes(void);
/* the two gradient buckets inside the workspace ([q1 | q2] and the policy) for a data-parallel all-reduce between the phases */
int erlfill_esac_grad_buckets(void *d_workspace, int batch, float **q_grad, float **policy_grad);
/* EmotionSACAgent.act (:322-348) for n rows: d_state [n][2], d_emotion [n][3] (or [3] with emotion_broadcast), d_eps [n][10] the
 * rsample() normals (NULL: Philox (row_id_base + row, step_idx)); deterministic: tanh(mean).  Physical action and the
 * normalised action (what remember() stores, :350-363). */
int erlfill_esac_act(const float *d_policy, int n, const float *d_state, const float *d_emotion, int emotion_broadcast, const float *d_eps,
                     int deterministic, const float *d_lo, const float *d_hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx,
                     float *d_action, float *d_norm_action, void *d_workspace, void *stream);
/* EmotionSACAgent.update (:365-448) on gathered records d_batch [B][ERLFILL_REC] (the emotion columns are the stored emotions).
 * d_eps_next / d_eps_new [B][10]: the two rsample() draws (NULL: Philox); d_emotion_now / d_emotion_prev [3]: the operands of the
 * emotion regulariser (reported; it has no gradient).  d_report [4] = q1_loss, q2_loss, policy_loss, emo_reg. */
int erlfill_esac_update(const erlfill_esac_nets *nets, const erlfill_esac_hyper *hyper, const float *d_batch, const float *d_eps_next,
                        const float *d_eps_new, const float *d_emotion_now, const float *d_emotion_prev, void *d_workspace, int phases,
                        float *d_report, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ERLFILL_H_ */
