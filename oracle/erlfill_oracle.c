/*
 * erlfill_oracle.c — CPU restatement (plain C, fp64) of the ERL-Fill rollout hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * the library built from it.  The product path (erl-fill_b200/) never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py)
 * against fixtures produced by running the UNMODIFIED reference in the build container
 * (oracle/gen_golden.py -> tests/golden/ *.npz): bit-exact on tick count, total_time,
 * final_weight (fp64), state (fp32), done; reward to <= 1 fp32 ulp.
 *
 * What is restated (reference file:line, all under /root/reference):
 *   E1  VirtualWeightController.load_params/motor_simulation/weight_simulation/simulate_run
 *       (VirtualWeightController.py:64-83, 85-113, 115-125, 127-244), frozen clock => one
 *       inner iteration per outer iteration; the 10 s wall-clock stop (:230-233) is dropped.
 *   E2  MutiConditionEnv.WeightEnv.step reward/state/done (MutiConditionEnv.py:346-468)
 *   E3  WeightEnv.step reward v2 (WeightEnv.py:480-572)
 *   E4  reset (MutiConditionEnv.py:135-164, WeightEnv.py:113-148)
 *   M1  EmotionModule.update (EmotionModule.py:121-162)
 *   A2  OUNoise.sample/anneal (ddpg_agent.py:762-785) and act()'s clip (:352-356)
 *   R1  action normalisation in remember (ddpg_agent.py:373-375)
 *   N1  the on-board per-tick motor / weight models (WeightEnv.py:176-229, MutiConditionEnv.py:184-226)
 *       under the same stage logic (kind | 2 in the env-step entry points)
 *   R2  uniform-priority index draw: searchsorted(cumsum(p)/last, u, 'right') == floor(u*n)
 *       for equal priorities (ddpg_agent.py:228-254)
 *
 * Arithmetic contract: IEEE-754 binary64 with no FMA contraction (build with
 * -ffp-contract=off) so that the result is identical to CPython floats; the places where
 * the reference (under numpy >= 2 weak-scalar promotion, which is what the fixtures were
 * generated with) computes in binary32 are restated in binary32 and marked [f32].
 *
 * Noise: the reference draws from CPython's Mersenne Twister.  The oracle takes its noise
 * either (a) from caller arrays (parity against the reference: the test regenerates the
 * MT stream with random.Random(seed)) or (b) from the counter-based Philox-4x32-10 stream
 * the CUDA kernels use (parity CUDA <-> oracle at any size; contract in DESIGN.md).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------ */
/* Philox-4x32-10 (Salmon et al., SC'11), the same function the CUDA kernels evaluate.   */
/* counter = (env_global_id, step_index, block, stream), key = (seed_lo, seed_hi)         */
/* ------------------------------------------------------------------------------------ */
#define ORC_STREAM_TICK 0u   /* block b, lane l -> uniform for tick 4*b+l                  */
#define ORC_STREAM_MISC 1u   /* block 0: lane0 system_delay, lane1/2 reset uniforms         */
#define ORC_STREAM_OU   2u   /* Box-Muller normals for OU noise                            */

static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32(uint32_t seed_lo, uint32_t seed_hi, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint32_t out[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = seed_lo, k1 = seed_hi;
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

/* u32 -> U in [0,1) on the 2^-32 grid (exact in binary64) */
static inline double u32_to_unit(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }

/* ------------------------------------------------------------------------------------ */
/* E1: simulator                                                                         */
/* params order == load_params order (VirtualWeightController.py:71-83):                  */
/*   0 target_weight 1 fast_weight 2 medium_weight 3 slow_weight 4 fast_delay             */
/*   5 medium_delay 6 slow_delay 7 fast_opening 8 medium_opening 9 slow_opening           */
/*   10 unload_delay                                                                      */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const double *u;       /* mode (a): per-tick uniform(-1,1) draws, or NULL            */
    int n_u;
    uint32_t seed_lo, seed_hi, env_id, step_idx; /* mode (b)                              */
    uint32_t cache[4];
    int cache_block;
} orc_noise;

static inline double noise_tick(orc_noise *nz, int t /* 0-based tick index */) {
    if (nz->u) return (t < nz->n_u) ? nz->u[t] : 0.0;
    int b = t >> 2;
    if (b != nz->cache_block) {
        orc_philox4x32(nz->seed_lo, nz->seed_hi, nz->env_id, nz->step_idx, (uint32_t)b,
                       ORC_STREAM_TICK, nz->cache);
        nz->cache_block = b;
    }
    /* uniform(-1,1) = -1 + 2*U, exact on this grid */
    return -1.0 + 2.0 * u32_to_unit(nz->cache[t & 3]);
}

/* Returns the number of loop iterations executed (= draws consumed). */
static int simulate_core(const int p[11], double sysdelay, orc_noise *nz, double *total_time,
                         double *final_weight, int *tick_out) {
    const double fast_pos = p[7], medium_pos = p[8], slow_pos = p[9];
    const double fast_w = p[1], medium_w = p[2], slow_w = p[3];
    double w = 0.0, open = 0.0, speed = 0.0, dir = 1.0; /* reset(): :16-25 */
    double target = fast_pos;                             /* :153 */
    int delay_done = 0, tick = 0, iteration = 0, nochg = 0, n = 0;
    double prev = 0.0;                                    /* :163 */
    for (;;) {
        tick += 1;                                        /* :181 */
        /* motor_simulation(target): :92-113 */
        double dt = speed > 0 ? speed / 100000 : 0.0;
        double dd = 0.5 * 100000 * (dt * dt);             /* 0.5*100000 folds to 50000.0 */
        if (fabs(open - target) > 1) {
            if (open < target) {
                if (open < target - dd) speed += 5000000 / 1000; else speed -= 100000 / 1000;
                dir = 1.0;
            } else {
                if (open > dd) speed += 5000000 / 1000; else speed -= 100000 / 1000;
                dir = -1.0;
            }
        } else {
            speed = 0.0; open = target;
        }
        open += (speed * dir) / 1000;
        /* weight_simulation(open): :122-125 */
        double u = noise_tick(nz, n);
        n += 1;
        w += open * (1.0 + u);
        if (w < 0) w = 0.0;
        /* stage selection: :191-213 */
        if (w < fast_w) {
            target = fast_pos;
            if (!delay_done) { tick += p[4]; delay_done = 1; w += target * sysdelay; }
        } else if (w < medium_w) {
            target = medium_pos;
            if (!delay_done) { tick += p[5]; delay_done = 1; w += target * sysdelay; }
        } else if (w < slow_w) {
            target = slow_pos;
            if (!delay_done) { tick += p[6]; delay_done = 1; w += target * sysdelay; }
        } else {
            break;
        }
        /* safety stops: :216-229 */
        iteration += 1;
        if (fabs(w - prev) < 0.01) nochg += 1; else nochg = 0;
        prev = w;
        if (iteration > 5000) break;
        if (nochg > 500) break;
    }
    *total_time = (tick + p[10]) / 1000.0;                /* :239 */
    *final_weight = w;
    if (tick_out) *tick_out = tick;
    return n;
}

/* ------------------------------------------------------------------------------------ */
/* N1: the "on-board" per-tick models of the env classes (WeightEnv.py:176-229,            */
/* MutiConditionEnv.py:184-226), driven by the stage logic of simulate_run above (in the     */
/* reference the stage logic of this variant lives in an external PLC polled over Modbus,   */
/* WeightEnv.py:245-331; the offline controller's own loop is the scripted stand-in).        */
/*   motor_simulation: decelerate_time = speed / 100000 with NO `speed > 0` guard,          */
/*     np.power(dt, 2) (== dt*dt for every reachable speed, tests/test_oracle_golden.py),    */
/*     and the speed CAP: `if current_speed < motor_speed (10000): current_speed += 5000`.   */
/*   weight_simulation: current_weight = int(current_weight + opening * (1.0 + u)):         */
/*     truncation toward zero on every tick, no clamp at zero.                               */
/* ------------------------------------------------------------------------------------ */
static int simulate_core_onboard(const int p[11], double sysdelay, orc_noise *nz, double *total_time,
                                 double *final_weight, int *tick_out) {
    const double fast_pos = p[7], medium_pos = p[8], slow_pos = p[9];
    const double fast_w = p[1], medium_w = p[2], slow_w = p[3];
    double w = 0.0, open = 0.0, speed = 0.0, dir = 0.0;  /* WeightEnv.py:66-68,78 */
    double target = fast_pos;
    int delay_done = 0, tick = 0, iteration = 0, nochg = 0, n = 0;
    double prev = 0.0;
    for (;;) {
        tick += 1;
        /* motor_simulation(target): WeightEnv.py:188-218 */
        double dt = speed / 100000;
        double dd = 0.5 * 100000 * (dt * dt);
        if (fabs(open - target) > 1) {
            if (open < target) {
                dir = 1.0;
                if (open < target - dd) { if (speed < 10000) speed += 5000000 / 1000; }
                else speed -= 100000 / 1000;
            } else if (open > target) {
                dir = -1.0;
                if (open > dd) { if (speed < 10000) speed += 5000000 / 1000; }
                else speed -= 100000 / 1000;
            }
        } else {
            speed = 0.0; open = target;
        }
        open += (speed * dir) / 1000;
        /* weight_simulation(open, 1.0): WeightEnv.py:228-229, int() truncates toward zero */
        double u = noise_tick(nz, n);
        n += 1;
        w = trunc(w + open * (1.0 + u));
        /* stage selection / first-tick delay / safety stops: VirtualWeightController.py:191-229 */
        if (w < fast_w) {
            target = fast_pos;
            if (!delay_done) { tick += p[4]; delay_done = 1; w += target * sysdelay; }
        } else if (w < medium_w) {
            target = medium_pos;
            if (!delay_done) { tick += p[5]; delay_done = 1; w += target * sysdelay; }
        } else if (w < slow_w) {
            target = slow_pos;
            if (!delay_done) { tick += p[6]; delay_done = 1; w += target * sysdelay; }
        } else {
            break;
        }
        iteration += 1;
        if (fabs(w - prev) < 0.01) nochg += 1; else nochg = 0;
        prev = w;
        if (iteration > 5000) break;
        if (nochg > 500) break;
    }
    *total_time = (tick + p[10]) / 1000.0;
    *final_weight = w;
    if (tick_out) *tick_out = tick;
    return n;
}

int orc_simulate_onboard_noise(const int p[11], double sysdelay, const double *u, int n_u,
                               double *total_time, double *final_weight, int *tick_out) {
    orc_noise nz; memset(&nz, 0, sizeof nz);
    nz.u = u; nz.n_u = n_u; nz.cache_block = -1;
    return simulate_core_onboard(p, sysdelay, &nz, total_time, final_weight, tick_out);
}

int orc_simulate_onboard_philox(const int p[11], uint32_t seed_lo, uint32_t seed_hi, uint32_t env_id,
                                uint32_t step_idx, double *total_time, double *final_weight, int *tick_out) {
    orc_noise nz; memset(&nz, 0, sizeof nz);
    nz.seed_lo = seed_lo; nz.seed_hi = seed_hi; nz.env_id = env_id; nz.step_idx = step_idx;
    nz.cache_block = -1;
    uint32_t r[4];
    orc_philox4x32(seed_lo, seed_hi, env_id, step_idx, 0u, ORC_STREAM_MISC, r);
    const double sysdelay = 15.0 + 35.0 * u32_to_unit(r[0]);      /* uniform(15,50) */
    return simulate_core_onboard(p, sysdelay, &nz, total_time, final_weight, tick_out);
}

int orc_simulate_run_noise(const int p[11], double sysdelay, const double *u, int n_u,
                           double *total_time, double *final_weight, int *tick_out) {
    orc_noise nz; memset(&nz, 0, sizeof nz);
    nz.u = u; nz.n_u = n_u; nz.cache_block = -1;
    return simulate_core(p, sysdelay, &nz, total_time, final_weight, tick_out);
}

double orc_philox_sysdelay(uint32_t seed_lo, uint32_t seed_hi, uint32_t env_id, uint32_t step_idx) {
    uint32_t r[4];
    orc_philox4x32(seed_lo, seed_hi, env_id, step_idx, 0u, ORC_STREAM_MISC, r);
    return 15.0 + 35.0 * u32_to_unit(r[0]);               /* uniform(15,50): :137 */
}

int orc_simulate_run_philox(const int p[11], uint32_t seed_lo, uint32_t seed_hi, uint32_t env_id,
                            uint32_t step_idx, double *total_time, double *final_weight,
                            int *tick_out) {
    orc_noise nz; memset(&nz, 0, sizeof nz);
    nz.seed_lo = seed_lo; nz.seed_hi = seed_hi; nz.env_id = env_id; nz.step_idx = step_idx;
    nz.cache_block = -1;
    double sysdelay = orc_philox_sysdelay(seed_lo, seed_hi, env_id, step_idx);
    return simulate_core(p, sysdelay, &nz, total_time, final_weight, tick_out);
}

/* ------------------------------------------------------------------------------------ */
/* numpy reductions restated (numpy/_core/src/umath/loops_utils.h.src pairwise_sum, n<128) */
/* ------------------------------------------------------------------------------------ */
static float pairwise_sum_f32(const float *a, int n) {
    if (n < 8) { float r = 0.f; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    float r[8]; int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}
static double pairwise_sum_f64(const double *a, int n) {
    if (n < 8) { double r = 0.; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    double r[8]; int i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}
/* np.std (population, ddof=0): numpy/_core/_methods.py _var/_std */
static float np_std_f32(const float *a, int n) {
    float mean = pairwise_sum_f32(a, n) / (float)n, x[16];
    for (int i = 0; i < n; ++i) { float d = a[i] - mean; x[i] = d * d; }
    return sqrtf(pairwise_sum_f32(x, n) / (float)n);
}
static double np_std_f64(const double *a, int n) {
    double mean = pairwise_sum_f64(a, n) / (double)n, x[16];
    for (int i = 0; i < n; ++i) { double d = a[i] - mean; x[i] = d * d; }
    return sqrt(pairwise_sum_f64(x, n) / (double)n);
}

/* ------------------------------------------------------------------------------------ */
/* Per-condition constants and per-env persistent state                                   */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int   target_weight;        /* g */
    double target_weight_err;   /* g */
    double target_time;         /* s */
    float lo[10], hi[10];       /* bounds in action order (env.bounds insertion order) */
} orc_condition;

#define ORC_RING 20             /* deque(maxlen=int(100/5)) fixed at construction: MutiConditionEnv.py:43-45 */
typedef struct {
    int step_counter, episode_counter;
    int ring_len, ring_head;    /* ring_head = index of the oldest entry */
    double ring[ORC_RING];
} orc_env_state;

enum { ORC_ENV_MULTI = 0, ORC_ENV_SINGLE = 1 };

typedef struct {
    float next_state[2];
    float reward;               /* reference reward rounded to binary32 */
    double reward64;            /* same value before the final rounding (equal to reward on the [f32] path) */
    int done;
    double final_weight, total_time, weight_error;
    int ticks;                  /* loop iterations (draws consumed) */
    int iparams[11];
    float clipped[10];
} orc_step_out;

/* clip (MutiConditionEnv.py:365-367 / WeightEnv.py:405-407) then int() truncation toward zero
 * (VirtualWeightController.py:71-83); action arrives as binary32 (DDPGAgent.act returns f32). */
static void clip_and_truncate(const orc_condition *c, const float *action, float *clipped, int *ip) {
    for (int i = 0; i < 10; ++i) {
        float v = action[i];
        v = v < c->lo[i] ? c->lo[i] : v;   /* np.clip == minimum(maximum(a, lo), hi) */
        v = v > c->hi[i] ? c->hi[i] : v;
        clipped[i] = v;
    }
    /* dict order: fast_weight medium_weight slow_weight fast_opening medium_opening slow_opening
       fast_delay medium_delay slow_delay unload_delay  (MutiConditionEnv.py:350-362) */
    ip[0] = c->target_weight;
    ip[1] = (int)clipped[0]; ip[2] = (int)clipped[1]; ip[3] = (int)clipped[2];
    ip[7] = (int)clipped[3]; ip[8] = (int)clipped[4]; ip[9] = (int)clipped[5];
    ip[4] = (int)clipped[6]; ip[5] = (int)clipped[7]; ip[6] = (int)clipped[8];
    ip[10] = (int)clipped[9];
}

/* E2 reward (MutiConditionEnv.py:376-436), numpy>=2 promotion: the sum stays a Python float
 * until the first numpy scalar joins it. */
static void reward_multi(const orc_condition *c, const float *clipped, double fw, double tt,
                         double *we_out, double *te_out, float *r32, double *r64) {
    const double err = c->target_weight_err, T = c->target_time;
    double we = fabs(fw - c->target_weight), te = fabs(tt - T);
    double error_reward;
    if (we <= 10) error_reward = 5000;
    else if (we <= err) error_reward = 4000;
    else if (we <= err * 2) error_reward = 3000;
    else if (we <= err * 4) error_reward = 1500;
    else if (we <= err * 6) error_reward = -1000;
    else if (we <= err * 8) error_reward = -2000;
    else if (we <= err * 12) error_reward = -3000;
    else error_reward = -300 - (we - 300) * 0.5;
    double time_reward;
    if (te <= 0.3 * T) time_reward = 500;
    else if (te <= 0.5 * T) time_reward = 300;
    else if (te <= 0.8 * T) time_reward = 100;
    else if (te <= 1.1 * T) time_reward = 10;
    else if (te <= 2 * T) time_reward = -100;
    else if (te <= 3 * T) time_reward = -200;
    else time_reward = -300 - (te - 3 * T) * 200;
    /* diversity: np.std over 11 binary32 values (10 clipped + target_weight) [f32] (:419-421) */
    float arr[11];
    for (int i = 0; i < 10; ++i) arr[i] = clipped[i];
    arr[10] = (float)c->target_weight;
    float diversity = 0.05f * np_std_f32(arr, 11);       /* python float * np.float32 -> f32 */
    /* boundary penalty (:424-430); comparison in binary64 */
    int penalty = 0;
    for (int i = 0; i < 10; ++i) {
        double v = clipped[i];
        if (v <= c->lo[i] + 1e-3 || v >= c->hi[i] - 1e-3) penalty -= 300;
    }
    double base = ((-1000 + error_reward) + time_reward);
    if (we <= err) {
        double tf = (3 * T - tt) / T;
        tf = tf < 0 ? 0 : (tf > 3 ? 3 : tf);
        double target_reward = 2000 * tanh(tf);          /* np.float64 */
        double r = ((base + target_reward) + (double)diversity) + penalty;
        r = r < -3000 ? -3000 : r; r = r > 6000 ? 6000 : r;
        *r64 = r; *r32 = (float)r;
    } else {
        float r = ((float)base + diversity) + (float)penalty;   /* [f32] weak promotion */
        r = r < -3000.f ? -3000.f : r; r = r > 6000.f ? 6000.f : r;
        *r64 = r; *r32 = r;
    }
    *we_out = we; *te_out = te;
}

/* E3 reward v2 (WeightEnv.py:480-556); everything binary64 (the action dict holds np.float32
 * scalars and one Python int, np.array(...) of that list is float64). */
static void reward_single(const orc_condition *c, const float *clipped, double fw, double tt,
                          double *we_out, double *te_out, float *r32, double *r64) {
    const double T = c->target_time;
    double we = fabs(fw - c->target_weight), te = fabs(tt - T);
    double r1;
    if (we <= 5) r1 = 8000 - 200 * we;
    else if (we <= 10) r1 = 6000 - 120 * (we - 5);
    else if (we <= 15) r1 = 4000 - 40 * (we - 10);
    else if (we <= 20) r1 = 2000 - 20 * (we - 15);
    else if (we <= 25) r1 = -1000 - 10 * (we - 20);
    else if (we <= 50) r1 = -3000 - 5 * (we - 25);
    else r1 = -4000 - 0.5 * (we - 50);
    double r2;
    if (te <= 0.1 * T) r2 = 3000;
    else if (te <= 0.3 * T) r2 = 1000;
    else if (te <= 0.5 * T) r2 = 500;
    else if (te <= 0.8 * T) r2 = -1000;
    else if (te <= 1.2 * T) r2 = -2000;
    else if (te <= 2 * T) r2 = -3000;
    else r2 = -3000 - (te - 1.05 * T) * 500;
    double r3 = 0;
    if (we <= 25) {
        double tf = (3 * T - tt) / T;
        tf = tf < 0 ? 0 : (tf > 3 ? 3 : tf);
        r3 = 2000 * tanh(tf);
    }
    double arr[11];
    for (int i = 0; i < 10; ++i) arr[i] = clipped[i];
    arr[10] = c->target_weight;
    double r4 = np_std_f64(arr, 11) * 0.2;
    int pen = 0;
    for (int i = 0; i < 10; ++i) {
        double lo = c->lo[i], hi = c->hi[i], v = clipped[i];
        double buffer = (hi - lo) * 0.05;
        if (v <= lo + buffer || v >= hi - buffer) pen += 1;
    }
    double boundary = -300 * pen;
    double r = 3500 + 0.3 * r1 + 0.3 * r2 + 0.2 * r3 + 0.15 * r4 + 0.05 * boundary;
    r = r < -10000 ? -10000 : r; r = r > 10000 ? 10000 : r;
    *r64 = r; *r32 = (float)r; *we_out = we; *te_out = te;
}

/* state/done (MutiConditionEnv.py:439-452 == WeightEnv.py:559-572) */
static void finish_step(int kind, const orc_condition *c, orc_env_state *st, int max_steps,
                        double fw, double tt, orc_step_out *o) {
    double we, te;
    if (kind == ORC_ENV_MULTI) reward_multi(c, o->clipped, fw, tt, &we, &te, &o->reward, &o->reward64);
    else reward_single(c, o->clipped, fw, tt, &we, &te, &o->reward, &o->reward64);
    o->next_state[0] = (float)(we / c->target_weight_err);
    o->next_state[1] = (float)(te / c->target_time);
    /* recent_errors.append (deque maxlen 20) */
    if (st->ring_len < ORC_RING) {
        st->ring[(st->ring_head + st->ring_len) % ORC_RING] = we; st->ring_len += 1;
    } else {
        st->ring[st->ring_head] = we; st->ring_head = (st->ring_head + 1) % ORC_RING;
    }
    int done = st->step_counter >= max_steps;
    if (!done && st->step_counter >= max_steps / 2 && st->ring_len == ORC_RING) {
        double tmp[ORC_RING];
        for (int i = 0; i < ORC_RING; ++i) tmp[i] = st->ring[(st->ring_head + i) % ORC_RING];
        double mean = pairwise_sum_f64(tmp, ORC_RING) / ORC_RING;
        done = mean < 25;                                  /* literal 25, not target_weight_err */
    }
    o->done = done;
    st->step_counter += 1;
    o->final_weight = fw; o->total_time = tt; o->weight_error = we;
}

void orc_env_step_noise(int kind, const orc_condition *c, orc_env_state *st, int max_steps,
                        const float action[10], double sysdelay, const double *u, int n_u,
                        orc_step_out *o) {
    clip_and_truncate(c, action, o->clipped, o->iparams);
    double tt, fw;
    o->ticks = (kind & 2) ? orc_simulate_onboard_noise(o->iparams, sysdelay, u, n_u, &tt, &fw, NULL)
                          : orc_simulate_run_noise(o->iparams, sysdelay, u, n_u, &tt, &fw, NULL);
    finish_step(kind & 1, c, st, max_steps, fw, tt, o);
}

void orc_env_step_philox(int kind, const orc_condition *c, orc_env_state *st, int max_steps,
                         const float action[10], uint32_t seed_lo, uint32_t seed_hi,
                         uint32_t env_id, uint32_t step_idx, orc_step_out *o) {
    clip_and_truncate(c, action, o->clipped, o->iparams);
    double tt, fw;
    o->ticks = (kind & 2) ? orc_simulate_onboard_philox(o->iparams, seed_lo, seed_hi, env_id, step_idx, &tt, &fw, NULL)
                          : orc_simulate_run_philox(o->iparams, seed_lo, seed_hi, env_id, step_idx, &tt, &fw, NULL);
    finish_step(kind & 1, c, st, max_steps, fw, tt, o);
}

/* E4 reset, cold-start branch: raw = f32([U(0,err), U(0,T)]); state = f32 divisions [f32]. */
void orc_env_reset_cold(const orc_condition *c, orc_env_state *st, double u0, double u1,
                        float state[2]) {
    float raw0 = (float)(0.0 + (c->target_weight_err - 0.0) * u0);  /* random.uniform(0, err) */
    float raw1 = (float)(0.0 + (c->target_time - 0.0) * u1);
    state[0] = raw0 / (float)c->target_weight_err;
    state[1] = raw1 / (float)c->target_time;
    st->step_counter = 0; st->episode_counter += 1; st->ring_len = 0; st->ring_head = 0;
}
/* E4 reset, replay branch: mean over 10 sampled stored states (axis-0 reduction = row-sequential
 * binary32 adds, then /10), then normalised AGAIN (reference behaviour). */
void orc_env_reset_replay(const orc_condition *c, orc_env_state *st, const float *states10 /*[10][2]*/,
                          float state[2]) {
    float s0 = states10[0], s1 = states10[1];
    for (int i = 1; i < 10; ++i) { s0 += states10[2 * i]; s1 += states10[2 * i + 1]; }
    float raw0 = s0 / 10.f, raw1 = s1 / 10.f;
    state[0] = raw0 / (float)c->target_weight_err;
    state[1] = raw1 / (float)c->target_time;
    st->step_counter = 0; st->episode_counter += 1; st->ring_len = 0; st->ring_head = 0;
}
void orc_philox_reset_uniforms(uint32_t seed_lo, uint32_t seed_hi, uint32_t env_id,
                               uint32_t step_idx, double *u0, double *u1) {
    uint32_t r[4];
    orc_philox4x32(seed_lo, seed_hi, env_id, step_idx, 0u, ORC_STREAM_MISC, r);
    *u0 = u32_to_unit(r[1]); *u1 = u32_to_unit(r[2]);
}

/* ------------------------------------------------------------------------------------ */
/* M1: EmotionModule.update (EmotionModule.py:137-162) in binary64.                        */
/* cur[3] in/out = current_emotion; returns the vector pushed to the history (as f32).     */
/* ------------------------------------------------------------------------------------ */
static inline double clip01(double x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
void orc_emotion_update(double cur[3], double reward, const float movement[2], float pushed[3]) {
    double tr = tanh(reward);
    double reward_term = 1 - tr;
    double mm = sqrt((double)movement[0] * movement[0] + (double)movement[1] * movement[1]);
    double mf = clip01(mm);
    double exploration = 0.8 * reward_term + 0.2 * mf;
    double conserv = clip01(tr - 0.1 * mf);
    double scaled = tanh(reward * 0.001);
    double anx = clip01(cur[2] + (-0.2 * scaled + 0.05 * mf));
    double e0 = clip01(0.8 * cur[0] + 0.2 * exploration);
    double e1 = clip01(0.8 * cur[1] + 0.2 * conserv);
    cur[0] = e0; cur[1] = e1; cur[2] = anx;
    pushed[0] = (float)e0; pushed[1] = (float)e1; pushed[2] = (float)anx;
}

/* ------------------------------------------------------------------------------------ */
/* A2: OU noise step + act() clip (ddpg_agent.py:352-356, 768-778), binary32 like numpy.   */
/* ------------------------------------------------------------------------------------ */
void orc_ou_act(float ou_state[10], float sigma, const float randn[10], const float scaled_action[10],
                const float scale[10], const float offset[10], float final_action[10]) {
    for (int i = 0; i < 10; ++i) {
        float dx = 0.15f * (0.0f - ou_state[i]);     /* theta*(mu - x): python float * f32 array */
        dx += sigma * randn[i];
        ou_state[i] += dx;
        float a = scaled_action[i] + ou_state[i] * scale[i];
        float lo = offset[i] - scale[i], hi = offset[i] + scale[i];
        a = a < lo ? lo : a; a = a > hi ? hi : a;
        final_action[i] = a;
    }
}
/* R1: norm_a = clip((a - offset)/scale, -1, 1) in binary32 (ddpg_agent.py:373-375) */
void orc_normalize_action(const float a[10], const float scale[10], const float offset[10], float out[10]) {
    for (int i = 0; i < 10; ++i) {
        float v = (a[i] - offset[i]) / scale[i];
        v = v < -1.f ? -1.f : v; v = v > 1.f ? 1.f : v;
        out[i] = v;
    }
}
/* Philox normals for OU (Box-Muller on the (0,1] grid); contract shared with the CUDA kernel. */
void orc_philox_randn10(uint32_t seed_lo, uint32_t seed_hi, uint32_t env_id, uint32_t step_idx, float out[10]) {
    for (int b = 0; b < 3; ++b) {
        uint32_t r[4];
        orc_philox4x32(seed_lo, seed_hi, env_id, step_idx, (uint32_t)b, ORC_STREAM_OU, r);
        for (int h = 0; h < 2; ++h) {
            int k = 4 * b + 2 * h;
            if (k >= 10) break;
            double u1 = ((double)r[2 * h] + 1.0) * (1.0 / 4294967296.0);      /* (0,1] */
            double u2 = (double)r[2 * h + 1] * (1.0 / 4294967296.0);
            double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586476925286766559 * u2;
            out[k] = (float)(rad * cos(ang));
            if (k + 1 < 10) out[k + 1] = (float)(rad * sin(ang));
        }
    }
}

/* R2: equal-priority np.random.choice(n, B, p) == floor(u*n) */
void orc_sample_indices_uniform(const double *u, int B, int64_t n, int64_t *idx) {
    for (int i = 0; i < B; ++i) {
        int64_t k = (int64_t)floor(u[i] * (double)n);
        idx[i] = k >= n ? n - 1 : k;
    }
}
/* Philox variant used by the CUDA sampler: counter (draw_block, update_idx, 0, stream 3) */
#define ORC_STREAM_SAMPLE 3u
void orc_sample_indices_philox(uint32_t seed_lo, uint32_t seed_hi, uint32_t update_idx, int B,
                               int64_t n, int64_t *idx) {
    for (int i = 0; i < B; ++i) {
        uint32_t r[4];
        orc_philox4x32(seed_lo, seed_hi, (uint32_t)(i >> 2), update_idx, 0u, ORC_STREAM_SAMPLE, r);
        /* floor(U*n) with U on the 2^-32 grid: exact integer arithmetic */
        idx[i] = (int64_t)(((uint64_t)r[i & 3] * (uint64_t)n) >> 32);
    }
}

/* ------------------------------------------------------------------------------------ */
/* Batched driver: N envs, SoA in/out, used as the checker for the CUDA env-step kernel    */
/* and as the CPU baseline (OpenMP over envs).                                             */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int kind, n_env, n_cond, max_steps, auto_reset;
    const orc_condition *conds; const int *cond_id; orc_env_state *states; const float *actions;
    uint32_t seed_lo, seed_hi, env_id_base, step_idx;
    float *next_state, *obs, *reward; unsigned char *done; double *final_weight, *total_time; int *ticks;
    atomic_int next; atomic_llong total_ticks;
} orc_batch_job;

static void *batch_worker(void *arg) {
    orc_batch_job *j = (orc_batch_job *)arg;
    long long local = 0;
    for (;;) {
        int lo = atomic_fetch_add(&j->next, 16);
        if (lo >= j->n_env) break;
        int hi = lo + 16 < j->n_env ? lo + 16 : j->n_env;
        for (int i = lo; i < hi; ++i) {
            uint32_t gid = j->env_id_base + (uint32_t)i;
            const orc_condition *c = &j->conds[j->cond_id ? j->cond_id[i] : (int)(gid % (uint32_t)j->n_cond)];
            orc_step_out o;
            orc_env_step_philox(j->kind, c, &j->states[i], j->max_steps, j->actions + 10 * i, j->seed_lo,
                                j->seed_hi, gid, j->step_idx, &o);
            j->next_state[2 * i] = o.next_state[0]; j->next_state[2 * i + 1] = o.next_state[1];
            j->reward[i] = o.reward; j->done[i] = (unsigned char)o.done;
            j->final_weight[i] = o.final_weight; j->total_time[i] = o.total_time; j->ticks[i] = o.ticks;
            local += o.ticks;
            if (j->obs) {
                if (o.done && j->auto_reset) {
                    double u0, u1;
                    orc_philox_reset_uniforms(j->seed_lo, j->seed_hi, gid, j->step_idx, &u0, &u1);
                    orc_env_reset_cold(c, &j->states[i], u0, u1, j->obs + 2 * i);
                } else {
                    j->obs[2 * i] = o.next_state[0]; j->obs[2 * i + 1] = o.next_state[1];
                }
            }
        }
    }
    atomic_fetch_add(&j->total_ticks, local);
    return NULL;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* n_threads <= 0: all online cores.  Returns the total number of ticks simulated. */
long long orc_env_step_batch_philox(int kind, int n_env, const orc_condition *conds, int n_cond,
                                    const int *cond_id /* may be NULL: env_id % n_cond */,
                                    orc_env_state *states, int max_steps, const float *actions /*[N][10]*/,
                                    uint32_t seed_lo, uint32_t seed_hi, uint32_t env_id_base,
                                    uint32_t step_idx, int auto_reset,
                                    float *next_state /*[N][2]*/, float *obs /*[N][2]*/, float *reward,
                                    unsigned char *done, double *final_weight, double *total_time,
                                    int *ticks, int n_threads) {
    orc_batch_job j;
    j.kind = kind; j.n_env = n_env; j.n_cond = n_cond; j.max_steps = max_steps; j.auto_reset = auto_reset;
    j.conds = conds; j.cond_id = cond_id; j.states = states; j.actions = actions;
    j.seed_lo = seed_lo; j.seed_hi = seed_hi; j.env_id_base = env_id_base; j.step_idx = step_idx;
    j.next_state = next_state; j.obs = obs; j.reward = reward; j.done = done;
    j.final_weight = final_weight; j.total_time = total_time; j.ticks = ticks;
    atomic_init(&j.next, 0); atomic_init(&j.total_ticks, 0);
    if (n_threads <= 0) n_threads = orc_max_threads();
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) { batch_worker(&j); return atomic_load(&j.total_ticks); }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads; ++t) if (pthread_create(&th[started], NULL, batch_worker, &j) == 0) started++;
    if (started == 0) batch_worker(&j);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    return atomic_load(&j.total_ticks);
}

int orc_sizeof_condition(void) { return (int)sizeof(orc_condition); }
int orc_sizeof_env_state(void) { return (int)sizeof(orc_env_state); }
int orc_sizeof_step_out(void) { return (int)sizeof(orc_step_out); }
