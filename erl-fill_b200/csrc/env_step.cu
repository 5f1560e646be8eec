// env_step.cu — fused filling-cycle kernel: one thread per env, SoA state, binary64 simulator
// registers, Philox-4x32-10 noise generated in-kernel (or injected from HBM in parity mode),
// reward shaping + done + auto-reset fused in the epilogue.
//
// Replaces (reference file:line under /root/reference):
//   E1 VirtualWeightController.simulate_run + motor_simulation + weight_simulation
//      (VirtualWeightController.py:127-244, 85-113, 115-125) with load_params' int() truncation (:71-83)
//   E2 MutiConditionEnv.WeightEnv.step  (MutiConditionEnv.py:346-468)
//   E3 WeightEnv.step reward v2          (WeightEnv.py:480-572)
//   E4 reset                             (MutiConditionEnv.py:135-164 / WeightEnv.py:113-148)
//
// Compiled with -fmad=false: the reference is CPython binary64 with separately rounded * and +,
// and this kernel reproduces it bit for bit (tests/test_env_step_gpu.py), so no contraction.
//
// Bound: the tick loop is a serial dependency chain of ~14 FP64 instructions per 1 ms tick over
// T = 640..5001 ticks; it touches ~200 B of HBM per env-step.  It is FP64-pipe/latency bound,
// not HBM bound; the roofline reported for it is ticks/s against the FP64
// issue ceiling, with the HBM figure given for completeness.
#include <math.h>

#include <mutex>

#include "common.cuh"

namespace erlfill {

thread_local int g_last_cuda_error = 0;

// ------------------------------------------------------------------------------------------
// Motor table.  motor_simulation() only ever changes current_speed by +5000, -100 or to 0, so
// speed == 100*k for an integer k, and both per-tick quantities that need a division are pure
// functions of k:
//   .x  decelerate_distance(k) = 0.5*100000*(speed/100000)**2     (VirtualWeightController.py:92-93)
//   .y  |speed|/1000                                                (:112-113)
// The reference evaluates `**2` with libm pow(); the table is built ON THE HOST with the same
// libm so every entry is bit-identical to CPython's value (pow(dt,2) != dt*dt from k = 2759 on).
// k outside the table (speeds above 409,500 or negative; unreachable for opening targets <= 100,
// peak observed 9200) falls back to arithmetic evaluated on the device.
// ------------------------------------------------------------------------------------------
constexpr int kTabGlobal = 4096;
constexpr int kTabShared = 256;
__device__ double2 g_motor_tab[kTabGlobal];

static std::mutex g_tab_mutex;
static bool g_tab_ready[64] = {};

static int ensure_motor_table() {
    int dev = 0;
    ERLFILL_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return ERLFILL_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    if (g_tab_ready[dev]) return ERLFILL_OK;
    static double2 host_tab[kTabGlobal];
    double (*volatile libm_pow)(double, double) = pow;   // volatile: keep the libm call, no x*x folding
    for (int k = 0; k < kTabGlobal; ++k) {
        const double speed = 100.0 * k;
        const double dt = speed > 0 ? speed / 100000 : 0.0;
        volatile double half_decel = 0.5 * 100000;
        host_tab[k].x = half_decel * libm_pow(dt, 2.0);
        host_tab[k].y = speed / 1000;
    }
    ERLFILL_CUDA_TRY(cudaMemcpyToSymbol(g_motor_tab, host_tab, sizeof(host_tab)));
    g_tab_ready[dev] = true;
    return ERLFILL_OK;
}

__device__ __forceinline__ double2 motor_entry(const double2 *s_tab, int k) {
    if ((unsigned)k < (unsigned)kTabShared) return s_tab[k];
    if ((unsigned)k < (unsigned)kTabGlobal) return g_motor_tab[k];
    // out-of-table: evaluate on the device (dt*dt instead of libm pow; see header comment)
    const double speed = 100.0 * (double)k;
    const double dt = speed > 0 ? __ddiv_rn(speed, 100000.0) : 0.0;
    double2 e;
    e.x = __dmul_rn(50000.0, __dmul_rn(dt, dt));
    e.y = __ddiv_rn(fabs(speed), 1000.0);
    return e;
}

// ------------------------------------------------------------------------------------------
struct StepArgs {
    erlfill_env_batch env;
    erlfill_step_out out;
    const float *d_actions;
    const double *d_u;         // injected noise [rows][n_env] or nullptr
    const double *d_reset_u;   // [2][n_env] or nullptr
    int noise_rows;
    uint32_t step_idx;
    int auto_reset;
    int actions_aligned16;
};

struct SimResult {
    double final_weight;
    int tick;        // simulated-time ticks incl. the stage delay
    int draws;       // loop iterations == noise draws consumed
};

// numpy pairwise_sum for 8 <= n < 16 (numpy/_core/src/umath/loops_utils.h.src): 8 accumulators
// seeded with a[0..7], tree-combined, then the tail added sequentially.
template <typename T>
__device__ __forceinline__ T np_sum11(const T *a) {
    T res = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    res += a[8]; res += a[9]; res += a[10];
    return res;
}
template <typename T>
__device__ __forceinline__ T np_std11(const T *a) {   // population std, numpy/_core/_methods.py:_var
    const T mean = np_sum11(a) / (T)11;
    T x[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) { const T d = a[i] - mean; x[i] = d * d; }
    const T var = np_sum11(x) / (T)11;
    return sizeof(T) == 4 ? (T)sqrtf((float)var) : (T)sqrt((double)var);
}

// E2 reward: MutiConditionEnv.py:376-436 with numpy>=2 weak-scalar promotion (the [f32] branch)
__device__ float reward_multi(const erlfill_condition &c, const float *clipped, double fw, double tt,
                              double &we, double &te) {
    const double err = c.target_weight_err, T = c.target_time;
    we = fabs(fw - (double)c.target_weight);
    te = fabs(tt - T);
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
    float arr[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) arr[i] = clipped[i];
    arr[10] = (float)c.target_weight;
    const float diversity = 0.05f * np_std11<float>(arr);
    int penalty = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const double v = clipped[i];
        if (v <= (double)c.lo[i] + 1e-3 || v >= (double)c.hi[i] - 1e-3) penalty -= 300;
    }
    const double base = (-1000 + error_reward) + time_reward;
    if (we <= err) {
        double tf = (3 * T - tt) / T;
        tf = tf < 0 ? 0 : (tf > 3 ? 3 : tf);
        double r = ((base + 2000 * tanh(tf)) + (double)diversity) + (double)penalty;
        r = r < -3000 ? -3000 : r;
        r = r > 6000 ? 6000 : r;
        return (float)r;
    }
    float r = ((float)base + diversity) + (float)penalty;
    r = r < -3000.f ? -3000.f : r;
    r = r > 6000.f ? 6000.f : r;
    return r;
}

// E3 reward v2: WeightEnv.py:480-556, binary64 throughout
__device__ float reward_single(const erlfill_condition &c, const float *clipped, double fw, double tt,
                               double &we, double &te) {
    const double T = c.target_time;
    we = fabs(fw - (double)c.target_weight);
    te = fabs(tt - T);
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
#pragma unroll
    for (int i = 0; i < 10; ++i) arr[i] = clipped[i];
    arr[10] = c.target_weight;
    const double r4 = np_std11<double>(arr) * 0.2;
    int pen = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const double lo = c.lo[i], hi = c.hi[i], v = clipped[i];
        const double buffer = (hi - lo) * 0.05;
        if (v <= lo + buffer || v >= hi - buffer) pen += 1;
    }
    const double boundary = -300 * pen;
    double r = 3500 + 0.3 * r1 + 0.3 * r2 + 0.2 * r3 + 0.15 * r4 + 0.05 * boundary;
    r = r < -10000 ? -10000 : r;
    r = r > 10000 ? 10000 : r;
    return (float)r;
}

// E4 cold-start reset state (MutiConditionEnv.py:138-154): raw in binary32, binary32 divisions
__device__ __forceinline__ float2 reset_state_cold(const erlfill_condition &c, double u0, double u1) {
    const float raw0 = (float)(0.0 + (c.target_weight_err - 0.0) * u0);
    const float raw1 = (float)(0.0 + (c.target_time - 0.0) * u1);
    return make_float2(raw0 / (float)c.target_weight_err, raw1 / (float)c.target_time);
}

// ------------------------------------------------------------------------------------------
// The tick loop.  State per thread: w, open, prev (binary64), k = speed/100 (int), dir (+1/-1),
// target opening, counters.  `opu` = 1.0 + uniform(-1,1).
// ------------------------------------------------------------------------------------------
struct SimParams {
    double fast_w, medium_w, slow_w;
    double fast_pos, medium_pos, slow_pos;
    int fast_delay, medium_delay, slow_delay;
};

struct SimState {
    double w, open, prev, target, step_mag, dd;
    int k, tick, iter, nochg;
    bool neg_dir;
};

// One tick.  Returns true when the run has finished.
template <bool kFirst>
__device__ __forceinline__ bool sim_tick(SimState &s, const SimParams &p, const double2 *s_tab, double opu,
                                         double sysdelay) {
    s.tick += 1;                                                        // :181
    // motor_simulation(target): :92-113.  s.dd is decelerate_distance for the current speed.
    const double diff = s.open - s.target;
    if (fabs(diff) > 1.0) {
        bool accel;
        if (s.open < s.target) { accel = s.open < s.target - s.dd; s.neg_dir = false; }
        else                   { accel = s.open > s.dd;            s.neg_dir = true; }
        s.k += accel ? 50 : -1;                                         // +5000 / -100
    } else {
        s.k = 0;
        s.open = s.target;
    }
    const double2 e = motor_entry(s_tab, s.k);
    s.dd = e.x;
    // (speed*dir)/1000: the quotient of the negated product equals the negated quotient
    double stp = e.y;
    if (s.k < 0) stp = -stp;
    s.open = s.open + (s.neg_dir ? -stp : stp);
    // weight_simulation(open): :122-125
    s.w = __dadd_rn(s.w, __dmul_rn(s.open, opu));
    if (s.w < 0) s.w = 0.0;
    // stage selection :191-213; the one-time delay is taken on the first tick only (the flag is
    // set once and never cleared), whichever stage that tick lands in
    if (s.w < p.fast_w) {
        s.target = p.fast_pos;
        if (kFirst) { s.tick += p.fast_delay; s.w = __dadd_rn(s.w, __dmul_rn(s.target, sysdelay)); }
    } else if (s.w < p.medium_w) {
        s.target = p.medium_pos;
        if (kFirst) { s.tick += p.medium_delay; s.w = __dadd_rn(s.w, __dmul_rn(s.target, sysdelay)); }
    } else if (s.w < p.slow_w) {
        s.target = p.slow_pos;
        if (kFirst) { s.tick += p.slow_delay; s.w = __dadd_rn(s.w, __dmul_rn(s.target, sysdelay)); }
    } else {
        return true;
    }
    // safety stops :216-229 (the 10 s wall-clock stop :230-233 has no meaning on a frozen clock)
    s.iter += 1;
    s.nochg = (fabs(s.w - s.prev) < 0.01) ? s.nochg + 1 : 0;
    s.prev = s.w;
    return (s.iter > ERLFILL_MAX_ITER) | (s.nochg > ERLFILL_NOCHANGE_LIMIT);
}

// 1.0 + (-1.0 + 2*U) == x * 2^-31 exactly for U = x * 2^-32: place x in the top mantissa bits of
// a double in [2,4) and subtract 2.
__device__ __forceinline__ double opu_from_u32(uint32_t x) {
    return __hiloint2double((int)(0x40000000u | (x >> 12)), (int)(x << 20)) - 2.0;
}

template <bool kInjected>
__device__ __forceinline__ SimResult simulate(const SimParams &p, const double2 *s_tab, double sysdelay,
                                              uint32_t k0, uint32_t k1, uint32_t gid, uint32_t step_idx,
                                              const double *u_col, int n_env, int noise_rows) {
    SimState s;
    s.w = 0.0; s.open = 0.0; s.prev = 0.0; s.target = p.fast_pos; s.dd = 0.0; s.step_mag = 0.0;
    s.k = 0; s.tick = 0; s.iter = 0; s.nochg = 0; s.neg_dir = false;
    int draws = 0;
    bool fin;
    if (kInjected) {
        // row 1+t of the [rows][n_env] noise matrix is tick t's uniform(-1,1)
        auto u_at = [&](int t) -> double {
            return (1 + t) < noise_rows ? __ldg(u_col + (size_t)(1 + t) * n_env) : 0.0;
        };
        fin = sim_tick<true>(s, p, s_tab, 1.0 + u_at(0), sysdelay);
        draws = 1;
        while (!fin) {
            fin = sim_tick<false>(s, p, s_tab, 1.0 + u_at(draws), sysdelay);
            draws += 1;
        }
    } else {
        uint4 r = philox4x32_10(k0, k1, gid, step_idx, 0u, STREAM_TICK);
        fin = sim_tick<true>(s, p, s_tab, opu_from_u32(r.x), sysdelay);
        draws = 1;
        if (!fin) { fin = sim_tick<false>(s, p, s_tab, opu_from_u32(r.y), sysdelay); draws = 2; }
        if (!fin) { fin = sim_tick<false>(s, p, s_tab, opu_from_u32(r.z), sysdelay); draws = 3; }
        if (!fin) { fin = sim_tick<false>(s, p, s_tab, opu_from_u32(r.w), sysdelay); draws = 4; }
        uint32_t blk = 1;
        while (!fin) {
            r = philox4x32_10(k0, k1, gid, step_idx, blk, STREAM_TICK);
            ++blk;
            fin = sim_tick<false>(s, p, s_tab, opu_from_u32(r.x), sysdelay); draws += 1;
            if (fin) break;
            fin = sim_tick<false>(s, p, s_tab, opu_from_u32(r.y), sysdelay); draws += 1;
            if (fin) break;
            fin = sim_tick<false>(s, p, s_tab, opu_from_u32(r.z), sysdelay); draws += 1;
            if (fin) break;
            fin = sim_tick<false>(s, p, s_tab, opu_from_u32(r.w), sysdelay); draws += 1;
        }
    }
    SimResult res;
    res.final_weight = s.w;
    res.tick = s.tick;
    res.draws = draws;
    return res;
}

constexpr int kMaxBlock = 256;

template <bool kInjected>
__global__ void __launch_bounds__(kMaxBlock) env_step_kernel(const StepArgs a) {
    __shared__ __align__(16) double2 s_tab[kTabShared];
    __shared__ __align__(16) float s_act[kMaxBlock * ERLFILL_ACTION_DIM];

    const int n_env = a.env.n_env;
    const int block_base = blockIdx.x * blockDim.x;
    const int n_valid = min((int)blockDim.x, n_env - block_base);
    const int tid = threadIdx.x;

    for (int k = tid; k < kTabShared; k += blockDim.x) s_tab[k] = g_motor_tab[k];
    // stage this block's [n_valid][10] action rows with 128-bit coalesced loads
    {
        const float *src = a.d_actions + (size_t)block_base * ERLFILL_ACTION_DIM;
        const int nfloat = n_valid * ERLFILL_ACTION_DIM;
        int done_f = 0;
        if (a.actions_aligned16 && ((block_base * ERLFILL_ACTION_DIM) & 3) == 0) {
            const int nvec = nfloat >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(s_act);
            for (int v = tid; v < nvec; v += blockDim.x) dst4[v] = __ldg(src4 + v);
            done_f = nvec << 2;
        }
        for (int f = done_f + tid; f < nfloat; f += blockDim.x) s_act[f] = __ldg(src + f);
    }
    __syncthreads();

    const int i = block_base + tid;
    const bool valid = i < n_env;
    int draws = 0;
    if (valid) {
        const uint32_t gid = a.env.env_id_base + (uint32_t)i;
        const int cid = a.env.d_cond_id ? a.env.d_cond_id[i] : (int)(gid % (uint32_t)a.env.n_cond);
        const erlfill_condition c = a.env.d_cond[cid];
        const uint32_t k0 = (uint32_t)a.env.seed, k1 = (uint32_t)(a.env.seed >> 32);

        // clip (np.clip == min(max(a, lo), hi)) and int() truncation toward zero
        float clipped[ERLFILL_ACTION_DIM];
#pragma unroll
        for (int j = 0; j < ERLFILL_ACTION_DIM; ++j) {
            float v = s_act[tid * ERLFILL_ACTION_DIM + j];
            v = v < c.lo[j] ? c.lo[j] : v;
            v = v > c.hi[j] ? c.hi[j] : v;
            clipped[j] = v;
        }
        SimParams p;
        p.fast_w = (double)(int)clipped[0];
        p.medium_w = (double)(int)clipped[1];
        p.slow_w = (double)(int)clipped[2];
        p.fast_pos = (double)(int)clipped[3];
        p.medium_pos = (double)(int)clipped[4];
        p.slow_pos = (double)(int)clipped[5];
        p.fast_delay = (int)clipped[6];
        p.medium_delay = (int)clipped[7];
        p.slow_delay = (int)clipped[8];
        const int unload_delay = (int)clipped[9];

        // noise for this (env, step): misc lanes = {system_delay, reset u0, reset u1, -}
        const uint4 misc = philox4x32_10(k0, k1, gid, a.step_idx, 0u, STREAM_MISC);
        double sysdelay;
        if (kInjected) sysdelay = a.noise_rows > 0 ? __ldg(a.d_u + i) : 0.0;
        else sysdelay = 15.0 + 35.0 * u32_to_unit(misc.x);               // uniform(15, 50): :137

        const SimResult sim = simulate<kInjected>(p, s_tab, sysdelay, k0, k1, gid, a.step_idx,
                                                  kInjected ? a.d_u + i : nullptr, n_env, a.noise_rows);
        draws = sim.draws;
        const double total_time = (double)(sim.tick + unload_delay) / 1000.0;   // :239
        const double fw = sim.final_weight;

        double we, te;
        const float reward = a.env.kind == ERLFILL_ENV_MULTI ? reward_multi(c, clipped, fw, total_time, we, te)
                                                            : reward_single(c, clipped, fw, total_time, we, te);
        const float2 ns = make_float2((float)(we / c.target_weight_err), (float)(te / c.target_time));

        // recent_errors.append + done rule (MutiConditionEnv.py:445-452)
        int step_counter = a.env.d_step_counter[i];
        int ring_len = a.env.d_ring_len[i], ring_head = a.env.d_ring_head[i];
        double *ring = a.env.d_ring + i;
        if (ring_len < ERLFILL_RING) {
            ring[(size_t)((ring_head + ring_len) % ERLFILL_RING) * n_env] = we;
            ring_len += 1;
        } else {
            ring[(size_t)ring_head * n_env] = we;
            ring_head = (ring_head + 1) % ERLFILL_RING;
        }
        bool done = step_counter >= a.env.max_steps;
        if (!done && step_counter >= a.env.max_steps / 2 && ring_len == ERLFILL_RING) {
            // np.mean over the deque, oldest first: pairwise_sum for n = 20
            double r8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r8[j] = ring[(size_t)((ring_head + j) % ERLFILL_RING) * n_env];
#pragma unroll
            for (int j = 0; j < 8; ++j) r8[j] += ring[(size_t)((ring_head + 8 + j) % ERLFILL_RING) * n_env];
            double res = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
#pragma unroll
            for (int j = 16; j < 20; ++j) res += ring[(size_t)((ring_head + j) % ERLFILL_RING) * n_env];
            done = (res / 20.0) < 25.0;
        }
        step_counter += 1;

        float2 obs = ns;
        if (done && a.auto_reset) {
            double u0, u1;
            if (kInjected && a.d_reset_u) { u0 = a.d_reset_u[i]; u1 = a.d_reset_u[n_env + i]; }
            else { u0 = u32_to_unit(misc.y); u1 = u32_to_unit(misc.z); }
            obs = reset_state_cold(c, u0, u1);
            step_counter = 0; ring_len = 0; ring_head = 0;
            a.env.d_episode_counter[i] += 1;
        }
        a.env.d_step_counter[i] = step_counter;
        a.env.d_ring_len[i] = ring_len;
        a.env.d_ring_head[i] = ring_head;
        reinterpret_cast<float2 *>(a.env.d_obs)[i] = obs;
        reinterpret_cast<float2 *>(a.out.d_next_state)[i] = ns;
        a.out.d_reward[i] = reward;
        a.out.d_done[i] = done ? 1 : 0;
        if (a.out.d_final_weight) a.out.d_final_weight[i] = fw;
        if (a.out.d_total_time) a.out.d_total_time[i] = total_time;
        if (a.out.d_ticks) a.out.d_ticks[i] = sim.draws;
    }
    if (a.out.d_tick_sum) {
        __syncwarp();
        const int s = __reduce_add_sync(0xffffffffu, draws);
        if ((tid & 31) == 0 && s) atomicAdd(a.out.d_tick_sum, (unsigned long long)s);
    }
}

__global__ void env_reset_kernel(const erlfill_env_batch env, const uint8_t *d_mask, uint32_t step_idx,
                                 const double *d_reset_u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= env.n_env) return;
    if (d_mask && !d_mask[i]) return;
    const uint32_t gid = env.env_id_base + (uint32_t)i;
    const int cid = env.d_cond_id ? env.d_cond_id[i] : (int)(gid % (uint32_t)env.n_cond);
    const erlfill_condition c = env.d_cond[cid];
    double u0, u1;
    if (d_reset_u) { u0 = d_reset_u[i]; u1 = d_reset_u[env.n_env + i]; }
    else {
        const uint4 misc = philox4x32_10((uint32_t)env.seed, (uint32_t)(env.seed >> 32), gid, step_idx, 0u, STREAM_MISC);
        u0 = u32_to_unit(misc.y); u1 = u32_to_unit(misc.z);
    }
    reinterpret_cast<float2 *>(env.d_obs)[i] = reset_state_cold(c, u0, u1);
    env.d_step_counter[i] = 0;
    env.d_episode_counter[i] += 1;
    env.d_ring_len[i] = 0;
    env.d_ring_head[i] = 0;
}

__global__ void env_reset_replay_kernel(const erlfill_env_batch env, const uint8_t *d_mask,
                                        const float *d_replay_states, const int64_t *d_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= env.n_env) return;
    if (d_mask && !d_mask[i]) return;
    const uint32_t gid = env.env_id_base + (uint32_t)i;
    const int cid = env.d_cond_id ? env.d_cond_id[i] : (int)(gid % (uint32_t)env.n_cond);
    const erlfill_condition c = env.d_cond[cid];
    // np.mean(list of 10 f32[2], axis=0): row-sequential binary32 adds, then /10 (MutiConditionEnv.py:147-148)
    const float2 *st = reinterpret_cast<const float2 *>(d_replay_states);
    float2 acc = st[d_idx[(size_t)i * 10]];
    for (int j = 1; j < 10; ++j) {
        const float2 v = st[d_idx[(size_t)i * 10 + j]];
        acc.x += v.x; acc.y += v.y;
    }
    const float raw0 = acc.x / 10.f, raw1 = acc.y / 10.f;
    reinterpret_cast<float2 *>(env.d_obs)[i] =
        make_float2(raw0 / (float)c.target_weight_err, raw1 / (float)c.target_time);   // normalised again: :151-154
    env.d_step_counter[i] = 0;
    env.d_episode_counter[i] += 1;
    env.d_ring_len[i] = 0;
    env.d_ring_head[i] = 0;
}

static int pick_block(int n_env) {
    // one thread per env; spread small batches over all 148 SMs (32-thread CTAs), use fuller
    // CTAs once every SM already has several warps
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long warps = (n_env + 31) / 32;
    if (warps <= (long)sms * 4) return 32;
    if (warps <= (long)sms * 16) return 64;
    return 128;
}

static bool env_ok(const erlfill_env_batch *e) {
    return e && e->n_env > 0 && e->n_cond > 0 && e->d_cond && e->d_step_counter && e->d_episode_counter &&
           e->d_ring_len && e->d_ring_head && e->d_ring && e->d_obs;
}

}  // namespace erlfill

using namespace erlfill;

extern "C" {

int erlfill_version(void) { return ERLFILL_VERSION; }
int erlfill_last_cuda_error(void) { return g_last_cuda_error; }
const char *erlfill_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)g_last_cuda_error); }

int erlfill_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return (major == 10 && minor == 0) ? 1 : 0;
}

int erlfill_env_reset(const erlfill_env_batch *env, const uint8_t *d_mask, uint32_t step_idx,
                      const erlfill_noise *noise, void *stream) {
    if (!env_ok(env)) return ERLFILL_ERR_ARG;
    const int block = 128, grid = (env->n_env + block - 1) / block;
    env_reset_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*env, d_mask, step_idx,
                                                                noise ? noise->d_reset_u : nullptr);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_env_reset_from_replay(const erlfill_env_batch *env, const uint8_t *d_mask,
                                  const float *d_replay_states, const int64_t *d_idx, void *stream) {
    if (!env_ok(env) || !d_replay_states || !d_idx) return ERLFILL_ERR_ARG;
    const int block = 128, grid = (env->n_env + block - 1) / block;
    env_reset_replay_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*env, d_mask, d_replay_states, d_idx);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_env_step(const erlfill_env_batch *env, const float *d_actions, uint32_t step_idx, int auto_reset,
                     const erlfill_noise *noise, const erlfill_step_out *out, void *stream) {
    if (!env_ok(env) || !d_actions || !out || !out->d_next_state || !out->d_reward || !out->d_done)
        return ERLFILL_ERR_ARG;
    if (env->kind != ERLFILL_ENV_MULTI && env->kind != ERLFILL_ENV_SINGLE) return ERLFILL_ERR_ARG;
    const int rc = ensure_motor_table();
    if (rc != ERLFILL_OK) return rc;
    StepArgs a;
    a.env = *env;
    a.out = *out;
    a.d_actions = d_actions;
    a.d_u = noise ? noise->d_u : nullptr;
    a.d_reset_u = noise ? noise->d_reset_u : nullptr;
    a.noise_rows = noise ? noise->rows : 0;
    a.step_idx = step_idx;
    a.auto_reset = auto_reset;
    a.actions_aligned16 = ((uintptr_t)d_actions & 15u) == 0;
    const int block = pick_block(env->n_env);
    const int grid = (env->n_env + block - 1) / block;
    if (a.d_u) env_step_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(a);
    else env_step_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(a);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
