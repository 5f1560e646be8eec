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
constexpr int kTabShared = 512;
__device__ double2 g_motor_tab[kTabGlobal];

// Motor trajectories from a hold state: g_traj[(A*101 + B)*64 + j] = opening after tick j+1 of the move from
// opening A (speed 0) to target B, both integers 0..100; g_traj_len = number of ticks up to and including the
// snap to B (0: no usable entry -- A == B, an opening below zero on the way, or k outside the motor table).
constexpr int kTrajDim = 101;
constexpr int kTrajLen = 64;
__device__ double g_traj[kTrajDim * kTrajDim * kTrajLen];
__device__ uint8_t g_traj_len[kTrajDim * kTrajDim];

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
    // trajectories: the generic tick's motor update (generic_tick below) iterated on the host with the same table
    static double host_traj[kTrajDim * kTrajDim * kTrajLen];
    static uint8_t host_len[kTrajDim * kTrajDim];
    for (int A = 0; A < kTrajDim; ++A) {
        for (int B = 0; B < kTrajDim; ++B) {
            double *tr = host_traj + (size_t)(A * kTrajDim + B) * kTrajLen;
            uint8_t &len = host_len[A * kTrajDim + B];
            len = 0;
            for (int j = 0; j < kTrajLen; ++j) tr[j] = 0.0;
            if (A == B) continue;
            double open = (double)A, dd = 0.0;
            const double target = (double)B;
            int k = 0, m = 0;
            bool ok = true, snapped = false;
            while (m < kTrajLen) {
                const double diff = open - target;
                const bool far = fabs(diff) > 1.0, fwd = diff < 0;
                const bool accel = fwd ? (open < target - dd) : (open > dd);
                k = far ? k + (accel ? 50 : -1) : 0;
                if (k < 0 || k >= kTabGlobal) { ok = false; break; }
                const double base = far ? open : target;
                dd = host_tab[k].x;
                open = base + (fwd ? host_tab[k].y : -host_tab[k].y);
                if (open < 0.0) { ok = false; break; }
                tr[m++] = open;
                if (!far) { snapped = true; break; }
            }
            if (ok && snapped && m < kTrajLen) {
                len = (uint8_t)m;
                tr[kTrajLen - 1] = (double)m;      // the row carries its own length (the prefetched copy needs no second load)
            } else {
                tr[kTrajLen - 1] = 0.0;
            }
        }
    }
    ERLFILL_CUDA_TRY(cudaMemcpyToSymbol(g_traj, host_traj, sizeof(host_traj)));
    ERLFILL_CUDA_TRY(cudaMemcpyToSymbol(g_traj_len, host_len, sizeof(host_len)));
    g_tab_ready[dev] = true;
    return ERLFILL_OK;
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
    const uint32_t *d_step_idx;  // != nullptr: the lock-step index lives in device memory (graph replay), step_idx is ignored
    int auto_reset;
    int actions_aligned16;
    int warp_fast_mode;        // SimParams::fast_mode of the warp-per-env kernel
    __device__ __forceinline__ uint32_t step() const { return d_step_idx ? __ldg(d_step_idx) : step_idx; }
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
// The tick loop.
//
// Per-lane state: w, open, prev (binary64), k = speed/100 (int), the current target opening and the
// deceleration distance dd(k).  Every lane is in one of two phases:
//   MOVING  the motor is travelling to a new target: full branch-free motor model per tick
//   HOLD    |open - target| <= 1 has snapped open to target with speed 0 (about 85 % of all ticks):
//           the motor model is the identity, so a tick is  w += target * (1 + u)  plus the stage test
// A warp alternates between a moving loop and a hold loop; both consume the noise in blocks of four
// ticks (one Philox-4x32-10 call), and lanes are re-aligned to a block boundary at the top of the
// outer loop so that the generator runs warp-synchronously.  The arithmetic of every tick is the
// reference's, in the reference's order (VirtualWeightController.py:85-125, 172-229).
// ------------------------------------------------------------------------------------------
struct SimParams {
    double fast_w, medium_w, slow_w;
    double fast_pos, medium_pos, slow_pos;
    int fast_delay, medium_delay, slow_delay;
    bool sane;      // all three opening targets >= 0: the hold phase can be used
    int fast_mode;  // warp kernel only: 0 = multi-window fast runs on, 1 = off, 2 = computed but rejected (tests the redo path)
};

struct Lane {
    double w, open, prev, target, dd;
    int k, nochg;
};

__device__ __forceinline__ long long dbits(double x) { return __double_as_longlong(x); }

__device__ __noinline__ double2 motor_entry_slow(int k) {
    if ((unsigned)k < (unsigned)kTabGlobal) return g_motor_tab[k];
    // outside the table: evaluate on the device (dt*dt instead of libm pow; see the table comment)
    const double speed = 100.0 * (double)k;
    const double dt = speed > 0 ? __ddiv_rn(speed, 100000.0) : 0.0;
    double2 e;
    e.x = __dmul_rn(50000.0, __dmul_rn(dt, dt));
    e.y = __ddiv_rn(speed, 1000.0);          // signed: a negative speed moves against `dir`
    return e;
}

// One full tick (motor + weight + stage + safety stops).  n_after = number of draws consumed
// including this one.  Returns true when the run has finished.
template <bool kFirst, int kTabN = kTabShared>
__device__ __forceinline__ bool generic_tick(Lane &s, const SimParams &p, const double2 *s_tab, double opu,
                                             double sysdelay, int n_after, int &delay) {
    // motor_simulation(target): VirtualWeightController.py:92-113
    const double diff = s.open - s.target;
    const long long db = dbits(diff);
    const bool far = (db & 0x7fffffffffffffffLL) > 0x3ff0000000000000LL;   // abs(open - target) > 1
    const bool fwd = db < 0;                                               // open < target
    const bool p1 = s.open < s.target - s.dd;
    const bool p2 = s.open > s.dd;
    const bool accel = fwd ? p1 : p2;
    int k = s.k + (accel ? 50 : -1);                                       // += 5000 / -= 100
    k = far ? k : 0;
    const double base = far ? s.open : s.target;
    double2 e;
    if ((unsigned)k < (unsigned)kTabN) e = s_tab[k];
    else e = motor_entry_slow(k);
    s.k = k;
    s.dd = e.x;
    s.open = base + (fwd ? e.y : -e.y);                                    // += (speed*dir)/1000
    // weight_simulation(open): :122-125
    double w = __dadd_rn(s.w, __dmul_rn(s.open, opu));
    w = (w < 0) ? 0.0 : w;
    // stage selection :191-213 (w is never -0.0, so the tests are exact on the bit patterns too)
    const bool lf = w < p.fast_w, lm = w < p.medium_w, ls = w < p.slow_w;
    s.target = lf ? p.fast_pos : (lm ? p.medium_pos : p.slow_pos);
    if (kFirst) {
        // the one-time stage delay is taken on the first tick only (the flag is never cleared)
        if (lf | lm | ls) {
            delay = lf ? p.fast_delay : (lm ? p.medium_delay : p.slow_delay);
            w = __dadd_rn(w, __dmul_rn(s.target, sysdelay));
        }
    }
    s.w = w;
    if (!(lf | lm | ls)) return true;                                      // normal finish
    // safety stops :216-229 (iteration == n_after here)
    s.nochg = (fabs(w - s.prev) < 0.01) ? s.nochg + 1 : 0;
    s.prev = w;
    return (n_after > ERLFILL_MAX_ITER) | (s.nochg > ERLFILL_NOCHANGE_LIMIT);
}

// 1.0 + (-1.0 + 2*U) == x * 2^-31 exactly for U = x * 2^-32: place x in the top mantissa bits of
// a double in [2,4) and subtract 2.
__device__ __forceinline__ double opu_from_u32(uint32_t x) {
    return __hiloint2double((int)(0x40000000u | (x >> 12)), (int)(x << 20)) - 2.0;
}

struct PhiloxNoise {
    uint32_t k0, k1, gid, step_idx;
    __device__ __forceinline__ void fetch_raw(uint4 &r, int block) const {
        r = philox4x32_10(k0, k1, gid, step_idx, (uint32_t)block, STREAM_TICK);
    }
    __device__ __forceinline__ void convert(double (&o)[4], const uint4 &r) const {
        o[0] = opu_from_u32(r.x); o[1] = opu_from_u32(r.y); o[2] = opu_from_u32(r.z); o[3] = opu_from_u32(r.w);
    }
    __device__ __forceinline__ void fetch(double (&o)[4], int block) const {
        uint4 r;
        fetch_raw(r, block);
        convert(o, r);
    }
};
struct InjectedNoise {
    const double *u_col;   // column of this env in the [rows][n_env] matrix; row 1+t = tick t
    int n_env, rows;
    __device__ __forceinline__ void fetch(double (&o)[4], int block) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = 1 + 4 * block + j;
            o[j] = 1.0 + (row < rows ? __ldg(u_col + (size_t)row * n_env) : 0.0);
        }
    }
    // the "raw" form of a block is just its index; the loads happen in convert()
    __device__ __forceinline__ void fetch_raw(uint4 &r, int block) const { r.x = (uint32_t)block; }
    __device__ __forceinline__ void convert(double (&o)[4], const uint4 &r) const { fetch(o, (int)r.x); }
};

__device__ __forceinline__ bool in_hold(const Lane &s, const SimParams &p) {
    return p.sane & (s.k == 0) & (dbits(s.open) == dbits(s.target));
}

// `valid` == false lanes only take part in the warp votes.
template <class NoiseT>
__device__ __forceinline__ SimResult simulate(const SimParams &p, const double2 *s_tab, double sysdelay,
                                              const NoiseT &nz, bool valid) {
    Lane s;
    s.w = 0.0; s.open = 0.0; s.prev = 0.0; s.target = p.fast_pos; s.dd = 0.0; s.k = 0; s.nochg = 0;
    double o[4] = {0.0, 0.0, 0.0, 0.0};
    int n = 0, delay = 0;
    bool fin = !valid;
    if (!fin) {
        nz.fetch(o, 0);
        fin = generic_tick<true>(s, p, s_tab, o[0], sysdelay, 1, delay);
        n = 1;
    }
    for (;;) {
        if (__all_sync(0xffffffffu, fin)) break;
        // (1) re-align to a noise-block boundary with the values still cached in o[]
        while (!fin && (n & 3)) {
            const int j = n & 3;
            const double opu = j == 1 ? o[1] : (j == 2 ? o[2] : o[3]);
            n += 1;
            fin = generic_tick<false>(s, p, s_tab, opu, sysdelay, n, delay);
        }
        // (2) moving phase, four ticks per noise block
        while (!fin && !in_hold(s, p)) {
            nz.fetch(o, n >> 2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!fin) {
                    n += 1;
                    fin = generic_tick<false>(s, p, s_tab, o[j], sysdelay, n, delay);
                }
            }
        }
        // (3) hold phase: open == target >= 0, speed 0, so  w += target * opu  and w never decreases;
        //     the lane stays in its stage exactly while w < (that stage's threshold)
        if (!fin && in_hold(s, p)) {
            const double tgt = s.target;
            double w = s.w;
            const bool lf = w < p.fast_w, lm = w < p.medium_w;
            const long long thr = dbits(lf ? p.fast_w : (lm ? p.medium_w : p.slow_w));
            bool leave = false;
            uint4 raw;
            nz.fetch_raw(raw, n >> 2);
            while (!leave) {
                nz.convert(o, raw);
                // prefetch the next block's noise: independent of the weight chain below, so the
                // generator's serial rounds overlap the FP64 adds even with one warp per scheduler
                nz.fetch_raw(raw, (n >> 2) + 1);
                // Speculate on the whole 4-tick block: w never decreases here, so the stage test on
                // the last value covers all four.  An increment >= 0.02 makes |w - w_old| > 0.01
                // certain; smaller ones get the exact test.  Stage exits, the iteration cap and a
                // stagnation count near its limit replay the block tick by tick below.
                const double i0 = __dmul_rn(tgt, o[0]), i1 = __dmul_rn(tgt, o[1]);
                const double i2 = __dmul_rn(tgt, o[2]), i3 = __dmul_rn(tgt, o[3]);
                const double w1 = __dadd_rn(w, i0), w2 = __dadd_rn(w1, i1);
                const double w3 = __dadd_rn(w2, i2), w4 = __dadd_rn(w3, i3);
                const int imin = min(min(__double2hiint(i0), __double2hiint(i1)),
                                     min(__double2hiint(i2), __double2hiint(i3)));
                if ((dbits(w4) < thr) & (n + 4 <= ERLFILL_MAX_ITER) & (s.nochg + 4 <= ERLFILL_NOCHANGE_LIMIT)) {
                    int nc = 0;
                    if (imin <= 0x3F947AE1) {
                        const bool c0 = fabs(w1 - w) < 0.01, c1 = fabs(w2 - w1) < 0.01;
                        const bool c2 = fabs(w3 - w2) < 0.01, c3 = fabs(w4 - w3) < 0.01;
                        nc = c3 ? (c2 ? (c1 ? (c0 ? s.nochg + 4 : 3) : 2) : 1) : 0;
                    }
                    s.nochg = nc;
                    w = w4;
                    n += 4;
                    continue;
                }
#pragma unroll 1
                for (int j = 0; j < 4 && !leave; ++j) {
                    const double w_old = w;
                    const double opu = j == 0 ? o[0] : (j == 1 ? o[1] : (j == 2 ? o[2] : o[3]));
                    w = __dadd_rn(w, __dmul_rn(tgt, opu));
                    n += 1;
                    if (dbits(w) >= thr) {
                        // left the stage: full stage evaluation (:191-213), then hand over to (1)/(2)
                        const bool f = w < p.fast_w, m = w < p.medium_w, sl = w < p.slow_w;
                        s.target = f ? p.fast_pos : (m ? p.medium_pos : p.slow_pos);
                        leave = true;
                        if (!(f | m | sl)) { fin = true; break; }
                    }
                    s.nochg = (fabs(w - w_old) < 0.01) ? s.nochg + 1 : 0;
                    if ((n > ERLFILL_MAX_ITER) | (s.nochg > ERLFILL_NOCHANGE_LIMIT)) { fin = true; leave = true; }
                }
            }
            s.w = w;
            s.prev = w;
        }
    }
    SimResult res;
    res.final_weight = s.w;
    res.tick = n + delay;
    res.draws = n;
    return res;
}

// Row N1 — the "on-board" per-tick models of the env classes (WeightEnv.py:176-229, MutiConditionEnv.py:184-226) under the stage
// logic of simulate_run (in the reference that logic runs in an external PLC polled over Modbus, WeightEnv.py:245-331; the offline
// controller's loop is the scripted stand-in, exactly what oracle/gen_golden.py onboard runs).  Differences from generic_tick:
//   motor:  decelerate_time = speed / 100000 with no `speed > 0` guard; np.power(dt, 2) == dt * dt for every reachable speed
//           (tests/test_oracle_golden.py); the speed CAP `if current_speed < 10000: current_speed += 5000`
//   weight: current_weight = int(current_weight + opening * (1.0 + u)) — truncation toward zero every tick, no clamp at zero
// int() truncation breaks the exact-integer hold phase of the warp kernel, so this variant is one thread per env, one tick at a time.
template <class NoiseT>
__device__ __forceinline__ SimResult simulate_onboard(const SimParams &p, double sysdelay, const NoiseT &nz, bool valid) {
    double w = 0.0, open = 0.0, speed = 0.0, dir = 0.0, prev = 0.0, target = p.fast_pos;
    int n = 0, delay = 0, nochg = 0;
    bool first = true, fin = !valid;
    double o[4] = {0.0, 0.0, 0.0, 0.0};
    while (!fin) {
        if ((n & 3) == 0) nz.fetch(o, n >> 2);
        const int j = n & 3;
        const double opu = j == 0 ? o[0] : (j == 1 ? o[1] : (j == 2 ? o[2] : o[3]));     // 1.0 + u
        n += 1;
        // motor_simulation(target): WeightEnv.py:188-218
        const double dt = __ddiv_rn(speed, 100000.0);
        const double dd = __dmul_rn(50000.0, __dmul_rn(dt, dt));
        if (fabs(__dsub_rn(open, target)) > 1.0) {
            if (open < target) {
                dir = 1.0;
                if (open < __dsub_rn(target, dd)) { if (speed < 10000.0) speed = __dadd_rn(speed, 5000.0); }
                else speed = __dsub_rn(speed, 100.0);
            } else if (open > target) {
                dir = -1.0;
                if (open > dd) { if (speed < 10000.0) speed = __dadd_rn(speed, 5000.0); }
                else speed = __dsub_rn(speed, 100.0);
            }
        } else {
            speed = 0.0; open = target;
        }
        open = __dadd_rn(open, __ddiv_rn(__dmul_rn(speed, dir), 1000.0));
        // weight_simulation(open, 1.0): :228-229
        w = trunc(__dadd_rn(w, __dmul_rn(open, opu)));
        // stage selection, first-tick delay, safety stops: VirtualWeightController.py:191-229
        const bool lf = w < p.fast_w, lm = w < p.medium_w, ls = w < p.slow_w;
        if (!(lf | lm | ls)) break;
        target = lf ? p.fast_pos : (lm ? p.medium_pos : p.slow_pos);
        if (first) {
            delay = lf ? p.fast_delay : (lm ? p.medium_delay : p.slow_delay);
            w = __dadd_rn(w, __dmul_rn(target, sysdelay));
            first = false;
        }
        nochg = (fabs(__dsub_rn(w, prev)) < 0.01) ? nochg + 1 : 0;
        prev = w;
        fin = (n > ERLFILL_MAX_ITER) | (nochg > ERLFILL_NOCHANGE_LIMIT);
    }
    SimResult res;
    res.final_weight = w;
    res.tick = n + delay;
    res.draws = n;
    return res;
}

// Reward, state, done rule, auto-reset and all stores of one env (E2/E3/E4 after the simulator run).
template <bool kInjected>
__device__ __forceinline__ void finish_env(const StepArgs &a, int i, const erlfill_condition &c, const float *clipped,
                                           const SimResult &sim, int unload_delay, const uint4 &misc) {
    const int n_env = a.env.n_env;
    const double total_time = (double)(sim.tick + unload_delay) / 1000.0;   // :239
    const double fw = sim.final_weight;

    double we, te;
    const float reward = (a.env.kind & 1) == ERLFILL_ENV_MULTI ? reward_multi(c, clipped, fw, total_time, we, te)
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

constexpr int kMaxBlock = 256;

template <bool kInjected, bool kOnboard = false>
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
    const int ic = valid ? i : n_env - 1;                    // padding lanes mirror the last env, never store
    const uint32_t gid = a.env.env_id_base + (uint32_t)ic;
    const int cid = a.env.d_cond_id ? a.env.d_cond_id[ic] : (int)(gid % (uint32_t)a.env.n_cond);
    const erlfill_condition c = a.env.d_cond[cid];
    const uint32_t k0 = (uint32_t)a.env.seed, k1 = (uint32_t)(a.env.seed >> 32);

    // clip (np.clip == min(max(a, lo), hi)) and int() truncation toward zero
    float clipped[ERLFILL_ACTION_DIM];
#pragma unroll
    for (int j = 0; j < ERLFILL_ACTION_DIM; ++j) {
        float v = s_act[(valid ? tid : 0) * ERLFILL_ACTION_DIM + j];
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
    p.sane = (p.fast_pos >= 0) & (p.medium_pos >= 0) & (p.slow_pos >= 0);
    const int unload_delay = (int)clipped[9];

    // noise for this (env, step): misc lanes = {system_delay, reset u0, reset u1, -}
    const uint4 misc = philox4x32_10(k0, k1, gid, a.step(), 0u, STREAM_MISC);
    double sysdelay;
    SimResult sim;
    if (kInjected) {
        sysdelay = a.noise_rows > 0 ? __ldg(a.d_u + ic) : 0.0;
        InjectedNoise nz;
        nz.u_col = a.d_u + ic; nz.n_env = n_env; nz.rows = a.noise_rows;
        sim = kOnboard ? simulate_onboard(p, sysdelay, nz, valid) : simulate(p, s_tab, sysdelay, nz, valid);
    } else {
        sysdelay = 15.0 + 35.0 * u32_to_unit(misc.x);                 // uniform(15, 50): :137
        PhiloxNoise nz;
        nz.k0 = k0; nz.k1 = k1; nz.gid = gid; nz.step_idx = a.step();
        sim = kOnboard ? simulate_onboard(p, sysdelay, nz, valid) : simulate(p, s_tab, sysdelay, nz, valid);
    }
    const int draws = valid ? sim.draws : 0;
    if (valid) finish_env<kInjected>(a, i, c, clipped, sim, unload_delay, misc);
    if (a.out.d_tick_sum) {
        __syncwarp();
        const int s = __reduce_add_sync(0xffffffffu, draws);
        if ((tid & 31) == 0 && s) atomicAdd(a.out.d_tick_sum, (unsigned long long)s);
    }
}

// ------------------------------------------------------------------------------------------
// Warp-per-env kernel (Philox noise).  One warp runs one env; the 32 lanes are 32 consecutive
// Philox noise blocks (128 ticks), so nothing diverges and the generator runs 32-wide.
//
// HOLD phase (about 98 % of all ticks: open == target, an integer 0..255, speed 0), in exact
// integer arithmetic.  A tick adds inc = target * (1 + u) with 1 + u = x * 2^-31 (x the tick's
// 32-bit Philox word), so inc = P * 2^-31 with the integer P = target * x < 2^40: the binary64
// product is exact.  While w stays inside its binade [2^e, 2^(e+1)), e <= 20, both w and every
// inc are multiples of ulp(w) = 2^(e-52) and the sums stay below 2^(e+1), so every binary64 add of
// the reference is exact too and  w_j = w_0 + 2^-31 * sum_{t<=j} P_t  holds bit for bit.  The
// stage test w_j < thr therefore becomes the integer test  prefix_j < ceil((lim - w_0) * 2^31)
// with lim = min(thr, 2^(e+1)) (lim - w_0 is exact: both are multiples of ulp(w_0) and the
// difference is below 2^e).  A 128-tick super-block is then: one Philox call and four 64-bit
// multiply/adds per lane, one warp reduction, one compare.  Only the tick that reaches `lim`
// (stage exit, or w entering the next binade, where the reference's add may round) is replayed
// with a real binary64 add from the exact predecessor value; the stagnation counter
// (|dw| < 0.01  <=>  P <= 21474836) is kept with a warp max-reduction over tick indices.
// MOVING phase / first tick / anything outside those preconditions: the generic tick of the
// thread-per-env kernel, executed redundantly by all lanes with the noise word shuffled from the
// lane that holds its block.
// ------------------------------------------------------------------------------------------
constexpr unsigned kFullMask = 0xffffffffu;
constexpr unsigned long long kSmallIncMax = 21474836ull;   // P * 2^-31 < 0.01  <=>  P <= 21474836
constexpr int kNoiseBlocks = 2;                            // Philox blocks per lane
constexpr int kWindow = 128 * kNoiseBlocks;                // draws held by the warp

struct WarpNoise {
    uint32_t k0, k1, gid, step_idx;
    int b0;         // the warp holds the draw window [4*b0, 4*b0 + kWindow): lane L has noise blocks b0 + 32*h + L
    uint4 raw[kNoiseBlocks];
    uint32_t *sh;   // the same window in shared memory, word i = draw 4*b0 + i (random access for the motor moves and generic ticks)
#ifdef ERLFILL_SIM_PROF
    long long pf_fast = 0, pf_hold = 0, pf_cross = 0, pf_move = 0, pf_gen = 0;
    int pf_fastwin = 0, pf_win = 0, pf_nhold = 0, pf_nmove = 0, pf_ngen = 0, pf_movefail = 0;
#endif
    __device__ __forceinline__ void fill(int block0, int lane) {
        b0 = block0;
#pragma unroll
        for (int h = 0; h < kNoiseBlocks; ++h)
            raw[h] = philox4x32_10(k0, k1, gid, step_idx, (uint32_t)(block0 + 32 * h + lane), STREAM_TICK);
        __syncwarp();                                   // every lane is done reading the previous window
#pragma unroll
        for (int h = 0; h < kNoiseBlocks; ++h) reinterpret_cast<uint4 *>(sh)[32 * h + lane] = raw[h];
        __syncwarp();
    }
    // Philox word of draw t (0-based, inside the window); t may differ per lane
    __device__ __forceinline__ uint32_t word(int t) const {
        const int rel = t - 4 * b0;
        return (unsigned)rel < (unsigned)kWindow ? sh[rel] : 0u;
    }
};

__device__ __forceinline__ unsigned long long warp_sum_u44(unsigned long long v) {   // v < 2^44 per lane
    return (unsigned long long)__reduce_add_sync(kFullMask, (unsigned)v & 0x3FFFFFu) +
           ((unsigned long long)__reduce_add_sync(kFullMask, (unsigned)(v >> 22)) << 22);
}

__device__ __forceinline__ unsigned long long warp_sum_u59(unsigned long long v) {   // v < 2^59 per lane
    const unsigned long long lo = (unsigned long long)__reduce_add_sync(kFullMask, (unsigned)v & 0x3FFFFFu);
    const unsigned long long mid = (unsigned long long)__reduce_add_sync(kFullMask, (unsigned)(v >> 22) & 0x3FFFFFu);
    const unsigned long long hi = (unsigned long long)__reduce_add_sync(kFullMask, (unsigned)(v >> 44));
    return lo + (mid << 22) + (hi << 44);
}

// sum of this lane's increments P = ti * x over the 4 draws of one noise block; draws with window index < skip
// (already consumed) count as 0
template <bool kMasked>
__device__ __forceinline__ unsigned long long block_sum(uint32_t ti, const uint4 &r, int t0, int skip) {
    unsigned long long a = (unsigned long long)ti * r.x, b = (unsigned long long)ti * r.y;
    unsigned long long c = (unsigned long long)ti * r.z, d = (unsigned long long)ti * r.w;
    if (kMasked) {
        a = t0 + 0 >= skip ? a : 0ull; b = t0 + 1 >= skip ? b : 0ull;
        c = t0 + 2 >= skip ? c : 0ull; d = t0 + 3 >= skip ? d : 0ull;
    }
    return (a + b) + (c + d);
}

// ---- hold phase ---------------------------------------------------------------------------
// While w stays in one binade the reference's adds are exact (see the header of this section).  When w enters
// the next binade its ulp doubles and the one add that crosses may round (a tie, round-half-even).  With
// A = w/ulp0 the anchor in units of the ORIGINAL ulp (a multiple of 2^c after c crossings) the exact sum at
// crossing c+1 is congruent to A modulo 4*2^c as long as the new binade is below 2^21 (the increments,
// multiples of 2^-31, are multiples of 4*2^c*ulp0 there), so the tie resolves the same way whichever tick
// crosses: A -> A -/+ 2^c when (A >> c) & 3 is 1 / 3.  The corrections of all binades below the stage
// threshold are therefore folded into one anchor up front and the whole hold stretch is the integer test
//     acc_j = sum_{t<=j} P_t  <  Dc = thr*2^31 - floor(A*ulp0*2^31)
// over as many noise windows as it takes.  Only the draw that reaches thr is replayed with a real binary64 add
// from its exact predecessor (whose anchor is the one of the binade the predecessor is in).
// One boundary per tick is guaranteed by 2*target <= 2^e0.  tests/test_hold_superblock_model.py restates this
// on the CPU against the oracle's serial loop.
__device__ __forceinline__ double pow2d(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }

__device__ __forceinline__ unsigned long long round_anchor(unsigned long long A, int c) {
    const unsigned r = (unsigned)(A >> c) & 3u;
    const unsigned long long step = 1ull << c;
    return (r & 1u) ? (r == 1u ? A - step : A + step) : A;
}

// w after `inc` (= acc * 2^-31, exact) has been added to the stretch's start value M * ulp0
__device__ __forceinline__ double hold_value(unsigned long long M, int e0, double thr, double inc) {
    const double u0 = pow2d(e0 - 52);
    unsigned long long A = M;
    int c = 0;
    while (pow2d(e0 + c + 1) < thr && __dadd_rn(__dmul_rn((double)(long long)A, u0), inc) >= pow2d(e0 + c + 1)) {
        A = round_anchor(A, c);
        ++c;
    }
    return __dadd_rn(__dmul_rn((double)(long long)A, u0), inc);
}

__device__ __forceinline__ bool hold_caps_ok(const WarpNoise &nz, int n, int nochg) {
    const int base = n >= 4 * nz.b0 + kWindow ? (n & ~3) : 4 * nz.b0;
    return (base + kWindow <= ERLFILL_MAX_ITER) & (nochg + kWindow <= ERLFILL_NOCHANGE_LIMIT);
}

// Preconditions (checked by the caller): in_hold, target an integer 0..255, 1 <= w < thr <= 2^20,
// 2*target <= 2^floor(log2 w), hold_caps_ok.
__device__ __forceinline__ void hold_stretch(Lane &s, const SimParams &p, double thr, WarpNoise &nz, int lane, int &n,
                                             bool &fin) {
    const long long wb = dbits(s.w);
    const int e0 = (int)((wb >> 52) & 0x7ff) - 1023;
    const unsigned long long M = ((unsigned long long)wb & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    unsigned long long Dc;
    {
        unsigned long long A = M;
        int c = 0;
        while (pow2d(e0 + c + 1) < thr) { A = round_anchor(A, c); ++c; }
        Dc = ((unsigned long long)(long long)thr << 31) - (A >> (21 - e0));
    }
    const uint32_t ti = (uint32_t)(int)s.target;
    const uint32_t xs = ti ? (uint32_t)(kSmallIncMax / ti) : 0xffffffffu;   // increment < 0.01  <=>  x <= xs
    const double kScale = 1.0 / 2147483648.0;
    const int tb = 4 * lane;
    unsigned long long acc = 0;
    int nochg = s.nochg;
    bool try_fast = (p.fast_mode != 1) & (ti != 0u);
    for (;;) {
        // ---- fast run (at most once per stretch): what is left of the window the warp holds plus Kr fresh rows of
        // 128 draws (one noise block per lane) that stay below thr with overwhelming probability.  The draws left
        // until thr are R = (Dc - acc) / (ti * 2^31) on average, with a standard deviation of 0.577 * sqrt(R) draws
        // (x uniform on 32 bits); the run is sized to end 5 sigma short of that.  Every lane sums its raw noise
        // words (the increments are ti * x, so one multiply at the end gives their exact sum), consecutive rows'
        // Philox chains overlap, and ONE warp sum checks acc + total < Dc: all increments are >= 0, so every prefix
        // is then below thr too.  The stagnation rule (VirtualWeightController.py:216-221) needs (i) no run of more
        // than 500 small increments inside: guaranteed when no complete noise block is all-small (every run is then
        // at most 6 + 3 draws long; nochg + kWindow <= LIMIT covers the run that continues the previous count), and
        // (ii) the count at the end: the run that ends the last row.  If the check or (i) fails nothing is committed
        // and the draws are redone window by window below (fast_mode 2 forces that in the tests).
        if (try_fast) {
            try_fast = false;
            const int wend = 4 * nz.b0 + kWindow;
            const int L = n < wend ? wend - n : 0;                   // unused draws of the current window
            const int nstart = n < wend ? wend : n;                  // first draw of the fresh rows
            if (((nstart & 3) == 0) & (nochg + kWindow <= ERLFILL_NOCHANGE_LIMIT)) {
                const float R = __fdividef((float)(long long)(Dc - acc), (float)ti * 2147483648.0f);
                int Kr = (int)((R - 3.0f * sqrtf(R) - (float)L) * (1.0f / 128.0f));
                Kr = min(Kr, (ERLFILL_MAX_ITER - nstart) >> 7);
                if (Kr >= 1) {
#ifdef ERLFILL_SIM_PROF
                    const long long pf_f0 = clock64();
#endif
                    unsigned long long lx = 0;              // sum of this lane's raw words
                    uint32_t worst = 0xffffffffu;           // min over complete blocks of max(x): <= xs <=> one is all-small
                    if (L > 0) {
                        const int skip = kWindow - L;
#pragma unroll
                        for (int h = 0; h < kNoiseBlocks; ++h) {
                            const int t0 = 128 * h + tb;
                            const uint4 r = nz.raw[h];
                            lx += (unsigned long long)(t0 + 0 >= skip ? r.x : 0u) + (t0 + 1 >= skip ? r.y : 0u);
                            lx += (unsigned long long)(t0 + 2 >= skip ? r.z : 0u) + (t0 + 3 >= skip ? r.w : 0u);
                            worst = min(worst, t0 >= skip ? max(max(r.x, r.y), max(r.z, r.w)) : 0xffffffffu);
                        }
                    }
                    const uint32_t blk0 = (uint32_t)(nstart >> 2) + (uint32_t)lane;
                    uint4 rl = make_uint4(0u, 0u, 0u, 0u);
                    int row = 0;
#pragma unroll 1
                    for (; row + 4 <= Kr; row += 4) {
                        uint4 r[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            r[q] = philox4x32_10(nz.k0, nz.k1, nz.gid, nz.step_idx, blk0 + (uint32_t)(32 * (row + q)), STREAM_TICK);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            lx += ((unsigned long long)r[q].x + r[q].y) + ((unsigned long long)r[q].z + r[q].w);
                            worst = min(worst, max(max(r[q].x, r[q].y), max(r[q].z, r[q].w)));
                        }
                        rl = r[3];
                    }
#pragma unroll 1
                    for (; row < Kr; ++row) {
                        rl = philox4x32_10(nz.k0, nz.k1, nz.gid, nz.step_idx, blk0 + (uint32_t)(32 * row), STREAM_TICK);
                        lx += ((unsigned long long)rl.x + rl.y) + ((unsigned long long)rl.z + rl.w);
                        worst = min(worst, max(max(rl.x, rl.y), max(rl.z, rl.w)));
                    }
                    const unsigned long long total = (unsigned long long)ti * warp_sum_u44(lx);   // lx < 2^40
                    const bool ok = !__any_sync(kFullMask, worst <= xs) & (acc + total < Dc) & (p.fast_mode != 2);
                    if (ok) {
                        int tl = -1;                        // last draw of the last row that is not small
                        tl = rl.x > xs ? tb : tl; tl = rl.y > xs ? tb + 1 : tl;
                        tl = rl.z > xs ? tb + 2 : tl; tl = rl.w > xs ? tb + 3 : tl;
                        nochg = 127 - __reduce_max_sync(kFullMask, tl);
                        acc += total;
                        n = nstart + 128 * Kr;
#ifdef ERLFILL_SIM_PROF
                        nz.pf_fastwin += Kr; nz.pf_fast += clock64() - pf_f0;
#endif
                    }
                }
            }
        }
        if (n >= 4 * nz.b0 + kWindow) nz.fill(n >> 2, lane);
        const int base = 4 * nz.b0;
        if ((base + kWindow > ERLFILL_MAX_ITER) | (nochg + kWindow > ERLFILL_NOCHANGE_LIMIT)) break;
        const int skip = n - base;
        unsigned long long bx[kNoiseBlocks], lane_x = 0;      // raw-word sums per noise block (increment sums / ti)
#pragma unroll
        for (int h = 0; h < kNoiseBlocks; ++h) {
            const uint4 r = nz.raw[h];
            if (skip == 0) {
                bx[h] = ((unsigned long long)r.x + r.y) + ((unsigned long long)r.z + r.w);
            } else {
                const int t0 = 128 * h + tb;
                bx[h] = ((unsigned long long)(t0 + 0 >= skip ? r.x : 0u) + (t0 + 1 >= skip ? r.y : 0u)) +
                        ((unsigned long long)(t0 + 2 >= skip ? r.z : 0u) + (t0 + 3 >= skip ? r.w : 0u));
            }
            lane_x += bx[h];
        }
        const unsigned long long total = (unsigned long long)ti * warp_sum_u44(lane_x);
        if (acc + total < Dc) {
            // every remaining draw of the window stays below thr
            acc += total;
            // stagnation counter (VirtualWeightController.py:216-221): does any real draw have an increment < 0.01?
            bool small_here;
            if (skip == 0) {
                uint32_t mn = 0xffffffffu;
#pragma unroll
                for (int h = 0; h < kNoiseBlocks; ++h)
                    mn = min(mn, min(min(nz.raw[h].x, nz.raw[h].y), min(nz.raw[h].z, nz.raw[h].w)));
                small_here = mn <= xs;
            } else {
                small_here = false;
#pragma unroll
                for (int h = 0; h < kNoiseBlocks; ++h) {
                    const uint32_t xv[4] = {nz.raw[h].x, nz.raw[h].y, nz.raw[h].z, nz.raw[h].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) small_here |= (128 * h + tb + k >= skip) & (xv[k] <= xs);
                }
            }
            if (__any_sync(kFullMask, small_here)) {
                // run of small increments at the end of the window: highest real draw that is NOT small
                int tl = -1;
#pragma unroll
                for (int h = 0; h < kNoiseBlocks; ++h) {
                    const uint32_t xv[4] = {nz.raw[h].x, nz.raw[h].y, nz.raw[h].z, nz.raw[h].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int t = 128 * h + tb + k;
                        tl = ((t >= skip) & (xv[k] > xs)) ? t : tl;
                    }
                }
                const int t_last = __reduce_max_sync(kFullMask, tl);
                nochg = t_last < 0 ? nochg + (kWindow - skip) : kWindow - 1 - t_last;
            } else {
                nochg = 0;
            }
            n = base + kWindow;
#ifdef ERLFILL_SIM_PROF
            nz.pf_win++;
#endif
            continue;
        }
        // some draw of this window reaches thr: find the noise block row, then the draw
#ifdef ERLFILL_SIM_PROF
        const long long pf_c0 = clock64();
#endif
        int hx = 0;
        unsigned long long row_before = acc;
#pragma unroll
        for (int h = 0; h < kNoiseBlocks - 1; ++h) {
            const unsigned long long th = (unsigned long long)ti * warp_sum_u44(bx[h]);
            if (hx == h && row_before + th < Dc) { row_before += th; hx = h + 1; }
        }
        uint4 r = nz.raw[0];
#pragma unroll
        for (int h = 1; h < kNoiseBlocks; ++h) r = hx == h ? nz.raw[h] : r;
        const int t0 = 128 * hx + tb;
        const unsigned long long P0 = t0 + 0 >= skip ? (unsigned long long)ti * r.x : 0ull;
        const unsigned long long P1 = t0 + 1 >= skip ? (unsigned long long)ti * r.y : 0ull;
        const unsigned long long P2 = t0 + 2 >= skip ? (unsigned long long)ti * r.z : 0ull;
        const unsigned long long P3 = t0 + 3 >= skip ? (unsigned long long)ti * r.w : 0ull;
        const unsigned long long S1 = P0, S2 = S1 + P1, S3 = S2 + P2, S4 = S3 + P3;
        unsigned long long incl = S4;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= d) incl += t;
        }
        const unsigned long long excl = row_before + (incl - S4);
        const unsigned long long c0 = excl + S1, c1 = excl + S2, c2 = excl + S3, c3 = excl + S4;
        const int kk = (int)(c0 < Dc) + (int)(c1 < Dc) + (int)(c2 < Dc) + (int)(c3 < Dc);
        const unsigned cross = __ballot_sync(kFullMask, kk < 4);
        const int Lj = __ffs((int)cross) - 1;
        const int kj = __shfl_sync(kFullMask, kk, Lj);
        const unsigned long long before_l = kk == 0 ? excl : (kk == 1 ? c0 : (kk == 2 ? c1 : c2));
        const unsigned long long Pj_l = kk == 0 ? P0 : (kk == 1 ? P1 : (kk == 2 ? P2 : P3));
        const unsigned long long before = __shfl_sync(kFullMask, before_l, Lj);
        const unsigned long long Pj = __shfl_sync(kFullMask, Pj_l, Lj);
        const int j = 128 * hx + 4 * Lj + kj;                       // window index of the draw that reaches thr
        const double w_prev = hold_value(M, e0, thr, __dmul_rn((double)(long long)before, kScale));   // exact
        const double w_new = __dadd_rn(w_prev, __dmul_rn((double)(long long)Pj, kScale));   // the reference's (rounded) add
        // stagnation run ending at draw j-1
        int tlj = -1;
#pragma unroll
        for (int h = 0; h < kNoiseBlocks; ++h) {
            const uint32_t xv[4] = {nz.raw[h].x, nz.raw[h].y, nz.raw[h].z, nz.raw[h].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = 128 * h + tb + k;
                tlj = ((t >= skip) & (t < j) & (xv[k] > xs)) ? t : tlj;
            }
        }
        const int t_last = __reduce_max_sync(kFullMask, tlj);
        const int run_before = t_last < 0 ? nochg + (j - skip) : j - 1 - t_last;
        n = base + j + 1;
        s.w = w_new;
        const bool f = w_new < p.fast_w, m = w_new < p.medium_w, sl = w_new < p.slow_w;
        if (!(f | m | sl)) { fin = true; return; }              // normal finish
        s.target = f ? p.fast_pos : (m ? p.medium_pos : p.slow_pos);
        s.nochg = (fabs(w_new - w_prev) < 0.01) ? run_before + 1 : 0;
        s.prev = w_new;
        fin = (n > ERLFILL_MAX_ITER) | (s.nochg > ERLFILL_NOCHANGE_LIMIT);
#ifdef ERLFILL_SIM_PROF
        nz.pf_cross += clock64() - pf_c0;
#endif
        return;
    }
    // a safety stop is within reach: hand the exact state to the generic ticks
    s.w = hold_value(M, e0, thr, __dmul_rn((double)(long long)acc, kScale));
    s.prev = s.w;
    s.nochg = nochg;
}

// MOVING phase from a hold state (open == A, speed 0) to a new target B, A != B integers in 0..100: the motor
// trajectory open_1..open_m (m <= 55, the last entry is the snap to B) does not depend on the noise, so it is
// read from the host-built table g_traj.  Lane j forms the tick's increment open_j * (1 + u_j) and parks it in
// shared memory; the weight is then accumulated by m dependent binary64 adds in the reference's order (every
// lane redundantly, broadcast 128-bit loads, no shuffle on the chain; entries past m are +0.0 and w >= 0, so
// adding them changes nothing).  All increments are >= 0, so w is non-decreasing and "same stage flags at both
// ends" means no stage change happened inside the phase.  Stagnation counter (VirtualWeightController.py:216-221):
// with every increment >= 0.011 each |dw| >= 0.01 and the counter is 0 afterwards; otherwise the chain is redone
// with the reference's own test on every tick.  Returns false (nothing committed) when a precondition fails; the
// caller then runs generic ticks.
__device__ __forceinline__ bool moving_phase(Lane &s, const SimParams &p, WarpNoise &nz, int lane, int &n, int &delay,
                                             bool first, double sysdelay, double *sh_p, const double *sh_traj, int pre0,
                                             int pre1) {
    const int A = (int)s.open, B = (int)s.target;
    if (!((s.k == 0) & ((double)A == s.open) & ((double)B == s.target) & ((unsigned)A < (unsigned)kTrajDim) &
          ((unsigned)B < (unsigned)kTrajDim)))
        return false;
    const int pair = A * kTrajDim + B;
    int m;
    double o_lo, o_hi;
    if ((pair == pre0) | (pair == pre1)) {                     // the row was copied to shared memory by the prologue
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const double *row = sh_traj + (pair == pre0 ? 0 : kTrajLen);
        o_lo = row[lane]; o_hi = row[lane + 32];
        m = (int)row[kTrajLen - 1];
    } else {
        const double *tr = g_traj + (size_t)pair * kTrajLen;
        m = __ldg(g_traj_len + pair);                          // the three loads are independent of each other
        o_lo = __ldg(tr + lane); o_hi = __ldg(tr + lane + 32);
    }
    if (m == 0 || n + m > ERLFILL_MAX_ITER) return false;
    const double w0 = s.w;
    const bool lf = w0 < p.fast_w, lm = w0 < p.medium_w, ls = w0 < p.slow_w;
    const double stage_target = lf ? p.fast_pos : (lm ? p.medium_pos : p.slow_pos);
    if (!(lf | lm | ls) || dbits(stage_target) != dbits(s.target)) return false;
    if (n + m > 4 * nz.b0 + kWindow) nz.fill(n >> 2, lane);
    const bool v_lo = lane < m, v_hi = lane + 32 < m;
    const double q_lo = __dmul_rn(o_lo, opu_from_u32(nz.word(n + lane)));
    const double p_lo = v_lo ? q_lo : 0.0;
    double p_hi = 0.0;
    if (m > 32) {
        const double q_hi = __dmul_rn(o_hi, opu_from_u32(nz.word(n + 32 + lane)));
        p_hi = v_hi ? q_hi : 0.0;
    }
    const bool small = __any_sync(kFullMask, (v_lo & (p_lo < 0.011)) | (v_hi & (p_hi < 0.011)));
    if (small && s.nochg + m > ERLFILL_NOCHANGE_LIMIT) return false;
    __syncwarp();
    sh_p[lane] = p_lo;
    sh_p[lane + 32] = p_hi;
    __syncwarp();
    const double2 *sh2 = reinterpret_cast<const double2 *>(sh_p);
    double2 v = sh2[0];
    double w = __dadd_rn(w0, v.x);
    if (first) w = __dadd_rn(w, __dmul_rn(s.target, sysdelay));            // one-time stage delay: :192-212
    int nc = 0;
    if (!small) {
        w = __dadd_rn(w, v.y);
#pragma unroll
        for (int j = 1; j < 16; ++j) { v = sh2[j]; w = __dadd_rn(w, v.x); w = __dadd_rn(w, v.y); }
        if (m > 32) {
#pragma unroll
            for (int j = 16; j < 32; ++j) { v = sh2[j]; w = __dadd_rn(w, v.x); w = __dadd_rn(w, v.y); }
        }
    } else {
        nc = (fabs(w - s.prev) < 0.01) ? s.nochg + 1 : 0;
#pragma unroll 1
        for (int j = 1; j < m; ++j) {
            const double w2 = __dadd_rn(w, sh_p[j]);
            nc = (fabs(w2 - w) < 0.01) ? nc + 1 : 0;
            w = w2;
        }
    }
    if (((w < p.fast_w) != lf) | ((w < p.medium_w) != lm) | ((w < p.slow_w) != ls)) return false;
    if (first) delay = lf ? p.fast_delay : (lm ? p.medium_delay : p.slow_delay);
    s.w = w; s.prev = w; s.nochg = nc;
    s.open = s.target; s.k = 0; s.dd = 0.0;
    n += m;
    return true;
}

__device__ __forceinline__ SimResult simulate_warp(const SimParams &p, double sysdelay, WarpNoise &nz, int lane,
                                                  double *sh_p, const double *sh_traj, int pre0, int pre1) {
    Lane s;
    s.w = 0.0; s.open = 0.0; s.prev = 0.0; s.target = p.fast_pos; s.dd = 0.0; s.k = 0; s.nochg = 0;
    int n = 0, delay = 0;
    bool fin = false, first = true;
    nz.fill(0, lane);
    while (!fin) {
        if (in_hold(s, p)) {
            const double thr = s.w < p.fast_w ? p.fast_w : (s.w < p.medium_w ? p.medium_w : p.slow_w);
            const double bin_start = __longlong_as_double(dbits(s.w) & 0x7ff0000000000000LL);
            if (!first & (s.target <= 255.0) & (s.w >= 1.0) & (thr <= 1048576.0) & (2.0 * s.target <= bin_start) &
                hold_caps_ok(nz, n, s.nochg)) {
#ifdef ERLFILL_SIM_PROF
                const long long pf_t = clock64();
#endif
                hold_stretch(s, p, thr, nz, lane, n, fin);
#ifdef ERLFILL_SIM_PROF
                nz.pf_hold += clock64() - pf_t; nz.pf_nhold++;
#endif
                continue;
            }
        } else {
#ifdef ERLFILL_SIM_PROF
            const long long pf_t = clock64();
#endif
            const bool moved = moving_phase(s, p, nz, lane, n, delay, first, sysdelay, sh_p, sh_traj, pre0, pre1);
#ifdef ERLFILL_SIM_PROF
            nz.pf_move += clock64() - pf_t; nz.pf_nmove++; nz.pf_movefail += !moved;
#endif
            if (moved) {
                first = false;
                continue;
            }
        }
#ifdef ERLFILL_SIM_PROF
        const long long pf_g = clock64();
#endif
        // generic tick (first tick without a table entry, openings outside 0..100, stage change while moving, ...)
        if (n >= 4 * nz.b0 + kWindow) nz.fill(n >> 2, lane);
        const double o = opu_from_u32(nz.word(n));
        n += 1;
        if (first) fin = generic_tick<true, kTabGlobal>(s, p, g_motor_tab, o, sysdelay, n, delay);
        else fin = generic_tick<false, kTabGlobal>(s, p, g_motor_tab, o, sysdelay, n, delay);
        first = false;
#ifdef ERLFILL_SIM_PROF
        nz.pf_gen += clock64() - pf_g; nz.pf_ngen++;
#endif
    }
    SimResult res;
    res.final_weight = s.w;
    res.tick = n + delay;
    res.draws = n;
    return res;
}

constexpr int kWarpsPerCta = 1;

// K1: the simulator, one warp per env.  Results travel to K2 in the step's own output slots (no workspace):
// final_weight (binary64) in next_state[i], tick count in reward[i], draws in obs[i].x.
#ifdef ERLFILL_SIM_PROF
__device__ unsigned long long *g_sim_prof = nullptr;    // [n_env][16], development only (tools/sim_prof.cu)
__device__ __forceinline__ unsigned long long pf_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// kFused: lane 0 also runs K2's per-env epilogue right here.  That costs a warp-instruction per scalar instruction
// (about +15 % issue slots per env) but saves the second launch and its cold loads: the host takes it when the
// whole batch is resident at once (one warp slot per env), where issue slots are free and latency is everything.
template <bool kFused>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 32 / kWarpsPerCta) env_sim_warp_kernel(const StepArgs a) {
#ifdef ERLFILL_SIM_PROF
    const unsigned long long pf_g0 = pf_globaltimer();
    const long long pf_c0 = clock64();
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kWarpsPerCta + warp;
    if (i >= a.env.n_env) return;                               // warp-uniform
    const uint32_t k0 = (uint32_t)a.env.seed, k1 = (uint32_t)(a.env.seed >> 32);
    const uint32_t gid = a.env.env_id_base + (uint32_t)i;
    const int cid = a.env.d_cond_id ? a.env.d_cond_id[i] : (int)(gid % (uint32_t)a.env.n_cond);
    const erlfill_condition *c = a.env.d_cond + cid;
    float v = 0.f;
    if (lane < ERLFILL_ACTION_DIM) {
        v = __ldg(a.d_actions + (size_t)i * ERLFILL_ACTION_DIM + lane);
        const float lo = c->lo[lane], hi = c->hi[lane];
        v = v < lo ? lo : v;
        v = v > hi ? hi : v;
    }
    const int iv = (int)v;                                      // int() truncation (VirtualWeightController.py:71-83)
    SimParams p;
    p.fast_w = (double)__shfl_sync(kFullMask, iv, 0);
    p.medium_w = (double)__shfl_sync(kFullMask, iv, 1);
    p.slow_w = (double)__shfl_sync(kFullMask, iv, 2);
    p.fast_pos = (double)__shfl_sync(kFullMask, iv, 3);
    p.medium_pos = (double)__shfl_sync(kFullMask, iv, 4);
    p.slow_pos = (double)__shfl_sync(kFullMask, iv, 5);
    p.fast_delay = __shfl_sync(kFullMask, iv, 6);
    p.medium_delay = __shfl_sync(kFullMask, iv, 7);
    p.slow_delay = __shfl_sync(kFullMask, iv, 8);
    p.sane = (p.fast_pos >= 0) & (p.medium_pos >= 0) & (p.slow_pos >= 0);
    p.fast_mode = a.warp_fast_mode;
    // the two motor trajectories this env will almost surely need (0 -> fast opening -> slow opening): copy their
    // table rows into shared memory (cp.async, L1-allocating: the envs of an SM mostly share rows) while the first
    // noise window is generated
    __shared__ __align__(16) double sh_p[kWarpsPerCta][64];
    __shared__ __align__(16) double sh_traj[kWarpsPerCta][2 * kTrajLen];
    __shared__ __align__(16) uint32_t sh_noise[kWarpsPerCta][kWindow];
    int pre0 = -1, pre1 = -1;
    {
        const int fo = __shfl_sync(kFullMask, iv, 3), so = __shfl_sync(kFullMask, iv, 5);
        if (((unsigned)fo < (unsigned)kTrajDim) & ((unsigned)so < (unsigned)kTrajDim)) {
            pre0 = fo; pre1 = fo * kTrajDim + so;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sh_traj[warp][lane]);
            const double *s0 = g_traj + (size_t)pre0 * kTrajLen + lane, *s1 = g_traj + (size_t)pre1 * kTrajLen + lane;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(s0) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 256u), "l"(s0 + 32) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 512u), "l"(s1) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 768u), "l"(s1 + 32) : "memory");
        }
    }
    const uint4 misc = philox4x32_10(k0, k1, gid, a.step(), 0u, STREAM_MISC);
    const double sysdelay = 15.0 + 35.0 * u32_to_unit(misc.x);             // uniform(15, 50): :137
    WarpNoise nz;
    nz.k0 = k0; nz.k1 = k1; nz.gid = gid; nz.step_idx = a.step(); nz.b0 = 0; nz.sh = sh_noise[warp];
#ifdef ERLFILL_SIM_PROF
    const long long pf_c1 = clock64();
#endif
    const SimResult sim = simulate_warp(p, sysdelay, nz, lane, sh_p[warp], sh_traj[warp], pre0, pre1);
    if (kFused) {
        float clipped[ERLFILL_ACTION_DIM];
#pragma unroll
        for (int j = 0; j < ERLFILL_ACTION_DIM; ++j) clipped[j] = __shfl_sync(kFullMask, v, j);
        if (lane == 0) {
            const erlfill_condition cc = *c;
            finish_env<false>(a, i, cc, clipped, sim, (int)clipped[9], misc);
            if (a.out.d_tick_sum && sim.draws) atomicAdd(a.out.d_tick_sum, (unsigned long long)sim.draws);
            if (a.d_step_idx) {
                // graph-replay mode: the last env to finish advances the lock-step index (every warp read it in its
                // prologue, so nobody still needs the old value) and re-arms the ticket word next to it
                uint32_t *cnt = const_cast<uint32_t *>(a.d_step_idx);
                if (atomicAdd(cnt + 1, 1u) == (uint32_t)a.env.n_env - 1u) {
                    cnt[1] = 0u;
                    cnt[0] += 1u;
                }
            }
        }
    } else if (lane == 0) {
        reinterpret_cast<double *>(a.out.d_next_state)[i] = sim.final_weight;
        reinterpret_cast<int *>(a.out.d_reward)[i] = sim.tick;
        reinterpret_cast<int *>(a.env.d_obs)[2 * i] = sim.draws;
    }
#ifdef ERLFILL_SIM_PROF
    if (lane == 0 && g_sim_prof) {
        unsigned long long *o = g_sim_prof + (size_t)i * 16;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        o[0] = pf_g0; o[1] = pf_globaltimer(); o[2] = smid;
        o[3] = (unsigned long long)(pf_c1 - pf_c0);              // prologue
        o[4] = (unsigned long long)(clock64() - pf_c1);          // simulate_warp + stores
        o[5] = nz.pf_hold; o[6] = nz.pf_cross; o[7] = nz.pf_move; o[8] = nz.pf_gen;
        o[9] = nz.pf_win; o[10] = nz.pf_nhold; o[11] = nz.pf_nmove; o[12] = nz.pf_ngen; o[13] = nz.pf_movefail;
        o[14] = ((unsigned long long)nz.pf_fast << 24) | (unsigned)sim.draws; o[15] = nz.pf_fastwin;
    }
#endif
}

// K2: reward / state / done / auto-reset / stores, one thread per env (E2/E3/E4)
__global__ void __launch_bounds__(128) env_finish_kernel(const StepArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int draws = 0;
    if (i < a.env.n_env) {
        const uint32_t gid = a.env.env_id_base + (uint32_t)i;
        const int cid = a.env.d_cond_id ? a.env.d_cond_id[i] : (int)(gid % (uint32_t)a.env.n_cond);
        const erlfill_condition c = a.env.d_cond[cid];
        float clipped[ERLFILL_ACTION_DIM];
#pragma unroll
        for (int j = 0; j < ERLFILL_ACTION_DIM; ++j) {
            float v = __ldg(a.d_actions + (size_t)i * ERLFILL_ACTION_DIM + j);
            v = v < c.lo[j] ? c.lo[j] : v;
            v = v > c.hi[j] ? c.hi[j] : v;
            clipped[j] = v;
        }
        SimResult sim;
        sim.final_weight = reinterpret_cast<const double *>(a.out.d_next_state)[i];
        sim.tick = reinterpret_cast<const int *>(a.out.d_reward)[i];
        sim.draws = reinterpret_cast<const int *>(a.env.d_obs)[2 * i];
        draws = sim.draws;
        const uint4 misc = philox4x32_10((uint32_t)a.env.seed, (uint32_t)(a.env.seed >> 32), gid, a.step(), 0u, STREAM_MISC);
        finish_env<false>(a, i, c, clipped, sim, (int)clipped[9], misc);
    }
    if (a.out.d_tick_sum) {
        const int sum = __reduce_add_sync(kFullMask, draws);
        if ((threadIdx.x & 31) == 0 && sum) atomicAdd(a.out.d_tick_sum, (unsigned long long)sum);
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

__global__ void env_reset_replay_kernel(const erlfill_env_batch env, const uint8_t *d_mask, const float *d_replay,
                                        int row_stride, int64_t n_valid, const int64_t *d_idx, uint32_t step_idx,
                                        int update_counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= env.n_env) return;
    if (d_mask && !d_mask[i]) return;
    const uint32_t gid = env.env_id_base + (uint32_t)i;
    const int cid = env.d_cond_id ? env.d_cond_id[i] : (int)(gid % (uint32_t)env.n_cond);
    const erlfill_condition c = env.d_cond[cid];
    // np.mean(list of 10 f32[2], axis=0): row-sequential binary32 adds, then /10 (MutiConditionEnv.py:147-148)
    float2 acc = make_float2(0.f, 0.f);
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int j = 0; j < 10; ++j) {
        int64_t row;
        if (d_idx) row = d_idx[(size_t)i * 10 + j];
        else {
            // batched loop: 10 draws with replacement, floor(U * n_valid), Philox (env, step, 1 + j/4, MISC)
            if ((j & 3) == 0) r = philox4x32_10((uint32_t)env.seed, (uint32_t)(env.seed >> 32), gid, step_idx, 1u + (uint32_t)(j >> 2), STREAM_MISC);
            const uint32_t x = (j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w));
            row = (int64_t)(((unsigned long long)x * (unsigned long long)n_valid) >> 32);
        }
        const float2 v = *reinterpret_cast<const float2 *>(d_replay + (size_t)row * row_stride);
        if (j == 0) acc = v; else { acc.x += v.x; acc.y += v.y; }
    }
    const float raw0 = acc.x / 10.f, raw1 = acc.y / 10.f;
    reinterpret_cast<float2 *>(env.d_obs)[i] =
        make_float2(raw0 / (float)c.target_weight_err, raw1 / (float)c.target_time);   // normalised again: :151-154
    if (update_counters) {
        env.d_step_counter[i] = 0;
        env.d_episode_counter[i] += 1;
        env.d_ring_len[i] = 0;
        env.d_ring_head[i] = 0;
    }
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

// envs (= warps) the device holds at once with one warp per CTA: 32 CTAs per SM
static int resident_warp_slots() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * (32 / kWarpsPerCta) * kWarpsPerCta;
}

static int g_env_step_mode = ERLFILL_SIM_AUTO;
static int g_warp_fast_mode = 0;

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

int erlfill_env_step_set_mode(int mode) {
    if (mode < ERLFILL_SIM_AUTO || mode > ERLFILL_SIM_WARP_REJECT_RUNS) return ERLFILL_ERR_ARG;
    g_warp_fast_mode = mode == ERLFILL_SIM_WARP_WINDOWS ? 1 : (mode == ERLFILL_SIM_WARP_REJECT_RUNS ? 2 : 0);
    g_env_step_mode = mode >= ERLFILL_SIM_WARP_WINDOWS ? ERLFILL_SIM_WARP_PER_ENV : mode;
    return ERLFILL_OK;
}

int erlfill_env_step_fused_max_envs(void) { return resident_warp_slots(); }

int erlfill_env_reset(const erlfill_env_batch *env, const uint8_t *d_mask, uint32_t step_idx,
                      const erlfill_noise *noise, void *stream) {
    if (!env_ok(env)) return ERLFILL_ERR_ARG;
    const int block = 128, grid = (env->n_env + block - 1) / block;
    env_reset_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*env, d_mask, step_idx,
                                                                noise ? noise->d_reset_u : nullptr);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_env_reset_from_replay(const erlfill_env_batch *env, const uint8_t *d_mask, const float *d_replay,
                                  int row_stride, int64_t n_valid, const int64_t *d_idx, uint32_t step_idx,
                                  int update_counters, void *stream) {
    if (!env_ok(env) || !d_replay || row_stride < 2 || (row_stride & 1)) return ERLFILL_ERR_ARG;
    if (!d_idx && (n_valid <= 0 || n_valid > 0xFFFFFFFFLL)) return ERLFILL_ERR_ARG;
    const int block = 128, grid = (env->n_env + block - 1) / block;
    env_reset_replay_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*env, d_mask, d_replay, row_stride, n_valid, d_idx,
                                                                       step_idx, update_counters);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

__global__ void step_idx_increment_kernel(uint32_t *p) { *p += 1u; }

static int env_step_impl(const erlfill_env_batch *env, const float *d_actions, uint32_t step_idx, uint32_t *d_step_idx, int auto_reset,
                         const erlfill_noise *noise, const erlfill_step_out *out, void *stream) {
    if (!env_ok(env) || !d_actions || !out || !out->d_next_state || !out->d_reward || !out->d_done)
        return ERLFILL_ERR_ARG;
    if (env->kind < 0 || env->kind > (ERLFILL_ENV_SINGLE | ERLFILL_ENV_ONBOARD)) return ERLFILL_ERR_ARG;
    const bool onboard = (env->kind & ERLFILL_ENV_ONBOARD) != 0;
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
    a.d_step_idx = d_step_idx;
    a.auto_reset = auto_reset;
    a.actions_aligned16 = ((uintptr_t)d_actions & 15u) == 0;
    a.warp_fast_mode = g_warp_fast_mode;
    const int block = pick_block(env->n_env);
    const int grid = (env->n_env + block - 1) / block;
    if (onboard) {
        // row N1: int()-truncated weight per tick -> thread-per-env tick loop (Philox or injected noise)
        if (a.d_u) env_step_kernel<true, true><<<grid, block, 0, (cudaStream_t)stream>>>(a);
        else env_step_kernel<false, true><<<grid, block, 0, (cudaStream_t)stream>>>(a);
    } else if (a.d_u) {
        // injected binary64 noise is not on the 2^-31 grid: thread-per-env kernel, binary64 throughout
        env_step_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(a);
    } else if (g_env_step_mode == ERLFILL_SIM_THREAD_PER_ENV) {
        env_step_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(a);
    } else {
        const int wgrid = (env->n_env + kWarpsPerCta - 1) / kWarpsPerCta;
        if (env->n_env <= resident_warp_slots()) {
            env_sim_warp_kernel<true><<<wgrid, kWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(a);
        } else {
            env_sim_warp_kernel<false><<<wgrid, kWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(a);
            ERLFILL_CUDA_TRY(cudaGetLastError());
            env_finish_kernel<<<(env->n_env + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
        }
    }
    const bool fused_warp = !onboard && !a.d_u && g_env_step_mode != ERLFILL_SIM_THREAD_PER_ENV && env->n_env <= resident_warp_slots();
    if (d_step_idx && !fused_warp) step_idx_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_step_idx);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_env_step(const erlfill_env_batch *env, const float *d_actions, uint32_t step_idx, int auto_reset,
                     const erlfill_noise *noise, const erlfill_step_out *out, void *stream) {
    return env_step_impl(env, d_actions, step_idx, nullptr, auto_reset, noise, out, stream);
}

int erlfill_env_step_counted(const erlfill_env_batch *env, const float *d_actions, uint32_t *d_step_idx, int auto_reset,
                             const erlfill_noise *noise, const erlfill_step_out *out, void *stream) {
    if (!d_step_idx) return ERLFILL_ERR_ARG;
    return env_step_impl(env, d_actions, 0u, d_step_idx, auto_reset, noise, out, stream);
}

}  // extern "C"
