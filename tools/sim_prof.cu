// tools/sim_prof.cu -- development probe (not product): phase profile of env_sim_warp_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --std=c++17 -fmad=false -DERLFILL_SIM_PROF \
//        -I include -I erl-fill_b200/csrc tools/sim_prof.cu -o tools/bin/sim_prof
//   tools/bin/sim_prof [n_env] [steps] [random]
// Per env: start/end globaltimer, SM id, cycles in prologue / hold windows / crossing / moving / generic ticks.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../erl-fill_b200/csrc/env_step.cu"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int probe_main();
int main(int argc, char **argv) {
    if (argc > 1 && argv[1][0] == 'p') return probe_main();
    const int n = argc > 1 ? atoi(argv[1]) : 4096;
    const int steps = argc > 2 ? atoi(argv[2]) : 6;
    const bool rnd = argc > 3 && argv[3][0] == 'r';
    const bool noflush = argc > 4;
    erlfill_condition c{};
    c.target_weight = 25000; c.target_weight_err = 25; c.target_time = 2.5;
    const float lo[10] = {18000, 0, 24900, 35, 3, 3, 100, 100, 100, 300};
    const float hi[10] = {23000, 1, 25025, 60, 5, 7, 300, 200, 200, 500};
    for (int j = 0; j < 10; ++j) { c.lo[j] = lo[j]; c.hi[j] = hi[j]; }
    std::vector<float> act((size_t)n * 10);
    srand(1);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 10; ++j) act[(size_t)i * 10 + j] = rnd ? lo[j] + (hi[j] - lo[j]) * (rand() / (float)RAND_MAX) : lo[j];
    erlfill_env_batch e{};
    e.n_env = n; e.kind = ERLFILL_ENV_SINGLE; e.max_steps = 100; e.n_cond = 1; e.env_id_base = 0; e.seed = 0;
    erlfill_condition *d_c; CK(cudaMalloc(&d_c, sizeof(c))); CK(cudaMemcpy(d_c, &c, sizeof(c), cudaMemcpyHostToDevice));
    e.d_cond = d_c; e.d_cond_id = nullptr;
    CK(cudaMalloc(&e.d_step_counter, n * 4)); CK(cudaMalloc(&e.d_episode_counter, n * 4));
    CK(cudaMalloc(&e.d_ring_len, n * 4)); CK(cudaMalloc(&e.d_ring_head, n * 4));
    CK(cudaMalloc(&e.d_ring, (size_t)n * 20 * 8)); CK(cudaMalloc(&e.d_obs, (size_t)n * 8));
    CK(cudaMemset(e.d_step_counter, 0, n * 4)); CK(cudaMemset(e.d_episode_counter, 0, n * 4));
    CK(cudaMemset(e.d_ring_len, 0, n * 4)); CK(cudaMemset(e.d_ring_head, 0, n * 4));
    CK(cudaMemset(e.d_ring, 0, (size_t)n * 20 * 8)); CK(cudaMemset(e.d_obs, 0, (size_t)n * 8));
    erlfill_step_out o{};
    CK(cudaMalloc(&o.d_next_state, (size_t)n * 8)); CK(cudaMalloc(&o.d_reward, n * 4)); CK(cudaMalloc(&o.d_done, n));
    CK(cudaMalloc(&o.d_tick_sum, 8)); CK(cudaMemset(o.d_tick_sum, 0, 8));
    float *d_act; CK(cudaMalloc(&d_act, act.size() * 4)); CK(cudaMemcpy(d_act, act.data(), act.size() * 4, cudaMemcpyHostToDevice));
    unsigned long long *d_prof; CK(cudaMalloc(&d_prof, (size_t)n * 16 * 8)); CK(cudaMemset(d_prof, 0, (size_t)n * 16 * 8));
    CK(cudaMemcpyToSymbol(erlfill::g_sim_prof, &d_prof, sizeof(d_prof)));
    void *d_flush; const size_t fb = 256u << 20; CK(cudaMalloc(&d_flush, fb));
    if (erlfill_env_reset(&e, nullptr, 0, nullptr, nullptr) != 0) { printf("reset failed\n"); return 1; }
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float ms_last = 0, ms_sim = 0;
    for (int s = 0; s < steps; ++s) {
        if (!noflush) CK(cudaMemsetAsync(d_flush, 0, fb));
        CK(cudaEventRecord(e0));
        if (erlfill_env_step(&e, d_act, (uint32_t)(s + 1), 1, nullptr, &o, nullptr) != 0) { printf("step failed\n"); return 1; }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms_last, e0, e1));
        printf("step %d: %.2f us\n", s, ms_last * 1e3f);
    }
    (void)ms_sim; (void)e2;
    std::vector<unsigned long long> pr((size_t)n * 16);
    CK(cudaMemcpy(pr.data(), d_prof, pr.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long g0 = ~0ull, g1 = 0;
    for (int i = 0; i < n; ++i) { g0 = std::min(g0, pr[i * 16]); g1 = std::max(g1, pr[i * 16 + 1]); }
    auto stat = [&](const char *name, auto f) {
        std::vector<double> v(n);
        for (int i = 0; i < n; ++i) v[i] = f(&pr[(size_t)i * 16]);
        std::sort(v.begin(), v.end());
        double sum = 0; for (double x : v) sum += x;
        printf("%-28s mean %10.1f  min %10.1f  p50 %10.1f  p90 %10.1f  max %10.1f\n", name, sum / n, v[0], v[n / 2], v[n * 9 / 10], v[n - 1]);
    };
    printf("n_env %d, kernel span (globaltimer) %.2f us\n", n, (g1 - g0) * 1e-3);
    stat("start offset us", [&](const unsigned long long *p) { return (p[0] - g0) * 1e-3; });
    stat("end offset us", [&](const unsigned long long *p) { return (p[1] - g0) * 1e-3; });
    stat("env duration us", [&](const unsigned long long *p) { return (p[1] - p[0]) * 1e-3; });
    stat("cyc prologue", [&](const unsigned long long *p) { return (double)p[3]; });
    stat("cyc simulate_warp", [&](const unsigned long long *p) { return (double)p[4]; });
    stat("cyc hold total", [&](const unsigned long long *p) { return (double)p[5]; });
    stat("cyc  of which crossing", [&](const unsigned long long *p) { return (double)p[6]; });
    stat("cyc moving", [&](const unsigned long long *p) { return (double)p[7]; });
    stat("cyc generic", [&](const unsigned long long *p) { return (double)p[8]; });
    stat("n full windows", [&](const unsigned long long *p) { return (double)p[9]; });
    stat("n hold stretches", [&](const unsigned long long *p) { return (double)p[10]; });
    stat("n moving calls", [&](const unsigned long long *p) { return (double)p[11]; });
    stat("n generic ticks", [&](const unsigned long long *p) { return (double)p[12]; });
    stat("n moving failures", [&](const unsigned long long *p) { return (double)p[13]; });
    stat("draws", [&](const unsigned long long *p) { return (double)(p[14] & 0xFFFFFF); });
    stat("cyc fast runs", [&](const unsigned long long *p) { return (double)(p[14] >> 24); });
    stat("n fast windows", [&](const unsigned long long *p) { return (double)p[15]; });
    // warps per SM
    int cnt[256] = {};
    for (int i = 0; i < n; ++i) cnt[pr[(size_t)i * 16 + 2] & 255]++;
    int mn = 1 << 30, mx = 0, used = 0;
    for (int s = 0; s < 256; ++s) if (cnt[s]) { mn = std::min(mn, cnt[s]); mx = std::max(mx, cnt[s]); ++used; }
    printf("SMs used %d, envs per SM min %d max %d\n", used, mn, mx);
    return 0;
}

// ---- fast-run probe: the multi-window loop of hold_stretch in isolation -------------------------------------
namespace erlfill {
template <int V>
__global__ void fast_probe(unsigned long long *out, long long *cyc, uint32_t ti, uint32_t xs, int K, uint32_t k0, uint32_t k1) {
    const int lane = threadIdx.x & 31;
    const uint32_t blk0 = 5u + (uint32_t)lane, gid = blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned long long lsum = 0;
    uint32_t worst = 0xffffffffu;
    uint4 rl[2];
    const long long t0 = clock64();
    if (V == 0 || V == 1 || V == 2) {
#pragma unroll 2
        for (int wdx = 0; wdx < K; ++wdx) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                rl[h] = philox4x32_10(k0, k1, gid, 3u, blk0 + (uint32_t)(32 * (2 * wdx + h)), STREAM_TICK);
                if (V == 2) lsum += (unsigned long long)rl[h].x + rl[h].y + rl[h].z + rl[h].w;
                else lsum += block_sum<false>(ti, rl[h], 0, 0);
                if (V != 1) worst = min(worst, max(max(rl[h].x, rl[h].y), max(rl[h].z, rl[h].w)));
            }
        }
    } else {
#pragma unroll 1
        for (int wdx = 0; wdx < K; wdx += 2) {
            uint4 r[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) r[h] = philox4x32_10(k0, k1, gid, 3u, blk0 + (uint32_t)(32 * (2 * wdx + h)), STREAM_TICK);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                lsum += (unsigned long long)r[h].x + r[h].y + r[h].z + r[h].w;
                worst = min(worst, max(max(r[h].x, r[h].y), max(r[h].z, r[h].w)));
            }
            rl[0] = r[2]; rl[1] = r[3];
        }
    }
    if (V == 2 || V == 3) lsum *= ti;
    const unsigned long long total = warp_sum_u59(lsum);
    const bool bad = __any_sync(kFullMask, worst <= xs);
    const long long t1 = clock64();
    if (lane == 0) { out[gid] = total + bad + rl[0].x + rl[1].y; if (gid == 0) cyc[V] = t1 - t0; }
}
}  // namespace erlfill

int probe_main() {
    unsigned long long *o; long long *c; cudaMalloc(&o, 148 * 64 * 8); cudaMallocManaged(&c, 64);
    for (int K = 8; K <= 72; K += 64)
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            erlfill::fast_probe<0><<<148, warps * 32>>>(o, c, 3u, 7000000u, K, 1u, 2u);
            erlfill::fast_probe<1><<<148, warps * 32>>>(o, c, 3u, 7000000u, K, 1u, 2u);
            erlfill::fast_probe<2><<<148, warps * 32>>>(o, c, 3u, 7000000u, K, 1u, 2u);
            erlfill::fast_probe<3><<<148, warps * 32>>>(o, c, 3u, 7000000u, K, 1u, 2u);
            cudaDeviceSynchronize();
            printf("warps/SM %2d K=%d rep %d: fast run cycles: as-is %lld, no worst %lld, add-only sums %lld, 4 chains add-only %lld\n", warps, K, rep, c[0], c[1], c[2], c[3]);
        }
    }
    return 0;
}
