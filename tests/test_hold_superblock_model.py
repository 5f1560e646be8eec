"""CPU model of the warp-per-env simulator's integer hold phase (csrc/env_step.cu: hold_superblock) checked
against the C oracle's serial binary64 loop on the same Philox noise.

The CUDA kernel replaces the reference's per-tick `w += target * (1 + u)` (VirtualWeightController.py:122-125)
by exact integer prefix sums over 128-tick super-blocks.  This test restates that algorithm with Python
integers/floats (no GPU) and shows it reproduces the oracle's tick count and final weight bit for bit,
including the cases the exactness argument singles out: binade crossings of w, stage thresholds on a power of
two, tiny openings (stagnation counting), opening 0 (stagnation stop) and the iteration cap.
"""
import math
import random
import struct

from oracle import oracle_lib as O

MAX_ITER, NOCHG_LIMIT, SUPER = 5000, 500, 128
SMALL_MAX = 21474836          # P * 2^-31 < 0.01  <=>  P <= 21474836
SCALE = 1.0 / 2147483648.0


def _bits(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


def _from_bits(b):
    return struct.unpack("<d", struct.pack("<q", b))[0]


class Noise:
    def __init__(self, seed, env_id, step_idx):
        self.seed, self.env_id, self.step_idx, self.cache = seed, env_id, step_idx, {}

    def block(self, b):
        if b not in self.cache:
            self.cache[b] = O.philox(self.seed, self.env_id, self.step_idx, b, 0)
        return self.cache[b]

    def word(self, n):
        return self.block(n >> 2)[n & 3]


def _generic_tick(s, p, x, sysdelay, n_after, first):
    """One tick of the reference loop (same statement order as oracle/py_port.py); s is a dict."""
    opening, target, speed = s["open"], s["target"], s["speed"]
    dt = speed / 100000 if speed > 0 else 0
    dd = 0.5 * 100000 * dt ** 2
    if abs(opening - target) > 1:
        if opening < target:
            speed += 5000.0 if opening < target - dd else -100.0
            direction = 1
        else:
            speed += 5000.0 if opening > dd else -100.0
            direction = -1
        opening += speed * direction / 1000
    else:
        speed = 0
        opening = target
    s["open"], s["speed"] = opening, speed
    w = s["w"] + opening * (x * 2.0 ** -31)
    if w < 0:
        w = 0.0
    lf, lm, ls = w < p["fast_w"], w < p["medium_w"], w < p["slow_w"]
    if lf or lm or ls:
        s["target"] = p["fast_pos"] if lf else (p["medium_pos"] if lm else p["slow_pos"])
    if first and (lf or lm or ls):
        s["delay"] = p["fast_delay"] if lf else (p["medium_delay"] if lm else p["slow_delay"])
        w = w + s["target"] * sysdelay
    s["w"] = w
    if not (lf or lm or ls):
        return True
    s["nochg"] = s["nochg"] + 1 if abs(w - s["prev"]) < 0.01 else 0
    s["prev"] = w
    return n_after > MAX_ITER or s["nochg"] > NOCHG_LIMIT


def _hold_ok(s, p, n):
    sane = p["fast_pos"] >= 0 and p["medium_pos"] >= 0 and p["slow_pos"] >= 0
    return (sane and s["speed"] == 0 and s["open"] == s["target"] and s["target"] <= 255.0 and 1.0 <= s["w"] < 2097152.0
            and (n & ~3) + SUPER <= MAX_ITER and s["nochg"] + SUPER <= NOCHG_LIMIT)


def _hold_superblock(s, p, nz, n, stats):
    w0 = s["w"]
    thr = p["fast_w"] if w0 < p["fast_w"] else (p["medium_w"] if w0 < p["medium_w"] else p["slow_w"])
    bin_end = _from_bits((_bits(w0) & 0x7FF0000000000000) + 0x0010000000000000)
    lim = min(thr, bin_end)
    d = (lim - w0) * 2147483648.0
    Dc = int(math.ceil(d))
    b0, skip = n >> 2, n & 3
    ti = int(s["target"])
    P = []
    for t in range(SUPER):
        P.append(0 if t < skip else ti * nz.word(4 * b0 + t))
    total = sum(P)
    if total < Dc:
        stats["fast"] += 1
        s["w"] = w0 + float(total) * SCALE
        last_ns = max([t for t in range(SUPER) if P[t] > SMALL_MAX], default=-1)
        if any(t >= skip and P[t] <= SMALL_MAX for t in range(SUPER)):
            s["nochg"] = s["nochg"] + (SUPER - skip) if last_ns < 0 else SUPER - 1 - last_ns
        else:
            s["nochg"] = 0
        s["prev"] = s["w"]
        return 4 * b0 + SUPER, False
    stats["cross"] += 1
    acc = 0
    for j in range(SUPER):
        if acc + P[j] >= Dc:
            break
        acc += P[j]
    w_prev = w0 + float(acc) * SCALE
    w_new = w_prev + float(P[j]) * SCALE
    last_ns = max([t for t in range(j) if P[t] > SMALL_MAX], default=-1)
    run_before = s["nochg"] + (j - skip) if last_ns < 0 else j - 1 - last_ns
    n = 4 * b0 + j + 1
    s["w"] = w_new
    lf, lm, ls = w_new < p["fast_w"], w_new < p["medium_w"], w_new < p["slow_w"]
    if not (lf or lm or ls):
        return n, True
    s["target"] = p["fast_pos"] if lf else (p["medium_pos"] if lm else p["slow_pos"])
    s["nochg"] = run_before + 1 if abs(w_new - w_prev) < 0.01 else 0
    s["prev"] = w_new
    return n, n > MAX_ITER or s["nochg"] > NOCHG_LIMIT


def simulate_model(params, seed, env_id, step_idx, stats):
    """params in load_params order: target, fast_w, medium_w, slow_w, fast/medium/slow delay, fast/medium/slow opening, unload."""
    p = dict(fast_w=float(params[1]), medium_w=float(params[2]), slow_w=float(params[3]), fast_delay=params[4],
             medium_delay=params[5], slow_delay=params[6], fast_pos=float(params[7]), medium_pos=float(params[8]),
             slow_pos=float(params[9]))
    misc = O.philox(seed, env_id, step_idx, 0, 1)
    sysdelay = 15.0 + 35.0 * (misc[0] * (1.0 / 4294967296.0))
    nz = Noise(seed, env_id, step_idx)
    s = dict(w=0.0, open=0.0, speed=0, prev=0.0, target=p["fast_pos"], nochg=0, delay=0)
    fin = _generic_tick(s, p, nz.word(0), sysdelay, 1, True)
    n = 1
    while not fin:
        if _hold_ok(s, p, n):
            n, fin = _hold_superblock(s, p, nz, n, stats)
        else:
            stats["generic"] += 1
            x = nz.word(n)
            n += 1
            fin = _generic_tick(s, p, x, sysdelay, n, False)
    return n, (n + s["delay"] + params[10]) / 1000.0, s["w"]


def _check(params, seed, env_id, step_idx, stats):
    n, tt, fw = simulate_model(params, seed, env_id, step_idx, stats)
    rn, rtt, rfw, _ = O.simulate_run_philox(params, seed, env_id, step_idx)
    assert (n, tt) == (rn, rtt), (params, n, rn)
    assert _bits(fw) == _bits(rfw), (params, fw, rfw)
    return n


def test_hold_superblock_model_matches_oracle():
    rng = random.Random(5)
    stats = dict(fast=0, cross=0, generic=0)
    seed = 0x1234ABCD5678
    cases = []
    # shipped bounds (experiment_runner.py:417-478, WeightEnv.py:32-46)
    for _ in range(60):
        tw = rng.choice([25000, 20000, 15000])
        cases.append([tw, rng.randint(5000, 23000), rng.randint(0, 1), rng.randint(tw - 200, tw + 25), rng.randint(60, 300),
                      rng.randint(60, 200), rng.randint(60, 200), rng.randint(25, 75), rng.randint(2, 5), rng.randint(3, 20),
                      rng.randint(250, 500)])
    # thresholds on powers of two, reachable medium stage, tiny and zero openings, equal openings
    cases += [[25000, 16384, 0, 24900, 100, 100, 100, 40, 3, 3, 300],
              [25000, 8192, 16384, 24900, 100, 100, 100, 40, 17, 3, 300],
              [25000, 16383, 16385, 16390, 100, 100, 100, 60, 30, 30, 300],
              [25000, 12000, 0, 24950, 200, 150, 150, 55, 4, 1, 400],
              [25000, 12000, 0, 24950, 200, 150, 150, 55, 4, 0, 400],      # opening 0: stagnation stop
              [25000, 12000, 0, 24950, 200, 150, 150, 55, 4, 2, 400],
              [3000, 1000, 2000, 3000, 10, 10, 10, 1, 1, 1, 100],          # iteration cap
              [25000, 20000, 0, 24990, 100, 100, 100, 200, 3, 255, 300],
              [25000, 20000, 0, 24990, 100, 100, 100, 300, 3, 5, 300],     # opening above the fast-path limit
              [600000, 300000, 0, 600000, 100, 100, 100, 255, 3, 250, 300],
              [25000, 0, 0, 24000, 100, 100, 100, 50, 3, 50, 300],
              [100, 50, 0, 90, 1, 1, 1, 2, 2, 2, 1]]
    ticks = 0
    for i, c in enumerate(cases):
        ticks += _check(c, seed, 1000 + i, i % 7, stats)
    assert stats["fast"] > 300 and stats["cross"] > 100, stats
    assert stats["generic"] < 0.2 * ticks
