"""CPU model of the warp-per-env simulator (csrc/env_step.cu: hold_superblock, moving_phase) checked
against the C oracle's serial binary64 loop on the same Philox noise.

The CUDA kernel replaces the reference's per-tick `w += target * (1 + u)` (VirtualWeightController.py:122-125)
by exact integer prefix sums over 256-draw noise windows (with the round-to-nearest-even corrections of
every binade w passes folded into one precomputed anchor), and the motor model (:85-113) by noise-independent
trajectory tables between hold states.  This test restates both with Python
integers/floats (no GPU) and shows it reproduces the oracle's tick count and final weight bit for bit,
including the cases the exactness argument singles out: binade crossings of w, stage thresholds on a power of
two, tiny openings (stagnation counting), opening 0 (stagnation stop) and the iteration cap.
"""
import math
import random
import struct

from oracle import oracle_lib as O

MAX_ITER, NOCHG_LIMIT = 5000, 500
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


_TRAJ = {}


def _trajectory(A, B):
    """Motor trajectory from the hold state (open=A, speed 0) to target B (csrc/env_step.cu: g_traj), or None."""
    if (A, B) not in _TRAJ:
        st = dict(open=float(A), target=float(B), speed=0)
        out, ok = [], A != B
        while ok and len(out) < 64:
            opening, target, speed = st["open"], st["target"], st["speed"]
            dt = speed / 100000 if speed > 0 else 0
            dd = 0.5 * 100000 * dt ** 2
            far = abs(opening - target) > 1
            if far:
                if opening < target:
                    speed += 5000.0 if opening < target - dd else -100.0
                    opening += speed / 1000
                else:
                    speed += 5000.0 if opening > dd else -100.0
                    opening += speed * -1 / 1000
            else:
                speed, opening = 0, target
            st["open"], st["speed"] = opening, speed
            if opening < 0 or speed < 0:
                ok = False
            out.append(opening)
            if not far:
                break
        else:
            ok = False
        _TRAJ[(A, B)] = out if ok else None
    return _TRAJ[(A, B)]


def _moving_phase(s, p, nz, n, first, sysdelay, stats):
    A, B = int(s["open"]), int(s["target"])
    if not (s["speed"] == 0 and A == s["open"] and B == s["target"] and 0 <= A <= 100 and 0 <= B <= 100):
        return None
    tr = _trajectory(A, B)
    if tr is None or n + len(tr) > MAX_ITER:
        return None
    w0 = s["w"]
    lf, lm, ls = w0 < p["fast_w"], w0 < p["medium_w"], w0 < p["slow_w"]
    if not (lf or lm or ls):
        return None
    if (p["fast_pos"] if lf else (p["medium_pos"] if lm else p["slow_pos"])) != s["target"]:
        return None
    inc = [tr[j] * (nz.word(n + j) * 2.0 ** -31) for j in range(len(tr))]
    small = min(inc) < 0.011        # some |dw| may be below 0.01: the stagnation count is then kept tick by tick
    if small and s["nochg"] + len(tr) > NOCHG_LIMIT:
        return None
    w = w0 + inc[0]
    if first:
        w = w + s["target"] * sysdelay
    nc = 0
    if small:
        nc = s["nochg"] + 1 if abs(w - s["prev"]) < 0.01 else 0
        stats["table_small"] = stats.get("table_small", 0) + 1
    for j in range(1, len(tr)):
        w2 = w + inc[j]
        if small:
            nc = nc + 1 if abs(w2 - w) < 0.01 else 0
        w = w2
    if (w < p["fast_w"]) != lf or (w < p["medium_w"]) != lm or (w < p["slow_w"]) != ls:
        return None
    if first:
        s["delay"] = p["fast_delay"] if lf else (p["medium_delay"] if lm else p["slow_delay"])
    s.update(w=w, prev=w, nochg=nc, open=s["target"], speed=0)
    stats["table"] += 1
    return n + len(tr)


WINDOW = 256


def _hold_ok(s, p, n, base):
    """Preconditions of the integer hold phase (csrc/env_step.cu: simulate_warp)."""
    sane = p["fast_pos"] >= 0 and p["medium_pos"] >= 0 and p["slow_pos"] >= 0
    if not (sane and s["speed"] == 0 and s["open"] == s["target"] and s["target"] <= 255.0 and 1.0 <= s["w"]):
        return False
    w = s["w"]
    thr = p["fast_w"] if w < p["fast_w"] else (p["medium_w"] if w < p["medium_w"] else p["slow_w"])
    bin_start = _from_bits(_bits(w) & 0x7FF0000000000000)
    return (thr <= 1048576.0 and 2.0 * s["target"] <= bin_start
            and base + WINDOW <= MAX_ITER and s["nochg"] + WINDOW <= NOCHG_LIMIT)


def _round_anchor(A, c):
    """Round-to-nearest-even of the anchor when w enters the next binade: A is a multiple of 2^c (units of the
    original ulp); the exact sum is congruent to A modulo 4*2^c, so the tie goes to the multiple of 4*2^c."""
    r = (A >> c) & 3
    if r & 1:
        A += (-1 if r == 1 else 1) << c
    return A


def _hold_stretch(s, p, nz, n, b0, stats):
    """Whole windows in integer arithmetic until the stage threshold is reached.  Returns (n, b0, fin)."""
    w_a = s["w"]
    thr = p["fast_w"] if w_a < p["fast_w"] else (p["medium_w"] if w_a < p["medium_w"] else p["slow_w"])
    e0 = ((_bits(w_a) >> 52) & 0x7FF) - 1023
    M = (_bits(w_a) & 0xFFFFFFFFFFFFF) | (1 << 52)
    u0 = 2.0 ** (e0 - 52)
    A, c = M, 0
    while 2.0 ** (e0 + c + 1) < thr:
        A = _round_anchor(A, c)
        c += 1
    Dc = (int(thr) << 31) - (A >> (21 - e0))
    ti = int(s["target"])
    xs = SMALL_MAX // ti if ti else 0xFFFFFFFF
    acc, nochg = 0, s["nochg"]
    try_run = ti != 0
    while True:
        if try_run:
            # multi-window run (csrc/env_step.cu, "fast run"): the rest of the current window plus Kr rows of 128 draws,
            # one sum, one compare; no complete noise block may be all-small; the count at the end comes from the last row
            try_run = False
            wend = 4 * b0 + WINDOW
            L = wend - n if n < wend else 0
            nstart = wend if n < wend else n
            if nstart & 3 == 0 and nochg + WINDOW <= NOCHG_LIMIT:
                R = (Dc - acc) / (ti * 2147483648.0)
                Kr = min(int((R - 3.0 * math.sqrt(R) - L) / 128.0), (MAX_ITER - nstart) >> 7)
                if Kr >= 1:
                    end = nstart + 128 * Kr
                    total = ti * sum(nz.word(t) for t in range(n, end))
                    bad = any(all(nz.word(4 * b + k) <= xs for k in range(4)) for b in range((n + 3) >> 2, end >> 2))
                    if not bad and acc + total < Dc:
                        t_last = max(t for t in range(128) if nz.word(end - 128 + t) > xs)
                        nochg, acc, n = 127 - t_last, acc + total, end
                        stats["run"] = stats.get("run", 0) + 1
                        stats["run_rows"] = stats.get("run_rows", 0) + Kr
        if n >= 4 * b0 + WINDOW:
            b0 = n >> 2
        base = 4 * b0
        if base + WINDOW > MAX_ITER or nochg + WINDOW > NOCHG_LIMIT:
            # a safety stop is within reach: materialise w (the anchor of the binade w is in) and leave
            A2, c2 = M, 0
            while 2.0 ** (e0 + c2 + 1) < thr and float(A2) * u0 + float(acc) * SCALE >= 2.0 ** (e0 + c2 + 1):
                A2 = _round_anchor(A2, c2)
                c2 += 1
            s["w"] = float(A2) * u0 + float(acc) * SCALE
            s["prev"], s["nochg"] = s["w"], nochg
            return n, b0, False
        skip = n - base
        x = [nz.word(base + t) for t in range(WINDOW)]
        P = [0 if t < skip else ti * x[t] for t in range(WINDOW)]
        total = sum(P)
        if acc + total < Dc:
            stats["fast"] += 1
            acc += total
            last_ns = max([t for t in range(skip, WINDOW) if x[t] > xs], default=-1)
            if any(x[t] <= xs for t in range(skip, WINDOW)):
                nochg = nochg + (WINDOW - skip) if last_ns < 0 else WINDOW - 1 - last_ns
            else:
                nochg = 0
            n = base + WINDOW
            continue
        stats["cross"] += 1
        before = acc
        for j in range(WINDOW):
            if before + P[j] >= Dc:
                break
            before += P[j]
        # the predecessor's anchor: boundaries that w_{j-1} has already passed
        A2, c2 = M, 0
        while 2.0 ** (e0 + c2 + 1) < thr and float(A2) * u0 + float(before) * SCALE >= 2.0 ** (e0 + c2 + 1):
            A2 = _round_anchor(A2, c2)
            c2 += 1
        w_prev = float(A2) * u0 + float(before) * SCALE
        w_new = w_prev + float(P[j]) * SCALE
        last_ns = max([t for t in range(skip, j) if x[t] > xs], default=-1)
        run_before = nochg + (j - skip) if last_ns < 0 else j - 1 - last_ns
        n = base + j + 1
        s["w"] = w_new
        lf, lm, ls = w_new < p["fast_w"], w_new < p["medium_w"], w_new < p["slow_w"]
        if not (lf or lm or ls):
            return n, b0, True
        assert w_new >= thr
        s["target"] = p["fast_pos"] if lf else (p["medium_pos"] if lm else p["slow_pos"])
        s["nochg"] = run_before + 1 if abs(w_new - w_prev) < 0.01 else 0
        s["prev"] = w_new
        return n, b0, n > MAX_ITER or s["nochg"] > NOCHG_LIMIT


def simulate_model(params, seed, env_id, step_idx, stats):
    """params in load_params order: target, fast_w, medium_w, slow_w, fast/medium/slow delay, fast/medium/slow opening, unload."""
    p = dict(fast_w=float(params[1]), medium_w=float(params[2]), slow_w=float(params[3]), fast_delay=params[4],
             medium_delay=params[5], slow_delay=params[6], fast_pos=float(params[7]), medium_pos=float(params[8]),
             slow_pos=float(params[9]))
    misc = O.philox(seed, env_id, step_idx, 0, 1)
    sysdelay = 15.0 + 35.0 * (misc[0] * (1.0 / 4294967296.0))
    nz = Noise(seed, env_id, step_idx)
    s = dict(w=0.0, open=0.0, speed=0, prev=0.0, target=p["fast_pos"], nochg=0, delay=0)
    n, fin, first, b0 = 0, False, True, 0
    while not fin:
        in_hold = s["speed"] == 0 and s["open"] == s["target"]
        if in_hold and not first and _hold_ok(s, p, n, (n & ~3) if n >= 4 * b0 + WINDOW else 4 * b0):
            n, b0, fin = _hold_stretch(s, p, nz, n, b0, stats)
            continue
        if not in_hold:
            n2 = _moving_phase(s, p, nz, n, first, sysdelay, stats)
            if n2 is not None:
                n, first = n2, False
                continue
        stats["generic"] += 1
        x = nz.word(n)
        n += 1
        fin = _generic_tick(s, p, x, sysdelay, n, first)
        first = False
    return n, (n + s["delay"] + params[10]) / 1000.0, s["w"]


def _check(params, seed, env_id, step_idx, stats):
    n, tt, fw = simulate_model(params, seed, env_id, step_idx, stats)
    rn, rtt, rfw, _ = O.simulate_run_philox(params, seed, env_id, step_idx)
    assert (n, tt) == (rn, rtt), (params, n, rn)
    assert _bits(fw) == _bits(rfw), (params, fw, rfw)
    return n


def test_hold_superblock_model_matches_oracle():
    rng = random.Random(5)
    stats = dict(fast=0, cross=0, generic=0, table=0)
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
    # adversarial sweep: thresholds on / next to powers of two, so stage exits coincide with binade crossings
    for k in range(400):
        pw = 2 ** rng.randint(9, 14)
        fw = pw + rng.choice([-2, -1, 0, 1, 2, 3, rng.randint(-200, 200)])
        sw = fw + rng.choice([1, 2, 5, rng.randint(1, 4000)])
        cases.append([sw, fw, rng.choice([0, fw + 1]), sw, rng.randint(1, 300), 100, 100, rng.randint(1, 100), rng.randint(0, 100),
                      rng.randint(1, 100), 300])
    for i, c in enumerate(cases[-400:]):
        ticks += _check(c, seed + 1, 5000 + i, i % 5, stats)
    assert stats["fast"] + stats["run_rows"] // 2 > 300 and stats["cross"] > 300, stats
    assert stats["table"] > 60 and stats["generic"] < 0.2 * ticks, stats
    assert stats["run"] > 60 and stats["run_rows"] > 300, stats


def test_moving_phase_small_increment_count():
    """Motor moves in which an increment falls below 0.011 (the stagnation count is kept tick by tick instead of being
    reset): the PID-effective constant action of configs[1] over many envs hits it in about 1 % of the runs."""
    stats = dict(fast=0, cross=0, generic=0, table=0)
    c = [25000, 18000, 0, 24900, 100, 100, 100, 35, 3, 3, 300]
    for i in range(600):
        _check(c, 0, i, 1, stats)
    assert stats.get("table_small", 0) >= 2 and stats["generic"] == 0, stats
