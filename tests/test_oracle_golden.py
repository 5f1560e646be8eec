"""Pins the CPU oracle (oracle/erlfill_oracle.c) to fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, generator: oracle/gen_golden.py).  CPU only.

The reference draws from CPython's Mersenne Twister; the tests regenerate the identical stream
with random.Random(seed) (stdlib, platform independent) and feed it to the oracle as arrays.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle_lib as O

MAXDRAW = 5002


def _stream(rng, n):
    """n draws of random.uniform(-1, 1) == -1 + 2*random() (CPython Lib/random.py)."""
    return np.array([-1.0 + 2.0 * rng.random() for _ in range(n)], np.float64)


def test_pow_equals_mul_for_reachable_speeds():
    # VirtualWeightController.py:93 uses decelerate_time ** 2 (libm pow); the oracle and the CUDA
    # table use dt*dt.  They agree for every speed/100 < 2759 (first libm mis-rounding), far above
    # the peak speed any opening target <= 100 can reach (9200 in the fixtures).
    for k in range(0, 2759):
        dt = float(k * 100) / 100000
        assert 0.5 * 100000 * dt ** 2 == 50000.0 * (dt * dt)


def test_sim_golden_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "sim_golden.npz"))
    assert g["peak_speed"].max() < 275900
    stops = {"cap": 0, "stagnant": 0, "normal": 0}
    for i in range(len(g["seeds"])):
        rng = random.Random(int(g["seeds"][i]))
        sysdelay = 15 + (50 - 15) * rng.random()          # random.uniform(15, 50)
        u = _stream(rng, int(g["iters"][i]) + 8)
        n, tt, fw, tick = O.simulate_run_noise(g["params"][i], sysdelay, u)
        assert n == int(g["iters"][i]), (i, n, g["iters"][i])
        assert tt == g["total_time"][i], i
        assert fw == g["final_weight"][i], i                # binary64 bit-exact
        stops["cap" if n == 5001 else "stagnant" if n <= 502 and fw < 1 else "normal"] += 1
    assert stops["cap"] > 0 and stops["stagnant"] > 0 and stops["normal"] > 300


def test_survey_kats():
    # SURVEY.md section 4 simulator KATs
    p = [25000, 12000, 0, 24950, 200, 150, 150, 55, 4, 12, 400]
    want = {0: (1257, 1.857, 24959.936875831205), 1: (1242, 1.842, 24956.853720840143),
            2: (1230, 1.83, 24958.48352671813)}
    for s, (n_w, tt_w, fw_w) in want.items():
        rng = random.Random(s)
        sysdelay = 15 + 35 * rng.random()
        n, tt, fw, _ = O.simulate_run_noise(p, sysdelay, _stream(rng, 1300))
        assert (n, tt, fw) == (n_w, tt_w, fw_w)


def _conds(bounds_json):
    B = json.loads(bounds_json)
    return {k: O.make_condition(v["target_weight"], v["target_weight_err"], v["target_time"], v["bounds"])
            for k, v in B.items()}, B


def test_env_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "env_golden.npz"))
    conds, _ = _conds(str(g["bounds_json"]))
    n_checked = n_done = n_f32 = 0
    for key in g["runs"]:
        kind = O.ENV_SINGLE if str(g[key + "_kind"]) == "single" else O.ENV_MULTI
        c = conds[str(g[key + "_name"])]
        max_steps = int(g[key + "_max_steps"])
        rng = random.Random(int(g[key + "_seed"]))
        st = O.EnvState()
        resets = g[key + "_reset_state"]
        ri = 0
        A = g[key + "_actions"]
        for t in range(len(A)):
            if g[key + "_reset_before"][t]:
                u0, u1 = rng.random(), rng.random()         # uniform(0, x) = 0 + x*random()
                s0 = O.env_reset_cold(c, st, u0, u1)
                assert np.array_equal(s0, resets[ri]), (key, t)
                ri += 1
            sysdelay = 15 + 35 * rng.random()
            # draw generously, then rewind the MT stream to what the step consumed
            state_before = rng.getstate()
            u = _stream(rng, MAXDRAW)
            o = O.env_step_noise(kind, c, st, max_steps, A[t], sysdelay, u)
            rng.setstate(state_before)
            _stream(rng, o.ticks)
            assert o.final_weight == g[key + "_final_weight"][t], (key, t)
            assert o.total_time == g[key + "_total_time"][t], (key, t)
            assert o.weight_error == g[key + "_weight_error"][t]
            assert np.array_equal(np.array(o.next_state, np.float32), g[key + "_next_state"][t])
            assert bool(o.done) == bool(g[key + "_done"][t]), (key, t)
            ref_r = g[key + "_reward"][t]
            if g[key + "_reward_is_f32"][t]:
                n_f32 += 1
                assert np.float32(o.reward) == np.float32(ref_r), (key, t, o.reward, ref_r)
            else:
                assert abs(o.reward64 - ref_r) <= 1e-9 * max(1.0, abs(ref_r)), (key, t, o.reward64, ref_r)
            n_checked += 1
            n_done += int(o.done)
        assert ri == len(resets)
    assert n_checked >= 300 and n_done >= 15 and n_f32 >= 50


def test_emotion_update_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "emotion_update_golden.npz"))
    cur = np.array([0.5, 0.5, 0.0])
    pushed_hist = []
    for t in range(len(g["rewards"])):
        cur, pushed = O.emotion_update(cur, float(g["rewards"][t]), g["moves"][t])
        pushed_hist.append(pushed)
        # the reference (numpy>=2) runs parts of M1 in binary32 when the reward is np.float32;
        # the oracle is binary64 throughout: contract is 1e-5 relative / 2e-7 absolute
        np.testing.assert_allclose(cur, g["current"][t], rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(np.array(pushed_hist[-20:]), g["history_last20"], rtol=1e-5, atol=2e-7)


def test_reset_replay_matches_numpy():
    rng = np.random.RandomState(3)
    c = O.make_condition(20000, 20, 2.0, {k: (0, 1) for k in O.ACTION_ORDER})
    for _ in range(200):
        s10 = (rng.uniform(0, 6, (10, 2))).astype(np.float32)
        raw = np.mean([s for s in s10], axis=0).astype(np.float32)     # MutiConditionEnv.py:147-148
        want = np.array([raw[0] / 20, raw[1] / 2.0], dtype=np.float32)  # :151-154
        st = O.EnvState()
        got = O.env_reset_replay(c, st, s10)
        assert np.array_equal(got, want)
        assert st.episode_counter == 1


def test_ou_and_normalize_match_numpy():
    rng = np.random.RandomState(5)
    scale = np.array([5500, .5, 50, 20, 1, 7.5, 100, 50, 50, 100], np.float32)
    offset = np.array([12500, .5, 24950, 55, 4, 12.5, 200, 150, 150, 400], np.float32)
    state = np.zeros(10, np.float32)
    sigma = 0.6
    for t in range(50):
        randn = rng.randn(10).astype(np.float32)
        scaled = (offset + scale * rng.uniform(-1, 1, 10)).astype(np.float32)
        # ddpg_agent.py:774-778 and :352-356 restated with numpy
        ref_state = state.copy()
        dx = 0.15 * (np.zeros(10, np.float32) - ref_state)
        dx += sigma * randn
        ref_state += dx
        ref_final = np.clip(scaled + ref_state * scale, offset - scale, offset + scale)
        state, final = O.ou_act(state, sigma, randn, scaled, scale, offset)
        assert np.array_equal(state, ref_state) and np.array_equal(final, ref_final)
        norm = O.normalize_action(final, scale, offset)
        ref_norm = np.clip((final - offset) / scale, -1.0, 1.0)
        assert np.array_equal(norm, ref_norm)
        sigma = max(0.1, sigma * 0.9995)


def test_uniform_priority_choice_equals_floor():
    # ddpg_agent.py:240-251 with all priorities 1e-4
    for n in (130, 1000, 4097):
        pri = np.array([1e-4 ** 0.6] * n, dtype=np.float32) * 0 + np.float32(1e-4)
        probs = pri / pri.sum()
        np.random.seed(11)
        ref = np.random.choice(n, 128, p=probs)
        np.random.seed(11)
        u = np.random.random_sample(128)
        assert np.array_equal(O.sample_indices_uniform(u, n), ref)


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter = key = 0 and the all-ones vector
    assert O.philox(0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox(0xffffffffffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


def test_python_port_matches_golden(golden_dir):
    from oracle import py_port
    g = np.load(os.path.join(golden_dir, "sim_golden.npz"))
    names = ["target_weight", "fast_weight", "medium_weight", "slow_weight", "fast_delay", "medium_delay",
             "slow_delay", "fast_opening", "medium_opening", "slow_opening", "unload_delay"]
    for i in range(0, len(g["seeds"]), 7):
        p = {k: int(v) for k, v in zip(names, g["params"][i])}
        n, tt, fw = py_port.simulate_run(p, random.Random(int(g["seeds"][i])))
        assert (n, tt, fw) == (int(g["iters"][i]), g["total_time"][i], g["final_weight"][i])


def test_np_power_equals_mul_for_onboard_speeds():
    # WeightEnv.py:189 / MutiConditionEnv.py:192 use np.power(decelerate_time, 2); the oracle and the CUDA kernel use dt*dt.
    # With the speed cap the on-board motor never exceeds 10000 + 4900; checked far beyond, both signs.
    for k in range(-400, 401):
        dt = float(k * 100) / 100000
        assert float(0.5 * 100000 * np.power(dt, 2)) == 50000.0 * (dt * dt)


def test_onboard_golden_bit_exact(golden_dir):
    """Row N1: the reference's own simulate_run driving the env classes' on-board motor_simulation / weight_simulation
    (oracle/gen_golden.py onboard) against oracle/erlfill_oracle.c:simulate_core_onboard."""
    g = np.load(os.path.join(golden_dir, "onboard_golden.npz"))
    assert g["peak_speed"].max() == 10000.0                # the speed cap binds in the fixtures
    stops = {"cap": 0, "stagnant": 0, "normal": 0}
    differs = 0
    for i in range(len(g["seeds"])):
        rng = random.Random(int(g["seeds"][i]))
        sysdelay = 15 + (50 - 15) * rng.random()
        u = _stream(rng, int(g["iters"][i]) + 8)
        n, tt, fw, tick = O.simulate_onboard_noise(g["params"][i], sysdelay, u)
        assert n == int(g["iters"][i]), (i, n, g["iters"][i])
        assert tt == g["total_time"][i], i
        assert fw == g["final_weight"][i], i
        stops["cap" if n == 5001 else "stagnant" if n <= 502 and fw < 1 else "normal"] += 1
        n0, tt0, fw0, _ = O.simulate_run_noise(g["params"][i], sysdelay, u)          # the offline controller on the same noise
        differs += (n0, fw0) != (n, fw)
    assert stops["cap"] > 0 and stops["stagnant"] > 0 and stops["normal"] > 150
    assert differs > 150                                   # int() truncation makes it a different trajectory, not a relabelling
