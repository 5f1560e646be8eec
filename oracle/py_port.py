"""Pure-Python restatement of VirtualWeightController.simulate_run (VirtualWeightController.py:127-244)
with a frozen clock — TEST INFRASTRUCTURE ONLY.  It exists to time "the reference's own language" next
to the C oracle in bench.py's cpu_baseline (the reference itself is a CPython loop) and to cross-check
the C oracle on small cases.  Noise comes from a random.Random instance, like the reference."""


def simulate_run(p, rng):
    """p: dict with the 11 integer parameters of load_params (:71-83); rng: random.Random."""
    system_delay_time = rng.uniform(15, 50)
    fast_pos, medium_pos, slow_pos = p["fast_opening"], p["medium_opening"], p["slow_opening"]
    fast_weight, medium_weight, slow_weight = p["fast_weight"], p["medium_weight"], p["slow_weight"]
    delays = (p["fast_delay"], p["medium_delay"], p["slow_delay"])
    w, opening, speed, direction = 0, 0, 0, 1
    target = fast_pos
    delay_finish = False
    tick = iteration = no_change = n = 0
    prev = w
    while True:
        tick += 1
        dt = speed / 100000 if speed > 0 else 0
        dd = 0.5 * 100000 * dt ** 2
        if abs(opening - target) > 1:
            if opening < target:
                speed += 5000000 / 1000 if opening < target - dd else -(100000 / 1000)
                direction = 1
            else:
                speed += 5000000 / 1000 if opening > dd else -(100000 / 1000)
                direction = -1
        else:
            speed = 0
            opening = target
        opening += speed * direction / 1000
        w += opening * (1.0 + rng.uniform(-1, 1))
        n += 1
        if w < 0:
            w = 0
        if w < fast_weight:
            stage = 0
            target = fast_pos
        elif w < medium_weight:
            stage = 1
            target = medium_pos
        elif w < slow_weight:
            stage = 2
            target = slow_pos
        else:
            break
        if not delay_finish:
            tick += delays[stage]
            delay_finish = True
            w += target * system_delay_time
        iteration += 1
        no_change = no_change + 1 if abs(w - prev) < 0.01 else 0
        prev = w
        if iteration > 5000 or no_change > 500:
            break
    return n, (tick + p["unload_delay"]) / 1000.0, w
