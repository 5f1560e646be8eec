"""Operating conditions of the filling line (bounds tables), restated from the reference:
WeightEnv.py:32-46, MutiConditionEnv.py:47-61, experiment_runner.py:417-478."""

SINGLE_25KG = {
    "target_weight": 25000, "target_weight_err": 25, "target_time": 2.5,
    "bounds": {"fast_weight": (18000, 23000), "medium_weight": (0, 1), "slow_weight": (24900, 25025),
               "fast_opening": (35, 60), "medium_opening": (3, 5), "slow_opening": (3, 7),
               "fast_delay": (100, 300), "medium_delay": (100, 200), "slow_delay": (100, 200),
               "unload_delay": (300, 500)}}

MULTI_DEFAULT_BOUNDS = {"fast_weight": (7000, 18000), "medium_weight": (0, 1), "slow_weight": (24900, 25000),
                        "fast_opening": (35, 75), "medium_opening": (3, 5), "slow_opening": (5, 20),
                        "fast_delay": (100, 300), "medium_delay": (100, 200), "slow_delay": (100, 200),
                        "unload_delay": (300, 500)}

# class defaults of MutiConditionEnv.WeightEnv.__init__ (note target_time ends up 2, :80-82)
MULTI_CLASS_DEFAULT = {"target_weight": 25000, "target_weight_err": 25, "target_time": 2,
                       "bounds": MULTI_DEFAULT_BOUNDS}

EXPERIMENT_3 = {
    "variant_25kg": {"target_weight": 25000, "target_weight_err": 25, "target_time": 2.5,
                     "bounds": dict(MULTI_DEFAULT_BOUNDS)},
    "variant_20kg": {"target_weight": 20000, "target_weight_err": 20, "target_time": 2.0,
                     "bounds": {"fast_weight": (6000, 15000), "medium_weight": (0, 1), "slow_weight": (19800, 20000),
                                "fast_opening": (30, 70), "medium_opening": (3, 5), "slow_opening": (5, 18),
                                "fast_delay": (80, 250), "medium_delay": (80, 180), "slow_delay": (80, 180),
                                "unload_delay": (280, 480)}},
    "variant_15kg": {"target_weight": 15000, "target_weight_err": 15, "target_time": 1.5,
                     "bounds": {"fast_weight": (5000, 12000), "medium_weight": (0, 1), "slow_weight": (14800, 15000),
                                "fast_opening": (25, 65), "medium_opening": (2, 4), "slow_opening": (4, 15),
                                "fast_delay": (60, 200), "medium_delay": (60, 150), "slow_delay": (60, 150),
                                "unload_delay": (250, 450)}},
}

# The RLS-PID agent's effective (constant) output on WeightEnv: every action clipped to its lower
# bound (rls_pidagent.py:142-174 + WeightEnv.py:405-407)
PID_EFFECTIVE_ACTION = [18000.0, 0.0, 24900.0, 35.0, 3.0, 3.0, 100.0, 100.0, 100.0, 300.0]
