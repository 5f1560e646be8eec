"""The callers either side of the hot path: the training-loop body of `train_ddpg`
and the evaluation loops of `experiment_runner.py`, without the TensorBoard / logging control plane.

  train_ddpg(env, agent, ...)       ddpg_agent.py:628-729 — the exact call order across the env/agent boundary
                                    (reset, ou reset+anneal, act, step, emotion.update, remember, learn, get_emotion),
                                    for the N = 1 drop-ins of dropin.py / envs.py
  train_erl_fill(env, agent, ...)   ERL_Fill_agent.py:526-541 — EmotionSAC's loop body
  test_agent(env, agent, ...)       experiment_runner.py:541-556 (and :160-180) — deterministic test episodes and the
                                    per-episode record the reference writes to its CSV (`results_detail`, :566-577)
  evaluate_batched(agent, ...)      the same evaluation for N envs in lock-step on the device: one episode per env,
                                    per-condition means of reward / weight error / time (`results`, :591)
  write_results_csv / write_summary the reference's two output formats (:595-604)
"""
import csv

import numpy as np
import torch


def train_ddpg(env, agent, episodes=1000, max_steps=500, on_step=None):
    """Loop body of train_ddpg; returns the list of episode rewards (ddpg_agent.py:628-732)."""
    rewards = []
    env.max_steps = max_steps
    for ep in range(episodes):
        state = env.reset()
        episode_reward = 0
        agent.ou_noise.reset()
        agent.ou_noise.anneal()
        for step in range(max_steps):
            action = np.transpose(agent.act(state))
            next_state, reward, done, info = env.step(action)
            movement = next_state - state
            agent.emotion.update(reward, movement)
            agent.remember(state, action, reward, next_state, done)
            state = next_state
            episode_reward += reward
            loss = agent.learn()
            cur_emotion = np.array(agent.emotion.get_emotion()).flatten()
            if on_step is not None:
                on_step(ep, step, reward, loss, cur_emotion, info)
            if done:
                break
        rewards.append(episode_reward)
        agent.train_step += 1
    return rewards


def train_erl_fill(env, agent, episodes=1000, max_steps=500, on_step=None):
    """Loop body of train_erl_fill (ERL_Fill_agent.py:526-541, 619): note `emotion.update(reward, action)` — the action is
    the "movement" — and `float(done)` in the stored tuple."""
    rewards = []
    env.max_steps = max_steps
    for ep in range(episodes):
        state = env.reset()
        ep_reward = 0
        for step in range(max_steps):
            action = agent.act(state)
            next_state, reward, done, info = env.step(action)
            agent.emotion.update(reward, action)
            agent.remember(state, action, reward, next_state, float(done))
            update_info = agent.update()
            state = next_state
            ep_reward += reward
            if on_step is not None:
                on_step(ep, step, reward, update_info, np.array(agent.emotion.get_emotion()).flatten(), info)
            if done:
                break
        rewards.append(ep_reward)
        agent.train_step += 1
    return rewards


def _select(agent, algo, state):
    """Action selection of the reference's test loops (experiment_runner.py:160-172, 545-552)."""
    if algo in ("sac", "emotion_sac", "ppo"):
        a = agent.act(state, deterministic=True)
    else:
        a = agent.act(state, add_noise=False)
    return a[0] if isinstance(a, tuple) else a


def test_agent(env, agent, algo="er_ddpg", episodes=10, max_steps=100, condition="Default"):
    """experiment_runner.py:541-592: (avg_reward, avg_weight_err, avg_time), results_detail rows."""
    detail, tot = [], np.zeros(3)
    for ep in range(episodes):
        state = env.reset()
        ep_reward, info = 0.0, {}
        for _ in range(max_steps):
            state, reward, done, info = env.step(_select(agent, algo, state))
            ep_reward += reward
            if done:
                break
        em = np.array(agent.emotion.get_emotion()).flatten() if hasattr(agent, "emotion") else np.full(3, np.nan)
        row = {"Condition": condition, "Episode": ep + 1, "Reward": ep_reward, "WeightErr": info.get("weight_error", 0.0),
               "Time": info.get("total_time", 0.0), "SlowWeight": info.get("action", {}).get("slow_weight", 0.0),
               "Curiosity": float(em[0]), "Conservativeness": float(em[1]), "Anxiety": float(em[2])}
        detail.append(row)
        tot += (ep_reward, row["WeightErr"], row["Time"])
    return tuple(tot / episodes), detail


test_agent.__test__ = False     # not a pytest test


def evaluate_batched(agent, max_steps=100, condition_names=None):
    """Deterministic evaluation (`add_noise=False` / `deterministic=True`) of a batched agent (VecERDDPG, VecTD3,
    VecSAC) on its VecFillEnv: every env plays ONE episode from a fresh reset; an env that finishes stops
    accumulating.  Returns {condition: (mean episode reward, mean final weight error, mean final time, n_env)} —
    the per-condition triple of experiment_runner.py:591 — and the per-env tensors."""
    env = agent.env
    n, dev = env.n_env, env.device
    env.max_steps = max_steps
    prev_auto = env.auto_reset
    env.auto_reset = False
    if hasattr(agent, "reset"):
        agent.reset()
    else:
        env.reset()
    alive = torch.ones(n, dtype=torch.bool, device=dev)
    ep_reward = torch.zeros(n, dtype=torch.float64, device=dev)
    werr = torch.zeros(n, dtype=torch.float64, device=dev)
    ftime = torch.zeros(n, dtype=torch.float64, device=dev)
    target = env.target_weight_per_env.to(torch.float64)
    for _ in range(max_steps):
        if hasattr(agent, "rollout_eval_step"):
            _, reward, done, info = agent.rollout_eval_step()
        else:
            a = agent.act_batch(env.obs, **({"deterministic": True} if "SAC" in type(agent).__name__ else {"noise_scale": 0.0}))
            _, reward, done, info = env.step(a)
        ep_reward += torch.where(alive, reward.to(torch.float64), torch.zeros_like(ep_reward))
        werr = torch.where(alive, (info["final_weight"].to(torch.float64) - target).abs(), werr)
        ftime = torch.where(alive, info["total_time"].to(torch.float64), ftime)
        alive &= ~done.bool()
        if not bool(alive.any()):
            break
    env.auto_reset = prev_auto
    cid = env.cond_id.long()
    out = {}
    for c in range(len(env.conditions)):
        m = cid == c
        k = int(m.sum())
        if k:
            name = condition_names[c] if condition_names else f"condition_{c}"
            out[name] = (float(ep_reward[m].mean()), float(werr[m].mean()), float(ftime[m].mean()), k)
    return out, {"reward": ep_reward, "weight_error": werr, "total_time": ftime}


DETAIL_COLUMNS = ["Condition", "Episode", "Reward", "WeightErr", "Time", "SlowWeight", "Curiosity", "Conservativeness", "Anxiety"]


def write_results_csv(path, detail):
    """`df_detail.to_csv(..., index=False, encoding="utf-8-sig")` (experiment_runner.py:595-596)."""
    with open(path, "w", newline="", encoding="utf-8-sig") as f:
        w = csv.DictWriter(f, fieldnames=DETAIL_COLUMNS, lineterminator="\n")
        w.writeheader()
        # pandas infers one dtype per column: a column whose values are all np.float32 (SlowWeight: the clipped action) stays binary32
        # and prints with the shortest binary32 repr ('24944.707'); any other numeric column becomes binary64, so a np.float32 reward in
        # a column that also holds Python floats prints as its exact binary64 value (2985.782470703125, not '2985.7825')
        f32_cols = {k for k in DETAIL_COLUMNS if detail and all(isinstance(r[k], np.float32) for r in detail)}
        for row in detail:
            w.writerow({k: (v if isinstance(v, (str, int)) else (str(v) if k in f32_cols else float(v))) for k, v in row.items()})


def write_summary(path, results):
    """exp3_summary.log (experiment_runner.py:599-603); results = [(name, reward, weight_err, time), ...]."""
    with open(path, "w", encoding="utf-8") as f:
        f.write("Condition\tReward\tWeightErr(g)\tTime(s)\n")
        for name, score, err, t in results:
            f.write(f"{name}\t{score:.2f}\t{err:.2f}\t{t:.2f}\n")
