"""Development: where the time of one N = 1 train_ddpg step goes (host wall clock per call, device synchronised)."""
import os, sys, time, random, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import erlfill_b200 as E
from erlfill_b200 import loops
cfg = E.conditions.EXPERIMENT_3["variant_25kg"]
torch.manual_seed(0); random.seed(0); np.random.seed(0)
env = E.MultiConditionWeightEnv({k: cfg[k] for k in ("target_weight", "target_weight_err", "target_time", "bounds")})
agent = E.DDPGAgent(env, use_td_error=True)
agent.emotion = E.EmotionModule()
env.attach_agent(agent)
acc = collections.defaultdict(lambda: [0, 0.0])
def wrap(obj, name, tag, sync=True):
    f = getattr(obj, name)
    def g(*a, **k):
        if sync: torch.cuda.synchronize()
        t = time.perf_counter(); r = f(*a, **k)
        if sync: torch.cuda.synchronize()
        acc[tag][0] += 1; acc[tag][1] += time.perf_counter() - t
        return r
    setattr(obj, name, g)
loops.train_ddpg(env, agent, episodes=4, max_steps=50)
sync = "--nosync" not in sys.argv
wrap(agent, "act", "agent.act", sync); wrap(env, "step", "env.step", sync); wrap(agent.emotion, "update", "emotion.update", sync)
wrap(agent, "remember", "agent.remember", sync); wrap(agent, "learn", "agent.learn", sync); wrap(agent.emotion, "get_emotion", "emotion.get_emotion", sync)
wrap(env, "reset", "env.reset", sync)
n = [0]
t0 = time.perf_counter()
loops.train_ddpg(env, agent, episodes=6, max_steps=50, on_step=lambda *a: n.__setitem__(0, n[0] + 1))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("steps %d, %.3f ms/step" % (n[0], dt / n[0] * 1e3))
for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("%-22s calls/step %.2f  ms/call %.3f  ms/step %.3f" % (k, c / n[0], t / c * 1e3, t / n[0] * 1e3))
