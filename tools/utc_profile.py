"""Development: in-kernel phase profile of ddpg_critic_grad_tc (CTA 0, epilogue thread 0), ERLFILL_UTC_PROF=1."""
import ctypes as C
import os
import sys

os.environ["ERLFILL_UTC_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import erlfill_b200 as E
from erlfill_b200 import _lib as L, ddpg, init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
a_sd, c_sd, _ = init.er_ddpg_state_dicts(E.conditions.EXPERIMENT_3["variant_25kg"]["bounds"], seed=0)
lrn = ddpg.DDPGLearner(a_sd, c_sd, batch_size=B, precision="fast")
lrn.use_graph = False
g = torch.Generator().manual_seed(0)
batch = torch.zeros(B, 20)
batch[:, 0:2] = torch.rand(B, 2, generator=g) * 5; batch[:, 2:12] = torch.rand(B, 10, generator=g) * 2 - 1
batch[:, 12] = torch.rand(B, generator=g) * 9000 - 3000; batch[:, 13:15] = torch.rand(B, 2, generator=g) * 5
batch[:, 16:19] = torch.rand(B, 3, generator=g)
batch = batch.cuda()
ne = torch.tensor([0.4, 0.5, 0.2], device="cuda")
for _ in range(3):
    lrn.update(batch, ne, 1)
torch.cuda.synchronize()
if os.environ.get("UTC_COLD") in ("1", "2"):      # the update as bench.py times it: on a cold L2 (1: L2 full of dirty lines, 2: clean lines)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush.fill_(1)
    torch.cuda.synchronize()
    if os.environ.get("UTC_COLD") == "2":
        big = torch.ones(256 << 18, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        print(float(big.sum()))
    torch.cuda.synchronize()
    lrn.update(batch, ne, 1)
    torch.cuda.synchronize()
    print("COLD L2")
out = (C.c_longlong * 256)()
L.lib().erlfill_debug_utc_profile(out)
t = np.array(out[:128], dtype=np.int64)
names = ["start", "stats", "T0 img", "T1 mma", "T1 epi", "T2 mma", "T2 epi", "T3 mma", "T3 epi", "T4 mma", "T4 head", "T5 mma", "T5 ln", "T6 mma", "T6 ln", "T7 mma",
         "T7 head", "C0 img", "C1 mma", "C1 ln+save", "C2 mma", "C2 ln+save", "C3 mma", "C3 loss", "dW8 drain", "B3 mma", "B3 lnbwd", "dW4a drain", "dW4b drain",
         "B2 mma", "B2 lnbwd", "(unused)", "dW0 drain"]
print("B = %d, ddpg_critic_grad_tc: cycles since kernel start / delta (CTA 0 of pair 0; the target chain T* runs on its own pairs in the split launch)" % B)
for i in range(1, 33):
    prev = max([t[j] for j in range(i) if t[j] > 0] or [0])
    if t[i] > 0 and prev > 0:
        print("%2d %-12s %9d %8d" % (i, names[i], t[i] - t[0], t[i] - prev))
t = np.array(out[128:256], dtype=np.int64)
an = ["start", "A0 img", "A1 mma", "A1 epi", "A2 mma", "A2 epi", "A3 mma", "A3 epi", "A4 mma", "A4 head", "C1 mma", "C1 ln", "C2 mma", "C2 ln", "C3 mma", "C3 dz3",
      "B3 mma", "B3 lnbwd", "B2 mma", "B2 lnbwd", "dx0 mma", "dz4", "dW4 drain", "da3 mma", "da3", "dW3 drain", "da2 mma", "da2", "dW2 drain", "da1 mma", "da1", "dW1 drain"]
print("ddpg_actor_grad_tc")
for i in range(1, 32):
    print("%2d %-12s %9d %8d" % (i, an[i], t[i] - t[0], t[i] - t[i - 1]))


cta = (C.c_longlong * 2048)()
L.lib().erlfill_debug_utc_cta(cta)
c = np.array(cta[:], dtype=np.int64).reshape(2, 256, 4)
for k, name in ((0, "critic"), (1, "actor")):
    used = c[k][:, 0] > 0
    t0 = c[k][used, 0].min()
    d = (c[k][used] - t0) / 1000.0
    print("%s kernel, %d CTAs, microseconds since the first CTA entered: entry min/max %.1f/%.1f  after setup %.1f/%.1f  epilogue done %.1f/%.1f  exit %.1f/%.1f"
          % (name, used.sum(), d[:, 0].min(), d[:, 0].max(), d[:, 1].min(), d[:, 1].max(), d[:, 2].min(), d[:, 2].max(), d[:, 3].min(), d[:, 3].max()))
    if k == 0:
        n = used.sum() // 2
        print("   critic-chain CTAs epilogue done %.1f..%.1f, target-chain CTAs %.1f..%.1f" % (d[:n, 2].min(), d[:n, 2].max(), d[n:, 2].min(), d[n:, 2].max()))
