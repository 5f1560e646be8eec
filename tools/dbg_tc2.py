import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import erlfill_b200
from erlfill_b200 import nets
G = np.load("tests/golden/nets_golden.npz")
W = {k[len("trans."):]: torch.from_numpy(np.array(G[k])) for k in G.files if k.startswith("trans.")}
P = nets.TransformerParams(W)
n, reps = int(sys.argv[1]), int(sys.argv[2])
seq = torch.rand(n, 20, 3, device="cuda")
for i in range(reps):
    out = nets.emotion_forward(P, seq=seq, precision="bf16")
    torch.cuda.synchronize()
    print("launch", i, "ok", float(out.sum()))
