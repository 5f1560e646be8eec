import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import erlfill_b200
from erlfill_b200 import nets
G = np.load("tests/golden/nets_golden.npz")
W = {k[len("trans."):]: torch.from_numpy(np.array(G[k])) for k in G.files if k.startswith("trans.")}
P = nets.TransformerParams(W)
g = torch.Generator().manual_seed(5)
for n in (12, 13, 24, 148 * 12, 148 * 12 * 3 + 5):
    seq = torch.rand(n, 20, 3, generator=g).cuda()
    ref = nets.emotion_forward(P, seq=seq)
    outs = [nets.emotion_forward(P, seq=seq, precision="bf16").clone() for _ in range(4)]
    torch.cuda.synchronize()
    for i in range(1, 4):
        d = (outs[i] != outs[0]).any(dim=1).nonzero().flatten().cpu().numpy()
        if len(d):
            grp = d // 12
            print("n=%d run %d: %d envs differ; env%%12 hist %s; tile hist %s; iteration hist %s; max diff %.2e" % (
                n, i, len(d), np.bincount(d % 12, minlength=12), np.bincount((d % 12) // 6, minlength=2),
                np.bincount(grp // 148, minlength=4), (outs[i] - outs[0]).abs().max().item()))
        else:
            print("n=%d run %d identical" % (n, i))
    print("  max err vs fp32", (outs[0] - ref).abs().max().item())
