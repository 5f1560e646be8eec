"""Development timing of the EmotionTransformer kernels (not the bench contract): python tools/bench_emotion.py [n_env]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import erlfill_b200  # noqa: F401
from erlfill_b200 import nets

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 65536
from erlfill_b200 import _lib as L0
DROP = L0.DROPOUT_PHILOX if "--philox" in sys.argv else L0.DROPOUT_OFF        # --philox: train() mode
G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "nets_golden.npz"))
W = {k[len("trans."):]: torch.from_numpy(np.array(G[k])) for k in G.files if k.startswith("trans.")}
P = nets.TransformerParams(W)
seq = torch.rand(n, 20, 3, device="cuda")
out = torch.empty(n, 3, device="cuda")
import ctypes as C
from erlfill_b200 import _lib as L
prof = (C.c_ulonglong * 16)()
for prec, reps in (("bf16", 10),) + ((("fp32", 2),) if "--fp32" in sys.argv else ()):
    for _ in range(2):
        nets.emotion_forward(P, seq=seq, precision=prec, out=out, dropout=DROP, seed=5)
    torch.cuda.synchronize()
    L.lib().erlfill_debug_tc_profile(None, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        nets.emotion_forward(P, seq=seq, precision=prec, out=out, dropout=DROP, seed=5)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%s: n=%d %.3f ms  %.2f M env/s  %.1f TFLOP/s (21.83 MFLOP/env)" % (prec, n, ms, n / ms / 1e3, n * 21.83e6 / ms / 1e9))
    if prec == "bf16":
        L.lib().erlfill_debug_tc_profile(prof, 0)
        names = ["image store", "wait qkv", "qkv epilogue", "attention", "wait out_proj", "LN1+store", "FFN loop", "wait Y+LN2"]
        names += ["I:non-FFN", "I:wait W", "I:issue MMA1", "I:wait P", "I:issue MMA2", "P:issue", "P:wait empty", "I:W-not-ready count"]
        tot = sum(prof[:8])
        tiles, cps = C.c_int(), C.c_int()
        L.lib().erlfill_debug_tc_config(C.byref(tiles), C.byref(cps))
        groups = -(-n // (6 * tiles.value))
        n_it = -(-groups // min(groups, 148 * cps.value))
        print("kernel shape: %d tile(s)/CTA, %d CTA(s)/SM" % (tiles.value, cps.value))
        print("block 0 phase cycles per layer-step (%d layer-steps/launch):" % (n_it * 4), ", ".join(
            "%s %.0f" % (nm, v / (reps * n_it * 4)) for nm, v in zip(names, prof)), "| total %.0f" % (tot / (reps * n_it * 4)))
