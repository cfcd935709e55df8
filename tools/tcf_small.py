"""smallest run of the tensor-core kernel (development): python tools/tcf_small.py [T] [B] [full]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vqvae_wavenet_b200 as pkg
from oracle import oracle as O
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
full = len(sys.argv) > 3
wn = None if full else dict(num_cycles=2, num_cycle_layers=3, dilation_rates=[1, 2, 4, 1, 2, 4])
cfg = O.Config(wavenet=wn)
w = O.make_weights(cfg, seed=1234)
eng = pkg.Engine(pkg.EngineConfig(wavenet=wn), device=0, max_batch=16)
eng.set_weights(w)
ze = O.synthetic_z_e(cfg, w, B, max((T + 63) // 64, 1), seed=1235, kind="scaled")
_, cond = eng.encode_condition(ze, [b % 4 for b in range(B)])
x = O.synthetic_audio(B, T, seed=1237)
out = {}
for prec in ("fp32", "tc"):
    eng.set_precision(prec)
    out[prec] = eng.teacher_forced(x, cond)
    print(prec, eng.last_kernel_name, "%.3f ms" % eng.last_kernel_ms, flush=True)
scale = np.abs(out["fp32"]).max()
d = np.abs(out["tc"] - out["fp32"]).max(axis=(0, 2)) / scale
print("per-step max |tc - fp32| / max|logit|:", np.array2string(d[:16], precision=2), "worst", d.max(), flush=True)
eng.close()
