"""bf16 tensor-core path vs the float32 path and the oracle's golden logits (run on a B200): logit error, KL, greedy divergence."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
import vqvae_wavenet_b200 as pkg
g = np.load(os.path.join(ROOT, "tests", "golden", "full.npz"))
cfg = O.Config(); w = O.make_weights(cfg, seed=1234, peaked=True)
eng = pkg.Engine(pkg.EngineConfig(), device=0, max_batch=64); eng.set_weights(w)
B, Tt = 4, 512
ze = O.synthetic_z_e(cfg, w, B, 64, seed=1235, kind="scaled")
_, cond = eng.encode_condition(ze, [0, 1, 2, 3])
x = O.synthetic_audio(B, Tt, seed=1237)
want = g["teacher_logits"]
for prec in ("fp32", "bf16"):
    eng.set_precision(prec)
    lg = eng.teacher_forced(x, cond[:, :Tt // 64])
    print(prec, eng.last_kernel_name, "max|dlogit|/max|logit| =", np.abs(lg[:, ::32] - want).max() / np.abs(want).max())
eng.set_precision("fp32")
a0, i0 = eng.generate(cond, 4096, mode="greedy")
eng.set_precision("bf16")
a1, i1 = eng.generate(cond, 4096, mode="greedy")
agree = (i0 == i1)
first = [int(np.argmin(r)) if not r.all() else len(r) for r in agree]
print("greedy bf16 vs fp32: first divergence per stream", first, "overall agreement", agree.mean())

# distribution-level view of the same teacher-forced run: KL(oracle || kernel) per step, top-1 agreement
def _logp(l):
    l = l.astype(np.float64)
    l = l - l.max(-1, keepdims=True)
    return l - np.log(np.exp(l).sum(-1, keepdims=True))
for prec in ("fp32", "bf16"):
    eng.set_precision(prec)
    lg = eng.teacher_forced(x, cond[:, :Tt // 64])[:, ::32]
    lp_ref, lp = _logp(want), _logp(lg)
    kl = (np.exp(lp_ref) * (lp_ref - lp)).sum(-1)
    print(prec, "KL(oracle||kernel) mean %.3e max %.3e nats; top-1 agreement %.4f"
          % (kl.mean(), kl.max(), (lg.argmax(-1) == want.argmax(-1)).mean()))
