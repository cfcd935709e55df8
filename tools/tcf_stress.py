"""stress (development): repeated launches of a generation kernel.
tcf_stress.py rounds T [precision] [same|fresh] [B]   same: one engine for all rounds; fresh: a new engine per round"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vqvae_wavenet_b200 as pkg
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
prec = sys.argv[3] if len(sys.argv) > 3 else "tc"
how = sys.argv[4] if len(sys.argv) > 4 else "fresh"
B = int(sys.argv[5]) if len(sys.argv) > 5 else 64
cfg = O.Config()
w = O.make_weights(cfg, seed=1234, peaked=True)
ze = O.synthetic_z_e(cfg, w, B, T // 64, seed=1235, kind="scaled")


def make():
    eng = pkg.Engine(pkg.EngineConfig(), device=0, max_batch=B)
    eng.set_weights(w)
    _, cond = eng.encode_condition(ze, [b % 4 for b in range(B)])
    eng.upload_condition(cond)
    eng.set_precision(prec)
    return eng


eng = make() if how == "same" else None
for i in range(n):
    if how != "same":
        eng = make()
    t0 = time.time()
    try:
        eng.generate_resident(B, T // 64, T, mode="greedy")
    except Exception as e:
        print("round %d FAILED after %.2f s: %s" % (i, time.time() - t0, e), flush=True)
        sys.exit(1)
    print("round %d ok %.2f s kernel %.2f ms = %.2f us/step" % (i, time.time() - t0, eng.last_kernel_ms, eng.last_kernel_ms * 1e3 / T), flush=True)
    if how != "same":
        eng.close()
print("all rounds ok", flush=True)
