import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import vqvae_wavenet_b200 as pkg
from oracle import oracle as O
cfg = O.Config(); w = O.make_weights(cfg, seed=1234)
eng = pkg.Engine(pkg.EngineConfig(), device=0, max_batch=1)
eng.set_weights(w)
n = 1 << 20
for kind in ("normal", "scaled", "near_code"):
    z = O.synthetic_z_e(cfg, w, n // 1024, 1024, seed=1235, kind=kind)
    eng.vq_upload(z.reshape(-1, 64))
    eng.set_vq_kernel("tensor")
    for _ in range(3):
        eng.vq_resident(n)
    print(kind, "ms", eng.last_kernel_ms, flush=True)
