"""per-step time of the tensor-core kernel against the number of layers (development): the weight stream of a cluster is
2.77 MB per layer, so the slope shows what a stage costs while the stream fits the L2 and after it stops fitting.
python tools/tcf_layers.py [B] [T] [L ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vqvae_wavenet_b200 as pkg
from oracle import oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
Ls = [int(x) for x in sys.argv[3:]] or [6, 12, 18, 24, 30]
for L in Ls:
    rates = [1 << (i % 10) for i in range(L)]
    wn = dict(num_cycles=1, num_cycle_layers=L, dilation_rates=rates)
    cfg = O.Config(wavenet=wn)
    w = O.make_weights(cfg, seed=1234, peaked=True)
    eng = pkg.Engine(pkg.EngineConfig(wavenet=wn), device=0, max_batch=B)
    eng.set_weights(w)
    ze = O.synthetic_z_e(cfg, w, B, T // 64, seed=1235, kind="scaled")
    _, cond = eng.encode_condition(ze, [b % 4 for b in range(B)])
    eng.upload_condition(cond)
    eng.set_precision("tc")
    for rep in range(2):
        eng.generate_resident(B, T // 64, T, mode="greedy")
        ms = eng.last_kernel_ms
    print("L=%2d B=%d T=%d: %s %.2f us/step (%.1f MB weight stream per cluster)" % (L, B, T, eng.last_kernel_name, ms * 1e3 / T, L * 2.77), flush=True)
    eng.close()
