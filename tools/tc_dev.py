"""Development check of the split-bf16 tensor-core kernel on a GPU box: accuracy against the float32 kernel and the
oracle on short runs, then timing of the three precisions at the benchmark shape.  Not part of the product or tests."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vqvae_wavenet_b200 as pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

SMALL = dict(num_cycles=2, num_cycle_layers=3, dilation_rates=[1, 2, 4, 1, 2, 4])


def accuracy(wavenet, B, T, peaked, label):
    cfg = O.Config(wavenet=wavenet)
    w = O.make_weights(cfg, seed=1234, peaked=peaked)
    eng = pkg.Engine(pkg.EngineConfig(wavenet=wavenet), device=0, max_batch=max(B, 16))
    eng.set_weights(w)
    ze = O.synthetic_z_e(cfg, w, B, max(T // 64, 1), seed=1235, kind="scaled")
    _, cond = eng.encode_condition(ze, [b % 4 for b in range(B)])
    x = O.synthetic_audio(B, T, seed=1237)
    out = {}
    for prec in ("fp32", "tc"):
        eng.set_precision(prec)
        t0 = time.time()
        out[prec] = eng.teacher_forced(x, cond)
        print("%s %s teacher_forced: kernel %s %.2f ms wall %.2fs" % (label, prec, eng.last_kernel_name, eng.last_kernel_ms, time.time() - t0), flush=True)
    scale = np.abs(out["fp32"]).max()
    d = np.abs(out["tc"] - out["fp32"])
    print("%s: max |tc - fp32| / max|logit| = %.3g (per step max: first %s ... worst step %d)" % (
        label, d.max() / scale, np.array2string(d.max(axis=(0, 2))[:6] / scale, precision=2), int(d.max(axis=(0, 2)).argmax())), flush=True)
    if not np.isfinite(out["tc"]).all():
        print("NON-FINITE logits in tc output", flush=True)
    Tg = min(T, 256)
    eng.set_precision("fp32")
    g0 = eng.generate(cond[:, :max(Tg // 64, 1)], Tg, mode="greedy")[1]
    eng.set_precision("tc")
    g1 = eng.generate(cond[:, :max(Tg // 64, 1)], Tg, mode="greedy")[1]
    print("%s: greedy %d steps equal fraction %.4f" % (label, Tg, (g0 == g1).mean()), flush=True)
    u = np.random.default_rng(3).random((Tg, B))
    eng.set_precision("fp32")
    s0 = eng.generate(cond[:, :max(Tg // 64, 1)], Tg, mode="sample", uniforms=u)[1]
    eng.set_precision("tc")
    s1 = eng.generate(cond[:, :max(Tg // 64, 1)], Tg, mode="sample", uniforms=u)[1]
    print("%s: sample %d steps equal fraction %.4f" % (label, Tg, (s0 == s1).mean()), flush=True)
    eng.close()
    return d.max() / scale


def timing(B, T, precisions=("tc", "fp32", "bf16")):
    cfg = O.Config()
    w = O.make_weights(cfg, seed=1234, peaked=True)
    eng = pkg.Engine(pkg.EngineConfig(), device=0, max_batch=B)
    eng.set_weights(w)
    ze = O.synthetic_z_e(cfg, w, B, T // 64, seed=1235, kind="scaled")
    _, cond = eng.encode_condition(ze, [b % 4 for b in range(B)])
    eng.upload_condition(cond)
    for prec in precisions:
        eng.set_precision(prec)
        for rep in range(2):
            eng.generate_resident(B, T // 64, T, mode="greedy")
            ms = eng.last_kernel_ms
        print("timing B=%d T=%d %s: %s %.2f ms = %.2f us/step, %.0f samples/s" % (
            B, T, prec, eng.last_kernel_name, ms, ms * 1e3 / T, B * T / (ms * 1e-3)), flush=True)
    eng.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "acc"):
        accuracy(SMALL, 3, 128, False, "small B=3")
        accuracy(SMALL, 23, 128, False, "small B=23")
        accuracy(None, 4, 256, True, "full B=4")
    if what in ("all", "time"):
        timing(64, 2048)
        timing(112, 2048, ("tc",))
        timing(16, 2048, ("tc",))
        timing(1, 2048, ("tc", "fp32"))
