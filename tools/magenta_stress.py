"""stress (development): long free-running launches of the 50-layer Magenta topology on the tensor-core kernel.
magenta_stress.py rounds T"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from vqvae_wavenet_b200 import magenta
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
B = 64
mw = O.make_magenta_fastgen_weights(peaked=True)
for i in range(n):
    gen = magenta.FastGenerationConfig(batch_size=B, precision="tc")
    gen.restore(mw)
    onehot = np.zeros((B, 109), np.float32)
    onehot[np.arange(B), np.arange(B) % 109] = 1
    gen.build(onehot)
    codes = np.random.default_rng(i).integers(0, 512, size=(B, T // 64))
    cond = gen.condition_from_codes(mw["embedding"][codes])
    t0 = time.time()
    audio, idx = gen.generate(cond, T, mode="sample", seed=i)
    print("magenta round %d ok %.2f s kernel %s %.2f ms = %.2f us/step, distinct %d" % (
        i, time.time() - t0, gen.engine.last_kernel_name, gen.engine.last_kernel_ms, gen.engine.last_kernel_ms * 1e3 / T, len(np.unique(idx))), flush=True)
    gen.close()
print("all rounds ok", flush=True)
