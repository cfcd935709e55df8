"""Generates the committed fixtures in tests/golden/.

Run in the BUILD container (it reads /root/reference/results/*.wav for the mu-law grid pin;
everything else comes from the NumPy oracle with seeded synthetic weights).  The GPU box has
no /root/reference: tests only read the .npz files written here.

    python tests/golden/make_golden.py [--skip-long]
"""
import os
import sys
import glob

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

SMALL_WAVENET = dict(num_cycles=2, num_cycle_layers=3, dilation_rates=[1, 2, 4, 1, 2, 4])


def wav_grid():
    """Weak known-answer pin: the shipped WAVs are float32 and every sample is a
    mu_law_decode_np(k) value."""
    from scipy.io import wavfile
    vals, names, rates, lens = [], [], [], []
    for p in sorted(glob.glob("/root/reference/results/VCTK/p225_001/*.wav")):
        sr, x = wavfile.read(p)
        assert x.dtype == np.float32
        vals.append(np.unique(x))
        names.append(os.path.basename(p))
        rates.append(sr)
        lens.append(x.shape[0])
    allv = np.unique(np.concatenate(vals))
    np.savez_compressed(os.path.join(HERE, "wav_grid.npz"), values=allv, names=np.array(names),
                        rates=np.array(rates), lengths=np.array(lens))
    print("wav_grid: %d distinct values from %d files" % (allv.size, len(names)))


def small():
    cfg = O.Config(wavenet=SMALL_WAVENET)
    w = O.make_weights(cfg, seed=1234)
    B, T, F = 3, 256, 4
    x = O.synthetic_audio(B, T, seed=1237)
    ze = O.synthetic_z_e(cfg, w, B, F, seed=1235, kind="scaled")
    idx, cond = O.encode_condition(ze, [0, 1, 2], w)
    logits_conv, labels = O.wavenet_teacher_forced(cfg, w, x[:, :, None], cond)
    _, _, logits_fast = O.generate(cfg, w, cond, T, mode="greedy", teacher=x, return_logits=True)
    _, gidx, gmar = O.generate(cfg, w, cond, T, mode="greedy", return_margins=True)
    u = np.random.default_rng(1236).random((T, B))
    _, sidx, smar = O.generate(cfg, w, cond, T, mode="sample", uniforms=u, return_margins=True)
    np.savez_compressed(os.path.join(HERE, "small.npz"), vq_idx=idx.astype(np.int16),
                        greedy_margin=gmar, sample_margin=smar,
                        logits_conv=logits_conv.reshape(B, T, -1)[:, ::16].astype(np.float32),
                        logits_fast=logits_fast[:, ::16].astype(np.float32),
                        labels=labels.astype(np.int16), greedy_idx=gidx.astype(np.int16),
                        sample_idx=sidx.astype(np.int16))
    print("small: conv-vs-fast max diff", np.abs(logits_conv.reshape(B, T, -1) - logits_fast).max())


def vq():
    cfg = O.Config()
    w = O.make_weights(cfg, seed=1234)
    out = {}
    for kind in ("normal", "near_code", "scaled"):
        ze = O.synthetic_z_e(cfg, w, 64, 104, seed=1235, kind=kind)
        idx, _, _ = O.vq_discretise(ze, w["embedding/embedding"])
        out["idx_" + kind] = idx.astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "vq_cfg2.npz"), **out)
    print("vq: wrote", list(out))


def full(long_steps=4096):
    """Full default configuration, peaked weight set: greedy 4096 samples, sample
    mode with rng(1236) uniforms, teacher-forced logits at a few steps."""
    cfg = O.Config()
    w = O.make_weights(cfg, seed=1234, peaked=True)
    B, T = 4, long_steps
    F = T // 64
    ze = O.synthetic_z_e(cfg, w, B, F, seed=1235, kind="scaled")
    vq_idx, cond = O.encode_condition(ze, [b % 4 for b in range(B)], w)
    _, gidx, gmar = O.generate(cfg, w, cond, T, mode="greedy", return_margins=True)
    Ts = 1024
    u = np.random.default_rng(1236).random((Ts, B))
    _, sidx, smar = O.generate(cfg, w, cond[:, :Ts // 64], Ts, mode="sample", uniforms=u, return_margins=True)
    Tt = 512
    x = O.synthetic_audio(B, Tt, seed=1237)
    _, _, lg = O.generate(cfg, w, cond[:, :Tt // 64], Tt, mode="greedy", teacher=x, return_logits=True)
    np.savez_compressed(os.path.join(HERE, "full.npz"), vq_idx=vq_idx.astype(np.int16),
                        greedy_margin=gmar.astype(np.float16), sample_margin=smar,
                        greedy_idx=gidx.astype(np.int16), sample_idx=sidx.astype(np.int16),
                        teacher_logits=lg[:, ::32].astype(np.float32))
    print("full: done")


if __name__ == "__main__":
    wav_grid()
    small()
    vq()
    if "--skip-long" not in sys.argv:
        full()
