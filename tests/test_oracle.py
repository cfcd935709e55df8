"""CPU tests of the oracle itself (no GPU): the pins listed in oracle/oracle.py's header."""
import os

import numpy as np
import pytest

from oracle import oracle as O

SMALL_WAVENET = dict(num_cycles=2, num_cycle_layers=3, dilation_rates=[1, 2, 4, 1, 2, 4])


def test_wav_grid_pin(golden_dir):
    """Every sample of the reference's five shipped WAVs is a mu_law_decode_np(k) value
    (<= 1 ulp) -- the only reference-produced numbers available for this path."""
    g = np.load(os.path.join(golden_dir, "wav_grid.npz"))
    lut = O.decode_lut(256)
    vals = g["values"]
    nearest = np.abs(vals[:, None] - lut[None, :]).min(axis=1)
    assert nearest.max() <= 6e-8
    assert list(g["rates"]) == [16000] * 5
    assert list(g["lengths"]) == [32768] * 5


def test_mu_law_roundtrip_and_luts():
    dec = O.decode_lut()
    enc = O.encode_lut()
    assert dec.dtype == np.float32 and dec.shape == (257,)
    assert abs(float(dec[256]) - 1.0446261) < 1e-6
    assert enc[256] == np.float32(1.0)                    # clipped
    k = np.arange(256)
    assert np.array_equal(O.mu_law_encode(dec[:256], to_int=True), k)
    assert np.abs(enc[:256] - (2 * k / 255.0 - 1)).max() < 1e-6


def test_product_luts_equal_oracle_luts():
    """host-side tables of the product path are bit-identical to the oracle's restatement"""
    from vqvae_wavenet_b200 import mu_law_ops
    assert np.array_equal(mu_law_ops.decode_lut(), O.decode_lut())
    assert np.array_equal(mu_law_ops.encode_lut(), O.encode_lut())


def test_sample_semantics():
    pdf = np.full((3, 256), 1.0 / 256, dtype=np.float32)
    idx = O.sample_indices(pdf, [0.0, 0.5, 0.99999999])
    assert idx[0] == 0
    assert idx[1] in (127, 128)
    pdf2 = np.zeros((1, 256), dtype=np.float32)
    pdf2[0, :10] = 0.0999999
    assert O.sample_indices(pdf2, [0.9999999])[0] == 256   # cdf ends below the draw -> index q
    with pytest.raises(NotImplementedError):
        O.decode(pdf, mode="beam")
    flat = np.zeros((1, 256), dtype=np.float32)
    flat[0, [7, 9]] = 0.5
    assert O.decode_indices(flat, "greedy")[0] == 7        # first maximum


def test_vq_direct_vs_expanded_and_ties():
    cfg = O.Config()
    w = O.make_weights(cfg)
    E = w["embedding/embedding"]
    z = O.synthetic_z_e(cfg, w, 8, 13, kind="near_code")
    i1, e_k, z_q = O.vq_discretise(z, E)
    i2 = O.vq_discretise_expanded(z, E)
    assert np.array_equal(i1, i2)
    assert np.array_equal(z_q, z + (e_k - z))
    # duplicated codebook rows: lowest index wins
    E2 = E.copy()
    E2[300] = E2[17]
    zz = E2[[17, 300]] + np.float32(1e-3)
    i3, _, _ = O.vq_discretise(zz, E2)
    assert list(i3) == [17, 17]


def test_golden_vq_reproducible(golden_dir):
    cfg = O.Config()
    w = O.make_weights(cfg)
    g = np.load(os.path.join(golden_dir, "vq_cfg2.npz"))
    z = O.synthetic_z_e(cfg, w, 64, 104, kind="scaled")
    idx, _, _ = O.vq_discretise(z, w["embedding/embedding"])
    assert np.array_equal(idx, g["idx_scaled"])


def test_fast_equals_conv_small(golden_dir):
    """queue form (wavenet.py:103-172) == padded dilated conv form (wavenet.py:24-100)"""
    cfg = O.Config(wavenet=SMALL_WAVENET)
    w = O.make_weights(cfg)
    B, T, F = 3, 256, 4
    x = O.synthetic_audio(B, T)
    ze = O.synthetic_z_e(cfg, w, B, F, kind="scaled")
    _, cond = O.encode_condition(ze, [0, 1, 2], w)
    lc, labels = O.wavenet_teacher_forced(cfg, w, x[:, :, None], cond)
    _, _, lf = O.generate(cfg, w, cond, T, mode="greedy", teacher=x, return_logits=True)
    assert np.abs(lc.reshape(B, T, -1) - lf).max() < 2e-5
    g = np.load(os.path.join(golden_dir, "small.npz"))
    assert np.allclose(lf[:, ::16], g["logits_fast"], atol=1e-5)
    assert np.array_equal(labels, g["labels"])


def test_conv_form_against_torch_witness():
    """independent library (torch conv1d) for the dilated causal conv + condition broadcast"""
    import torch
    import torch.nn.functional as Fn
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 40, 8)).astype(np.float32)
    k = rng.standard_normal((3, 8, 6)).astype(np.float32)
    b = rng.standard_normal(6).astype(np.float32)
    for d in (1, 2, 4):
        ours = O.conv1d_v2(x, k, b, d)
        xt = torch.from_numpy(x).permute(0, 2, 1)
        xt = Fn.pad(xt, (d * 2, 0))
        ref = Fn.conv1d(xt, torch.from_numpy(k).permute(2, 1, 0), torch.from_numpy(b), dilation=d)
        assert np.allclose(ours, ref.permute(0, 2, 1).numpy(), atol=1e-5)


def test_receptive_field():
    cfg = O.Config(wavenet=SMALL_WAVENET)
    assert cfg.receptive_field == 14 * 2 + 1 + 31
    assert O.Config().receptive_field == 6170             # wavenet.py:15-17 with the shipped JSON
    w = O.make_weights(cfg)
    B, T, F = 1, 128, 2
    x = O.synthetic_audio(B, T)
    ze = O.synthetic_z_e(cfg, w, B, F, kind="scaled")
    _, cond = O.encode_condition(ze, [0], w)
    base, _ = O.wavenet_teacher_forced(cfg, w, x[:, :, None], cond)
    x2 = x.copy()
    x2[0, 10] = 0.9
    pert, _ = O.wavenet_teacher_forced(cfg, w, x2[:, :, None], cond)
    diff = np.abs(base - pert).reshape(T, -1).max(-1)
    rf = cfg.receptive_field
    # input x[10] enters at step 11 (shift_right) and can influence steps 11 .. 10+rf
    assert diff[:11].max() == 0
    assert diff[11 + rf:].max() == 0
    assert diff[11:11 + rf].max() > 0


def test_speaker_none_is_row_zero():
    table = np.arange(12, dtype=np.float32).reshape(4, 3)
    onehot = np.zeros((2, 1, 4), dtype=np.float32)
    onehot[1, 0, 2] = 1
    h = O.speaker_rows(onehot, table)
    assert np.array_equal(h[0, 0], table[0]) and np.array_equal(h[1, 0], table[2])


def test_product_synthetic_weights_equal_oracle():
    """bench.py / smoke use the package's seeded weights; the oracle has its own copy: same numbers"""
    import vqvae_wavenet_b200 as pkg
    from vqvae_wavenet_b200 import synthetic
    ours = synthetic.make_weights(pkg.EngineConfig(wavenet=SMALL_WAVENET), peaked=True)
    theirs = O.make_weights(O.Config(wavenet=SMALL_WAVENET), peaked=True)
    assert list(ours) == list(theirs)
    for k in ours:
        assert np.array_equal(ours[k], theirs[k]), k
    cfg = O.Config()
    z = synthetic.synthetic_z_e(pkg.EngineConfig(), 3, 5)
    assert np.array_equal(z, O.synthetic_z_e(cfg, {"embedding/embedding": None}, 3, 5, kind="scaled"))


def test_encoder_magenta_against_torch_witness():
    """Encoder_Magenta restatement (Encoder/encoder.py:29-64) vs an independent float64 torch.nn.functional.conv1d
    evaluation of the same graph (causal left padding, dilation, stride-2 1x1 subsampling)"""
    import torch
    import torch.nn.functional as Fn
    from oracle import oracle as O
    cfg = O.Config()
    w = O.make_encoder_magenta_weights(cfg)
    assert [n for n, _ in O.encoder_magenta_specs(cfg)][:4] == ["encoder/preprocess/kernel", "encoder/preprocess/bias",
                                                                "encoder/cycle_1/layer_1/dilated/kernel",
                                                                "encoder/cycle_1/layer_1/dilated/bias"]
    B, T = 2, 1024
    x = O.synthetic_audio(B, T, seed=3)[:, :, None]
    z = O.encoder_magenta_forward(cfg, w, x)
    assert z.shape == (B, T // 64, 64)

    def tconv(net, name, dil=1, stride=1):
        K, b = w[name + "/kernel"], w[name + "/bias"]
        k = K.shape[0]
        xin = Fn.pad(torch.from_numpy(np.asarray(net, dtype=np.float64)).permute(0, 2, 1), (dil * (k - 1), 0))
        y = Fn.conv1d(xin, torch.from_numpy(K).permute(2, 1, 0).double(), torch.from_numpy(b).double(), stride=stride, dilation=dil)
        return y.permute(0, 2, 1).numpy()

    en = tconv(O.mu_law_encode(O.shift_right(x.astype(np.float32))), "encoder/preprocess")
    for i, dil in enumerate(O.MAGENTA_DILATIONS):
        sc = "encoder/cycle_1/layer_%d" % (i + 1)
        d = tconv(en, sc + "/dilated", stride=2)
        gated = np.tanh(tconv(d, sc + "/gate", dil=dil)) / (1.0 + np.exp(-tconv(d, sc + "/filter", dil=dil)))
        en = d + tconv(gated, sc + "/residual")
    zt = tconv(en, "encoder/postprocess")
    assert np.abs(zt - z).max() < 2e-5
    # causality of the restatement
    x2 = x.copy()
    x2[:, 64 * 10:] = 0
    assert np.array_equal(O.encoder_magenta_forward(cfg, w, x2)[:, :10], z[:, :10])


def test_mfcc_front_end_properties():
    """Encoder/encoder_ops.py:14-43 restated (tf.contrib.signal): shapes, the mel matrix's structure, the DCT's
    orthogonality up to its scale, a pure tone landing in the right mel band, pad_end framing"""
    W = O.linear_to_mel_weight_matrix()
    assert W.shape == (201, 80) and W.dtype == np.float32
    assert np.all(W >= 0) and np.all(W[0] == 0)                       # DC bin zeroed
    assert np.all(W.max(0) > 0)                                       # every band has support
    centers = (W * np.arange(201)[:, None]).sum(0) / W.sum(0)
    assert np.all(np.diff(centers) >= 0) and centers[-1] > 150       # bands ordered in frequency (the lowest share 40 Hz bins)
    assert W[:, 0].nonzero()[0].min() >= 1 and W[:, -1].nonzero()[0].max() <= 200
    D = O.dct2_matrix(80, 80).astype(np.float64)
    G = D.T @ D
    assert np.allclose(np.diag(G)[1:], 1.0, atol=1e-5) and np.allclose(G - np.diag(np.diag(G)), 0.0, atol=1e-5)
    assert np.allclose(G[0, 0], 2.0, atol=1e-5)                       # TF's unnormalised type-2 DCT: c = 0 is not rescaled
    win = O.hann_window_periodic(400)
    assert win[0] == 0 and abs(win[200] - 1.0) < 1e-7 and abs(win[1] - win[399]) < 1e-7
    t = np.arange(1600)
    x = (0.5 * np.sin(2 * np.pi * 1000.0 * t / 16000.0)).astype(np.float32)[None]
    m = O.mfcc(x)
    assert m.shape == (1, 10, 13) and m.dtype == np.float32 and np.isfinite(m).all()
    # a direct (float64 DFT) evaluation of one interior frame
    fr = x[0, 160:560].astype(np.float64) * win
    n = np.arange(400)
    mag = np.abs(np.array([np.sum(fr * np.exp(-2j * np.pi * k * n / 400)) for k in range(201)]))
    assert mag.argmax() == 25                                         # 1000 Hz / (16000 / 400)
    # (a pure tone leaves most bands at rounding-noise level, where log(. + 1e-6) amplifies float32 FFT noise: the
    # value check uses the broadband synthetic signal)
    xs = O.synthetic_audio(1, 1600, seed=9)
    ms = O.mfcc(xs)
    fr = xs[0, 160:560].astype(np.float64) * win
    mag = np.abs(np.array([np.sum(fr * np.exp(-2j * np.pi * k * n / 400)) for k in range(201)]))
    ref = np.log(mag @ W.astype(np.float64) + 1e-6) @ O.dct2_matrix().astype(np.float64)
    assert np.abs(ref - ms[0, 1]).max() < 2e-3
    # the last frames see zero padding (pad_end=True): frame 9 covers samples 1440..1839, 160 real ones
    assert O.mfcc(x[:, :1500]).shape == (1, 10, 13)


def test_encoder2019_structure():
    """Encoder/encoder.py:72-98: hop 320, ten convs in creation order, `relu + relu` doubles (quirk Q17)"""
    cfg = O.Config()
    w = O.make_encoder2019_weights(cfg)
    names = [n for n, _ in O.encoder2019_specs(cfg)]
    assert names[0] == "encoder/conv1d/kernel" and names[-1] == "encoder/conv1d_9/bias" and len(names) == 20
    assert w["encoder/conv1d/kernel"].shape == (3, 13, 768) and w["encoder/conv1d_2/kernel"].shape == (4, 768, 768)
    x = O.synthetic_audio(2, 1280, seed=3)[:, :, None]
    z = O.encoder2019_forward(cfg, w, x)
    assert z.shape == (2, 4, 64) and np.isfinite(z).all()
    # doubling the last block's kernel and bias == what `relu + relu` does to a plain conv: scale the final linear map instead
    w2 = dict(w)
    w2["encoder/conv1d_9/kernel"] = w["encoder/conv1d_9/kernel"] * np.float32(0.5)
    z2 = O.encoder2019_forward(cfg, w2, x)
    b = w["encoder/conv1d_9/bias"]
    assert np.allclose((z - b) * 0.5, z2 - b, atol=1e-4)
