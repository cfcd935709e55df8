"""The oracle against golden vectors produced by the reference's OWN code (CPU, no GPU).

tests/golden/ref_*.npz were written by tests/golden/make_ref_golden.py, which executes the unmodified
reference files (model.py, Decoder/*, Decoder/WaveNet/*, mu_law_ops.py, utils.py, Encoder/encoder.py) on top
of a NumPy stand-in for the TensorFlow leaf operators (tests/golden/tf_shim.py).  These tests pin
oracle/oracle.py to those outputs: bit-exact for indices / queue-form arithmetic, 1e-5 for the conv form
(the shim evaluates conv2d with torch, the oracle with NumPy matmuls).

Long runs are sampled so the CPU suite stays within minutes; VQWN_LONG=1 checks every step.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

SMALL_WAVENET = dict(num_cycles=2, num_cycle_layers=3, dilation_rates=[1, 2, 4, 1, 2, 4])
LONG = os.environ.get("VQWN_LONG", "0") == "1"


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _uniforms(seed, T, B):
    """what utils.py:22 consumed: np.random.rand(B) per step from the seeded global RNG"""
    rs = np.random.RandomState(int(seed))
    return np.stack([rs.rand(B) for _ in range(T)])


def test_variable_names_and_shapes(golden_dir):
    """the weight list the oracle (and vqwn_set_tensor) uses == what the reference's graph creates"""
    g = _load(golden_dir, "ref_vars.npz")
    ref = {n: tuple(int(x) for x in s.split(",")) for n, s in zip(g["variable_names"], g["variable_shapes"])}
    mine = {n: tuple(s) for n, s in O.tensor_specs(O.Config())}
    assert ref == mine
    assert int(g["receptive_field"]) == O.Config().receptive_field == 6170


@pytest.mark.parametrize("kind", ["normal", "near_code", "scaled"])
def test_vq_matches_reference(kind, golden_dir):
    cfg = O.Config()
    w = O.make_weights(cfg, seed=1234)
    g = _load(golden_dir, "ref_vq.npz")
    ze = O.synthetic_z_e(cfg, w, 64, 104, seed=1235, kind=kind)
    idx, _, zq = O.vq_discretise(ze, w["embedding/embedding"])
    assert np.array_equal(idx, g["vq_idx_" + kind])
    assert np.array_equal(zq.astype(np.float64).sum((1, 2)), g["vq_zq_sum_" + kind])
    # the older oracle-made fixture holds the same indices
    assert np.array_equal(_load(golden_dir, "vq_cfg2.npz")["idx_" + kind], g["vq_idx_" + kind])


def test_condition_and_ties_match_reference(golden_dir):
    cfg = O.Config()
    w = O.make_weights(cfg, seed=1234)
    g = _load(golden_dir, "ref_vq.npz")
    ze = O.synthetic_z_e(cfg, w, 64, 104, seed=1235, kind="scaled")[:8]
    spk = [0, 0, 1, 2, 3, 108, 5, 7]                       # the reference's 'None' row is index 0
    _, cond = O.encode_condition(ze, spk, w)
    assert np.array_equal(cond[:, ::8], g["vq_encoding_first8"])
    E = w["embedding/embedding"].copy()
    for dup, src in g["vq_tie_rows"]:
        E[dup] = E[src]
    idx, _, _ = O.vq_discretise(g["vq_tie_z"], E)
    assert np.array_equal(idx, g["vq_tie_idx"])
    assert list(idx[0][:4]) == [7, 7, 12, 12]               # duplicated rows: lowest index


def test_codec_and_decode_kats(golden_dir):
    g = _load(golden_dir, "ref_kat.npz")
    assert np.array_equal(O.decode_lut(), g["kat_decode_lut"])
    assert np.array_equal(O.mu_law_encode(g["kat_encode_x"]), g["kat_encode_float"])
    assert np.array_equal(O.mu_law_encode(g["kat_encode_x"], to_int=True), g["kat_encode_int"])
    pdf = g["kat_pdf"]
    assert np.array_equal(O.decode(pdf, mode="greedy"), g["kat_greedy"])
    for seed, want in zip(g["kat_sample_seeds"], g["kat_sample"]):
        u = np.random.RandomState(int(seed)).rand(pdf.shape[0])
        assert np.array_equal(O.decode(pdf, mode="sample", uniforms=u), want)
    hi = O.decode(pdf, mode="sample", uniforms=np.full(pdf.shape[0], 0.99999999))
    assert np.array_equal(hi, g["kat_sample_u_high"])
    assert hi[3] == O.decode_lut()[256]                      # index 256 overflow


def _run_config(golden_dir, tag, wav, peaked, steps_greedy, steps_sample, steps_teacher):
    g = _load(golden_dir, "ref_%s.npz" % tag)
    cfg = O.Config(wavenet=wav)
    w = O.make_weights(cfg, seed=1234, peaked=peaked)
    B, T = g[tag + "_greedy_idx"].shape
    spk = [max(int(s), 0) for s in g[tag + "_speakers"]]
    ze = O.synthetic_z_e(cfg, w, B, T // 64, seed=1235, kind="scaled")
    idx, cond = O.encode_condition(ze, spk, w)
    assert np.array_equal(idx, g[tag + "_vq_idx"])
    stride = int(g[tag + "_logit_stride"])
    Tt = min(steps_teacher, g[tag + "_teacher_logits"].shape[1] * stride)
    x = O.synthetic_audio(B, g[tag + "_teacher_logits"].shape[1] * stride, seed=1237)
    _, _, lg = O.generate(cfg, w, cond[:, :max(Tt // 64, 1)], Tt, mode="greedy", teacher=x, return_logits=True)
    want = g[tag + "_teacher_logits"][:, : (Tt + stride - 1) // stride]
    assert np.array_equal(lg[:, ::stride], want), "queue-form logits differ from the reference's"
    assert np.array_equal(O.softmax(lg[:, ::stride]), g[tag + "_teacher_probs"][:, : want.shape[1]])
    Tg = min(steps_greedy, T)
    _, gidx = O.generate(cfg, w, cond[:, :max(Tg // 64, 1)], Tg, mode="greedy")
    assert np.array_equal(gidx, g[tag + "_greedy_idx"][:, :Tg])
    Ts = min(steps_sample, g[tag + "_sample_idx"].shape[1])
    u = _uniforms(g[tag + "_sample_seed"], g[tag + "_sample_idx"].shape[1], B)
    _, sidx = O.generate(cfg, w, cond[:, :max(Ts // 64, 1)], Ts, mode="sample", uniforms=u[:Ts])
    assert np.array_equal(sidx, g[tag + "_sample_idx"][:, :Ts])


def test_small_config_matches_reference(golden_dir):
    _run_config(golden_dir, "small", SMALL_WAVENET, False, 256, 256, 256)


def test_full_config_matches_reference(golden_dir):
    n = 10 ** 9 if LONG else 128
    _run_config(golden_dir, "full", None, True, n, n, 64 if not LONG else n)


def test_small_conv_form_matches_reference(golden_dir):
    g = _load(golden_dir, "ref_small.npz")
    cfg = O.Config(wavenet=SMALL_WAVENET)
    w = O.make_weights(cfg, seed=1234)
    B, T = 3, 256
    ze = O.synthetic_z_e(cfg, w, B, T // 64, seed=1235, kind="scaled")
    _, cond = O.encode_condition(ze, [0, 1, 2], w)
    x = O.synthetic_audio(B, T, seed=1237)
    lg, labels = O.wavenet_teacher_forced(cfg, w, x[:, :, None], cond)
    lg = lg.reshape(B, T, -1)
    s = int(g["small_conv_stride"])
    assert np.array_equal(labels, g["small_conv_labels"])
    assert np.abs(lg[:, s - 1::s] - g["small_conv_logits"]).max() < 1e-5


def test_cfg5_conv_form_matches_reference(golden_dir):
    """BASELINE config 5 (teacher-forced, length 6656): one of the 8 streams through the oracle's conv form
    (all 8 with VQWN_LONG=1)"""
    g = _load(golden_dir, "ref_cfg5.npz")
    cfg = O.Config()
    w = O.make_weights(cfg, seed=1234)
    B, T = 8, 6656
    ze = O.synthetic_z_e(cfg, w, B, T // 64, seed=1235, kind="scaled")
    _, cond = O.encode_condition(ze, [b % 4 for b in range(B)], w)
    x = O.synthetic_audio(B, T, seed=1237)
    sel = list(range(B)) if LONG else [5]
    lg, labels = O.wavenet_teacher_forced(cfg, w, x[sel][:, :, None], cond[sel])
    lg = lg.reshape(len(sel), T, -1)
    s = int(g["cfg5_conv_stride"])
    scale = np.abs(g["cfg5_conv_logits"]).max()
    assert np.abs(lg[:, s - 1::s] - g["cfg5_conv_logits"][sel]).max() < 2e-5 * max(scale, 1.0)
    assert np.abs(lg[:, -8:] - g["cfg5_conv_last_logits"][sel]).max() < 2e-5 * max(scale, 1.0)
    assert np.array_equal(labels.reshape(len(sel), T), g["cfg5_conv_labels"].reshape(B, T)[sel])


def test_encoders_match_reference(golden_dir):
    g = _load(golden_dir, "ref_encoders.npz")
    cfg = O.Config()
    x = O.synthetic_audio(2, 2048, seed=1237)[:, :, None]
    z = O.encoder64_forward(cfg, O.make_encoder64_weights(cfg), x)
    assert np.abs(z - g["enc64_z_e"]).max() < 1e-5
    z = O.encoder_magenta_forward(cfg, O.make_encoder_magenta_weights(cfg), x)
    assert np.abs(z - g["encmag_z_e"]).max() < 2e-5


def test_encoder2019_matches_reference(golden_dir):
    """Encoder_2019 (Encoder/encoder.py:66-98 run by make_ref_golden.py; its tf.contrib.signal leaf operators are
    tf_shim's restatement of TensorFlow's published MFCC pipeline): oracle == reference output"""
    g = _load(golden_dir, "ref_enc2019.npz")
    cfg = O.Config()
    x = O.synthetic_audio(2, 7680, seed=1237)[:, :, None]
    z = O.encoder2019_forward(cfg, O.make_encoder2019_weights(cfg), x)
    assert z.shape == g["enc2019_z_e"].shape == (2, 24, 64)
    assert np.abs(z - g["enc2019_z_e"]).max() < 1e-4


def test_cfg5_end_to_end_fixture_front_half(golden_dir):
    """ref_cfg5_e2e.npz (the reference's whole graph per encoder variant): the oracle's encoders and VQ reproduce its
    z_e and code indices; the decoder half is checked on the GPU (tests/test_gpu_parity.py) and, for the conv form
    itself, by the cfg5 test above"""
    g = _load(golden_dir, "ref_cfg5_e2e.npz")
    cfg = O.Config()
    E = O.make_weights(cfg, seed=1234)["embedding/embedding"]
    for tag, mk, fwd, tol in (("e2e64", O.make_encoder64_weights, O.encoder64_forward, 2e-5),
                              ("e2emag", O.make_encoder_magenta_weights, O.encoder_magenta_forward, 5e-5),
                              ("e2e2019", O.make_encoder2019_weights, O.encoder2019_forward, 2e-4)):
        T = int(g[tag + "_T"])
        x = O.synthetic_audio(8, T, seed=1237)[:2, :, None]
        z = fwd(cfg, mk(cfg), x)
        assert np.abs(z - g[tag + "_z_e"][:2]).max() < tol * max(1.0, np.abs(z).max()), tag
        idx = O.vq_discretise(g[tag + "_z_e"], E)[0]
        assert np.array_equal(idx, g[tag + "_idx"].astype(np.int64)), tag


def test_magenta_fastgen_oracle_matches_reference(golden_dir):
    """Magenta/ fast generation (Magenta/config.py:18-138 + masked.py:133-174 run unmodified by make_ref_golden.py, full size:
    50 layers): the oracle's restatement gives the same logits bit for bit and the same greedy / seeded-sample sequences"""
    g = _load(golden_dir, "ref_magenta.npz")
    mw = O.make_magenta_fastgen_weights(peaked=True)
    enc = mw["embedding"][g["magenta_codes"].astype(np.int64)]
    spk = [int(v) for v in g["magenta_speakers"]]
    B, T = g["magenta_greedy_idx"].shape
    n = T if LONG else 48
    x = O.synthetic_audio(B, T, seed=1237)
    _, _, lg, _ = O.magenta_generate(mw, enc, spk, n, mode="greedy", teacher=x[:, :n]) if n == T else \
        O.magenta_generate(mw, np.repeat(enc, 64, axis=1)[:, :n], spk, n, mode="greedy", teacher=x[:, :n])
    assert np.array_equal(lg, g["magenta_teacher_logits"][:, :n])
    _, gi, _, _ = O.magenta_generate(mw, np.repeat(enc, 64, axis=1)[:, :n], spk, n, mode="greedy")
    assert np.array_equal(gi, g["magenta_greedy_idx"][:, :n])
    u = _uniforms(g["magenta_sample_seed"], T, B)
    _, si, _, _ = O.magenta_generate(mw, np.repeat(enc, 64, axis=1)[:, :n], spk, n, mode="sample", uniforms=u[:n])
    assert np.array_equal(si, g["magenta_sample_idx"][:, :n])


def test_magenta_weight_mapping_onto_default_topology(golden_dir):
    """the product's re-arrangement of a Magenta checkpoint (vq-vae-wavenet_b200/magenta.py: zero oldest tap, swapped gate
    halves, [cond_map | gc] stacked as one local-condition kernel, biases folded) evaluated by the DEFAULT generator's
    restatement (wavenet.py:103-172) reproduces the Magenta reference's logits: the mapping is exact up to float32
    summation order"""
    from vqvae_wavenet_b200 import magenta
    g = _load(golden_dir, "ref_magenta.npz")
    mw = O.make_magenta_fastgen_weights(peaked=True)
    assert {k: v.shape for k, v in mw.items()} == magenta.variable_shapes()
    w = magenta.convert_weights(mw)
    cfg = O.Config(wavenet=magenta.wavenet_parameters())
    assert cfg.receptive_field == sum(cfg.dilations) * 2 + 1 + 31
    spk = g["magenta_speakers"].astype(np.int64)
    enc = mw["embedding"][g["magenta_codes"].astype(np.int64)]
    B = enc.shape[0]
    n = 32
    cond = np.concatenate([np.repeat(enc, 64, axis=1)[:, :n], np.repeat(mw["speaker_emb"][spk][:, None], n, axis=1)], -1)
    x = O.synthetic_audio(B, g["magenta_greedy_idx"].shape[1], seed=1237)[:, :n]
    _, _, lg = O.generate(cfg, w, cond, n, mode="greedy", teacher=x, return_logits=True)
    want = g["magenta_teacher_logits"][:, :n]
    assert np.abs(lg - want).max() <= 2e-5 * np.abs(want).max()
    with pytest.raises(KeyError):
        magenta.convert_weights({k: v for k, v in mw.items() if k != "gc_7/bias"})
