"""Golden vectors produced by the reference's OWN code (tests/golden/ref_*.npz).

Run in the BUILD container only (it imports the unmodified reference from /root/reference):

    python tests/golden/make_ref_golden.py [--quick]

TensorFlow 1.x is not installable here, so `tf_shim.py` (a deferred-execution NumPy stand-in for
the handful of TF leaf operators the path uses) is installed as `tensorflow`; everything above the
leaf operators - graph construction, variable names and shapes, queue chaining, tap order, slicing,
conditioning, the mu-law codec, the host sampling loop - is executed from the reference's files:

    model.py:7-33,36-74,85-87,133-142      VQVAE.__init__ / _build / _discretise / _build_decoder(_generator)
    Decoder/decoder.py:12-62               WavenetDecoder.build / build_generator
    Decoder/decoder_ops.py:39-43           concat
    Decoder/WaveNet/wavenet.py:10-172      Wavenet.__init__ / build / build_generator
    Decoder/WaveNet/wavenet_ops.py:9-14,59-138,147-267
    mu_law_ops.py:5-31, utils.py:13-46     codec, sample / decode
    Encoder/encoder.py:8-64                Encoder_64, Encoder_Magenta
    generate.py:103-113                    the host loop is re-typed here (it is module-level script code that cannot be
                                           imported); it calls the reference's decode() and session handles exactly as written

The weights and inputs are the seeded synthetic ones of SURVEY 8d (oracle.make_weights & co. only
GENERATE random numbers here; no oracle arithmetic ends up in a fixture except the `*_margin`
arrays, which are an analysis of the reference's probabilities used by tests to recognise near-ties).
Each section also prints how far the oracle is from the reference output.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

tf_shim.install()
sys.path.insert(0, REF)
import tensorflow as tf  # noqa: E402  (the shim)
import model as ref_model  # noqa: E402
import utils as ref_utils  # noqa: E402
import mu_law_ops as ref_mu  # noqa: E402
from Decoder.decoder import WavenetDecoder  # noqa: E402
from Encoder.encoder import Encoder_64, Encoder_Magenta, Encoder_2019  # noqa: E402

sys.path.insert(1, ROOT)
from oracle import oracle as O  # noqa: E402

SMALL_WAVENET = dict(num_cycles=2, num_cycle_layers=3, dilation_rates=[1, 2, 4, 1, 2, 4])
QUICK = "--quick" in sys.argv


def wavenet_json(overrides=None):
    """the reference's wavenet_parameters.json, optionally with a shorter dilation stack"""
    with open(os.path.join(REF, "wavenet_parameters.json")) as f:
        args = json.load(f)
    args["verbose"] = False
    if overrides:
        args.update(overrides)
    fd, path = tempfile.mkstemp(suffix=".json")
    with os.fdopen(fd, "w") as f:
        json.dump(args, f)
    return path


class ConstantEncoder:
    """stands in for Encoder_*: build(x) returns a given z_e (the VQ input of BASELINE config 2)"""

    def __init__(self, z_e):
        self.z_e = z_e

    def build(self, x):
        return tf.constant(self.z_e)


def onehot(speaker_idx, n=109):
    """generate.py:46-61: [B,1,N] one-hot; index None -> all-zero row"""
    sp = np.zeros((len(speaker_idx), 1, n), dtype=np.float32)
    for i, s in enumerate(speaker_idx):
        if s is not None:
            sp[i, 0, s] = 1
    return sp


def build_model(weights, wjson, encoder, x, speaker_idx, num_speakers=109):
    """generate.py:63-86"""
    tf.reset_default_graph()
    tf.set_variable_values(weights)
    args = {"x": tf.constant(x), "speaker": tf.constant(onehot(speaker_idx, num_speakers)), "encoder": encoder,
            "decoder": WavenetDecoder(wjson), "k": 512, "beta": 0.25, "verbose": False, "use_vq": True,
            "speaker_embedding": 64, "num_speakers": num_speakers}
    return ref_model.VQVAE(args)


def host_loop(sess, wavenet, encoding, length, mode, teacher=None, want_logits=False, seed=None):
    """generate.py:103-113, with the optional teacher-forcing used by the logit-parity fixtures.
    Sample mode draws np.random.rand(B) per step from the global RNG exactly as utils.py:22 does; the
    fixture records the seed so a test can regenerate the same uniforms."""
    B = encoding.shape[0]
    audio = np.zeros([B, 1], dtype=np.float32)
    to_write = np.zeros([B, length], dtype=np.float32)
    probs_all = np.zeros([B, length, 256], dtype=np.float32)
    logits_all = np.zeros([B, length, 256], dtype=np.float32) if want_logits else None
    logits_node = wavenet.predictions.inputs[0]          # the tensor tf.nn.softmax was applied to (wavenet.py:171)
    sess.run(wavenet.init_ops)
    if seed is not None:
        np.random.seed(seed)
    ratio = length // encoding.shape[1]
    for i in range(length):
        fetch = [wavenet.predictions, wavenet.push_ops] + ([logits_node] if want_logits else [])
        out = sess.run(fetch, {wavenet.input_t: audio, wavenet.local_condition_t: encoding[:, i // ratio]})
        probs = out[0]
        decoded = ref_utils.decode(probs, mode=mode, quantization_channels=wavenet.args["quantization_channels"])
        to_write[:, i] = decoded
        probs_all[:, i] = probs
        if want_logits:
            logits_all[:, i] = out[2]
        audio = np.expand_dims(decoded, -1) if teacher is None else np.asarray(teacher[:, i:i + 1], dtype=np.float32)
    return to_write, probs_all, logits_all


def audio_to_index(audio):
    """invert mu_law_decode_np on its own grid (exact: the grid is strictly increasing)"""
    lut = ref_mu.mu_law_decode_np(np.arange(257))
    idx = np.searchsorted(lut, audio)
    idx = np.clip(idx, 0, 256)
    assert np.array_equal(lut[idx], audio)
    return idx.astype(np.int16)


def section_variables(out):
    """names + shapes the reference creates for the generator graph (SURVEY 8a weight list)"""
    cfg = O.Config()
    w = O.make_weights(cfg)
    m = build_model(w, wavenet_json(), ConstantEncoder(np.zeros((1, 2, 64), np.float32)), np.zeros((1, 128, 1), np.float32), [0])
    m.build_generator()
    created = tf_shim.created_variables()
    out["variable_names"] = np.array([n for n, _ in created])
    out["variable_shapes"] = np.array([",".join(map(str, s)) for _, s in created])
    assert created == [(n, tuple(s)) for n, s in O.tensor_specs(cfg)
                       ] or sorted(created) == sorted((n, tuple(s)) for n, s in O.tensor_specs(cfg)), "oracle.tensor_specs differs"
    out["receptive_field"] = np.int64(m.decoder.wavenet.receptive_field)
    print("variables: %d tensors, receptive field %d" % (len(created), m.decoder.wavenet.receptive_field))


def section_vq(out):
    """model.py:57-74 + :19-27 + decoder_ops.py:39-43 on BASELINE config 2 (64 x 104 vectors), three
    distributions, plus an exact-tie codebook"""
    cfg_small = O.Config(wavenet=SMALL_WAVENET)
    w = O.make_weights(cfg_small, seed=1234)            # same codebook / speaker table as the full set (drawn first)
    w_full = O.make_weights(O.Config(), seed=1234)
    assert np.array_equal(w["embedding/embedding"], w_full["embedding/embedding"])
    wj = wavenet_json(SMALL_WAVENET)
    sess = tf.Session()
    spk = [None, 0, 1, 2, 3, 108, 5, 7]
    for kind in ("normal", "near_code", "scaled"):
        ze = O.synthetic_z_e(cfg_small, w, 64, 104, seed=1235, kind=kind)
        idx = np.zeros((64, 104), np.int64)
        zq_sum = np.zeros((64,), np.float64)
        enc0 = None
        for b0 in range(0, 64, 8):
            m = build_model(w, wj, ConstantEncoder(ze[b0:b0 + 8]), np.zeros((8, 104 * 64, 1), np.float32), spk)
            m.build_generator()
            q, zq, enc = sess.run([m.q_z_x, m.z_q, m.encoding])
            assert q.dtype == np.int64 and enc.shape == (8, 104, 128)
            idx[b0:b0 + 8] = q
            zq_sum[b0:b0 + 8] = zq.astype(np.float64).sum((1, 2))
            assert np.array_equal(enc[..., :64], zq)
            if b0 == 0:
                enc0 = enc
        oi, _, ozq = O.vq_discretise(ze, w["embedding/embedding"])
        print("vq %-9s: oracle index mismatches %d, z_q sum diff %.3g" % (kind, int((oi != idx).sum()),
                                                                          np.abs(ozq.astype(np.float64).sum((1, 2)) - zq_sum).max()))
        out["vq_idx_" + kind] = idx.astype(np.int16)
        out["vq_zq_sum_" + kind] = zq_sum
        if kind == "scaled":
            out["vq_encoding_first8"] = enc0[:, ::8].astype(np.float32)      # incl. the speaker half, rows None,0,1,2,3,108,5,7
    # exact ties: codebook rows 7 and 300 identical, 12 and 13 identical; z on those rows and on midpoints
    w2 = dict(w)
    E = w["embedding/embedding"].copy()
    E[300] = E[7]
    E[13] = E[12]
    w2["embedding/embedding"] = E
    z = np.stack([E[7], E[300], E[12], E[13], (E[1] + E[2]) * np.float32(0.5), (E[500] + E[40]) * np.float32(0.5),
                  E[511], np.zeros(64, np.float32)])[None].astype(np.float32)
    m = build_model(w2, wj, ConstantEncoder(z), np.zeros((1, 8 * 64, 1), np.float32), [None])
    m.build_generator()
    out["vq_tie_z"] = z
    out["vq_tie_rows"] = np.array([[300, 7], [13, 12]])
    out["vq_tie_idx"] = sess.run(m.q_z_x).astype(np.int16)
    print("vq ties ->", out["vq_tie_idx"].tolist())


def run_config(tag, out, wav_over, B, T, T_sample, T_teacher, peaked, logit_stride, spk):
    cfg = O.Config(wavenet=wav_over)
    w = O.make_weights(cfg, seed=1234, peaked=peaked)
    F = T // 64
    ze = O.synthetic_z_e(cfg, w, B, F, seed=1235, kind="scaled")
    wj = wavenet_json(wav_over)
    sess = tf.Session()

    def fresh():
        """a new graph per run: init_ops may run only once per graph (a second run would block on full queues)"""
        mm = build_model(w, wj, ConstantEncoder(ze), np.zeros((B, T, 1), np.float32), spk)
        mm.build_generator()
        return mm, mm.decoder.wavenet
    m, wn = fresh()
    encoding = sess.run(m.encoding)                                    # generate.py:92
    out[tag + "_vq_idx"] = sess.run(m.q_z_x).astype(np.int16)
    out[tag + "_speakers"] = np.array([-1 if s is None else s for s in spk])
    o_idx, o_cond = O.encode_condition(ze, [0 if s is None else s for s in spk], w)
    print("%s: encoding oracle diff %.3g" % (tag, np.abs(o_cond - encoding).max()))

    # teacher-forced step logits (the per-step logit parity target)
    x = O.synthetic_audio(B, T_teacher, seed=1237)
    t0 = time.time()
    _, probs, logits = host_loop(sess, wn, encoding[:, :T_teacher // 64], T_teacher, "greedy", teacher=x, want_logits=True)
    _, _, o_logits = O.generate(cfg, w, o_cond[:, :T_teacher // 64], T_teacher, mode="greedy", teacher=x, return_logits=True)
    print("%s: teacher-forced %d steps in %.1fs; oracle logit diff %.3g (max |logit| %.3g)"
          % (tag, T_teacher, time.time() - t0, np.abs(o_logits - logits).max(), np.abs(logits).max()))
    out[tag + "_teacher_logits"] = logits[:, ::logit_stride].astype(np.float32)
    out[tag + "_teacher_probs"] = probs[:, ::logit_stride].astype(np.float32)
    out[tag + "_logit_stride"] = np.int64(logit_stride)

    # greedy free-running sequence
    t0 = time.time()
    m, wn = fresh()
    audio, probs, _ = host_loop(sess, wn, encoding, T, "greedy")
    gidx = audio_to_index(audio)
    out[tag + "_greedy_idx"] = gidx
    out[tag + "_greedy_margin"] = np.stack([O.draw_margin(probs[:, i], "greedy") for i in range(T)], 1).astype(np.float16)
    _, o_gidx = O.generate(cfg, w, o_cond, T, mode="greedy")
    print("%s: greedy %d steps x %d streams in %.1fs; oracle mismatching draws %d"
          % (tag, T, B, time.time() - t0, int((o_gidx != gidx).sum())))

    # sample mode, global NumPy RNG seeded (utils.py:22 draws np.random.rand(B) per step)
    seed = 1236
    m, wn = fresh()
    audio, probs, _ = host_loop(sess, wn, encoding[:, :T_sample // 64], T_sample, "sample", seed=seed)
    sidx = audio_to_index(audio)
    np.random.seed(seed)
    u = np.stack([np.random.rand(B) for _ in range(T_sample)])        # what the loop consumed
    out[tag + "_sample_idx"] = sidx
    out[tag + "_sample_seed"] = np.int64(seed)
    out[tag + "_sample_margin"] = np.stack([O.draw_margin(probs[:, i], "sample", u[i]) for i in range(T_sample)], 1).astype(np.float32)
    _, o_sidx = O.generate(cfg, w, o_cond[:, :T_sample // 64], T_sample, mode="sample", uniforms=u)
    print("%s: sample %d steps; oracle mismatching draws %d; index-256 draws %d"
          % (tag, T_sample, int((o_sidx != sidx).sum()), int((sidx == 256).sum())))
    return cfg, w, ze, spk


def conv_form(tag, out, wav_over, B, T, peaked, stride, spk):
    """wavenet.py:24-100 through model.py:76-83 + decoder.py:12-37 (the teacher-forced / training forward)"""
    cfg = O.Config(wavenet=wav_over)
    w = O.make_weights(cfg, seed=1234, peaked=peaked)
    F = T // 64
    ze = O.synthetic_z_e(cfg, w, B, F, seed=1235, kind="scaled")
    x = O.synthetic_audio(B, T, seed=1237)
    m = build_model(w, wavenet_json(wav_over), ConstantEncoder(ze), x[:, :, None], spk)
    m._build()                                                         # model.py:146
    with tf.variable_scope("decoder"):                                 # model.py:147-148
        m._build_decoder()
    t0 = time.time()
    logits, labels = tf.Session().run([m.x_z_q, m.labels])
    logits = logits.reshape(B, T, 256)
    print("%s: conv form [%d x %d] in %.1fs" % (tag, B, T, time.time() - t0))
    out[tag + "_conv_logits"] = logits[:, stride - 1::stride].astype(np.float32)
    out[tag + "_conv_last_logits"] = logits[:, -8:].astype(np.float32)
    out[tag + "_conv_stride"] = np.int64(stride)
    out[tag + "_conv_labels"] = labels.astype(np.int16)
    out[tag + "_conv_logit_sum"] = logits.astype(np.float64).sum(-1).astype(np.float64)[:, ::16]
    return cfg, w, ze, x, logits


def section_small(out):
    spk = [None, 1, 2]
    cfg, w, ze, spk = run_config("small", out, SMALL_WAVENET, B=3, T=256, T_sample=256, T_teacher=256, peaked=False,
                                 logit_stride=4, spk=spk)
    cfg, w, ze, x, logits = conv_form("small", out, SMALL_WAVENET, 3, 256, False, 4, spk)
    o_idx, o_cond = O.encode_condition(ze, [0, 1, 2], w)
    o_logits, o_labels = O.wavenet_teacher_forced(cfg, w, x[:, :, None], o_cond)
    print("small: conv form oracle logit diff %.3g, labels equal %s"
          % (np.abs(o_logits.reshape(3, 256, 256) - logits).max(), np.array_equal(o_labels, out["small_conv_labels"])))


def section_full(out):
    B = 16
    spk = [b % 4 for b in range(B)]
    T = 512 if QUICK else 4096
    run_config("full", out, None, B=B, T=T, T_sample=256 if QUICK else 1024, T_teacher=128 if QUICK else 512,
               peaked=True, logit_stride=32, spk=spk)


def section_cfg5(out):
    """BASELINE config 5: teacher-forced forward, length 6656, batch 8 (encoders '64' / 'Magenta' share hop 64)"""
    B, T = (2, 1024) if QUICK else (8, 6656)
    spk = [b % 4 for b in range(B)]
    conv_form("cfg5", out, None, B, T, False, 64, spk)


def section_kat(out):
    """utils.py:13-46 + mu_law_ops.py known answers, incl. index 256 and first-index ties"""
    out["kat_decode_lut"] = ref_mu.mu_law_decode_np(np.arange(257))
    sess = tf.Session()
    xs = np.concatenate([np.linspace(-1.5, 1.5, 301), ref_mu.mu_law_decode_np(np.arange(257))]).astype(np.float32)
    out["kat_encode_x"] = xs
    out["kat_encode_float"] = sess.run(ref_mu.mu_law_encode(tf.constant(xs.reshape(-1, 1)))).reshape(-1)
    out["kat_encode_int"] = sess.run(ref_mu.mu_law_encode(tf.constant(xs.reshape(-1, 1)), to_int=True)).reshape(-1).astype(np.int32)
    rng = np.random.default_rng(77)
    pdf = rng.random((12, 256)).astype(np.float32) ** 8
    pdf /= pdf.sum(1, keepdims=True, dtype=np.float32)
    pdf[0] = np.float32(1.0 / 256)                         # flat: greedy -> 0
    pdf[1] = 0
    pdf[1, 255] = 1                                        # one-hot at the end
    pdf[2] = 0
    pdf[2, [17, 200]] = 0.5                                # two-way tie: greedy -> 17
    pdf[3] = np.float32(0.00390624)                        # cdf ends below 1 -> index 256 possible
    out["kat_pdf"] = pdf
    out["kat_greedy"] = ref_utils.decode(pdf, mode="greedy")
    seeds, draws = [], []
    for seed in range(40):
        np.random.seed(seed)
        draws.append(ref_utils.decode(pdf, mode="sample"))
        seeds.append(seed)
    out["kat_sample_seeds"] = np.array(seeds)
    out["kat_sample"] = np.stack(draws)
    # force the overflow: uniform just below 1 against row 3's short cdf
    real = np.random.rand
    try:
        np.random.rand = lambda n: np.full(n, 0.99999999)
        out["kat_sample_u_high"] = ref_utils.decode(pdf, mode="sample")
    finally:
        np.random.rand = real
    print("kat: sample(u=0.99999999) ->", out["kat_sample_u_high"][:4], " (1.0446 = index 256)")


def section_encoders(out):
    """Encoder/encoder.py:8-64 on a short input"""
    cfg = O.Config()
    B, T = 2, 2048
    x = O.synthetic_audio(B, T, seed=1237)[:, :, None]
    sess = tf.Session()
    for name, cls, mk, fwd in (("enc64", Encoder_64, O.make_encoder64_weights, O.encoder64_forward),
                               ("encmag", Encoder_Magenta, O.make_encoder_magenta_weights, O.encoder_magenta_forward)):
        w = dict(O.make_weights(O.Config(wavenet=SMALL_WAVENET)))
        ew = mk(cfg)
        w.update(ew)
        m = build_model(w, wavenet_json(SMALL_WAVENET), cls(64), x, [0, 1])
        m.build_generator()
        z = sess.run(m.z_e)
        created = dict(tf_shim.created_variables())
        assert all(k in created and created[k] == v.shape for k, v in ew.items()), "encoder variable names differ"
        out[name + "_z_e"] = z.astype(np.float32)
        print("%s: z_e %s, oracle diff %.3g (max |z| %.3g)" % (name, z.shape, np.abs(fwd(cfg, ew, x) - z).max(), np.abs(z).max()))


def section_enc2019(out):
    """Encoder/encoder.py:66-98 + Encoder/encoder_ops.py:14-69 on a short input (T a multiple of 320, quirk Q10).  The
    conv stack, its residual / `relu + relu` structure and the variable names are the reference's code; the
    tf.contrib.signal leaf operators are tf_shim's restatement of TensorFlow's published algorithms."""
    cfg = O.Config()
    B, T = 2, 7680
    x = O.synthetic_audio(B, T, seed=1237)[:, :, None]
    sess = tf.Session()
    w = dict(O.make_weights(O.Config(wavenet=SMALL_WAVENET)))
    ew = O.make_encoder2019_weights(cfg)
    w.update(ew)
    m = build_model(w, wavenet_json(SMALL_WAVENET), Encoder_2019(64), x, [0, 1])
    m.build_generator()
    z = sess.run(m.z_e)
    created = dict(tf_shim.created_variables())
    assert all(k in created and created[k] == v.shape for k, v in ew.items()), "encoder variable names differ"
    out["enc2019_z_e"] = z.astype(np.float32)
    zo = O.encoder2019_forward(cfg, ew, x)
    print("enc2019: z_e %s, oracle diff %.3g (max |z| %.3g)" % (z.shape, np.abs(zo - z).max(), np.abs(z).max()))


def section_cfg5_e2e(out):
    """BASELINE config 5 end to end for the three encoder variants: audio -> Encoder_* -> VQ -> speaker concat ->
    teacher-forced conv-form decoder (model.py:36-42,57-87,76-83; decoder.py:12-37; wavenet.py:24-100), default 30-layer
    WaveNet, batch 8.  '64' and 'Magenta': T = 6656 (hop 64, 104 frames).  '2019': T = 7680 (hop 320, 24 frames) - the
    reference cannot run this encoder at 6656 (21 frames do not divide 6656)."""
    cfg = O.Config()
    B = 2 if QUICK else 8
    spk = [b % 4 for b in range(B)]
    for tag, cls, mk, T, hop in (("e2e64", Encoder_64, O.make_encoder64_weights, 6656, 64),
                                 ("e2emag", Encoder_Magenta, O.make_encoder_magenta_weights, 6656, 64),
                                 ("e2e2019", Encoder_2019, O.make_encoder2019_weights, 7680, 320)):
        if QUICK:
            T = hop * 4
        w = dict(O.make_weights(cfg, seed=1234))
        w.update(mk(cfg))
        x = O.synthetic_audio(B, T, seed=1237)
        t0 = time.time()
        parts = []
        for b0 in range(0, B, 2):                                          # streams are independent: two at a time (memory)
            m = build_model(w, wavenet_json(None), cls(64), x[b0:b0 + 2, :, None], spk[b0:b0 + 2])
            m._build()
            with tf.variable_scope("decoder"):
                m._build_decoder()
            parts.append(tf.Session().run([m.z_e, m.q_z_x, m.x_z_q]))
        z_e = np.concatenate([p_[0] for p_ in parts])
        idx = np.concatenate([p_[1] for p_ in parts])
        logits = np.concatenate([p_[2].reshape(-1, T, 256) for p_ in parts])
        print("%s: end-to-end conv form [%d x %d], %d frames, in %.1fs" % (tag, B, T, z_e.shape[1], time.time() - t0), flush=True)
        out[tag + "_T"] = np.int64(T)
        out[tag + "_z_e"] = z_e.astype(np.float32)
        out[tag + "_idx"] = idx.astype(np.int16)
        out[tag + "_logits"] = logits[:, 127::128].astype(np.float32)
        out[tag + "_last_logits"] = logits[:, -8:].astype(np.float32)
        out[tag + "_logit_sum"] = logits.astype(np.float64).sum(-1)[:, ::16]


def section_magenta(out):
    """Magenta/ fast generation: the reference's FastGenerationConfig.build (Magenta/config.py:18-138 over
    Magenta/masked.py:31-35,133-174) is imported unmodified and driven by Magenta/generate.py:73-84's loop (re-typed: it is
    module-level script code), full size: 50 layers, width 256, skip 512, k = 2.  The local condition fed per step is
    e_k (config.py:242), here rows of the synthetic codebook."""
    import importlib.util
    mdir = os.path.join(REF, "Magenta")
    sys.path.append(mdir)                                              # `from masked import *`
    spec = importlib.util.spec_from_file_location("magenta_config", os.path.join(mdir, "config.py"))
    mcfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mcfg)
    B, T = 3, 192
    spk = [0, 5, 17]
    mw = O.make_magenta_fastgen_weights(peaked=True)
    rng = np.random.default_rng(1239)
    codes = rng.integers(0, 512, size=(B, T // 64))
    encoding = mw["embedding"][codes]                                  # e_k rows
    out["magenta_codes"] = codes.astype(np.int16)
    out["magenta_speakers"] = np.array(spk)
    sess = tf.Session()

    def fresh():
        tf.reset_default_graph()
        tf.set_variable_values({k: v for k, v in mw.items() if k != "embedding"})
        gen = mcfg.FastGenerationConfig(batch_size=B)
        x_t = tf.placeholder(tf.float32, shape=[B, 1])
        net = gen.build(inputs=x_t, gc=tf.constant(onehot(spk, 109)[:, 0]))
        created = dict(tf_shim.created_variables())
        want = dict((k, v.shape) for k, v in mw.items() if k != "embedding")
        assert created == want, "Magenta variable names / shapes differ from the oracle's table"
        return x_t, net

    def loop(x_t, net, length, mode, teacher=None, seed=None):        # Magenta/generate.py:73-84
        audio = np.zeros([B, 1], dtype=np.float32)
        to_write = np.zeros([B, length], dtype=np.float32)
        probs_all = np.zeros([B, length, 256], dtype=np.float32)
        logits_all = np.zeros([B, length, 256], dtype=np.float32)
        logits_node = net["predictions"].inputs[0]
        sess.run(net["init_ops"])
        if seed is not None:
            np.random.seed(seed)
        ratio = length // encoding.shape[1]
        for i in range(length):
            probs, _, lg = sess.run([net["predictions"], net["push_ops"], logits_node],
                                    {x_t: audio, net["encoding"]: encoding[:, i // ratio]})
            decoded = ref_utils.decode(probs, mode=mode)
            to_write[:, i], probs_all[:, i], logits_all[:, i] = decoded, probs, lg
            audio = decoded.reshape([B, 1]) if teacher is None else np.asarray(teacher[:, i:i + 1], dtype=np.float32)
        return to_write, probs_all, logits_all

    x = O.synthetic_audio(B, T, seed=1237)
    t0 = time.time()
    x_t, net = fresh()
    _, probs, logits = loop(x_t, net, T, "greedy", teacher=x)
    _, _, o_logits, _ = O.magenta_generate(mw, encoding, spk, T, mode="greedy", teacher=x)
    print("magenta: teacher-forced %d steps in %.1fs; oracle logit diff %.3g (max |logit| %.3g)"
          % (T, time.time() - t0, np.abs(o_logits - logits).max(), np.abs(logits).max()))
    out["magenta_teacher_logits"] = logits.astype(np.float32)
    x_t, net = fresh()
    audio, probs, _ = loop(x_t, net, T, "greedy")
    gidx = audio_to_index(audio)
    out["magenta_greedy_idx"] = gidx
    out["magenta_greedy_margin"] = np.stack([O.draw_margin(probs[:, i], "greedy") for i in range(T)], 1).astype(np.float32)
    _, o_gidx, _, _ = O.magenta_generate(mw, encoding, spk, T, mode="greedy")
    print("magenta: greedy oracle mismatching draws %d, distinct values %d" % (int((o_gidx != gidx).sum()), len(np.unique(gidx))))
    seed = 1240
    x_t, net = fresh()
    audio, probs, _ = loop(x_t, net, T, "sample", seed=seed)
    sidx = audio_to_index(audio)
    np.random.seed(seed)
    u = np.stack([np.random.rand(B) for _ in range(T)])
    out["magenta_sample_idx"] = sidx
    out["magenta_sample_seed"] = np.int64(seed)
    out["magenta_sample_margin"] = np.stack([O.draw_margin(probs[:, i], "sample", u[i]) for i in range(T)], 1).astype(np.float32)
    _, o_sidx, _, _ = O.magenta_generate(mw, encoding, spk, T, mode="sample", uniforms=u)
    print("magenta: sample oracle mismatching draws %d" % int((o_sidx != sidx).sum()))


def main():
    sections = [("ref_magenta", section_magenta), ("ref_enc2019", section_enc2019), ("ref_cfg5_e2e", section_cfg5_e2e), ("ref_vars", section_variables), ("ref_vq", section_vq), ("ref_kat", section_kat),
                ("ref_small", section_small), ("ref_encoders", section_encoders), ("ref_cfg5", section_cfg5),
                ("ref_full", section_full)]
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for fname, fn in sections:
        if only and fname not in only:
            continue
        out = {}
        t0 = time.time()
        fn(out)
        path = os.path.join(HERE, fname + ("_quick" if QUICK else "") + ".npz")
        np.savez_compressed(path, **out)
        print("wrote %s (%.0f KB) in %.0fs" % (path, os.path.getsize(path) / 1024, time.time() - t0))


if __name__ == "__main__":
    main()
