"""CPU oracle for the VQ-VAE-WaveNet inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's algorithm for the path named in
BASELINE.json (VQ nearest-codebook lookup -> conditioning -> WaveNet fast generation ->
mu-law softmax draw).  It is imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product path (the CUDA library
behind include/vqwn.h) never routes through it.

PARITY PINNED against the reference's own code: TensorFlow 1.x cannot be installed in this image, so
tests/golden/make_ref_golden.py executes the UNMODIFIED reference files (model.py, Decoder/decoder.py,
Decoder/decoder_ops.py, Decoder/WaveNet/wavenet.py, Decoder/WaveNet/wavenet_ops.py, mu_law_ops.py, utils.py,
Encoder/encoder.py) on a NumPy stand-in for the TF leaf operators (tests/golden/tf_shim.py: matmul, FIFOQueue,
conv2d via torch, get_variable/variable_scope, ...), and writes tests/golden/ref_*.npz.  tests/test_ref_golden.py
proves this restatement equals those outputs: variable names/shapes and receptive field, VQ indices / z_q /
condition on BASELINE config 2 (bit-exact), queue-form logits and softmax (bit-exact), greedy and seeded-sample
sequences incl. 16 streams x 4096 steps on the full configuration (identical), conv-form logits incl. BASELINE
config 5 at 8 x 6656 (<= 2e-5, the shim's conv2d is torch), decode/sample known answers incl. index 256, both
encoders.  What stays unverifiable without TensorFlow: TF's own kernels' rounding (MatMul summation order) and
the EMA shadow-variable names inside checkpoints.
Further pins (tests/test_oracle.py):
  * the five shipped WAVs (results/VCTK/p225_001/*.wav) lie on the mu-law decode grid
    (fixture tests/golden/wav_grid.npz) -> pins mu_law_decode_np and the float32 WAV format;
  * queue form (wavenet.py:103-172) == padded dilated-conv form (wavenet.py:24-100);
  * two VQ distance formulations (model.py:60-65 direct, Magenta/sonnet.py:91-95 expanded);
  * an independent torch.nn.functional.conv1d witness of the conv form.

All citations are file:line relative to /root/reference.
Everything is float32 unless the reference itself uses another type at that point.
"""
from __future__ import annotations

import json
from collections import deque

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# Configuration (model_parameters.json / wavenet_parameters.json, read as generate.py:63-64
# and wavenet.py:10-21 do)
# --------------------------------------------------------------------------------------
class Config:
    """Hyper-parameters of the path.  Defaults = the two JSON files of the reference."""

    def __init__(self, wavenet=None, model=None, num_speakers=109):
        w = dict(
            quantization_channels=256, num_cycles=3, num_cycle_layers=10,
            dilation_rates=[2 ** i for i in range(10)] * 3, kernel_size=3,
            dilation_filters=256, skip_filters=512, residual_filters=256,
            preprocess=dict(kernel_size=32, filters=256))
        if wavenet:
            w.update(wavenet)
        m = dict(encoder="64", use_vq=True, speaker_embedding=64, k=512, latent_dim=64, beta=0.25)
        if model:
            m.update(model)
        # wavenet.py:13
        assert len(w["dilation_rates"]) == w["num_cycles"] * w["num_cycle_layers"]
        self.wavenet = w
        self.model = m
        self.num_speakers = num_speakers
        self.q = w["quantization_channels"]
        self.dilations = list(w["dilation_rates"])
        self.ksize = w["kernel_size"]
        self.R = w["residual_filters"]
        self.G = w["dilation_filters"]            # gate width; conv makes 2*G channels
        self.S = w["skip_filters"]
        self.pre_k = w["preprocess"]["kernel_size"]
        self.pre_f = w["preprocess"]["filters"]
        self.K = m["k"]
        self.D = m["latent_dim"]
        self.spk_dim = m["speaker_embedding"]
        self.C = self.D + self.spk_dim            # decoder.py:49-50 concat
        # wavenet.py:15-17
        self.receptive_field = sum(self.dilations) * (self.ksize - 1) + 1 + self.pre_k - 1

    @classmethod
    def from_files(cls, model_path, num_speakers=109):
        with open(model_path) as f:
            m = json.load(f)
        import os
        wp = m["wavenet_parameters"]
        if not os.path.isabs(wp) and not os.path.exists(wp):
            wp = os.path.join(os.path.dirname(model_path), wp)
        with open(wp) as f:
            w = json.load(f)
        return cls(w, m, num_speakers)

    def layer_scope(self, i):
        # wavenet.py:134-135
        n = self.wavenet["num_cycle_layers"]
        return "decoder/cycle_%d/layer_%d" % (1 + i // n, 1 + i % n)


def tensor_specs(cfg):
    """(tf_variable_name, shape) in the fixed order used for synthetic weights."""
    R, G, S, C, q = cfg.R, cfg.G, cfg.S, cfg.C, cfg.q
    specs = [
        ("embedding/embedding", (cfg.K, cfg.D)),                       # model.py:47-49
        ("speaker_embedding", (cfg.num_speakers, cfg.spk_dim)),        # model.py:23-26
        ("decoder/preprocess/kernel", (cfg.pre_k, 1, cfg.pre_f)),      # wavenet_ops.py:173-176
        ("decoder/preprocess/bias", (cfg.pre_f,)),
        ("decoder/skip/kernel", (1, cfg.pre_f, S)),                    # wavenet.py:127-128
        ("decoder/skip/bias", (S,)),
    ]
    for i in range(len(cfg.dilations)):
        s = cfg.layer_scope(i)
        specs += [
            (s + "/gated/kernel", (cfg.ksize, R, 2 * G)),
            (s + "/gated/bias", (2 * G,)),
            (s + "/gated/local_condition/kernel", (1, C, 2 * G)),
            (s + "/skip/kernel", (1, G, S)),
            (s + "/skip/bias", (S,)),
            (s + "/residual/kernel", (1, G, R)),
            (s + "/residual/bias", (R,)),
        ]
    specs += [
        ("decoder/postprocess1/kernel", (1, S, S)),
        ("decoder/postprocess1/bias", (S,)),
        ("decoder/postprocess1/local_condition/kernel", (1, C, S)),
        ("decoder/postprocess2/kernel", (1, S, q)),
        ("decoder/postprocess2/bias", (q,)),
    ]
    return specs


def make_weights(cfg, seed=1234, peaked=False):
    """Seeded synthetic weights keyed by reference variable name.

    kernels U(+-sqrt(3/fan_in)) as tf.uniform_unit_scaling_initializer(1.0) (wavenet_ops.py:69),
    biases U(+-0.05) (reference inits 0; non-zero so bias paths are exercised),
    codebook factor 1.7 (model.py:49), speaker table factor 2.0 (model.py:26).
    peaked=True scales postprocess2/kernel by 8 so the softmax is peaked (greedy tests)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in tensor_specs(cfg):
        if name.endswith("bias"):
            a = rng.uniform(-0.05, 0.05, size=shape)
        else:
            fan_in = int(np.prod(shape[:-1]))
            lim = np.sqrt(3.0 / fan_in)
            if name == "embedding/embedding":
                lim *= 1.7
            elif name == "speaker_embedding":
                lim *= 2.0
            a = rng.uniform(-lim, lim, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=F32)
    if peaked:
        out["decoder/postprocess2/kernel"] = (out["decoder/postprocess2/kernel"] * F32(8)).astype(F32)
    return out


# --------------------------------------------------------------------------------------
# mu-law codec (mu_law_ops.py) and sampling (utils.py:13-46)
# --------------------------------------------------------------------------------------
def mu_law_encode(x, quantization_channels=256, to_int=False):
    """mu_law_ops.py:5-15 (float path :6-8, int path :9-11).  float32 arithmetic."""
    mu = F32(quantization_channels - 1)
    x = np.clip(np.asarray(x, dtype=F32), F32(-1.0), F32(1.0))
    y = (np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)).astype(F32)
    if to_int:
        # tf.cast(float->int32) truncates toward zero; argument is >= 0.5 here
        y = ((y + F32(1)) / F32(2) * mu + F32(0.5)).astype(F32).astype(np.int32)
    return y


def mu_law_decode_np(y, quantization_channels=256):
    """mu_law_ops.py:26-31, float32 NumPy."""
    mu = np.asarray(quantization_channels - 1, dtype=F32)
    y = (2 * np.asarray(y, dtype=F32) / mu) - 1
    x = np.sign(y) * ((1 + mu) ** abs(y) - 1) / mu
    return x.astype(F32)


def decode_lut(quantization_channels=256):
    """audio value for every index the draw can return: 0..q (q itself happens in sample mode
    when the float32 cdf ends below the uniform draw, utils.py:20-25)."""
    return mu_law_decode_np(np.arange(quantization_channels + 1), quantization_channels)


def encode_lut(quantization_channels=256):
    """network input for every index: mu_law_encode(decode(k)) (generate.py:112-113 feeds the
    decoded float back; wavenet.py:113 re-encodes it, clipping entry q to 1.0)."""
    return mu_law_encode(decode_lut(quantization_channels), quantization_channels)


def sample_indices(pdf, uniforms):
    """utils.py:19-25 with the np.random.rand(batch) draw made injectable.
    cdf is a sequential float32 running sum; the comparison happens in float64."""
    cdf = np.cumsum(np.asarray(pdf, dtype=F32), axis=1)
    pred = np.zeros(cdf.shape[0], dtype=F32)
    for i, prob in enumerate(np.asarray(uniforms, dtype=np.float64)):
        pred[i] = cdf[i].searchsorted(prob)
    return pred


def decode_indices(predictions, mode="sample", uniforms=None):
    """utils.py:30-46 up to (not including) the mu-law decode: returns indices [B]."""
    if mode == "sample":
        return sample_indices(predictions, uniforms).astype(np.int32)
    elif mode == "greedy":
        return np.argmax(predictions, axis=-1).astype(np.int32)
    raise NotImplementedError("decode mode %s not implemented" % mode)


def decode(predictions, mode="sample", quantization_channels=256, uniforms=None):
    """utils.py:30-46.  NB sample() ignores quantization_channels (utils.py:41)."""
    idx = decode_indices(predictions, mode, uniforms)
    qc = 256 if mode == "sample" else quantization_channels
    return mu_law_decode_np(idx.astype(F32), qc)


# --------------------------------------------------------------------------------------
# VQ bottleneck (model.py:45-74) + speaker condition (model.py:19-27) + concat
# --------------------------------------------------------------------------------------
def vq_distances(z_e, embedding):
    """model.py:60-61 direct form, float32, materialising [..., K, D] like the reference."""
    z = np.asarray(z_e, dtype=F32)
    e = np.asarray(embedding, dtype=F32)
    diff = z[..., None, :] - e
    return np.sum(diff * diff, axis=-1, dtype=F32)


def vq_discretise(z_e, embedding, chunk=2048):
    """model.py:57-74: returns (q_z_x int64, e_k, z_q) with z_q = z_e + (e_k - z_e)."""
    z = np.asarray(z_e, dtype=F32)
    flat = z.reshape(-1, z.shape[-1])
    idx = np.empty(flat.shape[0], dtype=np.int64)
    for s in range(0, flat.shape[0], chunk):           # chunked only to bound memory
        idx[s:s + chunk] = np.argmin(vq_distances(flat[s:s + chunk], embedding), axis=-1)
    e_k = np.asarray(embedding, dtype=F32)[idx]
    z_q = (flat + (e_k - flat)).astype(F32)
    shp = z.shape[:-1]
    return idx.reshape(shp), e_k.reshape(z.shape), z_q.reshape(z.shape)


def vq_distances_f64(z_e, embedding):
    """float64 distances, used by tests to decide whether an index mismatch is a near-tie."""
    z = np.asarray(z_e, dtype=np.float64)
    e = np.asarray(embedding, dtype=np.float64)
    return (z * z).sum(-1, keepdims=True) - 2.0 * z @ e.T + (e * e).sum(-1)[None, :]


def vq_discretise_expanded(z_e, embedding):
    """Magenta/sonnet.py:91-98: ||z||^2 - 2 z.W + ||w||^2, argmax(-d)."""
    z = np.asarray(z_e, dtype=F32)
    flat = z.reshape(-1, z.shape[-1])
    w = np.asarray(embedding, dtype=F32).T
    d = (np.sum(flat ** 2, 1, keepdims=True, dtype=F32) - F32(2) * (flat @ w)
         + np.sum(w ** 2, 0, keepdims=True, dtype=F32))
    return np.argmax(-d, 1).reshape(z.shape[:-1])


def speaker_rows(speaker_onehot, speaker_embedding):
    """model.py:19-27: argmax over the one-hot (all-zero row -> index 0) then gather.
    speaker_onehot [B,1,N] -> h [B,1,spk_dim]."""
    idx = np.argmax(np.asarray(speaker_onehot), axis=-1)
    return np.asarray(speaker_embedding, dtype=F32)[idx]


def concat_condition(net, global_condition):
    """Decoder/decoder_ops.py:39-43: tile speaker row over frames, concat on channels."""
    g = np.tile(global_condition, [1, net.shape[1], 1])
    return np.concatenate([net, g], axis=-1).astype(F32)


def encode_condition(z_e, speaker_idx, weights):
    """model.py:133-142 + model.py:85-87 + decoder.py:49-50: what generate.py:92 evaluates
    as model.encoding ([B,F,D+spk])."""
    idx, _, z_q = vq_discretise(z_e, weights["embedding/embedding"])
    h = weights["speaker_embedding"][np.asarray(speaker_idx)][:, None, :]
    return idx, concat_condition(z_q, h)


# --------------------------------------------------------------------------------------
# Fast generation (Decoder/WaveNet/wavenet.py:103-172, wavenet_ops.py:147-267)
# --------------------------------------------------------------------------------------
class _FastConv:
    """wavenet_ops.py:163-195: one stride of a dilated causal conv with (k-1) chained
    FIFO queues of capacity d, pre-filled with d zero items (:181-184)."""

    def __init__(self, kernel, bias, dilation, batch):
        self.kernel = kernel
        self.bias = bias
        self.k = kernel.shape[0]
        self.d = dilation
        self.batch = batch
        self.cin = kernel.shape[1]
        self.queues = None

    def init(self):
        z = np.zeros((self.batch, self.cin), dtype=F32)
        self.queues = [deque([z] * self.d) for _ in range(self.k - 1)]

    def __call__(self, current):
        k = self.k
        new_state = current @ self.kernel[k - 1] + self.bias          # :178
        for i in range(1, k):
            q = self.queues[i - 1]
            past = q.popleft()                                        # :187
            q.append(current)                                         # :188
            current = past                                            # :189
            new_state = new_state + past @ self.kernel[k - i - 1]      # :193
        return new_state.astype(F32)


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)


def softmax(x):
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp((x - m).astype(F32)).astype(F32)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=F32)).astype(F32)


class FastWavenet:
    """wavenet.py:103-172.  step() == sess.run([predictions, push_ops]) (generate.py:109)."""

    def __init__(self, cfg, weights, batch):
        self.cfg, self.w, self.batch = cfg, weights, batch
        w = weights
        self.pre = _FastConv(w["decoder/preprocess/kernel"], w["decoder/preprocess/bias"], 1, batch)
        self.layers = []
        for i, d in enumerate(cfg.dilations):
            s = cfg.layer_scope(i)
            self.layers.append(_FastConv(w[s + "/gated/kernel"], w[s + "/gated/bias"], d, batch))
        self.init_ops()

    def init_ops(self):
        """generate.py:105"""
        self.pre.init()
        for c in self.layers:
            c.init()

    def step(self, input_t, local_condition_t):
        """input_t [B,1] float audio in [-1,1]; local_condition_t [B,C] -> (probs, logits)"""
        cfg, w = self.cfg, self.w
        x = mu_law_encode(np.asarray(input_t, dtype=F32).reshape(self.batch, 1))   # wavenet.py:113 (always 256)
        lc = np.asarray(local_condition_t, dtype=F32)
        current = self.pre(x)                                                       # :119-124
        skip = current @ w["decoder/skip/kernel"][0] + w["decoder/skip/bias"]       # :127-128
        G = cfg.G
        for i, conv in enumerate(self.layers):
            s = cfg.layer_scope(i)
            net = conv(current)                                                     # wavenet_ops.py:228-229
            net = net + lc @ w[s + "/gated/local_condition/kernel"][0]              # :230-231, :208
            gated = (np.tanh(net[:, :G]) * _sigmoid(net[:, G:])).astype(F32)        # :235-236
            skip = skip + (gated @ w[s + "/skip/kernel"][0] + w[s + "/skip/bias"])  # :261-262; wavenet.py:144
            current = current + (gated @ w[s + "/residual/kernel"][0] + w[s + "/residual/bias"])  # :264-265; :145
        net = np.maximum(skip, 0)                                                   # wavenet.py:153
        net = net @ w["decoder/postprocess1/kernel"][0] + w["decoder/postprocess1/bias"]
        net = net + lc @ w["decoder/postprocess1/local_condition/kernel"][0]        # :157-159
        net = np.maximum(net, 0)
        logits = (net @ w["decoder/postprocess2/kernel"][0] + w["decoder/postprocess2/bias"]).astype(F32)
        return softmax(logits), logits


def draw_margin(probs, mode, uniforms=None):
    """How far each stream's draw is from flipping: greedy -> (p1 - p2) / p1 of the two largest
    probabilities; sample -> distance from the uniform to the nearest float32 cdf boundary.
    Tests use it to tell a near-tie (documented tolerance) from a real divergence."""
    if mode == "greedy":
        s = np.sort(np.asarray(probs, dtype=np.float64), axis=-1)
        return ((s[:, -1] - s[:, -2]) / s[:, -1]).astype(F32)
    cdf = np.cumsum(np.asarray(probs, dtype=F32), axis=1).astype(np.float64)
    return np.min(np.abs(cdf - np.asarray(uniforms, dtype=np.float64)[:, None]), axis=1).astype(F32)


def generate(cfg, weights, encoding, length, mode="greedy", uniforms=None, teacher=None,
             return_logits=False, return_margins=False):
    """generate.py:103-113.  encoding [B,F,C]; returns (audio [B,T] float32, idx [B,T] int32
    [, logits [B,T,q]]).  teacher [B,T]: feed teacher[:, i-1] instead of the model's own draw
    (per-step logit parity; first input is 0 as shift_right does, wavenet_ops.py:9-14).
    uniforms [T,B] float64 replaces the unseeded np.random.rand(B) of utils.py:22."""
    B = encoding.shape[0]
    net = FastWavenet(cfg, weights, B)
    audio = np.zeros([B, 1], dtype=F32)
    to_write = np.zeros([B, length], dtype=F32)
    idx_out = np.zeros([B, length], dtype=np.int32)
    logits_out = np.zeros([B, length, cfg.q], dtype=F32) if return_logits else None
    margins = np.zeros([B, length], dtype=F32) if return_margins else None
    ratio = length // encoding.shape[1]                                             # generate.py:107
    for i in range(length):
        probs, logits = net.step(audio, encoding[:, i // ratio])
        u = None if uniforms is None else uniforms[i]
        idx = decode_indices(probs, mode, u)
        if return_margins:
            margins[:, i] = draw_margin(probs, mode, u)
        decoded = mu_law_decode_np(idx.astype(F32), 256 if mode == "sample" else cfg.q)
        to_write[:, i] = decoded
        idx_out[:, i] = idx
        if return_logits:
            logits_out[:, i] = logits
        if teacher is None:
            audio = np.expand_dims(decoded, -1)
        else:
            audio = np.asarray(teacher[:, i:i + 1], dtype=F32)
    out = (to_write, idx_out)
    if return_logits:
        out = out + (logits_out,)
    if return_margins:
        out = out + (margins,)
    return out


# --------------------------------------------------------------------------------------
# Teacher-forced conv form (wavenet.py:24-100, wavenet_ops.py:9-14,59-138) -- second witness
# --------------------------------------------------------------------------------------
def shift_right(x):
    """wavenet_ops.py:9-14"""
    return np.concatenate([np.zeros_like(x[:, :1]), x[:, :-1]], axis=1)


def conv1d_v2(net, kernel, bias=None, dilations=1):
    """wavenet_ops.py:59-90 with padding='CAUSAL': left-pad d*(k-1) zeros, VALID dilated conv."""
    k = kernel.shape[0]
    T = net.shape[1]
    pad = dilations * (k - 1)
    xp = np.concatenate([np.zeros((net.shape[0], pad, net.shape[2]), dtype=F32), net], axis=1)
    out = None
    for j in range(k):
        term = xp[:, j * dilations: j * dilations + T] @ kernel[j]
        out = term if out is None else out + term
    if bias is not None:
        out = out + bias
    return out.astype(F32)


def add_condition(net, condition, kernel):
    """wavenet_ops.py:93-101: bias-free 1x1 on [B,F,C], each frame broadcast over T//F samples."""
    B, T, C = net.shape
    F = condition.shape[1]
    enc = condition @ kernel[0]
    net = net.reshape(B, F, T // F, C) + enc[:, :, None, :]
    return net.reshape(B, T, C).astype(F32)


def wavenet_teacher_forced(cfg, weights, x, local_condition):
    """wavenet.py:24-100.  x [B,T,1] float audio, local_condition [B,F,C] -> (logits [B*T,q],
    labels [B*T])."""
    w = weights
    x = np.asarray(x, dtype=F32)
    labels = mu_law_encode(x, to_int=True).reshape(-1)                              # :33-34
    inputs = mu_law_encode(shift_right(x))                                          # :36-37
    net = conv1d_v2(inputs, w["decoder/preprocess/kernel"], w["decoder/preprocess/bias"])   # :42-45
    skip = conv1d_v2(net, w["decoder/skip/kernel"], w["decoder/skip/bias"])         # :53-55
    G = cfg.G
    for i, d in enumerate(cfg.dilations):
        s = cfg.layer_scope(i)
        a = conv1d_v2(net, w[s + "/gated/kernel"], w[s + "/gated/bias"], d)          # wavenet_ops.py:106
        a = add_condition(a, local_condition, w[s + "/gated/local_condition/kernel"])
        gated = (np.tanh(a[:, :, :G]) * _sigmoid(a[:, :, G:])).astype(F32)          # :112-113
        skip = skip + conv1d_v2(gated, w[s + "/skip/kernel"], w[s + "/skip/bias"])
        net = net + conv1d_v2(gated, w[s + "/residual/kernel"], w[s + "/residual/bias"])
    net = np.maximum(skip, 0)
    net = conv1d_v2(net, w["decoder/postprocess1/kernel"], w["decoder/postprocess1/bias"])
    net = add_condition(net, local_condition, w["decoder/postprocess1/local_condition/kernel"])
    net = np.maximum(net, 0)
    net = conv1d_v2(net, w["decoder/postprocess2/kernel"], w["decoder/postprocess2/bias"])
    return net.reshape(-1, cfg.q), labels


# --------------------------------------------------------------------------------------
# Encoder_64 (Encoder/encoder.py:8-26) -- SURVEY 8f #1: the step right before the path
# --------------------------------------------------------------------------------------
def encoder64_specs(cfg):
    """keras auto-names inside variable_scope('encoder') (model.py:134-135): conv1d, conv1d_1, ...,
    batch_normalization, batch_normalization_1, ...; kernels [k, in, out]."""
    specs = []
    cin = 1
    for i in range(7):
        sfx = "" if i == 0 else "_%d" % i
        cout = 768 if i < 6 else cfg.D
        k = 5 if i < 6 else 1
        specs += [("encoder/conv1d%s/kernel" % sfx, (k, cin, cout)), ("encoder/conv1d%s/bias" % sfx, (cout,)),
                  ("encoder/batch_normalization%s/gamma" % sfx, (cout,)), ("encoder/batch_normalization%s/beta" % sfx, (cout,)),
                  ("encoder/batch_normalization%s/moving_mean" % sfx, (cout,)),
                  ("encoder/batch_normalization%s/moving_variance" % sfx, (cout,))]
        cin = cout
    return specs


def make_encoder64_weights(cfg, seed=4321):
    """seeded synthetic encoder weights: glorot-uniform kernels (keras default), small biases, BN statistics
    away from the identity so every term is exercised"""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in encoder64_specs(cfg):
        if name.endswith("kernel"):
            lim = np.sqrt(6.0 / (shape[0] * shape[1] + shape[0] * shape[2]))
            a = rng.uniform(-lim, lim, size=shape)
        elif name.endswith("moving_variance"):
            a = rng.uniform(0.5, 1.5, size=shape)
        elif name.endswith("gamma"):
            a = rng.uniform(0.8, 1.2, size=shape)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=F32)
    return out


def encoder64_forward(cfg, weights, x):
    """Encoder/encoder.py:13-26.  x [B,T,1] -> z_e [B,T/64,latent_dim].
    Conv1D(768, k=5, strides=2, padding='same', relu): TF 'same' with stride 2 pads
    total = max((ceil(T/2)-1)*2 + 5 - T, 0), left = total // 2 (1 left / 2 right for even T).
    BatchNormalization() is called without training= -> inference form with the moving statistics,
    epsilon 1e-3 (keras default)."""
    net = np.asarray(x, dtype=F32)
    for i in range(7):
        sfx = "" if i == 0 else "_%d" % i
        K = weights["encoder/conv1d%s/kernel" % sfx]
        b = weights["encoder/conv1d%s/bias" % sfx]
        k = K.shape[0]
        stride = 2 if i < 6 else 1
        T = net.shape[1]
        To = (T + stride - 1) // stride
        total = max((To - 1) * stride + k - T, 0)
        left = total // 2
        xp = np.pad(net, [(0, 0), (left, total - left), (0, 0)])
        out = None
        for j in range(k):
            term = xp[:, j: j + (To - 1) * stride + 1: stride] @ K[j]
            out = term if out is None else out + term
        out = out + b
        if i < 6:
            out = np.maximum(out, 0)
        g = weights["encoder/batch_normalization%s/gamma" % sfx]
        be = weights["encoder/batch_normalization%s/beta" % sfx]
        mu = weights["encoder/batch_normalization%s/moving_mean" % sfx]
        var = weights["encoder/batch_normalization%s/moving_variance" % sfx]
        net = ((out - mu) * (g / np.sqrt(var + F32(1e-3))) + be).astype(F32)
    return net


# --------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY 8d
# --------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------
# Encoder_Magenta (Encoder/encoder.py:29-64) - SURVEY 8f #1, second encoder
# ---------------------------------------------------------------------------------------------
MAGENTA_DILATIONS = [1, 2, 4, 8, 16, 16]        # encoder.py:34
MAGENTA_FILTERS = 128                            # encoder.py:39
MAGENTA_KERNEL = 5                               # encoder.py:40


def encoder_magenta_specs(cfg):
    """variables under variable_scope('encoder') (model.py:134-135): conv1d_v2 creates 'kernel' [k,in,out] and 'bias'
    in the enclosing scope (wavenet_ops.py:66-76)"""
    C, k = MAGENTA_FILTERS, MAGENTA_KERNEL
    specs = [("encoder/preprocess/kernel", (k, 1, C)), ("encoder/preprocess/bias", (C,))]
    for i in range(len(MAGENTA_DILATIONS)):
        sc = "encoder/cycle_%d/layer_%d" % (1 + i // 6, 1 + i % 6)           # encoder.py:49-50
        specs += [(sc + "/dilated/kernel", (1, C, C)), (sc + "/dilated/bias", (C,)),
                  (sc + "/gate/kernel", (k, C, C)), (sc + "/gate/bias", (C,)),
                  (sc + "/filter/kernel", (k, C, C)), (sc + "/filter/bias", (C,)),
                  (sc + "/residual/kernel", (1, C, C)), (sc + "/residual/bias", (C,))]
    specs += [("encoder/postprocess/kernel", (1, C, cfg.D)), ("encoder/postprocess/bias", (cfg.D,))]
    return specs


def make_encoder_magenta_weights(cfg, seed=4322):
    """seeded synthetic weights: kernels U(+-sqrt(3/fan_in)) as uniform_unit_scaling_initializer(1.0)
    (wavenet_ops.py:69), small non-zero biases"""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in encoder_magenta_specs(cfg):
        if name.endswith("kernel"):
            lim = np.sqrt(3.0 / (shape[0] * shape[1]))
            a = rng.uniform(-lim, lim, size=shape)
        else:
            a = rng.uniform(-0.05, 0.05, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=F32)
    return out


def encoder_magenta_forward(cfg, weights, x):
    """Encoder/encoder.py:37-64.  x [B,T,1] -> z_e [B,T/64,latent_dim] (T a multiple of 64).
    shift_right + mu_law_encode (float path), causal k=5 preprocess conv, then 6 x [1x1 stride-2 conv 'dilated'
    (wavenet_ops.py:83-86: VALID with stride 2 keeps samples 0,2,4,...), gate / filter = causal dilated k=5 convs of
    it, tanh(gate) * sigmoid(filter), residual 1x1 added to the strided signal], then a 1x1 postprocess conv."""
    net = mu_law_encode(shift_right(np.asarray(x, dtype=F32)))
    en = conv1d_v2(net, weights["encoder/preprocess/kernel"], weights["encoder/preprocess/bias"])
    for i, dil in enumerate(MAGENTA_DILATIONS):
        sc = "encoder/cycle_%d/layer_%d" % (1 + i // 6, 1 + i % 6)
        d = conv1d_v2(en[:, ::2], weights[sc + "/dilated/kernel"], weights[sc + "/dilated/bias"])
        g = conv1d_v2(d, weights[sc + "/gate/kernel"], weights[sc + "/gate/bias"], dilations=dil)
        f = conv1d_v2(d, weights[sc + "/filter/kernel"], weights[sc + "/filter/bias"], dilations=dil)
        gated = (np.tanh(g) * (F32(1) / (F32(1) + np.exp(-f)))).astype(F32)
        en = (d + conv1d_v2(gated, weights[sc + "/residual/kernel"], weights[sc + "/residual/bias"])).astype(F32)
    return conv1d_v2(en, weights["encoder/postprocess/kernel"], weights["encoder/postprocess/bias"])


# ---------------------------------------------------------------------------------------------
# Encoder_2019 (Encoder/encoder.py:66-98, Encoder/encoder_ops.py:14-69) - SURVEY 8f #1, third encoder
# ---------------------------------------------------------------------------------------------
# The conv stack is the reference's own code; the MFCC front end calls tf.contrib.signal (TensorFlow r1.12-1.14, a
# third-party dependency absent from /root/reference), whose published algorithms are restated here:
#   stft(frame_length=400, frame_step=160, fft_length=400, periodic hann window, pad_end=True) -> |rfft| (201 bins);
#   linear_to_mel_weight_matrix(80, 201, 16000, 20, 8000): HTK mel scale 1127 ln(1 + f / 700), triangles on the mel
#     axis between 82 equally spaced edges, DC bin zeroed;
#   log(mel + 1e-6); mfccs_from_log_mel_spectrograms = DCT-II (unnormalised, factor 2) * rsqrt(2 * 80); first 13.
# PARITY of this front end is pinned only to the restatement that tests/golden/tf_shim.py shares (no TensorFlow here).
MFCC_SR, MFCC_FRAME, MFCC_STEP, MFCC_MELS, MFCC_COEFS = 16000, 400, 160, 80, 13
MFCC_LOWER_HZ, MFCC_UPPER_HZ = 20.0, 8000.0
ENC2019_HOP = 2 * MFCC_STEP          # one stride-2 conv after the 10 ms frames (encoder.py:82)


def hann_window_periodic(n):
    """tf.contrib.signal.hann_window(periodic=True): 0.5 - 0.5 cos(2 pi i / n)"""
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(F32)


def linear_to_mel_weight_matrix(num_mel_bins=MFCC_MELS, num_spectrogram_bins=MFCC_FRAME // 2 + 1, sample_rate=MFCC_SR,
                                lower_edge_hertz=MFCC_LOWER_HZ, upper_edge_hertz=MFCC_UPPER_HZ):
    """tf.contrib.signal.linear_to_mel_weight_matrix (mel_ops.py), evaluated in float64 and rounded to float32"""
    def hz_to_mel(f):
        return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    nyquist = sample_rate / 2.0
    lin = np.linspace(0.0, nyquist, num_spectrogram_bins)[1:]                 # bands_to_zero = 1 (the DC bin)
    spec_mel = hz_to_mel(lin)[:, None]
    edges = np.linspace(hz_to_mel(lower_edge_hertz), hz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    lower, center, upper = edges[:-2][None, :], edges[1:-1][None, :], edges[2:][None, :]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    w = np.maximum(0.0, np.minimum(lower_slopes, upper_slopes))
    return np.pad(w, [(1, 0), (0, 0)]).astype(F32)                             # [201, 80]


def dct2_matrix(n=MFCC_MELS, k=MFCC_COEFS):
    """mfccs_from_log_mel_spectrograms: dct(type=2) (X_c = 2 sum_m x_m cos(pi c (2m+1) / (2n))) * rsqrt(2n); [n, k]"""
    m = np.arange(n)[:, None]
    c = np.arange(k)[None, :]
    return (2.0 * np.cos(np.pi * c * (2 * m + 1) / (2.0 * n)) / np.sqrt(2.0 * n)).astype(F32)


def mfcc(batch_wav):
    """Encoder/encoder_ops.py:14-43.  batch_wav [B,T] -> [B, ceil(T/160), 13]"""
    x = np.asarray(batch_wav, dtype=F32)
    B, T = x.shape
    nfr = -(-T // MFCC_STEP)                                                   # pad_end=True
    xp = np.pad(x, [(0, 0), (0, (nfr - 1) * MFCC_STEP + MFCC_FRAME - T)])
    idx = np.arange(nfr)[:, None] * MFCC_STEP + np.arange(MFCC_FRAME)[None, :]
    frames = xp[:, idx] * hann_window_periodic(MFCC_FRAME)
    mag = np.abs(np.fft.rfft(frames.astype(F32), n=MFCC_FRAME, axis=-1)).astype(F32)
    mel = (mag @ linear_to_mel_weight_matrix()).astype(F32)
    logmel = np.log(mel + F32(1e-6)).astype(F32)
    return (logmel @ dct2_matrix()).astype(F32)


def encoder2019_specs(cfg):
    """keras auto-names inside variable_scope('encoder'): conv1d, conv1d_1, ..., conv1d_9 in creation order
    (encoder.py:75-96): 2 x conv_3_768, strided_conv_4_768, 2 + 4 x conv_3_768, linear_64"""
    specs = []
    shapes = [(3, MFCC_COEFS, 768), (3, 768, 768), (4, 768, 768)] + [(3, 768, 768)] * 6 + [(1, 768, cfg.D)]
    for i, shp in enumerate(shapes):
        sfx = "" if i == 0 else "_%d" % i
        specs += [("encoder/conv1d%s/kernel" % sfx, shp), ("encoder/conv1d%s/bias" % sfx, (shp[2],))]
    return specs


def make_encoder2019_weights(cfg, seed=4323):
    """seeded synthetic weights: glorot-uniform kernels (keras default) scaled down so the 2x 'relu + relu' blocks
    (encoder.py:91-93) keep activations O(1), small biases"""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in encoder2019_specs(cfg):
        if name.endswith("kernel"):
            lim = np.sqrt(6.0 / (shape[0] * shape[1] + shape[0] * shape[2]))
            a = rng.uniform(-lim, lim, size=shape)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=F32)
    return out


def _keras_conv1d_same(net, K, b, stride, relu):
    k = K.shape[0]
    T = net.shape[1]
    To = (T + stride - 1) // stride
    total = max((To - 1) * stride + k - T, 0)
    left = total // 2
    xp = np.pad(net, [(0, 0), (left, total - left), (0, 0)])
    out = None
    for j in range(k):
        term = xp[:, j: j + (To - 1) * stride + 1: stride] @ K[j]
        out = term if out is None else out + term
    out = (out + b).astype(F32)
    return np.maximum(out, 0) if relu else out


def encoder2019_forward(cfg, weights, x):
    """Encoder/encoder.py:72-98.  x [B,T,1] (T a multiple of 320) -> z_e [B,T/320,latent_dim].  Quirk Q17: the four
    'relu layers' are `relu + relu` (twice the conv output, no skip connection)."""
    def conv(i, net, stride=1, relu=True):
        sfx = "" if i == 0 else "_%d" % i
        return _keras_conv1d_same(net, weights["encoder/conv1d%s/kernel" % sfx], weights["encoder/conv1d%s/bias" % sfx], stride, relu)
    net = mfcc(np.asarray(x, dtype=F32)[:, :, 0])
    net = conv(0, net)
    net = conv(1, net) + net
    net = conv(2, net, stride=2)
    i = 3
    for _ in range(2):
        net = conv(i, net) + net
        i += 1
    for _ in range(4):
        r = conv(i, net)
        net = r + r
        i += 1
    return conv(9, net, relu=False)


# ---------------------------------------------------------------------------------------------
# Magenta/ fast-generation topology: Magenta/config.py:18-138 (FastGenerationConfig.build),
# Magenta/masked.py:133-174 (causal_linear, linear), Magenta/generate.py:52-87 (host loop)
# ---------------------------------------------------------------------------------------------
MAGENTA_NUM_STAGES, MAGENTA_NUM_LAYERS, MAGENTA_FILTER_LENGTH = 10, 50, 2      # config.py:5-7
MAGENTA_WIDTH, MAGENTA_SKIP_WIDTH, MAGENTA_BOTTLENECK, MAGENTA_K = 256, 512, 64, 512      # config.py:8-9,15-16
MAGENTA_SPEAKERS = 109                                                         # config.py:42


def magenta_fastgen_specs(num_layers=MAGENTA_NUM_LAYERS):
    """variables FastGenerationConfig.build creates, by name (config.py:40-130; masked.py:146-151,167-171), plus the
    codebook Config.build owns (config.py:229)"""
    W, S, E = MAGENTA_WIDTH, MAGENTA_SKIP_WIDTH, MAGENTA_BOTTLENECK
    specs = [("embedding", (MAGENTA_K, E)), ("speaker_emb", (MAGENTA_SPEAKERS, E)),
             ("startconv/W", (1, MAGENTA_FILTER_LENGTH, 1, W)), ("startconv/biases", (W,)),
             ("skip_start/W", (1, 1, W, S)), ("skip_start/biases", (S,))]
    for i in range(1, num_layers + 1):
        specs += [("dilatedconv_%d/W" % i, (1, MAGENTA_FILTER_LENGTH, W, 2 * W)), ("dilatedconv_%d/biases" % i, (2 * W,)),
                  ("cond_map_%d/W" % i, (1, 1, E, 2 * W)), ("cond_map_%d/biases" % i, (2 * W,)),
                  ("gc_%d/kernel" % i, (1, E, 2 * W)), ("gc_%d/bias" % i, (2 * W,)),
                  ("res_%d/W" % i, (1, 1, W, W)), ("res_%d/biases" % i, (W,)),
                  ("skip_%d/W" % i, (1, 1, W, S)), ("skip_%d/biases" % i, (S,))]
    specs += [("out1/W", (1, 1, S, S)), ("out1/biases", (S,)), ("cond_map_out1/W", (1, 1, E, S)), ("cond_map_out1/biases", (S,)),
              ("gc_final/kernel", (1, E, S)), ("gc_final/bias", (S,)), ("logits/W", (1, 1, S, 256)), ("logits/biases", (256,))]
    return specs


def make_magenta_fastgen_weights(seed=4324, num_layers=MAGENTA_NUM_LAYERS, peaked=False):
    """seeded synthetic weights: kernels U(+-sqrt(3 / fan_in)), biases U(+-0.05) except the gc biases around their
    initial value 1.0 (config.py:33), codebook / speaker table as make_weights"""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in magenta_fastgen_specs(num_layers):
        if name == "embedding":
            a = rng.uniform(-0.130, 0.130, size=shape)
        elif name == "speaker_emb":
            a = rng.uniform(-0.332, 0.332, size=shape)
        elif name.endswith("/W") or name.endswith("/kernel"):
            lim = np.sqrt(3.0 / float(np.prod(shape[:-1])))
            a = rng.uniform(-lim, lim, size=shape)
            if peaked and name == "logits/W":
                a = a * 8.0
        elif name.startswith("gc_"):
            a = 1.0 + rng.uniform(-0.05, 0.05, size=shape)
        else:
            a = rng.uniform(-0.05, 0.05, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=F32)
    return out


class MagentaFastWavenet:
    """FastGenerationConfig.build as one step (config.py:36-138): queues of capacity `rate` zero-filled (masked.py:135-137),
    y = state_1 . W[0,0] + x . W[0,1] + b (masked.py:165-169), local condition cond_map_i(en) and global condition gc_i(gc)
    each WITH bias (the gc bias initialised to 1), gate = sigmoid(first half) * tanh(second half) (config.py:103)."""

    def __init__(self, weights, batch, speaker_idx, num_layers=MAGENTA_NUM_LAYERS):
        self.w, self.batch, self.L = weights, batch, num_layers
        self.gc = weights["speaker_emb"][np.asarray(speaker_idx, dtype=np.int64)]          # config.py:41-43
        self.init_ops()

    def init_ops(self):
        z1 = np.zeros((self.batch, 1), dtype=F32)
        zw = np.zeros((self.batch, MAGENTA_WIDTH), dtype=F32)
        self.q0 = deque([z1])                                                               # startconv: rate 1
        self.q = [deque([zw] * (2 ** (i % MAGENTA_NUM_STAGES))) for i in range(self.L)]     # config.py:75

    @staticmethod
    def _causal_linear(q, x, W, b):
        state_1 = q.popleft()                                                               # masked.py:137
        q.append(x)                                                                         # masked.py:138
        return ((state_1 @ W[0, 0] + x @ W[0, 1]) + b).astype(F32)                          # masked.py:165-169

    def step(self, input_t, encoding_t):
        w = self.w
        x = mu_law_encode(np.asarray(input_t, dtype=F32).reshape(self.batch, 1))            # config.py:47 (masked.py:31-35)
        en = np.asarray(encoding_t, dtype=F32)
        lin = lambda v, name: ((v @ w[name + "/W"][0, 0]) + w[name + "/biases"]).astype(F32)        # masked.py:163-172
        gcl = lambda scope: ((self.gc @ w[scope + "/kernel"][0]) + w[scope + "/bias"]).astype(F32)   # config.py:22-35
        l = self._causal_linear(self.q0, x, w["startconv/W"], w["startconv/biases"])
        s = lin(l, "skip_start")
        m = MAGENTA_WIDTH
        for i in range(self.L):
            d = self._causal_linear(self.q[i], l, w["dilatedconv_%d/W" % (i + 1)], w["dilatedconv_%d/biases" % (i + 1)])
            d = d + lin(en, "cond_map_%d" % (i + 1))                                        # config.py:93
            d = d + gcl("gc_%d" % (i + 1))                                                  # config.py:96-97
            d = (_sigmoid(d[:, :m]) * np.tanh(d[:, m:])).astype(F32)                        # config.py:103
            l = l + lin(d, "res_%d" % (i + 1))                                              # config.py:106
            s = s + lin(d, "skip_%d" % (i + 1))                                             # config.py:109
        s = np.maximum(s, 0)
        s = lin(s, "out1") + lin(en, "cond_map_out1")                                       # config.py:112-113
        s = s + gcl("gc_final")
        s = np.maximum(s, 0)
        logits = lin(s, "logits")
        return softmax(logits), logits


def magenta_generate(weights, encoding, speaker_idx, length, mode="greedy", uniforms=None, teacher=None,
                     num_layers=MAGENTA_NUM_LAYERS):
    """Magenta/generate.py:73-84 (the same loop as generate.py:103-113).  encoding [B,F,64] = e_k (config.py:242).
    Returns (audio, idx, logits, margins)."""
    B = encoding.shape[0]
    net = MagentaFastWavenet(weights, B, speaker_idx, num_layers)
    audio = np.zeros([B, 1], dtype=F32)
    to_write = np.zeros([B, length], dtype=F32)
    idx_out = np.zeros([B, length], dtype=np.int32)
    logits_out = np.zeros([B, length, 256], dtype=F32)
    margins = np.zeros([B, length], dtype=F32)
    ratio = length // encoding.shape[1]
    for i in range(length):
        probs, logits = net.step(audio, encoding[:, i // ratio])
        u = None if uniforms is None else uniforms[i]
        idx = decode_indices(probs, mode, u)
        margins[:, i] = draw_margin(probs, mode, u)
        decoded = mu_law_decode_np(idx.astype(F32), 256)
        to_write[:, i], idx_out[:, i], logits_out[:, i] = decoded, idx, logits
        audio = np.expand_dims(decoded, -1) if teacher is None else np.asarray(teacher[:, i:i + 1], dtype=F32)
    return to_write, idx_out, logits_out, margins


def synthetic_z_e(cfg, weights, B, F, seed=1235, kind="normal"):
    rng = np.random.default_rng(seed)
    E = weights["embedding/embedding"]
    if kind == "normal":
        return rng.standard_normal((B, F, cfg.D)).astype(F32)
    if kind == "near_code":
        j = rng.integers(0, cfg.K, size=(B, F))
        return (E[j] + F32(0.02) * rng.standard_normal((B, F, cfg.D)).astype(F32)).astype(F32)
    if kind == "scaled":   # N(0,1) scaled to the codebook's magnitude: many distinct codes get used
        return (F32(0.13) * rng.standard_normal((B, F, cfg.D))).astype(F32)
    raise ValueError(kind)


def synthetic_audio(B, T, seed=1237):
    """cfg 5: 0.5*sum of 3 sinusoids + 0.05*N, clipped."""
    rng = np.random.default_rng(seed)
    t = np.arange(T, dtype=np.float64)[None, :]
    f = rng.uniform(80.0, 1200.0, size=(B, 3, 1))
    ph = rng.uniform(0, 2 * np.pi, size=(B, 3, 1))
    x = 0.5 * np.sin(2 * np.pi * f * t[:, None, :] / 16000.0 + ph).sum(1) / 3.0
    x = x + 0.05 * rng.standard_normal((B, T))
    return np.clip(x, -1, 1).astype(F32)
