"""Seeded synthetic weights and inputs for benchmarks / smoke runs: there is no
network for checkpoints or datasets.  Weights are keyed by the reference's variable names."""
import numpy as np


def tensor_specs(config):
    """(tf_variable_name, shape) in the order the reference's graph creates them."""
    w, m = config.wavenet, config.model
    R, G, S, q = w["residual_filters"], w["dilation_filters"], w["skip_filters"], w["quantization_channels"]
    C = config.cond_channels
    pk, pf = w["preprocess"]["kernel_size"], w["preprocess"]["filters"]
    specs = [("embedding/embedding", (m["k"], m["latent_dim"])),
             ("speaker_embedding", (config.num_speakers, m["speaker_embedding"])),
             ("decoder/preprocess/kernel", (pk, 1, pf)), ("decoder/preprocess/bias", (pf,)),
             ("decoder/skip/kernel", (1, pf, S)), ("decoder/skip/bias", (S,))]
    for i in range(len(w["dilation_rates"])):
        s = config.layer_scope(i)
        specs += [(s + "/gated/kernel", (w["kernel_size"], R, 2 * G)), (s + "/gated/bias", (2 * G,)),
                  (s + "/gated/local_condition/kernel", (1, C, 2 * G)),
                  (s + "/skip/kernel", (1, G, S)), (s + "/skip/bias", (S,)),
                  (s + "/residual/kernel", (1, G, R)), (s + "/residual/bias", (R,))]
    specs += [("decoder/postprocess1/kernel", (1, S, S)), ("decoder/postprocess1/bias", (S,)),
              ("decoder/postprocess1/local_condition/kernel", (1, C, S)),
              ("decoder/postprocess2/kernel", (1, S, q)), ("decoder/postprocess2/bias", (q,))]
    return specs


def make_weights(config, seed=1234, peaked=False):
    """kernels U(+-sqrt(3/fan_in)) (uniform_unit_scaling, wavenet_ops.py:69), biases U(+-0.05),
    codebook factor 1.7 (model.py:49), speaker table factor 2 (model.py:26); peaked: postprocess2 x8."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in tensor_specs(config):
        if name.endswith("bias"):
            a = rng.uniform(-0.05, 0.05, size=shape)
        else:
            lim = np.sqrt(3.0 / int(np.prod(shape[:-1])))
            lim *= {"embedding/embedding": 1.7, "speaker_embedding": 2.0}.get(name, 1.0)
            a = rng.uniform(-lim, lim, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    if peaked:
        out["decoder/postprocess2/kernel"] = (out["decoder/postprocess2/kernel"] * np.float32(8)).astype(np.float32)
    return out


def synthetic_z_e(config, B, F, seed=1235, scale=0.13):
    """encoder-output stand-in: N(0,1) scaled to the codebook's magnitude"""
    rng = np.random.default_rng(seed)
    return (np.float32(scale) * rng.standard_normal((B, F, config.model["latent_dim"]))).astype(np.float32)


def encoder64_specs(config):
    """keras auto-names inside variable_scope('encoder'): conv1d[_i] / batch_normalization[_i]"""
    specs, cin = [], 1
    for i in range(7):
        sfx = "" if i == 0 else "_%d" % i
        cout = 768 if i < 6 else config.model["latent_dim"]
        k = 5 if i < 6 else 1
        specs += [("encoder/conv1d%s/kernel" % sfx, (k, cin, cout)), ("encoder/conv1d%s/bias" % sfx, (cout,)),
                  ("encoder/batch_normalization%s/gamma" % sfx, (cout,)), ("encoder/batch_normalization%s/beta" % sfx, (cout,)),
                  ("encoder/batch_normalization%s/moving_mean" % sfx, (cout,)),
                  ("encoder/batch_normalization%s/moving_variance" % sfx, (cout,))]
        cin = cout
    return specs


def make_encoder64_weights(config, seed=4321):
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in encoder64_specs(config):
        if name.endswith("kernel"):
            lim = np.sqrt(6.0 / (shape[0] * shape[1] + shape[0] * shape[2]))
            a = rng.uniform(-lim, lim, size=shape)
        elif name.endswith("moving_variance"):
            a = rng.uniform(0.5, 1.5, size=shape)
        elif name.endswith("gamma"):
            a = rng.uniform(0.8, 1.2, size=shape)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out


MAGENTA_DILATIONS = [1, 2, 4, 8, 16, 16]        # Encoder/encoder.py:34


def encoder_magenta_specs(config):
    """conv1d_v2 variables of Encoder_Magenta under variable_scope('encoder') (Encoder/encoder.py:37-64)"""
    C, k = 128, 5
    specs = [("encoder/preprocess/kernel", (k, 1, C)), ("encoder/preprocess/bias", (C,))]
    for i in range(len(MAGENTA_DILATIONS)):
        sc = "encoder/cycle_%d/layer_%d" % (1 + i // 6, 1 + i % 6)
        specs += [(sc + "/dilated/kernel", (1, C, C)), (sc + "/dilated/bias", (C,)),
                  (sc + "/gate/kernel", (k, C, C)), (sc + "/gate/bias", (C,)),
                  (sc + "/filter/kernel", (k, C, C)), (sc + "/filter/bias", (C,)),
                  (sc + "/residual/kernel", (1, C, C)), (sc + "/residual/bias", (C,))]
    specs += [("encoder/postprocess/kernel", (1, C, config.model["latent_dim"])),
              ("encoder/postprocess/bias", (config.model["latent_dim"],))]
    return specs


def make_encoder_magenta_weights(config, seed=4322):
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in encoder_magenta_specs(config):
        if name.endswith("kernel"):
            lim = np.sqrt(3.0 / (shape[0] * shape[1]))
            a = rng.uniform(-lim, lim, size=shape)
        else:
            a = rng.uniform(-0.05, 0.05, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out


def encoder2019_specs(config):
    """keras Conv1D variables of Encoder_2019 under variable_scope('encoder') (Encoder/encoder.py:75-96): conv1d ... conv1d_9"""
    D = config.model["latent_dim"]
    shapes = [(3, 13, 768), (3, 768, 768), (4, 768, 768)] + [(3, 768, 768)] * 6 + [(1, 768, D)]
    specs = []
    for i, shp in enumerate(shapes):
        sfx = "" if i == 0 else "_%d" % i
        specs += [("encoder/conv1d%s/kernel" % sfx, shp), ("encoder/conv1d%s/bias" % sfx, (shp[2],))]
    return specs


def make_encoder2019_weights(config, seed=4323):
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in encoder2019_specs(config):
        if name.endswith("kernel"):
            lim = np.sqrt(6.0 / (shape[0] * shape[1] + shape[0] * shape[2]))
            a = rng.uniform(-lim, lim, size=shape)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out
