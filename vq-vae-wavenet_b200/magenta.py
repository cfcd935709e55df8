"""Magenta/ fast generation on the same device path.

The reference's second model family (Magenta/config.py:18-138 `FastGenerationConfig`, Magenta/masked.py:133-174,
Magenta/generate.py:52-87) is the same computation as the default generator with a different topology:

    Magenta                                              default generator (wavenet.py:103-172)
    startconv: causal_linear k = 2, 1 -> 256             preprocess fast_conv1d k = 32: taps 31 (current) and 30, the rest zero
    50 layers, dilation 2^(i mod 10), k = 2              5 cycles x 10 layers, kernel_size 3 with a zero oldest tap
    cond_map_i(e_k) + bias  and  gc_i(speaker) + bias    ONE local_condition kernel [128, 512] on [e_k | speaker row]
                                                         (decoder_ops.py:39-43 concat), the two biases folded into gated/bias
    sigmoid(first half) * tanh(second half)              tanh(first) * sigmoid(second): output columns of the layer swapped
    out1 + cond_map_out1 + gc_final, logits              postprocess1 (+ its local_condition kernel), postprocess2
    speaker_emb [109, 64], encoding = e_k                speaker_embedding, vqwn_set_vq_output(VQWN_VQ_OUT_CODE)

so it runs through the same kernels after a pure re-arrangement of the checkpoint's tensors: no arithmetic is added or
removed except `+ 0 * x(t - 2d)` for the zero tap and the bias sums, which are done once here in float32.
`convert_weights` is that re-arrangement; `FastGenerationConfig` mirrors the reference class for Magenta/generate.py.
"""
import numpy as np

from .engine import Engine, EngineConfig

NUM_STAGES, NUM_LAYERS, FILTER_LENGTH = 10, 50, 2          # Magenta/config.py:5-7
WIDTH, SKIP_WIDTH, BOTTLENECK, K = 256, 512, 64, 512       # Magenta/config.py:8-9,15-16
NUM_SPEAKERS = 109                                         # Magenta/config.py:42
PREPROCESS_TAPS = 32


def wavenet_parameters(num_layers=NUM_LAYERS):
    """wavenet_parameters.json content that expresses FastGenerationConfig's topology"""
    assert num_layers % NUM_STAGES == 0
    return dict(quantization_channels=256, num_cycles=num_layers // NUM_STAGES, num_cycle_layers=NUM_STAGES,
                dilation_rates=[2 ** (i % NUM_STAGES) for i in range(num_layers)],       # Magenta/config.py:75
                kernel_size=3, dilation_filters=WIDTH, skip_filters=SKIP_WIDTH, residual_filters=WIDTH,
                preprocess=dict(kernel_size=PREPROCESS_TAPS, filters=WIDTH))


def engine_config(num_layers=NUM_LAYERS):
    return EngineConfig(model=dict(encoder="None", use_vq=True, speaker_embedding=BOTTLENECK, k=K, latent_dim=BOTTLENECK),
                        wavenet=wavenet_parameters(num_layers), num_speakers=NUM_SPEAKERS)


def variable_shapes(num_layers=NUM_LAYERS):
    """the variables of a Magenta checkpoint that fast generation reads (Magenta/config.py:40-130,229), by name"""
    W, S, E = WIDTH, SKIP_WIDTH, BOTTLENECK
    out = {"embedding": (K, E), "speaker_emb": (NUM_SPEAKERS, E), "startconv/W": (1, FILTER_LENGTH, 1, W), "startconv/biases": (W,),
           "skip_start/W": (1, 1, W, S), "skip_start/biases": (S,)}
    for i in range(1, num_layers + 1):
        out.update({"dilatedconv_%d/W" % i: (1, FILTER_LENGTH, W, 2 * W), "dilatedconv_%d/biases" % i: (2 * W,),
                    "cond_map_%d/W" % i: (1, 1, E, 2 * W), "cond_map_%d/biases" % i: (2 * W,),
                    "gc_%d/kernel" % i: (1, E, 2 * W), "gc_%d/bias" % i: (2 * W,),
                    "res_%d/W" % i: (1, 1, W, W), "res_%d/biases" % i: (W,),
                    "skip_%d/W" % i: (1, 1, W, S), "skip_%d/biases" % i: (S,)})
    out.update({"out1/W": (1, 1, S, S), "out1/biases": (S,), "cond_map_out1/W": (1, 1, E, S), "cond_map_out1/biases": (S,),
                "gc_final/kernel": (1, E, S), "gc_final/bias": (S,), "logits/W": (1, 1, S, 256), "logits/biases": (256,)})
    return out


def convert_weights(mw, num_layers=NUM_LAYERS):
    """Magenta variables -> the tensors vqwn_set_tensor expects (reference variable names of the default generator)"""
    f32 = np.float32
    shapes = variable_shapes(num_layers)
    for name, shp in shapes.items():
        if name not in mw:
            raise KeyError("Magenta checkpoint lacks %s" % name)
        if tuple(np.shape(mw[name])) != shp:
            raise ValueError("%s: expected shape %s, got %s" % (name, shp, np.shape(mw[name])))
    g = lambda n: np.asarray(mw[n], dtype=f32)
    # gate halves: Magenta sigmoid(first) * tanh(second) (config.py:103) -> tanh(first) * sigmoid(second) (wavenet_ops.py:235-236)
    perm = np.concatenate([np.arange(WIDTH, 2 * WIDTH), np.arange(0, WIDTH)])
    out = {"embedding/embedding": g("embedding"), "speaker_embedding": g("speaker_emb")}
    pre = np.zeros((PREPROCESS_TAPS, 1, WIDTH), dtype=f32)
    pre[PREPROCESS_TAPS - 1] = g("startconv/W")[0, 1]          # current sample (masked.py:166: w_x = w[0, 1])
    pre[PREPROCESS_TAPS - 2] = g("startconv/W")[0, 0]          # queue of rate 1 (masked.py:165: w_q_1 = w[0, 0])
    out["decoder/preprocess/kernel"] = pre
    out["decoder/preprocess/bias"] = g("startconv/biases")
    out["decoder/skip/kernel"] = g("skip_start/W")[0]
    out["decoder/skip/bias"] = g("skip_start/biases")
    for i in range(num_layers):
        sc = "decoder/cycle_%d/layer_%d" % (1 + i // NUM_STAGES, 1 + i % NUM_STAGES)
        n = i + 1
        w = g("dilatedconv_%d/W" % n)[0]                        # [2, 256, 512]: [0] multiplies x(t - d), [1] x(t)
        k3 = np.zeros((3, WIDTH, 2 * WIDTH), dtype=f32)         # fast_conv1d: kernel[2] current, [1] t - d, [0] t - 2d
        k3[2], k3[1] = w[1][:, perm], w[0][:, perm]
        out[sc + "/gated/kernel"] = k3
        out[sc + "/gated/bias"] = ((g("dilatedconv_%d/biases" % n) + g("cond_map_%d/biases" % n)) + g("gc_%d/bias" % n))[perm]
        out[sc + "/gated/local_condition/kernel"] = np.concatenate([g("cond_map_%d/W" % n)[0, 0], g("gc_%d/kernel" % n)[0]], 0)[None][:, :, perm]
        out[sc + "/skip/kernel"] = g("skip_%d/W" % n)[0]
        out[sc + "/skip/bias"] = g("skip_%d/biases" % n)
        out[sc + "/residual/kernel"] = g("res_%d/W" % n)[0]
        out[sc + "/residual/bias"] = g("res_%d/biases" % n)
    out["decoder/postprocess1/kernel"] = g("out1/W")[0]
    out["decoder/postprocess1/bias"] = (g("out1/biases") + g("cond_map_out1/biases")) + g("gc_final/bias")
    out["decoder/postprocess1/local_condition/kernel"] = np.concatenate([g("cond_map_out1/W")[0, 0], g("gc_final/kernel")[0]], 0)[None]
    out["decoder/postprocess2/kernel"] = g("logits/W")[0]
    out["decoder/postprocess2/bias"] = g("logits/biases")
    return {k: np.ascontiguousarray(v, dtype=f32) for k, v in out.items()}


class FastGenerationConfig:
    """Mirror of Magenta/config.py:18-138 as Magenta/generate.py:60-84 uses it: `build` takes the speaker one-hots and
    returns the handles of the loop; `generate` is that loop (init_ops, then `length` steps of predictions + push_ops +
    decode) as ONE persistent kernel launch."""

    def __init__(self, batch_size=1, device=0, num_layers=NUM_LAYERS, precision="fp32"):
        self.batch_size = batch_size
        self.num_layers = num_layers
        self.engine = Engine(engine_config(num_layers), device=device, max_batch=batch_size)
        self.engine.set_precision(precision)
        self.engine.set_vq_output("code")                     # Magenta/config.py:242: the decoder sees e_k
        self.speaker_idx = None

    def restore(self, magenta_variables):
        """saver.restore (Magenta/generate.py:65): arrays keyed by the checkpoint's variable names"""
        self.engine.set_weights(convert_weights(magenta_variables, self.num_layers))

    def build(self, gc):
        """gc [B, 109] one-hot (Magenta/generate.py:52-56); config.py:40-43 takes its argmax into speaker_emb"""
        gc = np.asarray(gc)
        if gc.shape[0] != self.batch_size:
            raise ValueError("gc must hold one row per stream")
        self.speaker_idx = np.argmax(gc, axis=-1).astype(np.int32).reshape(-1)
        return self

    @property
    def speaker_emb(self):                                    # Magenta/generate.py:70
        return self.engine.get_tensor("speaker_embedding")

    @property
    def embedding(self):                                      # Magenta/generate.py:67
        return self.engine.get_tensor("embedding/embedding")

    def quantise(self, z_e):
        """Magenta/config.py:233-242: direct-form distance, first-index argmin, e_k -> (indices, condition [B,F,128])"""
        return self.engine.encode_condition(np.asarray(z_e, dtype=np.float32), self.speaker_idx)

    def condition_from_codes(self, e_k):
        """encoding = e_k [B,F,64] (already quantised) -> condition rows [e_k | speaker row]"""
        return self.engine.build_condition(np.asarray(e_k, dtype=np.float32), self.speaker_idx)

    def generate(self, cond, length, mode="sample", uniforms=None, seed=0):
        """Magenta/generate.py:73-84 -> (audio [B,length] float32, indices)"""
        return self.engine.generate(cond, length, mode=mode, uniforms=uniforms, seed=seed)

    def teacher_forced(self, x, cond):
        return self.engine.teacher_forced(x, cond)

    def close(self):
        self.engine.close()
