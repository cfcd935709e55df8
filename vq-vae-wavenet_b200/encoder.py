"""Mirror of Encoder/encoder.py as generate.py uses it (generate.py:65-69): an object with .build(x) -> z_e.
Encoder_64, Encoder_Magenta and Encoder_2019 all run on the device."""
import numpy as np


class Encoder_64:
    """6 x [Conv1D(768, k=5, s=2, same, relu) + BatchNorm] + Conv1D(latent_dim, k=1) + BatchNorm, hop 64
    (Encoder/encoder.py:8-26), executed by vqwn_encode_audio."""

    def __init__(self, latent_dim, engine=None):
        self.latent_dim = latent_dim
        self.engine = engine

    def build(self, net):
        if self.engine is None:
            raise RuntimeError("Encoder_64 needs the Engine that holds its weights (no CPU fallback)")
        return self.engine.encode_audio(np.asarray(net, dtype=np.float32))


class Encoder_Magenta:
    """shift_right + mu_law_encode, causal k=5 preprocess conv (128 filters), 6 x [1x1 stride-2 conv, causal dilated k=5
    gate / filter convs (dilations 1,2,4,8,16,16), tanh*sigmoid, 1x1 residual], 1x1 postprocess to latent_dim; hop 64
    (Encoder/encoder.py:29-64), executed by vqwn_encode_audio with vqwn_config.encoder = VQWN_ENCODER_MAGENTA."""

    def __init__(self, latent_dim, engine=None):
        self.latent_dim = latent_dim
        self.engine = engine
        self.args = {'dilation_rates': [1, 2, 4, 8, 16, 16], 'num_cycles': 1, 'num_cycle_layers': 6}    # encoder.py:33-36

    def build(self, net):
        if self.engine is None:
            raise RuntimeError("Encoder_Magenta needs the Engine that holds its weights (no CPU fallback)")
        return self.engine.encode_audio(np.asarray(net, dtype=np.float32))


class Encoder_2019:
    """MFCC front end (25 ms periodic-hann frames every 10 ms, |DFT|, 80 mel bands, log, 13 DCT coefficients:
    Encoder/encoder_ops.py:14-43), conv k3 + (conv k3 + skip), conv k4 stride 2, 2 x (conv k3 + skip), 4 x `relu + relu`
    (twice the conv output, reference quirk), 1x1 to latent_dim; hop 320 (Encoder/encoder.py:66-98), executed by
    vqwn_encode_audio with vqwn_config.encoder = VQWN_ENCODER_2019.  T must be a multiple of 320."""

    def __init__(self, latent_dim, engine=None):
        self.latent_dim = latent_dim
        self.engine = engine

    def build(self, net):
        if self.engine is None:
            raise RuntimeError("Encoder_2019 needs the Engine that holds its weights (no CPU fallback)")
        return self.engine.encode_audio(np.asarray(net, dtype=np.float32))
