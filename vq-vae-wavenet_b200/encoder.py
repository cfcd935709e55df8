"""Mirror of Encoder/encoder.py as generate.py uses it (generate.py:65-69): an object with .build(x) -> z_e.
Only Encoder_64 runs on the device so far; the other two raise NotImplementedError."""
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
    def __init__(self, latent_dim, engine=None):
        raise NotImplementedError("encoder Magenta not implemented")     # SURVEY 8f #1 (after Encoder_64)


class Encoder_2019:
    def __init__(self, latent_dim, engine=None):
        raise NotImplementedError("encoder 2019 not implemented")
