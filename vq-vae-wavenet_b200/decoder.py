"""Mirror of Decoder/decoder.py:6-62 (generator side)."""
import numpy as np

from .wavenet import Wavenet


class WavenetDecoder:
    def __init__(self, args_file):
        self.args_file = args_file
        self.wavenet = None

    def build_generator(self, engine, z_e, speaker_idx):
        """decoder.py:40-62: concat(local_condition, tile(global_condition)) and bind the WaveNet
        generator; returns the full conditioning tensor [B,F,C] (what generate.py:92 evaluates as
        model.encoding) plus the VQ indices.  VQ + gather + concat run fused on the device."""
        idx, cond = engine.encode_condition(z_e, speaker_idx)
        self.wavenet = Wavenet(self.args_file)
        self.wavenet.build_generator(engine, batch_size=np.asarray(z_e).shape[0])
        return idx, cond
