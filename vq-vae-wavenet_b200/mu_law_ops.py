"""Host-side mu-law tables (mirrors mu_law_ops.py of the reference: encode :5-15, NumPy decode
:26-31).  The device never evaluates log1p/pow inside the sample loop: the network input after a
draw is always encode(decode(k)), k in 0..q, so two (q+1)-entry float32 tables built here with
NumPy -- the library the reference itself decodes with -- are uploaded once per handle."""
import numpy as np

_F = np.float32


def mu_law_decode_np(y, quantization_channels=256):
    mu = _F(quantization_channels - 1)
    centred = _F(2) * np.asarray(y, dtype=_F) / mu - _F(1)          # (0, mu) -> (-1, 1)
    magnitude = (np.power(_F(1) + mu, np.abs(centred), dtype=_F) - _F(1)) / mu
    return (np.sign(centred) * magnitude).astype(_F)


def mu_law_encode_np(x, quantization_channels=256, to_int=False):
    mu = _F(quantization_channels - 1)
    x = np.clip(np.asarray(x, dtype=_F), _F(-1), _F(1))
    y = (np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)).astype(_F)
    if to_int:
        y = ((y + _F(1)) / _F(2) * mu + _F(0.5)).astype(np.int32)
    return y


def decode_lut(quantization_channels=256):
    """index -> audio for 0..q inclusive (index q appears in sample mode when the float32 cdf
    ends below the draw; reference utils.py:20-25)."""
    return mu_law_decode_np(np.arange(quantization_channels + 1), quantization_channels)


def encode_lut(quantization_channels=256):
    """index -> next-step network input mu_law_encode(decode(k)) (generate.py:112-113, wavenet.py:113)."""
    return mu_law_encode_np(decode_lut(quantization_channels), quantization_channels)
