"""B200-native VQ-VAE-WaveNet inference hot path: VQ lookup + WaveNet fast generation.

Host side (Python, mirrors the reference's generate.py-facing interface) over a C-ABI CUDA
library (include/vqwn.h, csrc/).  There is no CPU fallback: every compute entry point fails
loudly when libvqwn.so is missing or no sm_100 device is present.
"""
from ._lib import load_library, library_path, VqwnError  # noqa: F401
from .engine import Engine, EngineConfig  # noqa: F401
from .model import VQVAE  # noqa: F401
from .decoder import WavenetDecoder  # noqa: F401
from .wavenet import Wavenet  # noqa: F401
from .encoder import Encoder_64, Encoder_Magenta, Encoder_2019  # noqa: F401
from . import mu_law_ops, utils  # noqa: F401

__all__ = ["load_library", "library_path", "VqwnError", "Engine", "EngineConfig", "VQVAE",
           "WavenetDecoder", "Wavenet", "mu_law_ops", "utils"]
