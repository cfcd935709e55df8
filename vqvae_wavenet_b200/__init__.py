"""Importable alias of the package directory `vq-vae-wavenet_b200/` (a hyphen is not a valid
module name).  `import vqvae_wavenet_b200` executes vq-vae-wavenet_b200/__init__.py with this
module's __path__ pointing at that directory, so submodules resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vq-vae-wavenet_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
