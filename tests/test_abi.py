"""The C-ABI library loads and exports exactly the symbols include/vqwn.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vqwn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vqwn_[a-z0-9_]+)\s*\(", text)))


def test_build_and_symbols():
    import __graft_entry__ as g
    g.build()
    from vqvae_wavenet_b200 import _lib
    lib = _lib.load_library()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "library does not export %s" % s
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes binding and header disagree"
    assert b"sm_100a" in lib.vqwn_version()


def test_config_struct_matches_header():
    from vqvae_wavenet_b200 import _lib
    text = open(os.path.join(ROOT, "include", "vqwn.h")).read()
    body = re.search(r"typedef struct vqwn_config \{(.*?)\} vqwn_config;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"int32_t\s+([a-z_]+)", body)
    assert fields == [f[0] for f in _lib.Config._fields_]
    assert ctypes.sizeof(_lib.Config) == 4 * (len(fields) - 1 + 64)


def test_no_cpu_fallback_without_device():
    """On a box without a GPU, creating an engine must fail loudly (never a silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vqvae_wavenet_b200 import Engine, VqwnError
    with pytest.raises(VqwnError):
        Engine()


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    from vqvae_wavenet_b200 import _lib
    monkeypatch.setenv("VQWN_LIBRARY", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load_library()
