"""NumPy-facing wrapper of one vqwn handle (one GPU).  Thin: argument checking, dtype/contiguity
and error translation only; all arithmetic happens in libvqwn.so."""
import ctypes as C
import json
import os

import numpy as np

from . import _lib
from ._lib import VqwnError
from . import mu_law_ops

_MODES = {"greedy": _lib.MODE_GREEDY, "sample": _lib.MODE_SAMPLE}


class EngineConfig:
    """model_parameters.json + wavenet_parameters.json (same keys as the reference's files,
    generate.py:63-64, wavenet.py:10-21) + the speaker count generate.py:46-57 derives."""

    def __init__(self, model=None, wavenet=None, num_speakers=109):
        w = dict(quantization_channels=256, num_cycles=3, num_cycle_layers=10,
                 dilation_rates=[2 ** i for i in range(10)] * 3, kernel_size=3, dilation_filters=256,
                 skip_filters=512, residual_filters=256, preprocess=dict(kernel_size=32, filters=256))
        m = dict(encoder="64", use_vq=True, speaker_embedding=64, k=512, latent_dim=64, beta=0.25)
        if wavenet:
            w.update(wavenet)
        if model:
            m.update(model)
        assert len(w["dilation_rates"]) == w["num_cycles"] * w["num_cycle_layers"]   # wavenet.py:13
        self.model, self.wavenet, self.num_speakers = m, w, int(num_speakers)

    @classmethod
    def from_files(cls, model_path="model_parameters.json", num_speakers=109):
        with open(model_path) as f:
            m = json.load(f)
        wp = m["wavenet_parameters"]
        if not os.path.exists(wp):
            wp = os.path.join(os.path.dirname(os.path.abspath(model_path)), wp)
        with open(wp) as f:
            w = json.load(f)
        return cls(m, w, num_speakers)

    @property
    def cond_channels(self):
        return self.model["latent_dim"] + self.model["speaker_embedding"]

    @property
    def receptive_field(self):                                                      # wavenet.py:15-17
        w = self.wavenet
        return sum(w["dilation_rates"]) * (w["kernel_size"] - 1) + 1 + w["preprocess"]["kernel_size"] - 1

    def layer_scope(self, i):                                                       # wavenet.py:134-135
        n = self.wavenet["num_cycle_layers"]
        return "decoder/cycle_%d/layer_%d" % (1 + i // n, 1 + i % n)

    def to_c(self):
        w, m = self.wavenet, self.model
        c = _lib.Config()
        c.quantization_channels = w["quantization_channels"]
        c.num_layers = len(w["dilation_rates"])
        c.num_cycle_layers = w["num_cycle_layers"]
        for i, d in enumerate(w["dilation_rates"]):
            c.dilations[i] = int(d)
        c.kernel_size = w["kernel_size"]
        c.dilation_filters = w["dilation_filters"]
        c.skip_filters = w["skip_filters"]
        c.residual_filters = w["residual_filters"]
        c.pre_kernel_size = w["preprocess"]["kernel_size"]
        c.pre_filters = w["preprocess"]["filters"]
        c.k = m["k"]
        c.latent_dim = m["latent_dim"]
        c.speaker_dim = m["speaker_embedding"]
        c.num_speakers = self.num_speakers
        c.use_vq = 1 if m["use_vq"] else 0
        c.encoder = {"64": 64, "Magenta": 1, "2019": 2019}.get(str(m.get("encoder")), 0)   # VQWN_ENCODER_64 / _MAGENTA / _2019 / _NONE
        return c


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class Engine:
    def __init__(self, config=None, device=0, max_batch=64):
        self.lib = _lib.load_library()
        self.config = config or EngineConfig()
        if len(self.config.wavenet["dilation_rates"]) > _lib.MAX_LAYERS:
            raise ValueError("too many layers")
        self._h = C.c_void_p()
        cc = self.config.to_c()
        rc = self.lib.vqwn_create(C.byref(cc), int(device), int(max_batch), C.byref(self._h))
        if rc != 0:
            msg = self.lib.vqwn_last_error(None).decode()
            self._h = None
            self._raise(rc, msg)
        self.device, self.max_batch = device, max_batch
        self._B = None          # batch the library's queues are sized for (vqwn_reset); None: no run in progress
        self.q = self.config.wavenet["quantization_channels"]
        self.C = self.config.cond_channels
        self.D = self.config.model["latent_dim"]
        self.encoder_hop = 320 if str(self.config.model.get("encoder")) == "2019" else 64      # audio samples per z_e frame
        # NumPy-built mu-law tables (same expressions as the reference's NumPy decode)
        self.set_tensor("lut/mu_law_decode", mu_law_ops.decode_lut(self.q))
        self.set_tensor("lut/mu_law_encode", mu_law_ops.encode_lut(self.q))

    # ------------------------------------------------------------------ plumbing
    @staticmethod
    def _raise(rc, msg):
        if rc == _lib.ERR_NOTIMPL:
            raise NotImplementedError(msg)        # same exception type as generate.py:69 / utils.py:46
        if rc == _lib.ERR_INVALID:
            raise ValueError(msg)
        if rc == _lib.ERR_NOMEM:
            raise MemoryError(msg)
        raise VqwnError(rc, msg)

    def _ck(self, rc):
        if rc != 0:
            self._raise(rc, self.lib.vqwn_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.vqwn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.vqwn_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def set_precision(self, name):
        """"fp32": CUDA-core float32; "tc": split-bf16 tensor cores, float32-grade; "bf16": plain bf16 tensor cores.
        The dilation-queue layout differs between them: the step API needs reset() afterwards."""
        self._ck(self.lib.vqwn_set_precision(self._h, {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "tc": _lib.PREC_TC}[name]))
        self._B = None

    def set_reproducible(self, on=True):
        """"tc" precision only: fixed accumulation order of the four MMA-issuing warps - bit-reproducible results (a run
        equals its prefix, a shard the unsharded run, the step API the loop) at ~25 % lower throughput.  Default off: the
        last bits of the float32 accumulation vary from run to run (~1e-6 relative)."""
        self._ck(self.lib.vqwn_set_reproducible(self._h, 1 if on else 0))

    def set_stream_offset(self, offset):
        """global index of this engine's stream 0 in a sharded run (keys the seeded sample-mode generator)"""
        self._ck(self.lib.vqwn_set_stream_offset(self._h, int(offset)))

    def set_vq_kernel(self, name):
        self._ck(self.lib.vqwn_set_vq_kernel(self._h, {"auto": _lib.VQ_AUTO, "direct": _lib.VQ_DIRECT, "tensor": _lib.VQ_TENSOR, "tensor_bf16": _lib.VQ_TENSOR_BF16, "expanded": _lib.VQ_EXPANDED}[name]))

    def set_vq_output(self, name):
        """'straight_through': z_e + (e_k - z_e) (model.py:73); 'code': e_k itself (Magenta/config.py:242)"""
        self._ck(self.lib.vqwn_set_vq_output(self._h, {"straight_through": 0, "code": 1}[name]))

    # ------------------------------------------------------------------ weights
    def set_tensor(self, name, array):
        a = _f32(array)
        shape = (C.c_int64 * a.ndim)(*a.shape)
        self._ck(self.lib.vqwn_set_tensor(self._h, name.encode(), _ptr(a, C.c_float), shape, a.ndim))

    def set_weights(self, weights):
        for k, v in weights.items():
            self.set_tensor(k, v)

    def tensor_table(self):
        out = []
        for i in range(self.lib.vqwn_num_tensors(self._h)):
            name = C.create_string_buffer(256)
            shape = (C.c_int64 * 4)()
            nd, st = C.c_int(), C.c_int()
            self._ck(self.lib.vqwn_tensor_info(self._h, i, name, 256, shape, C.byref(nd), C.byref(st)))
            out.append((name.value.decode(), tuple(shape[:nd.value]), bool(st.value)))
        return out

    def get_tensor(self, name):
        shape = dict((n, s) for n, s, _ in self.tensor_table())[name]
        out = np.empty(shape, dtype=np.float32)
        self._ck(self.lib.vqwn_get_tensor(self._h, name.encode(), _ptr(out, C.c_float), out.size))
        return out

    # ------------------------------------------------------------------ encoder
    def encode_audio(self, x):
        """Encoder_*.build: x [B,T] or [B,T,1] float audio -> z_e [B,T/hop,latent_dim] (hop 64; Encoder_2019: 320)"""
        xx = _f32(x)
        if xx.ndim == 3:
            xx = np.ascontiguousarray(xx[:, :, 0])
        B, T = xx.shape
        z = np.empty((B, T // self.encoder_hop, self.D), dtype=np.float32)
        self._ck(self.lib.vqwn_encode_audio(self._h, _ptr(xx, C.c_float), B, T, _ptr(z, C.c_float)))
        return z

    # ------------------------------------------------------------------ VQ + conditioning
    def vq_lookup(self, z_e):
        z = _f32(z_e)
        assert z.shape[-1] == self.D
        n = z.size // self.D
        idx = np.empty(z.shape[:-1], dtype=np.int64)
        zq = np.empty_like(z)
        self._ck(self.lib.vqwn_vq_lookup(self._h, _ptr(z, C.c_float), n, _ptr(idx, C.c_int64), _ptr(zq, C.c_float)))
        return idx, zq

    def build_condition(self, z_q, speaker_idx):
        z = _f32(z_q)
        B, F, _ = z.shape
        s = np.ascontiguousarray(speaker_idx, dtype=np.int32).reshape(-1)
        if s.shape[0] != B:
            raise ValueError("speaker_idx must hold one index per stream (%d), got %d" % (B, s.shape[0]))
        cond = np.empty((B, F, self.C), dtype=np.float32)
        self._ck(self.lib.vqwn_build_condition(self._h, _ptr(z, C.c_float), _ptr(s, C.c_int32), B, F, _ptr(cond, C.c_float)))
        return cond

    def encode_condition(self, z_e, speaker_idx, want_indices=True):
        z = _f32(z_e)
        B, F, _ = z.shape
        s = np.ascontiguousarray(speaker_idx, dtype=np.int32).reshape(-1)
        if s.shape[0] != B:
            raise ValueError("speaker_idx must hold one index per stream (%d), got %d" % (B, s.shape[0]))
        cond = np.empty((B, F, self.C), dtype=np.float32)
        idx = np.empty((B, F), dtype=np.int64) if (want_indices and self.config.model["use_vq"]) else None
        self._ck(self.lib.vqwn_encode_condition(self._h, _ptr(z, C.c_float), _ptr(s, C.c_int32), B, F,
                                                _ptr(idx, C.c_int64) if idx is not None else None,
                                                _ptr(cond, C.c_float)))
        return idx, cond

    # ------------------------------------------------------------------ WaveNet
    def reset(self, batch):
        self._ck(self.lib.vqwn_reset(self._h, int(batch)))
        self._B = int(batch)

    def step(self, audio_t, cond_t):
        a = _f32(audio_t).reshape(-1)
        c = _f32(cond_t)
        B = a.shape[0]
        # the library reads and writes the batch of the last reset(), not the batch of these arrays
        if self._B is None:
            raise VqwnError(_lib.ERR_STATE, "step() needs reset(batch) first (no run in progress)")
        if B != self._B or c.shape != (B, self.C):
            raise ValueError("step(): audio_t must be [%d] and cond_t [%d,%d] (the batch given to reset), got %s and %s"
                             % (self._B, self._B, self.C, a.shape, c.shape))
        logits = np.empty((B, self.q), dtype=np.float32)
        probs = np.empty((B, self.q), dtype=np.float32)
        self._ck(self.lib.vqwn_step(self._h, _ptr(a, C.c_float), _ptr(c, C.c_float), _ptr(logits, C.c_float), _ptr(probs, C.c_float)))
        return probs, logits

    def decode(self, probs, mode="sample", uniforms=None):
        if mode not in _MODES:
            raise NotImplementedError("decode mode %s not implemented" % mode)      # utils.py:46
        p = _f32(probs)
        B = p.shape[0]
        u = None
        if mode == "sample":
            u = np.ascontiguousarray(np.random.rand(B) if uniforms is None else uniforms, dtype=np.float64)  # utils.py:22
        idx = np.empty(B, dtype=np.int32)
        audio = np.empty(B, dtype=np.float32)
        self._ck(self.lib.vqwn_decode(self._h, _ptr(p, C.c_float), B, _MODES[mode],
                                      _ptr(u, C.c_double) if u is not None else None, _ptr(idx, C.c_int32), _ptr(audio, C.c_float)))
        return idx, audio

    def generate(self, cond, length, mode="sample", uniforms=None, seed=0, out_audio=None, out_idx=None):
        if mode not in _MODES:
            raise NotImplementedError("decode mode %s not implemented" % mode)
        c = cond if (isinstance(cond, np.ndarray) and cond.dtype == np.float32 and cond.flags.c_contiguous) else _f32(cond)
        B, F, _ = c.shape
        u = None
        if mode == "sample" and uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64)
            assert u.shape == (length, B), "uniforms must be [T,B]"
        audio = out_audio if out_audio is not None else np.empty((B, length), dtype=np.float32)
        idx = out_idx if out_idx is not None else np.empty((B, length), dtype=np.int32)
        for name, arr, dt in (("out_audio", audio, np.float32), ("out_idx", idx, np.int32)):
            if not (isinstance(arr, np.ndarray) and arr.dtype == dt and arr.shape == (B, length) and arr.flags.c_contiguous
                    and arr.flags.writeable):
                raise ValueError("%s must be a writeable C-contiguous %s array of shape (%d, %d)" % (name, np.dtype(dt).name, B, length))
        if c.shape[2] != self.C:
            raise ValueError("cond must be [B,F,%d]" % self.C)
        self._B = B             # vqwn_generate resets the queues to this batch
        self._ck(self.lib.vqwn_generate(self._h, _ptr(c, C.c_float), B, F, int(length), _MODES[mode],
                                        _ptr(u, C.c_double) if u is not None else None, C.c_uint64(seed),
                                        _ptr(audio, C.c_float), _ptr(idx, C.c_int32)))
        return audio, idx

    def teacher_forced(self, x, cond):
        xx = _f32(x)
        c = _f32(cond)
        B, T = xx.shape
        F = c.shape[1]
        if c.shape[0] != B or c.shape[2] != self.C:
            raise ValueError("cond must be [%d,F,%d], got %s" % (B, self.C, c.shape))
        self._B = B
        logits = np.empty((B, T, self.q), dtype=np.float32)
        self._ck(self.lib.vqwn_teacher_forced(self._h, _ptr(xx, C.c_float), _ptr(c, C.c_float), B, F, T, _ptr(logits, C.c_float)))
        return logits

    # ------------------------------------------------------------------ resident variants (timing)
    def upload_condition(self, cond):
        c = _f32(cond)
        self._ck(self.lib.vqwn_upload_condition(self._h, _ptr(c, C.c_float), c.shape[0], c.shape[1]))

    def upload_uniforms(self, uniforms):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        self._ck(self.lib.vqwn_upload_uniforms(self._h, _ptr(u, C.c_double), u.shape[0], u.shape[1]))

    def generate_resident(self, B, F, T, mode="greedy", seed=0):
        self._B = int(B)
        self._ck(self.lib.vqwn_generate_resident(self._h, B, F, int(T), _MODES[mode], C.c_uint64(seed)))

    def download_output(self, B, T):
        audio = np.empty((B, T), dtype=np.float32)
        idx = np.empty((B, T), dtype=np.int32)
        self._ck(self.lib.vqwn_download_output(self._h, B, int(T), _ptr(audio, C.c_float), _ptr(idx, C.c_int32)))
        return audio, idx

    def vq_upload(self, z_e):
        z = _f32(z_e)
        self._ck(self.lib.vqwn_vq_upload(self._h, _ptr(z, C.c_float), z.size // self.D))

    def vq_resident(self, n):
        self._ck(self.lib.vqwn_vq_resident(self._h, int(n)))

    def vq_download(self, n):
        idx = np.empty(n, dtype=np.int64)
        zq = np.empty((n, self.D), dtype=np.float32)
        self._ck(self.lib.vqwn_vq_download(self._h, int(n), _ptr(idx, C.c_int64), _ptr(zq, C.c_float)))
        return idx, zq

    # ------------------------------------------------------------------ instrumentation
    @property
    def last_kernel_ms(self):
        return float(self.lib.vqwn_last_kernel_ms(self._h))

    @property
    def launch_count(self):
        return int(self.lib.vqwn_launch_count(self._h))

    @property
    def last_kernel_name(self):
        return self.lib.vqwn_last_kernel_name(self._h).decode()
