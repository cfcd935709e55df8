"""A minimal deferred-execution NumPy stand-in for the TensorFlow-1.x API surface that the
reference's inference path touches.  TEST INFRASTRUCTURE ONLY (used by make_ref_golden.py).

Purpose: TensorFlow 1.x cannot be installed in this image, so the reference's graph code cannot
run as shipped.  This module is installed into ``sys.modules['tensorflow']`` so that the
UNMODIFIED reference files (model.py, Decoder/*, Decoder/WaveNet/*, mu_law_ops.py, utils.py,
Encoder/encoder.py under /root/reference) import and build their graphs; ``Session.run``
then evaluates them with NumPy (torch only for conv2d, as an independent library witness).
The algorithm (which ops, in which order, with which variable names, queue chaining, tap
order, slicing) is therefore the reference's own code; only the leaf operators are supplied
here, each with the semantics of the TF op it replaces (stated per function).

Graph model: a Tensor is a node (function + inputs) with a static shape; Session.run
memoises node values per call.  FIFOQueue.dequeue pops when first evaluated in a run;
enqueue ops are committed at the end of the run (TF blocks an enqueue on a full queue until
the same run's dequeue frees the slot, which is the same order).
"""
from __future__ import annotations

import contextlib
import sys
import types
from collections import deque

import numpy as np

__version__ = "1.14.0-numpy-shim"
float32 = np.float32
int32 = np.int32
int64 = np.int64


# ----------------------------------------------------------------------------------------
# graph state
# ----------------------------------------------------------------------------------------
class _Graph:
    def __init__(self):
        self.scope = []
        self.variables = {}          # full name -> Variable (creation order preserved)
        self.trainable = []
        self.initial_values = {}     # full name -> ndarray supplied by the harness


_G = _Graph()


def reset_default_graph():
    global _G
    keep = _G.initial_values
    _G = _Graph()
    _G.initial_values = keep


def set_variable_values(values):
    """harness hook: values by full variable name; get_variable fails on a name not present"""
    _G.initial_values = dict(values)


def created_variables():
    return [(n, tuple(v.shape.as_list())) for n, v in _G.variables.items()]


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    parts = [p for p in name.split("/") if p]
    _G.scope.extend(parts)
    try:
        yield
    finally:
        del _G.scope[len(_G.scope) - len(parts):]


def _full_name(name):
    return "/".join(_G.scope + [name])


# ----------------------------------------------------------------------------------------
# tensors
# ----------------------------------------------------------------------------------------
class TensorShape:
    def __init__(self, dims):
        self.dims = [None if d is None else int(d) for d in dims]

    def as_list(self):
        return list(self.dims)

    def __getitem__(self, i):
        r = self.dims[i]
        return TensorShape(r) if isinstance(i, slice) else r

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(self.dims)

    def concatenate(self, other):
        return TensorShape(self.dims + list(other))

    def __repr__(self):
        return "(%s)" % ", ".join(str(d) for d in self.dims)


class Operation:
    """a node evaluated for its side effect (enqueue)"""

    def __init__(self, fn, inputs):
        self.fn, self.inputs = fn, list(inputs)


class Tensor:
    def __init__(self, fn, inputs, shape, dtype=np.float32, name=None):
        self.fn = fn
        self.inputs = list(inputs)
        self._shape = TensorShape(shape)
        self.dtype = dtype
        self.name = name

    # -- static shape API used by the reference
    @property
    def shape(self):
        return self._shape

    def get_shape(self):
        return self._shape

    def set_shape(self, shape):
        shape = list(shape)
        assert len(shape) == len(self._shape)
        for a, b in zip(shape, self._shape):
            assert a is None or b is None or int(a) == b, (shape, self._shape)

    # -- operators
    def __add__(self, o):
        return _binary(np.add, self, o)

    __radd__ = __add__

    def __sub__(self, o):
        return _binary(np.subtract, self, o)

    def __rsub__(self, o):
        return _binary(np.subtract, o, self)

    def __mul__(self, o):
        return _binary(np.multiply, self, o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return _binary(np.true_divide, self, o)

    def __rtruediv__(self, o):
        return _binary(np.true_divide, o, self)

    def __pow__(self, o):
        return _binary(np.power, self, o)

    def __rpow__(self, o):
        return _binary(np.power, o, self)

    def __neg__(self):
        return _unary(np.negative, self)

    def __lt__(self, o):
        return _binary(np.less, self, o, dtype=np.bool_)

    def __ge__(self, o):
        return _binary(np.greater_equal, self, o, dtype=np.bool_)

    def __getitem__(self, idx):
        probe = np.broadcast_to(np.zeros((), np.int8), tuple(self._shape.dims))[idx]
        return Tensor(lambda x: x[idx], [self], probe.shape, self.dtype)

    def __repr__(self):
        return "<shim Tensor %s %s>" % (self.name or "", self._shape)


class Variable(Tensor):
    def __init__(self, name, shape, value):
        super().__init__(None, [], shape, np.float32, name)
        self.value = value
        self.op = types.SimpleNamespace(name=name)


def _wrap(x, like_dtype=None):
    if isinstance(x, Tensor):
        return x
    a = np.asarray(x)
    if a.dtype == np.float64 and (like_dtype is None or like_dtype == np.float32):
        a = a.astype(np.float32)          # python float literals are weakly typed in TF
    elif a.dtype == np.int64 and like_dtype is not None:
        a = a.astype(like_dtype)
    return Tensor(lambda: a, [], a.shape, a.dtype.type)


def _bshape(a, b):
    return np.broadcast_shapes(tuple(a), tuple(b))


def _binary(f, a, b, dtype=None):
    ad = a.dtype if isinstance(a, Tensor) else None
    bd = b.dtype if isinstance(b, Tensor) else None
    a, b = _wrap(a, bd), _wrap(b, ad)
    out_dtype = dtype or a.dtype

    def run(x, y):
        r = f(x, y)
        return r if dtype is not None else r.astype(out_dtype, copy=False)
    return Tensor(run, [a, b], _bshape(a._shape.dims, b._shape.dims), out_dtype)


def _unary(f, a, dtype=None):
    a = _wrap(a)
    return Tensor(lambda x: f(x).astype(dtype or a.dtype, copy=False), [a], a._shape.dims, dtype or a.dtype)


def constant(value, dtype=None, shape=None, name=None):
    a = np.asarray(value)
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64:
        a = a.astype(np.float32)
    return Tensor(lambda: a, [], a.shape, a.dtype.type)


class _Placeholder(Tensor):
    pass


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(None, [], shape, dtype, name)


def zeros(shape, dtype=np.float32):
    a = np.zeros(shape, dtype)
    return Tensor(lambda: a, [], a.shape, dtype)


# ----------------------------------------------------------------------------------------
# variables / initializers (initial values come from the harness, by full name)
# ----------------------------------------------------------------------------------------
def uniform_unit_scaling_initializer(factor=1.0):
    return ("uniform_unit_scaling", factor)


def constant_initializer(v=0.0):
    return ("constant", v)


def get_variable(name, shape=None, dtype=np.float32, initializer=None, regularizer=None, trainable=True):
    full = _full_name(name)
    if full in _G.variables:
        raise ValueError("Variable %s already exists" % full)      # TF1 default (reuse=False)
    if full not in _G.initial_values:
        raise KeyError("the harness supplied no value for variable %r" % full)
    val = np.asarray(_G.initial_values[full])
    shape = [int(s) for s in shape]
    if list(val.shape) != shape:
        raise ValueError("variable %s: reference asks for shape %s, harness supplied %s" % (full, shape, val.shape))
    v = Variable(full, shape, val.astype(np.float32))
    _G.variables[full] = v
    if trainable:
        _G.trainable.append(v)
    return v


def trainable_variables():
    return list(_G.trainable)


# ----------------------------------------------------------------------------------------
# ops (each states the TF op it stands for)
# ----------------------------------------------------------------------------------------
def matmul(a, b):
    """tf.matmul on rank-2 float32 operands"""
    a, b = _wrap(a), _wrap(b)
    assert len(a._shape) == 2 and len(b._shape) == 2 and a._shape[1] == b._shape[0], (a._shape, b._shape)
    return Tensor(lambda x, y: np.matmul(x, y), [a, b], [a._shape[0], b._shape[1]])


def cast(x, dtype):
    """tf.cast: float -> int truncates toward zero"""
    x = _wrap(x)
    return Tensor(lambda v: np.asarray(v).astype(dtype), [x], x._shape.dims, dtype)


def clip_by_value(x, lo, hi):
    return _unary(lambda v: np.clip(v, np.float32(lo), np.float32(hi)), x)


def sign(x):
    return _unary(np.sign, x)


def abs(x):  # noqa: A001 - mirrors tf.abs
    x = _wrap(x)
    return _unary(np.abs, x, dtype=np.float32 if x.dtype == np.complex64 else None)      # tf.abs(complex64) is float32


def reshape(x, shape):
    x = _wrap(x)
    shape = [int(s) for s in shape]
    n = int(np.prod(x._shape.dims))
    if -1 in shape:
        known = -int(np.prod(shape))
        shape[shape.index(-1)] = n // known
    assert int(np.prod(shape)) == n, (x._shape, shape)
    return Tensor(lambda v: np.reshape(v, shape), [x], shape, x.dtype)


def pad(x, paddings):
    """tf.pad, mode CONSTANT (zeros)"""
    x = _wrap(x)
    paddings = [[int(a), int(b)] for a, b in paddings]
    shp = [d + a + b for d, (a, b) in zip(x._shape.dims, paddings)]
    return Tensor(lambda v: np.pad(v, paddings), [x], shp, x.dtype)


def expand_dims(x, axis):
    x = _wrap(x)
    shp = list(np.expand_dims(np.broadcast_to(np.int8(0), tuple(x._shape.dims)), axis).shape)
    return Tensor(lambda v: np.expand_dims(v, axis), [x], shp, x.dtype)


def squeeze(x, axis=None):
    x = _wrap(x)
    shp = list(np.squeeze(np.broadcast_to(np.int8(0), tuple(x._shape.dims)), axis).shape)
    return Tensor(lambda v: np.squeeze(v, axis), [x], shp, x.dtype)


def reduce_sum(x, axis=None):
    x = _wrap(x)
    shp = list(np.sum(np.broadcast_to(np.int8(0), tuple(x._shape.dims)), axis=axis).shape)
    return Tensor(lambda v: np.sum(v, axis=axis, dtype=v.dtype), [x], shp, x.dtype)


def argmin(x, axis=None):
    """tf.argmin -> int64; lowest index on ties (the de-facto kernel behaviour)"""
    x = _wrap(x)
    shp = list(np.argmin(np.broadcast_to(np.int8(0), tuple(x._shape.dims)), axis=axis).shape)
    return Tensor(lambda v: np.argmin(v, axis=axis).astype(np.int64), [x], shp, np.int64)


def argmax(x, axis=None):
    x = _wrap(x)
    shp = list(np.argmax(np.broadcast_to(np.int8(0), tuple(x._shape.dims)), axis=axis).shape)
    return Tensor(lambda v: np.argmax(v, axis=axis).astype(np.int64), [x], shp, np.int64)


def stop_gradient(x):
    return _unary(lambda v: v, x)


def tile(x, multiples):
    x = _wrap(x)
    mult = [int(_static(m)) for m in multiples]
    shp = [d * m for d, m in zip(x._shape.dims, mult)]
    return Tensor(lambda v: np.tile(v, mult), [x], shp, x.dtype)


def concat(values, axis):
    values = [_wrap(v) for v in values]
    shp = list(values[0]._shape.dims)
    ax = axis % len(shp)
    shp[ax] = sum(v._shape.dims[ax] for v in values)
    return Tensor(lambda *vs: np.concatenate(vs, axis=axis), values, shp, values[0].dtype)


def _static(x):
    if isinstance(x, Tensor):
        assert not x.inputs or all(not isinstance(i, _Placeholder) for i in x.inputs)
        return Session().run(x)
    return x


def shape(x):  # noqa: A001 - mirrors tf.shape
    x = _wrap(x)
    a = np.asarray(x._shape.dims, dtype=np.int32)
    return Tensor(lambda: a, [], a.shape, np.int32)


def transpose(x, perm):
    x = _wrap(x)
    return Tensor(lambda v: np.transpose(v, perm), [x], [x._shape.dims[p] for p in perm], x.dtype)


def tensordot(a, b, axes):
    a, b = _wrap(a), _wrap(b)
    assert axes == 1
    return Tensor(lambda x, y: np.tensordot(x, y, 1).astype(np.float32), [a, b],
                  list(a._shape.dims[:-1]) + list(b._shape.dims[1:]))


def _conv2d(x, kernel, strides=None, padding="VALID", dilations=None):
    """tf.nn.conv2d, NHWC input, HWIO filter, cross-correlation (no filter flip), VALID only here;
    evaluated with torch.nn.functional.conv2d (independent library implementation)."""
    x, kernel = _wrap(x), _wrap(kernel)
    assert padding == "VALID"
    strides = strides or [1, 1, 1, 1]
    dilations = dilations or [1, 1, 1, 1]
    n, h, w, c = x._shape.dims
    fh, fw, ci, co = kernel._shape.dims
    assert ci == c
    oh = (h - (fh - 1) * dilations[1] - 1) // strides[1] + 1
    ow = (w - (fw - 1) * dilations[2] - 1) // strides[2] + 1

    def run(v, k):
        import torch
        with torch.no_grad():
            r = torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(v)).permute(0, 3, 1, 2),
                                           torch.from_numpy(np.ascontiguousarray(k)).permute(3, 2, 0, 1),
                                           stride=(strides[1], strides[2]), dilation=(dilations[1], dilations[2]))
        return np.ascontiguousarray(r.permute(0, 2, 3, 1).numpy())
    return Tensor(run, [x, kernel], [n, oh, ow, co])


def _sigmoid(v):
    return (np.float32(1) / (np.float32(1) + np.exp(-v))).astype(np.float32)


def _softmax(v):
    m = np.max(v, axis=-1, keepdims=True)
    e = np.exp((v - m).astype(np.float32)).astype(np.float32)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)


def _embedding_lookup(params, ids):
    params, ids = _wrap(params), _wrap(ids)
    return Tensor(lambda p, i: p[i], [params, ids], list(ids._shape.dims) + list(params._shape.dims[1:]))


def _moments(x, axes):
    x = _wrap(x)
    m = Tensor(lambda v: v.mean(tuple(axes)), [x], list(np.mean(np.broadcast_to(np.int8(0), tuple(x._shape.dims)), tuple(axes)).shape))
    return m, m


def sigmoid(x):
    """tf.sigmoid (Magenta/config.py:103)"""
    return _unary(_sigmoid, x)


def tanh(x):
    """tf.tanh (Magenta/config.py:103)"""
    return _unary(np.tanh, x)


nn = types.SimpleNamespace(
    tanh=lambda x: _unary(np.tanh, x),
    sigmoid=lambda x: _unary(_sigmoid, x),
    relu=lambda x: _unary(lambda v: np.maximum(v, np.float32(0)), x),
    softmax=lambda x, name=None: _unary(_softmax, x),
    bias_add=lambda value, bias: value + bias,            # tf.nn.bias_add on [batch, channels] (Magenta/masked.py:159,169,172)
    conv2d=_conv2d,
    embedding_lookup=_embedding_lookup,
    moments=_moments,
)
math = types.SimpleNamespace(log1p=lambda x: _unary(np.log1p, x), log=lambda x: _unary(np.log, x))
summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None,
                                merge_all=lambda: None)
logging = types.SimpleNamespace(set_verbosity=lambda *a: None, ERROR=0)
compat = types.SimpleNamespace(v1=types.SimpleNamespace(logging=logging))


# ----------------------------------------------------------------------------------------
# FIFOQueue (wavenet_ops.py:181-188 is the only user)
# ----------------------------------------------------------------------------------------
class FIFOQueue:
    def __init__(self, capacity, dtypes=None, shapes=None):
        self.capacity = int(capacity)
        self.item_shape = tuple(int(s) for s in shapes)
        self.items = deque()

    def enqueue_many(self, vals):
        vals = _wrap(vals)

        def run(v):
            assert v.shape[1:] == self.item_shape
            return ("enqueue", self, [np.array(x) for x in v])
        return Operation(run, [vals])

    def enqueue(self, vals):
        if isinstance(vals, (list, tuple)):       # wavenet_ops.py:188 passes a one-element list, Magenta/masked.py:138 the tensor
            (vals,) = vals
        val = _wrap(vals)

        def run(v):
            assert v.shape == self.item_shape, (v.shape, self.item_shape)
            return ("enqueue", self, [np.array(v)])
        return Operation(run, [val])

    def dequeue(self):
        def run():
            if not self.items:
                raise RuntimeError("dequeue on an empty FIFOQueue would block forever")
            return self.items.popleft()
        return Tensor(run, [], self.item_shape)


# ----------------------------------------------------------------------------------------
# Session
# ----------------------------------------------------------------------------------------
class Session:
    def run(self, fetches, feed_dict=None):
        feed = {id(k): np.asarray(v) for k, v in (feed_dict or {}).items()}
        cache, pending = {}, []

        def ev(node):
            key = id(node)
            if key in cache:
                return cache[key]
            if isinstance(node, Variable):
                r = node.value
            elif isinstance(node, _Placeholder):
                if key not in feed:
                    raise RuntimeError("placeholder %r was not fed" % node.name)
                r = feed[key].astype(node.dtype)
                assert list(r.shape) == node._shape.as_list(), (r.shape, node._shape)
            else:
                r = node.fn(*[ev(i) for i in node.inputs])
                if isinstance(node, Operation):
                    pending.append(r)
                    r = None
            cache[key] = r
            return r

        def walk(f):
            if isinstance(f, (list, tuple)):
                return [walk(x) for x in f]
            return ev(f)

        out = walk(fetches)
        for _, q, items in pending:            # enqueues land after this run's dequeues
            q.items.extend(items)
            if len(q.items) > q.capacity:
                raise RuntimeError("enqueue on a full FIFOQueue would block forever")
        return out


# ----------------------------------------------------------------------------------------
# train (EMA shadow names cannot be verified without TensorFlow: identity mapping)
# ----------------------------------------------------------------------------------------
class _EMA:
    def __init__(self, decay):
        self.decay = decay
        self.applied = []

    def apply(self, var_list=None):
        self.applied = list(var_list or [])
        return None

    def variables_to_restore(self):
        return {v.name + "/ExponentialMovingAverage": v for v in self.applied}


train = types.SimpleNamespace(ExponentialMovingAverage=_EMA)


# ----------------------------------------------------------------------------------------
# keras layers used by Encoder/encoder.py:8-26 (auto-named conv1d, conv1d_1, ... like tf.keras
# inside a variable_scope)
# ----------------------------------------------------------------------------------------
_layer_counts = {}


def _layer_name(base):
    key = ("/".join(_G.scope), base, id(_G))
    n = _layer_counts.get(key, 0)
    _layer_counts[key] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


class _Conv1D:
    """tf.keras.layers.Conv1D, channels_last; 'same' padding follows TF's SAME rule:
    out = ceil(T/s); total = max((out-1)*s + k - T, 0); left = total // 2"""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, dilation_rate=1):
        self.filters, self.k, self.s = filters, kernel_size, strides
        self.padding, self.activation = padding, activation

    def __call__(self, net):
        net = _wrap(net)
        name = _layer_name("conv1d")
        cin = net._shape[-1]
        with variable_scope(name):
            kernel = get_variable("kernel", [self.k, cin, self.filters])
            bias = get_variable("bias", [self.filters])
        T = net._shape[1]
        if self.padding == "same":
            out = -(-T // self.s)
            total = max((out - 1) * self.s + self.k - T, 0)
            net = pad(net, [[0, 0], [total // 2, total - total // 2], [0, 0]])
        y = _conv2d(expand_dims(net, 1), expand_dims(kernel, 0), [1, 1, self.s, 1], "VALID")
        y = squeeze(y, 1) + bias
        if self.activation == "relu":
            y = nn.relu(y)
        return y


class _BatchNormalization:
    """tf.keras.layers.BatchNormalization called without training= : inference form with the moving
    statistics, epsilon 1e-3, y = (x - mean) * gamma / sqrt(var + eps) + beta"""

    def __call__(self, net):
        net = _wrap(net)
        name = _layer_name("batch_normalization")
        c = net._shape[-1]
        with variable_scope(name):
            gamma = get_variable("gamma", [c])
            beta = get_variable("beta", [c])
            mean = get_variable("moving_mean", [c], trainable=False)
            var = get_variable("moving_variance", [c], trainable=False)
        inv = _unary(lambda v: (np.float32(1) / np.sqrt(v + np.float32(1e-3))).astype(np.float32), var)
        return (net - mean) * (gamma * inv) + beta


# ---- tf.contrib.signal (TensorFlow r1.12-1.14), used by Encoder/encoder_ops.py:14-43.  The library itself is absent:
# these are restatements of its published algorithms (signal/python/ops: spectral_ops.stft, window_ops.hann_window,
# mel_ops.linear_to_mel_weight_matrix, mfcc_ops.mfccs_from_log_mel_spectrograms), float32 like the TF kernels.
def _hann_window(window_length, periodic=True, dtype=np.float32):
    n = window_length if periodic else window_length - 1
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(window_length) / n)).astype(np.float32)


def _stft(signals, frame_length, frame_step, fft_length=None, window_fn=_hann_window, pad_end=False):
    signals = _wrap(signals)
    fft_length = fft_length or frame_length
    T = signals._shape[-1]
    nfr = -(-T // frame_step) if pad_end else 1 + (T - frame_length) // frame_step
    win = window_fn(frame_length)

    def run(x):
        xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, max((nfr - 1) * frame_step + frame_length - T, 0))])
        idx = np.arange(nfr)[:, None] * frame_step + np.arange(frame_length)[None, :]
        frames = (xp[..., idx] * win).astype(np.float32)
        return np.fft.rfft(frames, n=fft_length, axis=-1).astype(np.complex64)
    return Tensor(run, [signals], list(signals._shape.dims[:-1]) + [nfr, fft_length // 2 + 1], np.complex64)


def _linear_to_mel_weight_matrix(num_mel_bins=20, num_spectrogram_bins=129, sample_rate=8000, lower_edge_hertz=125.0,
                                 upper_edge_hertz=3800.0, dtype=np.float32):
    def hz_to_mel(f):
        return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    nsb = int(num_spectrogram_bins)
    lin = np.linspace(0.0, sample_rate / 2.0, nsb)[1:]
    spec_mel = hz_to_mel(lin)[:, None]
    edges = np.linspace(hz_to_mel(lower_edge_hertz), hz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    lower, center, upper = edges[:-2][None, :], edges[1:-1][None, :], edges[2:][None, :]
    w = np.maximum(0.0, np.minimum((spec_mel - lower) / (center - lower), (upper - spec_mel) / (upper - center)))
    return constant(np.pad(w, [(1, 0), (0, 0)]).astype(np.float32))


def _mfccs_from_log_mel_spectrograms(log_mel):
    log_mel = _wrap(log_mel)
    n = log_mel._shape[-1]
    m = np.arange(n)[:, None]
    c = np.arange(n)[None, :]
    dct = (2.0 * np.cos(np.pi * c * (2 * m + 1) / (2.0 * n)) / np.sqrt(2.0 * n)).astype(np.float32)
    return Tensor(lambda v: (v @ dct).astype(np.float32), [log_mel], log_mel._shape.dims)


contrib = types.SimpleNamespace(signal=types.SimpleNamespace(
    stft=_stft, hann_window=_hann_window, linear_to_mel_weight_matrix=_linear_to_mel_weight_matrix,
    mfccs_from_log_mel_spectrograms=_mfccs_from_log_mel_spectrograms))


keras = types.SimpleNamespace(
    layers=types.SimpleNamespace(Conv1D=_Conv1D, BatchNormalization=_BatchNormalization,
                                 UpSampling1D=None),
    regularizers=types.SimpleNamespace(l2=lambda *_: None))


def install():
    """make `import tensorflow as tf` resolve to this module"""
    sys.modules["tensorflow"] = sys.modules[__name__]
