"""ctypes binding of libvqwn.so (include/vqwn.h).  No fallback: a missing library is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LAYERS = 64

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOTIMPL, ERR_NOMEM = 0, -1, -2, -3, -4, -5
MODE_GREEDY, MODE_SAMPLE = 0, 1
PREC_FP32, PREC_BF16, PREC_TC = 0, 1, 2
VQ_AUTO, VQ_DIRECT, VQ_TENSOR, VQ_TENSOR_BF16, VQ_EXPANDED = 0, 1, 2, 3, 4


class VqwnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vqwn error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("quantization_channels", C.c_int32),
        ("num_layers", C.c_int32),
        ("num_cycle_layers", C.c_int32),
        ("dilations", C.c_int32 * MAX_LAYERS),
        ("kernel_size", C.c_int32),
        ("dilation_filters", C.c_int32),
        ("skip_filters", C.c_int32),
        ("residual_filters", C.c_int32),
        ("pre_kernel_size", C.c_int32),
        ("pre_filters", C.c_int32),
        ("k", C.c_int32),
        ("latent_dim", C.c_int32),
        ("speaker_dim", C.c_int32),
        ("num_speakers", C.c_int32),
        ("use_vq", C.c_int32),
        ("encoder", C.c_int32),
    ]


def library_path():
    return os.environ.get("VQWN_LIBRARY", os.path.join(HERE, "libvqwn.so"))


_lib = None

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_H = C.c_void_p

# name -> (restype, argtypes): exactly the entry points include/vqwn.h declares
PROTOTYPES = {
    "vqwn_version": (C.c_char_p, []),
    "vqwn_create": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.POINTER(_H)]),
    "vqwn_destroy": (C.c_int, [_H]),
    "vqwn_last_error": (C.c_char_p, [_H]),
    "vqwn_set_stream": (C.c_int, [_H, C.c_void_p]),
    "vqwn_set_precision": (C.c_int, [_H, C.c_int]),
    "vqwn_set_reproducible": (C.c_int, [_H, C.c_int]),
    "vqwn_set_vq_kernel": (C.c_int, [_H, C.c_int]),
    "vqwn_set_vq_output": (C.c_int, [_H, C.c_int]),
    "vqwn_set_stream_offset": (C.c_int, [_H, C.c_int64]),
    "vqwn_set_tensor": (C.c_int, [_H, C.c_char_p, _f32p, _i64p, C.c_int]),
    "vqwn_get_tensor": (C.c_int, [_H, C.c_char_p, _f32p, C.c_int64]),
    "vqwn_num_tensors": (C.c_int, [_H]),
    "vqwn_tensor_info": (C.c_int, [_H, C.c_int, C.c_char_p, C.c_int, _i64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vqwn_encode_audio": (C.c_int, [_H, _f32p, C.c_int, C.c_int64, _f32p]),
    "vqwn_vq_lookup": (C.c_int, [_H, _f32p, C.c_int64, _i64p, _f32p]),
    "vqwn_build_condition": (C.c_int, [_H, _f32p, _i32p, C.c_int, C.c_int, _f32p]),
    "vqwn_encode_condition": (C.c_int, [_H, _f32p, _i32p, C.c_int, C.c_int, _i64p, _f32p]),
    "vqwn_reset": (C.c_int, [_H, C.c_int]),
    "vqwn_step": (C.c_int, [_H, _f32p, _f32p, _f32p, _f32p]),
    "vqwn_decode": (C.c_int, [_H, _f32p, C.c_int, C.c_int, _f64p, _i32p, _f32p]),
    "vqwn_generate": (C.c_int, [_H, _f32p, C.c_int, C.c_int, C.c_int64, C.c_int, _f64p, C.c_uint64, _f32p, _i32p]),
    "vqwn_teacher_forced": (C.c_int, [_H, _f32p, _f32p, C.c_int, C.c_int, C.c_int64, _f32p]),
    "vqwn_upload_condition": (C.c_int, [_H, _f32p, C.c_int, C.c_int]),
    "vqwn_upload_uniforms": (C.c_int, [_H, _f64p, C.c_int64, C.c_int]),
    "vqwn_generate_resident": (C.c_int, [_H, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_uint64]),
    "vqwn_download_output": (C.c_int, [_H, C.c_int, C.c_int64, _f32p, _i32p]),
    "vqwn_vq_upload": (C.c_int, [_H, _f32p, C.c_int64]),
    "vqwn_vq_resident": (C.c_int, [_H, C.c_int64]),
    "vqwn_vq_download": (C.c_int, [_H, C.c_int64, _i64p, _f32p]),
    "vqwn_last_kernel_ms": (C.c_double, [_H]),
    "vqwn_launch_count": (C.c_int64, [_H]),
    "vqwn_last_kernel_name": (C.c_char_p, [_H]),
}


def load_library():
    """Loads libvqwn.so and binds every prototype.  Raises if the library was not built --
    the product path never substitutes anything for it."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "CUDA library %s is missing: build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
