"""ctypes binding of libivf.so (the C ABI declared in include/ivf.h).

The product path has no CPU fallback: if the shared library is missing, or a call is made
without an sm_100 GPU, this module raises.  Nothing here imports `oracle/`.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libivf.so")

IVF_F32, IVF_BF16, IVF_U8 = 0, 1, 2
EP_AFFINE, EP_RELU, EP_ACCUM, EP_MASK, EP_OUT_F32, EP_LSTM = 1, 2, 4, 8, 16, 32
POOL_NONNEG = 64  # ivf_pool_desc.flags, forward: no negative input (IVF_POOL_NONNEG)
PFMT_NCDHW_F32, PFMT_NDHWC_F32, PFMT_S2D_BF16, PFMT_TBHWC_F32, PFMT_S2D2_BF16 = 0, 1, 2, 3, 4
PACK_KMAJOR, PACK_TAPMAJOR = 0, 1


class IvfError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n", "id", "ih", "iw", "od", "oh", "ow", "cin", "cout", "kd", "kh", "kw", "sd", "sh", "sw",
        "pd", "ph", "pw", "transposed", "in_ld", "in_coff", "out_ld", "out_coff", "mask_ld",
        "mask_coff", "flags", "dtype", "plan_kwm", "plan_mt", "plan_acc", "plan_ncta", "plan_ntiles", "plan_ds")]


class ConvSplit(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("split_cout", "out2_ld", "out2_coff", "split_cin", "in2_ld", "in2_coff")]


class PackDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "co", "ci", "kd", "kh", "kw", "ci_stride", "s2d_d", "s2d_h", "s2d_w", "dgrad", "layout", "dtype", "n_pad",
        "k_pad", "n_off", "k_off", "zero_first", "n_stride", "k_stride")]


class PoolDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n", "id", "ih", "iw", "c", "od", "oh", "ow", "kd", "kh", "kw", "sd", "sh", "sw", "pd", "ph",
        "pw", "in_ld", "in_coff", "out_ld", "out_coff", "mask_ld", "mask_coff", "flags", "dtype")]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float

# name -> (restype, argtypes); every symbol include/ivf.h declares
SIGNATURES = {
    "ivf_create": (_I, [_I, C.POINTER(_P)]),
    "ivf_destroy": (_I, [_P]),
    "ivf_last_error": (C.c_char_p, []),
    "ivf_version": (C.c_char_p, []),
    "ivf_launch_count": (C.c_int64, [_P]),
    "ivf_conv_bf16_kchunk": (_I, [_I]),
    "ivf_conv_bf16_cin_pad": (_I, [_I]),
    "ivf_conv_bf16_ntile": (_I, [_I]),
    "ivf_conv_bf16_cout_pad": (_I, [_I]),
    "ivf_conv_slab_plan": (_I, [C.POINTER(ConvDesc), _I, C.POINTER(C.c_int)]),
    "ivf_conv_slab_plan_ds": (_I, [C.POINTER(ConvDesc), _I]),
    "ivf_conv3d": (_I, [_P, C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ivf_conv3d_pair": (_I, [_P, C.POINTER(ConvDesc)] + [_P] * 8 + [C.POINTER(ConvDesc)] + [_P] * 8 + [_P]),
    "ivf_conv3d_split": (_I, [_P, C.POINTER(ConvDesc), C.POINTER(ConvSplit), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                _P]),
    "ivf_debug_read_scratch": (_I, [_P, _P, C.c_size_t]),
    "ivf_pack_weights": (_I, [_P, C.POINTER(PackDesc), _P, _P, _P]),
    "ivf_bn_fold": (_I, [_P, _P, _P, _P, _P, _F, _P, _I, _P, _P, _P]),
    "ivf_fill_u32": (_I, [_P, _P, C.c_size_t, C.c_uint32, _P]),
    "ivf_u8_to_f32": (_I, [_P, _P, _P, C.c_size_t, _P]),
    "ivf_one_hot": (_I, [_P, _P, _I, _I, _P, _P]),
    "ivf_argmax_rows": (_I, [_P, _P, _I, _I, _P, _P]),
    "ivf_maxpool3d_fwd": (_I, [_P, C.POINTER(PoolDesc), _P, _P, _P, _P]),
    "ivf_maxpool3d_bwd": (_I, [_P, C.POINTER(PoolDesc), _P, _P, _P, _P, _P, _P, _P]),
    "ivf_maxpool3d_fwd_bits": (_I, [_P, C.POINTER(PoolDesc), _P, _P, _P, _P, _P]),
    "ivf_maxpool3d_bwd_bits": (_I, [_P, C.POINTER(PoolDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "ivf_i3d_head_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "ivf_i3d_head_fwd": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "ivf_i3d_head_bwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P, _P, _I, _P, _I, _I, _P, _P, _P]),
    "ivf_perturb_fwd": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "ivf_perturb_bwd": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "ivf_mask_loss_adam": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _F, _F, _F, _F, _F, _F, _P, _P, _P]),
    "ivf_sigmoid": (_I, [_P, _P, _P, _I, _P]),
    "ivf_select_scores": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "ivf_init_mask_select": (_I, [_P, _P, _I, _I, _F, _P, _P, _P]),
    "ivf_tv_norm": (_I, [_P, _P, _I, _F, _F, _P, _P, _P]),
    "ivf_gradcam": (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "ivf_clstm_gates_fwd": (_I, [_P, _I, _P, _P, _I, _I, _P, _P, _P, _I, _P]),
    "ivf_clstm_gates_bwd": (_I, [_P, _I, _P, _P, _P, _P, _P, _I, _I, _P, _I, _P]),
    "ivf_conv3d_lstm": (_I, [_P, C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "ivf_bn_pool2d_fwd": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "ivf_bn_pool2d_bwd": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "ivf_bn_train_fwd": (_I, [_P, _I, _P, _I, _I, C.c_longlong, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _I, _I, _I,
                                _P]),
    "ivf_bn_train_bwd": (_I, [_P, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I, C.c_longlong, _I, _P, _P, _P, _P, _P, _I,
                                _I, _P, _P, _P]),
    "ivf_conv3d_wgrad": (_I, [_P, C.POINTER(ConvDesc), _I, _P, _P, _P, _P]),
    "ivf_conv3d_wgrad_s2d": (_I, [_P, C.POINTER(ConvDesc), _P, _P, _P, _I, _I, _I, _I, _P]),
    "ivf_head_train_fwd": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "ivf_head_train_bwd": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _P]),
    "ivf_dropout_mask": (_I, [_P, _P, C.c_longlong, _F, C.c_ulonglong, _P]),
    "ivf_optim_step": (_I, [_P, _I, _P, _P, _P, _P, C.c_longlong, _F, _F, _F, _F, _F, _I, _P]),
    "ivf_optim_step_multi": (_I, [_P, _I, _P, _I, _F, _F, _F, _F, _F, _I, _F, _P]),
    "ivf_viz_triptych": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "ivf_probe_im2col": (_I, [_P, C.POINTER(ConvDesc), _P, _I, _I, _I, _P, _P]),
}

_lib = None
_lib_lock = threading.Lock()


def load():
    """dlopen libivf.so and bind every symbol; raises if the library was not built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise IvfError(
                "libivf.so not found at %s: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error():
    return load().ivf_last_error().decode()


def check(rc, what=""):
    if rc != 0:
        raise IvfError("%s failed (code %d): %s" % (what or "libivf call", rc, last_error()))


_handles = {}
_handles_lock = threading.Lock()


def handle(device=None):
    """One ivf_handle per (device, thread): the C ABI's threading contract."""
    if not torch.cuda.is_available():
        raise IvfError("libivf needs an sm_100 GPU (torch.cuda.is_available() is False); "
                       "there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    elif isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    key = (int(device), threading.get_ident())
    with _handles_lock:
        h = _handles.get(key)
        if h is None:
            lib = load()
            out = _P()
            check(lib.ivf_create(int(device), C.byref(out)), "ivf_create")
            h = out
            _handles[key] = h
        return h


def launch_count(device=None):
    return int(load().ivf_launch_count(handle(device)))


def stream_ptr(device=None):
    """The caller's stream ON THE DEVICE THE BUFFERS LIVE ON (not the thread's current device): an engine built
    for cuda:1 can be driven while cuda:0 is current."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(fn):
    """Method decorator: run with self.device current (streams, graph capture and allocations made inside follow
    the current device; the C ABI guards itself by handle, this covers the torch side)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def dtype_code(t):
    if t.dtype == torch.float32:
        return IVF_F32
    if t.dtype == torch.bfloat16:
        return IVF_BF16
    raise IvfError("unsupported activation dtype %s" % t.dtype)
