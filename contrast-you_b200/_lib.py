"""ctypes binding of libcontrastyou_b200.so — the only door from Python into the CUDA kernels.

The signatures mirror include/contrastyou_b200.h one to one.  Loading is lazy (first kernel call or an explicit
``load()``) so that host-side helpers can be imported on machines without the library, but every compute entry
point goes through :func:`lib` and raises ``RuntimeError`` if the shared object is missing — there is no fallback.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcontrastyou_b200.so")

# dtype / variant / path codes (contrastyou_b200.h)
CY_F32, CY_BF16, CY_F16, CY_F32_SPLIT = 0, 1, 2, 3
CY_SUPCON, CY_SUPCON_EXCLUDE, CY_SELFPACED_HARD, CY_SELFPACED_SOFT = 0, 1, 2, 3
CY_PATH_AUTO, CY_PATH_SIMT, CY_PATH_TCGEN05 = 0, 1, 2
CY_NSTAT = 8
CY_ABI_VERSION = 7
CY_STAT_LOGDEN, CY_STAT_INVC, CY_STAT_COEF, CY_STAT_AUX = 0, 1, 2, 3

_c = ctypes
_vp, _i64, _i32, _f32, _f64, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float, _c.c_double, _c.c_size_t

# name -> (restype, argtypes); kept in header order so tests can diff it against include/contrastyou_b200.h
SIGNATURES = {
    "cy_abi_version": (_i32, []),
    "cy_last_error": (_c.c_char_p, []),
    "cy_device_sm_count": (_i32, []),
    "cy_launch_count": (_c.c_ulonglong, []),
    "cy_infonce_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "cy_infonce_fwd": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _f32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "cy_infonce_fwd_pass2": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _f32, _i32, _f32, _i32, _vp, _vp, _vp,
                                    _sz, _vp]),
    "cy_infonce_loss": (_i32, [_i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cy_infonce_bwd": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _f32, _i32, _f32, _i32, _vp, _vp, _vp,
                              _i64, _vp, _sz, _vp]),
    "cy_infonce_masks": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "cy_labels_canonicalize": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp]),
    "cy_infonce_pack": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "cy_infonce_unpack": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cy_infonce_pack_split": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "cy_infonce_pack_gather": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "cy_infonce_unpack_scatter": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "cy_iic_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "cy_iic_joint": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "cy_iic_epilogue_workspace_bytes": (_sz, [_i32, _i32]),
    "cy_iic_epilogue": (_i32, [_vp, _i32, _i32, _i32, _i32, _f32, _f32, _f64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cy_iic_bwd": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cy_iic_joint_heads": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _c.c_longlong, _vp, _sz, _vp]),
    "cy_iic_epilogue_heads": (_i32, [_vp, _c.c_longlong, _c.c_longlong, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f64, _vp, _vp, _vp,
                                     _c.c_longlong, _vp, _sz, _vp]),
    "cy_iic_bwd_heads": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _c.c_longlong, _vp, _vp, _vp, _vp]),
    "cy_softmax_t_fwd": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "cy_iic_bwd_logits_heads": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _c.c_longlong, _vp, _f32, _vp, _vp,
                                       _vp]),
    "cy_imsat_workspace_bytes": (_sz, [_i32]),
    "cy_imsat_fwd": (_i32, [_vp, _i32, _i64, _i32, _i64, _f32, _vp, _vp, _vp, _sz, _vp]),
    "cy_imsat_bwd": (_i32, [_vp, _i32, _i64, _i32, _i64, _f32, _vp, _vp, _vp, _vp]),
    "cy_p2p_push": (_i32, [_vp, _i32, _i32, _vp, _i32, _vp]),
    "cy_p2p_push_barrier": (_i32, [_vp, _i32, _i32, _vp, _i32, _c.c_ulonglong, _vp, _c.c_uint, _vp]),
}

_LIB = None


def load():
    """dlopen the library and bind every symbol of SIGNATURES.  Raises RuntimeError when it is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  contrast-you_b200 has no CPU / eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here == header / library mismatch
            fn.restype, fn.argtypes = res, args
        if handle.cy_abi_version() != CY_ABI_VERSION:
            raise RuntimeError("libcontrastyou_b200.so: ABI version mismatch")
        _LIB = handle
    return _LIB


def lib():
    return load()


def check(rc, what):
    if rc != 0:
        msg = load().cy_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return {torch.float32: CY_F32, torch.bfloat16: CY_BF16, torch.float16: CY_F16}[t.dtype]
    except KeyError:
        raise TypeError(f"contrast-you_b200 kernels take float32 / bfloat16 / float16 tensors, got {t.dtype}") from None


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("contrast-you_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                               f"got a tensor on {t.device}")


def stream_ptr(device=None) -> int:
    """Raw cudaStream_t of torch's current stream (the kernels must launch there: SURVEY.md §8b).  Goes through the C
    accessor: ``torch.cuda.current_stream()`` builds a Python Stream object and costs ~20 us per call."""
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    return torch._C._cuda_getCurrentRawStream(idx)


def ptr(t):
    return None if t is None else t.data_ptr()


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def guard(t: torch.Tensor):
    """Context that makes ``t``'s device current for the launches inside (the kernels launch in the current CUDA
    context; a model living on cuda:1 without ``torch.cuda.set_device(1)`` must still work, like the reference does on
    any device).  Free when the device is already current."""
    idx = t.device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)
