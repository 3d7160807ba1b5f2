"""ctypes binding of libtnb.so (the C ABI declared in include/tnb.h).

There is no CPU fallback: importing this module without the built library, or
calling into it without a CUDA device, raises.  PyTorch is used only to own
device buffers and streams.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TNB_LIB_PATH") or os.path.join(_HERE, "lib", "libtnb.so")  # override: kernel experiments

MAX_RANK = 24
F64, C128 = 0, 1
OP_N, OP_T, OP_C, OP_J = 0, 1, 2, 3
E_NOCONV = -4


class TnbError(RuntimeError):
    pass


class TensorDesc(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("shape", ctypes.c_int64 * MAX_RANK), ("stride", ctypes.c_int64 * MAX_RANK)]


_c = ctypes
_vp, _i32, _i64, _dbl, _sz = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_double, _c.c_size_t
_pd = _c.POINTER(TensorDesc)
_pi32 = _c.POINTER(_c.c_int32)
_pdbl = _c.POINTER(_c.c_double)

# name -> (restype, argtypes); must list every symbol of include/tnb.h
SIGNATURES = {
    "tnb_version": (_c.c_int, []),
    "tnb_error_string": (_c.c_char_p, [_c.c_int]),
    "tnb_device_info": (_c.c_int, [_pi32, _pi32, _pi32]),
    "tnb_launch_count": (_c.c_longlong, [_c.c_int]),
    "tnb_profile_enable": (_c.c_int, [_c.c_int]),
    "tnb_profile_get": (_c.c_int, [_c.c_int, _pdbl, _pdbl, _c.POINTER(_c.c_longlong), _c.POINTER(_c.c_longlong)]),
    "tnb_permute": (_c.c_int, [_pd, _pi32, _vp, _dbl, _dbl, _c.c_int, _vp]),
    "tnb_scale_inplace": (_c.c_int, [_pd, _dbl, _dbl, _vp]),
    "tnb_axpby": (_c.c_int, [_pd, _pd, _vp, _dbl, _dbl, _dbl, _dbl, _vp]),
    "tnb_norm2_workspace": (_sz, []),
    "tnb_norm2": (_c.c_int, [_pd, _vp, _vp, _sz, _vp]),
    "tnb_diag_embed": (_c.c_int, [_c.c_int, _vp, _i64, _vp, _c.c_int, _vp]),
    "tnb_diag_extract": (_c.c_int, [_pd, _vp, _vp]),
    "tnb_diag_scale": (_c.c_int, [_c.c_int, _vp, _i64, _i64, _i64, _vp, _c.c_int, _c.c_int, _vp]),
    "tnb_trace": (_c.c_int, [_pd, _c.c_int, _c.c_int, _vp, _vp]),
    "tnb_real_to_complex": (_c.c_int, [_vp, _i64, _vp, _vp]),
    "tnb_fill_uniform": (_c.c_int, [_vp, _i64, _c.c_ulonglong, _c.c_ulonglong, _vp]),
    "tnb_gemm": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int, _i64, _i64, _i64, _pdbl, _vp, _i64, _i64, _vp, _i64, _i64,
                            _pdbl, _vp, _i64, _i64, _i64, _vp]),
    "tnb_gemm_ws": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int, _i64, _i64, _i64, _pdbl, _vp, _i64, _vp, _i64, _pdbl, _vp, _i64,
                               _vp, _sz, _vp]),
    "tnb_tensordot_workspace": (_sz, [_pd, _pd, _c.c_int, _pi32, _pi32]),
    "tnb_tensordot": (_c.c_int, [_pd, _pd, _c.c_int, _pi32, _pi32, _c.c_int, _c.c_int, _vp, _vp, _sz, _vp]),
    "tnb_mps_mpo_site": (_c.c_int, [_pd, _pd, _vp, _vp]),
    "tnb_qr_workspace": (_sz, [_c.c_int, _i64, _i64]),
    "tnb_qr": (_c.c_int, [_c.c_int, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "tnb_svd_workspace": (_sz, [_c.c_int, _i64, _i64]),
    "tnb_svd": (_c.c_int, [_c.c_int, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _pi32, _vp]),
    "tnb_svd_project": (_c.c_int, [_c.c_int, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _pi32, _vp]),
    "tnb_truncation_count": (_c.c_int, [_vp, _i64, _i64, _dbl, _c.c_int, _vp, _vp, _vp]),
    # batched path
    "tnb_svd_project_batched_workspace": (_sz, [_c.c_int, _i64, _i64, _i64]),
    "tnb_svd_project_batched": (_c.c_int, [_c.c_int, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64,
                                           _vp, _sz, _pi32, _vp]),
    "tnb_truncation_count_batched": (_c.c_int, [_vp, _i64, _i64, _i64, _i64, _dbl, _c.c_int, _vp, _vp, _vp]),
    "tnb_norm2_batched": (_c.c_int, [_c.c_int, _vp, _i64, _i64, _i64, _vp, _vp]),
    "tnb_fill_uniform_batched": (_c.c_int, [_vp, _i64, _i64, _vp, _vp]),
}

_lib = None


def load():
    """Load libtnb.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TnbError("libtnb.so not found at %s -- run `python tncontract_b200/build.py` "
                       "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    msg = load().tnb_error_string(int(rc)).decode()
    if rc == E_NOCONV:
        raise np.linalg.LinAlgError(msg)          # tensor.py:914-929 error type
    if rc < 0:
        raise ValueError("libtnb: " + msg)
    raise TnbError("libtnb CUDA error %d: %s" % (rc, msg))


def dtype_code(np_dtype):
    if np_dtype == np.float64:
        return F64
    if np_dtype == np.complex128:
        return C128
    raise TypeError("libtnb supports float64 and complex128 only, got %r" % (np_dtype,))


def make_desc(ptr, dtype, shape, strides):
    d = TensorDesc()
    d.ptr = ptr
    d.dtype = dtype
    r = len(shape)
    if r > MAX_RANK:
        raise ValueError("tensor rank %d exceeds libtnb limit %d" % (r, MAX_RANK))
    d.rank = r
    for i in range(r):
        d.shape[i] = int(shape[i])
        d.stride[i] = int(strides[i])
    return d
