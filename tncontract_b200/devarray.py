"""Device array: the ndarray subset that tncontract uses on ``Tensor.data``
("Array surface under Tensor.data"), backed by a CUDA
buffer.  PyTorch owns the storage, the view metadata (shape/strides/offset)
and the stream; every arithmetic operation is a libtnb kernel called through
the C ABI.  There is no CPU fallback: constructing a DevArray without a CUDA
device raises.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.complex128): torch.complex128}
_T2NP = {torch.float64: np.dtype(np.float64), torch.complex128: np.dtype(np.complex128)}


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.TnbError("tncontract_b200 needs a CUDA device (B200); there is no CPU fallback")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def device():
    _require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _empty(shape, np_dtype):
    return torch.empty(tuple(int(s) for s in shape), dtype=_NP2T[np.dtype(np_dtype)], device=device())


def workspace(nbytes):
    """Scratch buffer (torch caching allocator); returns (tensor, void*)."""
    n = max(int(nbytes), 16)
    t = torch.empty(n, dtype=torch.uint8, device=device())
    return t, ctypes.c_void_p(t.data_ptr())


def _canon_dtype(a):
    """float64 / complex128 are the only device dtypes; everything else that
    NumPy would hold exactly in them is widened (ints, bools, f16/f32, c64)."""
    if a.dtype in (np.float64, np.complex128):
        return a
    if a.dtype.kind in "biuf":
        return a.astype(np.float64)
    if a.dtype.kind == "c":
        return a.astype(np.complex128)
    raise TypeError("cannot place dtype %r on the device" % (a.dtype,))


def _is_scalar(x):
    return isinstance(x, (int, float, complex, np.number)) and not isinstance(x, np.longdouble)


class DevArray:
    __array_priority__ = 1000
    __array_ufunc__ = None  # numpy scalars defer to our reflected operators

    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t

    # ---- construction / transfer -------------------------------------------
    @staticmethod
    def from_host(obj):
        a = _canon_dtype(np.asarray(obj))
        _require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device(), copy=True)  # always a private copy
        return DevArray(t.reshape(a.shape))

    @staticmethod
    def zeros(shape, dtype=np.float64):
        t = _empty(shape, dtype)
        t.zero_()  # cudaMemset on the current stream
        return DevArray(t)

    @staticmethod
    def empty(shape, dtype=np.float64):
        return DevArray(_empty(shape, dtype))

    def __array__(self, dtype=None, copy=None):
        src = self if self.t.is_contiguous() else self.copy()
        out = src.t.cpu().numpy()
        if dtype is not None:
            out = out.astype(dtype)
        return out

    def get(self, out=None):
        """Host copy.  With ``out`` (a torch CPU tensor -- pinned for full PCIe speed -- or a NumPy array of the
        same shape and dtype) the copy is enqueued on the current stream into that buffer and ``out`` is
        returned; the caller synchronises (``torch.cuda.synchronize()``) before reading it."""
        if out is None:
            return self.__array__()
        src = self if self.t.is_contiguous() else self.copy()
        dst = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
        if tuple(dst.shape) != tuple(src.t.shape) or dst.dtype != src.t.dtype:
            raise ValueError("get(out=...): shape/dtype mismatch %s %s vs %s %s" %
                             (tuple(dst.shape), dst.dtype, tuple(src.t.shape), src.t.dtype))
        dst.copy_(src.t, non_blocking=True)
        return out

    def item(self):
        return self.__array__().item()

    def __float__(self):
        return float(self.item())

    def __complex__(self):
        return complex(self.item())

    def __bool__(self):
        return bool(self.item())

    # ---- metadata -------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return _T2NP[self.t.dtype]

    @property
    def strides_elems(self):
        return tuple(self.t.stride())

    @property
    def ptr(self):
        return self.t.data_ptr()

    def desc(self):
        return _lib.make_desc(self.t.data_ptr(), _lib.dtype_code(self.dtype), self.t.shape, self.t.stride())

    def __len__(self):
        return self.t.shape[0]

    def __repr__(self):
        return "DevArray(%r)" % (self.__array__(),)

    def __str__(self):
        return str(self.__array__())

    # ---- views (metadata only, no kernels) -------------------------------------
    def transpose(self, *axes):
        if len(axes) == 1 and not isinstance(axes[0], int):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(range(self.ndim))[::-1]
        return DevArray(self.t.permute(*axes))

    def moveaxis(self, src, dst):
        order = list(range(self.ndim))
        order.insert(dst, order.pop(src))
        return DevArray(self.t.permute(*order))

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        for k in key:
            if not (isinstance(k, (int, slice, np.integer)) or k is None or k is Ellipsis):
                raise IndexError("DevArray supports basic indexing only (ints, slices, newaxis)")
        key = tuple(int(k) if isinstance(k, np.integer) else k for k in key)
        return DevArray(self.t[key])

    def is_contiguous(self):
        return self.t.is_contiguous()

    def reshape(self, *shape):
        """np.reshape semantics: a view when the strides allow it, otherwise a
        contiguous copy made by the permute kernel (tensor.py:295,315,363,818)."""
        if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)):
            shape = tuple(shape[0])
        shape = tuple(int(s) for s in shape)
        try:
            return DevArray(self.t.view(shape))
        except RuntimeError:
            return DevArray(self.copy().t.view(shape))

    def flatten(self):
        return DevArray(self.copy().t.view(-1))

    # ---- kernels ---------------------------------------------------------------------
    def _permute_copy(self, alpha=1.0, conj=False, out=None):
        """Contiguous copy of this (strided) view through tnb_permute; ``out``: an existing CONTIGUOUS DevArray of
        the same shape and dtype to write into (used to place a slab inside a larger buffer)."""
        if out is not None:
            assert out.shape == self.shape and out.dtype == self.dtype and out.t.is_contiguous()
            out = out.t
        else:
            out = _empty(self.shape, self.dtype)
        if self.size == 0:
            return DevArray(out)
        perm = (ctypes.c_int32 * max(self.ndim, 1))(*range(self.ndim))
        a = complex(alpha)
        d = self.desc()
        _lib.check(_lib.load().tnb_permute(ctypes.byref(d), perm, ctypes.c_void_p(out.data_ptr()),
                                           a.real, a.imag, int(conj), stream_ptr()))
        return DevArray(out)

    def copy(self):
        return self._permute_copy()

    def contiguous(self):
        return self if self.t.is_contiguous() else self.copy()

    def conjugate(self):
        return self._permute_copy(conj=(self.dtype == np.complex128))

    conj = conjugate

    def to_complex(self):
        if self.dtype == np.complex128:
            return self
        src = self.contiguous()
        out = _empty(self.shape, np.complex128)
        if self.size:
            _lib.check(_lib.load().tnb_real_to_complex(ctypes.c_void_p(src.ptr), self.size,
                                                       ctypes.c_void_p(out.data_ptr()), stream_ptr()))
        return DevArray(out)

    def _scaled(self, alpha):
        alpha = complex(alpha) if isinstance(alpha, (complex, np.complexfloating)) else float(alpha)
        if isinstance(alpha, complex) and self.dtype == np.float64:
            if alpha.imag == 0.0:
                return self._permute_copy(alpha.real)  # numpy would still promote; keep real
            return self.to_complex()._permute_copy(alpha)
        return self._permute_copy(alpha)

    def _as_scalar(self, other):
        """Python/NumPy number, or a size-1 DevArray (e.g. S.data[0, 0], onedim_utils.py:55)."""
        if _is_scalar(other):
            return other
        if isinstance(other, DevArray) and other.size == 1:
            return other.item()
        if isinstance(other, np.ndarray) and other.size == 1 and other.dtype != np.longdouble:
            return other.item()
        return None

    def __mul__(self, other):
        if isinstance(other, (np.longdouble, np.clongdouble)) or (
                isinstance(other, np.ndarray) and other.dtype in (np.longdouble, np.clongdouble)):
            # square_lattice.py:145,198: the float128 norm accumulator; host scalar result
            return self.__array__() * other
        s = self._as_scalar(other)
        if s is None:
            return NotImplemented
        if isinstance(s, (complex, np.complexfloating)) and self.dtype == np.float64:
            return self.to_complex()._permute_copy(complex(s))
        return self._scaled(s)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, (np.longdouble, np.clongdouble)):
            return self.__array__() / other
        s = self._as_scalar(other)
        if s is None:
            return NotImplemented
        return self.__mul__(1.0 / s)

    def __neg__(self):
        return self._scaled(-1.0)

    def _inplace_scale(self, s):
        a = complex(s)
        if a.imag != 0.0 and self.dtype == np.float64:
            raise TypeError("cannot scale a float64 device array in place by a complex number")
        if self.size:
            d = self.desc()
            _lib.check(_lib.load().tnb_scale_inplace(ctypes.byref(d), a.real, a.imag, stream_ptr()))
        return self

    @staticmethod
    def _narrow_longdouble(other):
        """``x.data *= norm`` with the float128 norm accumulator of twodim.mps_contract (square_lattice.py:162,177,
        188): NumPy keeps the float64 array and rounds the factor; so do we (the array stays on the device)."""
        if isinstance(other, (np.longdouble, np.clongdouble)):
            return complex(other) if isinstance(other, np.clongdouble) else float(other)
        if isinstance(other, np.ndarray) and other.size == 1 and other.dtype in (np.longdouble, np.clongdouble):
            return complex(other.item()) if other.dtype == np.clongdouble else float(other.item())
        return other

    def __imul__(self, other):
        s = self._as_scalar(self._narrow_longdouble(other))
        if s is None:
            return NotImplemented
        return self._inplace_scale(s)

    def __itruediv__(self, other):
        s = self._as_scalar(self._narrow_longdouble(other))
        if s is None:
            return NotImplemented
        return self._inplace_scale(1.0 / s)

    def _axpby(self, other, a, b):
        if not isinstance(other, DevArray):
            other = DevArray.from_host(other)
        if self.shape != other.shape:
            raise ValueError("operands could not be broadcast together with shapes %r %r" % (self.shape, other.shape))
        x, y = self, other
        if x.dtype != y.dtype:
            x, y = x.to_complex(), y.to_complex()
        out = _empty(x.shape, x.dtype)
        if x.size:
            dx, dy = x.desc(), y.desc()
            _lib.check(_lib.load().tnb_axpby(ctypes.byref(dx), ctypes.byref(dy), ctypes.c_void_p(out.data_ptr()),
                                             a, 0.0, b, 0.0, stream_ptr()))
        return DevArray(out)

    def __add__(self, other):
        return self._axpby(other, 1.0, 1.0)

    def __radd__(self, other):
        return self._axpby(other, 1.0, 1.0)

    def __sub__(self, other):
        return self._axpby(other, 1.0, -1.0)

    def norm(self):
        """Frobenius norm as a host float (np.linalg.norm; one 8-byte read-back)."""
        lib = _lib.load()
        n = lib.tnb_norm2_workspace()
        ws_t, ws = workspace(n)
        res = torch.empty(1, dtype=torch.float64, device=device())
        d = self.desc()
        _lib.check(lib.tnb_norm2(ctypes.byref(d), ctypes.c_void_p(res.data_ptr()), ws, n, stream_ptr()))
        return np.float64(res.item())


# ---- free functions on device arrays ---------------------------------------------
def tensordot(a, b, axes_a, axes_b, conj_a=False, conj_b=False):
    """np.tensordot(a, b, (axes_a, axes_b)) on device (tensor.py:735)."""
    if a.dtype != b.dtype:
        a, b = a.to_complex(), b.to_complex()
    n = len(axes_a)
    if n != len(axes_b):
        raise ValueError("shape-mismatch for sum")
    for x, y in zip(axes_a, axes_b):
        if a.shape[x] != b.shape[y]:
            raise ValueError("shape-mismatch for sum")
    oshape = ([s for i, s in enumerate(a.shape) if i not in axes_a] +
              [s for i, s in enumerate(b.shape) if i not in axes_b])
    out = _empty(oshape, a.dtype)
    if out.numel() == 0:
        return DevArray(out)
    lib = _lib.load()
    da, db = a.desc(), b.desc()
    ax = (ctypes.c_int32 * max(n, 1))(*axes_a)
    bx = (ctypes.c_int32 * max(n, 1))(*axes_b)
    need = lib.tnb_tensordot_workspace(ctypes.byref(da), ctypes.byref(db), n, ax, bx)
    ws_t, ws = workspace(need)
    _lib.check(lib.tnb_tensordot(ctypes.byref(da), ctypes.byref(db), n, ax, bx, int(conj_a), int(conj_b),
                                 ctypes.c_void_p(out.data_ptr()), ws, need, stream_ptr()))
    return DevArray(out)


def _as_matrix(a):
    """2-D row-major view with unit column stride (copy through the permute kernel if needed)."""
    assert a.ndim == 2
    st = a.strides_elems
    if a.shape[1] > 1 and st[1] != 1:
        a = a.copy()
    elif a.shape[0] > 1 and st[0] < a.shape[1]:
        a = a.copy()
    st = a.strides_elems
    lda = st[0] if a.shape[0] > 1 else max(a.shape[1], 1)
    return a, max(int(lda), 1)


def qr(a):
    """np.linalg.qr(a, mode='reduced') (tensor.py:1044) -> (q, r) device arrays."""
    m, n = a.shape
    k = min(m, n)
    a, lda = _as_matrix(a)
    q, r = _empty((m, k), a.dtype), _empty((k, n), a.dtype)
    if m * n == 0:
        return DevArray(q), DevArray(r)
    lib = _lib.load()
    code = _lib.dtype_code(a.dtype)
    need = lib.tnb_qr_workspace(code, m, n)
    ws_t, ws = workspace(need)
    _lib.check(lib.tnb_qr(code, m, n, ctypes.c_void_p(a.ptr), lda, ctypes.c_void_p(q.data_ptr()),
                          ctypes.c_void_p(r.data_ptr()), ws, need, stream_ptr()))
    return DevArray(q), DevArray(r)


last_svd_sweeps = 0


def svd(a):
    """np.linalg.svd(a, full_matrices=False) (tensor.py:915) -> (u, s, vh); s float64 descending."""
    global last_svd_sweeps
    m, n = a.shape
    k = min(m, n)
    a, lda = _as_matrix(a)
    u, s, vh = _empty((m, k), a.dtype), _empty((k,), np.float64), _empty((k, n), a.dtype)
    if m * n == 0:
        return DevArray(u), DevArray(s), DevArray(vh)
    lib = _lib.load()
    code = _lib.dtype_code(a.dtype)
    need = lib.tnb_svd_workspace(code, m, n)
    ws_t, ws = workspace(need)
    sweeps = ctypes.c_int32(0)
    _lib.check(lib.tnb_svd(code, m, n, ctypes.c_void_p(a.ptr), lda, ctypes.c_void_p(u.data_ptr()),
                           ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(vh.data_ptr()), ws, need,
                           ctypes.byref(sweeps), stream_ptr()))
    last_svd_sweeps = sweeps.value
    return DevArray(u), DevArray(s), DevArray(vh)


def svd_project(a):
    """(u, s, p) with p = u^H a = diag(s) vh: the form the MPS sweeps consume (tnb_svd_project)."""
    global last_svd_sweeps
    m, n = a.shape
    k = min(m, n)
    a, lda = _as_matrix(a)
    u, s, p = _empty((m, k), a.dtype), _empty((k,), np.float64), _empty((k, n), a.dtype)
    if m * n == 0:
        return DevArray(u), DevArray(s), DevArray(p)
    lib = _lib.load()
    code = _lib.dtype_code(a.dtype)
    need = lib.tnb_svd_workspace(code, m, n)
    ws_t, ws = workspace(need)
    sweeps = ctypes.c_int32(0)
    _lib.check(lib.tnb_svd_project(code, m, n, ctypes.c_void_p(a.ptr), lda, ctypes.c_void_p(u.data_ptr()),
                                   ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(p.data_ptr()), ws, need,
                                   ctypes.byref(sweeps), stream_ptr()))
    last_svd_sweeps = sweeps.value
    return DevArray(u), DevArray(s), DevArray(p)


def truncation(s, chi, threshold, relative):
    """Kept-rank rule on device; returns (kept:int, s0:float, s_scaled DevArray).
    One 16-byte read-back (the kept rank fixes the next tensor's shape)."""
    n = s.size
    info = torch.empty(2, dtype=torch.float64, device=device())
    scaled = _empty((n,), np.float64)
    _lib.check(_lib.load().tnb_truncation_count(ctypes.c_void_p(s.ptr), n, int(chi or 0), float(threshold),
                                                int(bool(relative)), ctypes.c_void_p(info.data_ptr()),
                                                ctypes.c_void_p(scaled.data_ptr()), stream_ptr()))
    kept, s0 = info.cpu().tolist()
    return int(kept), float(s0), DevArray(scaled)


def diag_embed(s, dtype=np.float64, mode=0):
    """np.diag(s) (mode 0), np.diag(sqrt(s)) (1), np.diag(1/s) (2)."""
    n = s.size
    out = _empty((n, n), dtype)
    if n:
        s = s.contiguous()
        _lib.check(_lib.load().tnb_diag_embed(_lib.dtype_code(np.dtype(dtype)), ctypes.c_void_p(s.ptr), n,
                                              ctypes.c_void_p(out.data_ptr()), mode, stream_ptr()))
    return DevArray(out)


def diag_extract(a):
    n = min(a.shape)
    out = _empty((n,), a.dtype)
    if n:
        d = a.desc()
        _lib.check(_lib.load().tnb_diag_extract(ctypes.byref(d), ctypes.c_void_p(out.data_ptr()), stream_ptr()))
    return DevArray(out)


def diag_scale_rows(x, s, mode=0):
    """In place: x[i, ...] *= f(s[i]) for a CONTIGUOUS x (first axis = bond)."""
    assert x.is_contiguous() and s.is_contiguous()
    rows = x.shape[0]
    cols = x.size // rows if rows else 0
    if rows and cols:
        _lib.check(_lib.load().tnb_diag_scale(_lib.dtype_code(x.dtype), ctypes.c_void_p(x.ptr), rows, cols, cols,
                                              ctypes.c_void_p(s.ptr), 0, mode, stream_ptr()))
    return x


def diag_scale_cols(x, s, mode=0):
    """In place: x[..., j] *= f(s[j]) for a CONTIGUOUS x (last axis = bond)."""
    assert x.is_contiguous() and s.is_contiguous()
    cols = x.shape[-1]
    rows = x.size // cols if cols else 0
    if rows and cols:
        _lib.check(_lib.load().tnb_diag_scale(_lib.dtype_code(x.dtype), ctypes.c_void_p(x.ptr), rows, cols, cols,
                                              ctypes.c_void_p(s.ptr), 1, mode, stream_ptr()))
    return x


def trace(a, axis1, axis2):
    """np.trace(a, axis1=..., axis2=...) (tensor.py:329)."""
    oshape = [s for i, s in enumerate(a.shape) if i not in (axis1, axis2)]
    out = _empty(oshape, a.dtype)
    if out.numel():
        if a.shape[axis1] == 0 or a.shape[axis2] == 0:
            out.zero_()
        else:
            d = a.desc()
            _lib.check(_lib.load().tnb_trace(ctypes.byref(d), axis1, axis2, ctypes.c_void_p(out.data_ptr()),
                                             stream_ptr()))
    return DevArray(out)


def mps_mpo_site(a, w):
    """Fused onedim_core.py:1702-1704 for A[d,Dl,Dr] and W[wl,wr,dout,d] views."""
    if a.dtype != w.dtype:
        a, w = a.to_complex(), w.to_complex()
    d, Dl, Dr = a.shape
    wl, wr, dout, d2 = w.shape
    assert d == d2
    out = _empty((Dl * wl, dout, Dr * wr), a.dtype)
    if out.numel():
        da, dw = a.desc(), w.desc()
        _lib.check(_lib.load().tnb_mps_mpo_site(ctypes.byref(da), ctypes.byref(dw),
                                                ctypes.c_void_p(out.data_ptr()), stream_ptr()))
    return DevArray(out)


def launch_count(reset=False):
    return int(_lib.load().tnb_launch_count(int(reset)))


def asdevarray(obj, copy=True):
    if isinstance(obj, DevArray):
        return obj.copy() if copy else obj
    if isinstance(obj, torch.Tensor):
        if obj.dtype not in _T2NP:
            obj = obj.to(torch.complex128 if obj.is_complex() else torch.float64)
        t = obj.to(device())
        d = DevArray(t)
        return d.copy() if (copy and t.data_ptr() == obj.data_ptr()) else d
    return DevArray.from_host(obj)
