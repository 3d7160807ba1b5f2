"""TEST INFRASTRUCTURE ONLY: a NumPy stand-in for libtnb.so that implements
the C ABI of include/tnb.h on HOST pointers, so the Python host layer of
``tncontract_b200`` (label algebra, sweep drivers, boundary-MPS loop) can be
exercised in the CPU-only container against the golden fixtures.

It is installed by the ``host_backend`` fixture (tests/conftest.py), which
monkey-patches ``tncontract_b200._lib`` / ``devarray`` for the duration of one
test.  Nothing in the product imports this file; without the fixture the
package refuses to run without CUDA.
"""
import ctypes

import numpy as np

F64, C128 = 0, 1


def _val(p):
    """Address carried by a ctypes pointer-ish argument."""
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return ctypes.addressof(p)


def _dt(code):
    return np.complex128 if code == C128 else np.float64


def _flat(addr, count, dtype):
    """1-D ndarray view of `count` elements at host address `addr`."""
    dtype = np.dtype(dtype)
    if count == 0:
        return np.empty(0, dtype)
    buf = (ctypes.c_char * (count * dtype.itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=count)


def _view(desc):
    """Strided ndarray view described by a tnb_tensor_t (byref or struct)."""
    d = getattr(desc, "_obj", desc)
    dtype = np.dtype(_dt(d.dtype))
    shape = tuple(d.shape[i] for i in range(d.rank))
    stride = tuple(d.stride[i] for i in range(d.rank))
    if any(s == 0 for s in shape):
        return np.empty(shape, dtype)
    extent = 1 + sum((s - 1) * st for s, st in zip(shape, stride))
    base = _flat(d.ptr, extent, dtype)
    return np.lib.stride_tricks.as_strided(base, shape, tuple(st * dtype.itemsize for st in stride))


def _mat(addr, rows, cols, ld, dtype):
    if rows == 0 or cols == 0:
        return np.empty((rows, cols), dtype)
    base = _flat(addr, (rows - 1) * ld + cols, dtype)
    return np.lib.stride_tricks.as_strided(base, (rows, cols), (ld * base.itemsize, base.itemsize))


def _fn(mode):
    return {0: lambda x: x, 1: np.sqrt, 2: lambda x: 1.0 / x}[mode]


class FakeLib:
    """Same entry points as libtnb.so; all 'device' pointers are host pointers."""

    def __init__(self):
        self.launches = 0
        self.calls = {}

    def _count(self, name):
        self.launches += 1
        self.calls[name] = self.calls.get(name, 0) + 1

    # ---- misc ---------------------------------------------------------------
    def tnb_version(self):
        return 100

    def tnb_error_string(self, code):
        return b"fake libtnb error %d" % code

    def tnb_launch_count(self, reset):
        v = self.launches
        if reset:
            self.launches = 0
        return v

    # ---- data movement / elementwise ---------------------------------------------
    def tnb_permute(self, desc, perm, out, ar, ai, conj, stream):
        self._count("permute")
        x = _view(desc)
        p = [perm[i] for i in range(x.ndim)]
        y = np.transpose(x, p)
        if conj:
            y = np.conj(y)
        if not (ar == 1.0 and ai == 0.0):
            y = y * (complex(ar, ai) if x.dtype == np.complex128 else ar)
        _flat(_val(out), y.size, x.dtype)[:] = np.ascontiguousarray(y).ravel()
        return 0

    def tnb_scale_inplace(self, desc, ar, ai, stream):
        self._count("scale")
        x = _view(desc)
        x *= (complex(ar, ai) if x.dtype == np.complex128 else ar)
        return 0

    def tnb_axpby(self, dx, dy, out, ar, ai, br, bi, stream):
        self._count("axpby")
        x, y = _view(dx), _view(dy)
        a = complex(ar, ai) if x.dtype == np.complex128 else ar
        b = complex(br, bi) if x.dtype == np.complex128 else br
        _flat(_val(out), x.size, x.dtype)[:] = (a * x + b * y).ravel()
        return 0

    def tnb_norm2_workspace(self):
        return 64

    def tnb_norm2(self, desc, out, ws, nbytes, stream):
        self._count("norm2")
        _flat(_val(out), 1, np.float64)[0] = np.linalg.norm(_view(desc).ravel())
        return 0

    def tnb_diag_embed(self, code, s, n, out, mode, stream):
        self._count("diag_embed")
        sv = _flat(_val(s), n, np.float64)
        _flat(_val(out), n * n, _dt(code))[:] = np.diag(_fn(mode)(sv)).astype(_dt(code)).ravel()
        return 0

    def tnb_diag_extract(self, desc, out, stream):
        self._count("diag_extract")
        x = _view(desc)
        d = np.diagonal(x)
        _flat(_val(out), d.size, x.dtype)[:] = d
        return 0

    def tnb_diag_scale(self, code, x, rows, cols, ld, s, axis, mode, stream):
        self._count("diag_scale")
        m = _mat(_val(x), rows, cols, ld, _dt(code))
        f = _fn(mode)(_flat(_val(s), rows if axis == 0 else cols, np.float64))
        m *= (f[:, None] if axis == 0 else f[None, :])
        return 0

    def tnb_trace(self, desc, a1, a2, out, stream):
        self._count("trace")
        x = _view(desc)
        t = np.trace(x, axis1=a1, axis2=a2)
        _flat(_val(out), t.size, x.dtype)[:] = np.asarray(t).ravel()
        return 0

    def tnb_real_to_complex(self, x, n, out, stream):
        self._count("r2c")
        _flat(_val(out), n, np.complex128)[:] = _flat(_val(x), n, np.float64)
        return 0

    def tnb_truncation_count(self, s, n, chi, threshold, relative, info, scaled, stream):
        self._count("truncation")
        sv = _flat(_val(s), n, np.float64)
        s0 = sv[0] if n else 0.0
        lim = sv[:chi] if (0 < chi < n) else sv
        with np.errstate(all="ignore"):
            if relative == 2:
                kept = int(np.sum(lim / s0 > threshold))
            else:
                kept = int(np.sum(lim > (threshold * s0 if relative else threshold)))
            inf = _flat(_val(info), 2, np.float64)
            inf[0], inf[1] = kept, s0
            if _val(scaled):
                _flat(_val(scaled), n, np.float64)[:] = sv / s0
        return 0

    # ---- contractions ------------------------------------------------------------------
    def tnb_tensordot_workspace(self, da, db, nctr, ax, bx):
        return 0

    def tnb_tensordot(self, da, db, nctr, ax, bx, conj_a, conj_b, out, ws, nbytes, stream):
        self._count("tensordot")
        a, b = _view(da), _view(db)
        if conj_a:
            a = np.conj(a)
        if conj_b:
            b = np.conj(b)
        r = np.tensordot(a, b, ([ax[i] for i in range(nctr)], [bx[i] for i in range(nctr)]))
        _flat(_val(out), r.size, a.dtype)[:] = np.asarray(r).ravel()
        return 0

    def tnb_mps_mpo_site(self, dA, dW, out, stream):
        self._count("mps_mpo_site")
        A, W = _view(dA), _view(dW)
        r = np.einsum("qlr,abpq->laprb", A, W)
        _flat(_val(out), r.size, A.dtype)[:] = r.ravel()
        return 0

    # ---- factorisations (LAPACK stands in for the device algorithms) ---------------------------
    def tnb_qr_workspace(self, code, m, n):
        return 0

    def tnb_qr(self, code, m, n, A, lda, Q, R, ws, nbytes, stream):
        self._count("qr")
        a = _mat(_val(A), m, n, lda, _dt(code))
        q, r = np.linalg.qr(a, mode="reduced")
        k = min(m, n)
        _flat(_val(Q), m * k, a.dtype)[:] = q.ravel()
        _flat(_val(R), k * n, a.dtype)[:] = r.ravel()
        return 0

    def tnb_svd_workspace(self, code, m, n):
        return 0

    def tnb_svd(self, code, m, n, A, lda, U, S, Vh, ws, nbytes, sweeps, stream):
        self._count("svd")
        a = _mat(_val(A), m, n, lda, _dt(code))
        u, s, vh = np.linalg.svd(a, full_matrices=False)
        k = min(m, n)
        _flat(_val(U), m * k, a.dtype)[:] = u.ravel()
        _flat(_val(S), k, np.float64)[:] = s
        _flat(_val(Vh), k * n, a.dtype)[:] = vh.ravel()
        return 0


def _svd_project(self, code, m, n, A, lda, U, S, P, ws, nbytes, sweeps, stream):
    self._count("svd_project")
    a = _mat(_val(A), m, n, lda, _dt(code))
    u, s, vh = np.linalg.svd(a, full_matrices=False)
    k = min(m, n)
    _flat(_val(U), m * k, a.dtype)[:] = u.ravel()
    _flat(_val(S), k, np.float64)[:] = s
    if _val(P):
        _flat(_val(P), k * n, a.dtype)[:] = (s[:, None] * vh).ravel()
    return 0


def _fill_uniform(self, out, n, key, offset, stream):
    from tncontract_b200.batch import host_uniform
    self._count("fill_uniform")
    k = key.value if hasattr(key, "value") else key
    o = offset.value if hasattr(offset, "value") else offset
    _flat(_val(out), n, np.float64)[:] = host_uniform(n, k, o)
    return 0


FakeLib.tnb_svd_project = _svd_project
FakeLib.tnb_fill_uniform = _fill_uniform


def install(monkeypatch):
    """Route tncontract_b200 to FakeLib + host torch buffers (one test only)."""
    import torch
    from tncontract_b200 import _lib, devarray as dv
    fake = FakeLib()
    monkeypatch.setattr(_lib, "_lib", fake)
    monkeypatch.setattr(_lib, "load", lambda: fake)
    monkeypatch.setattr(dv, "_require_cuda", lambda: None)
    monkeypatch.setattr(dv, "device", lambda: torch.device("cpu"))
    monkeypatch.setattr(dv, "stream_ptr", lambda: None)
    return fake
