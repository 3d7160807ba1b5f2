"""GPU parity tests of the libtnb C ABI (include/tnb.h) against NumPy on seeded
inputs.  Every call goes through ctypes into libtnb.so (tncontract_b200.devarray
is only the argument marshalling).  Tolerances: bit-exact for data movement,
1e-12 relative (north_star bound: 1e-10) for floating-point kernels."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _mods():
    import torch
    from tncontract_b200 import _lib, devarray as dv
    return torch, _lib, dv


def rnd(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return a


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d


def test_library_and_device():
    torch, _lib, dv = _mods()
    lib = _lib.load()
    assert lib.tnb_version() >= 100
    sms, major, minor = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    assert lib.tnb_device_info(ctypes.byref(sms), ctypes.byref(major), ctypes.byref(minor)) == 0
    assert major.value == 10 and sms.value > 100


@pytest.mark.parametrize("cplx", [False, True])
def test_permute_bit_exact(cplx):
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(0)
    cases = [((7,), (0,)), ((5, 9), (1, 0)), ((64, 48), (1, 0)), ((33, 65), (1, 0)), ((3, 4, 5), (2, 0, 1)),
             ((16, 3, 3, 2, 16), (0, 2, 3, 1, 4)), ((2, 40, 50), (1, 0, 2)), ((2, 40, 50), (2, 1, 0)),
             ((8, 8, 8, 8), (3, 2, 1, 0)), ((1, 5, 1, 6), (3, 2, 1, 0)), ((130, 2, 70), (2, 1, 0)),
             ((512, 3, 3, 2, 96), (0, 2, 3, 1, 4))]
    for shape, perm in cases:
        a = rnd(rng, shape, cplx)
        d = dv.DevArray.from_host(a)
        out = np.asarray(d.transpose(perm).copy())
        assert out.shape == tuple(shape[p] for p in perm)
        assert np.array_equal(out, np.transpose(a, perm)), (shape, perm)
        if cplx:
            assert np.array_equal(np.asarray(d.transpose(perm).conjugate()), np.conj(np.transpose(a, perm)))
    # strided / sliced views and scaling
    a = rnd(rng, (20, 30, 6), cplx)
    d = dv.DevArray.from_host(a)
    assert np.array_equal(np.asarray(d[2:17, :, 1:5].transpose(2, 0, 1).copy()), a[2:17, :, 1:5].transpose(2, 0, 1))
    assert np.array_equal(np.asarray(d[:, 3]), a[:, 3])
    assert np.array_equal(np.asarray(d * 2.0), a * 2.0)
    assert rel(np.asarray(d / 3.0), a / 3.0) < 1e-15
    assert np.array_equal(np.asarray(d.reshape(600, 6)), a.reshape(600, 6))
    assert np.array_equal(np.asarray(d.transpose(1, 0, 2).reshape(30, 120)), a.transpose(1, 0, 2).reshape(30, 120))


@pytest.mark.parametrize("cplx", [False, True])
def test_gemm_all_ops(cplx):
    torch, _lib, dv = _mods()
    lib = _lib.load()
    rng = np.random.default_rng(1)
    code = _lib.C128 if cplx else _lib.F64
    tdt = torch.complex128 if cplx else torch.float64

    def op(x, o):
        return {0: x, 1: x.T, 2: x.conj().T, 3: x.conj()}[o]

    shapes = [(1, 1, 1), (5, 7, 3), (64, 64, 64), (128, 128, 16), (129, 65, 33), (200, 300, 1), (17, 513, 40),
              (300, 31, 257), (256, 192, 128)]
    for (M, N, K) in shapes:
        for oa in range(4):
            for ob in range(4):
                # stored shapes: N/J -> (M,K); T/C -> (K,M)
                As = rnd(rng, (M, K) if oa in (0, 3) else (K, M), cplx)
                Bs = rnd(rng, (K, N) if ob in (0, 3) else (N, K), cplx)
                C0 = rnd(rng, (M, N), cplx)
                alpha = (0.7 - 0.2j) if cplx else 0.7
                beta = (0.3 + 0.5j) if cplx else -1.3
                A = torch.from_numpy(As).cuda()
                B = torch.from_numpy(Bs).cuda()
                C = torch.from_numpy(C0.astype(As.dtype)).cuda()
                al = (ctypes.c_double * 2)(np.real(alpha), np.imag(alpha))
                be = (ctypes.c_double * 2)(np.real(beta), np.imag(beta))
                rc = lib.tnb_gemm(code, oa, ob, M, N, K, al, A.data_ptr(), As.shape[1], 0, B.data_ptr(), Bs.shape[1], 0,
                                  be, C.data_ptr(), N, 0, 1, dv.stream_ptr())
                assert rc == 0
                ref = alpha * op(As, oa) @ op(Bs, ob) + beta * C0
                assert C.dtype == tdt
                assert rel(C.cpu().numpy(), ref) < TOL, (M, N, K, oa, ob)
    # strided batch with padded leading dimensions, beta = 0 on uninitialised C
    M, N, K, nb = 70, 50, 90, 5
    As = rnd(rng, (nb, M, K + 2), cplx)
    Bs = rnd(rng, (nb, K, N + 4), cplx)
    A, B = torch.from_numpy(As).cuda(), torch.from_numpy(Bs).cuda()
    C = torch.full((nb, M, N), float("nan"), dtype=tdt, device="cuda")
    one, zero = (ctypes.c_double * 2)(1, 0), (ctypes.c_double * 2)(0, 0)
    rc = lib.tnb_gemm(code, 0, 0, M, N, K, one, A.data_ptr(), K + 2, M * (K + 2), B.data_ptr(), N + 4, K * (N + 4), zero,
                      C.data_ptr(), N, M * N, nb, dv.stream_ptr())
    assert rc == 0
    ref = np.einsum("bmk,bkn->bmn", As[:, :, :K], Bs[:, :, :N])
    assert rel(C.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("cplx", [False, True])
def test_gemm_full_size_tiles_tma(cplx):
    """Products with >= 148 full-size tiles (128 x 64 complex128, 128 x 128 float64) take the TMA-fed kernels: ragged
    M / N / K edges (zero fill of out-of-bounds boxes), all operand layouts / conjugations, padded leading dimensions
    (even: tensor maps; odd for float64: no tensor map, cp.async kernel), a strided batch, and a broadcast operand
    (batch stride 0: cp.async kernel, same result)."""
    torch, _lib, dv = _mods()
    lib = _lib.load()
    rng = np.random.default_rng(14)
    code, tdt = (1, torch.complex128) if cplx else (0, torch.float64)
    op = lambda x, o: x if o == 0 else x.T if o == 1 else x.conj().T if o == 2 else x.conj()
    alpha, beta = (0.5 - 1.5j, 0.25 + 0.75j) if cplx else (0.5, -1.25)
    al = (ctypes.c_double * 2)(np.real(alpha), np.imag(alpha))
    be = (ctypes.c_double * 2)(np.real(beta), np.imag(beta))
    shapes = [(1500, 1100, 70), (1290, 1030, 257)] if cplx else [(1500, 1900, 70), (1290, 2000, 257)]
    for (M, N, K) in shapes:
        C0 = rnd(rng, (M, N), cplx)
        for oa in range(4):
            for ob in range(4):
                for pad_a, pad_b in ((2, 6), (3, 5)):
                    sa = (M, K) if oa in (0, 3) else (K, M)
                    sb = (K, N) if ob in (0, 3) else (N, K)
                    Af = rnd(rng, (sa[0], sa[1] + pad_a), cplx)     # padded leading dimensions
                    Bf = rnd(rng, (sb[0], sb[1] + pad_b), cplx)
                    A, B = torch.from_numpy(Af).cuda(), torch.from_numpy(Bf).cuda()
                    C = torch.from_numpy(C0.copy()).cuda()
                    rc = lib.tnb_gemm(code, oa, ob, M, N, K, al, A.data_ptr(), sa[1] + pad_a, 0, B.data_ptr(), sb[1] + pad_b, 0,
                                      be, C.data_ptr(), N, 0, 1, dv.stream_ptr())
                    assert rc == 0
                    ref = alpha * op(Af[:, :sa[1]], oa) @ op(Bf[:, :sb[1]], ob) + beta * C0
                    assert rel(C.cpu().numpy(), ref) < TOL, (M, N, K, oa, ob, pad_a)
    # strided batch and a broadcast B
    M, N, K, nb = (640, 760, 130, 3) if cplx else (900, 1000, 130, 3)
    As, Bs = rnd(rng, (nb, M, K), cplx), rnd(rng, (nb, K, N), cplx)
    A, B = torch.from_numpy(As).cuda(), torch.from_numpy(Bs).cuda()
    one, zero = (ctypes.c_double * 2)(1, 0), (ctypes.c_double * 2)(0, 0)
    for sB, ref in ((K * N, np.einsum("bmk,bkn->bmn", As, Bs)), (0, np.einsum("bmk,kn->bmn", As, Bs[0]))):
        C = torch.full((nb, M, N), float("nan"), dtype=tdt, device="cuda")
        rc = lib.tnb_gemm(code, 0, 0, M, N, K, one, A.data_ptr(), K, M * K, B.data_ptr(), N, sB, zero, C.data_ptr(), N, M * N, nb,
                          dv.stream_ptr())
        assert rc == 0
        assert rel(C.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("cplx", [False, True])
def test_gemm_split_k(cplx):
    """tnb_gemm_ws: Gram-like shapes (few C tiles, long K) cut along K over grid.z and reduced by the second kernel --
    the products under the Gram-Schmidt QR; all four operand layouts, conjugation, alpha / beta, K not a multiple of
    the split, and a scratch too small for any split (plain path)."""
    torch, _lib, dv = _mods()
    lib = _lib.load()
    rng = np.random.default_rng(13)
    code, tdt = (1, torch.complex128) if cplx else (0, torch.float64)
    op = lambda x, o: x if o == 0 else x.T if o == 1 else x.conj().T if o == 2 else x.conj()
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for M, N, K in [(200, 64, 3000), (1100, 64, 3072), (256, 256, 2050), (64, 64, 1500)]:
        for oa, ob in [(2, 0), (0, 0), (1, 3), (3, 2)]:
            As = rnd(rng, (M, K) if oa in (0, 3) else (K, M), cplx)
            Bs = rnd(rng, (K, N) if ob in (0, 3) else (N, K), cplx)
            C0 = rnd(rng, (M, N), cplx)
            alpha = (-1.0 + 0.25j) if cplx else -1.0
            beta = (1.0 + 0j) if cplx else 1.0
            A, B = torch.from_numpy(As).cuda(), torch.from_numpy(Bs).cuda()
            ref = alpha * op(As, oa) @ op(Bs, ob) + beta * C0
            for nbytes in (ws.numel(), 1024):
                C = torch.from_numpy(C0.astype(As.dtype)).cuda()
                al = (ctypes.c_double * 2)(np.real(alpha), np.imag(alpha))
                be = (ctypes.c_double * 2)(np.real(beta), np.imag(beta))
                rc = lib.tnb_gemm_ws(code, oa, ob, M, N, K, al, A.data_ptr(), As.shape[1], B.data_ptr(), Bs.shape[1], be,
                                     C.data_ptr(), N, ws.data_ptr(), nbytes, dv.stream_ptr())
                assert rc == 0
                assert rel(C.cpu().numpy(), ref) < TOL, (M, N, K, oa, ob, nbytes)


@pytest.mark.parametrize("cplx", [False, True])
def test_tensordot_layouts(cplx):
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(2)
    cases = [
        ((6, 5), (5, 7), [1], [0]),
        ((6, 5), (7, 5), [1], [1]),
        ((5, 6), (5, 7), [0], [0]),
        ((2, 9, 11), (9, 4), [1], [0]),                 # middle axis: batched over the leading axis
        ((4, 9), (2, 9, 11), [1], [1]),                 # R-absorb into A[phys,left,right]
        ((8, 12), (3, 7, 12), [1], [2]),                # V-absorb (NT)
        ((1, 1, 6, 10), (2, 6, 9), [2], [1]),           # ladder step 1
        ((1, 1, 10, 2, 9), (10, 2, 5), [2, 3], [0, 1]),  # ladder step 2
        ((3, 4, 5, 6), (6, 5, 2), [3, 2], [0, 1]),
        ((3, 4, 5, 6), (4, 6, 7), [1, 3], [0, 1]),      # needs a permutation copy
        ((3, 4), (5, 6), [], []),                       # outer product
        ((7,), (7,), [0], [0]),                         # scalar result
        ((2, 20, 30), (3, 3, 2, 2), [0], [3]),                # K = d, no fused path
        ((130, 140), (140, 150), [1], [0]),
    ]
    for sa, sb, aa, ab in cases:
        a, b = rnd(rng, sa, cplx), rnd(rng, sb, cplx)
        out = dv.tensordot(dv.DevArray.from_host(a), dv.DevArray.from_host(b), aa, ab)
        ref = np.tensordot(a, b, (aa, ab))
        assert out.shape == ref.shape
        assert rel(np.asarray(out), ref) < TOL, (sa, sb, aa, ab)
    # operands that are themselves transposed / sliced views, plus conj flags
    a, b = rnd(rng, (12, 10, 8), cplx), rnd(rng, (9, 12, 8), cplx)
    da, db = dv.DevArray.from_host(a).transpose(2, 0, 1)[:, 2:11], dv.DevArray.from_host(b).transpose(1, 2, 0)[2:11]
    ref = np.tensordot(a.transpose(2, 0, 1)[:, 2:11].conj(), b.transpose(1, 2, 0)[2:11], ([1, 0], [0, 1]))
    out = dv.tensordot(da, db, [1, 0], [0, 1], conj_a=True)
    assert rel(np.asarray(out), ref) < TOL
    # mixed dtypes promote to complex
    a, b = rnd(rng, (6, 5), False), rnd(rng, (5, 4), True)
    out = dv.tensordot(dv.DevArray.from_host(a), dv.DevArray.from_host(b), [1], [0])
    assert out.dtype == np.complex128 and rel(np.asarray(out), a @ b) < TOL


@pytest.mark.parametrize("cplx", [False, True])
def test_mps_mpo_site(cplx):
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(3)
    for d, Dl, Dr, wl, wr, dout in [(2, 5, 7, 3, 3, 2), (2, 1, 8, 1, 3, 2), (4, 16, 16, 4, 4, 4), (2, 64, 33, 3, 3, 2),
                                    (16, 8, 8, 16, 16, 16)]:
        A, W = rnd(rng, (d, Dl, Dr), cplx), rnd(rng, (wl, wr, dout, d), cplx)
        out = dv.mps_mpo_site(dv.DevArray.from_host(A), dv.DevArray.from_host(W))
        ref = np.einsum("qlr,abpq->laprb", A, W).reshape(Dl * wl, dout, Dr * wr)
        assert rel(np.asarray(out), ref) < TOL
    # strided operands (site stored as [left, phys, right])
    A0 = rnd(rng, (6, 2, 9), cplx)
    W0 = rnd(rng, (2, 2, 3, 3), cplx)
    out = dv.mps_mpo_site(dv.DevArray.from_host(A0).transpose(1, 0, 2), dv.DevArray.from_host(W0).transpose(2, 3, 0, 1))
    ref = np.einsum("qlr,abpq->laprb", A0.transpose(1, 0, 2), W0.transpose(2, 3, 0, 1)).reshape(18, 2, 27)
    assert rel(np.asarray(out), ref) < TOL


QR_SHAPES = [(1, 1), (5, 1), (1, 5), (8, 8), (40, 33), (33, 40), (128, 64), (257, 100), (100, 257), (512, 128),
             (700, 300), (1024, 512)]


@pytest.mark.parametrize("cplx", [False, True])
def test_qr(cplx):
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(4)
    for m, n in QR_SHAPES:
        a = rnd(rng, (m, n), cplx)
        q, r = dv.qr(dv.DevArray.from_host(a))
        q, r = np.asarray(q), np.asarray(r)
        k = min(m, n)
        assert q.shape == (m, k) and r.shape == (k, n)
        assert rel(q @ r, a) < TOL, (m, n)
        assert np.linalg.norm(q.conj().T @ q - np.eye(k)) < 1e-12 * max(1, k), (m, n)
        assert np.array_equal(np.tril(r, -1), np.zeros_like(r)), (m, n)
        assert np.max(np.abs(np.diag(r).imag)) == 0.0 if cplx else True
        # R is unique up to the signs of its rows: compare |diag| with LAPACK
        rr = np.linalg.qr(a, mode="r")
        assert rel(np.abs(np.diag(r)), np.abs(np.diag(rr))) < TOL, (m, n)
    # magnitudes whose squares overflow / underflow (LAPACK's scaled norms survive these)
    for scale in (1e170, 1e-170):
        a = rnd(rng, (30, 10), cplx)
        q, r = dv.qr(dv.DevArray.from_host(a * scale))
        assert rel(np.asarray(q) @ (np.asarray(r) / scale), a) < TOL
    # rank-deficient and zero columns must not produce NaNs
    a = rnd(rng, (30, 10), cplx)
    a[:, 3] = 0
    a[:, 7] = a[:, 2]
    q, r = dv.qr(dv.DevArray.from_host(a))
    assert np.all(np.isfinite(np.asarray(q))) and rel(np.asarray(q) @ np.asarray(r), a) < TOL



@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_qr_many_columns(cplx):
    """k >= 128 takes the Gram-Schmidt (BCGS2 + CholeskyQR) path of qr.cu; numerically dependent or very
    ill-conditioned inputs must trip its device checks and come out of the Householder path just as good."""
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(14)

    def check(a, tol=TOL):
        m, n = a.shape
        k = min(m, n)
        q, r = dv.qr(dv.DevArray.from_host(a))
        q, r = np.asarray(q), np.asarray(r)
        assert q.shape == (m, k) and r.shape == (k, n)
        assert np.all(np.isfinite(q)) and np.all(np.isfinite(r))
        assert rel(q @ r, a) < tol, (m, n)
        assert np.linalg.norm(q.conj().T @ q - np.eye(k)) < 1e-12 * k, (m, n)
        assert np.array_equal(np.tril(r, -1), np.zeros_like(r)), (m, n)
        return q, r

    for m, n in [(300, 200), (200, 300), (1000, 130), (129, 129), (640, 192)]:
        a = rnd(rng, (m, n), cplx)
        q, r = check(a)
        rr = np.linalg.qr(a, mode="r")
        k = min(m, n)
        assert rel(np.abs(np.diag(r)), np.abs(np.diag(rr))[:k]) < 1e-11, (m, n)
    for scale in (1e170, 1e-170):
        a = rnd(rng, (300, 200), cplx)
        q, r = dv.qr(dv.DevArray.from_host(a * scale))
        assert rel(np.asarray(q) @ (np.asarray(r) / scale), a) < TOL
    # dependent and zero columns
    a = rnd(rng, (400, 256), cplx)
    a[:, 70] = 0
    a[:, 150] = a[:, 20]
    a[:, 200] = a[:, 3] - 2 * a[:, 130]
    check(a)
    # columns graded over 12 decades: column-wise backward error
    a = rnd(rng, (400, 200), cplx) * np.logspace(0, -12, 200)[None, :]
    q, r = check(a)
    err = np.linalg.norm(q @ r - a, axis=0) / np.linalg.norm(a, axis=0)
    assert err.max() < 1e-11
    # condition number 1e12 (singular values graded, not the columns)
    u, _ = np.linalg.qr(rnd(rng, (300, 160), cplx))
    v, _ = np.linalg.qr(rnd(rng, (160, 160), cplx))
    a = (u * np.logspace(0, -12, 160)[None, :]) @ v.conj().T
    check(a)
    # moderate condition numbers: the second-pass Gram matrix leaves the neighbourhood of I where the
    # first-order factor is used (|G - I| ~ eps cond^2), so both branches of the 64 x 64 kernel are exercised
    for decades in (2, 4, 5):
        a = (u * np.logspace(0, -decades, 160)[None, :]) @ v.conj().T
        check(a, tol=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_qr_second_pass_skip(cplx):
    """Lagged Gram-Schmidt QR (qr.cu, m >= 1024, several 256-column groups): a group whose first pass left it
    orthonormal to 1e-14 entrywise keeps its first-pass columns (the update kernels of its second pass return
    at once on a device flag), every other group is corrected.  Either way no entry of Q^H Q - I may exceed the
    skip tolerance (plus the rounding of this test's own product), and R must be LAPACK's up to row signs."""
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(41)
    m, n = 1536, 768
    well = rnd(rng, (m, n), cplx)                       # groups come out of the first pass at rounding level
    u, _ = np.linalg.qr(rnd(rng, (m, n), cplx))
    v, _ = np.linalg.qr(rnd(rng, (n, n), cplx))
    graded = (u * np.logspace(0, -3, n)[None, :]) @ v.conj().T   # cond 1e3: |G - I| ~ eps cond^2 >> 1e-14, corrected
    mixed = np.concatenate([well[:, :256], graded[:, :256], well[:, 256:512]], axis=1)  # both kinds in one call
    for name, a in (("well", well), ("graded", graded), ("mixed", mixed)):
        q, r = dv.qr(dv.DevArray.from_host(a))
        q, r = np.asarray(q), np.asarray(r)
        assert rel(q @ r, a) < TOL, name
        e = np.abs(q.conj().T @ q - np.eye(n))
        assert e.max() < 3e-14, (name, e.max())
        rr = np.linalg.qr(a, mode="r")
        assert rel(np.abs(np.diag(r)), np.abs(np.diag(rr))) < 1e-11, name
        # the same factorisation twice gives the same bits (the device predicate is deterministic)
        q2, r2 = dv.qr(dv.DevArray.from_host(a))
        assert np.array_equal(np.asarray(q2), q) and np.array_equal(np.asarray(r2), r), name



@pytest.mark.gpu
def test_get_into_host_buffer():
    """DevArray.get(out=...) copies into a caller-provided (pinned) host buffer, also from strided views."""
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(15)
    a = rnd(rng, (6, 5, 4), True)
    d = dv.DevArray.from_host(a)
    buf = torch.empty((6, 5, 4), dtype=torch.complex128).pin_memory()
    assert d.get(out=buf) is buf
    torch.cuda.synchronize()
    assert np.array_equal(buf.numpy(), a)
    v = d.transpose((2, 0, 1)) if hasattr(d, "transpose") else None
    if v is not None:
        out = np.empty((4, 6, 5), dtype=np.complex128)
        v.get(out=out)
        torch.cuda.synchronize()
        assert np.array_equal(out, a.transpose(2, 0, 1))
    with pytest.raises(ValueError):
        d.get(out=np.empty((5, 6, 4), dtype=np.complex128))


SVD_SHAPES = [(1, 1), (1, 6), (6, 1), (2, 2), (12, 20), (20, 12), (16, 16), (17, 17), (33, 31), (64, 128), (128, 64),
              (100, 100), (256, 128), (300, 520), (512, 512)]


def _check_svd(a, u, s, vh, tol_s=1e-12):
    m, n = a.shape
    k = min(m, n)
    assert u.shape == (m, k) and s.shape == (k,) and vh.shape == (k, n)
    sref = np.linalg.svd(a, compute_uv=False)
    assert np.all(np.diff(s) <= 0)
    assert np.max(np.abs(s - sref)) <= tol_s * sref[0], np.max(np.abs(s - sref)) / sref[0]
    assert rel((u * s) @ vh, a) < 1e-11
    assert np.linalg.norm(u.conj().T @ u - np.eye(k)) < 1e-11 * max(1, k)
    assert np.linalg.norm(vh @ vh.conj().T - np.eye(k)) < 1e-11 * max(1, k)


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_shapes(cplx):
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(5)
    for m, n in SVD_SHAPES:
        a = rnd(rng, (m, n), cplx)
        u, s, vh = dv.svd(dv.DevArray.from_host(a))
        assert s.dtype == np.float64
        _check_svd(a, np.asarray(u), np.asarray(s), np.asarray(vh))


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_hard_spectra(cplx):
    """Graded, clustered and rank-deficient matrices: small singular values must
    keep RELATIVE accuracy (this is what rules out the Gram-matrix shortcut)."""
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(6)
    n = 96
    # graded columns: sigma spans 12 decades
    a = rnd(rng, (120, n), cplx) * np.logspace(0, -12, n)[None, :]
    u, s, vh = dv.svd(dv.DevArray.from_host(a))
    s = np.asarray(s)
    sref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - sref) / sref) < 1e-9
    _check_svd(a, np.asarray(u), s, np.asarray(vh))
    # exactly rank-deficient: the null singular values are at noise level, the rest accurate
    b = rnd(rng, (80, 20), cplx) @ rnd(rng, (20, 64), cplx)
    u, s, vh = dv.svd(dv.DevArray.from_host(b))
    s = np.asarray(s)
    sref = np.linalg.svd(b, compute_uv=False)
    assert np.max(np.abs(s[:20] - sref[:20]) / sref[:20]) < 1e-11
    assert np.all(s[20:] < 1e-13 * s[0])
    assert rel((np.asarray(u) * s) @ np.asarray(vh), b) < 1e-11
    # U[0,1) entries (init_mps_random): one dominant singular value
    c = rng.random((128, 64)) + (1j * rng.random((128, 64)) if cplx else 0)
    u, s, vh = dv.svd(dv.DevArray.from_host(c))
    _check_svd(c, np.asarray(u), np.asarray(s), np.asarray(vh))
    # huge / tiny magnitudes: the Gram matrices inside Jacobi must not overflow or underflow
    for scale in (1e170, 1e-170):
        e = rnd(rng, (40, 24), cplx)
        u, s, vh = dv.svd(dv.DevArray.from_host(e * scale))
        _check_svd(e, np.asarray(u), np.asarray(s) / scale, np.asarray(vh))
    # NaN input -> LinAlgError, like numpy.linalg.svd
    e = rnd(rng, (12, 9), cplx)
    e[3, 4] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        dv.svd(dv.DevArray.from_host(e))
    # zero matrix
    z = np.zeros((8, 5), dtype=complex if cplx else float)
    u, s, vh = dv.svd(dv.DevArray.from_host(z))
    assert np.array_equal(np.asarray(s), np.zeros(5)) and np.all(np.isfinite(np.asarray(u)))


@pytest.mark.parametrize("cplx", [False, True])
def test_elementwise(cplx):
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(7)
    a = rnd(rng, (37, 5, 11), cplx)
    d = dv.DevArray.from_host(a)
    assert abs(d.norm() - np.linalg.norm(a)) < 1e-13 * np.linalg.norm(a)
    assert abs(d.transpose(2, 0, 1)[1:9].norm() - np.linalg.norm(a.transpose(2, 0, 1)[1:9])) < 1e-12
    big = rnd(rng, (1 << 20,), cplx)
    assert abs(dv.DevArray.from_host(big).norm() - np.linalg.norm(big)) < 1e-13 * np.linalg.norm(big)
    e = d.copy()
    e *= 0.5
    assert np.array_equal(np.asarray(e), a * 0.5)
    v = d.transpose(1, 0, 2)[1:4]
    v *= 2.0  # in place through a strided view: the parent sees it
    ref = a.copy()
    ref[:, 1:4] *= 2.0
    assert np.array_equal(np.asarray(d), ref)
    b = rnd(rng, (37, 5, 11), cplx)
    assert rel(np.asarray(d + dv.DevArray.from_host(b)), ref + b) < 1e-15
    assert rel(np.asarray(d - dv.DevArray.from_host(b)), ref - b) < 1e-15
    s = np.abs(rng.standard_normal(9)) + 0.1
    ds = dv.DevArray.from_host(s)
    for mode, f in ((0, lambda x: x), (1, np.sqrt), (2, lambda x: 1.0 / x)):
        assert rel(np.asarray(dv.diag_embed(ds, a.dtype, mode)), np.diag(f(s))) < 1e-15
    m = rnd(rng, (9, 14), cplx)
    assert rel(np.asarray(dv.diag_scale_rows(dv.DevArray.from_host(m), ds)), s[:, None] * m) < 1e-15
    assert rel(np.asarray(dv.diag_scale_cols(dv.DevArray.from_host(m.T.copy()), ds)), m.T * s[None, :]) < 1e-15
    assert np.array_equal(np.asarray(dv.diag_extract(dv.DevArray.from_host(m))), np.diag(m))
    t = rnd(rng, (4, 6, 4, 3), cplx)
    assert rel(np.asarray(dv.trace(dv.DevArray.from_host(t), 0, 2)), np.trace(t, axis1=0, axis2=2)) < 1e-14
    sv = np.sort(np.abs(rng.standard_normal(40)))[::-1].copy()
    sv[30:] *= 1e-17
    dsv = dv.DevArray.from_host(sv)
    for chi, thr, relm in [(0, 1e-15, 0), (12, 1e-15, 0), (0, 1e-14, 1), (0, 1e-14, 2), (35, 1e-14, 2), (0, 0.5, 2)]:
        kept, s0, scaled = dv.truncation(dsv, chi, thr, relm)
        lim = sv[:chi] if chi else sv
        if relm == 0:
            want = int(np.sum(lim > thr))
        elif relm == 1:
            want = int(np.sum(lim > thr * sv[0]))
        else:
            want = int(np.sum(lim / sv[0] > thr))
        assert kept == want and s0 == sv[0]
        assert np.array_equal(np.asarray(scaled), sv / sv[0])


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_1024_columns(cplx):
    """k >= 1024 columns, the size class of the cfg 3 bulk sites: clusters of four CTAs per block pair, block
    counts with and without padding blocks, V accumulated (tnb_svd) and not (tnb_svd_project), and a graded
    matrix, whose pairs take the double-precision Gram domain."""
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(12)
    for m, n in [(1100, 1024), (1024, 1300), (1040, 1040)]:
        a = rnd(rng, (m, n), cplx)
        u, s, p = dv.svd_project(dv.DevArray.from_host(a))
        u, s, p = np.asarray(u), np.asarray(s), np.asarray(p)
        k = min(m, n)
        sref = np.linalg.svd(a, compute_uv=False)
        assert np.max(np.abs(s - sref)) <= 1e-12 * sref[0]
        assert np.linalg.norm(u.conj().T @ u - np.eye(k)) < 1e-11 * k
        assert rel(u @ p, a) < 1e-11
    a = rnd(rng, (1030, 1024), cplx)
    u, s, vh = dv.svd(dv.DevArray.from_host(a))
    _check_svd(a, np.asarray(u), np.asarray(s), np.asarray(vh))
    g = rnd(rng, (1100, 1024), cplx) * np.logspace(0, -11, 1024)[None, :]
    u, s, p = dv.svd_project(dv.DevArray.from_host(g))
    sref = np.linalg.svd(g, compute_uv=False)
    assert np.max(np.abs(np.asarray(s) - sref) / sref) < 1e-8
    assert np.linalg.norm(np.asarray(u).conj().T @ np.asarray(u) - np.eye(1024)) < 1e-9


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_project_rank_deficient(cplx):
    """Projection SVD of rank-deficient matrices (the boundary matrices of BASELINE.json config 5 and the sites of
    config 2 -- a random positive MPS is numerically of rank ~1 -- are of this kind).  The reference's truncation
    keeps whatever np.linalg.svd reports above 1e-15 (tensor.py:1146-1150), noise directions included, and relies on
    their vectors being orthonormal: U must be an isometry over ALL k columns, not only the live ones.  (A floor
    under which columns are left unrotated was tried for the sweep count -- cfg 5: 23 -> 19 sweeps per 4096 x 4096
    matrix -- and dropped because it breaks exactly this.)"""
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(23)
    for m, n, r in [(300, 260, 90), (260, 300, 200), (520, 512, 17)]:
        a = rnd(rng, (m, r), cplx) @ rnd(rng, (r, n), cplx)
        u, s, p = dv.svd_project(dv.DevArray.from_host(a))
        sweeps = int(dv.last_svd_sweeps)
        u, s, p = np.asarray(u), np.asarray(s), np.asarray(p)
        k = min(m, n)
        sref = np.linalg.svd(a, compute_uv=False)
        assert np.max(np.abs(s - sref)) <= 1e-12 * sref[0], (m, n, r)
        assert rel(u @ p, a) < 1e-12, (m, n, r)
        assert int(np.sum(s > 1e-12 * s[0])) == r, (m, n, r)
        assert np.max(np.abs(u.conj().T @ u - np.eye(k))) < 1e-11, (m, n, r)
        # Jacobi runs on R^H (wide) or on R2^H of the second QR (tall / square): the null space does not hold the
        # sweeps up (on the columns of R itself these matrices took 23-24 sweeps)
        assert sweeps <= 14, (m, n, r, sweeps)
    # degenerate inputs of the tall / square path (second QR of R^H): zero matrix, identity, a single column
    z = np.zeros((8, 5), dtype=complex if cplx else float)
    u, s, p = dv.svd_project(dv.DevArray.from_host(z))
    assert np.array_equal(np.asarray(s), np.zeros(5)) and np.all(np.isfinite(np.asarray(u))) and np.all(np.asarray(p) == 0)
    e = np.eye(70, dtype=complex if cplx else float)
    u, s, p = dv.svd_project(dv.DevArray.from_host(e))
    assert np.max(np.abs(np.asarray(s) - 1)) < 1e-14 and rel(np.asarray(u) @ np.asarray(p), e) < 1e-14
    c = rnd(rng, (200, 1), cplx)
    u, s, p = dv.svd_project(dv.DevArray.from_host(c))
    assert abs(np.asarray(s)[0] - np.linalg.norm(c)) < 1e-13 * np.linalg.norm(c) and rel(np.asarray(u) @ np.asarray(p), c) < 1e-14


@pytest.mark.parametrize("cplx", [False, True])
def test_svd_project(cplx):
    """tnb_svd_project: U, S and P = U^H A = diag(S) Vh, V never accumulated."""
    torch, _lib, dv = _mods()
    rng = np.random.default_rng(8)
    for m, n in [(1, 1), (6, 1), (1, 6), (20, 12), (12, 20), (64, 64), (96, 200), (200, 96), (256, 384)]:
        a = rnd(rng, (m, n), cplx)
        u, s, p = dv.svd_project(dv.DevArray.from_host(a))
        u, s, p = np.asarray(u), np.asarray(s), np.asarray(p)
        k = min(m, n)
        assert u.shape == (m, k) and s.shape == (k,) and p.shape == (k, n)
        sref = np.linalg.svd(a, compute_uv=False)
        assert np.max(np.abs(s - sref)) <= 1e-12 * sref[0]
        assert np.linalg.norm(u.conj().T @ u - np.eye(k)) < 1e-11 * max(1, k)
        assert rel(p, u.conj().T @ a) < 1e-12
        assert rel(u @ p, a) < 1e-11
        # rows of P are orthogonal with norms S
        g = p @ p.conj().T
        assert np.linalg.norm(g - np.diag(s ** 2)) < 1e-11 * sref[0] ** 2 * max(1, k)
    # graded spectrum: small singular values keep relative accuracy without V
    a = rnd(rng, (90, 140), cplx) * np.logspace(0, -10, 140)[None, :]
    u, s, p = dv.svd_project(dv.DevArray.from_host(a))
    sref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(np.asarray(s) - sref) / sref) < 1e-9
    b = (rnd(rng, (150, 70), cplx) * np.logspace(0, -10, 70)[None, :])
    u, s, p = dv.svd_project(dv.DevArray.from_host(b))
    sref = np.linalg.svd(b, compute_uv=False)
    assert np.max(np.abs(np.asarray(s) - sref) / sref) < 1e-9
    assert np.linalg.norm(np.asarray(u).conj().T @ np.asarray(u) - np.eye(70)) < 1e-10
