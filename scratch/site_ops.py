"""Time the building blocks of one cfg-3 bulk site in isolation (CUDA events) -- development aid.
usage: python scratch/site_ops.py [svd|qr|gemm|all] [reps]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv, _lib
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
def rn(*s): return dv.DevArray.from_host(rng.standard_normal(s) + 1j * rng.standard_normal(s))
def timeit(name, fn, flops=None):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = dv.launch_count()
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print("%-34s %9.3f ms  launches %6d %s" % (name, ms, dv.launch_count() - l0, ("%.2f TFLOP/s" % (flops / ms / 1e9)) if flops else ""), flush=True)
if what in ("svd", "all"):
    A = rn(1024, 1536)
    timeit("svd_project 1024x1536 c128", lambda: dv.svd_project(A)); print("   sweeps", dv.last_svd_sweeps)
    timeit("svd full    1024x1536 c128", lambda: dv.svd(A)); print("   sweeps", dv.last_svd_sweeps)
if what in ("qr", "all"):
    B = rn(3072, 1536)
    timeit("qr 3072x1536 c128", lambda: dv.qr(B), 8 * (2 * 3072 * 1536**2 - 2 / 3 * 1536**3))
    C = rn(1536, 1024)
    timeit("qr 1536x1024 c128", lambda: dv.qr(C), 8 * (2 * 1536 * 1024**2 - 2 / 3 * 1024**3))
if what in ("gemm", "all"):
    R, T = rn(1536, 1536), rn(1536, 2, 1536)
    timeit("R-absorb 1536x1536 . 1536x3072", lambda: dv.tensordot(R, T, [1], [0]), 8 * 1536 * 1536 * 3072)
    V, Q = rn(512, 1536), rn(2, 1536, 1536)
    timeit("P-absorb 512x1536 . (3072x1536)^T", lambda: dv.tensordot(V, Q, [1], [2]), 8 * 512 * 1536 * 3072)
    X, Y = rn(3072, 128), rn(128, 1408)
    timeit("rank-128 update 3072x1408", lambda: dv.tensordot(X, Y, [1], [0]), 8 * 3072 * 1408 * 128)
    a, b = torch.randn(1536, 1536, dtype=torch.complex128, device="cuda"), torch.randn(1536, 3072, dtype=torch.complex128, device="cuda")
    timeit("cuBLAS zgemm 1536x1536x3072", lambda: torch.matmul(a, b), 8 * 1536 * 1536 * 3072)
if what in ("qrprof",):
    import ctypes
    lib = _lib.load()
    names = ["gemm", "jacobi_round", "qr_panel", "permute", "mps_mpo_site", "elementwise"]
    for shape in ((3072, 1536), (1536, 1024)):
        B = rn(*shape)
        dv.qr(B); torch.cuda.synchronize()
        lib.tnb_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dv.qr(B); e1.record(); e1.synchronize()
        print("qr", shape, "total %.3f ms" % e0.elapsed_time(e1))
        for c, nm in enumerate(names):
            ms, w = ctypes.c_double(), ctypes.c_double(); l, s = ctypes.c_longlong(), ctypes.c_longlong()
            lib.tnb_profile_get(c, ctypes.byref(ms), ctypes.byref(w), ctypes.byref(l), ctypes.byref(s))
            if l.value:
                print("   %-14s %8.3f ms  launches %5d  scopes %5d  rate %.2f" % (nm, ms.value, l.value, s.value, w.value / max(ms.value, 1e-9) / 1e9))
        lib.tnb_profile_enable(0)
