"""The GEMM shapes of one first-pass block of the Gram-Schmidt QR (3072 x 1536 complex128, block at column j0), timed in
isolation through tnb_gemm_ws (split-K scratch as the QR has it): TNB_LIB_PATH selects the library variant."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv, _lib
lib = _lib.load()
m, n = 3072, 1536
Q = torch.randn(m, n, dtype=torch.complex128, device="cuda")
S = torch.zeros(n, n, dtype=torch.complex128, device="cuda")
P2 = torch.zeros(m, 256, dtype=torch.complex128, device="cuda")
ws = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
one, zero, mone = (ctypes.c_double * 2)(1, 0), (ctypes.c_double * 2)(0, 0), (ctypes.c_double * 2)(-1, 0)
def timeit(fn, reps=8, cold=False):
    for _ in range(2): fn()
    best = 1e9
    for _ in range(reps):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3
tag = os.environ.get("TNB_LIB_PATH", "default")[-18:]
es = 16
for j0 in (128, 512, 1088, 1472):
    b = 64
    qj, qp = Q.data_ptr(), Q.data_ptr() + j0 * es
    def sgemm():   # S = [Qj, W]^H W : (j0 + b) x b x m
        return lib.tnb_gemm_ws(1, 2, 0, j0 + b, b, m, one, qj, n, qp, n, zero, S.data_ptr() + j0 * es, n, ws.data_ptr(), ws.numel(), dv.stream_ptr())
    def proj():    # W -= Qj C : m x b x j0
        return lib.tnb_gemm_ws(1, 0, 0, m, b, j0, mone, qj, n, S.data_ptr() + j0 * es, n, one, qp, n, ws.data_ptr(), ws.numel(), dv.stream_ptr())
    def scale():   # P2 = W R^-1 : m x b x b
        return lib.tnb_gemm_ws(1, 0, 0, m, b, b, one, qp, n, S.data_ptr(), n, zero, P2.data_ptr(), 256, ws.data_ptr(), 0, dv.stream_ptr())
    for name, fn, fl in (("S  (j0+64)x64x3072", sgemm, 8.0 * (j0 + b) * b * m), ("P  3072x64xj0", proj, 8.0 * m * b * j0), ("WR 3072x64x64", scale, 8.0 * m * b * b)):
        assert fn() == 0
        t = timeit(fn)
        print("%s j0=%4d %-20s %7.1f us %6.2f TFLOP/s" % (tag, j0, name, t, fl / t / 1e6), flush=True)
for g0 in (256, 768, 1280):
    bg = 256
    qg = Q.data_ptr() + g0 * es
    def s2():      # (g0 + 256) x 256 x m
        return lib.tnb_gemm_ws(1, 2, 0, g0 + bg, bg, m, one, Q.data_ptr(), n, qg, n, zero, S.data_ptr(), 256, ws.data_ptr(), ws.numel(), dv.stream_ptr())
    def w2():      # m x 256 x (g0 + 256)
        return lib.tnb_gemm_ws(1, 0, 0, m, bg, g0 + bg, one, Q.data_ptr(), n, S.data_ptr(), 256, zero, P2.data_ptr(), 256, ws.data_ptr(), ws.numel(), dv.stream_ptr())
    for name, fn, fl in (("S2 (g0+256)x256x3072", s2, 8.0 * (g0 + bg) * bg * m), ("W2 3072x256x(g0+256)", w2, 8.0 * m * bg * (g0 + bg))):
        assert fn() == 0
        t = timeit(fn)
        print("%s g0=%4d %-20s %7.1f us %6.2f TFLOP/s" % (tag, g0, name, t, fl / t / 1e6), flush=True)
