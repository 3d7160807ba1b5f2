"""GEMM timing at the shapes of the cfg 3 sweep (CUDA events, L2 flushed), through tensordot (no split-K scratch) and
through whole QR / SVD calls (split-K inside): TNB_LIB_PATH selects the library variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv
def dev(*shape):
    return dv.DevArray(torch.randn(shape, dtype=torch.complex128, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=8):
    for _ in range(2): fn()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
tag = os.environ.get("TNB_LIB_PATH", "default")[-20:]
# C = A(MxK) B(KxN): NN;  A^H B through conj + axes
for (M, N, K) in [(4096, 4096, 4096), (1536, 3072, 1536), (1024, 1536, 1024), (3072, 512, 1024), (3072, 256, 1536), (3072, 64, 1536)]:
    a, b = dev(M, K), dev(K, N)
    ms = timeit(lambda: dv.tensordot(a, b, [1], [0]))
    print("%s NN %5d x %5d x %5d  %8.3f ms  %6.2f TFLOP/s" % (tag, M, N, K, ms, 8e-9 * M * N * K / ms), flush=True)
    at = dev(K, M)
    ms = timeit(lambda: dv.tensordot(at, b, [0], [0], conj_a=True))
    print("%s CN %5d x %5d x %5d  %8.3f ms  %6.2f TFLOP/s" % (tag, M, N, K, ms, 8e-9 * M * N * K / ms), flush=True)
q = dev(3072, 1536)
print("%s QR 3072x1536 %8.3f ms" % (tag, timeit(lambda: dv.qr(q), 5)), flush=True)
x = dev(1024, 1536)
print("%s svd_project 1024x1536 %8.3f ms (sweeps %d)" % (tag, timeit(lambda: dv.svd_project(x), 3), dv.last_svd_sweeps), flush=True)
