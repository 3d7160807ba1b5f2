"""float64 GEMM timing (CUDA events, L2 flushed) through tensordot: TNB_LIB_PATH selects the library variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv
def dev(*shape): return dv.DevArray(torch.randn(shape, dtype=torch.float64, device="cuda"))
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
for (M, N, K) in [(8192, 8192, 8192), (4096, 4096, 4096), (2048, 4096, 2048), (65536, 512, 4096)]:
    a, b = dev(M, K), dev(K, N)
    ms = timeit(lambda: dv.tensordot(a, b, [1], [0]))
    print("%s NN %5d x %5d x %5d  %8.3f ms  %6.2f TFLOP/s" % (tag, M, N, K, ms, 2e-9 * M * N * K / ms), flush=True)
    at = dev(K, M)
    ms = timeit(lambda: dv.tensordot(at, b, [0], [0]))
    print("%s TN %5d x %5d x %5d  %8.3f ms  %6.2f TFLOP/s" % (tag, M, N, K, ms, 2e-9 * M * N * K / ms), flush=True)
