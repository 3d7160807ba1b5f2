"""The HBM-bound kernels at the sizes the configurations give them, once each (for ncu) or timed with CUDA events.
python scratch/hbm_ops.py [time]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv, _lib
timed = len(sys.argv) > 1 and sys.argv[1] == "time"
def dev(shape, cplx=True):
    t = torch.randn(shape, dtype=torch.complex128 if cplx else torch.float64, device="cuda")
    return dv.DevArray(t)
A = dev((2, 512, 512)); W = dv.DevArray.from_host(np.random.default_rng(0).standard_normal((3, 3, 2, 2)) + 0j)
big = dev((512, 512, 3, 3, 2))                      # cfg 3 consolidate input (75 MB)
sq = dev((4096, 4096))                              # plain transpose, 268 MB
p5 = dev((256, 16, 256, 16), cplx=False)            # cfg 5 sized strided copy (134 MB)
x = dev((1536, 2, 1536))
ops = {
  "mps_mpo_site cfg3 bulk (75.5 MB out)": (lambda: dv.mps_mpo_site(A, W), (2*512*512 + 1536*2*1536) * 16),
  "permute consolidate (512,512,3,3,2)->(1536,2,1536) c128": (lambda: big.transpose([0, 2, 4, 1, 3]).copy(), 2 * big.size * 16),
  "permute transpose 4096x4096 c128": (lambda: sq.transpose([1, 0]).copy(), 2 * sq.size * 16),
  "permute (256,16,256,16)->(256,256,16,16) f64": (lambda: p5.transpose([0, 2, 1, 3]).copy(), 2 * p5.size * 8),
  "permute rows: contiguous copy 75 MB": (lambda: x.copy(), 2 * x.size * 16),
  "scale_inplace 75 MB": (lambda: x.__imul__(1.0000001), 2 * x.size * 16),
  "norm2 75 MB": (lambda: x.norm(), x.size * 16),
}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, (fn, nbytes) in ops.items():
    if not timed:
        fn(); torch.cuda.synchronize(); continue
    for _ in range(3): fn()
    best = 1e9
    for _ in range(10):
        flush.zero_()                                # evict L2 (126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-62s %8.1f us  %7.0f GB/s" % (name, best * 1e3, nbytes / (best * 1e-3) / 1e9), flush=True)
