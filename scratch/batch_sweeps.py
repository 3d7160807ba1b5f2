"""Jacobi sweeps per batched projection SVD of the cfg 4 unit of work (148 networks, N=64, d=4, chi=128 -> 64)."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import batched
orig = batched._svd_project_batched
log = []
def rec(a, want_p=True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = orig(a, want_p)
    torch.cuda.synchronize()
    log.append((tuple(a.shape[1:]), out[3], time.perf_counter() - t0))
    return out
batched._svd_project_batched = rec
batched.overlap_norm_compress_batched(3, list(range(148)))
log.clear()
t0 = time.perf_counter()
batched.overlap_norm_compress_batched(3, list(range(148, 296)))
print("unit of work (148 networks, synchronised per SVD): %.3f s" % (time.perf_counter() - t0))
by = collections.defaultdict(list)
for shape, sw, t in log: by[shape].append((sw, t))
for shape, v in sorted(by.items(), key=lambda kv: -sum(t for _, t in kv[1])):
    sw = [x for x, _ in v]
    print("%-12s calls %3d  sweeps min %2d mean %.1f max %2d  total %.3f s" % (shape, len(v), min(sw), np.mean(sw), max(sw), sum(t for _, t in v)))
