"""Device-time profile by kernel class of the batched path (cfg 4): python scratch/batch_prof.py [batch]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_batch
from tncontract_b200 import _lib
lib = _lib.load()
b = int(sys.argv[1]) if len(sys.argv) > 1 else 148
bench_batch.measure(0, 1, networks_per_gpu=b, batch_size=b)          # warm-up (allocator, plans)
lib.tnb_profile_enable(1)
rec = bench_batch.measure(0, 1, networks_per_gpu=b, batch_size=b)
torch.cuda.synchronize()
names = ["gemm", "jacobi_round", "qr_panel", "permute", "mps_mpo_site", "elementwise"]
tot = 0.0
for c, name in enumerate(names):
    ms, w = ctypes.c_double(), ctypes.c_double(); l, s = ctypes.c_longlong(), ctypes.c_longlong()
    lib.tnb_profile_get(c, ctypes.byref(ms), ctypes.byref(w), ctypes.byref(l), ctypes.byref(s))
    tot += ms.value
    print("%-14s %9.1f ms %8d launches  work %.3e  rate %.2f" % (name, ms.value, l.value, w.value, w.value / max(ms.value, 1e-9) / 1e9))
print("sum of classes %.1f ms; record: %s networks/s in %.1f ms" % (tot, rec["value"], rec["ms"]))
