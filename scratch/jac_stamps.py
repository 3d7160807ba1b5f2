"""clock64 stamps of the rotation phase of one Jacobi round (CTA 0): TNB_LIB_PATH=scratch/exp/libtnb_STAMPS.so"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv, _lib
rng = np.random.default_rng(0)
A = dv.DevArray.from_host(rng.standard_normal((1024, 1024)) + 1j * rng.standard_normal((1024, 1024)))
dv.svd_project(A)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["TNB_LIB_PATH"])
buf = (ctypes.c_longlong * 128)()
lib.tnb_debug_jacobi_stamps(buf, 128)
t = list(buf)
print("total rotation phase: %d cycles (%.2f us)" % (t[65] - t[0], (t[65] - t[0]) / 1965.0))
for s in range(16):
    b = 1 + 4 * s
    nxt = t[b + 4] if s < 15 else t[65]
    print("step %2d: start +%6d | rot warp done +%5d | G warp0 done +%5d | W warp8 done +%5d | step length %5d" %
          (s, t[b] - t[0], t[b + 1] - t[b], t[b + 2] - t[b], t[b + 3] - t[b], nxt - t[b]))
ph = ["entry", "prologue done", "griddep wait done", "load+Gram done", "reduction done", "rotation done", "apply+store done"]
for i in range(1, 7):
    print("%-20s +%6d cycles (%.2f us)" % (ph[i], t[70 + i] - t[70 + i - 1], (t[70 + i] - t[70 + i - 1]) / 1965.0))
print("kernel body total %.2f us" % ((t[76] - t[70]) / 1965.0))
if t[80]:
    print("G task of warp 0, step 5: begin +%d after the step's start stamp | operands loaded +%d | block computed +%d | done +%d" %
          (t[80] - t[21], t[81] - t[80], t[82] - t[80], t[23] - t[80]))
