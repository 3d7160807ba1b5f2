import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv, _lib
rng = np.random.default_rng(0)
A = dv.DevArray.from_host(rng.standard_normal((3072, 32)) + 1j * rng.standard_normal((3072, 32)))
for _ in range(3):
    dv.qr(A)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 20)()
_lib.load().tnb_debug_qr_stamps(buf, 20)
t = list(buf)
names = {0: "start", 1: "loaded", 2: "gram1", 3: "reduced1", 4: "chol1", 5: "inv1", 6: "mult1", 8: "gram2", 9: "reduced2", 10: "chol2", 12: "Rtot", 13: "Qtop", 14: "LU", 15: "T+Vtop", 16: "M inv", 17: "bcast", 18: "mult2", 19: "stored"}
prev = t[0]
for k in sorted(names):
    print("%-10s +%7d cyc  (%.2f us)  total %.2f us" % (names[k], t[k] - prev, (t[k] - prev) / 1965.0, (t[k] - t[0]) / 1965.0))
    prev = t[k]
