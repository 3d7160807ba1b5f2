"""Phase timing of the Jacobi round kernel: run with TNB_LIB_PATH=scratch/exp/libtnb_SKIP_X.so and
TNB_JACOBI_FIXED_SWEEPS=5; per-round time = total Jacobi class time / launches (device profile)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv, _lib
lib = _lib.load()
rng = np.random.default_rng(0)
for k in (1024, 1536):
    A = dv.DevArray.from_host(rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k)))
    for rep in range(2):
        lib.tnb_profile_enable(1)
        try:
            dv.svd_project(A)
        except Exception as e:
            pass
        torch.cuda.synchronize()
        ms, w = ctypes.c_double(), ctypes.c_double(); l, s = ctypes.c_longlong(), ctypes.c_longlong()
        lib.tnb_profile_get(1, ctypes.byref(ms), ctypes.byref(w), ctypes.byref(l), ctypes.byref(s))
        lib.tnb_profile_enable(0)
        if rep:
            print(os.environ.get("TNB_LIB_PATH", "default")[-22:], "k", k, "fixed", os.environ.get("TNB_JACOBI_FIXED_SWEEPS"),
                  "jacobi %.3f ms, %d launches, %.2f us/round" % (ms.value, l.value, 1e3 * ms.value / max(l.value, 1)), flush=True)
