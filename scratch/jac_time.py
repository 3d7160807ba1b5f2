"""Jacobi round timing through the device profile: TNB_LIB_PATH selects the library variant.
Prints class time, launches, sweeps and us/round for a projection SVD of a random k x k complex128 matrix."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv, _lib
lib = _lib.load()
rng = np.random.default_rng(0)
for k, cplx in ((1024, True), (512, True), (1024, False), (128, False)):
    a = rng.standard_normal((k, k)) + (1j * rng.standard_normal((k, k)) if cplx else 0)
    A = dv.DevArray.from_host(a)
    sref = np.linalg.svd(a, compute_uv=False)
    for rep in range(3):
        lib.tnb_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); u, s, p = dv.svd_project(A); e1.record()
        torch.cuda.synchronize()
        ms, w = ctypes.c_double(), ctypes.c_double(); l, sc = ctypes.c_longlong(), ctypes.c_longlong()
        lib.tnb_profile_get(1, ctypes.byref(ms), ctypes.byref(w), ctypes.byref(l), ctypes.byref(sc))
        lib.tnb_profile_enable(0)
    err = np.max(np.abs(np.asarray(s) - sref)) / sref[0]
    print(os.environ.get("TNB_LIB_PATH", "default")[-24:], "k", k, "c128" if cplx else "f64",
          "svd %.3f ms | jacobi %.3f ms, %d launches, sweeps %d, %.2f us/launch, work %.3e, sv err %.1e" %
          (e0.elapsed_time(e1), ms.value, l.value, dv.last_svd_sweeps, 1e3 * ms.value / max(l.value, 1), w.value, err), flush=True)
