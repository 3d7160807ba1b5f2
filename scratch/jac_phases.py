import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv
rng = np.random.default_rng(0)
A = dv.DevArray.from_host(rng.standard_normal((1024, 1024)) + 1j * rng.standard_normal((1024, 1024)))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = dv.launch_count()
    try:
        dv.svd_project(A)
    except Exception as e:
        pass
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(os.environ.get("TNB_LIB_PATH", "default")[-22:], "fixed", os.environ.get("TNB_JACOBI_FIXED_SWEEPS"), "%.2f ms" % (dt * 1e3), "launches", dv.launch_count() - l0, flush=True)
