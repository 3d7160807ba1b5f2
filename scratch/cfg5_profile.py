"""cfg 5 at its stated size (PEPS 8x8, D=4, boundary chi=256, float64): device time by kernel class (tnb_profile_*: CUDA
events around every launch, so the wall time of this pass is not a bench value) and Jacobi sweeps by matrix order."""
import collections, ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tncontract_b200 as tn
from tncontract_b200 import _lib, devarray as dv
L, D, chi, d = 8, 4, 256, 2
rng = np.random.default_rng(4)
grid = []
for r in range(L):
    row = []
    for c in range(L):
        shape = (d, 1 if r == 0 else D, 1 if r == L - 1 else D, 1 if c == 0 else D, 1 if c == L - 1 else D)
        row.append(tn.Tensor(rng.standard_normal(shape) / D, ["phys", "up", "down", "left", "right"]))
    grid.append(row)
peps = tn.twodim.SquareLatticePEPS(grid)
net = tn.twodim.inner_product_peps(peps, peps, contract_virtual=False)
lib = _lib.load()
KCLASS = ["gemm", "jacobi_round", "qr_panel", "permute", "mps_mpo_site", "elementwise"]
seen = collections.Counter()
calls = collections.defaultdict(list)
def wrap(name):
    orig = getattr(dv, name)
    def f(a_, *args, **kw):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = orig(a_, *args, **kw)
        torch.cuda.synchronize()
        calls[name].append((tuple(int(x) for x in a_.shape), time.perf_counter() - t0,
                            int(getattr(dv, "last_svd_sweeps", 0)) if "svd" in name else 0))
        return out
    setattr(dv, name, f)
for nm in ("svd_project", "svd", "qr"):
    if hasattr(dv, nm): wrap(nm)
torch.cuda.synchronize()
lib.tnb_profile_enable(1)
t0 = time.perf_counter()
cols = net.mps_contract(chi, return_all_columns=True, tolerance=1e-14)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
prof = {}
for c, name in enumerate(KCLASS):
    pms, pw = ctypes.c_double(), ctypes.c_double()
    pl, ps = ctypes.c_longlong(), ctypes.c_longlong()
    lib.tnb_profile_get(c, ctypes.byref(pms), ctypes.byref(pw), ctypes.byref(pl), ctypes.byref(ps))
    prof[name] = {"ms": round(pms.value, 1), "launches": pl.value, "rate": (pw.value / (pms.value * 1e-3) / 1e12 if name in ("gemm", "jacobi_round", "qr_panel") else pw.value / (pms.value * 1e-3) / 1e9) if pms.value > 0 else None}
lib.tnb_profile_enable(0)
top = {}
for nm, lst in calls.items():
    by = collections.defaultdict(lambda: [0, 0.0, 0])
    for shape, t, sw in lst:
        e = by[shape]; e[0] += 1; e[1] += t; e[2] += sw
    top[nm] = sorted(([list(k), v[0], round(v[1], 3), v[2]] for k, v in by.items()), key=lambda x: -x[2])[:6]
print(json.dumps({"workload": "cfg5 profiled pass", "seconds_profiled": dt, "value": repr(np.asarray(cols[-1].data).item()),
                  "kernel_classes": prof, "factorisations_top_by_time[shape, calls, s, sweeps]": top}))
