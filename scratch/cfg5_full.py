"""cfg 5 at its stated size on one B200: PEPS 8x8, D=4, d=2, boundary-MPS contraction with chi=256 (float64).
python scratch/cfg5_full.py [L D chi]   -> JSON line with time, per-column bonds, the scalar"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tncontract_b200 as tn
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4
chi = int(sys.argv[3]) if len(sys.argv) > 3 else 256
d = 2
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
torch.cuda.synchronize()
l0 = tn.launch_count()
t0 = time.perf_counter()
cols = net.mps_contract(chi, return_all_columns=True, tolerance=1e-14)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
val = cols[-1]
print(json.dumps({"workload": "cfg5: PEPS %dx%d D=%d d=2 boundary chi=%d float64" % (L, L, D, chi), "seconds": dt,
                  "launches": tn.launch_count() - l0, "value": repr(np.asarray(val.data).item()),
                  "dtype": str(np.asarray(val.data).dtype), "col_bonds": [c.bonddims() for c in cols[:-1]],
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
