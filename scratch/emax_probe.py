"""How orthonormal is a 256-column group after the first Gram-Schmidt pass?  (experiment build -DTNB_EXP_EMAX)
Runs the cfg 3 sweep on a short chain and prints the distribution of max |C2| and max |G0 - I| over the second passes."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import tncontract_b200 as tn
od = tn.onedim
lib = ctypes.CDLL(os.environ["TNB_LIB_PATH"])
n, d, chi = 24, 2, 512
sites = bench.make_host_sites(n, d, chi, seed=2)
w_host, w_labels = bench.mpo_host(n)
psi = od.MatrixProductState([tn.Tensor(a, ["phys", "left", "right"]) for a in sites], "left", "right", "phys")
H = od.MatrixProductOperator([tn.Tensor(np.ascontiguousarray(w), l) for w, l in zip(w_host, w_labels)], "left", "right", "physout", "physin")
psi.left_canonise(qr_decomposition=True, normalise=True)
buf = (ctypes.c_double * 8192)()
lib.tnb_debug_emax(buf, 4096)
phi = od.contract_mps_mpo(psi, H); phi.svd_compress(chi=chi)
torch.cuda.synchronize()
k = lib.tnb_debug_emax(buf, 4096)
v = np.array(buf[:2 * k]).reshape(k, 2)
print("second passes recorded:", k)
for name, col in (("max|C2| (against earlier groups)", v[:, 0]), ("max|G0 - I| (inside the group)", v[:, 1])):
    c = col[col > 0]
    print(name, "min %.1e median %.1e 90%% %.1e max %.1e" % (c.min(), np.median(c), np.quantile(c, 0.9), c.max()))
    for t in (1e-15, 3e-15, 1e-14, 3e-14, 1e-13, 1e-12):
        print("   <= %.0e: %5.1f %%" % (t, 100.0 * np.mean(col <= t)))
