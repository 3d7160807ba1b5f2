"""Projection SVD of tall / square matrices: time and Jacobi sweeps (random, rank-deficient, graded)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv
rng = np.random.default_rng(5)
def rn(m, n, c): return rng.standard_normal((m, n)) + (1j * rng.standard_normal((m, n)) if c else 0)
cases = [("random c128 1536x1024", rn(1536, 1024, True)), ("random c128 1024x1024", rn(1024, 1024, True)),
         ("random f64 2048x1024", rn(2048, 1024, False)),
         ("rank 300 c128 1536x1024", rn(1536, 300, True) @ rn(300, 1024, True)),
         ("rank 500 f64 2048x2048", rn(2048, 500, False) @ rn(500, 2048, False)),
         ("graded 1e-10 c128 1536x1024", rn(1536, 1024, True) * np.logspace(0, -10, 1024)[None, :]),
         ("random c128 1024x1536 (wide, unchanged)", rn(1024, 1536, True))]
for name, a in cases:
    d = dv.DevArray.from_host(a)
    dv.svd_project(d); torch.cuda.synchronize()
    t0 = time.perf_counter(); u, s, p = dv.svd_project(d); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    sw = int(dv.last_svd_sweeps)
    u, s, p = np.asarray(u), np.asarray(s), np.asarray(p)
    sref = np.linalg.svd(a, compute_uv=False)
    k = min(a.shape)
    print("%-42s %8.2f ms  sweeps %2d  sv err %.1e  |U^H U - I|max %.1e  recon %.1e" % (
        name, dt * 1e3, sw, np.max(np.abs(s - sref)) / sref[0], np.max(np.abs(u.conj().T @ u - np.eye(k))),
        np.linalg.norm(u @ p - a) / np.linalg.norm(a)), flush=True)
