import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv
rng = np.random.default_rng(0)
for m, n in [(5400, 40), (3300, 64)]:
    a = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    q, r = dv.qr(dv.DevArray.from_host(a))
    q, r = np.asarray(q), np.asarray(r)
    print(m, n, "resid %.2e orth %.2e" % (np.linalg.norm(q @ r - a) / np.linalg.norm(a), np.linalg.norm(q.conj().T @ q - np.eye(n))))
