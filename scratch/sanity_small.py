"""Small QR / SVD calls that reach every new kernel path (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv
rng = np.random.default_rng(0)
def rn(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
a = rn(1100, 330)                      # m >= 1024: split-K scratch, side stream, lagged groups (256 + 74)
q, r = dv.qr(dv.DevArray.from_host(a))
q, r = np.asarray(q), np.asarray(r)
print("qr", np.linalg.norm(q @ r - a) / np.linalg.norm(a), np.linalg.norm(q.conj().T @ q - np.eye(330)))
# graded columns (cond ~ 1e3): the first pass leaves |G - I| ~ eps cond^2 >> 1e-14, so the second pass of the groups is NOT skipped
g = rn(1100, 330) * np.logspace(0, -3, 330)[None, :]
g = g @ np.linalg.qr(rn(330, 330))[0]
q, r = dv.qr(dv.DevArray.from_host(g))
q, r = np.asarray(q), np.asarray(r)
print("qr graded", np.linalg.norm(q @ r - g) / np.linalg.norm(g), np.linalg.norm(q.conj().T @ q - np.eye(330)))
b = rn(300, 200)
q, r = dv.qr(dv.DevArray.from_host(b))
print("qr small", np.linalg.norm(np.asarray(q) @ np.asarray(r) - b) / np.linalg.norm(b))
for shape in ((260, 200), (96, 130), (64, 64)):
    c = rn(*shape)
    u, s, vh = dv.svd(dv.DevArray.from_host(c))
    u, s, vh = np.asarray(u), np.asarray(s), np.asarray(vh)
    print("svd", shape, np.linalg.norm((u * s) @ vh - c) / np.linalg.norm(c), np.max(np.abs(s - np.linalg.svd(c, compute_uv=False))) / s[0])
    up, sp, p = dv.svd_project(dv.DevArray.from_host(c))
    print("svd_project", shape, np.linalg.norm(np.asarray(up) @ np.asarray(p) - c) / np.linalg.norm(c))
torch.cuda.synchronize()
