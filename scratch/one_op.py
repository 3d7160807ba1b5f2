"""One isolated call of a site-level operation (for ncu captures): python scratch/one_op.py svd|qr|absorb"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tncontract_b200 import devarray as dv
rng = np.random.default_rng(0)
def rn(*s): return dv.DevArray.from_host(rng.standard_normal(s) + 1j * rng.standard_normal(s))
what = sys.argv[1]
if what == "svd":
    dv.svd_project(rn(1024, 1536))
elif what == "qr":
    dv.qr(rn(3072, 1536))
elif what == "absorb":
    dv.tensordot(rn(1536, 1536), rn(1536, 2, 1536), [1], [0])
    dv.tensordot(rn(512, 1536), rn(2, 1536, 1536), [1], [2])
torch.cuda.synchronize()
