#!/usr/bin/env python
"""Golden fixtures at the BASELINE.json sizes, written by the UNMODIFIED reference.

Run in the build container only (needs /root/reference; cfg 3 takes several
minutes of CPU):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fullsize.py [cfg2 cfg3 cfg4 cfg5r]

Same harness rules as make_golden.py (NumPy-2 shim before the import, nothing
else patched, ``np.linalg.svd`` wrapped only to record the singular values the
reference saw).  Only SMALL data is stored: per-bond singular values, kept bond
dimensions, label lists, norms, overlaps, energies -- the inputs are regenerated
from their seeds by the tests (bench.make_host_sites / batch.host_uniform /
numpy's seeded generators), exactly as here.

  fullsize_cfg2.npz   random MPS N=50 d=2 chi=64 float64 (np.random.seed(1); init_mps_random):
                      left_canonise() singular values, svd_compress(chi=32)
  fullsize_cfg3.npz   the bench.py workload for rank 0 (seed 2): N=100 d=2 chi=512 complex128,
                      TFI MPO apply, energy, svd_compress(chi=512)
  fullsize_cfg4.npz   cfg 4 networks 0..3 of seed 3 (N=64 d=4 chi=128 float64): overlap, norm,
                      svd_compress(chi=64)
  fullsize_cfg5r.npz  reduced cfg 5: PEPS 6x6, D=3, d=2, boundary chi=32 (rng 4)
"""
import json
import os
import sys
import time

import numpy as np

np.product = np.prod          # shim, see make_golden.py
np.float = float
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
import tncontract as tn  # noqa: E402
import tncontract.onedim as od  # noqa: E402
import tncontract.twodim as td  # noqa: E402

from make_golden import Bag, SvdSpy, tfi_mpo  # noqa: E402


def _labels(chain):
    return [[str(l) for l in t.labels] for t in chain]


def _spy_store(g, key, spy):
    g.meta[key + ".nsvd"] = len(spy.calls)
    for i, s in enumerate(spy.calls):
        g.scalar("%s.s%d" % (key, i), s)


def gen_cfg2():
    g = Bag("fullsize_cfg2")
    np.random.seed(1)
    psi = od.init_mps_random(50, 2, 64)
    g.meta["bonds0"] = [int(b) for b in psi.bonddims()]
    g.scalar("norm0", psi.norm())
    g.scalar("ip0", od.inner_product_mps(psi, psi))
    with SvdSpy() as spy:
        a = psi.copy(); a.left_canonise()
    _spy_store(g, "lc", spy)
    g.meta["lc.bonds"] = [int(b) for b in a.bonddims()]
    g.meta["lc.labels"] = _labels(a)
    a = psi.copy(); a.right_canonise()
    g.meta["rc.bonds"] = [int(b) for b in a.bonddims()]
    a = psi.copy(); a.left_canonise(qr_decomposition=True)
    g.meta["lcqr.bonds"] = [int(b) for b in a.bonddims()]
    g.scalar("lcqr.norm", a.norm(canonical_form="left"))
    with SvdSpy() as spy:
        c = psi.copy(); c.svd_compress(chi=32)
    _spy_store(g, "comp", spy)
    g.meta["comp.bonds"] = [int(b) for b in c.bonddims()]
    g.meta["comp.labels"] = _labels(c)
    g.scalar("comp.norm", c.norm())
    g.scalar("comp.overlap", od.inner_product_mps(psi, c))
    g.save()


def gen_cfg3():
    import bench
    g = Bag("fullsize_cfg3")
    n, d, chi = 100, 2, 512
    t0 = time.time()
    sites = bench.make_host_sites(n, d, chi, seed=2)
    psi = od.MatrixProductState([tn.Tensor(a, ["phys", "left", "right"]) for a in sites])
    psi.left_canonise(qr_decomposition=True, normalise=True)
    g.meta["psi.bonds"] = [int(b) for b in psi.bonddims()]
    g.meta["psi.labels"] = _labels(psi)
    H = tfi_mpo(n)
    phi = od.contract_mps_mpo(psi, H)
    g.meta["phi.bonds"] = [int(b) for b in phi.bonddims()]
    g.meta["phi.labels"] = _labels(phi)
    print("apply done %.0fs" % (time.time() - t0), flush=True)
    e = od.inner_product_mps(psi, phi) / od.inner_product_mps(psi, psi)
    g.scalar("energy", e)
    g.scalar("phi.norm", phi.norm())
    print("energy done %.0fs" % (time.time() - t0), flush=True)
    with SvdSpy() as spy:
        c = phi.copy(); c.svd_compress(chi=chi)
    print("compress done %.0fs" % (time.time() - t0), flush=True)
    _spy_store(g, "comp", spy)
    g.meta["comp.bonds"] = [int(b) for b in c.bonddims()]
    g.meta["comp.labels"] = _labels(c)
    g.scalar("comp.norm", c.norm())
    g.scalar("comp.norm_right", c.norm(canonical_form="right"))
    g.scalar("comp.overlap", od.inner_product_mps(phi, c))
    g.scalar("comp.energy", od.inner_product_mps(psi, c) / od.inner_product_mps(psi, psi))
    g.save()


def gen_cfg4():
    from tncontract_b200 import batch
    g = Bag("fullsize_cfg4")
    nets = [0, 1, 2, 3]
    g.meta["networks"] = nets
    g.meta["seed"] = 3
    for net in nets:
        a = od.MatrixProductState([tn.Tensor(t, l) for t, l in batch.random_mps(3, net, 0, 64, 4, 128, on_host=True)])
        b = od.MatrixProductState([tn.Tensor(t, l) for t, l in batch.random_mps(3, net, 1, 64, 4, 128, on_host=True)])
        g.scalar("n%d.overlap" % net, od.inner_product_mps(a, b))
        g.scalar("n%d.norm" % net, a.norm())
        with SvdSpy() as spy:
            a.svd_compress(chi=64)
        _spy_store(g, "n%d.comp" % net, spy)
        g.meta["n%d.bonds" % net] = [int(x) for x in a.bonddims()]
        g.meta["n%d.labels" % net] = _labels(a)
        g.scalar("n%d.norm_after" % net, a.norm())
        g.scalar("n%d.norm_after_right" % net, a.norm(canonical_form="right"))
    g.save()


def make_peps_host(L, D, d, seed):
    rng = np.random.default_rng(seed)
    grid = []
    for r in range(L):
        row = []
        for c in range(L):
            shape = (d, 1 if r == 0 else D, 1 if r == L - 1 else D, 1 if c == 0 else D, 1 if c == L - 1 else D)
            row.append(rng.standard_normal(shape) / D)
        grid.append(row)
    return grid


def gen_cfg5r():
    g = Bag("fullsize_cfg5r")
    L, D, d, chi = 6, 3, 2, 32
    g.meta.update({"L": L, "D": D, "d": d, "chi": chi, "seed": 4})
    grid = make_peps_host(L, D, d, 4)
    peps = td.SquareLatticePEPS([[tn.Tensor(a, ["phys", "up", "down", "left", "right"]) for a in row] for row in grid])
    net = td.inner_product_peps(peps, peps, contract_virtual=False)
    cols = net.mps_contract(chi, return_all_columns=True, tolerance=1e-14)
    g.meta["col.bonds"] = [[int(b) for b in c.bonddims()] for c in cols[:-1]]
    val = net.mps_contract(chi, tolerance=1e-14)
    g.meta["result_dtype"] = str(val.data.dtype)
    g.scalar("approx", np.asarray(val.data, dtype=np.float64))
    g.save()


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg4", "cfg5r", "cfg3"]
    for w in which:
        t0 = time.time()
        {"cfg2": gen_cfg2, "cfg3": gen_cfg3, "cfg4": gen_cfg4, "cfg5r": gen_cfg5r}[w]()
        print(w, "%.0fs" % (time.time() - t0), flush=True)
