"""GPU: the BASELINE.json configurations AT FULL SIZE against values the unmodified reference produced
(tests/golden/make_golden_fullsize.py -> tests/golden/fullsize_*.npz): per-bond singular values, kept bond
dimensions, label lists, norms, overlaps, energies.  Tolerance 1e-10 relative (BASELINE.json north_star) on
floating point; bond dimensions and labels exact.  These sizes drive the many-column QR path (k >= 128), the
4-CTA-cluster Jacobi rounds and the split-K GEMMs inside a golden-compared sweep."""
import os
import sys

import numpy as np
import pytest

from golden_io import Golden

pytestmark = pytest.mark.gpu
TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class SvdSpy:
    """Record s from every SVD the sweeps run (both entry points of the device layer)."""

    def __enter__(self):
        from tncontract_b200 import devarray as dv
        self.dv, self.seen = dv, []
        self.orig, self.orig_p = dv.svd, dv.svd_project

        def spy(fn):
            def f(a):
                out = fn(a)
                self.seen.append(np.asarray(out[1]))
                return out
            return f
        dv.svd, dv.svd_project = spy(self.orig), spy(self.orig_p)
        return self

    def __exit__(self, *exc):
        self.dv.svd, self.dv.svd_project = self.orig, self.orig_p


def check_singular_values(seen, g, key):
    assert len(seen) == g.meta[key + ".nsvd"]
    worst = 0.0
    for i, s in enumerate(seen):
        ref = g.scalar("%s.s%d" % (key, i))
        assert s.shape == ref.shape, (key, i, s.shape, ref.shape)
        worst = max(worst, float(np.max(np.abs(s - ref)) / ref[0]))
    assert worst <= TOL, (key, worst)
    return worst


def close(a, b, tol=TOL):
    return abs(complex(a) - complex(b)) <= tol * abs(complex(b))


def test_cfg2_full_size_vs_reference():
    """cfg 2: random MPS N=50, d=2, chi=64 float64 (np.random.seed(1)): left_canonise() singular values bond by
    bond, svd_compress(chi=32) singular values / bonds / labels / norm / overlap."""
    import tncontract_b200 as tn
    od = tn.onedim
    g = Golden("fullsize_cfg2")
    np.random.seed(1)
    psi = od.init_mps_random(50, 2, 64)
    assert psi.bonddims() == g.meta["bonds0"]
    assert close(psi.norm(), g.scalar("norm0"))
    assert close(od.inner_product_mps(psi, psi), g.scalar("ip0"))
    with SvdSpy() as spy:
        a = psi.copy(); a.left_canonise()
    check_singular_values(spy.seen, g, "lc")
    assert a.bonddims() == g.meta["lc.bonds"]
    assert [list(t.labels) for t in a] == g.meta["lc.labels"]
    a = psi.copy(); a.right_canonise()
    assert a.bonddims() == g.meta["rc.bonds"]
    a = psi.copy(); a.left_canonise(qr_decomposition=True)
    assert a.bonddims() == g.meta["lcqr.bonds"]
    assert close(a.norm(canonical_form="left"), g.scalar("lcqr.norm"))
    with SvdSpy() as spy:
        c = psi.copy(); c.svd_compress(chi=32)
    check_singular_values(spy.seen, g, "comp")
    assert c.bonddims() == g.meta["comp.bonds"]
    assert [list(t.labels) for t in c] == g.meta["comp.labels"]
    assert close(c.norm(), g.scalar("comp.norm"))
    assert close(od.inner_product_mps(psi, c), g.scalar("comp.overlap"))


def test_cfg3_headline_sweep_vs_reference():
    """cfg 3 = the bench.py workload of rank 0 (seed 2): N=100, d=2, chi=512 complex128.  TFI MPO apply, energy,
    svd_compress(chi=512): all 99 bonds' singular values (up to 1024 each), bonds, labels, norm, overlap, energy of
    the compressed state -- against one full sweep of the reference."""
    import tncontract_b200 as tn
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    od = tn.onedim
    g = Golden("fullsize_cfg3")
    n, d, chi = 100, 2, 512
    sites = bench.make_host_sites(n, d, chi, seed=2)
    psi = od.MatrixProductState([tn.Tensor(a, ["phys", "left", "right"]) for a in sites])
    psi.left_canonise(qr_decomposition=True, normalise=True)
    assert psi.bonddims() == g.meta["psi.bonds"]
    assert [list(t.labels) for t in psi] == g.meta["psi.labels"]
    W = bench.tfi_w()
    ws = [W[2] if i == 0 else (W[:, 0] if i == n - 1 else W) for i in range(n)]
    wl = [["right", "physout", "physin"] if i == 0 else (["left", "physout", "physin"] if i == n - 1 else
                                                         ["left", "right", "physout", "physin"]) for i in range(n)]
    H = od.MatrixProductOperator([tn.Tensor(w, l) for w, l in zip(ws, wl)], "left", "right", "physout", "physin")
    phi = od.contract_mps_mpo(psi, H)
    assert phi.bonddims() == g.meta["phi.bonds"]
    assert [list(t.labels) for t in phi] == g.meta["phi.labels"]
    pp = od.inner_product_mps(psi, psi)
    assert close(od.inner_product_mps(psi, phi) / pp, g.scalar("energy"))
    assert close(phi.norm(), g.scalar("phi.norm"))
    with SvdSpy() as spy:
        c = phi.copy(); c.svd_compress(chi=chi)
    worst = check_singular_values(spy.seen, g, "comp")
    assert c.bonddims() == g.meta["comp.bonds"]
    assert [list(t.labels) for t in c] == g.meta["comp.labels"]
    assert close(c.norm(), g.scalar("comp.norm"))
    assert close(c.norm(canonical_form="right"), g.scalar("comp.norm_right"))
    assert close(od.inner_product_mps(phi, c), g.scalar("comp.overlap"))
    assert close(od.inner_product_mps(psi, c) / pp, g.scalar("comp.energy"))
    print("cfg3 worst singular-value error (relative to s0):", worst)


def test_cfg4_networks_vs_reference():
    """cfg 4: networks 0..3 of seed 3 (N=64, d=4, chi=128 float64, counter-based generator): overlap, norm,
    svd_compress(chi=64) singular values / bonds / labels / norms -- single-network path."""
    import tncontract_b200 as tn
    from tncontract_b200 import batch
    od = tn.onedim
    g = Golden("fullsize_cfg4")
    for net in g.meta["networks"]:
        a = batch.random_mps(g.meta["seed"], net, 0, 64, 4, 128)
        b = batch.random_mps(g.meta["seed"], net, 1, 64, 4, 128)
        assert close(od.inner_product_mps(a, b), g.scalar("n%d.overlap" % net))
        assert close(a.norm(), g.scalar("n%d.norm" % net))
        with SvdSpy() as spy:
            a.svd_compress(chi=64)
        check_singular_values(spy.seen, g, "n%d.comp" % net)
        assert a.bonddims() == g.meta["n%d.bonds" % net]
        assert [list(t.labels) for t in a] == g.meta["n%d.labels" % net]
        assert close(a.norm(), g.scalar("n%d.norm_after" % net))
        assert close(a.norm(canonical_form="right"), g.scalar("n%d.norm_after_right" % net))


def test_cfg5_reduced_vs_reference():
    """cfg 5 reduced: PEPS 6x6, D=3 (double-layer bond 9), boundary chi=32, tolerance 1e-14: per-column bond
    dimensions exact, scalar to 1e-10, float128 result dtype kept."""
    import tncontract_b200 as tn
    td = tn.twodim
    g = Golden("fullsize_cfg5r")
    L, D, d, chi = g.meta["L"], g.meta["D"], g.meta["d"], g.meta["chi"]
    rng = np.random.default_rng(g.meta["seed"])
    grid = []
    for r in range(L):
        row = []
        for c in range(L):
            shape = (d, 1 if r == 0 else D, 1 if r == L - 1 else D, 1 if c == 0 else D, 1 if c == L - 1 else D)
            row.append(tn.Tensor(rng.standard_normal(shape) / D, ["phys", "up", "down", "left", "right"]))
        grid.append(row)
    peps = td.SquareLatticePEPS(grid)
    net = td.inner_product_peps(peps, peps, contract_virtual=False)
    cols = net.mps_contract(chi, return_all_columns=True, tolerance=1e-14)
    assert [c.bonddims() for c in cols[:-1]] == g.meta["col.bonds"]
    val = net.mps_contract(chi, tolerance=1e-14)
    assert str(np.asarray(val.data).dtype) == g.meta["result_dtype"]
    assert close(float(np.asarray(val.data)), float(g.scalar("approx")))
