#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference is imported from /root/reference after the two-line NumPy-2
shim (np.product / np.float were removed in NumPy 2; tensor.py:629,817,910,
1040 still use them).  Nothing else is patched; ``np.linalg.svd`` is wrapped
*in this harness* only to record the singular values the reference saw.
Outputs: tests/golden/*.npz, each holding arrays plus a JSON ``meta`` string
(labels, shapes, bond dimensions, scalars).  The GPU box never runs this.
"""
import json
import os
import pickle
import sys

import numpy as np

np.product = np.prod          # shim, see module docstring
np.float = float
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import tncontract as tn  # noqa: E402
import tncontract.onedim as od  # noqa: E402
import tncontract.twodim as td  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class Bag:
    """Collects arrays + JSON metadata for one fixture file."""

    def __init__(self, name):
        self.name, self.arrays, self.meta = name, {}, {}

    def tensor(self, key, t):
        self.arrays[key] = np.asarray(t.data)
        self.meta[key] = {"labels": [str(l) for l in t.labels], "shape": list(t.data.shape)}

    def chain(self, key, c):
        for i, t in enumerate(c):
            self.tensor("%s.%d" % (key, i), t)
        m = {"n": len(c), "left": c.left_label, "right": c.right_label, "bonds": [int(b) for b in c.bonddims()]}
        for attr in ("phys_label", "physout_label", "physin_label"):
            if hasattr(c, attr):
                m[attr] = getattr(c, attr)
        self.meta[key] = m

    def scalar(self, key, v):
        v = np.asarray(v)
        self.arrays[key] = v.astype(np.complex128 if np.iscomplexobj(v) else np.float64)

    def save(self):
        np.savez_compressed(os.path.join(HERE, self.name + ".npz"), meta=json.dumps(self.meta), **self.arrays)
        print("wrote", self.name, "%d arrays" % len(self.arrays))


class SvdSpy:
    """Record s from every np.linalg.svd call made by the reference."""

    def __enter__(self):
        self.calls, self._orig = [], np.linalg.svd

        def spy(a, *args, **kw):
            out = self._orig(a, *args, **kw)
            self.calls.append(np.array(out[1]))
            return out
        np.linalg.svd = spy
        return self

    def __exit__(self, *exc):
        np.linalg.svd = self._orig


# ---------------------------------------------------------------- contract
def gen_contract():
    g = Bag("contract")
    rng = np.random.default_rng(10)

    def rt(shape, labels, cplx=False):
        d = rng.standard_normal(shape)
        if cplx:
            d = d + 1j * rng.standard_normal(shape)
        return tn.Tensor(d, labels)

    cases = [
        ("single", rt((2, 2), ["spam", "eggs"]), rt((2, 3, 2, 4), ["i0", "i1", "i2", "i3"]), "spam", "i2", None, None),
        ("double", rt((2, 2), ["spam", "eggs"]), rt((2, 3, 2, 4), ["i0", "i1", "i2", "i3"]), ["spam", "eggs"], ["i0", "i2"], None, None),
        ("dup", rt((3, 3, 4), ["x", "x", "y"]), rt((3, 5, 3), ["x", "z", "x"]), "x", "x", None, None),
        ("dupslice", rt((3, 3, 4), ["x", "x", "y"]), rt((3, 5, 3), ["x", "z", "x"]), "x", "x", [0], [-1]),
        ("outer", rt((2, 3), ["a", "b"]), rt((4,), ["c"]), [], [], None, None),
        ("middle", rt((5, 7), ["q", "right"], True), rt((2, 7, 6), ["phys", "left", "right"], True), "right", "left", None, None),
        ("trailing", rt((4, 6), ["s", "left"], True), rt((2, 5, 6), ["phys", "left", "right"], True), "left", "right", None, None),
        ("ladder2", rt((1, 1, 3, 2, 4), ["l1", "l2", "r2", "rung", "r1"], True), rt((3, 2, 5), ["left", "rung", "right"], True), ["r2", "rung"], ["left", "rung"], None, None),
        ("scalar", rt((3, 4), ["a", "b"]), rt((3, 4), ["c", "d"]), ["a", "b"], ["c", "d"], None, None),
        ("permuted", rt((3, 4, 5, 6), ["a", "b", "c", "d"], True), rt((6, 7, 4), ["d2", "e", "b2"], True), ["d", "b"], ["d2", "b2"], None, None),
    ]
    names = []
    for name, A, B, l1, l2, s1, s2 in cases:
        C = tn.contract(A, B, l1, l2, index_slice1=s1, index_slice2=s2)
        g.tensor(name + ".A", A)
        g.tensor(name + ".B", B)
        g.tensor(name + ".C", C)
        g.meta[name] = {"l1": l1, "l2": l2, "s1": s1, "s2": s2}
        names.append(name)
    g.meta["cases"] = names
    # consolidate / move / fuse / split / trace
    t = rt((3, 2, 3, 2), ["l", "p", "l", "r"], True)
    g.tensor("cons.in", t)
    c = t.copy(); c.consolidate_indices(); g.tensor("cons.all", c)
    c = t.copy(); c.consolidate_indices(labels=["l"]); g.tensor("cons.l", c)
    t = rt((2, 3, 4, 5, 6), ["a", "b", "c", "b", "d"])
    g.tensor("move.in", t)
    c = t.copy(); c.move_indices(["d", "b", "c"], 0, preserve_relative_order=True); g.tensor("move.keep", c)
    c = t.copy(); c.move_indices(["d", "b", "c"], 0); g.tensor("move.given", c)
    c = t.copy(); c.move_index("c", 4); g.tensor("move.one", c)
    t = rt((2, 3, 4, 5, 6), ["a", "b", "c", "d", "a"])
    g.tensor("fuse.in", t)
    c = t.copy(); c.fuse_indices(["b", "d"], "new_index"); g.tensor("fuse.out", c)
    c.split_index("new_index", (3, 5), ["b", "d"]); g.tensor("fuse.split", c)
    t = rt((2, 3, 2, 4), ["i0", "i1", "i2", "i3"], True)
    g.tensor("trace.in", t)
    c = t.copy(); c.trace("i0", "i2"); g.tensor("trace.out", c)
    a, b = rt((2, 3), ["u", "v"]), rt((3, 2), ["v", "u"])
    g.tensor("add.a", a); g.tensor("add.b", b); g.tensor("add.sum", a + b)
    g.scalar("distance", tn.distance(a, a * 1.5))
    g.save()


# ---------------------------------------------------------------- factorisations
def gen_factor():
    g = Bag("factor")
    rng = np.random.default_rng(11)
    names = []
    for name, shape, labels, rows, cplx in [
        ("f64_tall", (4, 2, 3), ["a", "b", "c"], ["c", "a"], False),
        ("c128_tall", (2, 6, 5), ["phys", "left", "right"], ["phys", "left"], True),
        ("c128_wide", (2, 3, 9), ["phys", "left", "right"], ["phys", "left"], True),
        ("f64_mid", (6, 2, 7), ["left", "phys", "right"], ["phys", "left"], False),
        ("c128_rank4", (3, 4, 2, 5), ["p", "l", "q", "r"], ["q", "l"], True),
    ]:
        d = rng.standard_normal(shape)
        if cplx:
            d = d + 1j * rng.standard_normal(shape)
        t = tn.Tensor(d, labels)
        g.tensor(name + ".in", t)
        U, S, V = tn.tensor_svd(t, rows)
        g.tensor(name + ".U", U); g.tensor(name + ".S", S); g.tensor(name + ".V", V)
        Q, R = tn.tensor.tensor_qr(t, rows)
        g.tensor(name + ".Q", Q); g.tensor(name + ".R", R)
        L, Q2 = tn.tensor.tensor_lq(t, rows)
        g.tensor(name + ".L", L); g.tensor(name + ".LQ", Q2)
        for mode in ("left", "right", "both"):
            Ut, Vt, cut = tn.truncated_svd(t, rows, chi=2, absorb_singular_values=mode)
            g.tensor("%s.t%s.U" % (name, mode), Ut); g.tensor("%s.t%s.V" % (name, mode), Vt)
            g.scalar("%s.t%s.cut" % (name, mode), cut)
        Ut, St, Vt = tn.truncated_svd(t, rows, chi=0, threshold=0.5, absorb_singular_values=None, absolute=False)
        g.tensor(name + ".trel.U", Ut); g.tensor(name + ".trel.S", St); g.tensor(name + ".trel.V", Vt)
        g.meta[name] = {"rows": rows}
        names.append(name)
    g.meta["cases"] = names
    g.save()


# ---------------------------------------------------------------- cfg1 ring
def gen_ring():
    g = Bag("ring")
    np.random.seed(0)
    N = 100
    A = tn.Tensor(np.random.rand(2, 2), labels=["left", "right"])
    ts = [A.suf(str(i)) for i in range(N)]
    pairs = [("right" + str(j), "left" + str(j + 1)) for j in range(N - 1)]
    out = tn.con(ts, pairs, ("right" + str(N - 1), "left0"))
    g.tensor("A", A)
    g.tensor("out", out)
    g.scalar("trace_power", np.trace(np.linalg.matrix_power(A.data, N)))
    # README tn.con examples (README.md:78-118)
    rng = np.random.default_rng(5)
    a = tn.Tensor(rng.random((3, 2, 4)), labels=["a", "b", "c"])
    b = tn.Tensor(rng.random((3, 4)), labels=["d", "e"])
    c = tn.Tensor(rng.random((5, 5, 2)), labels=["f", "g", "h"])
    g.tensor("ex.a", a); g.tensor("ex.b", b); g.tensor("ex.c", c)
    g.tensor("ex.pair", tn.con(a, b, ("a", "d"), ("c", "e")))
    g.tensor("ex.internal", tn.con(c, ("f", "g")))
    g.tensor("ex.product", tn.con(a, b))
    g.tensor("ex.network", tn.con(a, b, c, ("a", "d"), ("c", "e"), ("f", "g"), ("h", "b")))
    g.save()


# ---------------------------------------------------------------- cfg2 reduced
def gen_mps_real():
    g = Bag("mps_real")
    np.random.seed(1)
    psi = od.init_mps_random(12, 2, 16)
    g.chain("psi", psi)
    g.scalar("psi.norm", psi.norm())
    a = psi.copy(); a.left_canonise(qr_decomposition=True); g.chain("lc_qr", a)
    with SvdSpy() as spy:
        a = psi.copy(); a.left_canonise()
    g.chain("lc_svd", a)
    for i, s in enumerate(spy.calls):
        g.scalar("lc_svd.s%d" % i, s)
    g.meta["lc_svd.nsvd"] = len(spy.calls)
    a = psi.copy(); a.right_canonise(); g.chain("rc_svd", a)
    a = psi.copy(); a.right_canonise(qr_decomposition=True, normalise=True); g.chain("rc_qr_n", a)
    with SvdSpy() as spy:
        a = psi.copy(); a.svd_compress(chi=8)
    g.chain("comp8", a)
    for i, s in enumerate(spy.calls):
        g.scalar("comp8.s%d" % i, s)
    g.meta["comp8.nsvd"] = len(spy.calls)
    g.scalar("comp8.norm", a.norm())
    g.scalar("comp8.overlap", od.inner_product_mps(psi, a))
    a = psi.copy(); a.svd_compress(chi=4, reverse=True, normalise=True); g.chain("comp4_rev_n", a)
    g.scalar("comp4_rev_n.overlap", od.inner_product_mps(psi, a))
    b = od.svd_compress_mps(psi, 6); g.chain("compmps6", b)
    g.scalar("compmps6.overlap", od.inner_product_mps(psi, b))
    # segment canonisation (start/end) as expval/ptrace use it
    a = psi.copy(); a.left_canonise(2, 7); g.chain("lc_seg", a)
    a = psi.copy(); a.right_canonise(3, 9); g.chain("rc_seg", a)
    g.scalar("frob", od.frob_distance_squared(psi, b))
    g.save()


# ---------------------------------------------------------------- cfg3 reduced
def tfi_mpo(N, J=1.0, h=0.5):
    I = np.eye(2); Z = np.diag([1.0, -1.0]); X = np.array([[0.0, 1.0], [1.0, 0.0]])
    W = np.zeros((3, 3, 2, 2))
    W[0, 0] = I; W[1, 0] = Z; W[2, 0] = -h * X; W[2, 1] = -J * Z; W[2, 2] = I
    ts = []
    for i in range(N):
        if i == 0:
            ts.append(tn.Tensor(W[2], ["right", "physout", "physin"]))
        elif i == N - 1:
            ts.append(tn.Tensor(W[:, 0], ["left", "physout", "physin"]))
        else:
            ts.append(tn.Tensor(W, ["left", "right", "physout", "physin"]))
    return od.MatrixProductOperator(ts, "left", "right", "physout", "physin")


def gen_mps_complex():
    g = Bag("mps_complex")
    rng = np.random.default_rng(2)
    N, d, chi = 8, 2, 8
    bonds = [1] + [chi] * (N - 1) + [1]
    ts = [tn.Tensor(rng.standard_normal((d, bonds[i], bonds[i + 1])) +
                    1j * rng.standard_normal((d, bonds[i], bonds[i + 1])),
                    ["phys", "left", "right"]) for i in range(N)]
    raw = od.MatrixProductState(ts)
    g.chain("raw", raw)
    psi = raw.copy(); psi.left_canonise(qr_decomposition=True, normalise=True)
    g.chain("psi", psi)
    H = tfi_mpo(N)
    g.chain("H", H)
    phi = od.contract_mps_mpo(psi, H)
    g.chain("phi", phi)
    e = od.inner_product_mps(psi, phi) / od.inner_product_mps(psi, psi)
    g.scalar("energy", e)
    with SvdSpy() as spy:
        c = phi.copy(); c.svd_compress(chi=chi)
    g.chain("phi_comp", c)
    for i, s in enumerate(spy.calls):
        g.scalar("phi_comp.s%d" % i, s)
    g.meta["phi_comp.nsvd"] = len(spy.calls)
    g.scalar("phi_comp.overlap", od.inner_product_mps(phi, c))
    g.scalar("phi_comp.norm", c.norm())
    g.scalar("phi.norm", phi.norm())
    # ladder_contract variants (Appendix A of SURVEY.md)
    inter = od.ladder_contract(psi, phi, "phys", "physout", return_intermediate_contractions=True,
                               complex_conjugate_array1=True)
    for i, t in enumerate(inter):
        g.tensor("ladder.inter.%d" % i, t)
    g.meta["ladder.inter.n"] = len(inter)
    g.tensor("ladder.mid", od.ladder_contract(psi, phi, "phys", "physout", start=1, end=3))
    g.tensor("ladder.right", od.ladder_contract(psi, phi, "phys", "physout", start=2, end=N - 1))
    g.save()


# ---------------------------------------------------------------- reference test fixture
def gen_fixture():
    g = Bag("fixture10")
    with open("/root/reference/tncontract/tests/random_10site_mps.dat", "rb") as f:
        psi = pickle.load(f, encoding="latin1")
    g.chain("psi", psi)
    # the same object pickled again by the reference's classes under this Python (protocol 2): the interchange
    # fixture of tests/test_persist.py (the original .dat stays in /root/reference)
    with open(os.path.join(HERE, "ref_pickle_10site.dat"), "wb") as f:
        pickle.dump(psi, f, protocol=2)
    g.scalar("norm", psi.norm())
    g.scalar("ip", od.inner_product_mps(psi, psi))
    a = psi.copy(); a.svd_compress(threshold=1e-12, normalise=False)
    g.chain("comp", a)
    g.scalar("comp.norm", a.norm())
    can = od.right_canonical_to_canonical(a, threshold=1e-12)
    g.chain("canon", can)
    g.scalar("canon.norm", can.norm())
    l, r, n = can.check_canonical_form(threshold=1e-10, print_output=False)
    g.meta["canon.check"] = [list(map(int, l)), list(map(int, r)), list(map(int, n))]
    b = psi.copy(); b.svd_compress(chi=2)
    g.chain("comp2", b)
    g.scalar("comp2.overlap_normalised", od.inner_product_mps(psi, b) / (psi.norm() * b.norm()))
    g.save()


# ---------------------------------------------------------------- cfg5 reduced
def gen_peps():
    g = Bag("peps")
    rng = np.random.default_rng(4)
    L, D, d, chi = 4, 3, 2, 12
    grid = []
    for r in range(L):
        row = []
        for c in range(L):
            shape = (d, 1 if r == 0 else D, 1 if r == L - 1 else D, 1 if c == 0 else D, 1 if c == L - 1 else D)
            row.append(tn.Tensor(rng.standard_normal(shape) / D, ["phys", "up", "down", "left", "right"]))
        grid.append(row)
    peps = td.SquareLatticePEPS(grid)
    for r in range(L):
        for c in range(L):
            g.tensor("peps.%d.%d" % (r, c), peps[r, c])
    g.meta["L"] = L; g.meta["chi"] = chi
    net = td.inner_product_peps(peps, peps, contract_virtual=False)
    for r in range(L):
        for c in range(L):
            g.tensor("net.%d.%d" % (r, c), net[r, c])
    exact = net.exact_contract()
    g.scalar("exact", np.asarray(exact.data, dtype=np.float64))
    cols = net.mps_contract(chi, return_all_columns=True, tolerance=1e-14)
    for i, c in enumerate(cols[:-1]):
        g.chain("col.%d" % i, c)
    val = net.mps_contract(chi, tolerance=1e-14)
    g.meta["result_dtype"] = str(val.data.dtype)
    g.meta["result_labels"] = list(val.labels)
    g.scalar("approx", np.asarray(val.data, dtype=np.float64))
    full = td.inner_product_peps(peps, peps, exact_contract=False, chi=81)
    g.scalar("approx_fullchi", np.asarray(full.data, dtype=np.float64))
    mpo1 = td.column_to_mpo(net, 1)
    g.chain("colmpo1", mpo1)
    g.save()


if __name__ == "__main__":
    gen_contract()
    gen_factor()
    gen_ring()
    gen_mps_real()
    gen_mps_complex()
    gen_fixture()
    gen_peps()
