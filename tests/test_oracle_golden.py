"""CPU: the oracle (oracle/tn_oracle.py) against fixtures produced by the real
reference (tests/golden/make_golden.py).  Labels / shapes / bond dimensions
exact; numbers to 1e-12 relative (same LAPACK underneath, so factors agree
too, not just gauge invariants)."""
import numpy as np
import pytest

from golden_io import Golden, rel_err
from oracle import tn_oracle as o

TOL = 1e-12


def OT(pair):
    return o.OT(pair[0], pair[1])


def check_tensor(t, gold, tol=TOL):
    data, labels = gold
    assert t.labels == labels
    assert tuple(t.shape) == tuple(data.shape)
    assert rel_err(t.data, data) <= tol


def to_chain(g, key):
    sites, m = g.chain(key)
    return o.Chain([OT(s) for s in sites], m["left"], m["right"], m.get("phys_label", "phys"),
                   m.get("physout_label", "physout"), m.get("physin_label", "physin")), m


def check_chain(ch, g, key, tol=TOL):
    sites, m = g.chain(key)
    assert ch.bonddims() == m["bonds"]
    assert (ch.left, ch.right) == (m["left"], m["right"])
    for t, gold in zip(ch.sites, sites):
        check_tensor(t, gold, tol)


def test_contract_cases():
    g = Golden("contract")
    for name in g.meta["cases"]:
        c = g.meta[name]
        out = o.contract(OT(g.tensor(name + ".A")), OT(g.tensor(name + ".B")), c["l1"], c["l2"], c["s1"], c["s2"])
        check_tensor(out, g.tensor(name + ".C"))


def test_consolidate_trace():
    g = Golden("contract")
    t = OT(g.tensor("cons.in"))
    check_tensor(o.consolidate(t), g.tensor("cons.all"))
    check_tensor(o.consolidate(t, ["l"]), g.tensor("cons.l"))
    check_tensor(o.trace_pair(OT(g.tensor("trace.in")), "i0", "i2"), g.tensor("trace.out"))


def test_factorisations():
    g = Golden("factor")
    for name in g.meta["cases"]:
        rows = g.meta[name]["rows"]
        t = OT(g.tensor(name + ".in"))
        U, S, V = o.tensor_svd(t, rows)
        for got, key in ((U, ".U"), (S, ".S"), (V, ".V")):
            check_tensor(got, g.tensor(name + key))
        Q, R = o.tensor_qr(t, rows)
        check_tensor(Q, g.tensor(name + ".Q"))
        check_tensor(R, g.tensor(name + ".R"))
        for mode in ("left", "right", "both"):
            Ut, Vt, cut = o.truncated_svd(t, rows, chi=2, absorb=mode)
            check_tensor(Ut, g.tensor("%s.t%s.U" % (name, mode)))
            check_tensor(Vt, g.tensor("%s.t%s.V" % (name, mode)))
            assert rel_err(cut, g.scalar("%s.t%s.cut" % (name, mode))) <= TOL
        Ut, St, Vt = o.truncated_svd(t, rows, chi=0, threshold=0.5, absorb=None, absolute=False)
        check_tensor(Ut, g.tensor(name + ".trel.U"))
        check_tensor(St, g.tensor(name + ".trel.S"))
        check_tensor(Vt, g.tensor(name + ".trel.V"))


def test_ring_and_con_examples():
    g = Golden("ring")
    A = OT(g.tensor("A"))
    N = 100
    ts = [o.OT(A.data, [l + str(i) for l in A.labels]) for i in range(N)]
    pairs = [("right%d" % j, "left%d" % (j + 1)) for j in range(N - 1)] + [("right%d" % (N - 1), "left0")]
    out = o.con(ts, pairs)
    check_tensor(out, g.tensor("out"))
    assert rel_err(out.data, g.scalar("trace_power")) < 1e-13
    a, b, c = (OT(g.tensor("ex." + k)) for k in "abc")
    check_tensor(o.con([a, b], [("a", "d"), ("c", "e")]), g.tensor("ex.pair"))
    check_tensor(o.con([c], [("f", "g")]), g.tensor("ex.internal"))
    check_tensor(o.con([a, b], []), g.tensor("ex.product"))
    check_tensor(o.con([a, b, c], [("a", "d"), ("c", "e"), ("f", "g"), ("h", "b")]), g.tensor("ex.network"))


def test_mps_real_sweeps():
    g = Golden("mps_real")
    psi, _ = to_chain(g, "psi")
    assert rel_err(o.chain_norm(psi), g.scalar("psi.norm")) < TOL
    a = psi.copy(); o.left_canonise(a, qr=True); check_chain(a, g, "lc_qr")
    rec = []
    a = psi.copy(); o.left_canonise(a, record=rec); check_chain(a, g, "lc_svd")
    assert len(rec) == g.meta["lc_svd.nsvd"]
    for i, s in enumerate(rec):
        ref = g.scalar("lc_svd.s%d" % i)
        assert rel_err(s, ref / ref[0]) < TOL
    a = psi.copy(); o.right_canonise(a); check_chain(a, g, "rc_svd")
    a = psi.copy(); o.right_canonise(a, qr=True, normalise=True); check_chain(a, g, "rc_qr_n")
    a = psi.copy(); o.svd_compress(a, chi=8); check_chain(a, g, "comp8")
    assert rel_err(o.inner_product_mps(psi, a), g.scalar("comp8.overlap")) < TOL
    a = psi.copy(); o.svd_compress(a, chi=4, reverse=True, normalise=True); check_chain(a, g, "comp4_rev_n")
    b = o.svd_compress_mps(psi, 6); check_chain(b, g, "compmps6")
    a = psi.copy(); o.left_canonise(a, 2, 7); check_chain(a, g, "lc_seg")
    a = psi.copy(); o.right_canonise(a, 3, 9); check_chain(a, g, "rc_seg")


def test_mps_complex_apply_compress_energy():
    g = Golden("mps_complex")
    raw, _ = to_chain(g, "raw")
    psi = raw.copy(); o.left_canonise(psi, qr=True, normalise=True); check_chain(psi, g, "psi")
    H, _ = to_chain(g, "H")
    check_chain(H, g, "H")
    phi = o.contract_mps_mpo(psi, H)
    check_chain(phi, g, "phi")
    assert phi.phys == g.meta["phi"]["phys_label"]
    e = o.inner_product_mps(psi, phi) / o.inner_product_mps(psi, psi)
    assert rel_err(e, g.scalar("energy")) < TOL
    c = phi.copy(); o.svd_compress(c, chi=8); check_chain(c, g, "phi_comp")
    assert rel_err(o.inner_product_mps(phi, c), g.scalar("phi_comp.overlap")) < TOL
    inter = o.ladder_contract(psi, phi, "phys", "physout", conj_a=True, intermediates=True)
    assert len(inter) == g.meta["ladder.inter.n"]
    for i, t in enumerate(inter):
        check_tensor(t, g.tensor("ladder.inter.%d" % i))
    check_tensor(o.ladder_contract(psi, phi, "phys", "physout", start=1, end=3), g.tensor("ladder.mid"))
    check_tensor(o.ladder_contract(psi, phi, "phys", "physout", start=2, end=len(psi) - 1), g.tensor("ladder.right"))


def test_reference_fixture_10site():
    """The pickled MPS of the reference's own test-suite; values also listed in SURVEY.md section 4."""
    g = Golden("fixture10")
    psi, m = to_chain(g, "psi")
    assert psi.bonddims() == [1, 3, 3, 3, 3, 3, 3, 3, 3, 3, 1]
    assert abs(o.chain_norm(psi) - 6.8326810768514090e-01) < 1e-13
    assert abs(o.inner_product_mps(psi, psi) - 0.46685530697963323) < 1e-13
    a = psi.copy(); o.svd_compress(a, threshold=1e-12); check_chain(a, g, "comp")
    # reference test: svd_compress preserves the norm to 10 decimals
    np.testing.assert_almost_equal(o.chain_norm(a), o.chain_norm(psi), decimal=10)
    b = psi.copy(); o.svd_compress(b, chi=2); check_chain(b, g, "comp2")
    ov = o.inner_product_mps(psi, b) / (o.chain_norm(psi) * o.chain_norm(b))
    assert abs(ov - 0.9996930862857119) < 1e-12


def test_peps_boundary_contraction():
    g = Golden("peps")
    L, chi = g.meta["L"], g.meta["chi"]
    peps = [[OT(g.tensor("peps.%d.%d" % (r, c))) for c in range(L)] for r in range(L)]
    net = o.inner_product_peps_network(peps, peps)
    for r in range(L):
        for c in range(L):
            check_tensor(net[r][c], g.tensor("net.%d.%d" % (r, c)))
    bonds = []
    val = o.boundary_mps_contract(net, chi, bond_log=bonds)
    assert val.dtype == np.longdouble
    assert rel_err(np.float64(val), g.scalar("approx")) < 1e-11
    for i, b in enumerate(bonds):
        assert b == g.meta["col.%d" % i]["bonds"]
    full = o.boundary_mps_contract(net, 81)
    assert rel_err(np.float64(full), g.scalar("exact")) < 1e-10
    mpo1 = o.column_chain([[o._with_dummy(o._with_dummy(o._with_dummy(o._with_dummy(t, "left"), "right"), "up"), "down")
                            for t in row] for row in net], 1)
    check_chain(mpo1, g, "colmpo1")
