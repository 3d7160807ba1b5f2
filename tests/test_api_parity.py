"""tncontract_b200's public API against the golden fixtures produced by the
reference (tests/golden/make_golden.py), on two backends (fixture ``backend``):

* host: the Python layer over a NumPy stand-in of the C ABI -- label algebra,
  shapes, sweep logic; factors agree elementwise (same LAPACK as the reference);
* gpu : the real CUDA path.  Labels, shapes and kept bond dimensions exact;
  contracted tensors, singular values and gauge-invariant quantities (norms,
  overlaps, energies) to 1e-10 relative (north_star tolerance); QR/SVD factors
  are gauge-dependent and are checked through products and isometry.
"""
import numpy as np
import pytest

from golden_io import Golden, rel_err

TOL = 1e-10       # north_star bound for the GPU path
TOL_HOST = 1e-12


def tn_():
    import tncontract_b200 as tn
    return tn


def T(pair):
    return tn_().Tensor(pair[0], pair[1])


def tol(backend):
    return TOL_HOST if backend == "host" else TOL


def check_tensor(t, gold, tolerance, values=True):
    data, labels = gold
    assert list(t.labels) == labels
    assert tuple(t.shape) == tuple(data.shape)
    if values:
        assert rel_err(np.asarray(t.data), data) <= tolerance


def to_chain(g, key):
    tn = tn_()
    sites, m = g.chain(key)
    ts = [T(s) for s in sites]
    if "physout_label" in m:
        return tn.onedim.MatrixProductOperator(ts, m["left"], m["right"], m["physout_label"], m["physin_label"])
    return tn.onedim.MatrixProductState(ts, m["left"], m["right"], m.get("phys_label", "phys"))


def check_chain(ch, g, key, backend, values=None):
    """Labels / shapes / bonds exact.  Site tensors elementwise on the host
    backend (or when ``values``); on the GPU the state as a whole is compared:
    <gold|ch> and <ch|ch> against <gold|gold>."""
    tn = tn_()
    sites, m = g.chain(key)
    assert [int(b) for b in ch.bonddims()] == m["bonds"]
    assert (ch.left_label, ch.right_label) == (m["left"], m["right"])
    elementwise = (backend == "host") if values is None else values
    for t, gold in zip(ch, sites):
        check_tensor(t, gold, tol(backend), values=elementwise)
    if not elementwise and "physout_label" not in m:
        gold = to_chain(g, key)
        gg = tn.onedim.inner_product_mps(gold, gold)
        ga = tn.onedim.inner_product_mps(gold, ch)
        aa = tn.onedim.inner_product_mps(ch, ch)
        assert abs(ga - gg) <= TOL * abs(gg)
        assert abs(aa - gg) <= TOL * abs(gg)


def test_contract_cases(backend):
    tn = tn_()
    g = Golden("contract")
    for name in g.meta["cases"]:
        c = g.meta[name]
        A, B = T(g.tensor(name + ".A")), T(g.tensor(name + ".B"))
        out = tn.contract(A, B, c["l1"], c["l2"], index_slice1=c["s1"], index_slice2=c["s2"])
        check_tensor(out, g.tensor(name + ".C"), tol(backend))
        if c["s1"] is None and not isinstance(c["l1"], list):
            check_tensor(A[c["l1"]] * B[c["l2"]], g.tensor(name + ".C"), tol(backend))
        # inputs untouched
        check_tensor(A, g.tensor(name + ".A"), 0.0)


def test_index_plumbing(backend):
    tn = tn_()
    g = Golden("contract")
    t = T(g.tensor("cons.in"))
    c = t.copy(); c.consolidate_indices(); check_tensor(c, g.tensor("cons.all"), 0.0)
    c = t.copy(); c.consolidate_indices(labels=["l"]); check_tensor(c, g.tensor("cons.l"), 0.0)
    t = T(g.tensor("move.in"))
    c = t.copy(); c.move_indices(["d", "b", "c"], 0, preserve_relative_order=True); check_tensor(c, g.tensor("move.keep"), 0.0)
    c = t.copy(); c.move_indices(["d", "b", "c"], 0); check_tensor(c, g.tensor("move.given"), 0.0)
    c = t.copy(); c.move_index("c", 4); check_tensor(c, g.tensor("move.one"), 0.0)
    t = T(g.tensor("fuse.in"))
    c = t.copy(); c.fuse_indices(["b", "d"], "new_index"); check_tensor(c, g.tensor("fuse.out"), 0.0)
    c.split_index("new_index", (3, 5), ["b", "d"]); check_tensor(c, g.tensor("fuse.split"), 0.0)
    t = T(g.tensor("trace.in"))
    c = t.copy(); c.trace("i0", "i2"); check_tensor(c, g.tensor("trace.out"), 1e-14)
    a, b = T(g.tensor("add.a")), T(g.tensor("add.b"))
    check_tensor(a + b, g.tensor("add.sum"), 1e-15)
    assert abs(tn.distance(a, a * 1.5) - g.scalar("distance")) < 1e-14
    # Tensor() owns a private copy; scalar products keep labels
    src = np.arange(6.0).reshape(2, 3)
    t = tn.Tensor(src, ["a", "b"])
    src[0, 0] = 99.0
    assert np.asarray(t.data)[0, 0] == 0.0
    assert (2 * t).labels == ["a", "b"] and np.array_equal(np.asarray((t * 2.0).data), 2 * np.arange(6.0).reshape(2, 3))
    with pytest.raises(ValueError):
        tn.Tensor(np.zeros((2, 2)), ["only_one"])
    with pytest.raises(ValueError):
        tn.contract(tn.Tensor(np.zeros((2, 3)), ["a", "b"]), tn.Tensor(np.zeros((4, 2)), ["c", "d"]), "b", "c")


def _isometry_defect(M, rows_first):
    M = np.asarray(M)
    k = M.shape[-1] if rows_first else M.shape[0]
    m = M.reshape(-1, k) if rows_first else M.reshape(k, -1).conj().T
    return np.linalg.norm(m.conj().T @ m - np.eye(k))


def test_factorisations(backend):
    tn = tn_()
    g = Golden("factor")
    exact = backend == "host"
    for name in g.meta["cases"]:
        rows = g.meta[name]["rows"]
        t = T(g.tensor(name + ".in"))
        U, S, V = tn.tensor_svd(t, rows)
        for got, key in ((U, ".U"), (S, ".S"), (V, ".V")):
            check_tensor(got, g.tensor(name + key), tol(backend), values=exact or key == ".S")
        # U S V reproduces the input (as a tensor with the rows first)
        rec = tn.contract(tn.contract(U, S, "svd_in", "svd_out"), V, "svd_in", "svd_out")
        want = t.copy(); want.move_indices(rows, 0)
        assert rec.labels == want.labels and rel_err(np.asarray(rec.data), np.asarray(want.data)) < TOL
        assert _isometry_defect(U.data, True) < 1e-11 and _isometry_defect(V.data, False) < 1e-11
        check_tensor(t, g.tensor(name + ".in"), 0.0)  # input untouched
        Q, R = tn.tensor_qr(t, rows)
        check_tensor(Q, g.tensor(name + ".Q"), tol(backend), values=exact)
        check_tensor(R, g.tensor(name + ".R"), tol(backend), values=exact)
        rec = tn.contract(Q, R, "qr_in", "qr_out")
        assert rec.labels == want.labels and rel_err(np.asarray(rec.data), np.asarray(want.data)) < TOL
        assert _isometry_defect(Q.data, True) < 1e-11
        L, Q2 = tn.tensor_lq(t, rows)
        check_tensor(L, g.tensor(name + ".L"), tol(backend), values=exact)
        check_tensor(Q2, g.tensor(name + ".LQ"), tol(backend), values=exact)
        for mode in ("left", "right", "both"):
            Ut, Vt, cut = tn.truncated_svd(t, rows, chi=2, absorb_singular_values=mode)
            check_tensor(Ut, g.tensor("%s.t%s.U" % (name, mode)), tol(backend), values=exact)
            check_tensor(Vt, g.tensor("%s.t%s.V" % (name, mode)), tol(backend), values=exact)
            assert rel_err(cut, g.scalar("%s.t%s.cut" % (name, mode))) <= tol(backend)
            # the truncated product is the gauge-invariant best rank-2 approximation
            prod = tn.contract(Ut, Vt, "svd_in", "svd_out")
            gU, gV = T(g.tensor("%s.t%s.U" % (name, mode))), T(g.tensor("%s.t%s.V" % (name, mode)))
            gprod = tn.contract(gU, gV, "svd_in", "svd_out")
            assert rel_err(np.asarray(prod.data), np.asarray(gprod.data)) < TOL
        Ut, St, Vt = tn.truncated_svd(t, rows, chi=0, threshold=0.5, absorb_singular_values=None, absolute=False)
        check_tensor(Ut, g.tensor(name + ".trel.U"), tol(backend), values=exact)
        check_tensor(St, g.tensor(name + ".trel.S"), tol(backend))
        check_tensor(Vt, g.tensor(name + ".trel.V"), tol(backend), values=exact)


def test_ring_and_con_examples(backend):
    tn = tn_()
    g = Golden("ring")
    A = T(g.tensor("A"))
    N = 100
    ts = [A.suf(str(i)) for i in range(N)]
    pairs = [("right" + str(j), "left" + str(j + 1)) for j in range(N - 1)]
    out = tn.con(ts, pairs, ("right" + str(N - 1), "left0"))
    check_tensor(out, g.tensor("out"), tol(backend))
    assert rel_err(np.asarray(out.data), g.scalar("trace_power")) < TOL
    a, b, c = (T(g.tensor("ex." + k)) for k in "abc")
    check_tensor(tn.con(a, b, ("a", "d"), ("c", "e")), g.tensor("ex.pair"), tol(backend))
    check_tensor(tn.con(c, ("f", "g")), g.tensor("ex.internal"), tol(backend))
    check_tensor(tn.con(a, b), g.tensor("ex.product"), tol(backend))
    check_tensor(tn.con(a, b, c, ("a", "d"), ("c", "e"), ("f", "g"), ("h", "b")), g.tensor("ex.network"), tol(backend))
    with pytest.raises(ValueError):
        tn.con(a, b, ("a", "d"), ("a", "e"))


def test_mps_real_sweeps(backend):
    tn = tn_()
    od = tn.onedim
    g = Golden("mps_real")
    psi = to_chain(g, "psi")
    check_chain(psi, g, "psi", backend, values=True)
    assert rel_err(psi.norm(), g.scalar("psi.norm")) < TOL
    a = psi.copy(); a.left_canonise(qr_decomposition=True); check_chain(a, g, "lc_qr", backend)
    a = psi.copy(); a.left_canonise(); check_chain(a, g, "lc_svd", backend)
    assert a.check_canonical_form(threshold=1e-9, print_output=False) == (len(a) - 1, len(a) - 1)
    a = psi.copy(); a.right_canonise(); check_chain(a, g, "rc_svd", backend)
    assert a.check_canonical_form(threshold=1e-9, print_output=False) == (0, 0)
    a = psi.copy(); a.right_canonise(qr_decomposition=True, normalise=True); check_chain(a, g, "rc_qr_n", backend)
    a = psi.copy(); a.svd_compress(chi=8); check_chain(a, g, "comp8", backend)
    assert rel_err(a.norm(), g.scalar("comp8.norm")) < TOL
    assert rel_err(od.inner_product_mps(psi, a), g.scalar("comp8.overlap")) < TOL
    a = psi.copy(); a.svd_compress(chi=4, reverse=True, normalise=True); check_chain(a, g, "comp4_rev_n", backend)
    assert rel_err(od.inner_product_mps(psi, a), g.scalar("comp4_rev_n.overlap")) < TOL
    b = od.svd_compress_mps(psi, 6); check_chain(b, g, "compmps6", backend)
    assert rel_err(od.inner_product_mps(psi, b), g.scalar("compmps6.overlap")) < TOL
    a = psi.copy(); a.left_canonise(2, 7); check_chain(a, g, "lc_seg", backend)
    a = psi.copy(); a.right_canonise(3, 9); check_chain(a, g, "rc_seg", backend)
    assert abs(od.frob_distance_squared(psi, b) - g.scalar("frob")) < 1e-10 * abs(g.scalar("psi.norm")) ** 2


def test_singular_values_per_bond(backend):
    """The singular values the reference's np.linalg.svd saw during left_canonise()
    and svd_compress(chi=8), bond by bond (recorded by SvdSpy in make_golden.py)."""
    tn = tn_()
    from tncontract_b200 import devarray as dv
    g = Golden("mps_real")
    seen = []
    orig, orig_p = dv.svd, dv.svd_project

    def spy(a, fn=None):
        out = (fn or orig)(a)
        seen.append(np.asarray(out[1]))
        return out

    dv.svd = spy
    dv.svd_project = lambda a: spy(a, orig_p)   # the sweeps use the projection form (U, s, diag(s) Vh)
    try:
        a = to_chain(g, "psi"); a.left_canonise()
        assert len(seen) == g.meta["lc_svd.nsvd"]
        for i, s in enumerate(seen):
            ref = g.scalar("lc_svd.s%d" % i)
            assert s.shape == ref.shape and np.max(np.abs(s - ref)) <= TOL * ref[0]
        del seen[:]
        a = to_chain(g, "psi"); a.svd_compress(chi=8)
        assert len(seen) == g.meta["comp8.nsvd"]
        for i, s in enumerate(seen):
            ref = g.scalar("comp8.s%d" % i)
            assert s.shape == ref.shape and np.max(np.abs(s - ref)) <= TOL * ref[0]
    finally:
        dv.svd, dv.svd_project = orig, orig_p


def test_mps_complex_apply_compress_energy(backend):
    tn = tn_()
    od = tn.onedim
    g = Golden("mps_complex")
    raw = to_chain(g, "raw")
    psi = raw.copy(); psi.left_canonise(qr_decomposition=True, normalise=True); check_chain(psi, g, "psi", backend)
    H = to_chain(g, "H")
    check_chain(H, g, "H", backend, values=True)
    phi = od.contract_mps_mpo(psi, H)
    # phi depends on psi's gauge: compare it elementwise only when psi itself is the golden one
    gpsi = to_chain(g, "psi")
    gphi = od.contract_mps_mpo(gpsi, H)
    check_chain(gphi, g, "phi", backend, values=True)
    check_chain(phi, g, "phi", backend)
    assert phi.phys_label == g.meta["phi"]["phys_label"]
    e = od.inner_product_mps(psi, phi) / od.inner_product_mps(psi, psi)
    assert rel_err(e, g.scalar("energy")) < TOL
    assert rel_err(phi.norm(), g.scalar("phi.norm")) < TOL
    c = phi.copy(); c.svd_compress(chi=8); check_chain(c, g, "phi_comp", backend)
    assert rel_err(od.inner_product_mps(phi, c), g.scalar("phi_comp.overlap")) < TOL
    assert rel_err(c.norm(), g.scalar("phi_comp.norm")) < TOL
    inter = od.ladder_contract(gpsi, gphi, "phys", "physout", return_intermediate_contractions=True,
                               complex_conjugate_array1=True)
    assert len(inter) == g.meta["ladder.inter.n"]
    for i, t in enumerate(inter):
        check_tensor(t, g.tensor("ladder.inter.%d" % i), TOL)
    check_tensor(od.ladder_contract(gpsi, gphi, "phys", "physout", start=1, end=3), g.tensor("ladder.mid"), TOL)
    check_tensor(od.ladder_contract(gpsi, gphi, "phys", "physout", start=2, end=len(gpsi) - 1),
                 g.tensor("ladder.right"), TOL)


def test_reference_fixture_10site(backend):
    """The reference's own test-suite (tests/test_canonical_form_mps_conversion.py)
    on its pickled 10-site MPS, plus the values derived from it in SURVEY.md section 4."""
    tn = tn_()
    od = tn.onedim
    g = Golden("fixture10")
    psi = to_chain(g, "psi")
    assert psi.bonddims() == [1, 3, 3, 3, 3, 3, 3, 3, 3, 3, 1]
    assert abs(psi.norm() - 6.8326810768514090e-01) < 1e-12
    assert abs(od.inner_product_mps(psi, psi) - 0.46685530697963323) < 1e-12
    a = psi.copy(); a.svd_compress(threshold=1e-12, normalise=False); check_chain(a, g, "comp", backend)
    np.testing.assert_almost_equal(a.norm(), psi.norm(), decimal=10)
    assert a.check_canonical_form(threshold=1e-10, print_output=False) == (0, 0)
    can = od.right_canonical_to_canonical(a, threshold=1e-12)
    sites, m = g.chain("canon")
    assert [int(b) for b in can.bonddims()] == m["bonds"]
    for i, (t, gold) in enumerate(zip(can, sites)):
        # Lambda (even) sites hold the gauge-invariant Schmidt values; Gamma sites carry SVD sign freedom
        check_tensor(t, gold, TOL, values=(i % 2 == 0))
    l, r, n = can.check_canonical_form(threshold=1e-10, print_output=False)
    assert [list(l), list(r), list(n)] == g.meta["canon.check"]
    np.testing.assert_almost_equal(can.norm(), g.scalar("canon.norm"), decimal=10)
    np.testing.assert_almost_equal(od.inner_product_mps(can, can), psi.norm() ** 2, decimal=10)
    np.testing.assert_almost_equal(od.inner_product_mps(psi, can), psi.norm() ** 2, decimal=10)
    # Schmidt values at the centre of the chain (gauge invariant)
    centre = np.diag(np.asarray(can[can.singular_site(5)].data))
    gold_centre = np.diag(g.arr["canon.%d" % (2 * 5)])
    assert np.max(np.abs(centre - gold_centre)) < 1e-10
    back = od.canonical_to_right_canonical(can)
    np.testing.assert_almost_equal(od.inner_product_mps(back, psi), psi.norm() ** 2, decimal=10)
    back = od.canonical_to_left_canonical(can)
    np.testing.assert_almost_equal(od.inner_product_mps(back, psi), psi.norm() ** 2, decimal=10)
    lc = od.left_canonical_to_canonical(od.left_canonical_form_mps(psi), threshold=1e-12)
    np.testing.assert_almost_equal(od.inner_product_mps(lc, psi), psi.norm() ** 2, decimal=10)
    b = psi.copy(); b.svd_compress(chi=2); check_chain(b, g, "comp2", backend)
    ov = od.inner_product_mps(psi, b) / (psi.norm() * b.norm())
    assert abs(ov - 0.9996930862857119) < 1e-10
    assert abs(ov - g.scalar("comp2.overlap_normalised")) < 1e-10


def test_peps_boundary_contraction(backend):
    tn = tn_()
    td = tn.twodim
    g = Golden("peps")
    L, chi = g.meta["L"], g.meta["chi"]
    grid = [[T(g.tensor("peps.%d.%d" % (r, c))) for c in range(L)] for r in range(L)]
    peps = td.SquareLatticePEPS(grid)
    net = td.inner_product_peps(peps, peps, contract_virtual=False)
    for r in range(L):
        for c in range(L):
            check_tensor(net[r, c], g.tensor("net.%d.%d" % (r, c)), tol(backend))
    assert net.can_contract()
    cols = net.mps_contract(chi, return_all_columns=True, tolerance=1e-14)
    for i, c in enumerate(cols[:-1]):
        # `mps_copy[0].data *= norm` (square_lattice.py:177) multiplies a float64 array by the long-double
        # accumulator IN PLACE: NumPy keeps float64, and so does the device array (it stays a working tensor).
        # The double-layer network has exactly degenerate singular values, so only the state as a whole is compared
        assert np.asarray(c[0].data).dtype == np.float64
        assert isinstance(c[0].data, tn.DevArray)
        check_chain(c, g, "col.%d" % i, backend, values=False)
    val = net.mps_contract(chi, tolerance=1e-14)
    assert str(np.asarray(val.data).dtype) == g.meta["result_dtype"]
    assert list(val.labels) == g.meta["result_labels"]
    assert rel_err(np.float64(val.data), g.scalar("approx")) < TOL
    full = td.inner_product_peps(peps, peps, exact_contract=False, chi=81)
    assert rel_err(np.float64(full.data), g.scalar("approx_fullchi")) < TOL
    exact = net.exact_contract()
    assert exact.labels == [] and rel_err(np.float64(np.asarray(exact.data)), g.scalar("exact")) < TOL
    check_chain(td.column_to_mpo(net, 1), g, "colmpo1", backend, values=True)


def test_utils_constructors_and_observables(backend):
    tn = tn_()
    od = tn.onedim
    np.random.seed(7)
    psi = od.init_mps_random(6, 2, 4)
    assert psi.bonddims() == [1, 4, 4, 4, 4, 4, 1]
    assert all(t.labels == ["phys", "left", "right"] for t in psi)
    z = od.init_mps_allzero(5, 3)
    assert abs(z.norm() - 1.0) < 1e-14 and z.bonddims() == [1] * 6
    lg = od.init_mps_logical(4, [1, 0, 1, 1], 2)
    assert abs(od.inner_product_mps(lg, od.init_mps_logical(4, [1, 0, 1, 1], 2)) - 1.0) < 1e-14
    assert abs(od.inner_product_mps(lg, od.init_mps_allzero(4, 2))) < 1e-14
    # <Z_i> through expvals_mps vs a dense state vector
    Z = tn.Tensor(np.diag([1.0, -1.0]), ["out", "in"])
    vals = od.expvals_mps(psi.copy(), Z)
    dense = od.contract_virtual_indices(psi, periodic_boundaries=False)
    dense.remove_all_dummy_indices()
    v = np.asarray(dense.data)
    nrm = np.vdot(v, v).real
    for i in range(6):
        zi = np.moveaxis(v, i, 0)
        want = (np.vdot(zi[0], zi[0]) - np.vdot(zi[1], zi[1])).real
        assert abs(vals[i] - want) < 1e-10 * nrm
    rhos = od.ptrace_mps(psi.copy())
    for i in (0, 3, 5):
        zi = np.moveaxis(v, i, 0).reshape(2, -1)
        assert rel_err(np.asarray(rhos[i].data), zi @ zi.conj().T) < 1e-10
    # one-body sum MPO: sum_i Z_i on |0000> gives N
    mpo = od.onebody_sum_mpo([np.diag([1.0, -1.0])] * 4)
    z4 = od.init_mps_allzero(4, 2)
    e = od.inner_product_mps(z4, od.contract_mps_mpo(z4, mpo))
    assert abs(e - 4.0) < 1e-12
    # variational compression of a compressible state reaches the SVD result's overlap
    np.random.seed(3)
    big = od.init_mps_random(6, 2, 6)
    small = od.svd_compress_mps(big, 3)
    big2 = od.contract_mps_mpo(small, od.onebody_sum_mpo([np.eye(2) * 0.5] * 6))  # bond 6, compressible to 3
    var = big2.variational_compress(3, max_iter=20, tolerance=1e-10)
    ov = abs(od.inner_product_mps(var, big2)) / (var.norm() * big2.norm())
    assert abs(ov - 1.0) < 1e-9
