"""SURVEY.md section 8(f) rows -- the callers either side of the hot path -- against fixtures written by the
unmodified reference (tests/golden/make_golden_frows.py), on both backends (see conftest.backend).  States are
gauge-dependent after an SVD, so chains are compared as dense tensors (all virtual indices contracted); bond
dimensions, labels and return types exact; numbers to 1e-10 (GPU) / 1e-12 (host stand-in)."""
import numpy as np
import pytest

from golden_io import Golden, rel_err
from test_api_parity import T, check_tensor, to_chain, tn_, tol


def dense(chain):
    od = tn_().onedim
    t = od.contract_virtual_indices(chain)
    t.remove_all_dummy_indices()
    return t


def check_dense(chain, g, key, backend, tolerance=None, phase_free=False):
    """phase_free: compare up to ONE global phase.  right_canonical_to_canonical drops the 1 x 1 unitary of its
    last SVD (onedim_core.py:1866-1869 "S and V are just scalars"), so the global phase of everything derived from
    the canonical form is whatever sign / phase convention the SVD routine has -- LAPACK's in the fixture."""
    data, labels = g.tensor(key)
    t = dense(chain)
    assert list(t.labels) == labels
    ours = np.asarray(t.data)
    if phase_free:
        ov = np.vdot(data, ours)
        assert abs(abs(ov) - np.linalg.norm(data) ** 2) <= (tolerance or tol(backend)) * np.linalg.norm(data) ** 2
        ours = ours * (np.conj(ov) / abs(ov))
    assert rel_err(ours, data) <= (tolerance or tol(backend))


def test_mps_apply_gate(backend):
    """MatrixProductState.apply_gate (onedim_core.py:662): 1-, 2- and 3-site gates, chi, canonise='right'."""
    g = Golden("frows")
    psi = to_chain(g, "psi")
    a = psi.copy(); a.apply_gate(T(g.tensor("gate2")), 2, gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"])
    assert a.bonddims() == g.meta["ag2"]["bonds"]
    assert [list(t.labels) for t in a] == [g.meta["ag2.%d" % i]["labels"] for i in range(len(a))]
    check_dense(a, g, "ag2.dense", backend)
    a = psi.copy(); a.apply_gate(T(g.tensor("gate3")), 1, gate_outputs=["o1", "o2", "o3"], gate_inputs=["i1", "i2", "i3"],
                                 chi=4, canonise="right")
    assert a.bonddims() == g.meta["ag3"]["bonds"]
    check_dense(a, g, "ag3.dense", backend)
    a = psi.copy(); a.apply_gate(T(g.tensor("gate1")), 4, gate_outputs=["out"], gate_inputs=["in"])
    assert a.bonddims() == g.meta["ag1"]["bonds"]
    check_dense(a, g, "ag1.dense", backend)


def test_mps_expval_ptrace(backend):
    """MatrixProductState.expval / ptrace (onedim_core.py:741, :825)."""
    g = Golden("frows")
    c = to_chain(g, "mixed")
    e = c.expval(T(g.tensor("gate2")), 2, left_canonised_up_to=2, right_canonised_up_to=4,
                 gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"])
    assert abs(complex(np.asarray(e.data)) - complex(g.scalar("expval2"))) <= tol(backend) * abs(g.scalar("expval2"))
    psi = to_chain(g, "psi")
    e = psi.expval(T(g.tensor("gate1")), 3, gate_outputs=["out"], gate_inputs=["in"])
    assert abs(complex(np.asarray(e.data)) - complex(g.scalar("expval1_uncanon"))) <= tol(backend) * abs(g.scalar("expval1_uncanon"))
    check_tensor(psi.ptrace(2, 3), g.tensor("ptrace23"), tol(backend))


def test_variational_compress_vs_reference(backend):
    """variational_compress (onedim_core.py:486): the alternating sweeps start from svd_compress and converge to
    the same state as the reference's (compared densely; 1e-8: twenty sweeps of local least-squares updates)."""
    g = Golden("frows")
    od = tn_().onedim
    psi = to_chain(g, "psi")
    v = psi.copy().variational_compress(3, max_iter=20, tolerance=1e-14)
    assert v.bonddims() == g.meta["varcomp"]["bonds"]
    check_dense(v, g, "varcomp.dense", backend, tolerance=1e-8)
    ov = od.inner_product_mps(psi, v)
    assert abs(ov - g.scalar("varcomp.overlap")) <= 1e-8 * abs(g.scalar("varcomp.overlap"))


def test_canonical_gates(backend):
    """MatrixProductStateCanonical: apply_gate (1 and 2 sites, chi), expval, swap_gate, compress_bond, ptrace and the
    converters (onedim_core.py:1067-1327, :1836-1928)."""
    g = Golden("frows")
    od = tn_().onedim
    r = to_chain(g, "psi"); r.right_canonise(normalise=True)
    can = od.right_canonical_to_canonical(r, threshold=1e-14)
    assert [int(b) for b in can.bonddims()] == g.meta["can0"]["bonds"]
    assert abs(can.norm() - g.scalar("can0.norm")) <= tol(backend)
    g1, g2 = T(g.tensor("gate1")), T(g.tensor("gate2"))
    can.apply_gate(g1, 2, gate_outputs=["out"], gate_inputs=["in"])
    assert [int(b) for b in can.bonddims()] == g.meta["can1.bonds"]
    e = can.expval(g1, 2, gate_outputs=["out"], gate_inputs=["in"])
    assert abs(complex(np.asarray(e.data)) - complex(g.scalar("can1.expval"))) <= tol(backend) * abs(g.scalar("can1.expval"))
    check_dense(od.canonical_to_right_canonical(can), g, "can1.dense", backend, phase_free=True)
    can.apply_gate(g2, 3, gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"], chi=4)
    assert [int(b) for b in can.bonddims()] == g.meta["can2.bonds"]
    assert [list(t.labels) for t in can] == g.meta["can2.labels"]
    check_dense(od.canonical_to_right_canonical(can), g, "can2.dense", backend, phase_free=True)
    e = can.expval(g2, 1, gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"])
    assert abs(complex(np.asarray(e.data)) - complex(g.scalar("can2.expval2"))) <= tol(backend) * abs(g.scalar("can2.expval2"))
    can.swap_gate(2)
    assert [int(b) for b in can.bonddims()] == g.meta["can3.bonds"]
    check_dense(od.canonical_to_right_canonical(can), g, "can3.dense", backend, phase_free=True)
    can.compress_bond(3, chi=2)
    assert [int(b) for b in can.bonddims()] == g.meta["can4.bonds"]
    check_dense(od.canonical_to_right_canonical(can), g, "can4.dense", backend, phase_free=True)
    check_tensor(can.ptrace(1, 2), g.tensor("can4.ptrace"), tol(backend))
    lcan = od.left_canonical_to_canonical(od.left_canonical_form_mps(to_chain(g, "psi"), normalise=True))
    assert [int(b) for b in lcan.bonddims()] == g.meta["lcan.bonds"]
    check_dense(od.canonical_to_left_canonical(lcan), g, "lcan.dense", backend, phase_free=True)
    # the generator's psi had been through expval() and ptrace() by then, which canonise IN PLACE (:741, :825):
    # replaying them checks that side effect too
    p2 = to_chain(g, "psi")
    p2.expval(g1, 3, gate_outputs=["out"], gate_inputs=["in"])
    p2.ptrace(2, 3)
    s = p2.copy(); s.swap_gate(3)
    assert [int(b) for b in s.bonddims()] == g.meta["swap.bonds"]
    check_dense(s, g, "swap.dense", backend)


def test_tensor_to_mps_mpo(backend):
    """tensor_to_mps / tensor_to_mpo (onedim_core.py:1711, :1764): exact and chi-truncated splitting."""
    g = Golden("frows")
    od = tn_().onedim
    big = T(g.tensor("big"))
    m = od.tensor_to_mps(big, phys_labels=["a", "b", "c", "dd", "e"])
    assert m.bonddims() == g.meta["t2mps"]["bonds"]
    assert [list(t.labels) for t in m] == [g.meta["t2mps.%d" % i]["labels"] for i in range(len(m))]
    check_dense(m, g, "t2mps.dense", backend)
    m = od.tensor_to_mps(big, phys_labels=["c", "a", "e", "b", "dd"], chi=3)
    assert m.bonddims() == g.meta["t2mps3"]["bonds"]
    check_dense(m, g, "t2mps3.dense", backend)
    w = od.tensor_to_mpo(T(g.tensor("op")), physout_labels=["o0", "o1", "o2"], physin_labels=["i0", "i1", "i2"])
    assert w.bonddims() == g.meta["t2mpo"]["bonds"]
    assert [list(t.labels) for t in w] == [g.meta["t2mpo.%d" % i]["labels"] for i in range(len(w))]
    wt = od.contract_virtual_indices(w); wt.remove_all_dummy_indices()
    check_tensor(wt, g.tensor("t2mpo.dense"), tol(backend))


def test_contract_multi_index_tensor_with_one_dim_array(backend):
    """onedim_core.py:1370: one tensor with N equally labelled indices against a chain of N tensors."""
    g = Golden("frows")
    od = tn_().onedim
    out = od.contract_multi_index_tensor_with_one_dim_array(T(g.tensor("multi")), to_chain(g, "arr3"), "x", "phys")
    check_tensor(out, g.tensor("multi.out"), tol(backend))


def _peps(g, L=3):
    tn = tn_()
    return tn.twodim.SquareLatticePEPS([[T(g.tensor("peps.%d.%d" % (r, c))) for c in range(L)] for r in range(L)])


def test_twodim_variational_and_pepo_helpers(backend):
    """mps_contract(compression_type='variational') (square_lattice.py:164-173), SquareLatticePEPS.outer_product
    (:251), SquareLatticePEPO.trace (:362), apply_pepo_to_peps (:377)."""
    g = Golden("frows")
    td = tn_().twodim
    peps = _peps(g)
    net = td.inner_product_peps(peps, peps, contract_virtual=False)
    ex = float(np.asarray(net.exact_contract().data))
    assert abs(ex - float(g.scalar("peps.exact"))) <= tol(backend) * abs(float(g.scalar("peps.exact")))
    for chi, key in ((3, "peps.var3"), (4, "peps.var4")):
        val = net.mps_contract(chi, compression_type="variational", max_iter=10, tolerance=1e-14)
        assert abs(float(np.asarray(val.data)) - float(g.scalar(key))) <= 1e-8 * abs(float(g.scalar(key)))
    pepo = peps.outer_product()
    L = 3
    for r in range(L):
        for c in range(L):
            assert list(pepo[r, c].labels) == g.meta["pepo.labels"][r][c]
            check_tensor(pepo[r, c], g.tensor("pepo.%d.%d" % (r, c)), tol(backend))
    tr = float(np.asarray(pepo.trace().exact_contract().data))
    assert abs(tr - float(g.scalar("pepo.trace"))) <= tol(backend) * abs(float(g.scalar("pepo.trace")))
    applied = td.apply_pepo_to_peps(peps, pepo)
    for r in range(L):
        for c in range(L):
            check_tensor(applied[r, c], g.tensor("applied.%d.%d" % (r, c)), tol(backend))
    n2 = float(np.asarray(td.inner_product_peps(applied, applied).data))
    assert abs(n2 - float(g.scalar("applied.norm2"))) <= tol(backend) * abs(float(g.scalar("applied.norm2")))


def test_mpo_apply_beyond_the_fused_kernel_limit(backend):
    """contract_mps_mpo with an MPO slice larger than the fused kernel's 48 KB of shared memory (wr * d * 16 B):
    the generic contract + consolidate_indices path takes over (it used to raise)."""
    tn = tn_()
    od = tn.onedim
    rng = np.random.default_rng(11)
    d, wr = 56, 64                                           # 56 * 64 * 16 B = 56 KB
    a = [rng.standard_normal((d, 1, 3)) + 1j * rng.standard_normal((d, 1, 3)),
         rng.standard_normal((d, 3, 1)) + 1j * rng.standard_normal((d, 3, 1))]
    w = [rng.standard_normal((1, wr, 2, d)) + 0j, rng.standard_normal((wr, 1, 2, d)) + 0j]
    psi = od.MatrixProductState([tn.Tensor(x, ["phys", "left", "right"]) for x in a])
    H = od.MatrixProductOperator([tn.Tensor(x, ["left", "right", "physout", "physin"]) for x in w], "left", "right",
                                 "physout", "physin")
    phi = od.contract_mps_mpo(psi, H)
    assert phi.bonddims() == [1, 3 * wr, 1]
    ref0 = np.einsum("qlr,abpq->lapr b".replace(" ", ""), a[0], w[0]).reshape(1, 2, 3 * wr)
    assert list(phi[0].labels) == ["left", "physout", "right"]
    assert rel_err(np.asarray(phi[0].data), ref0) <= tol(backend)


def test_inv_pad_and_high_rank(backend):
    """Tensor.inv() of a general matrix (tensor.py:487-488 inverts anything), pad_index through the
    C ABI (tensor.py:543-563), tensors with more than 12 indices (exact_contract of tall lattices)."""
    tn = tn_()
    rng = np.random.default_rng(21)
    a = rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))
    t = tn.Tensor(a, ["r", "c"]); t.inv()
    assert t.labels == ["r", "c"]
    assert rel_err(np.asarray(t.data), np.linalg.inv(a)) <= 1e-10
    d = tn.Tensor(np.diag([2.0, 4.0, 0.5]), ["r", "c"]); d.inv()
    assert np.array_equal(np.asarray(d.data), np.diag([0.5, 0.25, 2.0]))
    with pytest.raises(np.linalg.LinAlgError):
        z = tn.Tensor(np.zeros((3, 3)), ["r", "c"]); z.inv()
    b = rng.standard_normal((2, 3, 4))
    for before in (False, True):
        p = tn.Tensor(b, ["x", "y", "z"]); p.pad_index("y", 2, before=before)
        ref = np.pad(b, [(0, 0), (2, 0) if before else (0, 2), (0, 0)])
        assert p.labels == ["x", "y", "z"] and np.array_equal(np.asarray(p.data), ref)
        q = p.copy(); q.move_index("y", 0)                       # the padded tensor is an ordinary working tensor
        assert np.array_equal(np.asarray(q.data), np.moveaxis(ref, 1, 0))
    big = rng.standard_normal((2,) * 14)
    m = rng.standard_normal((2, 2))
    out = tn.contract(tn.Tensor(big, ["i%d" % k for k in range(14)]), tn.Tensor(m, ["a", "b"]), "i5", "a")
    assert out.labels == ["i%d" % k for k in range(14) if k != 5] + ["b"]
    assert rel_err(np.asarray(out.data), np.tensordot(big, m, ([5], [0]))) <= tol(backend)
