#!/usr/bin/env python
"""Golden fixtures for the rows of SURVEY.md section 8(f) (the callers either side of the hot path), written by
the UNMODIFIED reference -- same harness rules as make_golden.py.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_frows.py     ->  tests/golden/frows.npz

Covers: MatrixProductState.apply_gate / expval / ptrace / variational_compress, MatrixProductStateCanonical
apply_gate / swap_gate / compress_bond / expval / ptrace, tensor_to_mps / tensor_to_mpo,
contract_multi_index_tensor_with_one_dim_array, mps_contract(compression_type="variational"), PEPO helpers
(outer_product, apply_pepo_to_peps, SquareLatticePEPO.trace).
"""
import os
import sys

import numpy as np

np.product = np.prod
np.float = float
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import tncontract as tn  # noqa: E402
import tncontract.onedim as od  # noqa: E402
import tncontract.twodim as td  # noqa: E402

from make_golden import Bag  # noqa: E402


def dense(chain, g, key):
    """the state as one tensor (gauge-free comparison): contract all virtual indices"""
    t = od.contract_virtual_indices(chain)
    t.remove_all_dummy_indices()
    g.tensor(key, t)


def main():
    g = Bag("frows")
    rng = np.random.default_rng(7)

    def rt(shape, labels, cplx=True):
        a = rng.standard_normal(shape)
        if cplx:
            a = a + 1j * rng.standard_normal(shape)
        return tn.Tensor(a, labels)

    N, d, chi = 7, 2, 5
    bonds = [1] + [chi] * (N - 1) + [1]
    psi = od.MatrixProductState([rt((d, bonds[i], bonds[i + 1]), ["phys", "left", "right"]) for i in range(N)])
    g.chain("psi", psi)
    gate2 = rt((d, d, d, d), ["o1", "o2", "i1", "i2"])
    gate1 = rt((d, d), ["out", "in"])
    gate3 = rt((d, d, d, d, d, d), ["o1", "o2", "o3", "i1", "i2", "i3"])
    g.tensor("gate1", gate1); g.tensor("gate2", gate2); g.tensor("gate3", gate3)

    # ---- MatrixProductState.apply_gate (onedim_core.py:662) ---------------------------------------------
    a = psi.copy(); a.apply_gate(gate2, 2, gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"])
    g.chain("ag2", a); dense(a, g, "ag2.dense")
    a = psi.copy(); a.apply_gate(gate3, 1, gate_outputs=["o1", "o2", "o3"], gate_inputs=["i1", "i2", "i3"], chi=4,
                                 canonise="right")
    g.chain("ag3", a); dense(a, g, "ag3.dense")
    a = psi.copy(); a.apply_gate(gate1, 4, gate_outputs=["out"], gate_inputs=["in"])
    g.chain("ag1", a); dense(a, g, "ag1.dense")

    # ---- expval / ptrace (onedim_core.py:741, :825) ------------------------------------------------------
    c = psi.copy(); c.left_canonise(0, 2); c.right_canonise(4, N)
    g.chain("mixed", c)
    g.scalar("expval2", np.asarray(c.expval(gate2, 2, left_canonised_up_to=2, right_canonised_up_to=4,
                                            gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"]).data))
    g.scalar("expval1_uncanon", np.asarray(psi.expval(gate1, 3, gate_outputs=["out"], gate_inputs=["in"]).data))
    rho = psi.ptrace(2, 3)
    g.tensor("ptrace23", rho)

    # ---- variational_compress (onedim_core.py:486) -------------------------------------------------------
    v = psi.copy()
    v = v.variational_compress(3, max_iter=20, tolerance=1e-14)
    g.chain("varcomp", v); dense(v, g, "varcomp.dense")
    g.scalar("varcomp.overlap", od.inner_product_mps(psi, v))

    # ---- MatrixProductStateCanonical (onedim_core.py:880-1327) ------------------------------------------
    r = psi.copy(); r.right_canonise(normalise=True)
    can = od.right_canonical_to_canonical(r, threshold=1e-14)
    g.chain("can0", can)
    g.scalar("can0.norm", can.norm())
    can.apply_gate(gate1, 2, gate_outputs=["out"], gate_inputs=["in"])
    g.meta["can1.bonds"] = [int(b) for b in can.bonddims()]
    g.scalar("can1.expval", np.asarray(can.expval(gate1, 2, gate_outputs=["out"], gate_inputs=["in"]).data))
    dense(od.canonical_to_right_canonical(can), g, "can1.dense")
    can.apply_gate(gate2, 3, gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"], chi=4)
    g.meta["can2.bonds"] = [int(b) for b in can.bonddims()]
    g.meta["can2.labels"] = [[str(l) for l in t.labels] for t in can]
    dense(od.canonical_to_right_canonical(can), g, "can2.dense")
    g.scalar("can2.expval2", np.asarray(can.expval(gate2, 1, gate_outputs=["o1", "o2"], gate_inputs=["i1", "i2"]).data))
    can.swap_gate(2)
    g.meta["can3.bonds"] = [int(b) for b in can.bonddims()]
    dense(od.canonical_to_right_canonical(can), g, "can3.dense")
    can.compress_bond(3, chi=2)
    g.meta["can4.bonds"] = [int(b) for b in can.bonddims()]
    dense(od.canonical_to_right_canonical(can), g, "can4.dense")
    g.tensor("can4.ptrace", can.ptrace(1, 2))
    lcan = od.left_canonical_to_canonical(od.left_canonical_form_mps(psi, normalise=True))
    g.meta["lcan.bonds"] = [int(b) for b in lcan.bonddims()]
    dense(od.canonical_to_left_canonical(lcan), g, "lcan.dense")

    # OneDimensionalTensorNetwork.swap_gate on a plain MPS (onedim_core.py:95)
    s = psi.copy(); s.swap_gate(3)
    g.meta["swap.bonds"] = [int(b) for b in s.bonddims()]
    dense(s, g, "swap.dense")

    # ---- tensor_to_mps / tensor_to_mpo (onedim_core.py:1711, :1764) -------------------------------------
    big = rt((2, 3, 2, 3, 2), ["a", "b", "c", "dd", "e"])
    g.tensor("big", big)
    m = od.tensor_to_mps(big, phys_labels=["a", "b", "c", "dd", "e"])
    g.chain("t2mps", m); dense(m, g, "t2mps.dense")
    m = od.tensor_to_mps(big, phys_labels=["c", "a", "e", "b", "dd"], chi=3)
    g.chain("t2mps3", m); dense(m, g, "t2mps3.dense")
    op = rt((2, 2, 2, 2, 2, 2), ["o0", "o1", "o2", "i0", "i1", "i2"])
    g.tensor("op", op)
    w = od.tensor_to_mpo(op, physout_labels=["o0", "o1", "o2"], physin_labels=["i0", "i1", "i2"])
    g.chain("t2mpo", w)
    wt = od.contract_virtual_indices(w); wt.remove_all_dummy_indices(); g.tensor("t2mpo.dense", wt)

    # ---- contract_multi_index_tensor_with_one_dim_array (onedim_core.py:1370) ---------------------------
    multi = rt((2, 2, 2, 3), ["x", "x", "x", "free"])
    g.tensor("multi", multi)
    arr = od.MatrixProductState([rt((d, bonds[i], bonds[i + 1]), ["phys", "left", "right"]) for i in range(3)])
    g.chain("arr3", arr)
    out = od.contract_multi_index_tensor_with_one_dim_array(multi, arr, "x", "phys")
    g.tensor("multi.out", out)

    # ---- twodim: variational boundary contraction and PEPO helpers ----------------------------------------
    L, D = 3, 2
    grid = []
    for rr in range(L):
        row = []
        for cc in range(L):
            shape = (d, 1 if rr == 0 else D, 1 if rr == L - 1 else D, 1 if cc == 0 else D, 1 if cc == L - 1 else D)
            row.append(tn.Tensor(rng.standard_normal(shape), ["phys", "up", "down", "left", "right"]))
        grid.append(row)
    peps = td.SquareLatticePEPS(grid)
    for rr in range(L):
        for cc in range(L):
            g.tensor("peps.%d.%d" % (rr, cc), peps[rr, cc])
    net = td.inner_product_peps(peps, peps, contract_virtual=False)
    g.scalar("peps.exact", np.asarray(net.exact_contract().data, dtype=np.float64))
    val = net.mps_contract(3, compression_type="variational", max_iter=10, tolerance=1e-14)
    g.scalar("peps.var3", np.asarray(val.data, dtype=np.float64))
    val = net.mps_contract(4, compression_type="variational", max_iter=10, tolerance=1e-14)
    g.scalar("peps.var4", np.asarray(val.data, dtype=np.float64))
    pepo = peps.outer_product()
    g.meta["pepo.labels"] = [[[str(l) for l in pepo[rr, cc].labels] for cc in range(L)] for rr in range(L)]
    for rr in range(L):
        for cc in range(L):
            g.tensor("pepo.%d.%d" % (rr, cc), pepo[rr, cc])
    tr = pepo.trace()
    g.scalar("pepo.trace", np.asarray(tr.exact_contract().data, dtype=np.float64))
    applied = td.apply_pepo_to_peps(peps, pepo)
    for rr in range(L):
        for cc in range(L):
            g.tensor("applied.%d.%d" % (rr, cc), applied[rr, cc])
    g.scalar("applied.norm2", np.asarray(td.inner_product_peps(applied, applied).data, dtype=np.float64))
    g.save()


if __name__ == "__main__":
    main()
