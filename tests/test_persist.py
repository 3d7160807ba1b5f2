"""Row f4 of SURVEY.md section 8: pickle interchange with the reference.  ``tests/golden/ref_pickle_10site.dat`` is
the reference's own 10-site fixture (tncontract/tests/random_10site_mps.dat) pickled again by the reference's classes
(make_golden.py: gen_fixture)."""
import io
import os
import pickle
import sys

import numpy as np
import pytest

from golden_io import GOLDEN, Golden
from test_api_parity import check_tensor, tn_


def _load_fixture():
    from tncontract_b200 import persist
    with open(os.path.join(GOLDEN, "ref_pickle_10site.dat"), "rb") as f:
        return persist.load(f)


def test_reference_pickle_loads_into_our_classes(backend):
    tn = tn_()
    g = Golden("fixture10")
    psi = _load_fixture()
    assert type(psi) is tn.onedim.MatrixProductState
    assert all(type(t) is tn.Tensor for t in psi)
    sites, m = g.chain("psi")
    assert (psi.left_label, psi.right_label, psi.phys_label) == (m["left"], m["right"], m["phys_label"])
    assert [int(b) for b in psi.bonddims()] == m["bonds"]
    for t, gold in zip(psi, sites):
        check_tensor(t, gold, 0.0)                       # data moved, not recomputed: bit-exact
    assert abs(psi.norm() - g.scalar("norm")) <= 1e-10 * abs(g.scalar("norm"))
    psi.svd_compress(chi=2)                              # and it is a working network
    assert max(psi.bonddims()) == 2


def test_dump_writes_reference_class_paths_and_round_trips(backend):
    from tncontract_b200 import persist
    tn = tn_()
    psi = _load_fixture()
    raw = persist.dumps(psi)
    assert b"ctncontract.onedim.onedim_core\nMatrixProductState\n" in raw
    assert b"ctncontract.tensor\nTensor\n" in raw
    assert b"tncontract_b200" not in raw
    back = persist.loads(raw)
    assert type(back) is tn.onedim.MatrixProductState
    for a, b in zip(psi, back):
        assert a.labels == b.labels and np.array_equal(np.asarray(a.data), np.asarray(b.data))
    own = persist.dumps(psi, reference_paths=False)
    assert b"tncontract_b200.onedim.onedim_core" in own
    assert type(pickle.loads(own)) is tn.onedim.MatrixProductState      # plain pickle works for our own paths
    t = tn.Tensor(np.arange(6.0).reshape(2, 3) * (1 + 2j), ["a", "b"])
    t2 = persist.loads(persist.dumps(t))
    assert t2.labels == ["a", "b"] and np.array_equal(np.asarray(t2.data), np.asarray(t.data))


@pytest.mark.skipif(not os.path.isdir("/root/reference/tncontract"), reason="reference only in the build container")
def test_reference_reads_our_dump(backend):
    """The unmodified reference unpickles what dump() wrote (run in a subprocess: its import needs the NumPy-2 shim)."""
    import subprocess
    from tncontract_b200 import persist
    psi = _load_fixture()
    psi.left_canonise(qr_decomposition=True)
    raw = persist.dumps(psi)
    code = ("import sys, pickle, numpy as np; np.product = np.prod; np.float = float; sys.dont_write_bytecode = True;"
            "sys.path.insert(0, '/root/reference'); import tncontract;"
            "psi = pickle.load(sys.stdin.buffer);"
            "print(type(psi).__module__, type(psi).__name__, repr(float(psi.norm())), psi.bonddims())")
    out = subprocess.run([sys.executable, "-c", code], input=raw, capture_output=True, check=True).stdout.decode().split()
    assert out[0] == "tncontract.onedim.onedim_core" and out[1] == "MatrixProductState"
    assert abs(float(out[2]) - float(psi.norm())) <= 1e-10 * float(psi.norm())
