"""GPU: the batched path (BASELINE.json config 4) -- tnb_svd_project_batched and the BatchedMPS sweeps built on it --
against NumPy, against the per-network path and against values the unmodified reference produced for networks
0..3 of seed 3 at full size (tests/golden/fullsize_cfg4.npz).  Tolerance 1e-10 relative; bond dimensions exact."""
import ctypes

import numpy as np
import pytest

from golden_io import Golden

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _svd_batched(a):
    import torch
    from tncontract_b200 import batched, devarray as dv
    t = dv.DevArray.from_host(a).t
    u, s, p, sweeps = batched._svd_project_batched(t)
    torch.cuda.synchronize()
    return u.cpu().numpy(), s.cpu().numpy(), p.cpu().numpy(), sweeps


@pytest.mark.parametrize("shape", [(5, 64, 24), (3, 24, 64), (4, 256, 128), (2, 16, 128), (6, 40, 40), (1, 7, 3)])
@pytest.mark.parametrize("cplx", [False, True])
def test_svd_project_batched_vs_numpy(shape, cplx):
    rng = np.random.default_rng(hash((shape, cplx)) & 0xffff)
    a = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
    a[0] *= 1e3                                        # matrices of a batch are scaled independently
    if shape[0] > 1:
        a[1] *= 1e-5
    u, s, p, sweeps = _svd_batched(a)
    bsz, m, n = shape
    k = min(m, n)
    assert sweeps > 0
    for b in range(bsz):
        sref = np.linalg.svd(a[b], compute_uv=False)
        assert np.max(np.abs(s[b] - sref)) <= 1e-12 * sref[0]
        assert np.linalg.norm(u[b].conj().T @ u[b] - np.eye(k)) <= 1e-12 * k          # isometry
        assert np.linalg.norm(u[b].conj().T @ a[b] - p[b]) <= 1e-12 * sref[0] * k      # P = U^H A
        assert np.linalg.norm(u[b] @ p[b] - a[b]) <= 1e-12 * sref[0] * k               # A = U (S Vh)
        vh = p[b] / s[b][:, None]
        assert np.linalg.norm(vh @ vh.conj().T - np.eye(k)) <= 1e-11 * k


def test_svd_project_batched_flags_per_matrix():
    """A NaN in one matrix of the batch is reported (LinAlgError, as numpy.linalg.svd); a zero matrix is fine."""
    rng = np.random.default_rng(5)
    a = rng.standard_normal((3, 32, 16))
    a[1] = 0.0
    u, s, p, _ = _svd_batched(a)
    assert np.all(s[1] == 0.0) and np.all(np.isfinite(u)) and np.all(np.isfinite(p))
    a[2, 3, 3] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        _svd_batched(a)


def test_batched_cfg4_vs_reference_golden():
    """Networks 0..3 of cfg 4 (N=64, d=4, chi=128 -> 64, float64) as ONE batch: overlap, norm, every bond's
    singular values, bonds, norm after compression -- against the reference's values."""
    from tncontract_b200 import batched
    g = Golden("fullsize_cfg4")
    nets = g.meta["networks"]
    recs, a = batched.overlap_norm_compress_batched(g.meta["seed"], nets, 64, 4, 128, 64)
    for j, net in enumerate(nets):
        r = recs[j]
        assert r[0] == net
        assert abs(r[1] - g.scalar("n%d.overlap" % net).real) <= TOL * abs(g.scalar("n%d.overlap" % net))
        assert abs(r[2] - g.scalar("n%d.norm" % net).real) <= TOL * abs(g.scalar("n%d.norm" % net))
        assert abs(r[3] - g.scalar("n%d.norm_after_right" % net).real) <= TOL * abs(g.scalar("n%d.norm_after_right" % net))
        assert [int(x) for x in r[4:]] == g.meta["n%d.bonds" % net]
        assert len(a.singular_values) == g.meta["n%d.comp.nsvd" % net]
        for i, s in enumerate(a.singular_values):
            ref = g.scalar("n%d.comp.s%d" % (net, i))
            assert s[j].shape == ref.shape
            assert np.max(np.abs(s[j] - ref)) <= TOL * ref[0], (net, i)


def test_batched_complex_vs_per_network_path():
    """Complex128 batch of 5 small chains: overlap, norm and compression agree with the per-network public API
    (compressed states compared as vectors through overlaps)."""
    import tncontract_b200 as tn
    from tncontract_b200 import batched
    od = tn.onedim
    rng = np.random.default_rng(9)
    n, d, chi, keep, bsz = 9, 3, 12, 7, 5
    bonds = [1] + [chi] * (n - 1) + [1]

    def chain():
        return od.MatrixProductState([tn.Tensor(rng.standard_normal((d, bonds[i], bonds[i + 1])) +
                                                1j * rng.standard_normal((d, bonds[i], bonds[i + 1])),
                                                ["phys", "left", "right"]) for i in range(n)])
    xs, ys = [chain() for _ in range(bsz)], [chain() for _ in range(bsz)]
    bx, by = batched.BatchedMPS.from_mps_list(xs), batched.BatchedMPS.from_mps_list(ys)
    ov = bx.inner_product(by)
    nr = bx.norm()
    bx.svd_compress(chi=keep)
    outs = bx.to_mps_list()
    for j in range(bsz):
        ref_ov = od.inner_product_mps(xs[j], ys[j])
        assert abs(ov[j] - ref_ov) <= TOL * abs(ref_ov)
        assert abs(nr[j] - xs[j].norm()) <= TOL * abs(xs[j].norm())
        c = xs[j].copy(); c.svd_compress(chi=keep)
        assert outs[j].bonddims() == c.bonddims()
        cc = od.inner_product_mps(c, c)
        assert abs(od.inner_product_mps(c, outs[j]) - cc) <= TOL * abs(cc)
        assert abs(od.inner_product_mps(outs[j], outs[j]) - cc) <= TOL * abs(cc)


def test_ragged_batch_is_reported():
    """Networks that keep different bond dimensions cannot share a batch: RaggedBatchError (the caller falls back
    to the per-network path)."""
    import tncontract_b200 as tn
    from tncontract_b200 import batched
    od = tn.onedim
    rng = np.random.default_rng(3)
    n, d, chi = 5, 2, 4
    bonds = [1] + [chi] * (n - 1) + [1]
    full = od.MatrixProductState([tn.Tensor(rng.standard_normal((d, bonds[i], bonds[i + 1])), ["phys", "left", "right"])
                                  for i in range(n)])
    prod = od.MatrixProductState([tn.Tensor(np.ones((d, bonds[i], bonds[i + 1])), ["phys", "left", "right"])
                                  for i in range(n)])                      # a product state: every bond has rank 1
    b = batched.BatchedMPS.from_mps_list([full, prod])
    with pytest.raises(batched.RaggedBatchError):
        b.svd_compress(threshold=1e-10)
