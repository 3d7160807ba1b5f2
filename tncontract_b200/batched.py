"""Batched path: B independent matrix product states of ONE shape, processed site by site with one set of
kernel launches for the whole batch (BASELINE.json config 4; north_star item 4).

The reference handles one network per Python loop iteration (``svd_compress`` onedim_core.py:463-484,
``left_canonise`` :218-358, ``inner_product_mps`` :1666-1683, ``norm`` :642-660); a chi ~ 128 network can
neither fill a B200 nor amortise its ~17 000 kernel launches.  Here the same site of every network of a
shard is one strided-batched GEMM / one batched projection SVD (``tnb_svd_project_batched``: grid row =
network), so the launch count per network drops by the batch size and the SMs see ``B x (pairs per round)``
CTAs per Jacobi round.

Layout: site ``i`` of the batch is ONE contiguous device array of shape ``(B, Dl, d, Dr)`` (left, phys, right),
so both matricisations a sweep needs are plain views -- ``(Dl d) x Dr`` going right, ``Dl x (d Dr)`` going left.

Arithmetic follows the reference's sweeps step by step (norm carried out of the left sweep, singular values
normalised by the largest one for the threshold test, ``[:chi]`` after the threshold -- onedim_core.py:333-339);
the orthogonalising left sweep uses the projection SVD in place of the QR (same isometries up to the gauge
the following SVD sweep fixes anyway; no truncation happens there).  All networks of a batch must keep the same
number of singular values at every bond (true for the full-rank random networks of config 4); a ragged batch
raises ``RaggedBatchError`` and the caller falls back to the per-network path.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import devarray as dv

__all__ = ["BatchedMPS", "RaggedBatchError", "overlap_norm_compress_batched"]


class RaggedBatchError(RuntimeError):
    """The networks of a batch kept different bond dimensions (per-network results diverge in shape)."""


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _bgemm(opa, opb, m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch, alpha=1.0, beta=0.0):
    """C[i] = alpha op(A[i]) op(B[i]) + beta C[i], i < batch (tnb_gemm, strided-batched, row-major)."""
    if m == 0 or n == 0 or batch == 0:
        return
    al = (ctypes.c_double * 2)(float(np.real(alpha)), float(np.imag(alpha)))
    be = (ctypes.c_double * 2)(float(np.real(beta)), float(np.imag(beta)))
    _lib.check(_lib.load().tnb_gemm(_lib.dtype_code(dv._T2NP[c.dtype]), opa, opb, m, n, k, al, _ptr(a), lda, sa,
                                    _ptr(b), ldb, sb, be, _ptr(c), ldc, sc, batch, dv.stream_ptr()))


def _svd_project_batched(a, want_p=True):
    """a: (B, m, n) contiguous torch tensor -> u (B, m, k), s (B, k) float64, p (B, k, n) = u^H a."""
    bsz, m, n = a.shape
    k = min(m, n)
    npd = dv._T2NP[a.dtype]
    u = dv._empty((bsz, m, k), npd)
    s = dv._empty((bsz, k), np.float64)
    p = dv._empty((bsz, k, n), npd) if want_p else None
    lib = _lib.load()
    code = _lib.dtype_code(npd)
    need = lib.tnb_svd_project_batched_workspace(code, m, n, bsz)
    ws_t, ws = dv.workspace(need)
    sweeps = ctypes.c_int32(0)
    _lib.check(lib.tnb_svd_project_batched(code, m, n, bsz, _ptr(a), n, m * n, _ptr(u), m * k, _ptr(s), k,
                                           _ptr(p) if want_p else None, k * n, ws, need, ctypes.byref(sweeps),
                                           dv.stream_ptr()))
    return u, s, p, sweeps.value


def _contig(t):
    """contiguous copy of a strided torch view through tnb_permute"""
    return t if t.is_contiguous() else dv.DevArray(t).contiguous().t


def _scale_rows(x, c, mode=0):
    """x[b] *= f(c[b]) for a contiguous (B, ...) tensor and a contiguous float64 (B,) vector (tnb_diag_scale)."""
    bsz = x.shape[0]
    per = x.numel() // bsz if bsz else 0
    if bsz and per:
        _lib.check(_lib.load().tnb_diag_scale(_lib.dtype_code(dv._T2NP[x.dtype]), _ptr(x), bsz, per, per, _ptr(c), 0, mode,
                                              dv.stream_ptr()))


class BatchedMPS:
    """``B`` matrix product states with identical shapes; ``sites[i]``: torch CUDA tensor ``(B, Dl, d, Dr)``."""

    def __init__(self, sites):
        self.sites = list(sites)
        self.singular_values = []      # filled by svd_compress: one (B, kept...) host array per bond, right to left

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def random_uniform(cls, seed, networks, which, nsites, physdim, bonddim):
        """cfg 4 input for the given network indices: U[0,1) entries from the counter-based generator of
        tncontract_b200.batch (keyed by seed, network, mps, site; element order (phys, left, right) as
        batch.random_mps), generated on the device with one launch per site."""
        from .batch import network_key
        bonds = [1] + [bonddim] * (nsites - 1) + [1]
        bsz = len(networks)
        lib = _lib.load()
        sites = []
        for i in range(nsites):
            keys = np.array([network_key(seed, net, which, i) for net in networks], dtype=np.uint64)
            kd = torch.from_numpy(keys.view(np.int64)).to(dv.device())
            raw = dv._empty((bsz, physdim, bonds[i], bonds[i + 1]), np.float64)
            _lib.check(lib.tnb_fill_uniform_batched(_ptr(raw), physdim * bonds[i] * bonds[i + 1], bsz, _ptr(kd),
                                                    dv.stream_ptr()))
            sites.append(_contig(raw.permute(0, 2, 1, 3)))          # (B, p, l, r) -> (B, l, p, r)
        return cls(sites)

    @classmethod
    def from_mps_list(cls, mps_list):
        """Stack MatrixProductState objects of identical shapes (device or host data)."""
        n = len(mps_list[0])
        sites = []
        for i in range(n):
            per = []
            for psi in mps_list:
                t = psi[i].copy()
                t.move_indices([psi.left_label, psi.phys_label, psi.right_label], 0)
                per.append(np.asarray(t.data))
            sites.append(dv.DevArray.from_host(np.stack(per)).t)
        return cls(sites)

    def to_mps_list(self):
        """-> list of MatrixProductState (labels left / phys / right), one per network."""
        from . import tensor as tsr
        from .onedim import onedim_core as core
        out = []
        for b in range(self.batch):
            ts = [tsr.Tensor(dv.DevArray(s[b]), ["left", "phys", "right"]) for s in self.sites]
            out.append(core.MatrixProductState(ts, "left", "right", "phys"))
        return out

    def copy(self):
        return BatchedMPS([dv.DevArray(s).copy().t for s in self.sites])

    # ---- metadata ---------------------------------------------------------------------------------------
    @property
    def batch(self):
        return self.sites[0].shape[0]

    def __len__(self):
        return len(self.sites)

    def bonddims(self):
        return [self.sites[0].shape[1]] + [s.shape[3] for s in self.sites]

    # ---- <a|b>: ladder contraction (onedim_core.py:1491-1663 / :1666-1683), two batched GEMMs per site ------
    def inner_product(self, other):
        """<self|other> for every network -> host array (B,)."""
        bsz = self.batch
        npd = dv._T2NP[self.sites[0].dtype]
        env = dv.DevArray.from_host(np.ones((bsz, 1, 1), dtype=npd)).t
        for a, b in zip(self.sites, other.sites):
            la, d, ra = a.shape[1:]
            lb, _, rb = b.shape[1:]
            t = dv._empty((bsz, la, d * rb), npd)
            # T[la, (p rb)] = sum_lb E[la, lb] b[lb, (p rb)]
            _bgemm(_lib.OP_N, _lib.OP_N, la, d * rb, lb, env, lb, la * lb, b, d * rb, lb * d * rb, t, d * rb, la * d * rb, bsz)
            env = dv._empty((bsz, ra, rb), npd)
            # E'[ra, rb] = sum_(la p) conj(a[(la p), ra]) T[(la p), rb]
            _bgemm(_lib.OP_C, _lib.OP_N, ra, rb, la * d, a, ra, la * d * ra, t, rb, la * d * rb, env, rb, ra * rb, bsz)
        return np.asarray(dv.DevArray(env)).reshape(bsz)

    def norm(self):
        """sqrt(<psi|psi>) per network (MatrixProductState.norm(), onedim_core.py:660)."""
        return np.sqrt(np.real(self.inner_product(self)))

    def _frob(self, site):
        x = self.sites[site]
        bsz = x.shape[0]
        per = x.numel() // bsz
        out = dv._empty((bsz,), np.float64)
        _lib.check(_lib.load().tnb_norm2_batched(_lib.dtype_code(dv._T2NP[x.dtype]), _ptr(x), per, per, bsz, _ptr(out),
                                                 dv.stream_ptr()))
        return out

    def norm_canonical(self, form):
        """norm(canonical_form="left"/"right"): Frobenius norm of the last / first tensor (onedim_core.py:656-658)."""
        return np.asarray(dv.DevArray(self._frob(len(self) - 1 if form == "left" else 0)))

    # ---- sweeps -------------------------------------------------------------------------------------------
    def left_canonise(self):
        """left_canonise(qr_decomposition=True, normalise=False) for every network (onedim_core.py:267-294):
        site i <- isometry, the remainder is absorbed into site i + 1.  The isometry comes from the batched
        projection SVD (U, P = U^H A) instead of a QR: the same subspace in a gauge the later SVD sweep fixes."""
        bsz = self.batch
        for i in range(len(self) - 1):
            a = self.sites[i]
            dl, d, dr = a.shape[1:]
            u, s, p, _ = _svd_project_batched(a.reshape(bsz, dl * d, dr))
            k = u.shape[2]
            self.sites[i] = u.reshape(bsz, dl, d, k)
            nxt = self.sites[i + 1]
            _, d2, dr2 = nxt.shape[1:]
            out = dv._empty((bsz, k, d2, dr2), dv._T2NP[a.dtype])
            _bgemm(_lib.OP_N, _lib.OP_N, k, d2 * dr2, dr, p, dr, k * dr, nxt, d2 * dr2, dr * d2 * dr2, out, d2 * dr2,
                   k * d2 * dr2, bsz)
            self.sites[i + 1] = out

    def svd_compress(self, chi=None, threshold=1e-15, normalise=False):
        """svd_compress for every network (onedim_core.py:463-484): orthogonalising sweep to the right, norm taken
        out of the last tensor, truncating SVD sweep to the left (singular values normalised by the largest for
        the threshold test, then [:chi] -- :333-339), norm put back into the first tensor unless ``normalise``.
        Leaves the batch in right canonical form; per-bond singular values in ``self.singular_values``."""
        bsz = self.batch
        n = len(self)
        npd = dv._T2NP[self.sites[0].dtype]
        self.left_canonise()
        norm0 = self._frob(n - 1)                                  # norm(canonical_form="left")
        _scale_rows(self.sites[n - 1], norm0, mode=2)              # self[-1].data /= norm
        self.singular_values = []
        lib = _lib.load()
        # product of the largest singular values divided out bond by bond (`norm` of onedim_core.py:299,334)
        norm_acc = dv.DevArray.from_host(np.ones((bsz, 1))).t
        for i in range(n - 1, 0, -1):
            a = self.sites[i]
            dl, d, dr = a.shape[1:]
            # the reversed chain factorises M'[(p r), l] (onedim_core.py:317 on the reversed network)
            mt = _contig(a.reshape(bsz, dl, d * dr).permute(0, 2, 1))
            u, s, p, _ = _svd_project_batched(mt)                  # u (B, d dr, k), p = diag(s) vh (B, k, dl)
            k = u.shape[2]
            info = dv._empty((bsz, 2), np.float64)
            _lib.check(lib.tnb_truncation_count_batched(_ptr(s), k, k, bsz, int(chi or 0), float(threshold), 2, _ptr(info),
                                                        None, dv.stream_ptr()))
            info_h = info.cpu().numpy()                            # the kept ranks fix the next shapes: one read per site
            kept = info_h[:, 0].astype(np.int64)
            if np.any(info_h[:, 1] == 0.0):
                raise RaggedBatchError("a network of the batch has norm zero at site %d" % i)
            if np.any(kept != kept[0]):
                raise RaggedBatchError("kept bond dimensions differ inside the batch at site %d: %s" % (i, sorted(set(kept))))
            kk = int(kept[0])
            self.singular_values.append(np.asarray(dv.DevArray(s)))
            # site i <- rows of U'^T: (kk, d, dr), right-orthonormal
            self.sites[i] = _contig(u[:, :, :kk].permute(0, 2, 1)).reshape(bsz, kk, d, dr)
            # site i-1 <- site i-1 . (S V'h)^T[:, :kk] / s0  (V, then S normalised by its largest entry, absorbed:
            # onedim_core.py:333,347-349); the s0 factors are collected in norm_acc and put back at the end (:313)
            prev = self.sites[i - 1]
            pl, pd, _ = prev.shape[1:]
            out = dv._empty((bsz, pl, pd, kk), npd)
            _bgemm(_lib.OP_N, _lib.OP_T, pl * pd, kk, dl, prev, dl, pl * pd * dl, p, dl, k * dl, out, kk, pl * pd * kk, bsz)
            s0 = _contig(s[:, 0])
            _scale_rows(out, s0, mode=2)
            _scale_rows(norm_acc, s0, mode=0)
            self.sites[i - 1] = out
        _scale_rows(self.sites[0], norm_acc.reshape(bsz), mode=0)  # self[i].data *= norm (last tensor of the sweep)
        if not normalise:
            _scale_rows(self.sites[0], norm0, mode=0)              # self[0].data *= norm
        return self


def overlap_norm_compress_batched(seed, networks, nsites=64, physdim=4, bonddim=128, chi=64):
    """The cfg 4 unit of work for a list of network indices, batched: <a|b>, |a|, a.svd_compress(chi).
    -> list of flat float64 records [network, overlap, norm, norm_after, bonds...] (the layout of
    batch.overlap_norm_compress)."""
    a = BatchedMPS.random_uniform(seed, networks, 0, nsites, physdim, bonddim)
    b = BatchedMPS.random_uniform(seed, networks, 1, nsites, physdim, bonddim)
    ov = a.inner_product(b)
    nrm = a.norm()
    a.svd_compress(chi=chi)
    after = a.norm_canonical("right")
    bonds = [float(x) for x in a.bonddims()]
    return [np.array([float(net), float(np.real(ov[j])), float(nrm[j]), float(after[j])] + bonds)
            for j, net in enumerate(networks)], a
