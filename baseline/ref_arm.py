"""CPU arm of bench.py: the reference's own implementation of the hot path, timed on the host cores.

Two back ends, tried in this order:

* kind "reference": the UNMODIFIED reference package, installed by ``__graft_entry__.build()`` into
  ``baseline/_ref`` (``pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>``; the
  directory is git-ignored but travels to the GPU box).  The only harness-side patch is the two-line NumPy-2
  shim (``np.product = np.prod; np.float = float``) executed before the import -- NumPy >= 2 removed both and
  the reference's tensor.py:629,817,910,1040 still use them.  Everything timed goes through the reference's
  public API (``MatrixProductState``, ``contract_mps_mpo``, ``svd_compress``, ``tensor_qr``, ...).
* kind "port": ``oracle/tn_oracle.py``, the NumPy restatement (same LAPACK calls), when ``baseline/_ref`` is
  missing.

Nothing here touches the CUDA library.  Used only by ``bench.py --impl reference`` and by the
``cpu_baseline`` leg of the GPU arm.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is defined on ALL host cores."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def load_backend():
    """-> (kind, module-like namespace with the calls the arm needs)"""
    if os.path.isdir(os.path.join(REF_DIR, "tncontract")):
        if not hasattr(np, "product"):
            np.product = np.prod      # NumPy-2 shim, see module docstring
        if not hasattr(np, "float"):
            np.float = float
        sys.dont_write_bytecode = True
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import tncontract as tn
        import tncontract.onedim as od
        return "reference", _RefBackend(tn, od)
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import tn_oracle as o
    return "port", _PortBackend(o)


class _RefBackend:
    """The reference's public API (signatures of /root/reference/tncontract)."""

    def __init__(self, tn, od):
        self.tn, self.od = tn, od

    def mps(self, sites):
        return self.od.MatrixProductState([self.tn.Tensor(a, ["phys", "left", "right"]) for a in sites])

    def mpo(self, ws, wl):
        return self.od.MatrixProductOperator([self.tn.Tensor(w, l) for w, l in zip(ws, wl)], "left", "right",
                                             "physout", "physin")

    def canonise(self, psi):
        psi.left_canonise(qr_decomposition=True, normalise=True)

    def sweep(self, psi, H, chi):
        phi = self.od.contract_mps_mpo(psi, H)                  # onedim_core.py:1691
        phi.svd_compress(chi=chi)                               # onedim_core.py:463
        return phi

    def norm(self, phi):
        return float(phi.norm(canonical_form="right"))

    def bonds(self, phi):
        return [int(b) for b in phi.bonddims()]

    def bulk_site(self, A, W, nxt, Qs, Qn, chi):
        """One bulk site's work through the reference's tensor functions (the calls left_canonise /
        contract_mps_mpo make at onedim_core.py:1702-1704, :284-292, :317-349)."""
        tn = self.tn
        T = tn.contract(A, W, "phys", "physin")
        T.consolidate_indices()
        Q, R = tn.tensor.tensor_qr(T, ["physout", "left"])   # not in the package's __all__ (tensor.py:4-6)
        tn.contract(R, nxt, "right", "left")
        U, S, V = tn.tensor_svd(Qs, ["physout", "left"])
        s = np.diag(S.data)
        s = s / s[0]
        k = min(chi, int(np.sum(s > 1e-15)))
        V.data = V.data[:k]
        t = tn.contract(V, Qn, "right", "left")
        tn.contract(tn.Tensor(np.diag(s[:k]).astype(complex), ["a", "svd_out"]), t, "svd_out", "svd_out")

    def tensor(self, data, labels):
        return self.tn.Tensor(data, labels)

    def mps_labelled(self, sites):
        return self.od.MatrixProductState([self.tn.Tensor(a, l) for a, l in sites])

    def overlap_norm_compress(self, a, b, chi):
        self.od.inner_product_mps(a, b)          # onedim_core.py:1666
        a.norm()                                 # :642
        a.svd_compress(chi=chi)                  # :463


class _PortBackend:
    def __init__(self, o):
        self.o = o

    def mps(self, sites):
        o = self.o
        return o.Chain([o.OT(a, ["phys", "left", "right"]) for a in sites], "left", "right", "phys")

    def mpo(self, ws, wl):
        o = self.o
        return o.Chain([o.OT(w, l) for w, l in zip(ws, wl)], "left", "right", physout="physout", physin="physin")

    def canonise(self, psi):
        self.o.left_canonise(psi, qr_decomposition=True, normalise=True)

    def sweep(self, psi, H, chi):
        phi = self.o.contract_mps_mpo(psi, H)
        self.o.svd_compress(phi, chi=chi)
        return phi

    def norm(self, phi):
        return float(self.o.chain_norm(phi, canonical_form="right"))

    def bonds(self, phi):
        return [int(b) for b in phi.bonddims()]

    def bulk_site(self, A, W, nxt, Qs, Qn, chi):
        o = self.o
        T = o.consolidate(o.contract(A, W, "phys", "physin"))
        Q, R = o.tensor_qr(T, ["physout", "left"])
        o.contract(R, nxt, "right", "left")
        U, S, V = o.tensor_svd(Qs, ["physout", "left"])
        s = np.diag(S.data)
        s = s / s[0]
        k = min(chi, int(np.sum(s > 1e-15)))
        V.data = V.data[:k]
        t = o.contract(V, Qn, "right", "left")
        o.contract(o.OT(np.diag(s[:k]).astype(complex), ["a", "svd_out"]), t, "svd_out", "svd_out")

    def tensor(self, data, labels):
        return self.o.OT(data, labels)

    def mps_labelled(self, sites):
        o = self.o
        return o.Chain([o.OT(a, l) for a, l in sites], "left", "right", "phys")

    def overlap_norm_compress(self, a, b, chi):
        o = self.o
        o.inner_product_mps(a, b)
        o.chain_norm(a)
        o.svd_compress(a, chi=chi)


def network_seconds(host_a, host_b, chi_keep, reps=1):
    """The cfg 4 unit of work for ONE network on the CPU: <a|b>, |a|, a.svd_compress(chi).  host_a / host_b: lists
    of (ndarray, labels) as tncontract_b200.batch.random_mps(..., on_host=True) returns.  -> (kind, [seconds])"""
    use_all_host_threads()
    kind, be = load_backend()
    times = []
    for _ in range(reps):
        a = be.mps_labelled(host_a)
        b = be.mps_labelled(host_b)
        t0 = time.perf_counter()
        be.overlap_norm_compress(a, b, chi_keep)
        times.append(time.perf_counter() - t0)
    return kind, times


def full_sweep_seconds(host_sites, ws, wl, chi, max_steps, budget_s):
    """Time whole sweeps (the metric's unit) on the bench inputs: the canonisation of the input state is set-up
    and not timed, exactly as in the GPU arm.  Always runs one sweep; more (up to max_steps) only while the
    projected total stays inside budget_s.  -> (kind, [seconds per sweep], norm, bonds)"""
    use_all_host_threads()
    kind, be = load_backend()
    np.linalg.svd(np.random.default_rng(0).standard_normal((256, 256)))   # LAPACK / thread-pool warm-up
    psi = be.mps(host_sites)
    be.canonise(psi)
    H = be.mpo(ws, wl)
    times, phi = [], None
    t_start = time.perf_counter()
    while len(times) < max(1, max_steps):
        t0 = time.perf_counter()
        phi = be.sweep(psi, H, chi)
        times.append(time.perf_counter() - t0)
        if (time.perf_counter() - t_start) + times[-1] > budget_s:
            break
    return kind, times, be.norm(phi), be.bonds(phi)


def bulk_site_seconds(tfi_w, d, chi, D, reps):
    """The work of ONE bulk site of the sweep (chi, D=3: apply + consolidate, QR (d chi D) x (chi D), R-absorb,
    SVD (d chi) x (chi D), truncation, V/S-absorb) on synthetic operands of the bulk shapes.
    -> (kind, [seconds per site])"""
    use_all_host_threads()
    kind, be = load_backend()
    rng = np.random.default_rng(0)
    m = chi * D

    def rn(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    A = be.tensor(rn(d, chi, chi), ["phys", "left", "right"])
    W = be.tensor(tfi_w, ["left", "right", "physout", "physin"])
    nxt = be.tensor(rn(m, d, m), ["left", "physout", "right"])
    Qs = be.tensor(rn(d, m, chi), ["physout", "right", "left"])      # site as seen by the reversed SVD sweep
    Qn = be.tensor(rn(d, m, m), ["physout", "right", "left"])
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        be.bulk_site(A, W, nxt, Qs, Qn, chi)
        times.append(time.perf_counter() - t0)
    return kind, times
