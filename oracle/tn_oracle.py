"""CPU oracle for the tncontract hot path -- TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy restatement of the algorithms of the reference
(andrewdarmawan/tncontract, pure Python over NumPy/LAPACK) for the hot path
named in BASELINE.json: labelled pairwise contraction, QR / truncated SVD and
the MPS/MPO sweeps built on them.  It exists to CHECK the CUDA implementation
in ``tncontract_b200`` and to provide the CPU baseline timing in ``bench.py``.
Nothing under ``tncontract_b200/`` imports it; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm do.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (imported from /root/reference with the NumPy-2 shim) on seeded
inputs and stores inputs + outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those
fixtures (labels and bond dimensions exactly, numbers to 1e-12).

Representation: a tensor is ``OT(data: np.ndarray, labels: list[str])``; a
one-dimensional network is ``Chain(sites: list[OT], left, right, phys)``.
Every function cites the reference lines it restates (paths relative to
/root/reference/tncontract/).
"""
from __future__ import annotations

import uuid
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "OT", "Chain", "contract", "tensor_product", "trace_pair", "consolidate",
    "matricise", "tensor_svd", "tensor_qr", "truncated_svd", "left_canonise",
    "right_canonise", "svd_compress", "svd_compress_mps", "contract_mps_mpo",
    "ladder_contract", "inner_product_mps", "chain_norm", "con", "column_chain",
    "boundary_mps_contract", "inner_product_peps_network", "init_mps_random",
]


# --------------------------------------------------------------------------
# containers
# --------------------------------------------------------------------------
@dataclass
class OT:
    """ndarray + one label per axis (tensor.py:16-55; the ctor copies, :50)."""
    data: np.ndarray
    labels: list = field(default_factory=list)

    def __post_init__(self):
        self.data = np.array(self.data)
        self.labels = list(self.labels)
        if len(self.labels) != self.data.ndim:
            raise ValueError("Labels do not match shape of data.")  # tensor.py:151-155

    def copy(self):
        return OT(self.data.copy(), list(self.labels))

    @property
    def shape(self):
        return self.data.shape

    def dim(self, label):
        return self.data.shape[self.labels.index(label)]       # tensor.py:531-534

    def relabel(self, old, new):
        """tensor.py:164-179 (in place)."""
        old = old if isinstance(old, list) else [old]
        new = new if isinstance(new, list) else [new]
        self.labels = [new[old.index(l)] if l in old else l for l in self.labels]
        return self


@dataclass
class Chain:
    """1-D network: sites + the names of its virtual/physical labels
    (onedim_core.py:31-62, :186-199, :1329-1343).  ``phys`` is the MPS physical
    label; an MPO carries ``physout``/``physin`` instead."""
    sites: list
    left: str = "left"
    right: str = "right"
    phys: str = "phys"
    physout: str = "physout"
    physin: str = "physin"

    def __post_init__(self):
        out = []
        for t in self.sites:
            t = t.copy()                                        # onedim_core.py:54
            for lab in (self.left, self.right):                 # :59-61
                if lab not in t.labels:
                    t.data = t.data[np.newaxis]
                    t.labels.insert(0, lab)                     # tensor.py:506-513
            out.append(t)
        self.sites = out

    def copy(self):
        return Chain([t for t in self.sites], self.left, self.right, self.phys,
                     self.physout, self.physin)

    def __len__(self):
        return len(self.sites)

    def reverse(self):
        """onedim_core.py:84-88: reverse site order and swap the label NAMES."""
        self.sites = self.sites[::-1]
        self.left, self.right = self.right, self.left

    def bonddims(self):
        """onedim_core.py:168-175."""
        if not self.sites:
            return []
        return [self.sites[0].dim(self.left)] + [t.dim(self.right) for t in self.sites]

    def relabel(self, old, new):
        """onedim_core.py:380-398."""
        old = old if isinstance(old, list) else [old]
        new = new if isinstance(new, list) else [new]
        for t in self.sites:
            t.relabel(old, new)
        for attr in ("left", "right", "phys"):
            v = getattr(self, attr)
            if v in old:
                setattr(self, attr, new[old.index(v)])


def _uid():
    return str(uuid.uuid4())                                    # label.py:90-92


# --------------------------------------------------------------------------
# L1: labelled tensor primitives
# --------------------------------------------------------------------------
def _axes_for(labels, wanted):
    """Every axis whose label matches, label-list order (tensor.py:702-712)."""
    wanted = wanted if isinstance(wanted, list) else [wanted]
    ax = []
    for w in wanted:
        ax.extend(i for i, l in enumerate(labels) if l == w)
    return ax


def contract(a: OT, b: OT, la, lb, slice_a=None, slice_b=None) -> OT:
    """tensor.py:635-770: tensordot over label-selected axes; output labels are
    the free labels of ``a`` then those of ``b`` in stored order."""
    ia, ib = _axes_for(a.labels, la), _axes_for(b.labels, lb)
    if slice_a is not None:                                     # :716-731
        sel = [len(ia) - 1 if x == -1 else x for x in slice_a]
        ia = [ax for k, ax in enumerate(ia) if k in sel]
    if slice_b is not None:
        sel = [len(ib) - 1 if x == -1 else x for x in slice_b]
        ib = [ax for k, ax in enumerate(ib) if k in sel]
    out = np.tensordot(a.data, b.data, (ia, ib))                # :735
    labels = ([l for k, l in enumerate(a.labels) if k not in ia] +
              [l for k, l in enumerate(b.labels) if k not in ib])   # :764-768
    return OT(out, labels)


def tensor_product(*ts) -> OT:
    """tensor.py:773-779."""
    acc = ts[0]
    for t in ts[1:]:
        acc = contract(acc, t, [], [])
    return acc


def trace_pair(t: OT, l1, l2, k1=0, k2=0) -> OT:
    """tensor.py:318-334 (contract_internal / trace / tr); returns a new OT."""
    a1 = [i for i, l in enumerate(t.labels) if l == l1][k1]
    a2 = [i for i, l in enumerate(t.labels) if l == l2][k2]
    data = np.trace(t.data, axis1=a1, axis2=a2)
    return OT(data, [l for i, l in enumerate(t.labels) if i not in (a1, a2)])


def consolidate(t: OT, only=()) -> OT:
    """tensor.py:340-370: merge equal-label axes, labels sorted alphabetically
    (restricted to ``only`` when given).  Returns a new OT."""
    data, labels = t.data, list(t.labels)
    uniq = sorted(set(labels))
    if len(only):
        uniq = [u for u in uniq if u in only]
    for p, lab in enumerate(uniq):
        idx = [i for i, l in enumerate(labels) if l == lab]
        data = np.transpose(data, _perm_from_rolls(data.ndim, idx, p))   # :353-354
        merged = int(np.prod(data.shape[p:p + len(idx)]))
        data = np.reshape(data, data.shape[:p] + (merged,) + data.shape[p + len(idx):])
        labels = [l for l in labels if l != lab]
        labels.insert(p, lab)                                   # :367-370
    return OT(data, labels)


def _perm_from_rolls(ndim, idx, p):
    """Axis permutation equal to rolling axes idx[k] -> p+k in turn."""
    order = list(range(ndim))
    for k, q in enumerate(idx):
        # after the previous rolls axis q (q >= p+k always) is still at q
        order.insert(p + k, order.pop(q))
    return order


def matricise(t: OT, rows, dedupe=False):
    """Move the axes labelled ``rows`` to the front and flatten to 2-D.

    ``dedupe=False`` is the tensor_svd / tensor_to_matrix rule (tensor.py:
    903-905, 813-815: first axis with each label, one move_index per label);
    ``dedupe=True`` the tensor_qr rule (move_indices, tensor.py:444-482: every
    axis with the label, duplicates of the label list dropped).  Returns
    (matrix, row_shape, col_shape, col_labels)."""
    labels = list(t.labels)
    order = list(range(t.data.ndim))
    if not dedupe:
        for i, lab in enumerate(rows):
            k = labels.index(lab)
            labels.insert(i, labels.pop(k))
            order.insert(i, order.pop(k))
        nrow_axes = len(rows)
    else:
        uniq = []
        for lab in rows:
            if lab not in uniq:
                uniq.append(lab)
        picked = []
        for lab in uniq:
            picked.extend(i for i, l in enumerate(t.labels) if l == lab)
        rest = [i for i in range(t.data.ndim) if i not in picked]
        order = picked + rest
        labels = [t.labels[i] for i in order]
        nrow_axes = len(picked)
    data = np.transpose(t.data, order)
    rshape, cshape = data.shape[:nrow_axes], data.shape[nrow_axes:]
    nrow = int(np.prod(rshape)) if rshape else 1
    mat = np.reshape(data, (nrow, -1))
    col_labels = [l for l in labels if l not in rows]           # :907, :1037
    return mat, rshape, cshape, col_labels


def tensor_svd(t: OT, rows, svd_label="svd_"):
    """tensor.py:830-959 (absorb=None branch): U[rows..,svd_in],
    S[svd_out,svd_in] dense diagonal, V[svd_out,cols..] (V is V^H)."""
    rows = list(rows)
    mat, rshape, cshape, col_labels = matricise(t, rows)
    u, s, vh = np.linalg.svd(mat, full_matrices=False)          # :915
    U = OT(u.reshape(tuple(rshape) + (u.shape[1],)), rows + [svd_label + "in"])
    V = OT(vh.reshape((vh.shape[0],) + tuple(cshape)), [svd_label + "out"] + col_labels)
    S = OT(np.diag(s), [svd_label + "out", svd_label + "in"])
    return U, S, V


def tensor_qr(t: OT, rows, qr_label="qr_"):
    """tensor.py:962-1055."""
    rows = rows if isinstance(rows, list) else [rows]
    mat, rshape, cshape, col_labels = matricise(t, rows, dedupe=True)
    q, r = np.linalg.qr(mat, mode="reduced")                    # :1044
    # the reference reshapes Q with len(row_labels) leading axes (:1047)
    Q = OT(q.reshape(tuple(rshape) + (q.shape[1],)), rows + [qr_label + "in"])
    R = OT(r.reshape((r.shape[0],) + tuple(cshape)), [qr_label + "out"] + col_labels)
    return Q, R


def truncated_svd(t: OT, rows, chi=0, threshold=1e-15, absorb="right", absolute=True):
    """tensor.py:1112-1182.  Returns (U, V, discarded) or (U, S, V) when
    ``absorb is None``."""
    U, S, V = tensor_svd(t, rows)
    s = np.diag(S.data)
    keep, cut1 = (s[:chi], s[chi:]) if chi else (s, np.array([]))   # :1137-1142
    bar = threshold if absolute else s[0] * threshold               # :1146-1152
    cut2, keep = keep[keep <= bar], keep[keep > bar]
    discarded = np.concatenate((cut2, cut1), axis=0)
    k = len(keep)
    S = OT(np.diag(keep), S.labels)
    U = OT(U.data[..., :k], U.labels)                               # :1160-1162
    V = OT(V.data[:k], V.labels)
    if absorb is None:
        return U, S, V
    if absorb == "left":
        return contract(U, S, ["svd_in"], ["svd_out"]), V, discarded
    if absorb == "right":
        return U, contract(S, V, ["svd_in"], ["svd_out"]), discarded
    rt = OT(np.sqrt(S.data), S.labels)
    return (contract(U, rt, ["svd_in"], ["svd_out"]),
            contract(rt, V, ["svd_in"], ["svd_out"]), discarded)


# --------------------------------------------------------------------------
# L3: MPS sweeps
# --------------------------------------------------------------------------
def _zero_state(mps: Chain):
    """onedim_core.py:274-278 / 305-309 / 323-330."""
    for t in mps.sites:
        d = t.dim(mps.phys)
        t.data = np.zeros((d, 1, 1))
        t.labels = [mps.phys, mps.left, mps.right]


def left_canonise(mps: Chain, start=0, end=-1, chi=None, threshold=1e-14,
                  normalise=False, qr=False, record=None):
    """onedim_core.py:218-358 (in place).  ``record`` (a list) receives the
    normalised singular values of every SVD step, as kept before the chi cut."""
    N = len(mps)
    end = N if end == -1 else end
    S_ = mps.sites
    if qr:                                                       # :267-294
        for i in range(start, end):
            if i == N - 1:
                nrm = np.linalg.norm(S_[i].data)
                if nrm == 0.0:
                    _zero_state(mps)
                    return
                if normalise and start == 0:
                    S_[i].data = S_[i].data / nrm
                return
            tag = _uid()
            Q, R = tensor_qr(S_[i], [mps.phys, mps.left], qr_label=tag)
            S_[i] = Q.relabel(tag + "in", mps.right)
            nxt = contract(R, S_[i + 1], mps.right, mps.left)
            S_[i + 1] = nxt.relabel(tag + "out", mps.left)
        return
    acc = 1                                                      # :296-358
    for i in range(start, end):
        if i == N - 1:
            nrm = np.linalg.norm(S_[i].data)
            if nrm == 0.0:
                _zero_state(mps)
                return
            if normalise and start == 0:
                S_[i].data = S_[i].data / nrm
            else:
                S_[i].data = S_[i].data * acc
            return
        tag = _uid()
        U, S, V = tensor_svd(S_[i], [mps.phys, mps.left], svd_label=tag)
        s = np.diag(S.data)
        if s[0] == 0.0:
            _zero_state(mps)
            return
        acc = acc * s[0]
        s = s / s[0]                                             # :333-334
        if record is not None:
            record.append(s.copy())
        s = s[s > threshold]                                     # :336-339
        if chi:
            s = s[:chi]
        k = len(s)
        S_[i] = OT(U.data[:, :, :k], U.labels).relabel(tag + "in", mps.right)
        Vk = OT(V.data[:k], V.labels)
        nxt = contract(Vk, S_[i + 1], mps.right, mps.left)       # :347
        nxt = contract(OT(np.diag(s), S.labels), nxt, [tag + "in"], [tag + "out"])  # :349
        S_[i + 1] = nxt.relabel(tag + "out", mps.left)
        if i == end - 1:
            S_[i + 1].data = S_[i + 1].data * acc                # :357-358


def right_canonise(mps: Chain, start=0, end=-1, **kw):
    """onedim_core.py:360-378."""
    mps.reverse()
    N = len(mps)
    end = N if end == -1 else end
    left_canonise(mps, start=N - end, end=N - start, **kw)
    mps.reverse()


def chain_norm(mps: Chain, canonical_form=False):
    """onedim_core.py:642-660."""
    if canonical_form == "left":
        return np.linalg.norm(mps.sites[-1].data)
    if canonical_form == "right":
        return np.linalg.norm(mps.sites[0].data)
    return np.sqrt(inner_product_mps(mps, mps))


def svd_compress(mps: Chain, chi=None, threshold=1e-15, normalise=False,
                 reverse=False, record=None):
    """onedim_core.py:463-484 (in place): QR sweep, normalise, truncating SVD
    sweep back; result right-canonical (left-canonical when ``reverse``)."""
    if reverse:
        mps.reverse()
    left_canonise(mps, normalise=False, qr=True)
    nrm = chain_norm(mps, "left")
    mps.sites[-1].data = mps.sites[-1].data / nrm
    right_canonise(mps, chi=chi, threshold=threshold, normalise=False, record=record)
    if not normalise:
        mps.sites[0].data = mps.sites[0].data * nrm
    if reverse:
        mps.reverse()


def svd_compress_mps(mps: Chain, chi, threshold=1e-15, normalise=False, record=None):
    """onedim_core.py:1468-1473: SVD sweep without chi, then SVD sweep with chi."""
    a = mps.copy()
    left_canonise(a, chi=0, threshold=threshold, normalise=normalise)
    b = a.copy()
    right_canonise(b, chi=chi, threshold=threshold, normalise=normalise, record=record)
    return b


def contract_mps_mpo(mps: Chain, mpo: Chain) -> Chain:
    """onedim_core.py:1691-1708."""
    out = []
    for a, w in zip(mps.sites, mpo.sites):
        out.append(consolidate(contract(a, w, mps.phys, mpo.physin)))
    return Chain(out, mps.left, mps.right, mpo.physout)


def _drop_dummies(t: OT, only=None) -> OT:
    """tensor.py:515-529.  The reference walks the ORIGINAL label list and
    shape; on a hit it moves the FIRST axis currently carrying that label to
    the front (move_index, :380-394) and slices it away."""
    out = t.copy()
    for lab, d in zip(list(t.labels), t.shape):
        if d == 1 and (only is None or lab in only):
            k = out.labels.index(lab)
            out.data = np.moveaxis(out.data, k, 0)[0]
            out.labels.pop(k)
    return out


def ladder_contract(a: Chain, b: Chain, la, lb, start=0, end=None, conj_a=False,
                    left_out="left", right_out="right", intermediates=False):
    """onedim_core.py:1491-1663."""
    if end is None:
        end = min(len(a), len(b)) - 1
    if end < start:
        raise ValueError("Badly defined interval (end before start).")
    a, b = a.copy(), b.copy()
    if conj_a:
        for t in a.sites:
            t.data = t.data.conjugate()
    a.relabel([a.left, a.right], [_uid(), _uid()])
    b.relabel([b.left, b.right], [_uid(), _uid()])
    rung = _uid()
    a.relabel(la, rung)
    b.relabel(lb, rung)
    A, B = a.sites, b.sites
    steps = []

    def snapshot(C, olds, news, front):
        t = C.copy().relabel(olds, news)
        t = _drop_dummies(t, [x for x in t.labels if x not in news])
        steps.insert(0, t) if front else steps.append(t)

    if start == 0:                                               # :1587-1607
        olds, news = [a.right, b.right], [right_out + "1", right_out + "2"]
        for i in range(end + 1):
            if i == 0:
                C = contract(A[0], B[0], rung, rung)
            else:
                C = contract(C, A[i], a.right, a.left)
                C = contract(C, B[i], [b.right, rung], [b.left, rung])
            if intermediates:
                snapshot(C, olds, news, False)
        C = _drop_dummies(C.relabel(olds, news))
    elif end == len(a) - 1 and end == len(b) - 1:                # :1609-1629
        olds, news = [a.left, b.left], [left_out + "1", left_out + "2"]
        for i in range(end, start - 1, -1):
            if i == end:
                C = contract(A[end], B[end], rung, rung)
            else:
                C = contract(C, A[i], a.left, a.right)
                C = contract(C, B[i], [b.left, rung], [b.right, rung])
            if intermediates:
                snapshot(C, olds, news, True)
        C = _drop_dummies(C.relabel(olds, news))
    else:                                                        # :1631-1658
        olds = [a.right, b.right, a.left, b.left]
        news = [right_out + "1", right_out + "2", left_out + "1", left_out + "2"]
        for i in range(start, end + 1):
            t = contract(A[i], B[i], rung, rung)
            C = t if i == start else contract(C, t, [a.right, b.right], [a.left, b.left])
            if intermediates:
                s = C.copy().relabel(olds, news)
                s = _drop_dummies(s, [x for x in s.labels if x not in news])
                steps.append(_drop_dummies(s))
        C = _drop_dummies(C.relabel(olds, news))
    return steps if intermediates else C


def inner_product_mps(bra: Chain, ket: Chain, conj_bra=True, whole=False):
    """onedim_core.py:1666-1683."""
    t = ladder_contract(bra, ket, bra.phys, ket.phys, conj_a=conj_bra)
    return t if whole else t.data


# --------------------------------------------------------------------------
# L2: con
# --------------------------------------------------------------------------
def con(tensors, pairs) -> OT:
    """tncon.py:68-159 for the common case (list of tensors, list of label
    pairs, labels unique across tensors): internal edges traced first, then
    tensor pairs contracted in argument order (multi-edges between the same
    two tensors in one tensordot), finally the outer product of components."""
    ts = [t.copy() for t in tensors]
    flat = [x for p in pairs for x in p]
    if len(set(flat)) != len(flat):
        raise ValueError("Index found in more than one contraction pair.")
    where = {}
    for i, t in enumerate(ts):
        for lab in t.labels:
            if lab in where:
                raise ValueError("Index label " + lab + " found in two tensors.")
            where[lab] = i
    internal, pairwise, seen = [], [], []
    for c in pairs:
        i, j = where[c[0]], where[c[1]]
        if i == j:
            internal.append(c)
            continue
        key = (min(i, j), max(i, j))
        # tncon.py:112 looks the UNSORTED tuple up, so grouping only happens
        # when the pair is given in ascending tensor order
        if key in seen and (i, j) in seen:
            k = seen.index((i, j))
            if not isinstance(pairwise[k][0], list):
                pairwise[k] = [[pairwise[k][0]], [pairwise[k][1]]]
            pairwise[k][0].append(c[0])
            pairwise[k][1].append(c[1])
        elif key in seen:
            raise ValueError("tncon.py:112 raises here (unsorted tuple lookup)")
        else:
            pairwise.append(list(c))
            seen.append(key)
    for c in internal:
        i = where[c[0]]
        ts[i] = trace_pair(ts[i], c[0], c[1])
    comp = list(range(len(ts)))
    for c in pairwise:
        l0 = c[0][0] if isinstance(c[0], list) else c[0]
        l1 = c[1][0] if isinstance(c[1], list) else c[1]
        d, e = where[l0], where[l1]
        if d == e:
            ts[d] = trace_pair(ts[d], c[0], c[1])
            continue
        if d < e:
            ts[d] = contract(ts[d], ts[e], c[0], c[1])
            comp[e] = d
        else:
            ts[e] = contract(ts[e], ts[d], c[1], c[0])
            comp[d] = e
        for lab in ts[min(d, e)].labels:
            where[lab] = min(d, e)
    return tensor_product(*[ts[comp.index(x)] for x in set(comp)])


# --------------------------------------------------------------------------
# L4: square-lattice boundary-MPS contraction
# --------------------------------------------------------------------------
def _with_dummy(t: OT, lab):
    t = t.copy()
    if lab not in t.labels:
        t.data = t.data[np.newaxis]
        t.labels.insert(0, lab)
    return t


def column_chain(grid, col, up="up", right="right", down="down", left="left") -> Chain:
    """square_lattice.py:394-417: column -> MPS (edge columns) or MPO."""
    ncols = len(grid[0])
    sites = [grid[r][col] for r in range(len(grid))]
    if col == 0 or col == ncols - 1:
        phys, drop = (right, left) if col == 0 else (left, right)
        ch = Chain(sites, up, down, phys)
        ch.sites = [_drop_dummies(t, drop) for t in ch.sites]
        return ch
    return Chain(sites, up, down, physout=right, physin=left)


def boundary_mps_contract(grid, chi, tolerance=1e-14, labels=("up", "right", "down", "left"),
                          bond_log=None):
    """square_lattice.py:130-203 with compression_type='svd': returns the
    np.longdouble scalar data of the final rank-0 tensor."""
    up, right, down, left = labels
    grid = [[_with_dummy(_with_dummy(_with_dummy(_with_dummy(t, left), right), up), down)
             for t in row] for row in grid]                      # :49-53
    ncols = len(grid[0])
    norm = np.longdouble(1)
    cur = None
    for col in range(ncols - 1):
        if col == 0:
            todo = column_chain(grid, 0, up, right, down, left)
        else:
            todo = contract_mps_mpo(cur, column_chain(grid, col, up, right, down, left))
        cur = svd_compress_mps(todo, chi, normalise=False, threshold=tolerance)
        nrm = chain_norm(cur, "right")
        if nrm == 0.0:
            return 0.0
        cur.sites[0].data = cur.sites[0].data / nrm
        norm *= nrm
        if bond_log is not None:
            bond_log.append(cur.bonddims())
    last = column_chain(grid, ncols - 1, up, right, down, left)
    t = inner_product_mps(cur, last, conj_bra=False, whole=True)
    return t.data * norm                                         # :196-198


def inner_product_peps_network(ket, bra):
    """square_lattice.py:283-292: double-layer grid of consolidated tensors."""
    out = []
    for rk, rb in zip(ket, bra):
        row = []
        for k, b in zip(rk, rb):
            t = contract(OT(b.data.conjugate(), b.labels), k, "phys", "phys")
            row.append(consolidate(t))
        out.append(row)
    return out


# --------------------------------------------------------------------------
# constructors used by the benchmarks
# --------------------------------------------------------------------------
def init_mps_random(nsites, physdim, bonddim=1, left="left", right="right", phys="phys",
                    rand=None) -> Chain:
    """onedim_utils.py:23-60: U[0,1) entries, each site divided by its largest
    singular value through an SVD round trip."""
    rand = np.random.rand if rand is None else rand
    pd = physdim if np.iterable(physdim) else [physdim] * nsites
    bd = bonddim if np.iterable(bonddim) else [bonddim] * (nsites - 1)
    bd = [1] + list(bd) + [1]
    sites = []
    for i in range(nsites):
        rt = OT(rand(pd[i], bd[i], bd[i + 1]), [phys, left, right])
        U, S, V = tensor_svd(rt, [phys, left])
        S.data = S.data / S.data[0, 0]
        rt = contract(U, S, "svd_in", "svd_out")
        rt = contract(rt, V, "svd_in", "svd_out")
        sites.append(rt)
    return Chain(sites, left, right, phys)
