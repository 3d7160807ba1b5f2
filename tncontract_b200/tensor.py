"""Labelled tensors on the GPU: the ``tn.Tensor`` API of tncontract with the
data held in a :class:`~tncontract_b200.devarray.DevArray` and every numeric
operation executed by libtnb (hand-written sm_100a kernels behind the C ABI
in include/tnb.h).

Only the label algebra lives here; it is written so that label order, shapes,
return arities and exception types are those of the reference
(/root/reference/tncontract/tensor.py, lines cited per function).  Axis moves
are lazy views, exactly like the reference's ``np.rollaxis``; data is only
touched by kernels: tnb_permute (reshape copies), tnb_tensordot/tnb_gemm
(contract), tnb_qr, tnb_svd, tnb_trace and the small elementwise helpers.
"""
import numpy as np

from . import devarray as dv
from . import label as lbl
from .devarray import DevArray

__all__ = ['Tensor', 'contract', 'distance', 'matrix_to_tensor',
           'tensor_to_matrix', 'random_tensor', 'tensor_product', 'tensor_svd',
           'truncated_svd', 'zeros_tensor']


def _prod(xs):
    p = 1
    for x in xs:
        p *= int(x)
    return p


def _to_device(data):
    """Tensor() always owns a private copy of its input (tensor.py:50)."""
    if isinstance(data, np.ndarray) and data.dtype in (np.longdouble, np.clongdouble):
        return np.array(data)  # host-only tail of twodim.mps_contract (square_lattice.py:183-186)
    return dv.asdevarray(data, copy=True)


class Tensor():
    """A device array plus one label per axis (tensor.py:16-55)."""

    def __init__(self, data, labels=None, base_label="i"):
        labels = [] if labels is None else labels
        self.data = _to_device(data)
        if len(labels) == 0:
            self.assign_labels(base_label=base_label)
        else:
            self.labels = labels

    @classmethod
    def _wrap(cls, data, labels):
        """Adopt a freshly produced device array without the defensive copy."""
        t = cls.__new__(cls)
        t.data = data
        t.labels = labels
        return t

    # ---- persistence: the reference pickles a Tensor as {"data": ndarray, "_labels": list} -----------------
    def __getstate__(self):
        return {"data": np.asarray(self.data), "_labels": list(self._labels)}

    def __setstate__(self, state):
        self.data = _to_device(state["data"])
        self._labels = [l.decode() if isinstance(l, bytes) else l for l in state["_labels"]]

    # ---- printing / comparison (tensor.py:57-93) -------------------------------
    def __repr__(self):
        return "Tensor(data=%r, labels=%r)" % (np.asarray(self.data), self.labels)

    def __str__(self):
        rows = str(np.asarray(self.data)).splitlines()
        if len(rows) > 20:
            rows = rows[:20] + ["...", "Printed output of large array was truncated.\nString "
                                "representation of full data array returned by tensor.data.__str__()."]
        idx = "".join("   %d. (dim=%d) %s\n" % (i, self.shape[i], l) for i, l in enumerate(self.labels))
        return ("Tensor object: \nData type: " + str(self.data.dtype) + "\nNumber of indices: " +
                str(len(self.shape)) + "\n\nIndex labels:\n" + idx + "\nTensor data = \n" + "\n".join(rows))

    def __eq__(self, other):
        if not isinstance(other, Tensor):
            return False
        return self.labels == other.labels and np.array_equal(np.asarray(self.data), np.asarray(other.data))

    def __neq__(self, other):
        return not self.__eq__(other)

    # ---- scalar arithmetic (tensor.py:95-140) -------------------------------------
    def _scaled(self, other):
        try:
            res = self.data * other
        except TypeError:
            res = NotImplemented
        if res is NotImplemented:
            raise TypeError("unsupported operand type(s) *: for '" + self.__class__.__name__ + "' and '" +
                            other.__class__.__name__ + "'")
        return Tensor._wrap(res, list(self.labels))

    def __mul__(self, other):
        return self._scaled(other)

    def __rmul__(self, other):
        return self._scaled(other)

    def __add__(self, other):
        try:
            a, b = self.copy(), other.copy()
            a.consolidate_indices()
            b.consolidate_indices()
            return Tensor._wrap(a.data + b.data, a.labels)
        except Exception:
            raise TypeError("Can only add together tensors with the same"
                            " indices: labels and dimensions of each index must match.")

    def __getitem__(self, *args):
        """``A["a", "b"] * B["c", "d"]`` contraction shorthand (tensor.py:142-145)."""
        return ToContract(self, *args)

    # ---- labels (tensor.py:148-236) -----------------------------------------------------
    def get_labels(self):
        return self._labels

    def set_labels(self, labels):
        if len(labels) != len(self.data.shape):
            raise ValueError("Labels do not match shape of data.")
        self._labels = list(labels)

    labels = property(get_labels, set_labels)

    def assign_labels(self, base_label="i"):
        self.labels = [base_label + str(i) for i in range(len(self.data.shape))]

    def replace_label(self, old_labels, new_labels):
        old = old_labels if isinstance(old_labels, list) else [old_labels]
        new = new_labels if isinstance(new_labels, list) else [new_labels]
        for i, cur in enumerate(self.labels):
            if cur in old:
                self.labels[i] = new[old.index(cur)]

    def _reprime(self, labels, fn):
        # the reference aliases ``labels`` to self.labels when None and scans
        # that live list in the inner loop (tensor.py:194-201); kept as is
        if labels is None:
            labels = self.labels
        elif not isinstance(labels, list):
            labels = [labels]
        for i, cur in enumerate(self.labels):
            for plain in labels:
                if lbl.noprime_label(cur) == plain:
                    self.labels[i] = fn(self.labels[i])

    def prime_label(self, labels=None):
        self._reprime(labels, lbl.prime_label)

    def unprime_label(self, labels=None):
        self._reprime(labels, lbl.unprime_label)

    # ---- index plumbing: lazy axis moves, copies only in reshape ---------------------------
    def _permute(self, order):
        self.data = self.data.transpose(order)
        self._labels = [self._labels[i] for i in order]

    def move_index(self, label, position):
        """tensor.py:380-394: first axis labelled ``label`` goes to ``position``."""
        src = self.labels.index(label)
        order = list(range(self.rank))
        order.insert(position, order.pop(src))
        self._permute(order)

    def move_indices(self, labels, position, preserve_relative_order=False):
        """tensor.py:396-482: the selected axes end up as one block starting at
        ``position`` among the untouched axes (which keep their order)."""
        if not isinstance(labels, list):
            labels = [labels]
        if preserve_relative_order:
            picked = [i for i, l in enumerate(self.labels) if l in labels]
        else:
            seen, picked = [], []
            for l in labels:
                if l not in seen:
                    seen.append(l)
                    picked.extend(i for i, cur in enumerate(self.labels) if cur == l)
        if position + len(picked) > self.rank:
            # the reference has already pushed the axes to the back when it notices (tensor.py:470)
            self._permute([i for i in range(self.rank) if i not in picked] + picked)
            raise ValueError("Specified position too far right.")
        rest = [i for i in range(self.rank) if i not in picked]
        self._permute(rest[:position] + picked + rest[position:])

    def fuse_indices(self, indices_to_fuse, new_label, preserve_relative_order=False):
        """tensor.py:238-296."""
        self.move_indices(indices_to_fuse, 0, preserve_relative_order=preserve_relative_order)
        total, last = 1, None
        for i, l in enumerate(self.labels):
            if l in indices_to_fuse:
                total *= self.data.shape[i]
                last = i
        tail_labels = self.labels[last + 1:]
        self.data = self.data.reshape((total,) + tuple(self.data.shape[last + 1:]))
        self.labels = [new_label] + tail_labels

    def split_index(self, label, new_dims, new_labels):
        """tensor.py:298-316."""
        if len(new_dims) != len(new_labels):
            raise ValueError("Length of new_dims must equal length of new_labels")
        i = self.labels.index(label)
        shape = self.data.shape
        labels = self.labels[:i] + new_labels + self.labels[i + 1:]
        self.data = self.data.reshape(tuple(shape[:i]) + tuple(new_dims) + tuple(shape[i + 1:]))
        self.labels = labels

    def contract_internal(self, label1, label2, index1=0, index2=0):
        """tensor.py:318-334 (np.trace over two axes) -> tnb_trace."""
        a1 = [i for i, l in enumerate(self.labels) if l == label1][index1]
        a2 = [i for i, l in enumerate(self.labels) if l == label2][index2]
        keep = [l for i, l in enumerate(self.labels) if i not in (a1, a2)]
        if a1 == a2:
            raise ValueError("axis1 and axis2 cannot be the same")
        self.data = dv.trace(self.data, a1, a2)
        self.labels = keep

    trace = contract_internal
    tr = contract_internal

    def consolidate_indices(self, labels=[]):
        """tensor.py:340-370: equal labels merged, merged axes first in sorted
        label order.  One axis permutation + one reshape (a single tnb_permute)."""
        chosen = sorted(set(self.labels))
        if len(labels) != 0:
            chosen = [l for l in chosen if l in labels]
        groups = [[i for i, cur in enumerate(self.labels) if cur == l] for l in chosen]
        taken = [i for g in groups for i in g]
        rest = [i for i in range(self.rank) if i not in taken]
        shape = self.data.shape
        new_shape = [_prod(shape[i] for i in g) for g in groups] + [shape[i] for i in rest]
        new_labels = chosen + [self.labels[i] for i in rest]
        self.data = self.data.transpose(taken + rest).reshape(new_shape)
        self.labels = new_labels

    def sort_labels(self):
        self.consolidate_indices()

    def copy(self):
        return Tensor._wrap(self.data.copy(), list(self.labels))

    def conjugate(self):
        self.data = self.data.conjugate()

    def inv(self):
        """tensor.py:487-488.  The MPS code only ever inverts the diagonal
        matrices of singular values (onedim_core.py:1081-1084,1164-1167,1878);
        a diagonal matrix is inverted on device, anything else is refused."""
        if self.rank != 2 or self.shape[0] != self.shape[1]:
            raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")
        d = dv.diag_extract(self.data)
        full = float(self.data.norm())
        if d.dtype == np.float64 and np.isclose(float(d.norm()), full, rtol=1e-13, atol=0.0):
            if not np.all(np.asarray(d) != 0.0):
                raise np.linalg.LinAlgError("Singular matrix")    # as np.linalg.inv
            self.data = dv.diag_embed(d, np.float64, mode=2)      # real diagonal: 1 / s on the device
            return
        # general square matrix: A^-1 = V diag(1/s) U^H from the device SVD (np.linalg.inv raises on an exactly
        # singular matrix; so do we)
        u, s, vh = dv.svd(self.data.contiguous())
        if not float(np.asarray(s)[-1]) > 0.0:
            raise np.linalg.LinAlgError("Singular matrix")
        uh = u.transpose([1, 0]).conjugate().contiguous()         # k x m
        dv.diag_scale_rows(uh, s, mode=2)                          # rows scaled by 1 / s
        self.data = dv.tensordot(vh, uh, [0], [0], conj_a=True)    # (Vh)^H (S^-1 U^H)

    def add_suffix_to_labels(self, suffix):
        self.labels = [l + suffix for l in self.labels]

    def suf(self, suffix):
        t = self.copy()
        t.labels = [l + suffix for l in t.labels]
        return t

    def add_dummy_index(self, label, position=0):
        """tensor.py:506-513."""
        self.data = self.data[None]
        self._labels.insert(0, label)
        self.move_index(label, position)

    def remove_all_dummy_indices(self, labels=None):
        """tensor.py:515-529: walks the ORIGINAL labels/shape; each hit drops
        the first axis currently carrying that label."""
        for l, d in zip(list(self.labels), self.shape):
            if d == 1 and (labels is None or l in labels):
                k = self.labels.index(l)
                self.data = self.data.moveaxis(k, 0)[0]
                self._labels = self._labels[:k] + self._labels[k + 1:]

    def index_dimension(self, label):
        return self.data.shape[self.labels.index(label)]

    def to_matrix(self, row_labels):
        return tensor_to_matrix(self, row_labels)

    def pad_index(self, label, inc, before=False):
        """tensor.py:543-563 (np.pad with zeros along one axis)."""
        ax = self.labels.index(label)
        shape = list(self.shape)
        shape[ax] += inc
        # With the padded axis in front the old data is ONE contiguous slab of the new buffer: tnb_permute writes
        # it there directly; the result is handed back as a (lazy) view with the axis in its place.
        n = self.shape[ax]
        front = [ax] + [i for i in range(self.rank) if i != ax]
        out = DevArray.zeros([shape[ax]] + [shape[i] for i in front[1:]], self.data.dtype)
        slab = out[inc:] if before else out[:n]
        self.data.transpose(front)._permute_copy(out=slab)
        back = [front.index(i) for i in range(self.rank)]
        self.data = out.transpose(back)

    def contract(self, *args, **kwargs):
        t = contract(self, *args, **kwargs)
        self.data = t.data
        self.labels = t.labels

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def rank(self):
        return len(self.shape)

    def norm(self):
        """Frobenius norm (tensor.py:587-590) -> tnb_norm2."""
        return self.data.norm()


class ToContract():
    """Tensor + the labels to contract (tensor.py:593-613)."""

    def __init__(self, tensor, labels):
        self.tensor = tensor
        self.labels = labels

    def __mul__(self, other):
        l1 = list(self.labels) if isinstance(self.labels, tuple) else self.labels
        l2 = list(other.labels) if isinstance(other.labels, tuple) else other.labels
        return contract(self.tensor, other.tensor, l1, l2)


# ---- constructors ------------------------------------------------------------------
def random_tensor(*args, **kwargs):
    """tensor.py:618-623 (host RNG, then upload)."""
    labels = kwargs.pop("labels", [])
    base_label = kwargs.pop("base_label", "i")
    return Tensor(np.random.rand(*args), labels=labels, base_label=base_label)


def zeros_tensor(*args, **kwargs):
    """tensor.py:626-632."""
    labels = kwargs.pop("labels", [])
    dtype = kwargs.pop("dtype", float)
    base_label = kwargs.pop("base_label", "i")
    shape = args[0] if len(args) == 1 else args
    if isinstance(shape, (int, np.integer)):
        shape = (shape,)
    dt = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
    t = Tensor._wrap(DevArray.zeros(tuple(shape), dt), [base_label + str(i) for i in range(len(shape))])
    if len(labels):
        t.labels = labels
    return t


# ---- contraction ------------------------------------------------------------------------
def _matching_axes(tensor_labels, wanted):
    axes = []
    for w in wanted:
        axes.extend(i for i, l in enumerate(tensor_labels) if l == w)
    return axes


def contract(tensor1, tensor2, labels1, labels2, index_slice1=None, index_slice2=None):
    """tensor.py:635-770.  All axes carrying a listed label are contracted, in
    label-list order; output labels are free(tensor1) + free(tensor2).
    np.tensordot is replaced by tnb_tensordot: the operand permutations are
    folded into the DMMA GEMM's operand staging."""
    if not isinstance(labels1, list):
        labels1 = [labels1]
    if not isinstance(labels2, list):
        labels2 = [labels2]
    ax1 = _matching_axes(tensor1.labels, labels1)
    ax2 = _matching_axes(tensor2.labels, labels2)
    if index_slice1 is not None:
        pick = [len(ax1) - 1 if x == -1 else x for x in index_slice1]
        ax1 = [a for k, a in enumerate(ax1) if k in pick]
    if index_slice2 is not None:
        pick = [len(ax2) - 1 if x == -1 else x for x in index_slice2]
        ax2 = [a for k, a in enumerate(ax2) if k in pick]
    try:
        out = dv.tensordot(tensor1.data, tensor2.data, ax1, ax2)
    except ValueError as e:
        # same diagnostics as tensor.py:737-760 (including indexing the label
        # lists by axis position)
        if len(ax1) != len(ax2):
            raise ValueError('Number of indices in contraction does not match.')
        for i in range(len(ax1)):
            d1, d2 = tensor1.data.shape[ax1[i]], tensor2.data.shape[ax2[i]]
            if d1 != d2:
                raise ValueError(labels1[i] + ' with dim=' + str(d1) + ' does not match ' + labels2[i] +
                                 ' with dim=' + str(d2))
        for i in range(len(labels1)):
            if labels1[i] not in tensor1.labels:
                raise ValueError(labels1[i] + ' not in list of labels for tensor1')
            if labels2[i] not in tensor2.labels:
                raise ValueError(labels2[i] + ' not in list of labels for tensor2')
        raise e
    labels = ([l for i, l in enumerate(tensor1.labels) if i not in ax1] +
              [l for i, l in enumerate(tensor2.labels) if i not in ax2])
    return Tensor._wrap(out, labels)


def tensor_product(*args):
    """tensor.py:773-779."""
    t = args[0]
    for x in args[1:]:
        t = contract(t, x, [], [])
    return t


def distance(tensor1, tensor2):
    """tensor.py:782-802."""
    t1, t2 = tensor1.copy(), tensor2.copy()
    t1.consolidate_indices()
    t2.consolidate_indices()
    if t1.labels != t2.labels:
        raise ValueError("Input tensors must have the same labels.")
    return (t1.data - t2.data).norm()


def _rows_first(tensor, row_labels):
    """Axis order after ``move_index(label, i)`` for each row label in turn
    (tensor.py:813-815, 903-905): FIRST axis with each label."""
    labels = list(tensor.labels)
    order = list(range(len(labels)))
    for i, l in enumerate(row_labels):
        k = labels.index(l)
        labels.insert(i, labels.pop(k))
        order.insert(i, order.pop(k))
    return order, labels


def tensor_to_matrix(tensor, row_labels):
    """tensor.py:805-818; returns a 2-D device array."""
    order, _ = _rows_first(tensor, row_labels)
    data = tensor.data.transpose(order)
    nrow = _prod(data.shape[:len(row_labels)])
    ncol = int(_prod(data.shape) / nrow)
    return data.reshape((nrow, ncol))


def matrix_to_tensor(matrix, shape, labels=None):
    """tensor.py:821-827."""
    labels = [] if labels is None else labels
    data = dv.asdevarray(matrix, copy=True).reshape(tuple(shape))
    t = Tensor._wrap(data, ["i" + str(i) for i in range(len(shape))])
    if len(labels):
        t.labels = labels
    return t


# ---- factorisations ---------------------------------------------------------------------------
def _svd_parts(tensor, row_labels, svd_label, project=False):
    """Matricise + tnb_svd; returns U, s (device vector), V tensors.  With ``project`` the third
    tensor is P = diag(s) V (tnb_svd_project), the product the MPS sweeps absorb into the next site."""
    row_labels = list(row_labels)
    order, labels = _rows_first(tensor, row_labels)
    data = tensor.data.transpose(order)
    shape = data.shape
    nr = len(row_labels)
    m = _prod(shape[:nr])
    n = int(_prod(shape) / m) if m else 0
    col_labels = [l for l in labels if l not in row_labels]
    u, s, vh = (dv.svd_project if project else dv.svd)(data.reshape((m, n)))
    k = s.size
    U = Tensor._wrap(u.reshape(tuple(shape[:nr]) + (k,)), row_labels + [svd_label + "in"])
    V = Tensor._wrap(vh.reshape((k,) + tuple(shape[nr:])), [svd_label + "out"] + col_labels)
    return U, s, V


def tensor_svd(tensor, row_labels, svd_label="svd_", absorb_singular_values=None):
    """tensor.py:830-959.  U[rows.., svd_in], S[svd_out, svd_in] (dense real
    diagonal), V[svd_out, cols..] with V = V^H of the matrix SVD."""
    U, s, V = _svd_parts(tensor, row_labels, svd_label)
    S = Tensor._wrap(dv.diag_embed(s), [svd_label + "out", svd_label + "in"])
    if absorb_singular_values not in ("left", "right", "both"):
        return U, S, V
    # the reference contracts over the literal labels "svd_in"/"svd_out" here,
    # whatever svd_label is (tensor.py:944-957)
    if svd_label != "svd_":
        if absorb_singular_values == "left":
            return contract(U, S, ["svd_in"], ["svd_out"]), V
        if absorb_singular_values == "right":
            return U, contract(S, V, ["svd_in"], ["svd_out"])
        rt = Tensor._wrap(dv.diag_embed(s, mode=1), list(S.labels))
        return contract(U, rt, ["svd_in"], ["svd_out"]), contract(rt, V, ["svd_in"], ["svd_out"])
    mode = 1 if absorb_singular_values == "both" else 0
    if absorb_singular_values in ("left", "both"):
        U = Tensor._wrap(dv.diag_scale_cols(U.data.contiguous(), s, mode), U.labels)
    if absorb_singular_values in ("right", "both"):
        V = Tensor._wrap(dv.diag_scale_rows(V.data.contiguous(), s, mode), V.labels)
    return U, V


def tensor_qr(tensor, row_labels, qr_label="qr_"):
    """tensor.py:962-1055: Q[rows.., qr_in], R[qr_out, cols..] via tnb_qr."""
    t = Tensor._wrap(tensor.data, list(tensor.labels))  # views only; input untouched
    if not isinstance(row_labels, list):
        row_labels = [row_labels]
    t.move_indices(row_labels, 0)
    m = 1
    for i, l in enumerate(t.labels):
        if l not in row_labels:
            break
        m *= t.data.shape[i]
    col_labels = [l for l in t.labels if l not in row_labels]
    shape = t.data.shape
    n = int(_prod(shape) / m) if m else 0
    q, r = dv.qr(t.data.reshape((m, n)))
    nr = len(row_labels)
    k = q.shape[1]
    Q = Tensor._wrap(q.reshape(tuple(shape[:nr]) + (k,)), row_labels + [qr_label + "in"])
    R = Tensor._wrap(r.reshape((k,) + tuple(shape[nr:])), [qr_label + "out"] + col_labels)
    return Q, R


def tensor_lq(tensor, row_labels, lq_label="lq_"):
    """tensor.py:1058-1109: a QR on the column labels with the outputs renamed;
    returns (L, Q)."""
    col_labels = [l for l in tensor.labels if l not in row_labels]
    tmp = lbl.unique_label()
    Q, L = tensor_qr(tensor, col_labels, qr_label=tmp)
    Q.replace_label(tmp + "in", lq_label + "out")
    L.replace_label(tmp + "out", lq_label + "in")
    return L, Q


def truncated_svd(tensor, row_labels, chi=0, threshold=1e-15, absorb_singular_values="right", absolute=True):
    """tensor.py:1112-1182: keep ``[:chi]`` then the values above the
    (absolute or s0-relative) threshold; S absorbed left / right / sqrt both.
    Returns (U, V, discarded) or (U, S, V) when absorb_singular_values is None."""
    U, s_dev, V = _svd_parts(tensor, row_labels, "svd_")
    s = np.asarray(s_dev)  # the API returns the discarded values on the host
    keep, cut1 = (s[:chi], s[chi:]) if chi else (s, np.array([]))
    bar = threshold if absolute else s[0] * threshold
    cut2, keep = keep[keep <= bar], keep[keep > bar]
    discarded = np.concatenate((cut2, cut1), axis=0)
    k = len(keep)
    s_keep = s_dev[0:k]
    U = Tensor._wrap(U.data[..., 0:k], U.labels)   # a strided view, like tensor.py:1160-1162
    V = Tensor._wrap(V.data[0:k], V.labels)
    if absorb_singular_values is None:
        return U, Tensor._wrap(dv.diag_embed(s_keep), ["svd_out", "svd_in"]), V
    mode = 0 if absorb_singular_values in ("left", "right") else 1
    if absorb_singular_values != "right":
        U = Tensor._wrap(dv.diag_scale_cols(U.data.copy(), s_keep, mode), U.labels)
    if absorb_singular_values != "left":
        V = Tensor._wrap(dv.diag_scale_rows(V.data.copy(), s_keep, mode), V.labels)
    return U, V, discarded


def conjugate(tensor):
    """tensor.py:1185-1189."""
    t = tensor.copy()
    t.conjugate()
    return t
