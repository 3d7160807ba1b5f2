"""Square-lattice tensor networks: PEPS / PEPO containers and the
boundary-MPS contraction (reference: twodim/square_lattice.py; lines cited per
routine).  The column loop is host control flow; each column costs one
``contract_mps_mpo`` (fused tnb_mps_mpo_site kernels) and one
``svd_compress_mps`` (tnb_svd + tnb_tensordot sweeps).  The running norm is
kept on the host in ``np.longdouble`` exactly like the reference."""
import numpy as np

from .. import onedim as od
from .. import tensor as tsr


def _grid(rows, copy):
    """2-D object array of Tensors (built explicitly: np.array() would try to
    treat objects with __getitem__ as sequences)."""
    rows = [list(r) for r in rows]
    g = np.empty((len(rows), len(rows[0]) if rows else 0), dtype=object)
    for i, r in enumerate(rows):
        for j, x in enumerate(r):
            g[i, j] = x.copy() if copy else x
    return g


class SquareLatticeTensorNetwork():
    """Rectangular array of tensors joined up/down/left/right (square_lattice.py:17-232)."""

    def __init__(self, tensors, up_label="up", right_label="right", down_label="down", left_label="left",
                 copy_data=True):
        self.up_label, self.right_label = up_label, right_label
        self.down_label, self.left_label = down_label, left_label
        self.data = _grid(tensors, copy_data)
        for _, x in np.ndenumerate(self.data):  # open edges get dimension-1 bonds (:49-53)
            for lab in (left_label, right_label, up_label, down_label):
                if lab not in x.labels:
                    x.add_dummy_index(lab)

    def __iter__(self):
        return self.data.__iter__()

    def __len__(self):
        return self.data.__len__()

    def __getitem__(self, key):
        return self.data.__getitem__(key)

    def __setitem__(self, key, value):
        self.data.__setitem__(key, value)

    def _labels_kw(self):
        return dict(up_label=self.up_label, right_label=self.right_label, down_label=self.down_label,
                    left_label=self.left_label)

    def copy(self):
        return SquareLatticeTensorNetwork(self.data, copy_data=True, **self._labels_kw())

    @property
    def shape(self):
        return self.data.shape

    def is_left_right_periodic(self):
        return any(x.index_dimension(self.left_label) > 1 for x in self[:, 0])

    def can_contract(self):
        """All facing bonds of the bulk have equal dimension (:85-110)."""
        rows, cols = self.data.shape
        vertical, horizontal = [], []
        for i in range(1, rows):
            for j in range(1, cols):
                if self[i, j].index_dimension(self.up_label) != self[i - 1, j].index_dimension(self.down_label):
                    vertical.append((i, j))
                if self[i, j].index_dimension(self.left_label) != self[i, j - 1].index_dimension(self.right_label):
                    horizontal.append((i, j))
        if not vertical and not horizontal:
            return True
        print("Unmatched bonds found between the following sites:")
        for k in vertical:
            print("(" + str(k[0] - 1) + ", " + str(k[1]) + ")" + " and " + str(k))
        for k in horizontal:
            print("(" + str(k[0]) + ", " + str(k[1] - 1) + ")" + " and " + str(k))
        return False

    def exact_contract(self, until_column=-1):
        """Column-by-column exact contraction (:112-128); exponential cost."""
        C = od.contract_virtual_indices(column_to_mpo(self, 0))
        for i in range(1, self.data.shape[1]):
            if i == until_column + 1:
                return C
            C = od.contract_multi_index_tensor_with_one_dim_array(C, column_to_mpo(self, i), self.right_label,
                                                                  self.left_label)
            C.remove_all_dummy_indices([self.left_label, self.up_label, self.down_label])
        return C

    def mps_contract(self, chi, compression_type="svd", until_column=-1, max_iter=10, tolerance=1e-14,
                     return_all_columns=False):
        """Boundary-MPS contraction from the left (:130-203): absorb a column
        (MPO) into the boundary MPS, compress it to ``chi``, pull the norm out
        into a long-double accumulator, repeat; close with the last column."""
        ncols = self.shape[1]
        columns = []
        norm = np.longdouble(1)
        for col in range(ncols - 1):
            if col == 0:
                todo = column_to_mpo(self, 0)
            else:
                todo = od.contract_mps_mpo(boundary, column_to_mpo(self, col))
            if compression_type == "svd":
                boundary = od.svd_compress_mps(todo, chi, normalise=False, threshold=tolerance)
                centre = 0
                nrm = boundary.norm(canonical_form="right")
            elif compression_type == "variational":
                boundary = todo.variational_compress(chi, max_iter=max_iter, tolerance=tolerance)
                centre = -1
                nrm = boundary.norm(canonical_form="left")
            if nrm == 0.0:
                return 0.0
            boundary[centre].data = boundary[centre].data / nrm
            norm *= nrm
            if return_all_columns:
                snap = boundary.copy()
                snap[0].data *= norm
                columns.append(snap)
            if col == until_column:
                if return_all_columns:
                    return columns
                if compression_type == "svd":
                    boundary[0].data = np.asarray(boundary[0].data).astype(np.longdouble)
                boundary[centre].data *= norm
                return boundary
        closing = column_to_mpo(self, ncols - 1)
        full = od.inner_product_mps(boundary, closing, return_whole_tensor=True, complex_conjugate_bra=False) * norm
        if return_all_columns:
            columns.append(full)
            return columns
        return full

    def col_to_1D_array(self, col):
        return od.OneDimensionalTensorNetwork(self[:, col].copy(), left_label=self.up_label,
                                              right_label=self.down_label)

    def fliplr(self):
        mirror = self.copy()
        mirror.data = np.fliplr(mirror.data)
        mirror.right_label, mirror.left_label = self.left_label, self.right_label
        return mirror


class SquareLatticePEPS(SquareLatticeTensorNetwork):
    """PEPS: one physical index per site (square_lattice.py:234-281)."""

    def __init__(self, tensors, up_label="up", right_label="right", down_label="down", left_label="left",
                 phys_label="phys", copy_data=True):
        SquareLatticeTensorNetwork.__init__(self, tensors, up_label, right_label, down_label, left_label,
                                            copy_data=copy_data)
        self.phys_label = phys_label

    def copy(self):
        return SquareLatticePEPS(self.data, phys_label=self.phys_label, copy_data=True, **self._labels_kw())

    def outer_product(self, physin_label="physin", physout_label="physout"):
        return _outer(self, self, False, physin_label, physout_label, first_is_in=True)

    density_operator = outer_product


def _outer(peps1, peps2, conj2, physin_label, physout_label, first_is_in):
    """Site-wise tensor product with the virtual bonds fused (:251-281, :306-342)."""
    if peps1.shape != peps2.shape:
        raise ValueError("Peps input do not have same dimension.")
    virt = [peps1.left_label, peps1.right_label, peps1.up_label, peps1.down_label]
    rows = []
    for r in range(peps1.shape[0]):
        row = []
        for c in range(peps1.shape[1]):
            other = tsr.conjugate(peps2[r, c]) if conj2 else peps2[r, c]
            t = tsr.contract(peps1[r, c], other, [], [])
            names = (physin_label, physout_label) if first_is_in else (physout_label, physin_label)
            t.labels[t.labels.index(peps1.phys_label)] = names[0]
            t.labels[t.labels.index(peps2.phys_label)] = names[1]
            t.consolidate_indices(labels=virt)
            row.append(t)
        rows.append(row)
    return SquareLatticePEPO(rows, physin_label=physin_label, physout_label=physout_label, **peps1._labels_kw())


def inner_product_peps(peps_ket, peps_bra, exact_contract="True", complex_conjugate_bra=True,
                       compression_type="svd", chi=2, max_iter=10, tolerance=1e-14, contract_virtual=True):
    """<bra|ket> (square_lattice.py:283-304): double-layer network with the
    bond pairs fused by consolidate_indices, then exact or boundary-MPS
    contraction.  Like the reference, the physical label is the literal "phys",
    the bra is always conjugated, and the default ``exact_contract`` is the
    (truthy) string "True"."""
    rows = []
    for i in range(peps_ket.shape[0]):
        row = []
        for j in range(peps_ket.shape[1]):
            t = tsr.conjugate(peps_bra[i, j])["phys"] * peps_ket[i, j]["phys"]
            t.consolidate_indices()
            row.append(t)
        rows.append(row)
    ip = SquareLatticeTensorNetwork(rows)
    if not contract_virtual:
        return ip
    if exact_contract:
        return ip.exact_contract()
    return ip.mps_contract(chi, compression_type=compression_type, max_iter=max_iter, tolerance=tolerance)


def outer_product_peps(peps1, peps2, physin_label="physin", physout_label="physout"):
    """|peps1><peps2| as a PEPO (square_lattice.py:306-342)."""
    return _outer(peps1, peps2, True, physin_label, physout_label, first_is_in=False)


class SquareLatticePEPO(SquareLatticeTensorNetwork):
    """PEPO: physin/physout per site (square_lattice.py:344-375)."""

    def __init__(self, tensors, up_label="up", right_label="right", down_label="down", left_label="left",
                 physin_label="physin", physout_label="physout", copy_data=True):
        SquareLatticeTensorNetwork.__init__(self, tensors, up_label, right_label, down_label, left_label,
                                            copy_data=copy_data)
        self.physin_label = physin_label
        self.physout_label = physout_label

    def copy(self):
        return SquareLatticePEPO(self.data, physin_label=self.physin_label, physout_label=self.physout_label,
                                 copy_data=True, **self._labels_kw())

    def trace(self):
        rows = []
        for i in range(self.shape[0]):
            row = []
            for j in range(self.shape[1]):
                t = self[i, j].copy()
                t.trace(self.physin_label, self.physout_label)
                row.append(t)
            rows.append(row)
        return SquareLatticeTensorNetwork(rows, **self._labels_kw())


def apply_pepo_to_peps(peps, pepo):
    """Site-wise application, bond pairs fused (square_lattice.py:377-392)."""
    rows = []
    for i in range(peps.shape[0]):
        row = []
        for j in range(peps.shape[1]):
            t = peps[i, j][peps.phys_label] * pepo[i, j][pepo.physin_label]
            t.replace_label(pepo.physout_label, peps.phys_label)
            t.consolidate_indices()
            row.append(t)
        rows.append(row)
    return SquareLatticePEPS(rows, phys_label=peps.phys_label, **peps._labels_kw())


def column_to_mpo(square_tn, col):
    """Column -> MPS (first / last column, the outward dummy bond dropped) or
    MPO with physin = left, physout = right (square_lattice.py:394-417)."""
    sites = square_tn[:, col].copy()
    last = square_tn.shape[1] - 1
    if col == 0 or col == last:
        inward, outward = ((square_tn.right_label, square_tn.left_label) if col == 0
                           else (square_tn.left_label, square_tn.right_label))
        mps = od.MatrixProductState(sites, square_tn.up_label, square_tn.down_label, inward)
        for x in mps.data:
            x.remove_all_dummy_indices(outward)
        return mps
    return od.MatrixProductOperator(sites, square_tn.up_label, square_tn.down_label, square_tn.right_label,
                                    square_tn.left_label)
