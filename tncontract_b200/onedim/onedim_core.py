"""One-dimensional tensor networks (MPS / MPO) on the GPU.

Same classes, method names, keyword arguments and label conventions as the
reference module (onedim/onedim_core.py; line numbers cited per routine).  The
sweeps are host-side loops; every numerical step is a libtnb call:

  QR sweep   : tnb_qr + tnb_tensordot (R absorbed into the next site)
  SVD sweep  : tnb_svd + tnb_truncation_count (chi / threshold applied on the
               device, one 16-byte read-back for the kept rank) + tnb_tensordot
               (V absorbed) + tnb_diag_scale (S absorbed; the reference does a
               dense k x k diagonal GEMM here, onedim_core.py:349)
  MPO apply  : tnb_mps_mpo_site (tensordot + consolidate_indices fused)
  ladders    : tnb_tensordot with operand permutations folded into the GEMM
"""
import numpy as np

from .. import devarray as dv
from .. import tensor as tsr
from ..devarray import DevArray
from ..label import unique_label

__all__ = ['MatrixProductState', 'MatrixProductStateCanonical',
           'MatrixProductOperator', 'OneDimensionalTensorNetwork',
           'check_canonical_form_mps',
           'contract_mps_mpo', 'contract_multi_index_tensor_with_one_dim_array',
           'contract_virtual_indices', 'frob_distance_squared',
           'inner_product_mps', 'ladder_contract', 'left_canonical_form_mps',
           'mps_complex_conjugate', 'reverse_mps', 'right_canonical_form_mps',
           'svd_compress_mps', 'variational_compress_mps', 'tensor_to_mpo',
           'tensor_to_mps',
           'right_canonical_to_canonical', 'left_canonical_to_canonical',
           'canonical_to_right_canonical', 'canonical_to_left_canonical',
           ]


def _object_array(items):
    arr = np.empty(len(items), dtype=object)
    for i, x in enumerate(items):
        arr[i] = x
    return arr


def _as_list(x):
    return x if isinstance(x, list) else [x]


def _gate_io(gate, gate_outputs, gate_inputs):
    """Default output / input labels of a gate tensor (onedim_core.py:705-716)."""
    if gate_outputs is None and gate_inputs is None:
        half = int(len(gate.labels) / 2)
        gate_outputs, gate_inputs = gate.labels[:half], gate.labels[half:]
    elif gate_outputs is None:
        gate_outputs = [x for x in gate.labels if x not in gate_inputs]
    elif gate_inputs is None:
        gate_inputs = [x for x in gate.labels if x not in gate_outputs]
    if len(gate_outputs) != len(gate_inputs):
        raise ValueError("len(gate_outputs) != len(gate_inputs)")
    return gate_outputs, gate_inputs


class OneDimensionalTensorNetwork:
    """Array of tensors joined left-to-right (onedim_core.py:31-183)."""

    def __init__(self, tensors, left_label="left", right_label="right"):
        self.left_label = left_label
        self.right_label = right_label
        self.data = _object_array([x.copy() for x in tensors])
        for x in self.data:  # open boundaries get dimension-1 virtual axes (:59-61)
            if left_label not in x.labels:
                x.add_dummy_index(left_label)
            if right_label not in x.labels:
                x.add_dummy_index(right_label)

    def __iter__(self):
        return self.data.__iter__()

    def __len__(self):
        return self.data.__len__()

    def __getitem__(self, key):
        return self.data.__getitem__(key)

    def __setitem__(self, key, value):
        self.data.__setitem__(key, value)

    def copy(self):
        return OneDimensionalTensorNetwork([x.copy() for x in self], self.left_label, self.right_label)

    def reverse(self):
        """Reverse the site order and swap the NAMES of the virtual labels (:84-88)."""
        self.data = self.data[::-1]
        self.left_label, self.right_label = self.right_label, self.left_label

    def complex_conjugate(self):
        for x in self.data:
            x.conjugate()

    def swap_gate(self, i, threshold=1e-15):
        """Swap the physical indices of sites i, i+1 by an SVD (:95-127)."""
        A, B = self[i], self[i + 1]
        virt = (self.left_label, self.right_label)
        A_phys = [l for l in A.labels if l not in virt]
        B_phys = [l for l in B.labels if l not in virt]
        A.prime_label(A_phys)
        t = tsr.contract(A, B, self.right_label, self.left_label)
        U, V, _ = tsr.truncated_svd(t, [self.left_label] + B_phys, chi=0, threshold=threshold,
                                    absorb_singular_values='both')
        U.replace_label('svd_in', self.right_label)
        self[i] = U
        V.unprime_label(A_phys)
        V.replace_label('svd_out', self.left_label)
        self[i + 1] = V

    def _rename_attrs(self, old, new, attrs):
        for a in attrs:
            cur = getattr(self, a)
            if cur in old:
                setattr(self, a, new[old.index(cur)])

    def replace_labels(self, old_labels, new_labels):
        old, new = _as_list(old_labels), _as_list(new_labels)
        for x in self.data:
            x.replace_label(old, new)
        self._rename_attrs(old, new, ("left_label", "right_label"))

    def standard_virtual_labels(self, suffix=""):
        self.replace_labels([self.left_label, self.right_label], ["left" + suffix, "right" + suffix])

    def unique_virtual_labels(self):
        self.replace_labels([self.left_label, self.right_label], [unique_label(), unique_label()])

    def leftdim(self, site):
        return self.data[site].index_dimension(self.left_label)

    def rightdim(self, site):
        return self.data[site].index_dimension(self.right_label)

    def bonddims(self):
        if self.nsites == 0:
            return []
        return [self.leftdim(0)] + [self.rightdim(i) for i in range(self.nsites)]

    @property
    def nsites(self):
        return len(self.data)

    @property
    def nsites_physical(self):
        return self.nsites


class _PhysLabelMixin:
    """replace_labels / standard_labels shared by the two MPS classes (:380-406, :930-956)."""

    def replace_labels(self, old_labels, new_labels):
        old, new = _as_list(old_labels), _as_list(new_labels)
        for x in self.data:
            x.replace_label(old, new)
        self._rename_attrs(old, new, ("left_label", "right_label", "phys_label"))

    def standard_labels(self, suffix=""):
        self.replace_labels([self.left_label, self.right_label, self.phys_label],
                            ["left" + suffix, "right" + suffix, "phys" + suffix])


class MatrixProductState(_PhysLabelMixin, OneDimensionalTensorNetwork):
    """Matrix product state (onedim_core.py:186-877)."""

    def __init__(self, tensors, left_label="left", right_label="right", phys_label="phys"):
        OneDimensionalTensorNetwork.__init__(self, tensors, left_label=left_label, right_label=right_label)
        self.phys_label = phys_label

    def __repr__(self):
        return ("MatrixProductState(tensors=%r, left_label=%r, right_label=%r,"
                "phys_label=%r)" % (self.data, self.left_label, self.right_label, self.phys_label))

    def __str__(self):
        return ("MatrixProductState object: " + "sites = " + str(len(self)) + ", left_label = " + self.left_label +
                ", right_label = " + self.right_label + ", phys_label = " + self.phys_label)

    def copy(self):
        return MatrixProductState([x.copy() for x in self], self.left_label, self.right_label, self.phys_label)

    # ---- canonisation sweeps ---------------------------------------------------------
    def _reset_to_zero_state(self):
        """Zero-norm input: all-zero product state with bond dimension 1 (:274-278)."""
        for k in range(len(self)):
            d = self[k].index_dimension(self.phys_label)
            self[k].data = DevArray.zeros((d, 1, 1))
            self[k].labels = [self.phys_label, self.left_label, self.right_label]

    def left_canonise(self, start=0, end=-1, chi=None, threshold=1e-14, normalise=False, qr_decomposition=False):
        """onedim_core.py:218-358.  Sites ``start`` .. ``end``-1 become isometries
        from (phys, left) to right; QR (no truncation) or SVD (chi / threshold
        relative to the largest singular value of each bond)."""
        N = len(self)
        if end == -1:
            end = N
        rows = [self.phys_label, self.left_label]
        scale = 1  # product of the largest singular values divided out so far (SVD branch)
        for i in range(start, end):
            if i == N - 1:
                # nothing to the right: what is left on this site is the norm
                nrm = self[i].data.norm()
                if nrm == 0.0:
                    self._reset_to_zero_state()
                elif normalise == True and start == 0:
                    self[i].data = self[i].data / nrm
                elif not qr_decomposition:
                    self[i].data = self[i].data * scale
                return
            tag = unique_label()
            if qr_decomposition:
                Q, R = tsr.tensor_qr(self[i], rows, qr_label=tag)
                Q.replace_label(tag + "in", self.right_label)
                self[i] = Q
                nxt = tsr.contract(R, self[i + 1], self.right_label, self.left_label)
                nxt.replace_label(tag + "out", self.left_label)
                self[i + 1] = nxt
                continue
            # U, s and P = diag(s) V in one call: the reference contracts V and then diag(s / s0) into
            # the next site (:347-349); that product is (U^H A) / s0, so V itself is never needed
            U, s, P = tsr._svd_parts(self[i], rows, tag, project=True)
            # s/s0 > threshold, then [:chi] (:333-339), evaluated on the device
            kept, s0, s_rel = dv.truncation(s, chi, threshold, relative=2)
            if s0 == 0.0:
                self._reset_to_zero_state()
                return
            scale = scale * s0
            U.data = U.data[:, :, 0:kept]
            P.data = P.data[0:kept]
            P.data *= 1.0 / s0
            U.replace_label(tag + "in", self.right_label)
            self[i] = U
            nxt = tsr.contract(P, self[i + 1], self.right_label, self.left_label)
            nxt.replace_label(tag + "out", self.left_label)
            self[i + 1] = nxt
            if i == end - 1:
                self[i + 1].data *= scale

    def right_canonise(self, start=0, end=-1, chi=None, threshold=1e-14, normalise=False, qr_decomposition=False):
        """Mirror image of left_canonise (:360-378)."""
        self.reverse()
        N = len(self)
        if end == -1:
            end = N
        self.left_canonise(start=N - end, end=N - start, chi=chi, threshold=threshold, normalise=normalise,
                           qr_decomposition=qr_decomposition)
        self.reverse()

    def check_canonical_form(self, threshold=1e-14, print_output=True):
        """onedim_core.py:408-461."""
        cc = mps_complex_conjugate(self)
        n = len(self)

        def isometry_defect(i, virt, virt_cc):
            I = tsr.contract(self[i], cc[i], [self.phys_label, virt], [cc.phys_label, virt_cc])
            m = np.asarray(I.data)
            return np.linalg.norm(m - np.identity(m.shape[0]))

        first_not_left = n - 1
        for i in range(n - 1):
            if isometry_defect(i, self.left_label, cc.left_label) > threshold:
                first_not_left = i
                break
        first_not_right = 0
        for i in range(n - 1, 0, -1):
            if isometry_defect(i, self.right_label, cc.right_label) > threshold:
                first_not_right = i
                break
        if print_output:
            if first_not_left == first_not_right:
                if first_not_left == n - 1:
                    unnorm = abs(self[-1].data.norm() - 1) > threshold
                    print("MPS in left canonical form (" + ("unnormalised" if unnorm else "normalised") + ")")
                elif first_not_left == 0:
                    unnorm = abs(self[0].data.norm() - 1) > threshold
                    print("MPS in right canonical form (" + ("unnormalised" if unnorm else "normalised") + ")")
                else:
                    print("MPS in mixed canonical form with orthogonality centre at site " + str(first_not_right))
            else:
                print("No tensors left canonised" if first_not_left == 0
                      else "Tensors left canonised up to site " + str(first_not_left))
                print("No tensors right canonised" if first_not_right == n - 1
                      else "Tensors right canonised up to site " + str(first_not_right))
        return (first_not_left, first_not_right)

    def svd_compress(self, chi=None, threshold=1e-15, normalise=False, reverse=False):
        """onedim_core.py:463-484: QR sweep to the right, truncating SVD sweep
        back; the result is right-canonical (left-canonical if ``reverse``)."""
        if reverse:
            self.reverse()
        self.left_canonise(normalise=False, qr_decomposition=True)
        nrm = self.norm(canonical_form="left")
        self[-1].data /= nrm
        self.right_canonise(chi=chi, threshold=threshold, normalise=False)
        if normalise == False:
            self[0].data *= nrm
        if reverse:
            self.reverse()

    def variational_compress(self, chi, max_iter=10, initial_guess=None, tolerance=1e-15, normalise=False):
        """onedim_core.py:486-631: alternating single-site optimisation against
        the uncompressed state, starting from svd_compress (or ``initial_guess``)."""
        if initial_guess == None:
            mps = self.copy()
            mps.svd_compress(chi=chi, reverse=True)
        else:
            mps = initial_guess
            mps.left_canonise(qr_decomposition=True)
        mps.replace_labels([mps.left_label, mps.right_label, mps.phys_label],
                           [unique_label(), unique_label(), unique_label()])
        env_tag = unique_label()
        left_envs = ladder_contract(mps, self, mps.phys_label, self.phys_label,
                                    return_intermediate_contractions=True, right_output_label=env_tag,
                                    complex_conjugate_array1=True)

        def sweep(trial, target, left_envs):
            """One right-to-left pass; ``trial`` enters right-canonical... returns
            the right environments (which become the next pass's left ones) and
            the site norms (:540-606)."""
            le = left_envs[0].labels[0][:-1]
            re, lq = unique_label(), unique_label()
            right_envs = []
            norms = [trial[-1].norm()]
            last = target.nsites - 1
            for i in range(last, 0, -1):
                upd = tsr.contract(target[i], left_envs[i - 1], target.left_label, le + "2")
                if i != last:
                    upd = tsr.contract(upd, renv, target.right_label, re + "2")
                    upd.replace_label(re + "1", trial.right_label)
                upd.replace_label([le + "1", target.phys_label], [trial.left_label, trial.phys_label])
                L, Q = tsr.tensor_lq(upd, trial.left_label, lq_label=lq)
                Q.replace_label(lq + "out", trial.left_label)
                L.replace_label(lq + "in", trial.right_label)
                trial[i] = Q
                trial[i - 1] = tsr.contract(trial[i - 1], L, trial.right_label, trial.left_label)
                norms.append(trial[i - 1].norm())
                if i == last:
                    renv = tsr.contract(tsr.conjugate(trial[i]), target[i], trial.phys_label, self.phys_label)
                    renv.remove_all_dummy_indices(labels=[trial.right_label, target.right_label])
                else:
                    renv.contract(tsr.conjugate(trial[i]), re + "1", trial.right_label)
                    renv.contract(target[i], [trial.phys_label, re + "2"], [self.phys_label, self.right_label])
                renv.replace_label([trial.left_label, target.left_label], [re + "1", re + "2"])
                right_envs.append(renv.copy())
                if i == 1:
                    upd = tsr.contract(target[0], renv, target.right_label, re + "2")
                    upd.replace_label([target.phys_label, re + "1"], [trial.phys_label, trial.right_label])
                    trial[0] = upd
            return right_envs, np.array(norms)

        for it in range(max_iter):
            left_envs, _ = sweep(mps, self, left_envs)
            mps.reverse()
            self.reverse()
            left_envs, norms = sweep(mps, self, left_envs)
            mps.reverse()
            self.reverse()
            if np.all(np.abs(norms[1:] - norms[:-1]) / norms[1:] < tolerance):
                mps.replace_labels([mps.left_label, mps.right_label, mps.phys_label],
                                   [self.left_label, self.right_label, self.phys_label])
                if normalise == True:
                    mps[-1].data /= mps.norm(canonical_form="left")
                return mps
            elif it == max_iter - 1:
                raise RuntimeError("variational_compress did not converge.")

    def physical_site(self, n):
        return n

    def physdim(self, site):
        return self.data[site].index_dimension(self.phys_label)

    def norm(self, canonical_form=False):
        """onedim_core.py:642-660."""
        if canonical_form == "left":
            return self[-1].data.norm()
        elif canonical_form == "right":
            return self[0].data.norm()
        return np.sqrt(inner_product_mps(self, self))

    # ---- gates / local observables (onedim_core.py:662-877) ------------------------------
    def apply_gate(self, gate, firstsite, gate_outputs=None, gate_inputs=None, chi=None, threshold=1e-15,
                   canonise='left'):
        gate_outputs, gate_inputs = _gate_io(gate, gate_outputs, gate_inputs)
        nsites = len(gate_inputs)
        t = contract_virtual_indices(self, firstsite, firstsite + nsites, periodic_boundaries=False)
        t = tsr.contract(t, gate, self.phys_label, gate_inputs)
        if canonise == 'right':
            phys_labels, ll, rl = gate_outputs[::-1], 'right', 'left'
        else:
            phys_labels, ll, rl = gate_outputs, 'left', 'right'
        mps = tensor_to_mps(t, phys_labels=phys_labels, mps_phys_label=self.phys_label, left_label=ll,
                            right_label=rl, chi=chi, threshold=threshold)
        if canonise == 'right':
            mps.reverse()
        self.data[firstsite:firstsite + nsites] = mps.data

    def _centre_on(self, first, last_plus_one, left_canonised_up_to, right_canonised_up_to):
        if left_canonised_up_to < first:
            self.left_canonise(left_canonised_up_to, first)
        if right_canonised_up_to > last_plus_one:
            self.right_canonise(last_plus_one, right_canonised_up_to)

    def expval(self, gate, firstsite, left_canonised_up_to=0, right_canonised_up_to=-1, gate_outputs=None,
               gate_inputs=None):
        gate_outputs, gate_inputs = _gate_io(gate, gate_outputs, gate_inputs)
        nsites = len(gate_inputs)
        if right_canonised_up_to == -1:
            right_canonised_up_to = len(self)
        self._centre_on(firstsite, firstsite + nsites, left_canonised_up_to, right_canonised_up_to)
        t = contract_virtual_indices(self, firstsite, firstsite + nsites, periodic_boundaries=False)
        td = tsr.conjugate(t)
        exp = tsr.contract(t, gate, self.phys_label, gate_inputs)
        exp = tsr.contract(td, exp, self.phys_label, gate_outputs)
        exp.tr(self.left_label, self.left_label, index1=0, index2=1)
        exp.tr(self.right_label, self.right_label, index1=0, index2=1)
        return exp

    def ptrace(self, firstsite, lastsite=None, left_canonised_up_to=0, right_canonised_up_to=-1):
        if right_canonised_up_to == -1:
            right_canonised_up_to = self.nsites
        if lastsite is None:
            lastsite = firstsite
        self._centre_on(firstsite, lastsite + 1, left_canonised_up_to, right_canonised_up_to)
        t = contract_virtual_indices(self, self.physical_site(firstsite), self.physical_site(lastsite) + 1,
                                     periodic_boundaries=False)
        return _density_from_block(t, self)


def _density_from_block(t, net):
    """|t><t| with the outer virtual indices traced (:868-877, :1317-1326)."""
    td = tsr.conjugate(t)
    for i, l in enumerate(t.labels):
        if l == net.phys_label:
            t.labels[i] = l + "_out" + str(i)
            td.labels[i] = l + "_in" + str(i)
    return t[net.left_label, net.right_label] * td[net.left_label, net.right_label]


class MatrixProductStateCanonical(_PhysLabelMixin, OneDimensionalTensorNetwork):
    """Vidal form: Lambda Gamma Lambda ... Gamma Lambda (onedim_core.py:880-1326)."""

    def __init__(self, tensors, left_label="left", right_label="right", phys_label="phys"):
        OneDimensionalTensorNetwork.__init__(self, tensors, left_label=left_label, right_label=right_label)
        self.phys_label = phys_label

    def __repr__(self):
        return ("MatrixProductStateCanonical(tensors=%r, left_label=%r,"
                "right_label=%r, phys_label=%r)" % (self.data, self.left_label, self.right_label, self.phys_label))

    def __str__(self):
        return ("MatrixProductStateCanonical object: " + "sites (incl. singular value sites)= " + str(len(self)) +
                ", left_label = " + self.left_label + ", right_label = " + self.right_label + ", phys_label = " +
                self.phys_label)

    def copy(self):
        return MatrixProductStateCanonical([x.copy() for x in self], self.left_label, self.right_label,
                                           self.phys_label)

    def physical_site(self, n):
        return 2 * n + 1

    def singular_site(self, n):
        return 2 * n

    def physdim(self, site):
        return self.data[self.physical_site(site)].index_dimension(self.phys_label)

    def singulardim(self, site):
        return self.data[self.singular_site(site)].index_dimension(self.left_label)

    def bonddims(self):
        return super(MatrixProductStateCanonical, self).bonddims()

    @property
    def nsites_physical(self):
        return int((self.nsites - 1) / 2)

    def norm(self, canonical_form=True):
        if canonical_form is True:
            return self[-1].data.norm() * self[0].data.norm()
        return np.sqrt(inner_product_mps(self, self))

    def check_canonical_form(self, threshold=1e-14, print_output=True):
        """onedim_core.py:1004-1065."""
        bad_left, bad_right, bad_norm = [], [], []
        edge = (0, self.nsites_physical - 1)

        def classify(i, block, virt, sink):
            I = tsr.contract(block, tsr.conjugate(block), [self.phys_label, virt], [self.phys_label, virt])
            m = np.asarray(I.data)
            if np.linalg.norm(m - np.identity(m.shape[0])) > threshold:
                flat = m.flatten()
                if i in edge and not len(flat[np.abs(flat) > threshold]) > 1:
                    bad_norm.append(i)
                else:
                    sink.append(i)

        for i in range(self.nsites_physical):
            p = self.physical_site(i)
            classify(i, self[p - 1][self.right_label,] * self[p][self.left_label,], self.left_label, bad_left)
        for i in range(self.nsites_physical):
            p = self.physical_site(i)
            classify(i, self[p][self.right_label,] * self[p + 1][self.left_label,], self.right_label, bad_right)
        if print_output:
            if not bad_left and not bad_right:
                print("MPS in canonical form (" + ("normalised" if not bad_norm else "unnormalised") + ")")
            else:
                print("Physical sites not left-canonical:")
                print(bad_left)
                print("Physical sites not right-canonical:")
                print(bad_right)
        return bad_left, bad_right, bad_norm

    def _inverse_lambdas(self, start, end):
        a, b = self[start].copy(), self[end].copy()
        a.inv()
        b.inv()
        return a, b

    def _store_split(self, start, U, S, V, S1_inv, S2_inv):
        """Write Gamma Lambda Gamma back after an SVD of the two-site block."""
        S.replace_label(["svd_out", "svd_in"], [self.left_label, self.right_label])
        self[start + 1] = S1_inv[self.right_label,] * U[self.left_label,]
        self[start + 2] = S
        self[start + 3] = V[self.right_label,] * S2_inv[self.left_label,]

    def compress_bond(self, singular_site, chi=None, threshold=1e-15):
        """onedim_core.py:1067-1095."""
        site = self.singular_site(singular_site)
        if self.singulardim(site) == 1:
            return
        start, end = site - 2, site + 2
        self[end - 1].prime_label(self.phys_label)
        t = contract_virtual_indices(self, start, end + 1, periodic_boundaries=False)
        S1_inv, S2_inv = self._inverse_lambdas(start, end)
        U, S, V = tsr.truncated_svd(t, [self.phys_label, self.left_label], chi=chi, threshold=threshold,
                                    absorb_singular_values=None)
        U.replace_label("svd_in", self.right_label)
        V.replace_label("svd_out", self.left_label)
        self._store_split(start, U, S, V, S1_inv, S2_inv)
        self[end - 1].unprime_label(self.phys_label)

    def compress_all(self, chi=None, threshold=1e-15, normalise=False):
        raise NotImplementedError

    def apply_gate(self, gate, firstsite, gate_outputs=None, gate_inputs=None, chi=None, threshold=1e-15):
        """onedim_core.py:1100-1186 (one- and two-site gates)."""
        gate_outputs, gate_inputs = _gate_io(gate, gate_outputs, gate_inputs)
        nsites = len(gate_inputs)
        if nsites > 2:
            raise NotImplementedError("gate acting on more than two sites.")
        start = self.physical_site(firstsite) - 1
        end = self.physical_site(firstsite + nsites - 1) + 1
        t = contract_virtual_indices(self, start, end + 1, periodic_boundaries=False)
        t = tsr.contract(t, gate, self.phys_label, gate_inputs)
        S1_inv, S2_inv = self._inverse_lambdas(start, end)
        if nsites == 1:
            t.replace_label([gate_outputs[0]], [self.phys_label])
            t = S1_inv[self.right_label,] * t[self.left_label,]
            self[start + 1] = t[self.right_label,] * S2_inv[self.left_label,]
        elif nsites == 2:
            U, S, V = tsr.truncated_svd(t, [gate_outputs[0], self.left_label], chi=chi, threshold=threshold,
                                        absorb_singular_values=None)
            U.replace_label(["svd_in", gate_outputs[0]], [self.right_label, self.phys_label])
            V.replace_label(["svd_out", gate_outputs[1]], [self.left_label, self.phys_label])
            self._store_split(start, U, S, V, S1_inv, S2_inv)

    def swap_gate(self, i, chi=None, threshold=1e-15):
        """onedim_core.py:1188-1232."""
        start = self.physical_site(i) - 1
        end = self.physical_site(i + 1) + 1
        self[start + 1].prime_label(self.phys_label)
        t = contract_virtual_indices(self, start, end + 1, periodic_boundaries=False)
        S1_inv, S2_inv = self._inverse_lambdas(start, end)
        U, S, V = tsr.truncated_svd(t, [self.left_label, self.phys_label], chi=chi, threshold=threshold,
                                    absorb_singular_values=None)
        V.unprime_label(self.phys_label)
        U.replace_label("svd_in", self.right_label)
        V.replace_label('svd_out', self.left_label)
        self._store_split(start, U, S, V, S1_inv, S2_inv)

    def expval(self, gate, firstsite, gate_outputs=None, gate_inputs=None):
        """onedim_core.py:1234-1290."""
        gate_outputs, gate_inputs = _gate_io(gate, gate_outputs, gate_inputs)
        start = self.physical_site(firstsite) - 1
        end = self.physical_site(firstsite + len(gate_inputs) - 1) + 1
        t = contract_virtual_indices(self, start, end + 1, periodic_boundaries=False)
        td = tsr.conjugate(t)
        exp = t[self.phys_label,] * gate[gate_inputs]
        exp = td[self.phys_label,] * exp[gate_outputs]
        exp.tr(self.left_label, self.left_label, index1=0, index2=1)
        exp.tr(self.right_label, self.right_label, index1=0, index2=1)
        return exp

    def ptrace(self, firstsite, lastsite=None):
        """onedim_core.py:1292-1326."""
        if lastsite is None:
            lastsite = firstsite
        start = self.physical_site(firstsite) - 1
        end = self.physical_site(lastsite) + 1
        t = contract_virtual_indices(self, start, end + 1, periodic_boundaries=False)
        return _density_from_block(t, self)


class MatrixProductOperator(OneDimensionalTensorNetwork):
    """Matrix product operator (onedim_core.py:1329-1367); like the reference
    it keeps the base-class ``copy``."""

    def __init__(self, tensors, left_label="left", right_label="right", physout_label="physout",
                 physin_label="physin"):
        OneDimensionalTensorNetwork.__init__(self, tensors, left_label, right_label)
        self.physout_label = physout_label
        self.physin_label = physin_label

    def __repr__(self):
        return ("MatrixProductOperator(tensors=%r, left_label=%r, right_label=%r, physout_label=%r, phsin_labe=%r)"
                % (self.data, self.left_label, self.right_label, self.physout_label, self.physin_label))

    def __str__(self):
        return ("MatrixProductOperator object: " + "sites = " + str(len(self)) + ", left_label = " +
                self.left_label + ", right_label = " + self.right_label + ", physout_label = " +
                self.physout_label + ", physin_label = " + self.physin_label)

    def physoutdim(self, site):
        return self.data[site].index_dimension(self.physout_label)

    def physindim(self, site):
        return self.data[site].index_dimension(self.physin_label)


# ---- free functions --------------------------------------------------------------------
def contract_multi_index_tensor_with_one_dim_array(tensor, array, label1, label2):
    """onedim_core.py:1370-1396."""
    tmp = unique_label()
    tensor.replace_label(label1, tmp)
    C = tsr.contract(tensor, array[0], tmp, label2, index_slice1=[0])
    for i in range(1, len(array)):
        C = tsr.contract(C, array[i], [array.right_label, tmp], [array.left_label, label2], index_slice1=[0, 1])
    C.contract_internal(array.right_label, array.left_label)
    tensor.replace_label(tmp, label1)
    return C


def contract_virtual_indices(array_1d, start=0, end=None, periodic_boundaries=True):
    """onedim_core.py:1399-1422."""
    C = array_1d[start].copy()
    for x in array_1d[start + 1:end]:
        C = tsr.contract(C, x, array_1d.right_label, array_1d.left_label)
    if periodic_boundaries:
        C.contract_internal(array_1d.right_label, array_1d.left_label)
    return C


def left_canonical_form_mps(orig_mps, chi=0, threshold=1e-14, normalise=False):
    mps = orig_mps.copy()
    mps.left_canonise(chi=chi, threshold=threshold, normalise=normalise)
    return mps


def right_canonical_form_mps(orig_mps, chi=0, threshold=1e-14, normalise=False):
    mps = orig_mps.copy()
    mps.right_canonise(chi=chi, threshold=threshold, normalise=normalise)
    return mps


def canonical_form_mps(orig_mps, chi=0, threshold=1e-14, normalise=False):
    mps = orig_mps.copy()
    mps.right_canonise(chi=chi, threshold=threshold, normalise=normalise)
    return right_canonical_to_canonical(mps, threshold=threshold)


def reverse_mps(orig_mps):
    mps = orig_mps.copy()
    mps.reverse()
    return mps


def check_canonical_form_mps(mps, threshold=1e-14, print_output=True):
    return mps.check_canonical_form(threshold=threshold, print_output=print_output)


def svd_compress_mps(orig_mps, chi, threshold=1e-15, normalise=False):
    """onedim_core.py:1468-1473: SVD sweep without a chi cut, then the
    truncating SVD sweep back (this is what twodim's boundary contraction uses)."""
    mps = left_canonical_form_mps(orig_mps, threshold=threshold, normalise=normalise)
    return right_canonical_form_mps(mps, chi=chi, threshold=threshold, normalise=normalise)


def variational_compress_mps(mps, chi, max_iter=10, initial_guess=None, tolerance=1e-15):
    return mps.variational_compress(chi, max_iter=max_iter, initial_guess=initial_guess, tolerance=tolerance)


def mps_complex_conjugate(mps):
    new_mps = mps.copy()
    for x in new_mps.data:
        x.conjugate()
    return new_mps


def ladder_contract(array1, array2, label1, label2, start=0, end=None, complex_conjugate_array1=False,
                    left_output_label="left", right_output_label="right", return_intermediate_contractions=False):
    """onedim_core.py:1491-1663: contract two chains rung by rung.  From the
    left boundary if the interval contains it, else from the right boundary if
    it contains that, else pairwise then together."""
    if end == None:
        end = min(array1.nsites, array2.nsites) - 1
    if end < start:
        raise ValueError("Badly defined interval (end before start).")
    a1, a2 = array1.copy(), array2.copy()
    if complex_conjugate_array1:
        a1.complex_conjugate()
    a1.unique_virtual_labels()
    a2.unique_virtual_labels()
    rung = unique_label()
    a1.replace_labels(label1, rung)
    a2.replace_labels(label2, rung)
    ro = [right_output_label + "1", right_output_label + "2"]
    lo = [left_output_label + "1", left_output_label + "2"]
    steps = []

    def snapshot(C, olds, news, twice=False):
        t = C.copy()
        t.replace_label(olds, news)
        t.remove_all_dummy_indices(labels=[x for x in t.labels if x not in news])
        if twice:
            t.remove_all_dummy_indices()
        return t

    if start == 0:
        olds, news = [a1.right_label, a2.right_label], ro
        for i in range(0, end + 1):
            if i == 0:
                C = tsr.contract(a1[0], a2[0], rung, rung)
            else:
                C.contract(a1[i], a1.right_label, a1.left_label)
                C.contract(a2[i], [a2.right_label, rung], [a2.left_label, rung])
            if return_intermediate_contractions:
                steps.append(snapshot(C, olds, news))
    elif end == a1.nsites - 1 and end == a2.nsites - 1:
        olds, news = [a1.left_label, a2.left_label], lo
        for i in range(end, start - 1, -1):
            if i == end:
                C = tsr.contract(a1[end], a2[end], rung, rung)
            else:
                C.contract(a1[i], a1.left_label, a1.right_label)
                C.contract(a2[i], [a2.left_label, rung], [a2.right_label, rung])
            if return_intermediate_contractions:
                steps.insert(0, snapshot(C, olds, news))
    else:
        olds = [a1.right_label, a2.right_label, a1.left_label, a2.left_label]
        news = ro + lo
        for i in range(start, end + 1):
            t = tsr.contract(a1[i], a2[i], rung, rung)
            if i == start:
                C = t
            else:
                C.contract(t, [a1.right_label, a2.right_label], [a1.left_label, a2.left_label])
            if return_intermediate_contractions:
                steps.append(snapshot(C, olds, news, twice=True))
    C.replace_label(olds, news)
    C.remove_all_dummy_indices()
    return steps if return_intermediate_contractions else C


def inner_product_mps(mps_bra, mps_ket, complex_conjugate_bra=True, return_whole_tensor=False):
    """onedim_core.py:1666-1683.  Returns a host 0-d array (the reference
    returns ``t.data``) unless the whole rank-0 Tensor is requested."""
    bra = canonical_to_left_canonical(mps_bra) if isinstance(mps_bra, MatrixProductStateCanonical) else mps_bra
    ket = canonical_to_left_canonical(mps_ket) if isinstance(mps_ket, MatrixProductStateCanonical) else mps_ket
    t = ladder_contract(bra, ket, mps_bra.phys_label, mps_ket.phys_label,
                        complex_conjugate_array1=complex_conjugate_bra)
    if return_whole_tensor:
        return t
    return np.asarray(t.data)


def frob_distance_squared(mps1, mps2):
    ip = inner_product_mps
    return ip(mps1, mps1) + ip(mps2, mps2) - 2 * np.real(ip(mps1, mps2))


def _fused_apply_ok(a, w, mps, mpo):
    """The fused kernel covers the standard case: one axis per label, and an MPO slice W[a, :, p, :] (wr x d
    elements) that fits its 48 KB of shared memory with at most 2^31 - 1 output rows (mps_mpo.cu); anything else
    takes the generic contract + consolidate_indices path."""
    if sorted(w.labels) == sorted([mpo.left_label, mpo.right_label, mpo.physout_label, mpo.physin_label]) and \
            len(set(w.labels)) == 4 and len(set(a.labels)) == 3 and mps.left_label in a.labels:
        wr, d = w.index_dimension(mpo.right_label), w.index_dimension(mpo.physin_label)
        rows = a.index_dimension(mps.left_label) * w.index_dimension(mpo.left_label) * w.index_dimension(mpo.physout_label)
        if wr * d * 16 > 48 * 1024 or rows > 2 ** 31 - 1:
            return False
    return (sorted(a.labels) == sorted([mps.phys_label, mps.left_label, mps.right_label]) and
            sorted(w.labels) == sorted([mpo.left_label, mpo.right_label, mpo.physout_label, mpo.physin_label]) and
            mps.left_label == mpo.left_label and mps.right_label == mpo.right_label and
            len(set(a.labels)) == 3 and len(set(w.labels)) == 4 and
            mpo.physout_label not in (mps.left_label, mps.right_label))


def contract_mps_mpo(mps, mpo):
    """onedim_core.py:1691-1708: per site contract phys with physin, then
    consolidate_indices() (alphabetical label order, bonds fused mps-major).
    Standard sites go through the fused tnb_mps_mpo_site kernel, which writes
    the consolidated layout directly."""
    if isinstance(mps, MatrixProductStateCanonical):
        raise NotImplementedError(("Function not implemented for" + "MatrixProductStateCanonical"))
    out = []
    for i in range(len(mps)):
        a, w = mps[i], mpo[i]
        if _fused_apply_ok(a, w, mps, mpo):
            A = a.data.transpose([a.labels.index(l) for l in (mps.phys_label, mps.left_label, mps.right_label)])
            W = w.data.transpose([w.labels.index(l) for l in (mpo.left_label, mpo.right_label, mpo.physout_label,
                                                              mpo.physin_label)])
            fused = dv.mps_mpo_site(A, W)  # [(l, wl), physout, (r, wr)]
            names = [mps.left_label, mpo.physout_label, mps.right_label]
            order = sorted(range(3), key=lambda k: names[k])
            t = tsr.Tensor._wrap(fused.transpose(order) if order != [0, 1, 2] else fused, [names[k] for k in order])
            if order != [0, 1, 2]:
                t.data = t.data.copy()  # consolidate_indices leaves a contiguous array
        else:
            t = tsr.contract(a, w, mps.phys_label, mpo.physin_label)
            t.consolidate_indices()
        out.append(t)
    return MatrixProductState._adopt(out, mps.left_label, mps.right_label, mpo.physout_label)


def _adopt(cls, tensors, left_label, right_label, phys_label):
    """Build an MPS around freshly produced tensors without the defensive
    per-site copy of the public constructor (the tensors are not aliased)."""
    self = cls.__new__(cls)
    self.left_label, self.right_label, self.phys_label = left_label, right_label, phys_label
    self.data = _object_array(tensors)
    for x in self.data:
        if left_label not in x.labels:
            x.add_dummy_index(left_label)
        if right_label not in x.labels:
            x.add_dummy_index(right_label)
    return self


MatrixProductState._adopt = classmethod(_adopt)


def tensor_to_mps(tensor, phys_labels=None, mps_phys_label='phys', left_label='left', right_label='right', chi=0,
                  threshold=1e-15):
    """onedim_core.py:1711-1761: peel sites off a big tensor by truncated SVDs."""
    if phys_labels is None:
        phys_labels = [x for x in tensor.labels if x not in [left_label, right_label]]
    nsites = len(phys_labels)
    V = tensor.copy()
    sites = []
    for k in range(nsites - 1):
        U, V, _ = tsr.truncated_svd(V, [left_label] * (left_label in V.labels) + [phys_labels[k]], chi=chi,
                                    threshold=threshold, absorb_singular_values='right')
        U.replace_label('svd_in', right_label)
        U.replace_label(phys_labels[k], mps_phys_label)
        sites.append(U)
        V.replace_label('svd_out', left_label)
    V.replace_label(phys_labels[nsites - 1], mps_phys_label)
    sites.append(V)
    return MatrixProductState(sites, phys_label=mps_phys_label, left_label=left_label, right_label=right_label)


def tensor_to_mpo(tensor, physout_labels=None, physin_labels=None, mpo_physout_label='physout',
                  mpo_physin_label='physin', left_label='left', right_label='right', chi=0, threshold=1e-15):
    """onedim_core.py:1764-1833."""
    phys_labels = [x for x in tensor.labels if x not in [left_label, right_label]]
    if physout_labels is None and physin_labels is None:
        half = int(len(phys_labels) / 2)
        physout_labels, physin_labels = phys_labels[:half], phys_labels[half:]
    elif physout_labels is None:
        physout_labels = [x for x in phys_labels if x not in physin_labels]
    elif physin_labels is None:
        physin_labels = [x for x in phys_labels if x not in physout_labels]
    nsites = len(physin_labels)
    if len(physout_labels) != nsites:
        raise ValueError("len(physout_labels) != len(physin_labels)")
    V = tensor.copy()
    sites = []
    for k in range(nsites - 1):
        U, V, _ = tsr.truncated_svd(V, [left_label] * (left_label in V.labels) + [physout_labels[k], physin_labels[k]],
                                    chi=chi, threshold=threshold)
        U.replace_label('svd_in', right_label)
        U.replace_label(physout_labels[k], mpo_physout_label)
        U.replace_label(physin_labels[k], mpo_physin_label)
        sites.append(U)
        V.replace_label('svd_out', left_label)
    V.replace_label(physout_labels[nsites - 1], mpo_physout_label)
    V.replace_label(physin_labels[nsites - 1], mpo_physin_label)
    sites.append(V)
    return MatrixProductOperator(sites, physout_label=mpo_physout_label, physin_label=mpo_physin_label,
                                 left_label=left_label, right_label=right_label)


def right_canonical_to_canonical(mps, threshold=1e-14):
    """onedim_core.py:1836-1884: right-canonical MPS -> Lambda/Gamma form."""
    N = mps.nsites
    S_prev = tsr.Tensor([[1.0]], labels=[mps.left_label, mps.right_label])
    S_prev_inv = S_prev.copy()
    B = mps[0]
    tensors = []
    tag = unique_label()
    for i in range(N):
        U, s, V = tsr._svd_parts(B, [mps.phys_label, mps.left_label], tag)
        kept, _, _ = dv.truncation(s, 0, threshold, relative=0)
        s = s[0:kept]
        S = tsr.Tensor._wrap(dv.diag_embed(s), [mps.left_label, mps.right_label])
        U.data = U.data[:, :, 0:kept]
        V.data = V.data[0:kept]
        U.replace_label(tag + "in", mps.right_label)
        V.replace_label(tag + "out", mps.left_label)
        G = S_prev_inv[mps.right_label,] * U[mps.left_label,]
        tensors.append(S_prev)
        tensors.append(G)
        if i == N - 1:
            tensors.append(S)
        else:
            V = S[mps.right_label,] * V[mps.left_label,]
            B = V[mps.right_label,] * mps[i + 1][mps.left_label,]
            S_prev = S.copy()
            S_prev_inv = tsr.Tensor._wrap(dv.diag_embed(s, mode=2), list(S.labels))
    return MatrixProductStateCanonical(tensors, left_label=mps.left_label, right_label=mps.right_label,
                                       phys_label=mps.phys_label)


def left_canonical_to_canonical(mps, threshold=1e-14):
    mpsr = reverse_mps(mps)
    mpsc = right_canonical_to_canonical(mpsr, threshold=threshold)
    mpsc.reverse()
    return mpsc


def canonical_to_right_canonical(mps):
    """onedim_core.py:1897-1911."""
    N = mps.nsites_physical
    tensors = []
    for i in range(N - 1):
        p = mps.physical_site(i)
        tensors.append(mps[p][mps.right_label,] * mps[p + 1][mps.left_label,])
    tensors.append(mps[mps.physical_site(N - 1)])
    tensors[0].data = tensors[0].data * mps[-1].data.norm() * mps[0].data.norm()
    return MatrixProductState(tensors, left_label=mps.left_label, right_label=mps.right_label,
                              phys_label=mps.phys_label)


def canonical_to_left_canonical(mps):
    """onedim_core.py:1914-1928."""
    N = mps.nsites_physical
    tensors = [mps[mps.physical_site(0)]]
    for i in range(1, N):
        p = mps.physical_site(i)
        tensors.append(mps[p - 1][mps.right_label,] * mps[p][mps.left_label,])
    tensors[-1].data = tensors[-1].data * mps[-1].data.norm() * mps[0].data.norm()
    return MatrixProductState(tensors, left_label=mps.left_label, right_label=mps.right_label,
                              phys_label=mps.phys_label)
