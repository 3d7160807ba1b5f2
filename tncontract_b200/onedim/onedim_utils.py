"""Constructors and per-site observables for MPS / MPO objects
(reference: onedim/onedim_utils.py; lines cited per function).  Initial data
is produced on the host and uploaded once; everything after that runs through
the device Tensor API."""
import numpy as np

from .. import tensor as tsr
from . import onedim_core as core

__all__ = ['init_mps_random', 'init_mps_allzero', 'init_mps_logical',
           'onebody_sum_mpo', 'expvals_mps', 'ptrace_mps']


def _per_site(value, n):
    return list(value) if np.iterable(value) else [value] * n


def init_mps_random(nsites, physdim, bonddim=1, left_label='left', right_label='right', phys_label='phys'):
    """onedim_utils.py:23-60: U[0,1) site tensors, each rescaled so that its
    largest singular value (as a (phys,left) x right matrix) is 1.  The
    reference multiplies U S V back together; here S/s0 is folded into V as a
    row scaling (tnb_diag_scale) followed by one GEMM."""
    physdim = _per_site(physdim, nsites)
    bonds = [1] + _per_site(bonddim, nsites - 1) + [1]
    sites = []
    for i in range(nsites):
        raw = tsr.Tensor(np.random.rand(physdim[i], bonds[i], bonds[i + 1]), [phys_label, left_label, right_label])
        U, S, V = tsr.tensor_svd(raw, [phys_label, left_label])
        S.data = S.data / S.data[0, 0]
        site = U["svd_in",] * S["svd_out",]
        sites.append(site["svd_in",] * V["svd_out",])
    return core.MatrixProductState(sites, left_label=left_label, right_label=right_label, phys_label=phys_label)


def _product_state(levels, physdim, left_label, right_label, phys_label):
    sites = []
    for lvl, d in zip(levels, physdim):
        v = np.zeros((d, 1, 1))
        v[lvl, 0, 0] = 1.0
        sites.append(tsr.Tensor(v, [phys_label, left_label, right_label]))
    return core.MatrixProductState(sites, left_label=left_label, right_label=right_label, phys_label=phys_label)


def init_mps_allzero(nsites, physdim, left_label='left', right_label='right', phys_label='phys'):
    """|00...0> (onedim_utils.py:63-88)."""
    physdim = _per_site(physdim, nsites)
    return _product_state([0] * nsites, physdim, left_label, right_label, phys_label)


def init_mps_logical(nsites, basis_state, physdim, left_label='left', right_label='right', phys_label='phys'):
    """|ij...l> with site n in level ``basis_state[n]`` (onedim_utils.py:91-119)."""
    physdim = _per_site(physdim, nsites)
    return _product_state([basis_state[j] for j in range(nsites)], physdim, left_label, right_label, phys_label)


def onebody_sum_mpo(terms, output_label=None):
    """sum_i O_i as a bond-dimension-2 MPO (onedim_utils.py:122-172;
    Sanchez-Burillo et al., PRL 113, 263604 (2014), SM eqs. (3)-(4))."""
    sites = []
    last = len(terms) - 1
    for i, term in enumerate(terms):
        if output_label is not None:
            term = term.copy()
            term.move_index(output_label, 0)
        op = np.asarray(term.data if isinstance(term, tsr.Tensor) else term, dtype=complex)
        eye = np.identity(op.shape[0], dtype=complex)[:, :op.shape[1]] if op.shape[0] == op.shape[1] else \
            (np.arange(op.shape[0])[:, None] == np.arange(op.shape[1])[None, :]).astype(complex)
        if i == 0:
            B = np.stack([op, eye], axis=-1)
            sites.append(tsr.Tensor(B, ['physout', 'physin', 'right']))
        elif i == last:
            B = np.stack([eye, op], axis=-1)
            sites.append(tsr.Tensor(B, ['physout', 'physin', 'left']))
        else:
            B = np.zeros(op.shape + (2, 2), dtype=complex)
            B[:, :, 0, 0] = eye
            B[:, :, 1, 0] = op
            B[:, :, 1, 1] = eye
            sites.append(tsr.Tensor(B, ['physout', 'physin', 'left', 'right']))
    return core.MatrixProductOperator(sites, left_label='left', right_label='right', physin_label='physin',
                                      physout_label='physout')


def _prepare_centre(mps, canonised):
    """Bring the orthogonality centre to site 0 (of the possibly reversed chain)."""
    if canonised == 'left':
        mps.reverse()
    elif canonised != 'right':
        mps.right_canonise()


def expvals_mps(mps, oplist=[], sites=None, output_label=None, canonised=None):
    """<op_i> for each requested site (onedim_utils.py:175-255); the centre of
    orthogonality is walked along the chain so every value is a one-site
    contraction."""
    if sites is None:
        sites = range(len(mps))
    if not np.iterable(sites):
        sites = [sites]
    n = len(sites)
    values = np.zeros(n, dtype=complex)
    ops = oplist if isinstance(oplist, list) else [oplist] * n
    if canonised == 'left':
        ops = ops[::-1]
    _prepare_centre(mps, canonised)
    centre = 0
    for i, site in enumerate(sites):
        mps.left_canonise(centre, site)
        centre = site
        op, A = ops[i], mps[site]
        if output_label is None:
            out_label, in_label = op.labels[0], op.labels[1]
        else:
            out_label = output_label
            in_label = [x for x in op.labels if x is not out_label][0]
        Ad = tsr.conjugate(A)
        e = tsr.contract(A, op, mps.phys_label, in_label)
        e = tsr.contract(Ad, e, mps.phys_label, out_label)
        e.contract_internal(mps.left_label, mps.left_label, index1=0, index2=1)
        e.contract_internal(mps.right_label, mps.right_label, index1=0, index2=1)
        values[i] = complex(np.asarray(e.data))
    if canonised == 'left':
        mps.reverse()
        values = values[::-1]
    return values


def ptrace_mps(mps, sites=None, canonised=None):
    """One-site reduced density matrices (onedim_utils.py:258-313)."""
    _prepare_centre(mps, canonised)
    if sites is None:
        sites = range(len(mps))
    if not np.iterable(sites):
        sites = [sites]
    rhos = []
    centre = 0
    for site in sites:
        mps.left_canonise(centre, site)
        centre = site
        A = mps[centre]
        Ad = tsr.conjugate(A)
        Ad.prime_label(mps.phys_label)
        virt = [mps.left_label, mps.right_label]
        rhos.append(tsr.contract(A, Ad, virt, virt))
    if canonised == 'left':
        mps.reverse()
        rhos = rhos[::-1]
    return rhos
