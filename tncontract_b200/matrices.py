"""Small operator constants (host ``np.matrix`` objects, as in the reference's
matrices.py:17-80; no device work -- they are uploaded when wrapped in a Tensor)."""
import numpy as np


def _m(rows):
    return np.matrix(rows)


def sigmap():
    return _m([[0., 1.], [0., 0.]])


def sigmam():
    return _m([[0., 0.], [1., 0.]])


def sigmax():
    return sigmam() + sigmap()


def sigmay():
    return -1j * sigmap() + 1j * sigmam()


def sigmaz():
    return _m([[1., 0.], [0., -1.]])


def destroy(dim):
    """Lowering operator of a ``dim``-level system."""
    return _m(np.diag(np.sqrt(range(1, dim)), 1))


def create(dim):
    return destroy(dim).getH()


def identity(dim):
    return _m(np.identity(dim))


def basis(dim, i):
    """``dim`` x 1 unit column vector e_i."""
    v = np.zeros(dim)
    v[i] = 1.0
    return _m(v).T
