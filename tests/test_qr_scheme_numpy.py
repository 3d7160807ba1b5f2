"""NumPy restatement of the many-column QR scheme of csrc/qr.cu (block Gram-Schmidt BCGS-PIP+ with a lagged,
Cholesky-free second pass) -- checks the numerical claims the CUDA path relies on, on the CPU:
  * one Pythagorean pass per 64-column block + one first-order second pass per 256-column group gives
    |Q^H Q - I| ~ eps and |Q R - A| ~ eps |A| for moderately conditioned inputs;
  * the first-order factor R2 = I + U, R2^-1 = I - U (U = striu(E) + diag(E)/2) of a Gram matrix I + E is
    exact to O(|E|^2);
  * the second-pass Gram matrix leaves the 1e-8 neighbourhood of I when cond(A) grows (the device flag that
    sends the CUDA path to its two-pass / Householder fallbacks).
The GPU parity tests of the real kernels are in test_cabi_gpu.py (test_qr, test_qr_many_columns)."""
import numpy as np

CB, GB = 64, 256


def first_order_factor(G):
    n = G.shape[0]
    E = G - np.eye(n)
    U = np.triu(E, 1) + np.diag(np.real(np.diag(E))) / 2
    return np.eye(n) + U, np.eye(n) - U, np.max(np.abs(np.triu(E)))


def bcgs_pip_lagged(A, cb=CB, gb=GB):
    m, n = A.shape
    k = min(m, n)
    Q = A[:, :k].astype(complex).copy()
    R = np.zeros((k, n), dtype=complex)
    emax = 0.0
    for g0 in range(0, k, gb):
        g1 = min(g0 + gb, k)
        for j0 in range(g0, g1, cb):                      # first pass, block by block
            j1 = min(j0 + cb, g1)
            S = Q[:, :j1].conj().T @ Q[:, j0:j1]          # [C; G0]
            C, G = S[:j0], S[j0:] - S[:j0].conj().T @ S[:j0]
            R1 = np.linalg.cholesky(G).conj().T           # upper
            Ri = np.linalg.inv(R1)
            Q[:, j0:j1] = Q[:, :j1] @ np.vstack([-C @ Ri, Ri])
            R[:j0, j0:j1] = C
            R[j0:j1, j0:j1] = R1
        S = Q[:, :g1].conj().T @ Q[:, g0:g1]              # lagged second pass over the group
        C, G = S[:g0], S[g0:] - S[:g0].conj().T @ S[:g0]
        R2, Ri2, e = first_order_factor(G)
        emax = max(emax, e)
        Q[:, g0:g1] = Q[:, :g1] @ np.vstack([-C @ Ri2, Ri2])
        Rt = R[g0:g1, g0:g1].copy()
        R[g0:g1, g0:g1] = R2 @ Rt
        R[:g0, g0:g1] += C @ Rt
    if n > k:
        R[:, k:] = Q.conj().T @ A[:, k:]
    return Q, R, emax


def _rand(rng, m, n):
    return rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))


def test_first_order_factor_is_second_order_accurate():
    rng = np.random.default_rng(0)
    for eps in (1e-13, 1e-10, 1e-8):
        H = _rand(rng, 96, 96)
        G = np.eye(96) + eps * (H + H.conj().T) / np.linalg.norm(H, 2)
        R, Ri, e = first_order_factor(G)
        assert e <= 2.1 * eps
        assert np.array_equal(np.tril(R, -1), np.zeros_like(R)) and np.all(np.diag(R).imag == 0)
        assert np.linalg.norm(R.conj().T @ R - G, 2) <= 4 * eps * eps + 4e-16
        assert np.linalg.norm(R @ Ri - np.eye(96), 2) <= 4 * eps * eps + 4e-16


def test_lagged_scheme_on_well_conditioned_inputs():
    rng = np.random.default_rng(1)
    for m, n in ((700, 600), (400, 330), (300, 520)):
        A = _rand(rng, m, n)
        Q, R, emax = bcgs_pip_lagged(A)
        k = min(m, n)
        assert emax < 1e-8                                   # the device flag would stay clear
        assert np.linalg.norm(Q.conj().T @ Q - np.eye(k)) < 1e-13 * k
        assert np.linalg.norm(Q @ R - A) / np.linalg.norm(A) < 1e-14
        assert np.array_equal(np.tril(R[:, :k], -1), np.zeros((k, k)))
        rr = np.linalg.qr(A, mode="r")
        assert np.allclose(np.abs(np.diag(R)), np.abs(np.diag(rr))[:k], rtol=1e-11, atol=0)


def test_second_pass_gram_leaves_identity_neighbourhood_with_condition_number():
    rng = np.random.default_rng(2)
    u, _ = np.linalg.qr(_rand(rng, 500, 320))
    v, _ = np.linalg.qr(_rand(rng, 320, 320))
    seen = []
    for decades in (1, 3, 6):
        A = (u * np.logspace(0, -decades, 320)[None, :]) @ v.conj().T
        Q, R, emax = bcgs_pip_lagged(A)
        seen.append(emax)
        if emax < 1e-8:                                      # accepted by the lagged path: must be accurate
            assert np.linalg.norm(Q.conj().T @ Q - np.eye(320)) < 1e-12 * 320
            assert np.linalg.norm(Q @ R - A) / np.linalg.norm(A) < 1e-13
    assert seen[0] < 1e-10 and seen[-1] > 1e-8               # cond 1e6: the flag trips, fallback takes over
