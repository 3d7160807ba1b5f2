"""NumPy restatement of the rotation phase of jacobi_round_kernel (csrc/svd.cu): the parallel ordering of a
cross-pair / in-block round, the rotation formula, the one-step look-ahead (the three entries of G^(s+1) a
rotation needs, from three 2x2 blocks of G^(s) and the two rotations of step s) and the Hermitian update.
Checks, on the CPU, the identities the kernel relies on; the kernel itself is tested in test_cabi_gpu.py."""
import numpy as np

JB, JP = 16, 32


def rr_pair(n, r, p):
    if p == 0:
        return n - 1, r
    return (r + p) % (n - 1), (r - p + n - 1) % (n - 1)


def steps(diag):
    out = []
    for st in range(JB - 1 if diag else JB):
        pairs = []
        for pr in range(JP // 2):
            if diag:
                x, y = rr_pair(JB, st, pr & (JB // 2 - 1))
                if pr >= JB // 2:
                    x, y = x + JB, y + JB
            else:
                x, y = pr, JB + ((pr + st) & (JB - 1))
            pairs.append((min(x, y), max(x, y)))
        out.append(pairs)
    return out


def make_rot(alpha, beta, gam, tol2=0.0):
    ag2, ab = abs(gam) ** 2, alpha * beta
    if not (ab > 0.0 and ag2 > tol2 * ab):
        return 1.0, 0.0j
    tau = 0.5 * (beta - alpha)
    rh = 1.0 / np.sqrt(tau * tau + ag2)
    c2 = 0.5 + 0.5 * abs(tau) * rh
    rc = 1.0 / np.sqrt(c2)
    return c2 * rc, gam * np.copysign(0.5 * rh * rc, tau)


def rot_matrix(pairs, rots):
    J = np.eye(JP, dtype=complex)
    for (p, q), (c, sp) in zip(pairs, rots):
        J[p, p], J[p, q], J[q, p], J[q, q] = c, sp, -np.conj(sp), c
    return J


def rotated_entry(B, Ra, Rb, r, k):
    """entry (r, k) of J_a^H B J_b for a 2x2 block B"""
    (ca, spa), (cb, spb) = Ra, Rb
    Ja = np.array([[ca, spa], [-np.conj(spa), ca]])
    Jb = np.array([[cb, spb], [-np.conj(spb), cb]])
    return (Ja.conj().T @ B @ Jb)[r, k]


def test_orderings_cover_every_pair_once():
    cross = {pq for st in steps(False) for pq in st}
    assert cross == {(a, JB + b) for a in range(JB) for b in range(JB)}
    diag = [pq for st in steps(True) for pq in st]
    assert len(diag) == len(set(diag)) == 2 * (JB * (JB - 1) // 2)
    for st in steps(False) + steps(True):                     # 16 disjoint pairs covering all 32 columns
        assert sorted(i for pq in st for i in pq) == list(range(JP))


def test_rotation_is_unitary_and_annihilates_the_pivot():
    rng = np.random.default_rng(0)
    for _ in range(50):
        x = rng.standard_normal((40, 2)) + 1j * rng.standard_normal((40, 2))
        x[:, 1] *= 10.0 ** rng.uniform(-6, 6)
        G = x.conj().T @ x
        c, sp = make_rot(G[0, 0].real, G[1, 1].real, G[0, 1])
        J = np.array([[c, sp], [-np.conj(sp), c]])
        assert abs(c * c + abs(sp) ** 2 - 1) < 1e-15
        Gn = J.conj().T @ G @ J
        assert abs(Gn[0, 1]) <= 1e-15 * np.sqrt(G[0, 0].real * G[1, 1].real) + 1e-300


def test_lookahead_entries_equal_the_updated_gram_matrix():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((200, JP)) + 1j * rng.standard_normal((200, JP))
    for diag in (False, True):
        G = X.conj().T @ X
        W = np.eye(JP, dtype=complex)
        sts = steps(diag)
        rots = [make_rot(G[p, p].real, G[q, q].real, G[p, q]) for p, q in sts[0]]
        for s, pairs in enumerate(sts):
            pos = {}
            for a, (p, q) in enumerate(pairs):
                pos[p], pos[q] = (a, 0), (a, 1)
            nxt = None
            if s + 1 < len(sts):                              # what the rotation warp computes from G^(s)
                nxt = []
                for p2, q2 in sts[s + 1]:
                    (a1, r1), (a2, r2) = pos[p2], pos[q2]
                    i1, i2 = list(pairs[a1]), list(pairs[a2])
                    al = rotated_entry(G[np.ix_(i1, i1)], rots[a1], rots[a1], r1, r1).real
                    be = rotated_entry(G[np.ix_(i2, i2)], rots[a2], rots[a2], r2, r2).real
                    ga = rotated_entry(G[np.ix_(i1, i2)], rots[a1], rots[a2], r1, r2)
                    nxt.append((al, be, ga))
            J = rot_matrix(pairs, rots)
            G = J.conj().T @ G @ J                            # what the update warps compute
            W = W @ J
            if nxt is not None:
                scale = np.max(np.abs(G))
                for (al, be, ga), (p2, q2) in zip(nxt, sts[s + 1]):
                    assert abs(al - G[p2, p2].real) < 1e-13 * scale
                    assert abs(be - G[q2, q2].real) < 1e-13 * scale
                    assert abs(ga - G[p2, q2]) < 1e-13 * scale
                rots = [make_rot(al, be, ga) for al, be, ga in nxt]
        assert np.linalg.norm(W.conj().T @ W - np.eye(JP)) < 1e-13
        assert np.linalg.norm(G - G.conj().T) < 1e-12 * np.max(np.abs(G))
        Y = X @ W                                             # the apply: columns rotated by the accumulated W
        assert np.linalg.norm(Y.conj().T @ Y - G) < 1e-12 * np.max(np.abs(G))


def test_three_multiplication_complex_product():
    rng = np.random.default_rng(2)
    a = rng.standard_normal((32, 64)) + 1j * rng.standard_normal((32, 64))
    b = rng.standard_normal((64, 48)) + 1j * rng.standard_normal((64, 48))
    p1, p2 = a.real @ b.real, a.imag @ b.imag
    p3 = (a.real + a.imag) @ (b.real + b.imag)
    c = (p1 - p2) + 1j * (p3 - p1 - p2)
    assert np.linalg.norm(c - a @ b) <= 1e-14 * np.linalg.norm(a, 2) * np.linalg.norm(b, 2)
    # Gram form used by the round kernel: conj(a)^T a' with P3 = (ar + ai)(br - bi)
    g1, g2 = a.real.T @ a.real, a.imag.T @ a.imag
    g3 = (a.real + a.imag).T @ (a.real - a.imag)
    g = (g1 + g2) + 1j * (g1 - g2 - g3)
    assert np.linalg.norm(g - a.conj().T @ a) <= 1e-14 * np.linalg.norm(a, 2) ** 2

