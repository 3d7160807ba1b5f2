"""Sweep counts when every cross round also re-does the in-block rotations (31 steps per round instead of 16)."""
import numpy as np, sys, time
sys.path.insert(0, 'scratch'); sys.path.insert(0, 'tests')
from test_jacobi_round_numpy import steps, make_rot, rot_matrix, rr_pair, JB, JP
SC, SD = steps(False), steps(True)
def inner(G, sts, tol2):
    W = np.eye(JP, dtype=complex)
    for pairs in sts:
        rots = [make_rot(G[p, p].real, G[q, q].real, G[p, q], tol2) for p, q in pairs]
        J = rot_matrix(pairs, rots)
        G = J.conj().T @ G @ J; W = W @ J
    return W
def run(Xt, mode, maxsweeps=30):
    n, L = Xt.shape; Xt = Xt / np.linalg.norm(Xt); tol2 = L * 2.22e-16**2
    nblk = n // JB; hist = []
    for sw in range(maxsweeps):
        mx = 0.0
        for r in range(-1, nblk - 1):
            for p in range(nblk // 2):
                I, J = rr_pair(nblk, max(r, 0), p); I, J = min(I, J), max(I, J)
                idx = np.r_[I*JB:I*JB+JB, J*JB:J*JB+JB]
                P = Xt[idx]; G = P.conj() @ P.T
                d = np.sqrt(np.diag(G).real); C = np.abs(G) / np.outer(d, d); np.fill_diagonal(C, 0)
                if r < 0: C[:JB, JB:] = 0; C[JB:, :JB] = 0
                elif mode == 'cross': C[:JB, :JB] = 0; C[JB:, JB:] = 0
                mx = max(mx, C.max())
                if r < 0: sts = SD
                elif mode == 'cross': sts = SC
                else: sts = SC + SD          # cross then in-block again
                W = inner(G, sts, tol2)
                Xt[idx] = W.T @ P
        hist.append(mx)
        if mx < 1e-7: return sw + 1, hist
    return maxsweeps, hist
rng = np.random.default_rng(0)
for n in (256, 512):
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Q, R = np.linalg.qr(A)
    for mode in ('cross', 'cross+diag'):
        t = time.time(); sw, hist = run(np.conj(R).copy(), mode)
        print(n, mode, 'sweeps', sw, ['%.0e' % h for h in hist], '%.0fs' % (time.time() - t), flush=True)
