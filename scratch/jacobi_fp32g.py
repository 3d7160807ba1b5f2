"""Sweep counts of the block Jacobi when the rotation phase (G update + angles) runs in FP32 and only W (renormalised
rotations) and the Gram / apply products stay FP64."""
import numpy as np, sys, time
sys.path.insert(0, 'tests')
from test_jacobi_round_numpy import steps, rr_pair, JB, JP
SC, SD = steps(False), steps(True)
f32 = np.float32
def make_rot32(al, be, g, tol2):
    # normalised FP32 formulation
    if not (al > 0 and be > 0): return f32(1), np.complex64(0), 0
    inv = f32(1) / max(al, be)
    a_, b_, g_ = al * inv, be * inv, np.complex64(g * inv)
    ag2 = f32(g_.real * g_.real + g_.imag * g_.imag)
    cos2 = ag2 / (a_ * b_)
    if not (cos2 > f32(tol2)): return f32(1), np.complex64(0), 0
    st = 3 if cos2 > f32(1e-14) else 1
    tau = f32(0.5) * (b_ - a_)
    rh = f32(1) / np.sqrt(f32(tau * tau + ag2))
    c2 = f32(0.5) + f32(0.5) * abs(tau) * rh
    rc = f32(1) / np.sqrt(c2)
    return f32(c2 * rc), np.complex64(g_ * np.copysign(f32(0.5) * rh * rc, tau)), st
def make_rot64(al, be, g, tol2):
    ag2, ab = abs(g) ** 2, al * be
    if not (ab > 0.0 and ag2 > tol2 * ab): return 1.0, 0.0j, 0
    st = 3 if ag2 > 1e-14 * ab else 1
    tau = 0.5 * (be - al); rh = 1.0 / np.sqrt(tau * tau + ag2); c2 = 0.5 + 0.5 * abs(tau) * rh; rc = 1.0 / np.sqrt(c2)
    return c2 * rc, g * np.copysign(0.5 * rh * rc, tau), st
def inner(G, sts, tol2, fp32):
    W = np.eye(JP, dtype=complex); state = 0
    if fp32:
        G = (G / np.max(np.diag(G).real)).astype(np.complex64)
    for pairs in sts:
        J = np.eye(JP, dtype=np.complex64 if fp32 else complex); Jd = np.eye(JP, dtype=complex)
        for p, q in pairs:
            if fp32:
                c, sp, st = make_rot32(G[p, p].real, G[q, q].real, G[p, q], tol2)
                cd, sd = float(c), complex(sp); n2 = cd * cd + abs(sd) ** 2; d = n2 - 1.0; r = 1 - d / 2 + 3 * d * d / 8
                cd, sd = cd * r, sd * r
            else:
                c, sp, st = make_rot64(G[p, p].real, G[q, q].real, G[p, q], tol2); cd, sd = c, sp
            state |= st
            J[p, p], J[p, q], J[q, p], J[q, q] = c, sp, -np.conj(sp), c
            Jd[p, p], Jd[p, q], Jd[q, p], Jd[q, q] = cd, sd, -np.conj(sd), cd
        G = J.conj().T @ G @ J; W = W @ Jd
        if fp32: G = G.astype(np.complex64)
    return W, state
def run(Xt, fp32, maxsweeps=30):
    n, L = Xt.shape; Xt = Xt / np.linalg.norm(Xt); tol2 = L * 2.22e-16**2
    nblk = n // JB
    for sw in range(maxsweeps):
        state = 0
        for r in range(-1, nblk - 1):
            for p in range(nblk // 2):
                I, J = rr_pair(nblk, max(r, 0), p); I, J = min(I, J), max(I, J)
                idx = np.r_[I*JB:I*JB+JB, J*JB:J*JB+JB]
                P = Xt[idx]; G = P.conj() @ P.T
                W, st = inner(G, SD if r < 0 else SC, tol2, fp32); state |= st
                Xt[idx] = W.T @ P
        if not (state & 2): return sw + 1, Xt
    return maxsweeps, Xt
rng = np.random.default_rng(0)
for n, kind in ((128, 'rand'), (256, 'rand'), (256, 'graded'), (256, 'lowrank')):
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    if kind == 'graded': A = A * np.logspace(0, -12, n)[None, :]
    if kind == 'lowrank': A = A[:, :n // 3] @ (rng.standard_normal((n // 3, n)) + 1j * rng.standard_normal((n // 3, n)))
    sref = np.linalg.svd(A, compute_uv=False)
    Q, R = np.linalg.qr(A)
    for fp32 in (False, True):
        t = time.time(); sw, Y = run(np.conj(R).copy(), fp32)
        s = np.sort(np.linalg.norm(Y, axis=1))[::-1] * np.linalg.norm(R)
        keep = sref > 1e-13 * sref[0]
        rel = np.max(np.abs(s - sref)[keep] / sref[keep])
        Yn = Y[np.linalg.norm(Y, axis=1) > 1e-13 * np.linalg.norm(Y, axis=1).max()]
        Yn = Yn / np.linalg.norm(Yn, axis=1)[:, None]; orth = np.abs(Yn.conj() @ Yn.T - np.eye(len(Yn))).max()
        print(n, kind, 'fp32' if fp32 else 'fp64', 'sweeps', sw, 'sv relerr %.1e orth %.1e' % (rel, orth), '%.0fs' % (time.time() - t), flush=True)
