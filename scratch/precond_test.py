import numpy as np, sys
sys.path.insert(0,'scratch')
from jacobi_emul import block_jacobi
rng=np.random.default_rng(1)
def run(name, X):
    Xt,Vt,sw,hist=block_jacobi(X)
    print(name,'sweeps',sw,['%.0e'%h for h in hist])
for (m,n,kind) in [(192,192,'rand'),(128,192,'rand'),(192,192,'graded')]:
    A=rng.standard_normal((m,n))+1j*rng.standard_normal((m,n))
    if kind=='graded': A=A*np.logspace(0,-6,n)[None,:]
    wide=m<n
    Q,R=np.linalg.qr(A.conj().T if wide else A)
    print(m,n,kind)
    run(' R^H        ', np.conj(R))            # rows of Xt = columns of R^H  => Xt = conj(R)
    run(' R          ', R.T.copy())            # columns of R
    Q2,R2=np.linalg.qr(R.conj().T)             # R^H = Q2 R2
    run(' R2^H       ', np.conj(R2))
    # sorted by column norm (descending) variant on R^H
    Xt=np.conj(R); order=np.argsort(-np.linalg.norm(Xt,axis=1)); run(' R^H sorted ', Xt[order])
    # pivoted QR emulate: sort columns of A by norm first
    import scipy.linalg as sl
    Qp,Rp,P=sl.qr(A.conj().T if wide else A, mode='economic', pivoting=True)
    run(' Rp^H (piv) ', np.conj(Rp))
    Q3,R3=np.linalg.qr(Rp.conj().T)
    run(' (Rp^H)->R3^H', np.conj(R3))
