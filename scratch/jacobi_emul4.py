"""FP32 pre-pass experiment: run the block Jacobi sweeps in complex64 with V accumulated, re-orthonormalise V in
fp64, apply to X in fp64, count the remaining fp64 sweeps."""
import numpy as np, sys, time
sys.path.insert(0,'scratch')
import jacobi_emul as je, jacobi_emul2 as j2
JB=16; JP=32
sd=j2.steps_diag(); sc=j2.steps_cross()
def inner(G,tol,steps,dt):
    W=np.eye(JP,dtype=dt)
    for st in steps:
        J=np.eye(JP,dtype=dt)
        for (p,q) in st:
            al=G[p,p].real; be=G[q,q].real; g=G[p,q]; ag2=abs(g)**2; ab=al*be
            if ab>0 and ag2>tol*tol*ab:
                tau=0.5*(be-al); rh=1/np.sqrt(tau*tau+ag2); c2=0.5+0.5*abs(tau)*rh; c=np.sqrt(c2); sp=g*np.copysign(0.5*rh/c,tau)
                J[p,p]=c; J[p,q]=sp; J[q,p]=-np.conj(sp); J[q,q]=c
        G=(J.conj().T@G@J).astype(dt); W=(W@J).astype(dt)
    return W
def sweeps(Xt,dt,eps,maxsweeps,stop,Vt=None):
    n,L=Xt.shape; tol=np.sqrt(L)*eps
    nblk=n//JB; hist=[]
    for sw in range(maxsweeps):
        mx=0.0
        for r in range(-1,nblk-1):
            for p in range(nblk//2):
                I,J=je.rr_pair(nblk,max(r,0),p); I,J=min(I,J),max(I,J)
                idx=np.r_[I*JB:I*JB+JB, J*JB:J*JB+JB]
                P=Xt[idx]
                G=(P.conj()@P.T).astype(dt)
                d=np.sqrt(np.abs(np.diag(G).real)); C=np.abs(G)/np.maximum(np.outer(d,d),1e-300); np.fill_diagonal(C,0)
                if r<0: C[:JB,JB:]=0; C[JB:,:JB]=0
                else: C[:JB,:JB]=0; C[JB:,JB:]=0
                mx=max(mx,C.max())
                W=inner(G,tol,sd if r<0 else sc,dt)
                Xt[idx]=(W.T@P).astype(dt)
                if Vt is not None: Vt[idx]=(W.T@Vt[idx]).astype(dt)
        hist.append(mx)
        if mx<stop: break
    return hist
rng=np.random.default_rng(0)
for n in (256,512):
    A=rng.standard_normal((n,n))+1j*rng.standard_normal((n,n))
    if len(sys.argv)>1: A=A*np.logspace(0,-float(sys.argv[1]),n)[None,:]
    Q,R=np.linalg.qr(A)
    X=np.conj(R); X/=np.linalg.norm(X)
    X32=X.astype(np.complex64); V32=np.eye(n,dtype=np.complex64)
    nlow=int(sys.argv[2]) if len(sys.argv)>2 else 8
    h32=sweeps(X32,np.complex64,6e-8,nlow,3e-6,V32)
    print(n,'fp32 hist',['%.0e'%h for h in h32])
    V0=V32.astype(np.complex128)        # rows = columns of V
    Qv,Rv=np.linalg.qr(V0.T); 
    X1=(Qv.T@X)                         # Xt_new = V^T Xt
    h64=sweeps(X1,np.complex128,2.2e-16,20,1e-10)
    s=np.sort(np.linalg.norm(X1,axis=1))[::-1]*np.linalg.norm(R); sref=np.linalg.svd(A,compute_uv=False)
    print(n,'fp64 hist after prepass',['%.0e'%h for h in h64],'err %.1e'%(np.max(abs(s-sref))/sref[0]), 'relerr %.1e'%np.max(abs(s-sref)/sref))
