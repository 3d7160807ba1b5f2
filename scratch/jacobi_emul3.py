"""Sweep-count experiments: single pass per pair (current kernel) vs full diagonalisation of the pair Gram
matrix (block Jacobi proper), with / without norm sorting, for n up to 1024."""
import numpy as np, sys, time
sys.path.insert(0,'scratch')
import jacobi_emul as je
import jacobi_emul2 as j2
JB=16; JP=32
def run(Xt, mode, maxsweeps=30, JB=16, sort_inner=True):
    JP=2*JB
    n,L=Xt.shape; Xt=Xt.copy()/np.linalg.norm(Xt); tol=np.sqrt(L)*2.22e-16
    nb=(n+JB-1)//JB; nblk=max(2,nb+(nb&1))
    hist=[]
    for sw in range(maxsweeps):
        mx=0.0
        for r in range(-1 if mode=='single' else 0,nblk-1):
            for p in range(nblk//2):
                I,J=je.rr_pair(nblk,max(r,0),p); I,J=min(I,J),max(I,J)
                idx=np.r_[I*JB:I*JB+JB, J*JB:J*JB+JB]
                idx=idx[idx<n]
                P=Xt[idx]
                G=P.conj()@P.T
                d=np.sqrt(np.abs(np.diag(G).real)); C=np.abs(G)/np.maximum(np.outer(d,d),1e-300); np.fill_diagonal(C,0)
                if mode=='single':
                    if r<0: C2=C.copy(); C2[:JB,JB:]=0; C2[JB:,:JB]=0; mx=max(mx,C2.max())
                    else: mx=max(mx,C[:JB,JB:].max() if C.shape[0]>JB else 0)
                    if len(idx)<JP: 
                        Gp=np.zeros((JP,JP),dtype=G.dtype); Gp[:len(idx),:len(idx)]=G; 
                        W,_=j2.inner(Gp,tol,j2.sd if r<0 else j2.sc); W=W[:len(idx),:len(idx)]
                    else:
                        W,_=j2.inner(G,tol,j2.sd if r<0 else j2.sc)
                else:
                    mx=max(mx,C.max())
                    if C.max()<=tol: continue
                    w,W=np.linalg.eigh(G)
                    W=W[:,::-1]  # descending
                    # choose W close to identity ordering? keep sorted (de Rijk-like)
                Xt[idx]=W.T@P
        hist.append(mx)
        if mx<1e-10: return Xt,sw+1,hist
    return Xt,maxsweeps,hist
j2.sd=j2.steps_diag(); j2.sc=j2.steps_cross()
if __name__=='__main__':
    rng=np.random.default_rng(0)
    for (m,n) in [(256,256),(512,512)]:
        A=rng.standard_normal((m,n))+1j*rng.standard_normal((m,n))
        Q,R=np.linalg.qr(A)
        for mode in ('single','full'):
            t=time.time(); Xt,sw,hist=run(np.conj(R),mode)
            s=np.sort(np.linalg.norm(Xt,axis=1))[::-1]*np.linalg.norm(R); sref=np.linalg.svd(A,compute_uv=False)
            print(m,n,mode,'sweeps',sw,'err %.1e'%(np.max(abs(s-sref))/sref[0]),['%.0e'%h for h in hist],'%.1fs'%(time.time()-t),flush=True)
