"""Emulate the cross-round / diag-round ordering of svd.cu and compare sweep counts with the full 31-step ordering."""
import numpy as np, sys
sys.path.insert(0,'scratch')
import jacobi_emul as je
JB=16; JP=32
def steps_cross():
    return [[(a, JB+((a+s)%JB)) for a in range(JB)] for s in range(JB)]
def steps_diag():
    out=[]
    for s in range(JB-1):
        st=[]
        for t in range(JB//2):
            x,y=je.rr_pair(JB,s,t); st.append((min(x,y),max(x,y))); st.append((min(x,y)+JB,max(x,y)+JB))
        out.append(st)
    return out
def inner(G,tol,steps):
    W=np.eye(JP,dtype=G.dtype); big=False; rot=False
    for st in steps:
        J=np.eye(JP,dtype=G.dtype)
        for (p,q) in st:
            al=G[p,p].real; be=G[q,q].real; g=G[p,q]; ag2=abs(g)**2; ab=al*be
            if ab>0 and ag2>tol*tol*ab:
                rot=True
                if ag2>1e-20*ab: big=True
                tau=0.5*(be-al); rh=1/np.sqrt(tau*tau+ag2); c2=0.5+0.5*abs(tau)*rh; c=np.sqrt(c2); sp=g*np.copysign(0.5*rh/c,tau)
                J[p,p]=c; J[p,q]=sp; J[q,p]=-np.conj(sp); J[q,q]=c
        G=J.conj().T@G@J; W=W@J
    return W,big
def jacobi(Xt,maxsweeps=40):
    n,L=Xt.shape; Xt=Xt.copy()/np.linalg.norm(Xt); tol=np.sqrt(L)*2.22e-16
    nb=(n+JB-1)//JB; nblk=max(2,nb+(nb&1)); sc=steps_cross(); sd=steps_diag()
    for sw in range(maxsweeps):
        anybig=False
        for r in range(-1,nblk-1):
            for p in range(nblk//2):
                I,J=je.rr_pair(nblk,max(r,0),p); I,J=min(I,J),max(I,J)
                idx=list(range(I*JB,I*JB+JB))+list(range(J*JB,J*JB+JB))
                P=np.zeros((JP,L),dtype=Xt.dtype)
                for k,i in enumerate(idx):
                    if i<n: P[k]=Xt[i]
                G=P.conj()@P.T
                W,big=inner(G,tol,sd if r<0 else sc); anybig|=big
                Pn=W.T@P
                for k,i in enumerate(idx):
                    if i<n: Xt[i]=Pn[k]
        if not anybig: return Xt,sw+1
    return Xt,maxsweeps
rng=np.random.default_rng(0)
for (m,n) in [(96,96),(200,200),(256,256),(40,24),(10,10)]:
    A=rng.standard_normal((m,n))+1j*rng.standard_normal((m,n))
    Q,R=np.linalg.qr(A)
    Xt,sw=jacobi(np.conj(R))
    s=np.sort(np.linalg.norm(Xt,axis=1))[::-1]*np.linalg.norm(R); sref=np.linalg.svd(A,compute_uv=False)
    Y=Xt/np.linalg.norm(Xt,axis=1)[:,None]
    print(m,n,'sweeps(new order, early stop)',sw,'err %.1e'%(np.max(abs(s-sref))/sref[0]),'orth %.1e'%np.linalg.norm(Y.conj()@Y.T-np.eye(len(Y))))
    _,_,sw_old,hist=je.block_jacobi(np.conj(R))
    print('     old full-31 ordering sweeps', sw_old)
