"""NumPy emulation of svd.cu's block Jacobi (Gram + two-sided inner sweep + apply) to check sweep counts / accuracy."""
import numpy as np, sys, time
JB=16; JP=32
def rr_pair(n, r, p):
    if p == 0: return n-1, r
    return (r+p) % (n-1), (r-p+n-1) % (n-1)
def inner_sweep(G, tol):
    W=np.eye(JP,dtype=G.dtype); mx=0.0
    for step in range(JP-1):
        J=np.eye(JP,dtype=G.dtype)
        for a in range(JP//2):
            p,q=rr_pair(JP,step,a)
            if p>q: p,q=q,p
            al=G[p,p].real; be=G[q,q].real; g=G[p,q]; ag=abs(g)
            if al>0 and be>0 and ag>0:
                off=ag/(np.sqrt(al)*np.sqrt(be)); mx=max(mx,off)
                if off>tol:
                    z=(be-al)/(2*ag); t=np.copysign(1.0,z)/(abs(z)+np.sqrt(1+z*z)); c=1/np.sqrt(1+t*t); s=c*t
                    sp=g*(s/ag)
                    J[p,p]=c; J[p,q]=sp; J[q,p]=-np.conj(sp); J[q,q]=c
        G=J.conj().T@G@J; W=W@J
    return W,mx
def block_jacobi(Xt, maxsweeps=60):
    n,L=Xt.shape; Xt=Xt.copy(); Vt=np.eye(n,dtype=Xt.dtype)
    tol=np.sqrt(L)*2.22e-16
    nb=(n+JB-1)//JB; nblk=max(2,nb+(nb&1))
    hist=[]
    for sw in range(maxsweeps):
        mx=0.0
        for r in range(nblk-1):
            for p in range(nblk//2):
                I,J=rr_pair(nblk,r,p)
                if I>J: I,J=J,I
                idx=[i for i in list(range(I*JB,I*JB+JB))+list(range(J*JB,J*JB+JB))]
                valid=[i for i in idx if i<n]
                P=np.zeros((JP,L),dtype=Xt.dtype); PV=np.zeros((JP,n),dtype=Xt.dtype)
                for k,i in enumerate(idx):
                    if i<n: P[k]=Xt[i]; PV[k]=Vt[i]
                G=P.conj()@P.T
                W,m=inner_sweep(G,tol); mx=max(mx,m)
                Pn=W.T@P; PVn=W.T@PV
                for k,i in enumerate(idx):
                    if i<n: Xt[i]=Pn[k]; Vt[i]=PVn[k]
        hist.append(mx)
        if mx<=tol: break
    return Xt,Vt,sw+1,hist
def test(m,n,cplx,kind,pre=True):
    rng=np.random.default_rng(0)
    A=rng.standard_normal((m,n)) + (1j*rng.standard_normal((m,n)) if cplx else 0)
    if kind=='graded': A=A*np.logspace(0,-12,n)[None,:]
    if kind=='lowrank':
        r=max(1,min(m,n)//3); A=A[:, :r] @ (rng.standard_normal((r,n)) + (1j*rng.standard_normal((r,n)) if cplx else 0))
    if kind=='unif': A=rng.random((m,n))+ (1j*rng.random((m,n)) if cplx else 0)
    sref=np.linalg.svd(A,compute_uv=False)
    wide=m<n
    Q,R=np.linalg.qr(A.conj().T if wide else A)
    t=time.time(); Xt,Vt,sw,hist=block_jacobi(np.conj(R)); dt=time.time()-t
    s=np.linalg.norm(Xt,axis=1); order=np.argsort(-s,kind='stable'); s2=s[order]
    k=len(s2)
    Vs=Vt[order]; Y=Xt[order]
    with np.errstate(all='ignore'):
        inv=np.where(s2>0,1/np.where(s2>0,s2,1),0)
    if not wide:
        U=Q@Vs.T; Vh=np.conj(Y)*inv[:,None]
    else:
        U=(Y*inv[:,None]).T; Vh=np.conj(Vs)@Q.conj().T
    rec=np.linalg.norm(U*s2@Vh-A)/np.linalg.norm(A)
    err=np.max(np.abs(s2-sref))/sref[0]
    relerr=np.max(np.abs(s2-sref)[sref>1e-13*sref[0]]/sref[sref>1e-13*sref[0]])
    orthU=np.linalg.norm(U.conj().T@U-np.eye(k)); orthV=np.linalg.norm(Vh@Vh.conj().T-np.eye(k))
    print(m,n,'c' if cplx else 'r',kind,'sweeps',sw,'err %.1e rel %.1e rec %.1e orthU %.1e orthV %.1e t=%.1fs'%(err,relerr,rec,orthU,orthV,dt), ['%.0e'%h for h in hist])
if __name__=="__main__":
    test(20,12,False,'rand'); test(12,20,True,'rand'); test(3,3,False,'rand'); test(5,1,True,'rand'); test(2,6,False,'rand')
    test(128,64,False,'unif'); test(64,96,True,'rand'); test(64,64,True,'graded'); test(96,64,False,'lowrank'); test(100,130,True,'lowrank')
    test(256,256,True,'rand')
