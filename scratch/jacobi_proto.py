import numpy as np, sys, time
def rr_pair(n, r, p):
    # circle method: n even players, round r in [0,n-1), pair p in [0,n/2)
    if p == 0: return n-1, r
    return (r+p) % (n-1), (r-p+n-1) % (n-1)
def check_rr(n):
    seen=set()
    for r in range(n-1):
        used=set()
        for p in range(n//2):
            a,b=rr_pair(n,r,p); assert a!=b and a not in used and b not in used; used|={a,b}; seen.add((min(a,b),max(a,b)))
    assert len(seen)==n*(n-1)//2
for n in (2,4,6,8,32,64): check_rr(n)

def panel_sweep(P, W, tol):
    # P: (w, L) rows are jacobi columns; one full round-robin sweep; returns max off
    w=P.shape[0]; ww = w + (w&1); mx=0.0
    for r in range(ww-1):
        for p in range(ww//2):
            a,b=rr_pair(ww,r,p)
            if a>=w or b>=w: continue
            xa,xb=P[a],P[b]
            al=np.vdot(xa,xa).real; be=np.vdot(xb,xb).real; g=np.vdot(xa,xb)  # conj(xa).xb
            ag=abs(g)
            if al==0 or be==0: continue
            off=ag/np.sqrt(al*be); mx=max(mx,off)
            if off<=tol: continue
            zeta=(be-al)/(2*ag); t=np.sign(zeta)/(abs(zeta)+np.sqrt(1+zeta*zeta)) if zeta!=0 else 1.0
            c=1/np.sqrt(1+t*t); s=c*t; ph=g/ag
            na=c*xa - s*np.conj(ph)*xb; nb=s*ph*xa + c*xb
            P[a],P[b]=na,nb
            wa,wb=W[a].copy(),W[b].copy()
            W[a]=c*wa - s*np.conj(ph)*wb; W[b]=s*ph*wa + c*wb
    return mx

def block_jacobi(X, b=8, tol=None, maxsweeps=40):
    # X: (mm, n) -> orthogonalise columns. Xt rows = columns.
    mm,n=X.shape
    tol = tol or np.sqrt(mm)*2.2e-16
    Xt=X.T.copy(); Vt=np.eye(n,dtype=X.dtype)
    nblk=(n+b-1)//b; nblk+= nblk&1
    blocks=[list(range(i*b,min(n,(i+1)*b))) for i in range(nblk)]
    sweeps=0
    while sweeps<maxsweeps:
        mx=0.0
        for r in range(max(nblk-1,1)):
            for p in range(nblk//2):
                I,J=rr_pair(nblk,r,p) if nblk>1 else (0,0)
                idx=blocks[I]+blocks[J]
                if len(idx)<2: continue
                P=Xt[idx].copy(); W=np.eye(len(idx),dtype=X.dtype)
                mx=max(mx,panel_sweep(P,W,tol))
                Xt[idx]=P
                Vt[idx]=W @ Vt[idx]   # rows of W are rotated columns: Wrow_q = sum_p coeff * e_p -> new Vt[q] = sum_p W[q,p] Vt[p]
        sweeps+=1
        if mx<=tol: break
    s=np.linalg.norm(Xt,axis=1)
    return Xt,Vt,s,sweeps

def test(m,n,cplx,b,kind):
    rng=np.random.default_rng(0)
    A=rng.standard_normal((m,n)) + (1j*rng.standard_normal((m,n)) if cplx else 0)
    if kind=='graded':
        A=A*np.logspace(0,-12,n)[None,:]
    if kind=='lowrank':
        A=A[:, :n//3] @ (rng.standard_normal((n//3,n)) + (1j*rng.standard_normal((n//3,n)) if cplx else 0))
    sref=np.linalg.svd(A,compute_uv=False)
    Q,R=np.linalg.qr(A)
    out={}
    for mode in ('R','RH'):
        X = R if mode=='R' else R.conj().T
        t=time.time(); Xt,Vt,s,sw=block_jacobi(X,b); dt=time.time()-t
        order=np.argsort(-s); s2=s[order]
        err=np.linalg.norm(s2-sref)/np.linalg.norm(sref)
        # reconstruction: X V2 = U2 S, V2[:,j]=Vt[j,:]
        V2=Vt.T; rec=np.linalg.norm(X@V2 - Xt.T)/np.linalg.norm(X)
        orth=np.linalg.norm(V2.conj().T@V2-np.eye(n))
        out[mode]=(sw,err,rec,orth)
    print(m,n,'c' if cplx else 'r',kind,'b=%d'%b,{k:(v[0],'%.1e'%v[1],'%.1e'%v[2],'%.1e'%v[3]) for k,v in out.items()})
if __name__=="__main__":
    test(48,32,True,8,'rand'); test(48,32,False,8,'rand'); test(64,64,True,8,'graded'); test(64,48,True,8,'lowrank')
    test(128,128,True,16,'rand'); test(96,96,False,16,'graded')
