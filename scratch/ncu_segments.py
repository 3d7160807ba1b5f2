"""Segment an ncu SASS-level sample profile by barrier landmarks: python scratch/ncu_segments.py rep [ninstr]"""
import csv, subprocess, sys, io
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
his=[i for i,r in enumerate(rows) if r and r[0]=='Address']
hi=his[0]; hdr=rows[hi]
end=his[1]-1 if len(his)>1 else len(rows)
data=[r for r in rows[hi+1:end] if len(r)>5]
si=hdr.index('# Samples'); ei=hdr.index('Instructions Executed')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[si] or 0) for r in data)
print('instructions',len(data),'total samples',tot)
segstart=0; acc=0; agg={}; ndm=0; nfp=0
for k,r in enumerate(data):
    s=int(r[si] or 0); acc+=s
    src=r[1]
    if 'DMMA' in src: ndm+=1
    if any(t in src for t in ('DFMA','DMUL','DADD','MUFU')): nfp+=1
    for i in stall_cols:
        v=int(r[i] or 0)
        if v: agg[hdr[i]]=agg.get(hdr[i],0)+v
    if any(t in src for t in ('BAR.SYNC','UCGABAR_WAIT','EXIT','WARPSYNC')) or k==len(data)-1:
        if acc>0.004*tot:
            top=sorted(agg.items(),key=lambda x:-x[1])[:3]
            print('%5d..%5d %6d (%4.1f%%) dmma %3d fp64 %3d exec=%-7s %s | %s'%(segstart,k,acc,100*acc/tot,ndm,nfp,r[ei],src.strip()[:28],top))
        segstart=k+1; acc=0; agg={}; ndm=0; nfp=0
