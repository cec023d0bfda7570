import numpy as np
a=np.loadtxt('/root/repo/gpurun_out/times.txt',dtype=np.int64)
n=len(a)//2
A,B=a[:n],a[n:]
if A[:,1].min()>B[:,1].min(): A,B=B,A
for nm,X in (("first",A),("second",B)):
    t0=X[:,1].min(); print(nm,"span start->lastend %.1f us; ahead med %.1f max %.1f; tiles med %.1f"%((X[:,4].max()-t0)/1e3, np.median(X[:,2]-X[:,1])/1e3, (X[:,2]-X[:,1]).max()/1e3, np.median(X[:,3]-X[:,2])/1e3))
print("period (start to start) %.1f us"%((B[:,1].min()-A[:,1].min())/1e3))
print("gap (last end of first -> first start of second) %.1f us"%((B[:,1].min()-A[:,4].max())/1e3))
