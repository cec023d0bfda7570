"""Per-CTA timing of the last two k_neuron launches of a run: HHD_DEBUG_TIMES=out.txt python bench.py ... ; python tools/cta_times.py out.txt
(the engine writes out.txt = block, start, after the synaptic share, end of tiles, end [ns, %globaltimer] and out.txt.passes = first start / close per step)."""
import numpy as np, sys
a=np.loadtxt(sys.argv[1],dtype=np.int64)
n=len(a)//2; X=a[n:] if a[n:,1].min()>a[:n,1].min() else a[:n]
t0=X[:,1].min()
ah=(X[:,2]-X[:,1])/1e3; tl=(X[:,3]-X[:,2])/1e3; en=(X[:,4]-t0)/1e3
print("  span %.1f | ahead sum %.0f CTA-us (mean over active %.1f, max %.1f) | tiles sum %.0f CTA-us (mean %.1f) | end mean %.1f"%(en.max(), ah.sum(), ah[ah>0.5].mean() if (ah>0.5).any() else 0, ah.max(), tl.sum(), tl.mean(), en.mean()))
p=np.loadtxt(sys.argv[1]+".passes",dtype=np.int64); p=p[np.argsort(p[:,1])]
per=np.diff(p[:,1])/1e3; per=per[(per>0)&(per<1000)]
print("  period mean %.1f med %.1f"%(per.mean(), np.median(per)))
