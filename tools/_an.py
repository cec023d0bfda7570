import numpy as np
a=np.loadtxt('/root/repo/gpurun_out/times.txt',dtype=np.int64)
t0=a[:,1].min()
st=(a[:,1]-t0)/1e3; ah=(a[:,2]-a[:,1])/1e3; tl=(a[:,3]-a[:,2])/1e3; en=(a[:,4]-t0)/1e3
print("start  min/med/max", st.min(), np.median(st), st.max())
print("ahead  min/med/max", ah.min(), np.median(ah), ah.max())
print("tiles  min/med/max", tl.min(), np.median(tl), tl.max())
print("end    min/med/max", en.min(), np.median(en), en.max())
idx=np.argsort(-en)[:12]
for i in idx: print(a[i,0], "start %.1f ahead %.1f tiles %.1f end %.1f"%(st[i],ah[i],tl[i],en[i]))
print("late starters (>5us):", (st>5).sum())
h,_=np.histogram(ah,bins=[0,1,2,4,8,16,32,64,128,1e9]); print("ahead hist", h)
for lo in range(0,len(a),74): print(lo, "ahead mean %.1f max %.1f tiles mean %.1f"%(ah[lo:lo+74].mean(), ah[lo:lo+74].max(), tl[lo:lo+74].mean()))
