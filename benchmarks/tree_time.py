import os,sys,time
sys.path.insert(0,".")
import numpy as np
from spark_ensemble_b200 import _native as N
from spark_ensemble_b200.context import Context
ctx=Context(0); n,d=100_000_000,128
ctx.alloc(N.SLOT_X,d,n); ctx.fill_synthetic(N.SLOT_X,"normal",3,0,1); ctx.alloc(N.SLOT_H,1,n)
for bins in (1,):
    ctx.set_option("tree_bins", bins)
    for depth in (6,):
        nn=2**(depth+1)-1; idx=np.arange(nn); leaf=idx>=2**depth-1
        tree={"feature":np.where(leaf,-1,(idx*37)%d),"threshold":np.where(leaf,0.0,((idx*13)%7-3)*0.2),"left":np.where(leaf,0,2*idx+1),"right":np.where(leaf,0,2*idx+2),"value":np.linspace(-1,1,nn)}
        ctx.tree_predict(tree,N.SLOT_H,0); ctx.sync()
        ctx.kernel_timing(True); ctx.kernel_times_reset()
        for _ in range(10): ctx.tree_predict(tree,N.SLOT_H,0)
        kt=ctx.kernel_times(); ctx.kernel_timing(False)
        print("bins",bins,"depth",depth,{k:round(v["ms"]/v["launches"],4) for k,v in kt.items()}, flush=True)
ctx.close()
