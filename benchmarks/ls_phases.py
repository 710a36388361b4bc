import sys,time
sys.path.insert(0,".")
import numpy as np
from spark_ensemble_b200 import _native as N
from spark_ensemble_b200.context import Context
ctx=Context(0)
for n in (6_250_000, 50_000_000):
    ctx.gbm_configure(n,0,1,"bernoulli",0.0,False)
    ctx.fill_synthetic(N.SLOT_Y,"bernoulli",1,0.4,1.0); ctx.fill(N.SLOT_F,0.0); ctx.fill_synthetic(N.SLOT_H,"normal",3,0.0,1.0)
    for ct in (4,3):
        ctx.set_option("ls_ctas_per_sm", ct); ctx.set_option("fused_timing",1)
        for _ in range(3): ctx.gbm_linesearch_brent()
        a,l,ne=ctx.gbm_linesearch_brent()
        print(n, "ctas", ct, "evals", ne, "pass us/eval", ctx.get_option("last_fused_stats_us")/ne, "fold us/eval", ctx.get_option("last_fused_brent_us")/ne, "workers", ctx.get_option("last_ls_workers"))
        ctx.set_option("fused_timing",0)
        t0=time.perf_counter()
        for _ in range(20): ctx.gbm_linesearch_brent()
        print("   wall us per search", 1e6*(time.perf_counter()-t0)/20, "per eval", 1e6*(time.perf_counter()-t0)/20/ne)
ctx.close()
