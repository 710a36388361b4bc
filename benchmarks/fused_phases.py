#!/usr/bin/env python
"""Where a one-launch squared-loss round spends its time: in-kernel %globaltimer stamps (statistics phase, fold +
exchange + Brent, update phase), the kernel's CUDA-event duration and the host wall clock per round."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spark_ensemble_b200 import _native as N  # noqa: E402
from spark_ensemble_b200.context import Context  # noqa: E402

ctx = Context(0)
for n in (100_000_000, 50_000_000, 25_000_000, 12_500_000, 6_250_000):
    ctx.gbm_configure(n, 0, 1, "squared", 0.0, False)
    ctx.fill_synthetic(N.SLOT_Y, "normal", 1, 0.0, 1.0)
    ctx.fill_synthetic(N.SLOT_F, "normal", 2, 0.0, 0.5)
    ctx.copy_slot(N.SLOT_H, N.SLOT_Y)
    ctx.gbm_update([0.5], residual=False, loss=False)
    ctx.copy_slot(N.SLOT_H, N.SLOT_F)   # h = 0.5 y + N(0, 0.5)
    ctx.fill(N.SLOT_F, 0.0)
    ctx.gbm_pseudo_residuals(False)
    for fused, l2m in ((1, 0), (0, 0)):
        ctx.set_option("fused_round", fused)
        ctx.set_option("fused_l2_mode", l2m)
        ctx.set_option("fused_timing", 0)
        for _ in range(5):
            ctx.gbm_round(0.01, True, 1e-6, 100, residual=True)
        ctx.sync()
        t0 = time.perf_counter()
        R = 200
        for _ in range(R):
            ctx.gbm_round(0.01, True, 1e-6, 100, residual=True)
        ctx.sync()
        wall = 1e6 * (time.perf_counter() - t0) / R
        ctx.kernel_timing(True); ctx.kernel_times_reset()
        for _ in range(50):
            ctx.gbm_round(0.01, True, 1e-6, 100, residual=True)
        kt = ctx.kernel_times(); ctx.kernel_timing(False)
        out = {"rows": n, "fused": fused, "l2_mode": l2m, "wall_us_per_round": wall,
               "kernel_event_us": {k: 1e3 * v["ms"] / v["launches"] for k, v in kt.items()}}
        if fused:
            ctx.set_option("fused_timing", 1)
            ph = []
            for _ in range(20):
                _, _, ne = ctx.gbm_round(0.01, True, 1e-6, 100, residual=True)
                ph.append([ctx.get_option("last_fused_stats_us"), ctx.get_option("last_fused_brent_us"),
                           ctx.get_option("last_fused_update_us")])
            ph.sort(key=lambda p: sum(p))
            out["phases_us_median"] = ph[len(ph) // 2]
            out["brent_evals"] = ne
            ctx.set_option("fused_timing", 0)
        print(json.dumps(out), flush=True)
ctx.close()
