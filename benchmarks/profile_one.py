#!/usr/bin/env python
"""Launch a handful of kernels of one family for ncu (benchmarks/profile_one.py <what> [rows])."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spark_ensemble_b200 import _native as N  # noqa: E402
from spark_ensemble_b200.context import Context  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "bernoulli"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
ctx = Context(0)
if what in ("squared", "bernoulli", "exponential", "logcosh", "absolute"):
    ctx.gbm_configure(n, 0, 1, what, 0.9, False)
    ctx.fill_synthetic(N.SLOT_Y, "bernoulli" if what in ("bernoulli", "exponential") else "normal", 1, 0.4, 1)
    ctx.fill_synthetic(N.SLOT_F, "normal", 2, 0, 0.5)
    ctx.fill_synthetic(N.SLOT_H, "normal", 3, 0, 1)
    for _ in range(3):
        ctx.gbm_linesearch_eval([0.7])
        ctx.gbm_update([1e-3], residual=True, loss=True)
elif what in ("fused_round", "two_launch_round"):
    ctx.gbm_configure(n, 0, 1, "squared", 0.0, False)
    ctx.fill_synthetic(N.SLOT_Y, "normal", 1, 0, 1)
    ctx.fill_synthetic(N.SLOT_F, "normal", 2, 0, 0.5)
    ctx.copy_slot(N.SLOT_H, N.SLOT_Y)
    ctx.gbm_update([0.5], residual=False, loss=False)
    ctx.copy_slot(N.SLOT_H, N.SLOT_F)   # h = 0.5 y + N(0, 0.5)
    ctx.fill(N.SLOT_F, 0.0)
    ctx.gbm_pseudo_residuals(False)
    ctx.set_option("fused_round", 1 if what == "fused_round" else 0)
    for _ in range(8):
        ctx.gbm_round(0.1, True, 1e-6, 100, residual=True)
elif what == "ls_persist":
    ctx.gbm_configure(n, 0, 1, "bernoulli", 0.0, False)
    ctx.fill_synthetic(N.SLOT_Y, "bernoulli", 1, 0.4, 1)
    ctx.fill_synthetic(N.SLOT_F, "normal", 2, 0, 0.5)
    ctx.fill_synthetic(N.SLOT_H, "normal", 3, 0, 1)
    for _ in range(3):
        ctx.gbm_linesearch_brent()
elif what.startswith("logloss"):
    K = int(what[7:])
    ctx.gbm_configure(n, 0, K, "logloss", 0.0, False)
    ctx.fill_synthetic(N.SLOT_Y, "randint", 1, 0, K)
    ctx.fill_synthetic(N.SLOT_F, "normal", 2, 0, 0.5)
    ctx.fill_synthetic(N.SLOT_H, "normal", 3, 0, 1)
    for _ in range(3):
        ctx.gbm_linesearch_eval(np.full(K, 0.7))
        ctx.gbm_update(np.full(K, 1e-3), residual=True, loss=True)
elif what == "samme_r":
    K = 26
    ctx.boost_configure(n, K, True)
    ctx.fill_synthetic(N.SLOT_Y, "randint", 1, 0, K)
    ctx.fill_synthetic(N.SLOT_PROBA, "uniform", 2, 0.001, 0.08)
    for _ in range(3):
        ctx.fill(N.SLOT_BW, 1.0)
        ctx.boost_real_update(float(n))
elif what == "tree":
    d = 128
    ctx.alloc(N.SLOT_X, d, n)
    ctx.fill_synthetic(N.SLOT_X, "normal", 3, 0, 1)
    ctx.alloc(N.SLOT_H, 1, n)
    depth = 6
    nn = 2 ** (depth + 1) - 1
    idx = np.arange(nn)
    leaf = idx >= 2 ** depth - 1
    tree = {"feature": np.where(leaf, -1, (idx * 37) % d), "threshold": np.where(leaf, 0.0, ((idx * 13) % 7 - 3) * 0.2),
            "left": np.where(leaf, 0, 2 * idx + 1), "right": np.where(leaf, 0, 2 * idx + 2), "value": np.linspace(-1, 1, nn)}
    for _ in range(3):
        ctx.tree_predict(tree, N.SLOT_H, 0)
elif what == "wmedian":
    M = 32
    ctx.agg_configure(N.AGG_BOOSTING_REG_MEDIAN, M, 2, 1, 0, n)
    ctx.fill_synthetic(N.SLOT_P, "uniform", 5, 0.01, 1.0)
    for _ in range(3):
        ctx.agg_run(np.linspace(0.5, 1.5, M))   # distinct weights: the margin-checked fast path
elif what == "agg_real":
    M, K = 10, 26
    ctx.agg_configure(N.AGG_BOOSTING_REAL, M, K, 1, 0, n)
    ctx.fill_synthetic(N.SLOT_P, "uniform", 5, 0.01, 1.0)
    for _ in range(3):
        ctx.agg_run()
elif what in ("votes", "wvotes"):
    M, K = 64, 26
    ctx.agg_configure(N.AGG_BAGGING_HARD if what == "votes" else N.AGG_BOOSTING_DISCRETE, M, K, 1, 0, n)
    ctx.fill_synthetic(N.SLOT_P, "randint", 5, 0, K)
    for _ in range(3):
        ctx.agg_run(None if what == "votes" else np.linspace(0.5, 1.5, M))
ctx.sync()
ctx.close()
