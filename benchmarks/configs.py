#!/usr/bin/env python
"""The five BASELINE.json configurations, device-resident, on one GPU (per-GPU shard sizes for the 8-GPU ones).

    python benchmarks/configs.py [--out gpurun_out/configs.json]

C1 GBMRegressor cpusmall 20 rounds (host base learner: the reference's plumbing case, timed end to end)
C2 GBMRegressor 10 M x 64 squared, 100 rounds           (device rounds: Brent over one-pass statistics + K1)
C3 GBMClassifier 50 M bernoulli, 200 rounds              (device rounds: Brent with one K2 pass per evaluation + K1)
C4 BoostingClassifier SAMME.R 26 classes, 12.5 M rows/GPU (1/8 of 100 M): weight-update kernel per round
C5 BaggingRegressor.transform 512 models x 6.25 M rows/GPU (1/8 of 50 M): aggregation kernel
The direction of every device round is a fixed synthetic vector (the base learner is third party and not timed)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spark_ensemble_b200 import _native as N  # noqa: E402
from spark_ensemble_b200.context import Context  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    sc = 0.1 if args.quick else 1.0
    res = {}
    # ---- C1
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import GBMRegressor
    d = np.load(os.path.join(ROOT, "tests", "golden", "cpusmall.npz"))
    X, y = d["X"].astype(np.float32), d["y"].astype(np.float64)
    t0 = time.perf_counter()
    m = GBMRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=5)).setNumBaseLearners(20).fit(DataFrame(features=X, label=y))
    dt = time.perf_counter() - t0
    res["C1"] = {"config": "GBMRegressor cpusmall 8192x12, 20 rounds, squared (sklearn tree on host)", "seconds": dt,
                 "final_train_loss": m.trainingHistory[-1]["trainLoss"]}
    print("C1", res["C1"], flush=True)
    ctx = Context(0)
    # ---- C2 / C3: device rounds
    for name, loss, n, rounds in (("C2", "squared", int(10e6 * sc), 100), ("C3", "bernoulli", int(50e6 * sc), 200)):
        ctx.gbm_configure(n, 0, 1, loss, 0.0, False)
        ctx.fill_synthetic(N.SLOT_Y, "bernoulli" if loss == "bernoulli" else "normal", 1, 0.4, 1.0)
        ctx.fill(N.SLOT_F, 0.0)
        ctx.fill_synthetic(N.SLOT_H, "normal", 3, 0.0, 1.0)
        ctx.gbm_pseudo_residuals(False)
        ctx.sync()
        for _ in range(3):
            a, _, _ = ctx.gbm_linesearch_brent(0.0, 100.0, 1.0, 1e-6, 1e-6, 100)
            ctx.gbm_update([0.1 * a], residual=True, loss=True)
        ctx.sync()
        evals = 0
        t0 = time.perf_counter()
        for _ in range(rounds):
            a, _, ne = ctx.gbm_linesearch_brent(0.0, 100.0, 1.0, 1e-6, 1e-6, 100)
            ctx.gbm_update([0.1 * a], residual=True, loss=True)
            evals += ne
        ctx.sync()
        dt = time.perf_counter() - t0
        res[name] = {"config": f"{loss} {n} rows, {rounds} rounds, Brent tol 1e-6", "ms_per_round": 1e3 * dt / rounds,
                     "rows_per_s": n * rounds / dt, "brent_evals_per_round": evals / rounds}
        print(name, res[name], flush=True)
        # opt-in curvature-based line search (same minimiser to the same tolerance, not the reference's iterates)
        evals = 0
        t0 = time.perf_counter()
        for _ in range(rounds):
            a, _, ne = ctx.gbm_linesearch_newton(0.0, 100.0, 1.0, 1e-6, 1e-6, 100)
            ctx.gbm_update([0.1 * a], residual=True, loss=True)
            evals += ne
        ctx.sync()
        dt = time.perf_counter() - t0
        res[name + "_newton"] = {"config": f"{loss} {n} rows, {rounds} rounds, Newton line search tol 1e-6",
                                 "ms_per_round": 1e3 * dt / rounds, "rows_per_s": n * rounds / dt,
                                 "evals_per_round": evals / rounds}
        print(name + "_newton", res[name + "_newton"], flush=True)
        if loss == "squared":
            # opt-in device-resident round: closed-form alpha on the device, two launches, no host synchronisation
            for _ in range(3):
                ctx.gbm_round_squared_async(0.1)
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(rounds):
                ctx.gbm_round_squared_async(0.1)
            ctx.sync()
            dt = time.perf_counter() - t0
            res[name + "_async"] = {"config": f"{loss} {n} rows, {rounds} rounds, device-resident closed-form step",
                                    "ms_per_round": 1e3 * dt / rounds, "rows_per_s": n * rounds / dt}
            print(name + "_async", res[name + "_async"], flush=True)
    for s in (N.SLOT_F, N.SLOT_H, N.SLOT_R):
        ctx.free(s)
    # ---- C4
    K, n, rounds = 26, int(12.5e6 * sc), 20
    ctx.boost_configure(n, K, True)
    ctx.fill_synthetic(N.SLOT_Y, "randint", 1, 0, K)
    ctx.fill_synthetic(N.SLOT_PROBA, "uniform", 2, 0.001, 0.08)
    ctx.fill(N.SLOT_BW, 1.0)
    sw = ctx.slot_sum(N.SLOT_BW)
    t0 = time.perf_counter()
    for _ in range(rounds):
        e, sw = ctx.boost_real_update(sw)
    ctx.sync()
    dt = time.perf_counter() - t0
    res["C4"] = {"config": f"SAMME.R K={K}, {n} rows/GPU, {rounds} rounds", "ms_per_round": 1e3 * dt / rounds,
                 "rows_per_s": n * rounds / dt, "gbs": (4 * K + 12) * n * rounds / dt / 1e9}
    print("C4", res["C4"], flush=True)
    ctx.free(N.SLOT_PROBA)
    # ---- C5
    M, n, reps = 512, int(6.25e6 * sc), 10
    ctx.agg_configure(N.AGG_BAGGING_REGRESSOR, M, 0, 1, 0, n)
    ctx.fill_synthetic(N.SLOT_P, "normal", 7, 0.0, 1.0)
    ctx.agg_run()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.agg_run()
    ctx.sync()
    dt = time.perf_counter() - t0
    res["C5"] = {"config": f"BaggingRegressor.transform M={M}, {n} rows/GPU", "ms_per_pass": 1e3 * dt / reps,
                 "rows_per_s": n * reps / dt, "gbs": (4 * M + 4) * n * reps / dt / 1e9}
    print("C5", res["C5"], flush=True)
    ctx.close()
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
