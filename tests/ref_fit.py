"""Oracle-driven restatement of the reference's fit/predict CONTROL FLOW (test infrastructure).

The per-row arithmetic comes from oracle/ (fp64 C); the loop structure follows
regression/GBMRegressor.scala:287-474, classification/GBMClassifier.scala:272-494 and
classification/BoostingClassifier.scala:156-280.  Used (a) to generate tests/golden/*.json and
(b) to REPLAY a fit produced by the product with the product's own base models, so the two sides see
identical directions and differ only in the hot-path arithmetic.
"""
import math

import numpy as np

from oracle import oracle as O


def _f32(a):
    """Base-model outputs enter the device as fp32; the oracle sees the same rounded values."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def ref_gbm_regressor_fit(oracle, X, y, w, learner, loss="squared", alpha_q=0.9, num_learners=10, lr=1.0,
                          tol=1e-6, max_iter=100, optimized=True, models=None, subspaces=None, init=None):
    """GBMRegressor.train with gradient updates, no validation, ratio-1 bags.  When `models` is given
    they are replayed (no fitting)."""
    n, d = X.shape
    lid = O.LOSS_IDS[loss]
    param = alpha_q
    if init is None:
        init = float(np.mean(y)) if loss == "squared" else float(np.sort(y)[max(int(math.ceil(
            (0.5 if loss in ("absolute", "huber") else alpha_q) * n)) - 1, 0)])
    F = np.full((1, n), np.float64(np.float32(init)))
    alphas, losses, fitted = [], [], []
    for i in range(num_learners):
        sub = np.arange(d) if subspaces is None else subspaces[i]
        if models is None:
            r, _, _ = oracle.pseudo_residuals(lid, param, 1, y, w, F, False)
            m = learner.fit(X[:, sub], r[0], w)
        else:
            m = models[i]
        fitted.append(m)
        h = _f32(m.predict(X[:, sub])).reshape(1, n)
        if optimized:
            f = lambda a: oracle.linesearch_eval(lid, param, y, w, F, h, [a])[0]
            a, _, st = oracle.brent(f, 0.0, 100.0, 1.0, tol, tol, max_iter)
            assert st == 0
        else:
            a = 1.0
        oracle.update(F, h, [lr * a])
        alphas.append(a)
        losses.append(oracle.mean_loss(lid, param, 1, y, F))
    return {"init": init, "alpha": alphas, "train_loss": losses, "F": F[0].copy(), "models": fitted}


def ref_gbm_classifier_replay(oracle, X, y, w, loss, num_classes, init_raw, models, subspaces, alphas, lr=1.0):
    """Replays GBMClassifier's F-updates with given per-round coefficient vectors (the optimiser is third
    party); returns per-round mean loss, line-search objective at the given alphas and final F."""
    n = X.shape[0]
    lid = O.LOSS_IDS[loss]
    dim = num_classes if loss == "logloss" else 1
    F = np.repeat(_f32(init_raw).reshape(dim, 1), n, axis=1).copy()
    losses, objs = [], []
    for i, ims in enumerate(models):
        sub = subspaces[i]
        h = np.stack([_f32(m.predict(X[:, sub])) for m in ims])
        objs.append(oracle.linesearch_eval(lid, 0.0, y, w, F, h, alphas[i]))
        oracle.update(F, h, np.asarray(alphas[i]) * lr)
        losses.append(oracle.mean_loss(lid, 0.0, dim, y, F))
    return {"train_loss": losses, "objective": objs, "F": F}


def ref_boosting_replay(oracle, X, y, w0, K, real, models):
    """BoostingClassifier.train weight recursion with given models."""
    n = X.shape[0]
    w = np.ones(n) if w0 is None else _f32(w0)
    sum_w = oracle.sum(w)
    errs, sums, est_w = [], [], []
    for m in models:
        if real:
            P = _f32(m.predictProbability(X).T)
            w, e, s = oracle.samme_r_update(K, y, w, sum_w, P)
            est_w.append(1.0)
        else:
            pred = _f32(m.predict(X))
            e = oracle.samme_error(y, w, sum_w, pred)
            beta = e / ((1 - e) * (K - 1))
            est_w.append(1.0 if beta == 0.0 else math.log(1.0 / beta))
            w, s = oracle.samme_update(y, w, sum_w, pred, beta)
        errs.append(e)
        sums.append(s)
        sum_w = s
    return {"estimatorError": errs, "sumWeights": sums, "weights": w, "estimatorWeights": est_w}
