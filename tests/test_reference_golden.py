"""Replays golden vectors dumped FROM THE REAL REFERENCE (bench/dump_reference_golden.scala, run wherever a JVM + Spark
+ the spark-ensemble jar exist) through this repository's hot path.  tests/golden/reference_*.json are consumed
automatically when present; without them (this image has no JVM) the tests are skipped and parity stays pinned by the
published third-party vectors (tests/test_thirdparty_golden.py) and the reference's portable properties only.

Per round t the dump carries the third-party base model's training predictions h_t; everything downstream — the Brent
line search (weight_t = learningRate * alpha_t), the F update, the loss — is recomputed here and compared:
  * weight within the optimiser tolerance (tol = 1e-6 relative + absolute, the reference's own stopping rule),
  * train loss and prediction checksum within 1e-5 relative (north_star)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"reference_c1.json": ("squared", "cpusmall.npz"), "reference_c3.json": ("bernoulli", "adult8k.npz")}
RTOL = 1e-5


def _load(name):
    p = os.path.join(HERE, "golden", name)
    if not os.path.exists(p):
        pytest.skip(f"{name} absent: no JVM in this image — run bench/dump_reference_golden.scala on a Spark box and "
                    f"copy its output to tests/golden/ to pin parity to the reference itself")
    g = json.load(open(p))
    loss, data = CASES[name]
    d = np.load(os.path.join(HERE, "golden", data))
    y = d["y"].astype(np.float64)
    assert y.shape[0] == g["n"], "the dump must cover the fixture's rows in file order"
    return g, loss, y


def _check_round(t, rd, lr, tol, alpha, loss_mean, F):
    w_ref = rd["weight"]
    assert abs(lr * alpha - w_ref) <= lr * (4 * tol * max(abs(alpha), 1.0) + 4 * tol), (t, lr * alpha, w_ref)
    assert abs(loss_mean - rd["train_loss_after"]) <= RTOL * abs(rd["train_loss_after"]) + 8 * tol * abs(rd["train_loss_after"]), t
    cs = rd["prediction_checksum"]
    assert abs(float(np.sum(F)) - cs[0]) <= RTOL * max(abs(cs[0]), float(np.sum(np.abs(F))) * 1e-3), t
    assert abs(float(np.sum(F * F)) - cs[1]) <= 2 * RTOL * cs[1], t


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_replays_reference_dump(oracle, name):
    """CPU: the oracle's line search / update / loss against the reference's own per-round numbers."""
    g, loss, y = _load(name)
    lid = O.LOSS_IDS[loss]
    lr, tol, max_iter = g["config"]["learningRate"], g["config"]["tol"], g["config"]["maxIter"]
    F = np.full((1, y.shape[0]), g["init"][0])
    for t, rd in enumerate(g["rounds"]):
        h = np.asarray(rd["direction"], dtype=np.float64).reshape(1, -1)
        f = lambda a: oracle.linesearch_eval(lid, 0.0, y, None, F, h, [a])[0]
        alpha, _, st = oracle.brent(f, 0.0, 100.0, 1.0, tol, tol, max_iter)
        assert st == 0
        # follow the REFERENCE's weight so that rounding of the optimiser does not compound over rounds
        oracle.update(F, h, [rd["weight"]])
        _check_round(t, rd, lr, tol, alpha, oracle.mean_loss(lid, 0.0, 1, y, F), F[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_replays_reference_dump(name):
    """GPU: the same replay through the C ABI (fp32 device state)."""
    from spark_ensemble_b200 import _native as N
    from spark_ensemble_b200.context import Context
    g, loss, y = _load(name)
    lr, tol, max_iter = g["config"]["learningRate"], g["config"]["tol"], g["config"]["maxIter"]
    n = y.shape[0]
    with Context(0) as ctx:
        ctx.gbm_configure(n, 0, 1, loss, 0.0, False)
        ctx.upload(N.SLOT_Y, y.astype(np.float32))
        ctx.fill(N.SLOT_F, g["init"][0])
        for t, rd in enumerate(g["rounds"]):
            ctx.upload(N.SLOT_H, np.asarray(rd["direction"], dtype=np.float32))
            alpha, _, _ = ctx.gbm_linesearch_brent(0.0, 100.0, 1.0, tol, tol, max_iter)
            ls, _ = ctx.gbm_update([rd["weight"]], residual=True, loss=True)
            _check_round(t, rd, lr, tol, alpha, ls / n, ctx.download(N.SLOT_F).astype(np.float64))
