"""GPU tests of the host-side mirror (fit/transform through the reference-shaped API -> C ABI -> CUDA),
modelled on the reference's suites: GBMRegressorSuite, GBMClassifierSuite, BoostingClassifierSuite,
Bagging*Suite.  Numeric parity is checked by REPLAYING each fit with the product's own base models
through the oracle-driven control flow (tests/ref_fit.py): both sides then see identical directions and
differ only in the hot-path arithmetic (tolerance 1e-5 relative, north_star)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import ref_fit

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


def _cpusmall():
    d = np.load(os.path.join(GOLD, "cpusmall.npz"))
    return d["X"].astype(np.float32), d["y"].astype(np.float64)


def _letter(n=6000):
    d = np.load(os.path.join(GOLD, "letter.npz"))
    return (d["X"][:n].astype(np.float64) / 7.5 - 1.0).astype(np.float32), d["y"][:n].astype(np.float64)


def _adult():
    d = np.load(os.path.join(GOLD, "adult8k.npz"))
    return np.unpackbits(d["X"], axis=1)[:, :123].astype(np.float32), d["y"].astype(np.float64)


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-30)


def test_gbm_regressor_cpusmall_config1(oracle):
    """BASELINE config 1: GBMRegressor on cpusmall, 20 rounds, squared loss."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import GBMRegressor
    X, y = _cpusmall()
    gbm = GBMRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=5)).setNumBaseLearners(20)
    assert gbm.uid.startswith("GBMRegressor2_")
    model = gbm.fit(DataFrame(features=X, label=y))
    assert model.uid.startswith("GBMRegressionModel_") and model.numModels == 20
    hist = model.trainingHistory
    rep = ref_fit.ref_gbm_regressor_fit(oracle, X, y, None, None, num_learners=20, models=model.models,
                                        subspaces=model.subspaces, init=model.init.prediction)
    for t in range(20):
        assert abs(hist[t]["alpha"] - rep["alpha"][t]) <= 3 * (1e-6 * abs(rep["alpha"][t]) + 1e-6)  # Brent tol
        assert _rel(hist[t]["trainLoss"], rep["train_loss"][t]) < RTOL
    pred = model.transform(DataFrame(features=X))["prediction"]
    np.testing.assert_allclose(pred, rep["F"], rtol=RTOL, atol=RTOL * float(np.abs(rep["F"]).mean()))
    assert model.predict(X[5]) == pytest.approx(rep["F"][5], rel=RTOL)
    # committed oracle golden of the same config (trees fitted on fp64 residuals there: loose)
    gold = json.load(open(os.path.join(GOLD, "gbm_cpusmall_oracle.json")))
    assert model.init.prediction == pytest.approx(gold["init"], rel=1e-6)
    assert hist[-1]["trainLoss"] == pytest.approx(gold["train_loss"][-1], rel=2e-2)
    # GBMRegressorSuite.scala:51-76: GBM(10 trees) beats a single tree
    single = DecisionTreeRegressor(maxDepth=5).fit(X, y)
    rmse = lambda p: float(np.sqrt(np.mean((p - y) ** 2)))
    assert rmse(pred) < rmse(single.predict(X))


@pytest.mark.parametrize("loss", ["absolute", "quantile", "huber"])
def test_gbm_regressor_other_losses_replay(oracle, loss):
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import GBMRegressor
    X, y = _cpusmall()
    X, y = X[:3000], y[:3000]
    gbm = (GBMRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=4)).setNumBaseLearners(5)
           .setLoss(loss).setAlpha(0.8))
    model = gbm.fit(DataFrame(features=X, label=y))
    pred = model.transform(DataFrame(features=X))["prediction"]
    # fp64 recomputation of GBMRegressionModel.predict from the fitted members
    F = np.full(len(y), np.float64(np.float32(model.init.prediction)))
    for wgt, m, s in zip(model.weights, model.models, model.subspaces):
        F = F + np.float64(np.float32(wgt)) * m.predict(X[:, s]).astype(np.float32).astype(np.float64)
    np.testing.assert_allclose(pred, F, rtol=RTOL, atol=RTOL * float(np.abs(F).mean()))
    if loss != "huber":  # huber's delta is re-estimated per round (exact-quantile restatement)
        rep = ref_fit.ref_gbm_regressor_fit(oracle, X, y, None, None, loss=loss, alpha_q=0.8, num_learners=5,
                                            models=model.models, subspaces=model.subspaces,
                                            init=model.init.prediction)
        for t in range(5):
            # flat/kinked objectives: compare the objective value reached, not the abscissa
            assert _rel(model.trainingHistory[t]["trainLoss"], rep["train_loss"][t]) < 1e-4
    losses = [h["trainLoss"] for h in model.trainingHistory]
    assert losses[-1] < losses[0]


def test_gbm_regressor_learning_rate_monotone_and_early_stop():
    """GBMRegressorSuite.scala:126-164 (lr 0.1 => metric monotone) and :78-124 (early-stop index rule)."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import GBMRegressor
    X, y = _cpusmall()
    rng = np.random.default_rng(0)
    val = rng.random(len(y)) < 0.3
    df = DataFrame(features=X, label=y, val=val)
    gbm = (GBMRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=4)).setNumBaseLearners(12)
           .setLearningRate(0.1).setValidationIndicatorCol("val").setNumRounds(3).setValidationTol(0.01))
    model = gbm.fit(df)
    hist = model.trainingHistory
    vl = [h["validationLoss"] for h in hist]
    assert all(b <= a * (1 + 1e-9) for a, b in zip(vl, vl[1:]))  # monotone with lr = 0.1
    # recompute the early-stop bookkeeping from the recorded validation losses (reference rule :457-464)
    init_loss = float(np.mean((y[val] - model.init.prediction) ** 2) / 2)
    best, v, i = init_loss, 0, 0
    while i < len(vl) and v < 3:
        if best - vl[i] < 0.01 * max(vl[i], 0.01):
            v += 1
        elif vl[i] < best:
            best, v = vl[i], 0
        i += 1
    assert len(hist) == i and model.numModels == i - v


def test_gbm_regressor_newton_weights_and_resident_features(oracle):
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import GBMRegressor
    X, y = _cpusmall()
    w = np.random.default_rng(1).random(len(y)) + 0.5
    df = DataFrame(features=X, label=y, weight=w)
    base = GBMRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=4)).setNumBaseLearners(4).setWeightCol("weight")
    m_host = base.copy().setUpdates("newton").fit(df)
    m_dev = base.copy().setUpdates("newton").setResidentFeatures(True).fit(df)  # trees evaluated on device
    for a, b in zip(m_host.trainingHistory, m_dev.trainingHistory):
        assert a["alpha"] == pytest.approx(b["alpha"], rel=1e-5, abs=1e-5)
        assert a["trainLoss"] == pytest.approx(b["trainLoss"], rel=1e-5)
    rep = ref_fit.ref_gbm_regressor_fit(oracle, X, y, w, None, num_learners=4, models=m_host.models,
                                        subspaces=m_host.subspaces, init=m_host.init.prediction)
    for t in range(4):
        assert _rel(m_host.trainingHistory[t]["trainLoss"], rep["train_loss"][t]) < RTOL


@pytest.mark.parametrize("loss,updates", [("bernoulli", "newton"), ("exponential", "gradient"),
                                          ("logloss", "gradient")])
def test_gbm_classifier_binary_replay(oracle, loss, updates):
    """GBMClassifierSuite.scala:89-146 (adult, bernoulli/exponential)."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.classification import GBMClassifier
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    X, y = _adult()
    gbm = (GBMClassifier().setBaseLearner(DecisionTreeRegressor(maxDepth=4)).setNumBaseLearners(4)
           .setLoss(loss).setUpdates(updates))
    model = gbm.fit(DataFrame(features=X, label=y))
    assert model.uid.startswith("GBMClassificationModel_") and model.numClasses == 2
    assert model.dim == (2 if loss == "logloss" else 1)
    alphas = [h["alpha"] for h in model.trainingHistory]
    rep = ref_fit.ref_gbm_classifier_replay(oracle, X, y, None, loss, 2, model.init, model.models,
                                            model.subspaces, alphas)
    for t, h in enumerate(model.trainingHistory):
        assert _rel(h["trainLoss"], rep["train_loss"][t]) < RTOL
    out = model.transform(DataFrame(features=X))
    raw, prob, pred = out["rawPrediction"], out["probability"], out["prediction"]
    F = rep["F"]
    ref_raw = np.stack([-F[0], F[0]], axis=1) if model.dim == 1 else F.T
    np.testing.assert_allclose(raw, ref_raw, rtol=RTOL, atol=RTOL * float(np.abs(ref_raw).mean()))
    ref_prob = oracle.gbm_raw2prob(O.LOSS_IDS[loss], ref_raw.T).T
    np.testing.assert_allclose(prob, ref_prob, rtol=1e-4, atol=1e-7)
    acc = float(np.mean(pred == y))
    assert acc > 0.75
    if loss == "exponential":  # reference quirk 6: probability decreases with F, prediction does not use it
        assert np.mean(np.argmax(prob, axis=1) == pred) < 0.5


def test_gbm_classifier_multiclass_letter(oracle):
    """GBMClassifierSuite.scala:51-87 (letter, newton, 3 learners)."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.classification import GBMClassifier
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    X, y = _letter(4000)
    gbm = (GBMClassifier().setBaseLearner(DecisionTreeRegressor(maxDepth=5)).setNumBaseLearners(3)
           .setUpdates("newton").setMaxIter(30))
    model = gbm.fit(DataFrame(features=X, label=y))
    assert model.numClasses == 26 and model.dim == 26
    alphas = [h["alpha"] for h in model.trainingHistory]
    rep = ref_fit.ref_gbm_classifier_replay(oracle, X, y, None, "logloss", 26, model.init, model.models,
                                            model.subspaces, alphas)
    for t, h in enumerate(model.trainingHistory):
        assert _rel(h["trainLoss"], rep["train_loss"][t]) < RTOL
    losses = [h["trainLoss"] for h in model.trainingHistory]
    assert losses == sorted(losses, reverse=True)
    out = model.transform(DataFrame(features=X))
    np.testing.assert_allclose(out["rawPrediction"], rep["F"].T, rtol=RTOL, atol=RTOL)
    np.testing.assert_allclose(out["probability"].sum(axis=1), 1.0, rtol=1e-5)
    assert np.mean(out["prediction"] == y) > 0.5


@pytest.mark.parametrize("algorithm", ["discrete", "real"])
def test_boosting_classifier_letter(oracle, algorithm):
    """BoostingClassifierSuite.scala:52-154: weight recursion parity (replay) + zero-sum raw predictions."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.classification import BoostingClassifier
    from spark_ensemble_b200.learners import DecisionTreeClassifier
    X, y = _letter(5000)
    K = 26
    bc = (BoostingClassifier().setBaseLearner(DecisionTreeClassifier(maxDepth=8)).setNumBaseLearners(6)
          .setAlgorithm(algorithm))
    model = bc.fit(DataFrame(features=X, label=y))
    assert model.uid.startswith("BoostingClassificationModel_") and model.numModels >= 1
    rep = ref_fit.ref_boosting_replay(oracle, X, y, None, K, algorithm == "real", model.models)
    for t in range(model.numModels):
        h = model.trainingHistory[t]
        assert h["estimatorError"] == pytest.approx(rep["estimatorError"][t], rel=1e-4, abs=1e-7)
        assert h["sumWeights"] == pytest.approx(rep["sumWeights"][t], rel=1e-4)
        assert model.weights[t] == pytest.approx(rep["estimatorWeights"][t], rel=1e-4)
    out = model.transform(DataFrame(features=X))
    raw = out["rawPrediction"]
    assert np.max(np.abs(raw.sum(axis=1))) <= 1e-5 * np.abs(raw).sum(axis=1).max()  # symmetric constraint
    np.testing.assert_allclose(out["probability"].sum(axis=1), 1.0, rtol=1e-5)
    # boosting improves on its first member (BoostingClassifierSuite.scala:52-91)
    first = np.mean(model.models[0].predict(X) == y)
    assert np.mean(out["prediction"] == y) >= first - 0.02


def test_bagging_models_aggregate_members():
    """Bagging*Suite: the ensemble prediction is the mean / vote of its members."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.classification import BaggingClassifier
    from spark_ensemble_b200.learners import DecisionTreeClassifier, DecisionTreeRegressor
    from spark_ensemble_b200.regression import BaggingRegressor
    X, y = _cpusmall()
    br = BaggingRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=6)).setNumBaseLearners(7).setSubspaceRatio(0.7)
    m = br.fit(DataFrame(features=X, label=y))
    pred = m.transform(DataFrame(features=X))["prediction"]
    members = np.stack([mm.predict(X[:, s]).astype(np.float32).astype(np.float64)
                        for mm, s in zip(m.models, m.subspaces)])
    np.testing.assert_allclose(pred, members.sum(axis=0) / 7, rtol=RTOL, atol=RTOL * np.abs(members).mean())
    assert all(len(s) <= 12 and list(s) == sorted(s) for s in m.subspaces)
    Xl, yl = _letter(3000)
    for strategy in ("hard", "soft"):
        bcl = (BaggingClassifier().setBaseLearner(DecisionTreeClassifier(maxDepth=8)).setNumBaseLearners(5)
               .setVotingStrategy(strategy))
        mc = bcl.fit(DataFrame(features=Xl, label=yl))
        out = mc.transform(DataFrame(features=Xl))
        if strategy == "hard":
            votes = np.stack([mm.predict(Xl[:, s]) for mm, s in zip(mc.models, mc.subspaces)])
            cnt = np.stack([(votes == c).sum(axis=0) for c in range(26)], axis=1).astype(np.float64)
            np.testing.assert_array_equal(out["rawPrediction"], cnt)
            np.testing.assert_allclose(out["probability"], cnt / 5, rtol=1e-6)
        else:
            P = np.stack([mm.predictProbability(Xl[:, s]) for mm, s in zip(mc.models, mc.subspaces)])
            np.testing.assert_allclose(out["rawPrediction"], P.sum(axis=0), rtol=RTOL, atol=1e-6)
        assert np.mean(out["prediction"] == yl) > 0.4


@pytest.mark.parametrize("loss_type,voting", [("exponential", "median"), ("linear", "mean"), ("squared", "median")])
def test_boosting_regressor_adaboost_r2(oracle, loss_type, voting):
    """BoostingRegressorSuite.scala:51-132: beats a single tree; weight recursion replayed through the oracle."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import BoostingRegressor
    X, y = _cpusmall()
    br = (BoostingRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=5)).setNumBaseLearners(8)
          .setLossType(loss_type).setVotingStrategy(voting))
    assert br.uid.startswith("BoostingRegressor_")
    model = br.fit(DataFrame(features=X, label=y))
    assert model.uid.startswith("BoostingRegressionModel_") and 1 <= model.numModels <= 8
    w = np.ones(len(y))
    sw = float(len(y))
    for t, m in enumerate(model.models):
        pred = m.predict(X).astype(np.float32).astype(np.float64)
        mx = oracle.r2_max_error(y.astype(np.float32).astype(np.float64), pred)
        e = oracle.r2_estimator_error(loss_type, y, pred, w, sw, mx)
        h = model.trainingHistory[t]
        assert h["maxError"] == pytest.approx(mx, rel=1e-6)
        assert h["estimatorError"] == pytest.approx(e, rel=1e-4)
        beta = e / (1 - e)
        assert model.weights[t] == pytest.approx(np.log(1 / beta), rel=1e-3)
        w, sw = oracle.r2_update(loss_type, y, pred, w, sw, mx, beta)
        assert h["sumWeights"] == pytest.approx(sw, rel=1e-4)
    pred = model.transform(DataFrame(features=X))["prediction"]
    P = np.stack([m.predict(X).astype(np.float32) for m in model.models])
    ref = (oracle.agg_weighted_median(P, model.weights) if voting == "median"
           else oracle.agg_weighted_mean(P, model.weights))
    np.testing.assert_allclose(pred, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).mean())
    rmse = lambda p: float(np.sqrt(np.mean((p - y) ** 2)))
    assert rmse(pred) < rmse(DecisionTreeRegressor(maxDepth=5).fit(X, y).predict(X)) * 1.05


def test_gbm_regressor_with_row_subsampling():
    """subsampleRatio < 1 / replacement: line search on the bag, update on all rows (reference quirk 4)."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import GBMRegressor
    X, y = _cpusmall()
    base = GBMRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=4)).setNumBaseLearners(6).setLearningRate(0.5)
    full = base.copy().fit(DataFrame(features=X, label=y))
    for ratio, repl in ((0.5, False), (1.0, True)):
        m = base.copy().setSubsampleRatio(ratio).setReplacement(repl).fit(DataFrame(features=X, label=y))
        assert m.numModels == 6
        losses = [h["trainLoss"] for h in m.trainingHistory]
        assert losses[-1] < losses[0]
        # a bagged fit still tracks the full fit's loss closely on this dataset
        assert losses[-1] < 1.5 * full.trainingHistory[-1]["trainLoss"]


@pytest.mark.parametrize("algorithm", ["discrete", "real"])
def test_boosting_classifier_resident_features(algorithm):
    """Base trees evaluated on device (se_tree_predict / se_tree_predict_multi) give the same fit as host predict."""
    from spark_ensemble_b200 import DataFrame
    from spark_ensemble_b200.classification import BoostingClassifier
    from spark_ensemble_b200.learners import DecisionTreeClassifier
    X, y = _letter(3000)
    base = BoostingClassifier().setBaseLearner(DecisionTreeClassifier(maxDepth=6)).setNumBaseLearners(4).setAlgorithm(algorithm)
    host = base.copy().fit(DataFrame(features=X, label=y))
    dev = base.copy().setResidentFeatures(True).fit(DataFrame(features=X, label=y))
    assert host.numModels == dev.numModels
    for a, b in zip(host.trainingHistory, dev.trainingHistory):
        assert a["estimatorError"] == pytest.approx(b["estimatorError"], rel=1e-6, abs=1e-9)
        assert a["sumWeights"] == pytest.approx(b["sumWeights"], rel=1e-6)


def test_gbm_regressor_fit_sharded_over_two_gpus():
    """Param `devices`: the SAME fit with the rows sharded over two contexts in one process (sharded.ShardedContext;
    cross-GPU sums inside the kernels) reproduces the single-GPU fit: same number of learners, weights within the
    optimiser tolerance, predictions within 1e-5."""
    import numpy as np
    from spark_ensemble_b200 import DataFrame, _native as N
    from spark_ensemble_b200.learners import DecisionTreeRegressor
    from spark_ensemble_b200.regression import GBMRegressor
    if N.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    rng = np.random.default_rng(3)
    n, d = 20011, 9
    X = rng.standard_normal((n, d)).astype(np.float32)
    y = (X[:, 0] * 2 - X[:, 1] + 0.3 * X[:, 2] * X[:, 3] + 0.1 * rng.standard_normal(n)).astype(np.float64)
    val = rng.random(n) < 0.2
    df = DataFrame(features=X, label=y, validation=val)
    fits = []
    for devices in ([], [0, 1]):
        for resident in (False, True):
            g = (GBMRegressor().setBaseLearner(DecisionTreeRegressor(maxDepth=4)).setNumBaseLearners(8)
                 .setValidationIndicatorCol("validation").setLearningRate(0.5))
            g.set("devices", devices)
            g.set("residentFeatures", resident)
            m = g.fit(df)
            fits.append((devices, resident, m))
    base = fits[0][2]
    pb = base.transform(df)["prediction"]
    for devices, resident, m in fits[1:]:
        assert m.numModels == base.numModels
        np.testing.assert_allclose(m.weights, base.weights, rtol=1e-5, atol=4e-6)
        np.testing.assert_allclose(m.transform(df)["prediction"], pb, rtol=1e-5, atol=1e-5 * float(np.abs(pb).max()))
    # GBMClassifier (LogLoss over 3 classes: per-class line-search gradients summed across the GPUs)
    from spark_ensemble_b200.classification import GBMClassifier
    yc = (np.digitize(y, np.quantile(y, [0.33, 0.66]))).astype(np.float64)
    dfc = DataFrame(features=X, label=yc)
    ms = []
    for devices in ([], [0, 1]):
        g = GBMClassifier().setBaseLearner(DecisionTreeRegressor(maxDepth=3)).setNumBaseLearners(4).setLoss("logloss")
        g.set("devices", devices)
        ms.append(g.fit(dfc))
    assert len(ms[0].weights) == len(ms[1].weights)
    for w0, w1 in zip(ms[0].weights, ms[1].weights):
        np.testing.assert_allclose(w1, w0, rtol=1e-4, atol=1e-5)  # L-BFGS-B iterates see sums that differ in the last bits
    np.testing.assert_array_equal(ms[0].transform(dfc)["prediction"], ms[1].transform(dfc)["prediction"])


def test_tree_arrays_must_form_a_tree(monkeypatch):
    """A child reached twice (cycle / shared node) would make the device walk spin: rejected with SE_ERR_ARG
    before anything is launched."""
    import numpy as np
    from spark_ensemble_b200 import _native as N
    from spark_ensemble_b200.context import Context
    with Context(0) as ctx:
        n, d = 64, 3
        ctx.alloc(N.SLOT_X, d, n)
        ctx.fill(N.SLOT_X, 0.5)
        ctx.alloc(N.SLOT_H, 1, n)
        good = {"feature": [0, -1, -1], "threshold": [0.0, 0, 0], "left": [1, 0, 0], "right": [2, 0, 0], "value": [0, 1.0, 2.0]}
        ctx.tree_predict(good, N.SLOT_H, 0)
        assert np.all(ctx.download(N.SLOT_H) == 2.0)
        for bad in ({**good, "left": [0, 0, 0]},                      # node 0 is its own child
                    {**good, "left": [1, 0, 0], "right": [1, 0, 0]},  # shared child
                    {"feature": [0, 1, -1], "threshold": [0.0, 0, 0], "left": [1, 0, 0], "right": [2, 2, 0], "value": [0, 0, 1.0]}):
            with pytest.raises(ValueError, match="not a tree"):
                ctx.tree_predict(bad, N.SLOT_H, 0)
