"""Base learners for the host-side mirror.

In the reference the base learner is ANY third-party Spark ML Predictor (ensemble/ensembleParams.scala:
64-81) — its fit/predict are not part of the hot path; only its per-row outputs are the hot path's
inputs.  There is no Spark here, so scikit-learn estimators stand in for Spark's DecisionTree*/Linear*
(also third party).  Fitted trees / linear models expose array forms so the product can evaluate them
ON DEVICE over the column-major feature matrix (se_tree_predict / se_linear_predict).
"""
from __future__ import annotations

import numpy as np


def _tree_arrays(tree) -> dict:
    t = tree.tree_
    leaf = t.children_left < 0
    thr64 = t.threshold.astype(np.float64)
    thr32 = thr64.astype(np.float32)
    # largest fp32 <= fp64 threshold, so `x <= thr32` equals `x <= thr64` for every fp32 x
    up = thr32.astype(np.float64) > thr64
    thr32[up] = np.nextafter(thr32[up], np.float32(-np.inf))
    return {
        "feature": np.where(leaf, -1, t.feature).astype(np.int32),
        "threshold": np.where(leaf, 0.0, thr32).astype(np.float32),
        "left": np.maximum(t.children_left, 0).astype(np.int32),
        "right": np.maximum(t.children_right, 0).astype(np.int32),
    }


class _Model:
    def tree_arrays(self):
        return None

    def linear_arrays(self):
        return None


class DecisionTreeRegressionModel(_Model):
    def __init__(self, sk):
        self.sk = sk

    def predict(self, X) -> np.ndarray:
        return self.sk.predict(np.asarray(X, dtype=np.float32)).astype(np.float64)

    def tree_arrays(self):
        d = _tree_arrays(self.sk)
        d["value"] = self.sk.tree_.value.reshape(-1).astype(np.float32)
        return d


class DecisionTreeRegressor:
    """Stand-in for org.apache.spark.ml.regression.DecisionTreeRegressor (maxDepth default 5)."""

    def __init__(self, maxDepth: int = 5, minInstancesPerNode: int = 1, seed: int = 0):
        self.maxDepth, self.minInstancesPerNode, self.seed = maxDepth, minInstancesPerNode, seed

    def copy(self, extra=None):
        return DecisionTreeRegressor(self.maxDepth, self.minInstancesPerNode, self.seed)

    def fit(self, X, y, w=None) -> DecisionTreeRegressionModel:
        from sklearn.tree import DecisionTreeRegressor as SK
        sk = SK(max_depth=self.maxDepth, min_samples_leaf=self.minInstancesPerNode, random_state=self.seed)
        sk.fit(np.asarray(X, dtype=np.float32), np.asarray(y, dtype=np.float64),
               sample_weight=None if w is None else np.asarray(w, dtype=np.float64))
        return DecisionTreeRegressionModel(sk)


class DecisionTreeClassificationModel(_Model):
    def __init__(self, sk, num_classes: int):
        self.sk, self.numClasses = sk, num_classes

    def predictProbability(self, X) -> np.ndarray:
        p = self.sk.predict_proba(np.asarray(X, dtype=np.float32))
        out = np.zeros((p.shape[0], self.numClasses))
        out[:, self.sk.classes_.astype(int)] = p
        return out

    def predict(self, X) -> np.ndarray:
        return np.argmax(self.predictProbability(X), axis=1).astype(np.float64)

    def tree_arrays(self):
        """Array form with per-node class-probability vectors ("values" [n_nodes, K]) and the predicted label
        ("value" [n_nodes]) so both predictProbability and predict can be evaluated on device."""
        d = _tree_arrays(self.sk)
        v = self.sk.tree_.value[:, 0, :].astype(np.float64)
        v = v / np.maximum(v.sum(axis=1, keepdims=True), 1e-300)
        full = np.zeros((v.shape[0], self.numClasses))
        full[:, self.sk.classes_.astype(int)] = v
        d["values"] = full.astype(np.float32)
        d["value"] = np.argmax(full, axis=1).astype(np.float32)
        return d


class DecisionTreeClassifier:
    """Stand-in for org.apache.spark.ml.classification.DecisionTreeClassifier."""

    def __init__(self, maxDepth: int = 5, seed: int = 0):
        self.maxDepth, self.seed = maxDepth, seed
        self.numClasses = None

    def copy(self, extra=None):
        return DecisionTreeClassifier(self.maxDepth, self.seed)

    def fit(self, X, y, w=None, num_classes: int | None = None) -> DecisionTreeClassificationModel:
        from sklearn.tree import DecisionTreeClassifier as SK
        sk = SK(max_depth=self.maxDepth, random_state=self.seed)
        yi = np.asarray(y).astype(int)
        sk.fit(np.asarray(X, dtype=np.float32), yi, sample_weight=None if w is None else np.asarray(w, dtype=np.float64))
        return DecisionTreeClassificationModel(sk, int(num_classes or (yi.max() + 1)))


class LinearRegressionModel(_Model):
    def __init__(self, coef, intercept):
        self.coefficients = np.asarray(coef, dtype=np.float64)
        self.intercept = float(intercept)

    def predict(self, X) -> np.ndarray:
        return np.asarray(X, dtype=np.float64) @ self.coefficients + self.intercept

    def linear_arrays(self):
        return {"coef": self.coefficients.astype(np.float32), "intercept": np.float32(self.intercept)}


class LinearRegression:
    """Stand-in for org.apache.spark.ml.regression.LinearRegression (ridge via regParam)."""

    def __init__(self, regParam: float = 0.0):
        self.regParam = regParam

    def copy(self, extra=None):
        return LinearRegression(self.regParam)

    def fit(self, X, y, w=None) -> LinearRegressionModel:
        from sklearn.linear_model import Ridge
        sk = Ridge(alpha=max(self.regParam, 1e-12))
        sk.fit(np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64),
               sample_weight=None if w is None else np.asarray(w, dtype=np.float64))
        return LinearRegressionModel(sk.coef_, sk.intercept_)
