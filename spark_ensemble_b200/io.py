"""Data formats either side of the path: LIBSVM text -> dense fp32 row blocks -> the column-major
feature matrix in HBM.

The reference reads its fixtures with Spark's `libsvm` source (`data/*/*.svm`, e.g.
test/regression/GBMRegressorSuite.scala: `spark.read.format("libsvm").load(...)`), which yields 1-based sparse vectors
densified by the learners.  Here the file is parsed in row blocks on the host and every block goes through
`se_upload_rowmajor` (pinned double-buffered staging + transpose kernel), so a file larger than host memory never has
to be resident: peak host memory is one block.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import numpy as np


def iter_libsvm_dense(path: str, num_features: int, block_rows: int = 1 << 16,
                      zero_based: bool = False) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
    """Yield (X_block [rows, num_features] float32 row-major, y_block [rows] float64) from a LIBSVM file.

    Index semantics follow Spark's reader: indices are 1-based (`zero_based=False`), must be ascending within a line
    and may not exceed `num_features`; absent entries are 0; `#` starts a comment."""
    X = np.zeros((block_rows, num_features), dtype=np.float32)
    y = np.empty(block_rows, dtype=np.float64)
    r = 0
    off = 0 if zero_based else 1
    with open(path, "r") as fh:
        for lineno, line in enumerate(fh, 1):
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            parts = line.split()
            y[r] = float(parts[0])
            prev = -1
            for tok in parts[1:]:
                i_s, v_s = tok.split(":", 1)
                i = int(i_s) - off
                if i <= prev:
                    raise ValueError(f"{path}:{lineno}: indices must be ascending (got {i + off} after {prev + off})")
                if i < 0 or i >= num_features:
                    raise ValueError(f"{path}:{lineno}: feature index {i + off} outside 1..{num_features}")
                X[r, i] = float(v_s)
                prev = i
            r += 1
            if r == block_rows:
                yield X, y
                X = np.zeros((block_rows, num_features), dtype=np.float32)
                y = np.empty(block_rows, dtype=np.float64)
                r = 0
    if r:
        yield X[:r], y[:r]


def count_libsvm_rows(path: str) -> int:
    n = 0
    with open(path, "r") as fh:
        for line in fh:
            if line.split("#", 1)[0].strip():
                n += 1
    return n


def load_libsvm_to_device(ctx, slot: int, path: str, num_features: int, block_rows: int = 1 << 16,
                          num_rows: Optional[int] = None) -> np.ndarray:
    """Stream a LIBSVM file into the column-major [num_features][n] slot (SE_SLOT_X / SE_SLOT_VX) block by block;
    returns the labels.  `num_rows` saves the counting pass when the caller knows it."""
    n = count_libsvm_rows(path) if num_rows is None else int(num_rows)
    ctx.alloc(slot, num_features, n)
    labels = np.empty(n, dtype=np.float64)
    row = 0
    for Xb, yb in iter_libsvm_dense(path, num_features, block_rows):
        if row + len(yb) > n:
            raise ValueError(f"{path} has more than the announced {n} rows")
        ctx.upload_rowmajor(slot, Xb, row)
        labels[row:row + len(yb)] = yb
        row += len(yb)
    if row != n:
        raise ValueError(f"{path} has {row} rows, expected {n}")
    return labels
