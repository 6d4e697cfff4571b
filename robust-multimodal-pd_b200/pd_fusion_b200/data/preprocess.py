"""NaN-aware robust scaling of tabular features (reference: data/preprocess.py:5-70).  Out of the hot path
(runs once per fold on <= 1e4 x 1e3 floats); kept because `prep_info` carries the fitted scaler into
`evaluate_model`."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


class NaNRobustScaler:
    def __init__(self):
        self.medians = None
        self.iqrs = None

    def fit(self, X: np.ndarray):
        self.medians = np.nanmedian(X, axis=0)
        self.iqrs = np.nanpercentile(X, 75, axis=0) - np.nanpercentile(X, 25, axis=0)
        self.iqrs[self.iqrs == 0] = 1.0
        return self

    def transform(self, X: np.ndarray) -> np.ndarray:
        if self.medians is None:
            raise ValueError("Scaler not fitted")
        return (X - self.medians) / self.iqrs


def preprocess_features(df, feature_cols: List[str], imputer=None, scaler=None, strategy: str = "robust") -> Tuple[np.ndarray, object, object]:
    if not [c for c in feature_cols if c in df.columns]:
        return np.full((len(df), len(feature_cols)), np.nan), imputer, scaler
    frame = df[feature_cols].copy()
    X = frame[feature_cols].values
    if scaler is None:
        scaler = NaNRobustScaler().fit(X)
    return scaler.transform(X), None, scaler
