"""The missingness sweep ("eval_missingness"; reference: evaluation/evaluate.py:11-138).

Same signatures and the same dispatch on the TYPE of `prep_info` (("mil", col) tuple -> MIL; dict -> MoE;
3-tuple -> matrix models).  What changes is the execution: every scenario's masks are drawn first (host, the
reference's RNG order), uploaded once as uint8 [S, N, M], and the model is evaluated under ALL of them in one
launch; tabular preprocessing runs once instead of once per scenario (evaluate.py:58,67).
Models that do not expose a sweep entry point (e.g. a CalibratedModel wrapper) fall back to the reference's
per-scenario `predict_proba` calls -- each of which still runs on the device.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import pandas as pd

from ..data.feature_utils import apply_masks_to_matrix
from ..data.missingness import apply_missingness_scenario, get_modality_mask_matrix, scenario_mask_tensor
from ..data.preprocess import preprocess_features
from ..data.schema import MODALITIES, TARGET_COL
from ..utils.metrics import compute_metrics


def _kind(prep_info) -> str:
    if isinstance(prep_info, tuple) and len(prep_info) >= 2 and prep_info[0] == "mil":
        return "mil"
    return "moe" if isinstance(prep_info, dict) else "matrix"


def _sweep_probs(model, df_test, mask_test, prep_info, scenarios: List[Dict]) -> Tuple[np.ndarray, List[Dict[str, np.ndarray]]]:
    """Probabilities [S, N] for all scenarios + the per-scenario mask dicts (drawn in scenario order)."""
    kind = _kind(prep_info)
    if kind == "mil":
        _, per = scenario_mask_tensor(df_test, scenarios, mask_test, list(mask_test.keys()))
        bags = df_test[prep_info[1]].tolist()
        if hasattr(model, "predict_proba_sweep") and all(b is None or isinstance(b, np.ndarray) for b in bags):
            # one upload, one projection, one pooling pass for ALL scenarios (pdf_mil_sweep): the scenario only decides where
            # missing_prob replaces the bag's probability
            n = len(bags)
            mri = np.stack([np.asarray(cur["mri"]) if "mri" in cur else np.ones(n, dtype=int) for cur in per])
            return model.predict_proba_sweep(bags, mri), per
        probs = []
        for cur in per:   # the bag tensor is padded once per call; a missing bag never reaches the device
            b = [bag if m == 1 else None for bag, m in zip(bags, cur["mri"])] if "mri" in cur else bags
            probs.append(model.predict_proba(b, masks=cur))
        return np.stack(probs), per
    if kind == "moe":
        mods = list(prep_info.keys())
        if mods != sorted(mods):
            # MoENet orders its experts by sorted(modality) while the reference's evaluate feeds the router in caller order
            # (evaluation/evaluate.py:54-63): an unsorted prep_info would silently pair router outputs with the wrong experts
            raise ValueError(f"MoE prep_info modalities must be in sorted order, got {mods}")
        masks, per = scenario_mask_tensor(df_test, scenarios, mask_test, mods)
        X = {m: preprocess_features(df_test, prep_info[m][2], prep_info[m][0], prep_info[m][1])[0] for m in mods}
        if hasattr(model, "predict_proba_sweep"):
            return model.predict_proba_sweep(X, masks, mods), per
        import torch
        probs = []
        for s, cur in enumerate(per):
            Xd = {m: torch.FloatTensor(X[m] * cur[m].reshape(-1, 1) if m in cur else X[m]) for m in mods}
            probs.append(model.predict_proba(Xd, torch.FloatTensor(np.stack([cur[m] for m in mods], axis=1))))
        return np.stack(probs), per
    imputer, scaler, feature_cols = prep_info
    X, _, _ = preprocess_features(df_test, feature_cols, imputer, scaler)
    masks, per = scenario_mask_tensor(df_test, scenarios, mask_test, MODALITIES)
    if hasattr(model, "predict_proba_sweep") and not hasattr(model, "mask_dim"):
        return model.predict_proba_sweep(X, masks, MODALITIES), per
    probs = []
    for cur in per:
        Xm = apply_masks_to_matrix(X, cur, feature_cols)
        probs.append(model.predict_proba(Xm, masks=get_modality_mask_matrix(cur) if hasattr(model, "mask_dim") else cur))
    return np.stack(probs), per


def _with_groups(metrics: Dict, df_test, y_true, y_prob, group_col):
    if group_col and group_col in df_test.columns:
        t = pd.DataFrame({"group": df_test[group_col].values, "y_true": y_true, "y_prob": y_prob})
        g = compute_metrics(t.groupby("group")["y_true"].first().values, t.groupby("group")["y_prob"].mean().values)
        metrics.update({f"subject_{k}": v for k, v in g.items()})
    return metrics


def evaluate_model(model, df_test, mask_test, prep_info, config) -> Dict[str, Dict[str, float]]:
    scenarios = config.get("scenarios", [{"name": "baseline", "drop_modalities": []}])
    group_col = config.get("group_col")
    y_true = df_test[TARGET_COL].values
    probs, _ = _sweep_probs(model, df_test, mask_test, prep_info, scenarios)
    return {sc["name"]: _with_groups(compute_metrics(y_true, np.asarray(probs[s])), df_test, y_true, np.asarray(probs[s]), group_col)
            for s, sc in enumerate(scenarios)}


def predict_proba_for_scenario(model, df_test, mask_test, prep_info, scenario):
    """(y_true, y_prob) for a single scenario."""
    probs, _ = _sweep_probs(model, df_test, mask_test, prep_info, [scenario])
    return df_test[TARGET_COL].values, np.asarray(probs[0])


def compute_risk_coverage(y_true, y_prob, masks):
    """Risk-coverage curve by descending confidence (reference: evaluate.py:140-169)."""
    conf = np.maximum(y_prob, 1 - y_prob)
    order = np.argsort(conf)[::-1]
    hits = ((y_prob >= 0.5).astype(int) == y_true).astype(int)[order]
    n = len(y_true)
    k = np.arange(1, n + 1)
    return {"coverage": k / n, "risk": 1 - np.cumsum(hits) / k}
