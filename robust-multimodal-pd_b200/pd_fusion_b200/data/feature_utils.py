"""Column -> modality resolution (reference: data/feature_utils.py:6-61). Host-side bookkeeping only; the
mask multiply itself runs inside pdf_moddrop_sweep / pdf_moe_sweep on the device."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .schema import MODALITIES, MODALITY_FEATURES


def get_modality_feature_cols(df, modality: str) -> List[str]:
    pref = [c for c in df.columns if c.startswith(f"{modality}_")]
    return pref if pref else [c for c in MODALITY_FEATURES.get(modality, []) if c in df.columns]


def get_all_feature_cols(df) -> List[str]:
    cols: List[str] = []
    for mod in MODALITIES:
        cols.extend(get_modality_feature_cols(df, mod))
    return cols


def get_feature_slices(feature_cols: List[str]) -> Dict[str, List[int]]:
    out: Dict[str, List[int]] = {m: [] for m in MODALITIES}
    for i, col in enumerate(feature_cols):
        owner = next((m for m in MODALITIES if col.startswith(f"{m}_")), None)
        if owner is None:
            owner = next((m for m, names in MODALITY_FEATURES.items() if col in names), None)
        if owner is not None:
            out[owner].append(i)
    return out


def apply_masks_to_matrix(X: np.ndarray, masks: Dict[str, np.ndarray], feature_cols: List[str]) -> np.ndarray:
    """API parity helper (the evaluation path does not call it: the sweep kernels apply the masks on the device)."""
    Xm = np.array(X, copy=True)
    for mod, idxs in get_feature_slices(feature_cols).items():
        if idxs and mod in masks:
            Xm[:, idxs] = Xm[:, idxs] * np.asarray(masks[mod]).reshape(-1, 1)
    return Xm
