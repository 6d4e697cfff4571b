"""Scenario -> per-subject modality masks (reference: data/missingness.py:7-66).

Integer work done on the host with the reference's exact draws from the GLOBAL numpy RNG (one `rand(N)` per
dropped modality in list order; one `choice` per subject for `type: random`), so that after the same
`np.random.seed(k)` both implementations consume identical masks.  The masks are then uploaded once as
uint8 [S, N, M] and every scenario is evaluated in one launch (pdf_moddrop_sweep / pdf_moe_sweep).
"""
from __future__ import annotations

import logging
from typing import Dict, List, Sequence

import numpy as np

from .schema import MODALITIES

_log = logging.getLogger("pd_fusion")


def apply_missingness_scenario(df, scenario: Dict, maskdict: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    masks = {name: np.array(vec, copy=True) for name, vec in maskdict.items()}
    label = scenario.get("name", "unnamed")
    for mod in scenario.get("drop_modalities", []) if "drop_modalities" in scenario else []:
        if mod not in masks:
            _log.info("[missingness] scenario '%s': modality '%s' not found in masks; no-op.", label, mod)
            continue
        if "drop_rate" in scenario:
            rate = float(scenario.get("drop_rate", 0.0))
            if rate <= 0:
                continue
            hit = np.random.rand(len(masks[mod])) < rate
            if np.all(masks[mod] == 0):
                _log.info("[missingness] scenario '%s': modality '%s' already absent; no-op.", label, mod)
            masks[mod][hit] = 0
        else:
            if np.all(masks[mod] == 0):
                _log.info("[missingness] scenario '%s': modality '%s' already absent; no-op.", label, mod)
            masks[mod] = np.zeros_like(masks[mod])
    if scenario.get("type") == "random":
        k = scenario.get("n_drop", 1)
        names = list(masks.keys()) if masks else MODALITIES
        for row in range(len(df)):
            present = [m for m in names if m in masks and masks[m][row] == 1]
            if not present:
                continue
            for mod in np.random.choice(present, size=min(k, len(present)), replace=False):
                masks[mod][row] = 0
    return masks


def get_modality_mask_matrix(maskdict: Dict[str, np.ndarray]) -> np.ndarray:
    if not maskdict:
        raise ValueError("maskdict is empty")
    template = next(iter(maskdict.values()))
    return np.stack([maskdict[m] if m in maskdict else np.zeros_like(template) for m in MODALITIES], axis=1)


def scenario_mask_tensor(df, scenarios: Sequence[Dict], maskdict: Dict[str, np.ndarray], order: Sequence[str]):
    """All scenarios in config order -> (uint8 [S, N, len(order)], list of per-scenario mask dicts)."""
    per_scenario: List[Dict[str, np.ndarray]] = [apply_missingness_scenario(df, sc, maskdict) for sc in scenarios]
    n = len(df)
    out = np.zeros((len(scenarios), n, len(order)), dtype=np.uint8)
    for s, cur in enumerate(per_scenario):
        for j, mod in enumerate(order):
            out[s, :, j] = (np.asarray(cur[mod]) != 0) if mod in cur else 1
    return out, per_scenario
