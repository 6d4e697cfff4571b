"""ds001907 dataset loader: turns the cache files the embedding builders write into `(df, masks)` for `train_pipeline` /
`evaluate_model` (reference: data/openneuro_ds001907.py:17-82).

Same contract: the manifest path comes from `PD_FUSION_DS001907_MANIFEST` when set (`:17-22`), `feature_mode` selects the cache
(`:35-62`), `diagnosis` is filled from `label` (`:64-66`), and the availability masks are clinical = datspect = 0, mri = "any mri_
column non-null" / "mri_mil is not None" (`:68-81`).  Nothing here touches the GPU: it is the consumer side of the on-disk formats.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, Tuple

import numpy as np
import pandas as pd

from .openneuro_features import (load_cnn_embeddings, load_resnet2d_embeddings, load_resnet2d_mil_embeddings,
                                 load_simple_features)
from .schema import TARGET_COL


def _resolve_manifest_path(config: Dict) -> Path:
    env_path = os.environ.get("PD_FUSION_DS001907_MANIFEST")
    if env_path:
        return Path(env_path)
    return Path(config.get("manifest_path", "data/processed/openneuro_ds001907_manifest.csv"))


def load_openneuro_ds001907(config: Dict) -> Tuple[pd.DataFrame, Dict[str, np.ndarray]]:
    manifest_path = _resolve_manifest_path(config)
    if not manifest_path.exists():
        raise FileNotFoundError(f"Manifest not found at {manifest_path}")
    mode = config.get("feature_mode", "simple")
    resnet_dir = Path(config.get("resnet2d_cache_dir", "data/processed/openneuro_ds001907/embeddings_resnet2d"))
    if mode == "simple":
        df = load_simple_features(manifest_path, Path(config.get("feature_cache_dir", "data/processed/openneuro_ds001907/features_simple")),
                                  config.get("feature_config", {}))
    elif mode == "cnn3d":
        df = load_cnn_embeddings(manifest_path, Path(config.get("embedding_cache_dir", "data/processed/openneuro_ds001907/embeddings_cnn3d")),
                                 config.get("cnn_config", {}))
    elif mode == "resnet2d":
        df = load_resnet2d_embeddings(manifest_path, resnet_dir, config.get("resnet2d_config", {}))
    elif mode == "resnet2d_mil":
        df = load_resnet2d_mil_embeddings(manifest_path, resnet_dir, config.get("resnet2d_config", {}))
    elif mode == "resnet2d_mil_ft":
        df = pd.read_csv(manifest_path)
        if "t1wbrain_path" not in df.columns:
            raise ValueError("Manifest is missing t1wbrain_path for MIL fine-tune.")
        df["mri_mil"] = df["t1wbrain_path"]              # the fine-tune model loads the volumes itself
    else:
        raise ValueError(f"Unknown feature_mode: {mode}")

    if "label" in df.columns and TARGET_COL not in df.columns:
        df[TARGET_COL] = df["label"].astype(int)
    mri_cols = [c for c in df.columns if c.startswith("mri_")]       # (includes "mri_mil", as in the reference)
    if mri_cols:
        mri_mask = (~df[mri_cols].isna().all(axis=1)).astype(int).values
    elif "mri_mil" in df.columns:
        mri_mask = df["mri_mil"].apply(lambda x: int(x is not None)).values
    else:
        raise ValueError("No mri_ feature columns or mri_mil found in ds001907 dataframe.")
    n = len(df)
    return df, {"clinical": np.zeros(n, dtype=int), "datspect": np.zeros(n, dtype=int), "mri": mri_mask}
