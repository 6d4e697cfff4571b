"""Deterministic synthetic inputs for parity tests and bench.py.

Nothing here is on the product path: it only manufactures inputs of the shape the
reference consumes (T1 volumes ``f32[256,256,176]`` C-order, tabular modality
blocks, manifests).  The reference has no generator for volumes; its tabular
generator is ``pd_fusion/data/ppmi_loader.py:146-178`` (NaN-bearing, hence not
reused — see SURVEY.md §0).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, Tuple

import numpy as np

FULL_SHAPE = (256, 256, 176)


def synthetic_volume(index: int, shape: Tuple[int, int, int] = FULL_SHAPE, bad_fraction: float = 1e-5) -> np.ndarray:
    """Ellipsoid "brain" with Gamma(4, 100) intensities, exact zeros outside and a
    sprinkle of NaN/+Inf/-Inf voxels (exercises the nan_to_num scrub).

    Seeded by ``default_rng(1000 + index)`` so host, oracle and GPU box see identical bytes.
    """
    rng = np.random.default_rng(1000 + int(index))
    rx, ry, rz = 0.70 + 0.10 * rng.random(), 0.80 + 0.10 * rng.random(), 0.75 + 0.10 * rng.random()
    X, Y, Z = shape
    x = np.linspace(-1.0, 1.0, X, dtype=np.float32)[:, None, None]
    y = np.linspace(-1.0, 1.0, Y, dtype=np.float32)[None, :, None]
    z = np.linspace(-1.0, 1.0, Z, dtype=np.float32)[None, None, :]
    inside = (x / np.float32(rx)) ** 2 + (y / np.float32(ry)) ** 2 + (z / np.float32(rz)) ** 2 < 1.0
    vol = rng.gamma(4.0, 100.0, size=shape).astype(np.float32)
    vol *= inside
    n_bad = int(round(bad_fraction * vol.size))
    if n_bad > 0:
        flat = vol.reshape(-1)
        pos = rng.integers(0, flat.size, size=n_bad)
        kinds = rng.integers(0, 3, size=n_bad)
        flat[pos[kinds == 0]] = np.nan
        flat[pos[kinds == 1]] = np.inf
        flat[pos[kinds == 2]] = -np.inf
    return vol


def write_synthetic_manifest(out_dir: Path, n_subjects: int, shape=FULL_SHAPE, start: int = 0) -> Path:
    """Writes ``sub-XXXXX.npy`` volumes plus a manifest CSV with the reference's columns
    (``subject_id, session, label, t1wbrain_path`` — data/openneuro_features.py:226-274)."""
    import pandas as pd

    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    rows = []
    for i in range(start, start + n_subjects):
        p = out_dir / f"sub-{i:05d}.npy"
        np.save(p, synthetic_volume(i, shape))
        rows.append({"subject_id": f"sub-{i:05d}", "session": 1, "label": i % 2, "t1wbrain_path": str(p)})
    manifest = out_dir / "manifest.csv"
    pd.DataFrame(rows).to_csv(manifest, index=False)
    return manifest


def synthetic_table(n: int, dims: Dict[str, int], seed: int = 42, mask_seed: int = 7,
                    present_rate: float = 0.85):
    """NaN-free tabular features per modality + availability masks + labels with signal.

    Column naming follows the reference's prefixed convention ``<modality>_f<j>``
    (data/ppmi_loader.py:157) so ``get_modality_feature_cols`` resolves them.
    """
    import pandas as pd

    rng = np.random.default_rng(seed)
    data = {"patno": np.arange(n)}
    feats = {}
    for mod, d in dims.items():
        f = rng.standard_normal((n, d))
        feats[mod] = f
        for j in range(d):
            data[f"{mod}_f{j}"] = f[:, j]
    score = np.zeros(n)
    for k, mod in enumerate(dims):
        if dims[mod] > 0:
            score = score + (1.0 if k % 2 == 0 else -1.0) * feats[mod][:, 0]
    data["diagnosis"] = (score > 0).astype(int)
    mrng = np.random.default_rng(mask_seed)
    masks = {mod: (mrng.random(n) < present_rate).astype(int) for mod in dims}
    return pd.DataFrame(data), masks


def iter_volumes(indices: Iterable[int], shape=FULL_SHAPE):
    for i in indices:
        yield synthetic_volume(i, shape)
