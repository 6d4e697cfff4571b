"""Host side of test-time augmentation.

The reference draws, per subject and per pass, from ONE numpy Generator seeded with `abs(hash(str(subject_id))) % 2**32`
(data/openneuro_features.py:231-247, scripts/build_resnet2d_mil_embeddings.py:120-137), in this order:
    angle ~ U(-rot, rot); translate ~ U(-t, t, size=2) * [H, W]; scale = 1 + U(-s, s); shift ~ U(-s, s);
    noise ~ N(0, sigma, size=[n, H, W])   (only when sigma > 0)
Those draws are reproduced here with the same calls and handed to the device as data (`pdf_tta_params`, noise field); the
affine resampling / intensity / noise / clip arithmetic runs in libpdfusion_b200.so (`pdf_tta_augment`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from .. import _lib


@dataclass
class TtaDraw:
    rot: np.ndarray               # [2, 2] float64
    offset: np.ndarray            # [2] float64
    scale: float
    shift: float
    noise: Optional[np.ndarray]   # [n, H, W] float64 or None


def subject_seed(subject_id) -> int:
    """The reference's per-subject seed.  `hash` of a str is salted per process unless PYTHONHASHSEED is set -- exactly as in
    the reference, two runs only agree under a fixed PYTHONHASHSEED."""
    return abs(hash(str(subject_id))) % (2 ** 32)


def affine_matrix(shape_hw: Sequence[int], angle_deg: float, translate: np.ndarray):
    """matrix and offset `_apply_affine_2d` passes to scipy (openneuro_features.py:166-170)."""
    theta = np.deg2rad(angle_deg)
    rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    center = np.array(shape_hw) / 2.0
    offset = center - rot @ center + translate
    return rot, offset


def draw_passes(seed: int, n_pass: int, n_slices: int, shape_hw: Sequence[int], cfg: Dict) -> List[TtaDraw]:
    rng = np.random.default_rng(seed)
    H, W = int(shape_hw[0]), int(shape_hw[1])
    out = []
    for _ in range(n_pass):
        angle = rng.uniform(-cfg["max_rotation_deg"], cfg["max_rotation_deg"])
        translate = rng.uniform(-cfg["max_translation"], cfg["max_translation"], size=2)
        translate = translate * np.array([H, W])
        rot, offset = affine_matrix((H, W), angle, translate)
        scale = 1.0 + rng.uniform(-cfg["intensity_scale"], cfg["intensity_scale"])
        shift = rng.uniform(-cfg["intensity_shift"], cfg["intensity_shift"])
        noise = rng.normal(0.0, cfg["noise_std"], size=(n_slices, H, W)) if cfg["noise_std"] > 0 else None
        out.append(TtaDraw(rot, offset, float(scale), float(shift), noise))
    return out


def params_array(draws: Sequence[TtaDraw]):
    """ctypes array of pdf_tta_params, one per subject of the batch (one pass)."""
    arr = (_lib.TtaParams * len(draws))()
    for i, d in enumerate(draws):
        arr[i].rot[:] = [float(v) for v in np.asarray(d.rot, dtype=np.float64).reshape(4)]
        arr[i].offset[:] = [float(v) for v in np.asarray(d.offset, dtype=np.float64).reshape(2)]
        arr[i].scale = np.float32(d.scale)      # numpy multiplies the float32 slices by a weak python float: float32 arithmetic
        arr[i].shift = np.float32(d.shift)
    return arr


def params_bytes(draws: Sequence[TtaDraw]) -> np.ndarray:
    arr = params_array(draws)
    return np.frombuffer(bytes(arr), dtype=np.uint8).copy()


DEFAULTS = dict(max_rotation_deg=5.0, max_translation=0.05, intensity_scale=0.1, intensity_shift=0.1, noise_std=0.01)


def tta_config(cfg: Dict) -> Dict:
    return {k: float(cfg.get(k, v)) for k, v in DEFAULTS.items()}
