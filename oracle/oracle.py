"""CPU oracle for the imaging-embedding + fusion hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithms of the reference path
(`/root/reference`, cited as file:line relative to that root).  It is imported only by
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs.  The product package (`robust-multimodal-pd_b200/pd_fusion_b200`) never imports it.

Parity pin: the reference's own tests hold NO golden vectors for this path
(SURVEY.md §4), so every function here is pinned against outputs of the reference itself,
executed in the authoring container by `oracle/make_golden.py` and committed under
`tests/golden/` (`tests/test_oracle_golden.py` replays them).  Third-party arithmetic
the reference calls (not vendored under /root/reference; bare names in pyproject.toml:11-27,
versions installed here in brackets) is restated from its published algorithm:
  * scipy.ndimage.zoom(order=1)     [scipy 1.18.1]  -> `zoom_trilinear`
  * numpy.percentile / linspace      [numpy 2.3.5]   -> `percentile_linear`, `linspace_indices`
  * torch F.interpolate(bilinear)    [torch 2.11.0]  -> `bilinear_resize`
  * torchvision resnet18/50          [tv 0.26.0]     -> `resnet_forward` (torch fp32 functional ops)
  * scipy.ndimage.affine_transform(order=1, mode="constant")  [scipy 1.18.1]  -> `affine_transform_linear`
Integer/index work is numpy; floating-point contractions use torch fp32 on the CPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MODALITIES = ["clinical", "datspect", "mri"]  # data/schema.py:3


# ----------------------------------------------------------------------------------------
# a1  _load_volume  (data/openneuro_features.py:22-32)
# ----------------------------------------------------------------------------------------
def zoom_tables(src: int, dst: int):
    """Per-axis sample tables of ndimage.zoom(order=1, grid_mode=False):
    output index i samples input coordinate i*(src-1)/(dst-1) (float64)."""
    i = np.arange(dst, dtype=np.float64)
    step = (src - 1) / (dst - 1) if dst > 1 else 0.0
    c = i * step
    f = np.floor(c)
    w1 = c - f
    w0 = 1.0 - w1
    i0 = f.astype(np.int64)
    i1 = np.minimum(i0 + 1, src - 1)
    return i0, i1, w0, w1


def zoom_trilinear(vol: np.ndarray, target_shape: Sequence[int]) -> np.ndarray:
    """Bit-exact restatement of `ndimage.zoom(vol, [t/s...], order=1)` for f32 input
    (openneuro_features.py:30-31).  scipy accumulates the 8 taps in float64 as
    ((v*wx)*wy)*wz, last axis fastest, then rounds once to f32."""
    v = vol.astype(np.float64)
    tabs = [zoom_tables(s, t) for s, t in zip(vol.shape, target_shape)]
    out = np.zeros(tuple(target_shape), dtype=np.float64)
    for a in (0, 1):
        ia, wa = tabs[0][a], tabs[0][2 + a]
        for b in (0, 1):
            ib, wb = tabs[1][b], tabs[1][2 + b]
            for c in (0, 1):
                ic, wc = tabs[2][c], tabs[2][2 + c]
                term = v[np.ix_(ia, ib, ic)]
                term = term * wa[:, None, None]
                term = term * wb[None, :, None]
                term = term * wc[None, None, :]
                out = out + term
    return out.astype(np.float32)


def decode_stored_voxels(voxels: np.ndarray, shape, fortran: bool, slope: float = 1.0, inter: float = 0.0) -> np.ndarray:
    """`nib.load(p).get_fdata().astype(np.float32)` (openneuro_features.py:24-25) from the stored voxels.

    nibabel is NOT installed here and is not vendored by the reference (pyproject.toml lists the bare name): this restates
    nibabel 5.x's published behaviour -- `get_fdata()` returns float64; the array proxy applies `arr * scl_slope + scl_inter`
    (two float64 operations) unless the slope is 0 / non-finite or (slope, inter) == (1, 0); a non-finite inter counts as 0;
    data are Fortran-ordered in the file.  PARITY UNPINNED for this one function (no nibabel to run, no reference fixture)."""
    data = np.asarray(voxels).reshape(tuple(shape), order="F" if fortran else "C").astype(np.float64)
    if slope != 0 and np.isfinite(slope) and not (slope == 1.0 and inter == 0.0):
        data = data * np.float64(slope) + np.float64(inter if np.isfinite(inter) else 0.0)
    return np.ascontiguousarray(data.astype(np.float32))


def load_volume(data: np.ndarray, target_shape=(160, 160, 160)) -> np.ndarray:
    """openneuro_features.py:25-31 minus the nibabel decode (out of scope)."""
    d = np.asarray(data).astype(np.float32)
    d = np.nan_to_num(d, nan=0.0, posinf=0.0, neginf=0.0)
    if target_shape is not None:
        d = zoom_trilinear(d, target_shape)
    return d


# ----------------------------------------------------------------------------------------
# a2  _normalize_volume_for_resnet  (data/openneuro_features.py:121-132)
# ----------------------------------------------------------------------------------------
def percentile_linear(sorted_vals: np.ndarray, q: float) -> np.float32:
    """numpy.percentile(method='linear') on an f32 array with a Python-number q
    (numpy/lib/_function_base_impl.py: percentile -> _quantile -> _lerp, numpy 2.x / NEP 50):
    q is divided by float32(100), so the virtual index (n-1)*q is evaluated in FLOAT32
    ((n-1) is a weak Python int); gamma = virtual - floor(virtual) (exact, f32);
    lerp in f32 with the `t >= 0.5` branch."""
    n = int(sorted_vals.shape[0])
    q32 = np.float32(q) / np.float32(100)
    virtual = np.float32(np.float32(n - 1) * q32)
    prev = np.float32(np.floor(virtual))
    g = np.float32(virtual - prev)
    if virtual >= n - 1:
        f0 = f1 = n - 1
    elif virtual < 0:
        f0 = f1 = 0
    else:
        f0 = int(prev)
        f1 = f0 + 1
    a = np.float32(sorted_vals[f0])
    b = np.float32(sorted_vals[f1])
    d = np.float32(b - a)
    if g >= np.float32(0.5):
        return np.float32(b - np.float32(d * np.float32(np.float32(1.0) - g)))
    return np.float32(a + np.float32(d * g))


def percentile_bounds(vol: np.ndarray) -> Tuple[np.float32, np.float32]:
    """lo/hi of openneuro_features.py:122-129 (p1/p99 of voxels > 0, else min/max)."""
    vals = vol[vol > 0]
    if vals.size > 0:
        s = np.sort(vals.astype(np.float32))
        return percentile_linear(s, 1.0), percentile_linear(s, 99.0)
    return np.float32(np.min(vol)), np.float32(np.max(vol))


def normalize_volume_for_resnet(vol: np.ndarray) -> np.ndarray:
    lo, hi = percentile_bounds(vol)
    v = np.clip(vol.astype(np.float32), lo, hi)
    # NEP-50: python float 1e-6 is weak -> whole expression stays float32
    v = (v - lo) / np.float32(np.float32(hi - lo) + np.float32(1e-6))
    return v.astype(np.float32)


# ----------------------------------------------------------------------------------------
# a3  _select_slices  (data/openneuro_features.py:134-151)
# ----------------------------------------------------------------------------------------
def uniform_histogram_counts(vals: np.ndarray, lo: np.float32, hi: np.float32, bins: int):
    """Counts and float32 edges of np.histogram(vals, bins, range=(lo, hi)) for float32 values inside [lo, hi]
    (numpy/lib/_histograms_impl.py, uniform-bin path): equal outer edges are widened by 0.5; the edges are
    np.linspace(first, last, bins + 1) evaluated in float32 (arange * step + start, last edge = stop); after numpy's +-1
    corrections a value sits in the bin with edges[i] <= v < edges[i+1], the last bin closed on the right."""
    first, last = np.float32(lo), np.float32(hi)
    if first == last:
        first, last = np.float32(first - np.float32(0.5)), np.float32(last + np.float32(0.5))
    step = np.float32(np.float32(last - first) / np.float32(bins))
    edges = (np.arange(bins, dtype=np.float32) * step + first).astype(np.float32)
    edges = np.concatenate([edges, np.array([last], dtype=np.float32)])
    v = vals.astype(np.float32)
    idx = np.searchsorted(edges[1:bins], v, side="right")            # interior edges <= v
    return np.bincount(idx, minlength=bins).astype(np.int64), edges


def simple_features(volume: np.ndarray, hist_bins: int = 10, grid_size: int = 8, extra_stats: bool = False) -> np.ndarray:
    """`_compute_simple_features` (data/openneuro_features.py:34-73): statistics of the positive voxels (all voxels when there is
    none), a density histogram of the values clipped to [p1, p99], the trilinear zoom of the volume to grid_size^3, optionally
    skewness / excess kurtosis (scipy.stats defaults: biased moments, Fisher) and the entropy of the histogram."""
    vol = volume.astype(np.float32)
    sel = vol > 0
    vals = vol[sel] if sel.any() else vol.reshape(-1)
    s = np.sort(vals)
    n = s.shape[0]
    mean = np.float32(vals.astype(np.float64).sum() / n)
    var = ((vals.astype(np.float64) - vals.astype(np.float64).mean()) ** 2).mean()
    std = np.float32(np.sqrt(var))
    median = s[n // 2] if n % 2 else np.float32(np.float32(s[n // 2 - 1] + s[n // 2]) / np.float32(2))
    p10, p90 = percentile_linear(s, 10.0), percentile_linear(s, 90.0)
    lo, hi = percentile_linear(s, 1.0), percentile_linear(s, 99.0)
    counts, edges = uniform_histogram_counts(np.clip(vals, lo, hi), lo, hi, hist_bins)
    hist = counts / np.array(np.diff(edges), float) / counts.sum()
    feats = [float(mean), float(std), float(s[0]), float(s[-1]), float(median), float(p10), float(p90)] + hist.tolist()
    if grid_size:
        feats += zoom_trilinear(vol, (grid_size,) * 3).reshape(-1).tolist()
    if extra_stats:
        d = vals.astype(np.float64) - vals.astype(np.float64).mean()
        m2, m3, m4 = (d ** 2).mean(), (d ** 3).mean(), (d ** 4).mean()
        sk = float(np.nan_to_num(m3 / m2 ** 1.5 if m2 > 0 else np.nan, nan=0.0))
        kt = float(np.nan_to_num(m4 / m2 ** 2 - 3.0 if m2 > 0 else np.nan, nan=0.0))
        h = hist + 1e-12
        feats += [sk, kt, float(-(h * np.log(h)).sum())]
    return np.array(feats, dtype=np.float32)


def linspace_indices(lo: int, hi: int, count: int) -> np.ndarray:
    """np.linspace(lo, hi, count).astype(int): f64 `i*step + lo`, last forced to hi, truncation."""
    if count <= 0:
        return np.zeros((0,), dtype=np.int64)
    if count == 1:
        return np.array([lo], dtype=np.int64)
    step = (np.float64(hi) - np.float64(lo)) / np.float64(count - 1)
    y = np.arange(count, dtype=np.float64) * step + np.float64(lo)
    y[-1] = np.float64(hi)
    return y.astype(np.int64)


def select_slice_indices(vol_norm: np.ndarray, axis: int, slice_count: int) -> np.ndarray:
    other = tuple(i for i in range(3) if i != axis)
    nonzero = np.any(vol_norm > 0, axis=other)
    idxs = np.where(nonzero)[0]
    if len(idxs) == 0:
        idxs = np.arange(vol_norm.shape[axis])
    lo, hi = int(idxs[0]), int(idxs[-1])
    n = min(int(slice_count), hi - lo + 1)
    return linspace_indices(lo, hi, n)


def select_slices(vol_norm: np.ndarray, axis: int, slice_count: int) -> np.ndarray:
    idx = select_slice_indices(vol_norm, axis, slice_count)
    if axis == 0:
        return vol_norm[idx, :, :]
    if axis == 1:
        return vol_norm[:, idx, :].transpose(1, 0, 2)
    return vol_norm[:, :, idx].transpose(2, 0, 1)


# ----------------------------------------------------------------------------------------
# a4  slices -> network input (data/openneuro_features.py:250-255)
# ----------------------------------------------------------------------------------------
def bilinear_tables(src: int, dst: int):
    """ATen upsample_bilinear2d(align_corners=False): src = (dst+0.5)*(in/out) - 0.5 clamped at 0, f32."""
    scale = np.float32(src) / np.float32(dst)
    d = np.arange(dst, dtype=np.float32)
    s = (d + np.float32(0.5)) * scale - np.float32(0.5)
    s = np.maximum(s, np.float32(0.0)).astype(np.float32)
    i0 = np.minimum(np.floor(s).astype(np.int64), src - 1)
    i1 = np.minimum(i0 + 1, src - 1)
    w1 = (s - i0.astype(np.float32)).astype(np.float32)
    w0 = (np.float32(1.0) - w1).astype(np.float32)
    return i0, i1, w0, w1


def bilinear_resize(slices: np.ndarray, size: int) -> np.ndarray:
    """[n,H,W] f32 -> [n,size,size] f32."""
    s = slices.astype(np.float32)
    hi0, hi1, hw0, hw1 = bilinear_tables(s.shape[1], size)
    wi0, wi1, ww0, ww1 = bilinear_tables(s.shape[2], size)
    top = s[:, hi0][:, :, wi0] * ww0[None, None, :] + s[:, hi0][:, :, wi1] * ww1[None, None, :]
    bot = s[:, hi1][:, :, wi0] * ww0[None, None, :] + s[:, hi1][:, :, wi1] * ww1[None, None, :]
    return (top * hw0[None, :, None] + bot * hw1[None, :, None]).astype(np.float32)


def slices_to_input(slices: np.ndarray, input_size: int, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> np.ndarray:
    """[n,H,W] -> [n,3,S,S] f32 NCHW: resize, replicate x3, (x-mean)/std."""
    r = bilinear_resize(slices, input_size)
    x = np.repeat(r[:, None, :, :], 3, axis=1)
    m = np.asarray(mean, dtype=np.float32).reshape(1, 3, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(1, 3, 1, 1)
    return ((x - m) / s).astype(np.float32)


def preprocess_subject(raw: np.ndarray, target_shape, axes: Sequence[int], counts: Sequence[int]):
    """a1-a3 for one subject. Returns (normalised volume, [indices per axis], slices [L,H,W])."""
    vol = normalize_volume_for_resnet(load_volume(raw, target_shape))
    idx = [select_slice_indices(vol, a, c) for a, c in zip(axes, counts)]
    sl = np.concatenate([select_slices(vol, a, c) for a, c in zip(axes, counts)], axis=0)
    return vol, idx, sl


# ----------------------------------------------------------------------------------------
# a5  torchvision ResNet18/50 forward, eval mode, fc=Identity (openneuro_features.py:153-164,257-262)
# ----------------------------------------------------------------------------------------
RESNET_SPECS = {"resnet18": ("basic", [2, 2, 2, 2]), "resnet50": ("bottleneck", [3, 4, 6, 3])}


def resnet_forward(sd: Dict[str, "torch.Tensor"], arch: str, x: "torch.Tensor", batch_size: int = 0):
    """Functional restatement of torchvision.models.resnet.ResNet._forward_impl with BatchNorm in
    eval mode (running stats, eps 1e-5), returning the pooled [n, D] features (fc = Identity)."""
    import torch
    import torch.nn.functional as F

    kind, layers = RESNET_SPECS[arch]

    def bn(t, p):
        return F.batch_norm(t, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            training=False, eps=1e-5)

    def run(xb):
        t = F.conv2d(xb, sd["conv1.weight"], stride=2, padding=3)
        t = F.relu(bn(t, "bn1"))
        t = F.max_pool2d(t, kernel_size=3, stride=2, padding=1)
        for li, nblocks in enumerate(layers, start=1):
            for bi in range(nblocks):
                p = f"layer{li}.{bi}"
                stride = 2 if (li > 1 and bi == 0) else 1
                idn = t
                if kind == "basic":
                    o = F.relu(bn(F.conv2d(t, sd[p + ".conv1.weight"], stride=stride, padding=1), p + ".bn1"))
                    o = bn(F.conv2d(o, sd[p + ".conv2.weight"], padding=1), p + ".bn2")
                else:
                    o = F.relu(bn(F.conv2d(t, sd[p + ".conv1.weight"]), p + ".bn1"))
                    o = F.relu(bn(F.conv2d(o, sd[p + ".conv2.weight"], stride=stride, padding=1), p + ".bn2"))
                    o = bn(F.conv2d(o, sd[p + ".conv3.weight"]), p + ".bn3")
                if (p + ".downsample.0.weight") in sd:
                    idn = bn(F.conv2d(t, sd[p + ".downsample.0.weight"], stride=stride), p + ".downsample.1")
                t = F.relu(o + idn)
        return torch.flatten(F.adaptive_avg_pool2d(t, 1), 1)

    with torch.no_grad():
        if not batch_size:
            return run(x)
        return torch.cat([run(x[i:i + batch_size]) for i in range(0, x.shape[0], batch_size)], dim=0)


def embed_subject(raw: np.ndarray, sd, arch: str, target_shape=(160, 160, 160), axes=(2,), counts=(24,),
                  input_size: int = 224, batch_size: int = 32, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """tta=1 path of the builders: returns (indices, per-slice embeddings [L,D] f32)."""
    import torch

    _, idx, sl = preprocess_subject(raw, target_shape, axes, counts)
    x = torch.from_numpy(slices_to_input(sl, input_size, mean, std))
    emb = resnet_forward(sd, arch, x, batch_size).numpy()
    return idx, emb


# ----------------------------------------------------------------------------------------
# a6  test-time augmentation (data/openneuro_features.py:166-178, 231-248; scripts/build_resnet2d_mil_embeddings.py:120-139)
# ----------------------------------------------------------------------------------------
def affine_transform_linear(img: np.ndarray, matrix: np.ndarray, offset: np.ndarray) -> np.ndarray:
    """scipy.ndimage.affine_transform(img, matrix, offset, order=1, mode="constant", cval=0) for a 2-D float32 image:
    input coordinate = offset + matrix @ (oy, ox) accumulated left to right in float64; a point whose coordinate leaves
    [0, len-1] on either axis is 0; otherwise the 4 taps are summed as ((v * wy) * wx) in row-major tap order (a tap
    index one past the edge -- only reachable with weight 0 -- reads cval) and rounded once to float32."""
    H, W = img.shape
    oy, ox = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    c0 = (offset[0] + oy * matrix[0, 0]) + ox * matrix[0, 1]
    c1 = (offset[1] + oy * matrix[1, 0]) + ox * matrix[1, 1]
    inside = (c0 >= 0) & (c0 <= H - 1) & (c1 >= 0) & (c1 <= W - 1)
    f0, f1 = np.floor(c0), np.floor(c1)
    y, x = c0 - f0, c1 - f1
    i0 = np.clip(f0.astype(np.int64), 0, H - 1)
    j0 = np.clip(f1.astype(np.int64), 0, W - 1)
    v = img.astype(np.float64)

    def tap(ii, jj):
        ok = (ii < H) & (jj < W)
        return np.where(ok, v[np.minimum(ii, H - 1), np.minimum(jj, W - 1)], 0.0)

    t = np.zeros_like(c0)
    t = t + (tap(i0, j0) * (1 - y)) * (1 - x)
    t = t + (tap(i0, j0 + 1) * (1 - y)) * x
    t = t + (tap(i0 + 1, j0) * y) * (1 - x)
    t = t + (tap(i0 + 1, j0 + 1) * y) * x
    return np.where(inside, t, 0.0).astype(img.dtype)


def apply_affine_2d(slice_2d: np.ndarray, angle_deg: float, translate: np.ndarray) -> np.ndarray:
    """`_apply_affine_2d` (openneuro_features.py:166-178)."""
    theta = np.deg2rad(angle_deg)
    rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    center = np.array(slice_2d.shape) / 2.0
    offset = center - rot @ center + translate
    return affine_transform_linear(slice_2d, rot, offset)


def tta_passes(slices: np.ndarray, seed: int, n_pass: int, max_rotation_deg=5.0, max_translation=0.05, intensity_scale=0.1,
               intensity_shift=0.1, noise_std=0.01) -> List[np.ndarray]:
    """The augmented float32 slice stacks of the `tta > 1` loop, one per pass, drawing from ONE Generator(seed) in the
    reference's order (angle, translate[2], scale, shift, noise field); cast to float32 as the MIL script does (:139)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_pass):
        aug = slices.copy()
        angle = rng.uniform(-max_rotation_deg, max_rotation_deg)
        translate = rng.uniform(-max_translation, max_translation, size=2)
        translate = translate * np.array([aug.shape[1], aug.shape[2]])
        for i in range(aug.shape[0]):
            aug[i] = apply_affine_2d(aug[i], angle, translate)
        scale = 1.0 + rng.uniform(-intensity_scale, intensity_scale)
        shift = rng.uniform(-intensity_shift, intensity_shift)
        aug = aug * scale + shift                      # float32 array, weak python scalars: stays float32 (NEP 50)
        if noise_std > 0:
            aug = aug + rng.normal(0.0, noise_std, size=aug.shape)    # float64 from here
        aug = np.clip(aug, 0.0, 1.0)
        out.append(aug.astype(np.float32, copy=False))
    return out


def embed_subject_tta(raw: np.ndarray, sd, arch: str, seed: int, n_pass: int, tta_kwargs: Dict, target_shape=(160, 160, 160),
                      axes=(2,), counts=(24,), input_size: int = 224, batch_size: int = 32):
    """`tta > 1` path: per-slice embeddings averaged over the passes [L, D] f32."""
    import torch

    _, idx, sl = preprocess_subject(raw, target_shape, axes, counts)
    acc = None
    for aug in tta_passes(sl, seed, n_pass, **tta_kwargs):
        emb = resnet_forward(sd, arch, torch.from_numpy(slices_to_input(aug, input_size)), batch_size).numpy()
        acc = emb if acc is None else acc + emb
    return idx, acc / max(1, n_pass)


# ----------------------------------------------------------------------------------------
# a8  MILAttentionNet.forward (models/mil_attention.py:40-51), eval mode, one bag
# ----------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def mil_forward(sd: Dict[str, np.ndarray], bag: np.ndarray, gated: bool) -> float:
    g = {k: np.asarray(v, dtype=np.float32) for k, v in sd.items()}
    x = bag.astype(np.float32)
    h = np.maximum(x @ g["instance.0.weight"].T + g["instance.0.bias"], 0.0)
    if gated:
        v = np.tanh(h @ g["attn_v.0.weight"].T + g["attn_v.0.bias"])
        u = _sigmoid(h @ g["attn_u.0.weight"].T + g["attn_u.0.bias"])
        s = (v * u) @ g["attn_w.weight"].T + g["attn_w.bias"]
    else:
        t = np.tanh(h @ g["attn.0.weight"].T + g["attn.0.bias"])
        s = t @ g["attn.2.weight"].T + g["attn.2.bias"]
    s = s[:, 0].astype(np.float32)
    e = np.exp(s - s.max())
    w = (e / e.sum()).astype(np.float32)
    pooled = (w[:, None] * h).sum(axis=0)
    z = pooled @ g["classifier.0.weight"].T + g["classifier.0.bias"]
    return float(_sigmoid(z.astype(np.float32))[0])


def mil_predict_proba(sd, bags: List[Optional[np.ndarray]], gated: bool, masks=None, missing_prob: float = 0.5):
    """MilAttentionModel.predict_proba (models/mil_attention.py:157-178)."""
    mri = masks["mri"] if isinstance(masks, dict) and "mri" in masks else None
    out = []
    for i, bag in enumerate(bags):
        if bag is None or (mri is not None and mri[i] == 0):
            out.append(missing_prob)
        else:
            out.append(mil_forward(sd, bag, gated))
    return np.array(out)


# ----------------------------------------------------------------------------------------
# a9  ModalityDropoutModel.predict_proba (models/fusion_moddrop.py:93-114)
# ----------------------------------------------------------------------------------------
def moddrop_slices(modality_dims: Dict[str, int]) -> Dict[str, Tuple[int, int]]:
    """Feature blocks laid out in sorted(modality) order (fusion_moddrop.py:12-22)."""
    cur, out = 0, {}
    for mod in sorted(modality_dims):
        out[mod] = (cur, cur + modality_dims[mod])
        cur += modality_dims[mod]
    return out


def moddrop_predict_proba(sd, modality_dims, X: np.ndarray, masks: Optional[Dict[str, np.ndarray]]):
    x = np.asarray(X, dtype=np.float32).copy()
    if masks is not None:
        zm = np.ones_like(x)
        for mod, (a, b) in moddrop_slices(modality_dims).items():
            if mod in masks:
                zm[:, a:b] = np.asarray(masks[mod], dtype=np.float32)[:, None]
        x = x * zm
    keys = sorted({int(k.split(".")[1]) for k in sd if k.startswith("net.")})
    for j, li in enumerate(keys):
        x = x @ np.asarray(sd[f"net.{li}.weight"], np.float32).T + np.asarray(sd[f"net.{li}.bias"], np.float32)
        if j < len(keys) - 1:
            x = np.maximum(x, 0.0)
    return _sigmoid(x.astype(np.float32)).flatten()


# ----------------------------------------------------------------------------------------
# a10 MoENet.forward (models/moe.py:37-47)
# ----------------------------------------------------------------------------------------
def moe_predict_proba(sd, X_dict: Dict[str, np.ndarray], mask: np.ndarray):
    m = np.asarray(mask, dtype=np.float32)
    r = np.maximum(m @ np.asarray(sd["router.0.weight"], np.float32).T + np.asarray(sd["router.0.bias"], np.float32), 0.0)
    r = r @ np.asarray(sd["router.2.weight"], np.float32).T + np.asarray(sd["router.2.bias"], np.float32)
    e = np.exp(r - r.max(axis=1, keepdims=True))
    w = e / e.sum(axis=1, keepdims=True)
    out = np.zeros((m.shape[0],), dtype=np.float32)
    for i, mod in enumerate(sorted(X_dict)):
        x = np.asarray(X_dict[mod], dtype=np.float32)
        keys = sorted({int(k.split(".")[3]) for k in sd if k.startswith(f"experts.{mod}.net.")})
        for j, li in enumerate(keys):
            x = x @ np.asarray(sd[f"experts.{mod}.net.{li}.weight"], np.float32).T + np.asarray(
                sd[f"experts.{mod}.net.{li}.bias"], np.float32)
            if j < len(keys) - 1:
                x = np.maximum(x, 0.0)
        out = out + _sigmoid(x.astype(np.float32))[:, 0] * w[:, i].astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------
# a11 masks (data/missingness.py:7-66, data/feature_utils.py:49-61)
# ----------------------------------------------------------------------------------------
def apply_missingness_scenario(n_subjects: int, scenario: Dict, maskdict: Dict[str, np.ndarray]):
    """Same draws from the GLOBAL numpy RNG, in the same order, as data/missingness.py:7-51."""
    new = {k: np.array(v).copy() for k, v in maskdict.items()}
    if "drop_modalities" in scenario:
        for mod in scenario["drop_modalities"]:
            if mod not in new:
                continue
            if "drop_rate" in scenario:
                rate = float(scenario.get("drop_rate", 0.0))
                if rate <= 0:
                    continue
                drop = np.random.rand(len(new[mod])) < rate
                new[mod][drop] = 0
            else:
                new[mod] = np.zeros_like(new[mod])
    if scenario.get("type") == "random":
        n_drop = scenario.get("n_drop", 1)
        mods = list(new.keys()) if new else MODALITIES
        for i in range(n_subjects):
            avail = [m for m in mods if m in new and new[m][i] == 1]
            if not avail:
                continue
            for mod in np.random.choice(avail, size=min(n_drop, len(avail)), replace=False):
                new[mod][i] = 0
    return new


def modality_mask_matrix(maskdict: Dict[str, np.ndarray]) -> np.ndarray:
    """data/missingness.py:53-66."""
    if not maskdict:
        raise ValueError("maskdict is empty")
    first = next(iter(maskdict.values()))
    return np.stack([maskdict[m] if m in maskdict else np.zeros_like(first) for m in MODALITIES], axis=1)
