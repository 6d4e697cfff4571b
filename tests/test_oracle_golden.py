"""Pins oracle/oracle.py against outputs of the REFERENCE ITSELF (tests/golden/*.npz, produced
by oracle/make_golden.py in the authoring container).  CPU only."""
import hashlib
import json

import numpy as np
import pytest

from oracle import oracle as O
from pd_fusion_b200.synthetic import synthetic_table, synthetic_volume


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SMALL = ["small_a", "small_b", "small_c"]
FULL = ["full_c2", "full_c5"]


@pytest.mark.parametrize("case", SMALL + FULL)
def test_preprocess_bit_exact(golden, case):
    g = golden("preproc")
    sp = json.loads(str(g[f"{case}/spec"]))
    raw = synthetic_volume(sp["index"], tuple(sp["shape"]), bad_fraction=1e-4 if sp["shape"][0] < 100 else 1e-5)
    zoomed = O.load_volume(raw, tuple(sp["target"]))
    assert sha(zoomed) == str(g[f"{case}/zoomed_sha"])          # scipy zoom restated bit-for-bit
    lo, hi = O.percentile_bounds(zoomed)
    assert lo == g[f"{case}/lo"] and hi == g[f"{case}/hi"]      # np.percentile restated bit-for-bit
    vol = O.normalize_volume_for_resnet(zoomed)
    assert sha(vol) == str(g[f"{case}/vol_sha"])
    for a, c in zip(sp["axes"], sp["counts"]):
        assert np.array_equal(O.select_slice_indices(vol, a, c), g[f"{case}/idx{a}"])
    sl = np.concatenate([O.select_slices(vol, a, c) for a, c in zip(sp["axes"], sp["counts"])], axis=0)
    assert sha(sl) == str(g[f"{case}/slices_sha"])
    size = 56 if sl.shape[1] <= 64 else 224
    x = O.slices_to_input(sl[:2], size)
    np.testing.assert_allclose(x[:, :, ::4, ::4], g[f"{case}/input_sample"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("case,raw", [("zeros", None), ("negative", None)])
def test_preprocess_degenerate(golden, case, raw):
    g = golden("preproc")
    raw = np.zeros((20, 20, 20), np.float32) if case == "zeros" else -synthetic_volume(5, (20, 22, 24), 0.0) - 1.0
    zoomed = O.load_volume(raw, (16, 16, 16))
    lo, hi = O.percentile_bounds(zoomed)
    assert lo == g[f"{case}/lo"] and hi == g[f"{case}/hi"]
    vol = O.normalize_volume_for_resnet(zoomed)
    assert sha(vol) == str(g[f"{case}/vol_sha"])
    assert np.array_equal(O.select_slice_indices(vol, 2, 4), g[f"{case}/idx2"])


def _backbone_sd(arch):
    import torch
    import torchvision
    torch.manual_seed(1234)
    m = getattr(torchvision.models, arch)(weights=None)
    return {k: v for k, v in m.state_dict().items() if not k.startswith("fc.")}


@pytest.mark.parametrize("case", ["r18_small", "r50_small", "r18_c2"])
def test_embedding_vs_reference(golden, case):
    g = golden("embed")
    sp = json.loads(str(g[f"{case}/spec"]))
    sd = _backbone_sd(sp["arch"])
    fp = np.array([float(v.double().sum()) for k, v in sorted(sd.items()) if v.dtype.is_floating_point][:8])
    # fc.* sorts after conv1/bn1/layer*, so the first 8 entries are unaffected by dropping fc
    assert np.allclose(fp, g[f"{sp['arch']}/fingerprint"], rtol=1e-12), "seeded torchvision init drifted"
    raw = synthetic_volume(sp["index"], tuple(sp["shape"]), bad_fraction=1e-4 if sp["shape"][0] < 100 else 1e-5)
    _, emb = O.embed_subject(raw, sd, sp["arch"], tuple(sp["target"]), sp["axes"], sp["counts"], sp["input_size"], sp["bs"])
    ref = g[f"{case}/emb"]
    assert emb.shape == ref.shape
    rel = np.linalg.norm(emb - ref) / np.linalg.norm(ref)
    assert rel < 1e-5, rel


@pytest.mark.parametrize("tag", ["mil_gated", "mil_plain", "mil_c3"])
def test_mil_head(golden, tag):
    g = golden("heads")
    cfg = json.loads(str(g[f"{tag}/cfg"]))
    sd = {k.split("/sd/")[1]: g[k] for k in g.files if k.startswith(f"{tag}/sd/")}
    none = g[f"{tag}/none"]
    bags = [None if none[i] else g[f"{tag}/bag{i}"] for i in range(len(none))]
    p = O.mil_predict_proba(sd, bags, cfg["gated"], masks={"mri": g[f"{tag}/mask_mri"]})
    np.testing.assert_allclose(p, g[f"{tag}/prob"], atol=2e-6, rtol=0)


def _table(g):
    dims = json.loads(str(g["table/dims"]))
    df, masks = synthetic_table(int(g["table/n"]), dims, seed=42, mask_seed=7)
    return dims, df, masks


def _robust_scale(X):
    med = np.nanmedian(X, axis=0)
    iqr = np.nanpercentile(X, 75, axis=0) - np.nanpercentile(X, 25, axis=0)
    iqr[iqr == 0] = 1.0
    return (X - med) / iqr


def test_masks_bit_exact(golden):
    g = golden("heads")
    dims, df, masks = _table(g)
    scen = json.loads(str(g["scenarios"]))["scenarios"]
    np.random.seed(11)
    got = []
    for sc in scen:
        cur = O.apply_missingness_scenario(len(df), sc, masks)
        got.append(O.modality_mask_matrix(cur))
    assert np.array_equal(np.stack(got).astype(np.uint8), g["masks_seed11"])


def test_moddrop_and_moe_probs(golden):
    g = golden("heads")
    dims, df, masks = _table(g)
    scen = json.loads(str(g["scenarios"]))["scenarios"]
    cols = [c for m in O.MODALITIES for c in df.columns if c.startswith(m + "_")]
    X = _robust_scale(df[cols].values)
    sd = {k.split("/sd/")[1]: g[k] for k in g.files if k.startswith("moddrop/sd/")}
    for s in range(len(scen)):
        mk = {m: g["masks_seed11"][s][:, i] for i, m in enumerate(O.MODALITIES)}
        p = O.moddrop_predict_proba(sd, dims, X, mk)
        np.testing.assert_allclose(p, g["moddrop/probs"][s], atol=2e-6, rtol=0)
    mods = [str(m) for m in g["moe/mods"]]
    sdm = {k.split("/sd/")[1]: g[k] for k in g.files if k.startswith("moe/sd/")}
    for s in range(len(scen)):
        mk = g["moe/masks_seed12"][s]
        Xd = {}
        for i, m in enumerate(mods):
            cm = [c for c in df.columns if c.startswith(m + "_")]
            Xd[m] = _robust_scale(df[cm].values) * mk[:, i:i + 1]
        p = O.moe_predict_proba(sdm, Xd, mk.astype(np.float32))
        np.testing.assert_allclose(p, g["moe/probs"][s], atol=2e-6, rtol=0)


def test_tta_affine_and_passes_bit_exact(golden):
    """a6: scipy.ndimage.affine_transform (order 1, constant) restated bit for bit, and the augmented slice stacks of the
    reference's `tta > 1` loop reproduced from the stored per-subject seeds."""
    g = golden("tta")
    for k in range(int(g["affine/n"])):
        got = O.apply_affine_2d(g[f"affine/{k}/img"], float(g[f"affine/{k}/angle"]), g[f"affine/{k}/translate"])
        assert np.array_equal(got.view(np.uint32), g[f"affine/{k}/out"].view(np.uint32)), k
    targs = json.loads(str(g["script/targs"]))
    from pd_fusion_b200.synthetic import write_synthetic_manifest  # noqa: F401  (same generator the fixture used)
    for b, seed in enumerate(g["script/seeds"]):
        raw = synthetic_volume(b, (48, 40, 36))
        _, _, sl = O.preprocess_subject(raw, (32, 32, 32), [0, 2], [3, 2])
        for p, aug in enumerate(O.tta_passes(sl, int(seed), 2, **targs)):
            want = g[f"script/aug/{b}/{p}"]
            assert aug.dtype == np.float32 and np.array_equal(aug.view(np.uint32), want.view(np.uint32)), (b, p)


def test_product_tta_draws_match_oracle(golden):
    """The product's host-side draw logic (pd_fusion_b200/data/tta.py) consumes the Generator exactly like the reference:
    feeding its draws through the oracle's arithmetic reproduces the reference's augmented slices."""
    from pd_fusion_b200.data.tta import draw_passes, params_bytes, tta_config
    g = golden("tta")
    targs = tta_config(json.loads(str(g["script/targs"])))
    raw = synthetic_volume(0, (48, 40, 36))
    _, _, sl = O.preprocess_subject(raw, (32, 32, 32), [0, 2], [3, 2])
    draws = draw_passes(int(g["script/seeds"][0]), 2, sl.shape[0], sl.shape[1:], targs)
    for p, d in enumerate(draws):
        aug = np.stack([O.affine_transform_linear(s, d.rot, d.offset) for s in sl])
        aug = aug * np.float32(d.scale) + np.float32(d.shift)
        aug = np.clip(aug + d.noise, 0.0, 1.0).astype(np.float32)
        assert np.array_equal(aug.view(np.uint32), g[f"script/aug/0/{p}"].view(np.uint32)), p
    assert params_bytes([draws[0], draws[1]]).size == 2 * 56


def test_oracle_moe_c4_dims_vs_reference(golden):
    """The oracle's MoE restatement at BASELINE config 4's dims (512 + 10, N = 10 000) against the reference's own probabilities."""
    import json
    from pd_fusion_b200.synthetic import synthetic_table
    g = golden("c45")
    dims, N = json.loads(str(g["c4/dims"])), int(g["c4/n"])
    df, _ = synthetic_table(N, dims, seed=44, mask_seed=9)
    mods = [str(m) for m in g["c4/mods"]]
    sd = {k.split("/sd/")[1]: g[k] for k in g.files if k.startswith("c4/sd/")}
    masks = np.unpackbits(g["c4/masks_seed13"], axis=1)[:, :N, :]

    def scale(X):
        med = np.nanmedian(X, axis=0)
        iqr = np.nanpercentile(X, 75, axis=0) - np.nanpercentile(X, 25, axis=0)
        iqr[iqr == 0] = 1.0
        return ((X - med) / iqr).astype(np.float32)
    X = {m: scale(df[[c for c in df.columns if c.startswith(m + "_")]].values) for m in mods}
    for s in (0, 2, 5):
        mk = masks[s].astype(np.float32)
        p = O.moe_predict_proba(sd, {m: X[m] * mk[:, i:i + 1] for i, m in enumerate(mods)}, mk)
        np.testing.assert_allclose(p, g["c4/probs"][s], atol=5e-6, rtol=0)


SIMPLE_CASES = ["brain48", "brain96", "nonpositive", "constant", "sparse"]
SIMPLE_CFGS = [(10, 8, False), (16, 4, True), (10, 0, True)]


def check_simple_features(got, want, bins, grid, extra):
    """Layout: [mean, std, min, max, median, p10, p90 | hist[bins] | grid^3 | skew, kurtosis, entropy].  Order statistics, the grid
    zoom and the histogram (exact counts over exact float32 edges) are bit-exact; the float32 pairwise sums behind numpy's mean /
    std and scipy's moments are reproduced to 1e-5 relative (float64 sums here)."""
    assert got.shape == want.shape and got.dtype == np.float32
    np.testing.assert_allclose(got[:2], want[:2], rtol=2e-5, atol=1e-6)
    assert np.array_equal(got[2:7], want[2:7]), (got[2:7], want[2:7])
    assert np.array_equal(got[7:7 + bins], want[7:7 + bins]), (got[7:7 + bins], want[7:7 + bins])
    g3 = grid ** 3
    assert np.array_equal(got[7 + bins:7 + bins + g3], want[7 + bins:7 + bins + g3])
    if extra:
        np.testing.assert_allclose(got[-3:-1], want[-3:-1], rtol=2e-3, atol=2e-3)
        np.testing.assert_allclose(got[-1], want[-1], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("case", SIMPLE_CASES)
def test_simple_features_vs_reference(golden, case):
    """oracle.simple_features against the reference's `_compute_simple_features` outputs (data/openneuro_features.py:34-73)."""
    g = golden("simple")
    vol = O.load_volume(g[f"{case}/raw"], tuple(int(v) for v in g[f"{case}/target"]))
    for bins, grid, extra in SIMPLE_CFGS:
        check_simple_features(O.simple_features(vol, bins, grid, extra), g[f"{case}/feats_{bins}_{grid}_{int(extra)}"], bins, grid, extra)
