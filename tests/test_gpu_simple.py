"""f4, the "simple" feature mode on the device (csrc/simple_feats.cu + K1a's zoom) against the reference's own outputs
(tests/golden/simple.npz: `_compute_simple_features` / `load_simple_features`, data/openneuro_features.py:34-104) and the oracle."""
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from pd_fusion_b200.data import openneuro_features as onf
from pd_fusion_b200.synthetic import write_synthetic_manifest
from test_oracle_golden import SIMPLE_CASES, SIMPLE_CFGS, check_simple_features


@pytest.mark.parametrize("case", SIMPLE_CASES)
def test_simple_features_device_vs_reference(golden, case):
    g = golden("simple")
    vol = O.load_volume(g[f"{case}/raw"], tuple(int(v) for v in g[f"{case}/target"]))
    for bins, grid, extra in SIMPLE_CFGS:
        got = onf._compute_simple_features(vol, hist_bins=bins, grid_size=grid, extra_stats=extra)
        check_simple_features(got, g[f"{case}/feats_{bins}_{grid}_{int(extra)}"], bins, grid, extra)
        check_simple_features(got, O.simple_features(vol, bins, grid, extra), bins, grid, extra)


def test_simple_features_batch_equals_single():
    """Several subjects in one launch (one block each) give the rows the single-volume call gives."""
    rng = np.random.default_rng(3)
    vols = (rng.gamma(4.0, 100.0, size=(5, 40, 36, 32)) * (rng.random((5, 40, 36, 32)) > 0.4)).astype(np.float32)
    batch = onf._simple_features_device(torch.from_numpy(vols).cuda(), 10, 8, True)
    for b in range(5):
        assert np.array_equal(batch[b], onf._compute_simple_features(vols[b], 10, 8, True))


def test_load_simple_features_matches_reference(golden, tmp_path):
    g = golden("simple")
    manifest = write_synthetic_manifest(tmp_path / "vols", int(g["manifest/n"]), start=int(g["manifest/start"]))
    cfg = json.loads(str(g["manifest/cfg"]))
    df = onf.load_simple_features(manifest, tmp_path / "cache", cfg)
    assert list(df.columns) == [str(c) for c in g["manifest/columns"]]
    files = sorted(p.name for p in (tmp_path / "cache").iterdir())
    assert len(files) == 1 and files[0].startswith("features_") and files[0].endswith(".parquet")
    want = g["manifest/feats"]
    got = df[[c for c in df.columns if c.startswith("mri_feat_")]].values
    assert got.dtype == np.float64 and got.shape == want.shape
    for r in range(want.shape[0]):
        check_simple_features(got[r].astype(np.float32), want[r].astype(np.float32), 10, 8, True)
    again = onf.load_simple_features(manifest, tmp_path / "cache", cfg)          # second call reads the cache
    assert np.array_equal(again.values[:, 3:].astype(np.float64), got)
