"""BASELINE configs 4 and 5 at full size against outputs of the reference itself (tests/golden/c45.npz, oracle/make_golden.py c45)."""
import json

import numpy as np
import pandas as pd
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from pd_fusion_b200.backbone import ResNet2D
from pd_fusion_b200.data import openneuro_features as of
from pd_fusion_b200.heads import MoeSweep
from pd_fusion_b200.synthetic import synthetic_table, write_synthetic_manifest


def test_c5_full_size_embeddings_vs_reference_script(golden, tmp_path, monkeypatch):
    """resnet50, slice axes 0/1/2 x 24 = 72 slices of a 256x256x176 volume, --tta 2 (rotation 8 deg, noise 0.02): the bf16 tcgen05
    path (chained pointwise kernels included) against the embeddings the unmodified MIL script wrote; 1e-2 norm-wise per slice."""
    g = golden("c45")

    def _bb(backbone, pretrained=True):
        torch.manual_seed(1234)
        m = ResNet2D("resnet50")
        dim = m.fc.in_features
        m.fc = torch.nn.Identity()
        return m, dim, None
    monkeypatch.setattr(of, "_build_resnet_backbone", _bb)
    monkeypatch.setenv("PD_FUSION_B200_PRECISION", "bf16")
    monkeypatch.setenv("PD_FUSION_B200_SUBJECT_BATCH", "1")
    manifest = write_synthetic_manifest(tmp_path / "vols", 1, shape=(256, 256, 176), start=21)
    df = pd.read_csv(manifest)
    targs = json.loads(str(g["c5/targs"]))
    emb, _ = of.embed_manifest(df, "resnet50", (160, 160, 160), [0, 1, 2], [24, 24, 24], 224, tta=2, tta_cfg=targs,
                               tta_seeds=[int(s) for s in g["c5/seeds"]])
    want = g["c5/emb"]
    assert emb.shape == want.shape == (1, 72, 2048)
    rel = np.linalg.norm(emb - want, axis=2) / np.linalg.norm(want, axis=2)
    assert rel.max() < 1e-2, rel


def _robust_scale(X):
    med = np.nanmedian(X, axis=0)
    iqr = np.nanpercentile(X, 75, axis=0) - np.nanpercentile(X, 25, axis=0)
    iqr[iqr == 0] = 1.0
    return (X - med) / iqr


def test_c4_moe_sweep_full_dims(golden):
    """MoE over imaging (512) + clinical (10) features, N = 10 000, every scenario of configs/eval_missingness.yaml in one launch:
    probabilities within 5e-6 of the reference's per-scenario predict_proba, AUC identical to 3 decimals."""
    from sklearn.metrics import roc_auc_score
    g = golden("c45")
    dims, N = json.loads(str(g["c4/dims"])), int(g["c4/n"])
    df, _ = synthetic_table(N, dims, seed=44, mask_seed=9)
    y = df["diagnosis"].values
    mods = [str(m) for m in g["c4/mods"]]
    sd = {k.split("/sd/")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith("c4/sd/")}
    masks = np.unpackbits(g["c4/masks_seed13"], axis=1)[:, :N, :]
    moe = MoeSweep(sd, mods)
    Xd = {m: torch.from_numpy(_robust_scale(df[[c for c in df.columns if c.startswith(m + "_")]].values).astype(np.float32)).cuda() for m in mods}
    p = moe.forward(Xd, torch.from_numpy(np.ascontiguousarray(masks)).cuda())
    torch.cuda.synchronize()
    p = p.cpu().numpy()
    np.testing.assert_allclose(p, g["c4/probs"], atol=5e-6, rtol=0)
    ref = json.loads(str(g["c4/metrics"]))
    for s, sc in enumerate(json.loads(str(g["c4/scenarios"]))["scenarios"]):
        assert round(roc_auc_score(y, p[s]), 3) == round(ref[sc["name"]]["roc_auc"], 3)
