"""K3/K4 parity on the GPU: MIL head, ModDrop sweep and MoE sweep vs the golden outputs of the reference."""
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from pd_fusion_b200.heads import MilHead, ModDropSweep, MoeSweep
from pd_fusion_b200.synthetic import synthetic_table


@pytest.mark.parametrize("tag", ["mil_gated", "mil_plain", "mil_c3"])
def test_mil_head(golden, tag):
    g = golden("heads")
    cfg = json.loads(str(g[f"{tag}/cfg"]))
    sd = {k.split("/sd/")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{tag}/sd/")}
    none = g[f"{tag}/none"]
    mask = g[f"{tag}/mask_mri"]
    bags = [g[f"{tag}/bag{i}"] for i in range(len(none))]
    lmax = max(b.shape[0] for b in bags)
    X = np.zeros((len(bags), lmax, cfg["D"]), np.float32)
    lens = np.zeros(len(bags), np.int32)
    for i, b in enumerate(bags):
        if none[i] or mask[i] == 0:
            continue
        X[i, :b.shape[0]] = b
        lens[i] = b.shape[0]
    head = MilHead(sd, cfg["gated"], 0.5)
    p = head.forward(torch.from_numpy(X).cuda(), torch.from_numpy(lens).cuda())
    torch.cuda.synchronize()
    np.testing.assert_allclose(p.cpu().numpy(), g[f"{tag}/prob"], atol=5e-6, rtol=0)


def _robust_scale(X):
    med = np.nanmedian(X, axis=0)
    iqr = np.nanpercentile(X, 75, axis=0) - np.nanpercentile(X, 25, axis=0)
    iqr[iqr == 0] = 1.0
    return (X - med) / iqr


def test_moddrop_and_moe_sweeps(golden):
    from sklearn.metrics import roc_auc_score
    g = golden("heads")
    dims = json.loads(str(g["table/dims"]))
    df, masks = synthetic_table(int(g["table/n"]), dims, seed=42, mask_seed=7)
    y = df["diagnosis"].values
    cols = [c for m in O.MODALITIES for c in df.columns if c.startswith(m + "_")]
    X = _robust_scale(df[cols].values).astype(np.float32)
    sd = {k.split("/sd/")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith("moddrop/sd/")}
    sweep = ModDropSweep(sd, dims)
    assert sweep.mods == O.MODALITIES
    p = sweep.forward(torch.from_numpy(X).cuda(), torch.from_numpy(g["masks_seed11"]).cuda())
    torch.cuda.synchronize()
    p = p.cpu().numpy()
    np.testing.assert_allclose(p, g["moddrop/probs"], atol=5e-6, rtol=0)
    ref_metrics = json.loads(str(g["moddrop/metrics"]))
    scen = json.loads(str(g["scenarios"]))["scenarios"]
    for s, sc in enumerate(scen):   # AUC identical to 3 decimals (north_star)
        assert round(roc_auc_score(y, p[s]), 3) == round(ref_metrics[sc["name"]]["roc_auc"], 3)

    mods = [str(m) for m in g["moe/mods"]]
    sdm = {k.split("/sd/")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith("moe/sd/")}
    moe = MoeSweep(sdm, mods)
    Xd = {m: torch.from_numpy(_robust_scale(df[[c for c in df.columns if c.startswith(m + "_")]].values).astype(np.float32)).cuda()
          for m in mods}
    pm = moe.forward(Xd, torch.from_numpy(g["moe/masks_seed12"]).cuda())
    torch.cuda.synchronize()
    pm = pm.cpu().numpy()
    np.testing.assert_allclose(pm, g["moe/probs"], atol=5e-6, rtol=0)
    ref_metrics = json.loads(str(g["moe/metrics"]))
    for s, sc in enumerate(scen):
        assert round(roc_auc_score(y, pm[s]), 3) == round(ref_metrics[sc["name"]]["roc_auc"], 3)


def test_every_mask_sweep_matches_oracle():
    """The full power set of M=3 modality masks in one launch vs the oracle evaluated mask by mask."""
    dims = {"clinical": 7, "datspect": 3, "mri": 33}
    df, _ = synthetic_table(257, dims, seed=3)
    cols = [c for m in O.MODALITIES for c in df.columns if c.startswith(m + "_")]
    X = df[cols].values.astype(np.float32)
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(43, 48), torch.nn.ReLU(), torch.nn.Linear(48, 20), torch.nn.ReLU(), torch.nn.Linear(20, 1))
    sd = {}
    for key, layer in (("net.0", net[0]), ("net.3", net[2]), ("net.6", net[4])):
        sd[key + ".weight"], sd[key + ".bias"] = layer.weight.detach(), layer.bias.detach()
    masks = np.array([[(p >> b) & 1 for b in range(3)] for p in range(8)], np.uint8)       # [8,3]
    mk = np.repeat(masks[:, None, :], len(X), axis=1)
    p = ModDropSweep(sd, dims).forward(torch.from_numpy(X).cuda(), torch.from_numpy(mk).cuda())
    torch.cuda.synchronize()
    sdn = {k: v.detach().numpy() for k, v in sd.items()}
    for s in range(8):
        ref = O.moddrop_predict_proba(sdn, dims, X, {m: mk[s][:, i] for i, m in enumerate(O.MODALITIES)})
        np.testing.assert_allclose(p[s].cpu().numpy(), ref, atol=5e-6, rtol=0)
