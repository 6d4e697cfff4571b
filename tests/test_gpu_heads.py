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


def _mil_case(g, tag):
    cfg = json.loads(str(g[f"{tag}/cfg"]))
    sd = {k.split("/sd/")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{tag}/sd/")}
    none = g[f"{tag}/none"]
    bags = [None if none[i] else g[f"{tag}/bag{i}"] for i in range(len(none))]
    return cfg, sd, bags, g[f"{tag}/mask_mri"], g[f"{tag}/prob"]


@pytest.mark.parametrize("tag", ["mil_gated", "mil_plain", "mil_c3"])
def test_mil_sweep_equals_per_scenario_reference(golden, tag):
    """pdf_mil_sweep: ONE projection + pooling pass for all scenarios.  Scenario 0 is the golden mask (the reference's own
    probabilities), the others are further drop patterns; each row must equal what the reference's per-scenario call returns
    (its probability where the bag is live, missing_prob elsewhere) -- FP32 path, 5e-6."""
    from pd_fusion_b200.models.mil_attention import MilAttentionModel
    g = golden("heads")
    cfg, sd, bags, mask, ref = _mil_case(g, tag)
    n = len(bags)
    params = {"hidden_dim": sd["instance.0.weight"].shape[0], "attn_dim": (sd["attn_v.0.weight"] if cfg["gated"] else sd["attn.0.weight"]).shape[0],
              "gated": cfg["gated"], "missing_prob": 0.5}
    model = MilAttentionModel(cfg["D"], params)
    model.model.load_state_dict(sd)
    model.invalidate()
    rng = np.random.default_rng(3)
    mm = np.stack([mask, np.ones(n, int), np.zeros(n, int), rng.integers(0, 2, n), rng.integers(0, 2, n)])
    probs = model.predict_proba_sweep(bags, mm)
    assert probs.shape == (5, n) and probs.dtype == np.float64
    np.testing.assert_allclose(probs[0], ref, atol=5e-6, rtol=0)
    full = model.predict_proba(bags, masks={"mri": np.ones(n, int)})
    for s in range(5):
        want = np.where((mm[s] != 0) & np.array([b is not None for b in bags]), full, 0.5)
        np.testing.assert_allclose(probs[s], want, atol=1e-7, rtol=0)
    assert np.all(probs[2] == 0.5)


@pytest.mark.parametrize("D,H,A,gated,L", [(2048, 256, 128, True, 48), (2048, 256, 128, False, 72), (512, 128, 64, True, 24), (96, 64, 64, False, 5)])
def test_mil_tensor_path(D, H, A, gated, L):
    """tcgen05 kind::tf32 projection + attention GEMM (scores in the epilogue) against the FP32 path on ragged bags, several
    row tiles per CTA.  TF32 products carry a 10-bit mantissa: probabilities agree to 2e-3 (typically 1e-4)."""
    from pd_fusion_b200.models.mil_attention import MILAttentionNet
    torch.manual_seed(D + H + L)
    sd = MILAttentionNet(D, H, A, 0.2, gated=gated).state_dict()
    n = 700 if D <= 512 else 67
    rng = np.random.default_rng(L)
    lens = rng.integers(1, L + 1, n).astype(np.int32)
    lens[::9] = 0
    X = rng.standard_normal((n, L, D)).astype(np.float32) * np.float32(1.5)
    for i in range(n):
        X[i, lens[i]:] = 0
    live = rng.integers(0, 2, (7, n)).astype(np.uint8)
    Xd, ld, lv = torch.from_numpy(X).cuda(), torch.from_numpy(lens).cuda(), torch.from_numpy(live).cuda()
    p32 = MilHead(sd, gated, 0.37, precision="fp32").sweep(Xd, ld, lv)
    p32 = p32.cpu().numpy()
    dead = (live == 0) | (lens[None, :] == 0)
    assert np.all(p32[dead] == np.float32(0.37))
    assert p32[~dead].std() > 1e-3                        # the comparison is not between constants
    from pd_fusion_b200 import _lib
    lib = _lib.load()
    try:
        for mt in (1, 2, 0):                              # 128-row tiles, 256-row tiles (one weight pass per two tiles), automatic
            _lib.check(lib.pdf_debug_set_mil_mt(mt))
            ptf = MilHead(sd, gated, 0.37, precision="tf32").sweep(Xd, ld, lv).cpu().numpy()
            assert np.all(ptf[dead] == np.float32(0.37)), mt
            np.testing.assert_allclose(ptf, p32, atol=2e-3, rtol=0, err_msg=f"row tiles per weight pass = {mt}")
    finally:
        lib.pdf_debug_set_mil_mt(0)
    # and the FP32 path against the oracle on a few bags
    sdn = {k: v.numpy() for k, v in sd.items()}
    ref = O.mil_predict_proba(sdn, [X[i, :lens[i]] if lens[i] else None for i in range(8)], gated, None, 0.37)
    np.testing.assert_allclose(MilHead(sd, gated, 0.37).forward(Xd[:8].contiguous(), ld[:8]).cpu().numpy(), ref, atol=5e-6, rtol=0)


@pytest.mark.parametrize("dims,hidden,N", [({"clinical": 0, "datspect": 0, "mri": 512}, [256, 128, 64], 3001),
                                            ({"clinical": 10, "datspect": 5, "mri": 20}, [64, 32], 5000),
                                            ({"clinical": 7, "datspect": 3, "mri": 33}, [500, 40], 1100)])
def test_moddrop_tiled_path_equals_warp_path_and_oracle(dims, hidden, N):
    """From 4096 (scenario, subject) pairs up the sweep runs layers 2.. as in-block tiled GEMMs (weights fetched once per 64-pair
    tile): same probabilities as the warp-per-pair kernel (2e-6) and as the oracle evaluated scenario by scenario (5e-6)."""
    from pd_fusion_b200 import _lib
    from pd_fusion_b200.models.fusion_moddrop import ModalityDropoutNet
    lib = _lib.load()
    torch.manual_seed(N)
    sd = ModalityDropoutNet(dims, hidden, 0.3).state_dict()
    F = sum(dims.values())
    rng = np.random.default_rng(N)
    X = rng.standard_normal((N, F)).astype(np.float32)
    mk = (rng.random((7, N, 3)) < 0.7).astype(np.uint8)
    sweep = ModDropSweep(sd, dims)
    Xd, md = torch.from_numpy(X).cuda(), torch.from_numpy(mk).cuda()
    p_t = sweep.forward(Xd, md).cpu().numpy()
    lib.pdf_debug_set_moddrop_tiled(0)
    try:
        p_w = sweep.forward(Xd, md).cpu().numpy()
    finally:
        lib.pdf_debug_set_moddrop_tiled(1)
    np.testing.assert_allclose(p_t, p_w, atol=2e-6, rtol=0)
    sdn = {k: v.numpy() for k, v in sd.items()}
    for s in (0, 3, 6):
        ref = O.moddrop_predict_proba(sdn, dims, X, {m: mk[s][:, i] for i, m in enumerate(O.MODALITIES)})
        np.testing.assert_allclose(p_t[s], ref, atol=5e-6, rtol=0)
