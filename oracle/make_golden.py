"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF (authoring container only).

    python oracle/make_golden.py            # needs /root/reference; never runs on the GPU box

The reference imports from /root/reference/src with the two patches of SURVEY.md Appendix B
(nibabel is absent -> .npy loader with identical maths; no network -> pretrained=False after a
fixed torch seed).  Nothing from the reference is copied: only its OUTPUTS on seeded synthetic
inputs are stored, together with the seeds needed to regenerate the inputs.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import hashlib
import json
import os
import runpy
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(ROOT / "robust-multimodal-pd_b200"))

import torch  # noqa: E402
from scipy import ndimage  # noqa: E402

import pd_fusion.data.openneuro_features as of  # noqa: E402  (the reference)
from pd_fusion.data.missingness import apply_missingness_scenario  # noqa: E402
from pd_fusion.evaluation.evaluate import evaluate_model  # noqa: E402
from pd_fusion.models.fusion_moddrop import ModalityDropoutModel  # noqa: E402
from pd_fusion.models.mil_attention import MilAttentionModel  # noqa: E402
from pd_fusion.models.moe import MoEModel  # noqa: E402
from pd_fusion_b200.synthetic import synthetic_table, synthetic_volume, write_synthetic_manifest  # noqa: E402

GOLD = ROOT / "tests" / "golden"
BACKBONE_SEED = 1234
HEAD_SEED = 4321


def _load_volume_npy(path, target_shape=(96, 96, 96)):
    # same maths as of._load_volume (openneuro_features.py:25-31) without nibabel
    d = np.load(str(path)).astype(np.float32)
    d = np.nan_to_num(d, nan=0.0, posinf=0.0, neginf=0.0)
    if target_shape is not None:
        d = ndimage.zoom(d, [t / s for t, s in zip(target_shape, d.shape)], order=1)
    return d


_orig_bb = of._build_resnet_backbone


def _bb_seeded(backbone, pretrained=True):
    torch.manual_seed(BACKBONE_SEED)
    return _orig_bb(backbone, pretrained=False)


of._load_volume = _load_volume_npy
of._build_resnet_backbone = _bb_seeded


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def weight_fingerprint(sd) -> np.ndarray:
    return np.array([float(v.double().sum()) for k, v in sorted(sd.items()) if v.dtype.is_floating_point][:8])


def ref_preprocess(raw, target_shape, axes, counts):
    d = np.nan_to_num(raw.astype(np.float32), nan=0.0, posinf=0.0, neginf=0.0)
    zoomed = ndimage.zoom(d, [t / s for t, s in zip(target_shape, d.shape)], order=1)
    vals = zoomed[zoomed > 0]
    lo, hi = (np.percentile(vals, 1), np.percentile(vals, 99)) if vals.size else (np.min(zoomed), np.max(zoomed))
    vol = of._normalize_volume_for_resnet(zoomed)
    idx, sl = [], []
    for a, c in zip(axes, counts):
        s = of._select_slices(vol, a, c)
        sl.append(s)
        # recover the indices the reference used by matching against the volume
        other = tuple(i for i in range(3) if i != a)
        nz = np.where(np.any(vol > 0, axis=other))[0]
        if len(nz) == 0:
            nz = np.arange(vol.shape[a])
        l, h = int(nz[0]), int(nz[-1])
        n = min(c, h - l + 1)
        ind = np.linspace(l, h, n).astype(int)
        assert np.array_equal(np.take(vol, ind, axis=a).transpose((a,) + other) if a else vol[ind], s)
        idx.append(ind)
    return zoomed, np.float32(lo), np.float32(hi), vol, idx, np.concatenate(sl, axis=0)


def ref_input_tensor(slices, input_size):
    import torch.nn.functional as F
    t = torch.from_numpy(slices).unsqueeze(1).float()
    t = F.interpolate(t, size=(input_size, input_size), mode="bilinear", align_corners=False)
    t = t.repeat(1, 3, 1, 1)
    mean = torch.tensor([0.5, 0.5, 0.5]).view(1, 3, 1, 1)
    std = torch.tensor([0.5, 0.5, 0.5]).view(1, 3, 1, 1)
    return (t - mean) / std


def gold_preproc():
    cases = []
    specs = [
        dict(name="small_a", index=0, shape=(64, 48, 44), target=(40, 40, 40), axes=[0, 1, 2], counts=[6, 6, 6]),
        dict(name="small_b", index=1, shape=(50, 70, 36), target=(32, 32, 32), axes=[2], counts=[40]),  # count > extent
        dict(name="small_c", index=2, shape=(33, 35, 31), target=(48, 48, 48), axes=[1], counts=[1]),   # upsample, n==1
        dict(name="full_c2", index=3, shape=(256, 256, 176), target=(160, 160, 160), axes=[2], counts=[24]),
        dict(name="full_c5", index=4, shape=(256, 256, 176), target=(160, 160, 160), axes=[0, 1, 2], counts=[24, 24, 24]),
    ]
    out = {}
    for sp in specs:
        raw = synthetic_volume(sp["index"], sp["shape"], bad_fraction=1e-4 if sp["shape"][0] < 100 else 1e-5)
        zoomed, lo, hi, vol, idx, sl = ref_preprocess(raw, sp["target"], sp["axes"], sp["counts"])
        k = sp["name"]
        out[f"{k}/spec"] = np.array(json.dumps(sp))
        out[f"{k}/zoomed_sha"] = np.array(sha(zoomed))
        out[f"{k}/vol_sha"] = np.array(sha(vol))
        out[f"{k}/lo"], out[f"{k}/hi"] = lo, hi
        for a, ind in zip(sp["axes"], idx):
            out[f"{k}/idx{a}"] = ind.astype(np.int64)
        out[f"{k}/slices_sha"] = np.array(sha(sl))
        out[f"{k}/slices_sample"] = sl[:, ::4, ::4].copy() if sl.shape[1] > 64 else sl.copy()
        x = ref_input_tensor(sl[:2], 56 if sl.shape[1] <= 64 else 224).numpy()
        out[f"{k}/input_sample"] = x[:, :, ::4, ::4].copy()
        cases.append(k)
    # degenerate volumes: all zeros / all negative (the min/max branch, openneuro_features.py:127-129)
    for k, raw in [("zeros", np.zeros((20, 20, 20), np.float32)), ("negative", -synthetic_volume(5, (20, 22, 24), 0.0) - 1.0)]:
        zoomed, lo, hi, vol, idx, sl = ref_preprocess(raw, (16, 16, 16), [2], [4])
        out[f"{k}/lo"], out[f"{k}/hi"] = lo, hi
        out[f"{k}/idx2"] = idx[0].astype(np.int64)
        out[f"{k}/vol_sha"] = np.array(sha(vol))
        cases.append(k)
    out["cases"] = np.array(cases)
    np.savez_compressed(GOLD / "preproc.npz", **out)
    print("preproc.npz", cases)


def gold_embed():
    out = {}
    for arch, specs in [
        ("resnet18", [dict(name="r18_small", index=0, shape=(64, 48, 44), target=(40, 40, 40), axes=[2], counts=[6], input_size=64, bs=4),
                      dict(name="r18_c2", index=3, shape=(256, 256, 176), target=(160, 160, 160), axes=[2], counts=[24], input_size=224, bs=32)]),
        ("resnet50", [dict(name="r50_small", index=1, shape=(50, 70, 36), target=(32, 32, 32), axes=[0, 2], counts=[3, 3], input_size=64, bs=4),
                      dict(name="r50_c3", index=6, shape=(256, 256, 176), target=(160, 160, 160), axes=[2], counts=[48], input_size=224, bs=16)]),
    ]:
        model, emb_dim, _ = of._build_resnet_backbone(arch)
        model.eval()
        out[f"{arch}/fingerprint"] = weight_fingerprint(model.state_dict())
        for sp in specs:
            raw = synthetic_volume(sp["index"], sp["shape"], bad_fraction=1e-4 if sp["shape"][0] < 100 else 1e-5)
            _, _, _, _, idx, sl = ref_preprocess(raw, sp["target"], sp["axes"], sp["counts"])
            x = ref_input_tensor(sl, sp["input_size"])
            feats = []
            with torch.no_grad():
                for i in range(0, x.size(0), sp["bs"]):
                    feats.append(model(x[i:i + sp["bs"]]))
            emb = torch.cat(feats, 0).numpy().astype(np.float32)
            out[f"{sp['name']}/spec"] = np.array(json.dumps({**sp, "arch": arch}))
            out[f"{sp['name']}/emb"] = emb
            print(sp["name"], emb.shape, float(np.abs(emb).mean()))
    np.savez_compressed(GOLD / "embed.npz", **out)


def gold_scripts():
    """Runs the reference's two CLI scripts end-to-end on 3 small synthetic subjects."""
    out = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        manifest = write_synthetic_manifest(td / "vols", 3, shape=(48, 40, 36))
        out["manifest_bytes"] = np.frombuffer(manifest.read_bytes(), dtype=np.uint8)
        # the manifest embeds absolute paths; tests rewrite the path column but keep hashing rules
        out["manifest_dir"] = np.array(str(td / "vols"))
        for script, args, tag in [
            ("build_resnet2d_embeddings.py", ["--target-shape", "32", "32", "32", "--slice-count", "4", "--input-size", "64", "--batch-size", "3"], "c2"),
            ("build_resnet2d_mil_embeddings.py", ["--backbone", "resnet18", "--target-shape", "32", "32", "32", "--slice-axes", "0", "2",
                                                  "--slice-counts", "3", "2", "--input-size", "64", "--batch-size", "4"], "mil"),
        ]:
            od = td / f"out_{tag}"
            argv = sys.argv
            sys.argv = [script, "--manifest", str(manifest), "--out-dir", str(od)] + args
            try:
                runpy.run_path(str(REF / "scripts" / script), run_name="__main__")
            finally:
                sys.argv = argv
            files = sorted(p.name for p in od.iterdir())
            out[f"{tag}/files"] = np.array(files)
            out[f"{tag}/argv"] = np.array(args)
            for p in od.iterdir():
                if p.suffix == ".json":
                    out[f"{tag}/json"] = np.array(p.read_text())
                if p.suffix == ".parquet":
                    import pandas as pd
                    df = pd.read_parquet(p)
                    out[f"{tag}/columns"] = np.array(list(df.columns))
                    out[f"{tag}/emb"] = df[[c for c in df.columns if c.startswith("mri_resnet_")]].values
                    out[f"{tag}/dtypes"] = np.array([str(t) for t in df.dtypes])
                if p.suffix == ".npz":
                    d = np.load(p, allow_pickle=True)
                    out[f"{tag}/emb"] = d["embeddings"]
                    out[f"{tag}/subject_id"] = d["subject_id"].astype(str)
                    out[f"{tag}/label"] = d["label"]
    np.savez_compressed(GOLD / "scripts.npz", **out)
    print("scripts.npz", {k: out[k] for k in ("c2/files", "mil/files")})


def gold_tta():
    """Test-time augmentation: the reference's `_apply_affine_2d` on its own, the augmented slice stacks of its `tta > 1`
    loop (the loop body is inline script code, restated here with the reference's helper and the SAME numpy calls, then
    checked against the embeddings the unmodified MIL script writes), and that script's output.  Run with PYTHONHASHSEED=0."""
    out = {}
    rng = np.random.default_rng(5)
    # 1. _apply_affine_2d alone
    for k, (H, W, ang, tr) in enumerate([(160, 160, 4.2, (3.5, -6.25)), (37, 52, -8.0, (0.0, 0.0)), (32, 32, 0.0, (0.0, 0.0)), (48, 40, 2.5, (-2.4, 2.0))]):
        img = rng.random((H, W)).astype(np.float32)
        img[: H // 6] = 0
        out[f"affine/{k}/img"] = img
        out[f"affine/{k}/angle"], out[f"affine/{k}/translate"] = np.float64(ang), np.array(tr, dtype=np.float64)
        out[f"affine/{k}/out"] = of._apply_affine_2d(img, ang, np.array(tr, dtype=np.float64))
    out["affine/n"] = np.array(4)
    # 2. the MIL script with --tta 2 on two small subjects
    targs = dict(max_rotation_deg=8.0, max_translation=0.05, intensity_scale=0.1, intensity_shift=0.1, noise_std=0.02)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        manifest = write_synthetic_manifest(td / "vols", 2, shape=(48, 40, 36))
        args = ["--backbone", "resnet18", "--target-shape", "32", "32", "32", "--slice-axes", "0", "2", "--slice-counts", "3", "2",
                "--input-size", "64", "--batch-size", "4", "--tta", "2", "--max-rotation-deg", "8.0", "--noise-std", "0.02"]
        od = td / "out"
        argv = sys.argv
        sys.argv = ["build_resnet2d_mil_embeddings.py", "--manifest", str(manifest), "--out-dir", str(od)] + args
        try:
            runpy.run_path(str(REF / "scripts" / "build_resnet2d_mil_embeddings.py"), run_name="__main__")
        finally:
            sys.argv = argv
        npz = [p for p in od.iterdir() if p.suffix == ".npz"][0]
        d = np.load(npz, allow_pickle=True)
        out["script/emb"] = d["embeddings"]
        out["script/argv"] = np.array(args)
        out["script/files"] = np.array(sorted(p.name for p in od.iterdir()))
        import pandas as pd
        df = pd.read_csv(manifest)
        seeds = [abs(hash(str(s))) % (2 ** 32) for s in df["subject_id"]]
        out["script/seeds"] = np.array(seeds, dtype=np.int64)
        out["script/subject_ids"] = np.array([str(s) for s in df["subject_id"]])
        model, _, _ = of._build_resnet_backbone("resnet18")
        model.eval()
        for b, row in df.iterrows():
            vol = of._normalize_volume_for_resnet(of._load_volume(Path(row["t1wbrain_path"]), target_shape=(32, 32, 32)))
            sl = np.concatenate([of._select_slices(vol, 0, 3), of._select_slices(vol, 2, 2)], axis=0)
            r = np.random.default_rng(seeds[b])
            acc = None
            for p_ in range(2):
                aug = sl.copy()
                angle = r.uniform(-targs["max_rotation_deg"], targs["max_rotation_deg"])
                translate = r.uniform(-targs["max_translation"], targs["max_translation"], size=2)
                translate = translate * np.array([aug.shape[1], aug.shape[2]])
                for i in range(aug.shape[0]):
                    aug[i] = of._apply_affine_2d(aug[i], angle, translate)
                scale = 1.0 + r.uniform(-targs["intensity_scale"], targs["intensity_scale"])
                shift = r.uniform(-targs["intensity_shift"], targs["intensity_shift"])
                aug = aug * scale + shift
                aug = aug + r.normal(0.0, targs["noise_std"], size=aug.shape)
                aug = np.clip(aug, 0.0, 1.0).astype(np.float32, copy=False)
                out[f"script/aug/{b}/{p_}"] = aug
                with torch.no_grad():
                    emb = model(ref_input_tensor(aug, 64)).numpy()
                acc = emb if acc is None else acc + emb
            assert np.allclose(acc / 2, d["embeddings"][b], rtol=0, atol=1e-5), "restated TTA loop disagrees with the reference script"
        out["script/targs"] = np.array(json.dumps(targs))
    np.savez_compressed(GOLD / "tta.npz", **out)
    print("tta.npz", out["script/emb"].shape, seeds)


def gold_ft():
    """`MilAttentionFineTuneModel.predict_proba` of the reference (eval mode, no TTA) on path bags, a ready slice-array bag,
    a None bag and a masked-out bag; resnet18 backbone from the seeded builder, head weights stored."""
    from pd_fusion.models.mil_attention_finetune import MilAttentionFineTuneModel
    out = {}
    params = dict(backbone="resnet18", target_shape=(32, 32, 32), slice_axes=[0, 2], slice_counts=[3, 2], input_size=64, slice_batch_size=4,
                  hidden_dim=32, attn_dim=16, gated=True, pretrained=True, missing_prob=0.37)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        manifest = write_synthetic_manifest(td / "vols", 3, shape=(48, 40, 36))
        import pandas as pd
        paths = pd.read_csv(manifest)["t1wbrain_path"].tolist()
        torch.manual_seed(HEAD_SEED)
        m = MilAttentionFineTuneModel(params)          # backbone: patched builder -> manual_seed(1234); then the head is created
        rng = np.random.default_rng(3)
        arr_bag = rng.random((4, 32, 32)).astype(np.float32)
        bags = [paths[0], paths[1], None, arr_bag, paths[2]]
        masks = {"mri": np.array([1, 1, 1, 1, 0])}
        out["prob"] = np.asarray(m.predict_proba(bags, masks=masks), dtype=np.float64)
        out["arr_bag"] = arr_bag
        out["params"] = np.array(json.dumps(params))
        for k, v in _sd_np(m.attn.state_dict()).items():
            out[f"attn/{k}"] = v
        out["backbone_fingerprint"] = weight_fingerprint(m.backbone.state_dict())
    np.savez_compressed(GOLD / "ft.npz", **out)
    print("ft.npz", out["prob"])


def _sd_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def gold_heads():
    out = {}
    rng = np.random.default_rng(99)
    # ---- MIL heads (random init, seed 4321), gated (C3) and non-gated (C5), ragged bags + missing
    for tag, gated, D, H, A in [("mil_gated", True, 96, 32, 16), ("mil_plain", False, 64, 24, 8), ("mil_c3", True, 2048, 256, 128)]:
        torch.manual_seed(HEAD_SEED)
        m = MilAttentionModel(D, {"hidden_dim": H, "attn_dim": A, "gated": gated, "missing_prob": 0.5})
        lens = [48, 48, 7, 1, 48, 13] if D < 2048 else [48, 48, 48, 48]
        bags = [(rng.standard_normal((L, D)) * (1.0 if D < 2048 else 0.5)).astype(np.float32) for L in lens]
        masks = {"mri": np.array([1, 0, 1, 1, 1, 1][:len(bags)])}
        bl = list(bags)
        if D < 2048:
            bl[4] = None
        p = m.predict_proba(bl, masks=masks)
        for k, v in _sd_np(m.model.state_dict()).items():
            out[f"{tag}/sd/{k}"] = v
        for i, b in enumerate(bags):
            out[f"{tag}/bag{i}"] = b
        out[f"{tag}/none"] = np.array([b is None for b in bl])
        out[f"{tag}/mask_mri"] = masks["mri"]
        out[f"{tag}/prob"] = np.asarray(p, dtype=np.float64)
        out[f"{tag}/cfg"] = np.array(json.dumps(dict(gated=gated, D=D, H=H, A=A)))

    # ---- fusion heads trained BY THE REFERENCE on signal-bearing NaN-free tables, then its own sweep
    import yaml
    scen_cfg = yaml.safe_load((REF / "configs" / "eval_missingness.yaml").read_text())
    scen_cfg = {"scenarios": scen_cfg["scenarios"]}
    extra = [{"name": "mri_missing_50", "drop_modalities": ["mri"], "drop_rate": 0.5}]
    scen_cfg["scenarios"] = scen_cfg["scenarios"] + extra
    out["scenarios"] = np.array(json.dumps(scen_cfg))
    dims = {"clinical": 10, "datspect": 5, "mri": 20}
    df, masks = synthetic_table(400, dims, seed=42, mask_seed=7)
    out["table/n"], out["table/dims"] = np.array(400), np.array(json.dumps(dims))
    y = df["diagnosis"].values
    feat_cols = [c for m in ["clinical", "datspect", "mri"] for c in df.columns if c.startswith(m + "_")]
    from pd_fusion.data.preprocess import preprocess_features
    X, _, scaler = preprocess_features(df, feat_cols)

    # ModDrop
    torch.manual_seed(HEAD_SEED); np.random.seed(HEAD_SEED)
    md = ModalityDropoutModel(dims, {"hidden_dims": [64, 32], "dropout": 0.2, "lr": 1e-3, "batch_size": 32, "epochs": 15, "moddrop_rate": 0.3})
    md.train(X, y)
    for k, v in _sd_np(md.model.state_dict()).items():
        out[f"moddrop/sd/{k}"] = v
    np.random.seed(11)
    res = evaluate_model(md, df, masks, (None, scaler, feat_cols), scen_cfg)
    out["moddrop/metrics"] = np.array(json.dumps({k: {m: float(v) for m, v in d.items()} for k, d in res.items()}))
    np.random.seed(11)
    probs, mk = [], []
    for sc in scen_cfg["scenarios"]:
        cur = apply_missingness_scenario(df, sc, masks)
        from pd_fusion.data.feature_utils import apply_masks_to_matrix
        Xs, _, _ = preprocess_features(df, feat_cols, None, scaler)
        probs.append(md.predict_proba(apply_masks_to_matrix(Xs, cur, feat_cols), masks=cur))
        mk.append(np.stack([cur[m] for m in ["clinical", "datspect", "mri"]], axis=1))
    out["moddrop/probs"] = np.stack(probs).astype(np.float32)
    out["masks_seed11"] = np.stack(mk).astype(np.uint8)

    # MoE (imaging + clinical, BASELINE config 4 shape class) -- per-modality preprocessors
    mods = ["clinical", "mri"]
    prep, Xd = {}, {}
    for m in mods:
        cols = [c for c in df.columns if c.startswith(m + "_")]
        Xm, _, sc_m = preprocess_features(df, cols)
        prep[m] = (None, sc_m, cols)
        Xd[m] = torch.FloatTensor(Xm * masks[m].reshape(-1, 1))
    mt = torch.FloatTensor(np.stack([masks[m] for m in mods], axis=1))
    torch.manual_seed(HEAD_SEED)
    moe = MoEModel({m: Xd[m].shape[1] for m in mods}, {"expert_hidden_dims": [32, 16], "router_hidden_dims": [16], "lr": 3e-3, "epochs": 60})
    moe.train(Xd, y, mt)
    for k, v in _sd_np(moe.model.state_dict()).items():
        out[f"moe/sd/{k}"] = v
    masks2 = {m: masks[m] for m in mods}
    np.random.seed(12)
    res = evaluate_model(moe, df, masks2, prep, scen_cfg)
    out["moe/metrics"] = np.array(json.dumps({k: {m: float(v) for m, v in d.items()} for k, d in res.items()}))
    np.random.seed(12)
    probs, mk = [], []
    for sc in scen_cfg["scenarios"]:
        cur = apply_missingness_scenario(df, sc, masks2)
        Xs = {}
        for m in mods:
            Xm, _, _ = preprocess_features(df, prep[m][2], None, prep[m][1])
            Xs[m] = torch.FloatTensor(Xm * cur[m].reshape(-1, 1))
        probs.append(moe.predict_proba(Xs, torch.FloatTensor(np.stack([cur[m] for m in mods], axis=1))))
        mk.append(np.stack([cur[m] for m in mods], axis=1))
    out["moe/probs"] = np.stack(probs).astype(np.float32)
    out["moe/masks_seed12"] = np.stack(mk).astype(np.uint8)
    out["moe/mods"] = np.array(mods)
    np.savez_compressed(GOLD / "heads.npz", **out)
    print("heads.npz moddrop", json.loads(str(out["moddrop/metrics"]))["full_observation"]["roc_auc"],
          "moe", json.loads(str(out["moe/metrics"]))["full_observation"]["roc_auc"])


def gold_c45():
    """BASELINE configs 4 and 5 at their full sizes.
    C5: the UNMODIFIED MIL script on one synthetic 256x256x176 volume with resnet50, slice axes 0/1/2 x 24 (L = 72), --tta 2,
        batch size 8, the augmentation amplitudes of configs/data_openneuro_ds001907_resnet2d_mil_multi.yaml.
    C4: MoE over imaging (512) + clinical (10) features, N = 10 000 subjects, trained BY THE REFERENCE, its own evaluate_model sweep."""
    out = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        manifest = write_synthetic_manifest(td / "vols", 1, shape=(256, 256, 176), start=21)
        args = ["--backbone", "resnet50", "--target-shape", "160", "160", "160", "--slice-axes", "0", "1", "2", "--slice-counts", "24", "24", "24",
                "--input-size", "224", "--batch-size", "8", "--tta", "2", "--max-rotation-deg", "8.0", "--max-translation", "0.05",
                "--intensity-scale", "0.1", "--intensity-shift", "0.1", "--noise-std", "0.02"]
        od = td / "out"
        argv = sys.argv
        sys.argv = ["build_resnet2d_mil_embeddings.py", "--manifest", str(manifest), "--out-dir", str(od)] + args
        try:
            runpy.run_path(str(REF / "scripts" / "build_resnet2d_mil_embeddings.py"), run_name="__main__")
        finally:
            sys.argv = argv
        d = np.load([p for p in od.iterdir() if p.suffix == ".npz"][0], allow_pickle=True)
        import pandas as pd
        df = pd.read_csv(manifest)
        out["c5/emb"] = d["embeddings"].astype(np.float32)
        out["c5/argv"] = np.array(args)
        out["c5/seeds"] = np.array([abs(hash(str(s))) % (2 ** 32) for s in df["subject_id"]], dtype=np.int64)
        out["c5/targs"] = np.array(json.dumps(dict(max_rotation_deg=8.0, max_translation=0.05, intensity_scale=0.1, intensity_shift=0.1, noise_std=0.02)))
        print("c5", out["c5/emb"].shape, float(np.abs(out["c5/emb"]).mean()), out["c5/seeds"])

    import yaml
    scen_cfg = {"scenarios": yaml.safe_load((REF / "configs" / "eval_missingness.yaml").read_text())["scenarios"]}
    out["c4/scenarios"] = np.array(json.dumps(scen_cfg))
    dims = {"clinical": 10, "datspect": 0, "mri": 512}
    N = 10000
    df, masks = synthetic_table(N, dims, seed=44, mask_seed=9)
    y = df["diagnosis"].values
    from pd_fusion.data.preprocess import preprocess_features
    mods = ["clinical", "mri"]
    prep, Xd = {}, {}
    for m in mods:
        cols = [c for c in df.columns if c.startswith(m + "_")]
        Xm, _, sc_m = preprocess_features(df, cols)
        prep[m] = (None, sc_m, cols)
        Xd[m] = torch.FloatTensor(Xm * masks[m].reshape(-1, 1))
    mt = torch.FloatTensor(np.stack([masks[m] for m in mods], axis=1))
    torch.manual_seed(HEAD_SEED)
    params = yaml.safe_load((REF / "configs" / "model_moe.yaml").read_text())["params"]
    params = dict(params, epochs=40)
    moe = MoEModel({m: Xd[m].shape[1] for m in mods}, params)
    moe.train(Xd, y, mt)
    for k, v in _sd_np(moe.model.state_dict()).items():
        out[f"c4/sd/{k}"] = v
    masks2 = {m: masks[m] for m in mods}
    np.random.seed(13)
    res = evaluate_model(moe, df, masks2, prep, scen_cfg)
    out["c4/metrics"] = np.array(json.dumps({k: {m: float(v) for m, v in d.items()} for k, d in res.items()}))
    np.random.seed(13)
    probs, mk = [], []
    for sc in scen_cfg["scenarios"]:
        cur = apply_missingness_scenario(df, sc, masks2)
        Xs = {}
        for m in mods:
            Xm, _, _ = preprocess_features(df, prep[m][2], None, prep[m][1])
            Xs[m] = torch.FloatTensor(Xm * cur[m].reshape(-1, 1))
        probs.append(moe.predict_proba(Xs, torch.FloatTensor(np.stack([cur[m] for m in mods], axis=1))))
        mk.append(np.stack([cur[m] for m in mods], axis=1))
    out["c4/probs"] = np.stack(probs).astype(np.float32)
    out["c4/masks_seed13"] = np.packbits(np.stack(mk).astype(np.uint8), axis=1)       # [S, ceil(N/8), M]
    out["c4/n"], out["c4/dims"], out["c4/mods"] = np.array(N), np.array(json.dumps(dims)), np.array(mods)
    np.savez_compressed(GOLD / "c45.npz", **out)
    print("c45.npz moe auc", json.loads(str(out["c4/metrics"]))["full_observation"]["roc_auc"], out["c4/probs"].shape)


def gold_simple():
    """The reference's `_compute_simple_features` and `load_simple_features` (data/openneuro_features.py:34-104) on seeded
    synthetic volumes: the features of single resampled volumes (incl. a volume without positive voxels and a constant one) and
    the parquet of a 3-row manifest (256x256x176 -> 96^3, default config + extra_stats)."""
    out = {}
    cases = {"brain48": (synthetic_volume(3, (64, 48, 44)), (48, 48, 48)), "brain96": (synthetic_volume(5, (120, 110, 100)), (96, 96, 96))}
    rng = np.random.default_rng(11)
    cases["nonpositive"] = ((-rng.random((40, 40, 40))).astype(np.float32), (32, 32, 32))
    cases["constant"] = (np.full((36, 36, 36), 2.5, dtype=np.float32), (24, 24, 24))
    cases["sparse"] = ((rng.random((50, 40, 30)) > 0.97).astype(np.float32) * rng.random((50, 40, 30)).astype(np.float32) * 900, (40, 40, 40))
    for tag, (raw, tgt) in cases.items():
        vol = _load_volume_npy_arr(raw, tgt)
        out[f"{tag}/raw"] = raw
        out[f"{tag}/target"] = np.array(tgt)
        for bins, grid, extra in ((10, 8, False), (16, 4, True), (10, 0, True)):
            out[f"{tag}/feats_{bins}_{grid}_{int(extra)}"] = of._compute_simple_features(vol, hist_bins=bins, grid_size=grid, extra_stats=extra)
    of._load_volume = _load_volume_npy
    with tempfile.TemporaryDirectory() as td:
        manifest = write_synthetic_manifest(Path(td) / "vols", 3, start=40)
        cfg = {"hist_bins": 10, "grid_size": 8, "target_shape": [96, 96, 96], "extra_stats": True}
        df = of.load_simple_features(manifest, Path(td) / "cache", cfg)
        out["manifest/start"], out["manifest/n"] = np.array(40), np.array(3)
        out["manifest/cfg"] = np.array(json.dumps(cfg))
        out["manifest/columns"] = np.array(list(df.columns))
        out["manifest/feats"] = df[[c for c in df.columns if c.startswith("mri_feat_")]].values
        out["manifest/file"] = np.array(sorted(p.name for p in (Path(td) / "cache").iterdir())[0].split("_")[0])
    np.savez_compressed(GOLD / "simple.npz", **out)
    print("simple.npz", {k: v.shape for k, v in out.items() if k.endswith("_0") or k == "manifest/feats"})


def gold_cnn3d():
    """The UNMODIFIED reference script scripts/build_cnn3d_embeddings.py on six small synthetic volumes (CPU): the only patch is a
    stand-in `nibabel` module whose load(path).get_fdata() returns the .npy array as float64 (nibabel is not installed here; the
    script imports it at module top).  Stores the embeddings parquet's values and the state the run started from / ended with."""
    import types
    nib = types.ModuleType("nibabel")

    class _Img:
        def __init__(self, p):
            self._p = p

        def get_fdata(self):
            return np.load(self._p).astype(np.float64)

    nib.load = lambda p: _Img(p)
    sys.modules["nibabel"] = nib
    out = {}
    shape, target, n, epochs, bs, seed = (40, 36, 32), [16, 16, 16], 6, 2, 4, 42
    with tempfile.TemporaryDirectory() as td:
        manifest = write_synthetic_manifest(Path(td) / "vols", n, shape=shape, start=60)
        od = Path(td) / "out"
        args = ["--target-shape"] + [str(v) for v in target] + ["--embedding-dim", "32", "--epochs", str(epochs), "--batch-size", str(bs),
                                                                  "--lr", "0.001", "--seed", str(seed)]
        argv = sys.argv
        sys.argv = ["build_cnn3d_embeddings.py", "--manifest", str(manifest), "--out-dir", str(od)] + args
        try:
            runpy.run_path(str(REF / "scripts" / "build_cnn3d_embeddings.py"), run_name="__main__")
        finally:
            sys.argv = argv
        pq = sorted(od.glob("*.parquet"))[0]
        import pandas as pd
        df = pd.read_parquet(pq)
        out["columns"] = np.array(list(df.columns))
        out["emb"] = df[[c for c in df.columns if c.startswith("mri_cnn_")]].values.astype(np.float32)
        out["file_prefix"] = np.array(pq.name.split("_")[0])
        out["argv"] = np.array(args)
        out["spec"] = np.array(json.dumps({"shape": shape, "target": target, "n": n, "start": 60, "epochs": epochs, "batch_size": bs, "seed": seed,
                                           "embedding_dim": 32}))
    np.savez_compressed(GOLD / "cnn3d.npz", **out)
    print("cnn3d.npz", out["emb"].shape, float(np.abs(out["emb"]).mean()))


def _load_volume_npy_arr(raw, target_shape):
    d = np.nan_to_num(raw.astype(np.float32), nan=0.0, posinf=0.0, neginf=0.0)
    return ndimage.zoom(d, [t / s for t, s in zip(target_shape, d.shape)], order=1)


if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    which = sys.argv[1:] or ["preproc", "embed", "scripts", "heads", "tta", "ft", "c45", "simple", "cnn3d"]
    os.environ.setdefault("PYTHONHASHSEED", "0")
    for w in which:
        globals()[f"gold_{w}"]()
