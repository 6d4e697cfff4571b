"""Host-side logic (no GPU): scenario masks vs the reference's golden masks, feature bookkeeping, sharding,
the drop-in cache-name contract, NIfTI decode, FLOP accounting."""
import gzip
import json
import struct
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from oracle import oracle as O
from pd_fusion_b200.backbone import ResNet2D, conv_list, flops_per_image
from pd_fusion_b200.data import feature_utils as FU
from pd_fusion_b200.data.missingness import apply_missingness_scenario, get_modality_mask_matrix, scenario_mask_tensor
from pd_fusion_b200.data.preprocess import preprocess_features
from pd_fusion_b200.parallel import shard_range
from pd_fusion_b200.synthetic import synthetic_table


def test_masks_bit_exact_vs_reference(golden):
    g = golden("heads")
    dims = json.loads(str(g["table/dims"]))
    df, masks = synthetic_table(int(g["table/n"]), dims, seed=42, mask_seed=7)
    scen = json.loads(str(g["scenarios"]))["scenarios"]
    np.random.seed(11)
    mk, per = scenario_mask_tensor(df, scen, masks, O.MODALITIES)
    assert mk.dtype == np.uint8 and np.array_equal(mk, g["masks_seed11"])
    np.random.seed(11)
    seq = np.stack([get_modality_mask_matrix(apply_missingness_scenario(df, sc, masks)) for sc in scen])
    assert np.array_equal(seq.astype(np.uint8), g["masks_seed11"])
    # a modality that is not in the mask dict is a no-op; the input dict is never mutated
    before = {k: v.copy() for k, v in masks.items()}
    out = apply_missingness_scenario(df, {"name": "x", "drop_modalities": ["nope", "mri"]}, masks)
    assert all(np.array_equal(before[k], masks[k]) for k in masks) and out["mri"].sum() == 0


def test_feature_bookkeeping():
    df, _ = synthetic_table(10, {"clinical": 3, "datspect": 2, "mri": 4})
    cols = FU.get_all_feature_cols(df)
    assert cols[:3] == ["clinical_f0", "clinical_f1", "clinical_f2"] and len(cols) == 9
    sl = FU.get_feature_slices(cols)
    assert sl["mri"] == [5, 6, 7, 8]
    X, _, sc = preprocess_features(df, cols)
    Xm = FU.apply_masks_to_matrix(X, {"mri": np.array([0] * 5 + [1] * 5)}, cols)
    assert np.all(Xm[:5, 5:] == 0) and np.array_equal(Xm[5:], X[5:]) and np.array_equal(Xm[:, :5], X[:, :5])
    X2, _, _ = preprocess_features(df, cols, None, sc)
    assert np.array_equal(X, X2)


@pytest.mark.parametrize("n,ws", [(10, 4), (5, 4), (0, 2), (8, 8), (10000, 8), (3, 1)])
def test_shard_ranges_tile_rows_in_order(n, ws):
    spans = [shard_range(n, r, ws) for r in range(ws)]
    flat = [i for a, b in spans for i in range(a, b)]
    assert flat == list(range(n))
    per = -(-n // ws) if n else 0
    assert all(b - a <= per for a, b in spans)


def test_cache_name_contract(golden):
    from pd_fusion_b200.data.openneuro_features import _hash_config
    g = golden("scripts")
    for tag, prefix in (("c2", "resnet2d_"), ("mil", "resnet2d_mil_")):
        cfg = json.loads(str(g[f"{tag}/json"]))["config"]
        name = [str(f) for f in g[f"{tag}/files"] if str(f).endswith(".json")][0]
        assert name.endswith(f"_{_hash_config(cfg)}.json") and name.startswith(prefix)


def test_nifti_reader_roundtrip(tmp_path):
    from pd_fusion_b200.data.openneuro_features import _read_volume_host
    vol = (np.arange(5 * 4 * 3, dtype=np.int16).reshape(5, 4, 3) - 7)
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, 5, 4, 3, 1, 1, 1, 1)
    struct.pack_into("<h", hdr, 70, 4)
    struct.pack_into("<h", hdr, 72, 16)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 2.0, 1.5)
    p = tmp_path / "v.nii.gz"
    with gzip.open(p, "wb") as f:
        f.write(bytes(hdr) + vol.tobytes(order="F"))
    got = _read_volume_host(p)
    assert got.dtype == np.float32 and got.shape == (5, 4, 3) and np.array_equal(got, (vol * 2.0 + 1.5).astype(np.float32))
    np.save(tmp_path / "v.npy", vol.astype(np.float64))
    assert np.array_equal(_read_volume_host(tmp_path / "v.npy"), vol.astype(np.float32))
    # the stored form (what the batched builders upload and decode on the device) describes the same array
    from oracle import oracle as O
    from pd_fusion_b200.data.openneuro_features import _read_volume_stored
    sv = _read_volume_stored(p)
    assert (sv.code, sv.shape, sv.fortran, sv.slope, sv.inter) == (4, (5, 4, 3), True, 2.0, 1.5) and sv.voxels.dtype == np.int16
    assert np.array_equal(sv.to_float32(), got)
    assert np.array_equal(O.decode_stored_voxels(sv.voxels, sv.shape, sv.fortran, sv.slope, sv.inter), got)
    sn = _read_volume_stored(tmp_path / "v.npy")
    assert (sn.code, sn.fortran) == (64, False) and np.array_equal(sn.to_float32(), vol.astype(np.float32))
    # big-endian file, slope 0 (= "no scaling" in NIfTI)
    be = bytearray(352)
    struct.pack_into(">i", be, 0, 348)
    struct.pack_into(">8h", be, 40, 3, 5, 4, 3, 1, 1, 1, 1)
    struct.pack_into(">h", be, 70, 512)
    struct.pack_into(">h", be, 72, 16)
    struct.pack_into(">f", be, 108, 352.0)
    struct.pack_into(">2f", be, 112, 0.0, 9.0)
    u16 = (np.arange(60, dtype=np.uint16) * 1000).reshape(5, 4, 3)
    pb = tmp_path / "be.nii"
    pb.write_bytes(bytes(be) + u16.astype(">u2").tobytes(order="F"))
    sb = _read_volume_stored(pb)
    assert sb.code == 512 and sb.voxels.dtype == np.dtype("<u2") and np.array_equal(sb.to_float32(), u16.astype(np.float32))
    assert np.array_equal(_read_volume_host(pb), u16.astype(np.float32))


def test_read_ahead_keeps_manifest_order_and_surfaces_errors(tmp_path):
    """The builders read volume files with worker threads, a window ahead of the GPU: rows come back in manifest order whatever
    the completion order, released rows are dropped, and a reader's exception reaches the caller (as the reference's would)."""
    from pd_fusion_b200.data.openneuro_features import _ReadAhead
    paths = []
    for i in range(9):
        p = tmp_path / f"v{i}.npy"
        np.save(p, (np.arange(24, dtype=np.float32) + i).reshape(2, 3, 4))
        paths.append(str(p))
    r = _ReadAhead(paths, 2, 9, window=3)
    seen, i = [], 2
    while i < 9:
        j = min(i + 2, 9)
        peek = r.get(j) if j < 9 else None            # the batching loop peeks at the first row of the next batch ...
        seen += [float(r.get(k).voxels[0]) for k in range(i, j)]
        r.release(i, j)
        assert peek is None or r.get(j) is peek       # ... and finds the same object when that batch starts
        i = j
    assert seen == [2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0] and not r.futures
    bad = _ReadAhead(paths + [str(tmp_path / "missing.nii.gz")], 9, 10, window=2)
    with pytest.raises(FileNotFoundError):
        bad.get(9)


def test_parallel_npz_writer_is_read_back_like_numpys(tmp_path):
    """The MIL cache is written by `savez_compressed_parallel` (deflate spread over threads, pigz-style): every reader sees what
    np.savez_compressed would have written -- same members, dtypes, values (object arrays included), a sound ZIP."""
    import zipfile
    from pd_fusion_b200.utils import npz_writer
    rng = np.random.default_rng(3)
    emb = np.maximum(rng.standard_normal((37, 24, 512)).astype(np.float32), 0)            # 1.8 MB
    big = rng.integers(0, 1000, size=(5, 1 << 20), dtype=np.int32)                      # 20 MB: several chunks ...
    odd = rng.standard_normal(2 * (8 << 20) // 8 + 12345)                                # ... and one that ends mid-chunk
    ids = np.array([f"sub-{i:05d}" for i in range(37)], dtype=object)
    kw = dict(embeddings=emb, big=big, odd=odd, subject_id=ids, session=np.ones(37, dtype=int), label=np.arange(37) % 2, empty=np.zeros((0, 3)))
    p = npz_writer.savez_compressed_parallel(tmp_path / "cache", threads=3, **kw)
    assert p.name == "cache.npz" and zipfile.ZipFile(p).testzip() is None
    np.savez_compressed(tmp_path / "ref.npz", **kw)
    a, b = np.load(tmp_path / "ref.npz", allow_pickle=True), np.load(p, allow_pickle=True)
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    assert p.stat().st_size < 1.02 * (tmp_path / "ref.npz").stat().st_size              # independent chunks cost next to nothing
    for m in zipfile.ZipFile(p).infolist():
        assert m.compress_type == zipfile.ZIP_DEFLATED


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` needs no GPU (it times the reference's CPU path): one JSON line with the arm's keys."""
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mri_subjects_per_sec_resnet2d_embed_fuse" and line["unit"] == "subjects/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "subjects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "configs/openneuro_ds001907_resnet2d_mil.yaml"      # the metric's own config is the default


def test_flop_accounting_matches_survey():
    assert int(flops_per_image("resnet18")) == 3627122688 and int(flops_per_image("resnet50")) == 8174272512
    assert int(flops_per_image("resnet18") - flops_per_image("resnet18", folded_stem=True)) == 157351936
    assert len(conv_list("resnet18")) == 20 and len(conv_list("resnet50")) == 53


def test_resnet2d_matches_torchvision_init():
    import torch
    tv = pytest.importorskip("torchvision")
    for arch in ("resnet18", "resnet50"):
        torch.manual_seed(1234)
        a = ResNet2D(arch).state_dict()
        torch.manual_seed(1234)
        b = getattr(tv.models, arch)(weights=None).state_dict()
        assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in b)


def test_evaluate_dispatch_with_a_stub_model():
    """evaluate_model's dispatch on the type of prep_info and its scenario/RNG order, with a model that needs no GPU."""
    from pd_fusion_b200.evaluation.evaluate import evaluate_model, predict_proba_for_scenario

    class Stub:
        def __init__(self):
            self.calls = []

        def predict_proba(self, X, masks=None):
            self.calls.append({k: np.array(v) for k, v in masks.items()})
            return 1.0 / (1.0 + np.exp(-np.nan_to_num(X).sum(axis=1)))

    dims = {"clinical": 3, "datspect": 2, "mri": 4}
    df, masks = synthetic_table(64, dims)
    cols = FU.get_all_feature_cols(df)
    _, _, sc = preprocess_features(df, cols)
    cfg = {"scenarios": [{"name": "full", "drop_modalities": []}, {"name": "half", "drop_modalities": ["mri"], "drop_rate": 0.5},
                         {"name": "r1", "type": "random", "n_drop": 1}]}
    m = Stub()
    np.random.seed(5)
    res = evaluate_model(m, df, masks, (None, sc, cols), cfg)
    assert list(res) == ["full", "half", "r1"] and set(res["full"]) == {"roc_auc", "pr_auc", "balanced_accuracy", "f1", "brier_score", "ece"}
    np.random.seed(5)
    want = [O.apply_missingness_scenario(len(df), s, masks) for s in cfg["scenarios"]]
    for got, w in zip(m.calls, want):
        assert all(np.array_equal(got[k], w[k]) for k in w)
    y, p = predict_proba_for_scenario(m, df, masks, (None, sc, cols), cfg["scenarios"][0])
    assert len(y) == len(p) == 64


@pytest.mark.parametrize("arch,chunks", [("resnet18", [2, 2, 3, 4, 5]), ("resnet18", [5, 5, 5, 5, 5]), ("resnet50", [1, 1, 2, 4, 3])])
def test_lowered_op_list_computes_the_network(arch, chunks):
    """The depth-first, chunked op list (backbone.lower_resnet) is executed by a tiny CPU interpreter over byte buffers
    and must reproduce torch's forward of the same state_dict: checks buffer wiring, chunk offsets and scratch reuse
    (the kernels themselves are checked on the GPU)."""
    import torch
    import torch.nn.functional as F

    from pd_fusion_b200 import _lib
    from pd_fusion_b200.backbone import lower_resnet

    n, S = 5, 32
    torch.manual_seed(5)
    net = ResNet2D(arch).eval()
    with torch.no_grad():           # make BN non-trivial
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1)
    sd = net.state_dict()
    ops, extents, _ = lower_resnet(arch, n, S, bf16=False, fused_stem=False, chunks=chunks)
    bufs = {k: np.zeros(v, dtype=np.uint8) for k, v in extents.items()}
    x = torch.randn(n, 3, S, S)
    bufs["input"][:] = np.frombuffer(x.permute(0, 2, 3, 1).contiguous().numpy().tobytes(), dtype=np.uint8)

    def view(ref, shape):
        buf, off = ref
        cnt = int(np.prod(shape))
        return bufs[buf][off:off + 4 * cnt].view(np.float32).reshape(shape)

    bn = {cv["name"]: cv["bn"] for cv in conv_list(arch)}
    for name, f, refs, wname in ops:
        k = f["kind"]
        xin = torch.from_numpy(view(refs["d_in"], (f["n"], f["h"], f["w"], f["c"])).copy()).permute(0, 3, 1, 2)
        if k == _lib.OP_CONV:
            w = sd[wname + ".weight"]
            g, b, mu, var = (sd[bn[wname] + s] for s in (".weight", ".bias", ".running_mean", ".running_var"))
            y = F.conv2d(xin, w, stride=f["stride"], padding=f["pad"])
            sc = g / torch.sqrt(var + 1e-5)
            y = y * sc.view(1, -1, 1, 1) + (b - mu * sc).view(1, -1, 1, 1)
            if "d_residual" in refs:
                y = y + torch.from_numpy(view(refs["d_residual"], (f["n"], f["ho"], f["wo"], f["k"])).copy()).permute(0, 3, 1, 2)
            if f["relu"]:
                y = torch.relu(y)
            view(refs["d_out"], (f["n"], f["ho"], f["wo"], f["k"]))[:] = y.permute(0, 2, 3, 1).numpy()
        elif k == _lib.OP_MAXPOOL:
            y = F.max_pool2d(xin, 3, 2, 1)
            view(refs["d_out"], (f["n"], f["ho"], f["wo"], f["c"]))[:] = y.permute(0, 2, 3, 1).numpy()
        elif k == _lib.OP_AVGPOOL:
            view(refs["d_out"], (f["n"], f["c"]))[:] = xin.mean(dim=(2, 3)).numpy()
        else:
            raise AssertionError(k)
    net.fc = torch.nn.Identity()
    with torch.no_grad():
        ref = net(x).numpy()
    out = bufs["output"].view(np.float32).reshape(n, -1)
    assert np.abs(out - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_lowering_bf16_buffers_are_written_before_read():
    """Structural check of the bf16 / fused-stem op list: every byte an op reads was written by an earlier op -- also when the
    BasicBlock downsample convolutions are computed inside conv1's launch (`dual_stages`)."""
    from pd_fusion_b200.backbone import lower_resnet
    for arch in ("resnet18", "resnet50"):
        for fused in (True, False):
            for dual, chain in (((), ()), ((2, 3, 4), ()), ((), (1, 2, 3))):
                n, S = 7, 64
                ops, extents, _ = lower_resnet(arch, n, S, bf16=True, fused_stem=fused, chunks=[2, 2, 4, 4, 7], dual_stages=dual,
                                               chain_stages=chain)
                chained = {o[1]["_weight3"] for o in ops if "_weight3" in o[1]}
                if arch == "resnet50" and chain:         # conv1 of every block of stages 1..3 but the very first rides on the conv3 before
                    emitted = {o[0] for o in ops if o[0].endswith(".conv1")}          # it, ACROSS the stage boundaries 1->2 and 2->3 too
                    assert emitted == {"layer1.0.conv1", "layer4.0.conv1", "layer4.1.conv1", "layer4.2.conv1"}
                    assert len(chained) == 12 and {"layer2.0.conv1", "layer3.0.conv1"} <= chained
                else:
                    assert not chained
                n_down = sum("downsample" in o[0] for o in ops)
                n_dual = sum("_weight2" in o[1] for o in ops)
                if arch == "resnet18" and dual:          # BasicBlock: the three downsample convs ride inside conv1
                    assert n_down == 0 and n_dual > 0 and all("d_out2" in o[2] for o in ops if "_weight2" in o[1])
                else:                                    # Bottleneck blocks (stride on conv2) cannot share tiles: nothing fused
                    assert n_dual == 0 and n_down > 0
                written = {k: np.zeros(v, dtype=bool) for k, v in extents.items()}
                written["input"][:] = True
                for name, f, refs, _ in ops:
                    esz = 2
                    in_elems = f["n"] * f["h"] * f["w"] * f["c"]
                    in_bytes = in_elems * (4 if (f["kind"] == 2 and f.get("out_f32")) else esz)
                    buf, off = refs["d_in"]
                    assert written[buf][off:off + in_bytes].all(), (arch, fused, name, "input not produced")
                    if "d_residual" in refs:
                        rb, ro = refs["d_residual"]
                        assert written[rb][ro:ro + f["n"] * f["ho"] * f["wo"] * f["k"] * esz].all(), (arch, name, "residual")
                    ob, oo = refs["d_out"]
                    if f["kind"] == 2:
                        nbytes = f["n"] * f["c"] * 4
                    elif f["kind"] in (1,):
                        nbytes = f["n"] * f["ho"] * f["wo"] * f["c"] * esz
                    else:
                        nbytes = f["n"] * f["ho"] * f["wo"] * f["k"] * (4 if f.get("out_f32") else esz)
                    assert (ob, oo) != refs["d_in"] and oo + nbytes <= extents[ob]
                    written[ob][oo:oo + nbytes] = True
                    if "d_out3" in refs:                 # chained conv1 of the next block / next stage
                        b3, o3 = refs["d_out3"]
                        n3 = f["n"] * f["ho"] * f["wo"] * f["k3"] * esz
                        assert (b3, o3) not in (refs["d_in"], refs["d_out"], refs["d_residual"]) and o3 + n3 <= extents[b3]
                        written[b3][o3:o3 + n3] = True
                    if "d_out2" in refs:
                        b2, o2 = refs["d_out2"]
                        n2 = f["n"] * f["ho"] * f["wo"] * f["k"] * esz
                        assert (b2, o2) not in (refs["d_in"], refs["d_out"]) and o2 + n2 <= extents[b2]
                        written[b2][o2:o2 + n2] = True
                assert written["output"].all()


def test_load_openneuro_ds001907_modes(tmp_path, monkeypatch):
    """Cache consumer (reference: data/openneuro_ds001907.py:17-82): env override of the manifest, feature_mode dispatch onto the
    cache files the builders write, `diagnosis` from `label`, masks clinical = datspect = 0 and mri from the data."""
    import pandas as pd
    from pd_fusion_b200.data.openneuro_ds001907 import load_openneuro_ds001907
    from pd_fusion_b200.data.openneuro_features import _hash_config, _hash_file
    man = tmp_path / "m.csv"
    pd.DataFrame({"subject_id": ["a", "b", "c"], "session": [1, 1, 2], "label": [0, 1, 1], "t1wbrain_path": ["x", "y", "z"]}).to_csv(man, index=False)
    cfg = {"backbone": "resnet18", "slice_count": 4}
    cache = tmp_path / "cache"
    cache.mkdir()
    # resnet2d: parquet with mri_resnet_* columns (one subject without features)
    df = pd.DataFrame({"subject_id": ["a", "b", "c"], "session": [1, 1, 2], "label": [0, 1, 1],
                       "mri_resnet_0": [0.1, np.nan, 0.3], "mri_resnet_1": [1.0, np.nan, 3.0]})
    df.to_parquet(cache / f"resnet2d_{_hash_file(man)}_{_hash_config(cfg)}.parquet", index=False)
    monkeypatch.setenv("PD_FUSION_DS001907_MANIFEST", str(man))
    out, masks = load_openneuro_ds001907({"manifest_path": "does/not/exist.csv", "feature_mode": "resnet2d", "resnet2d_cache_dir": str(cache),
                                          "resnet2d_config": cfg})
    assert list(out["diagnosis"]) == [0, 1, 1]
    assert masks["mri"].tolist() == [1, 0, 1] and masks["clinical"].tolist() == [0, 0, 0] and masks["datspect"].tolist() == [0, 0, 0]
    # resnet2d_mil: npz with [S, L, D] bags
    emb = np.arange(3 * 2 * 5, dtype=np.float32).reshape(3, 2, 5)
    np.savez_compressed(cache / f"resnet2d_mil_{_hash_file(man)}_{_hash_config(cfg)}.npz", embeddings=emb,
                        subject_id=np.array(["a", "b", "c"], dtype=object), session=np.array([1, 1, 2]), label=np.array([0, 1, 1]))
    out, masks = load_openneuro_ds001907({"feature_mode": "resnet2d_mil", "resnet2d_cache_dir": str(cache), "resnet2d_config": cfg})
    assert np.array_equal(out["mri_mil"][1], emb[1]) and masks["mri"].tolist() == [1, 1, 1]
    # fine-tune mode: the manifest itself, paths in mri_mil
    out, masks = load_openneuro_ds001907({"feature_mode": "resnet2d_mil_ft"})
    assert list(out["mri_mil"]) == ["x", "y", "z"] and masks["mri"].tolist() == [1, 1, 1]
    with pytest.raises(ValueError):
        load_openneuro_ds001907({"feature_mode": "nope"})
    with pytest.raises(FileNotFoundError):
        load_openneuro_ds001907({"feature_mode": "cnn3d", "embedding_cache_dir": str(cache)})
    monkeypatch.delenv("PD_FUSION_DS001907_MANIFEST")
    with pytest.raises(FileNotFoundError):
        load_openneuro_ds001907({"manifest_path": str(tmp_path / "missing.csv"), "feature_mode": "resnet2d"})
