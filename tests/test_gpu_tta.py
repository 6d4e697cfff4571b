"""a6 on the GPU: test-time augmentation (affine + intensity + noise + clip) vs outputs of the reference (tests/golden/tta.npz)
and the oracle."""
import json
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
import pd_fusion_b200.data.openneuro_features as of
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D
from pd_fusion_b200.data.tta import draw_passes, params_bytes, tta_config
from pd_fusion_b200.preprocess import VolumePreprocessor
from pd_fusion_b200.synthetic import synthetic_volume, write_synthetic_manifest


def test_apply_affine_2d_bit_exact(golden):
    g = golden("tta")
    for k in range(int(g["affine/n"])):
        got = of._apply_affine_2d(g[f"affine/{k}/img"], float(g[f"affine/{k}/angle"]), g[f"affine/{k}/translate"])
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), g[f"affine/{k}/out"].view(np.uint32)), k


def test_tta_passes_bit_exact(golden):
    """slices -> augmented slices for both subjects and both passes of the reference run, and the resized network input
    against the oracle."""
    g = golden("tta")
    targs = tta_config(json.loads(str(g["script/targs"])))
    raws = np.stack([synthetic_volume(b, (48, 40, 36)) for b in range(2)])
    pre = VolumePreprocessor(raws[0].shape, (32, 32, 32), [0, 2], [3, 2], 64, out_mode=_lib.OUT_F32_NHWC3, max_batch=2)
    pre.resample(torch.from_numpy(raws).cuda())
    pre.select(2)
    slices = pre.gather_slices(2)
    torch.cuda.synchronize()
    for b in range(2):
        _, _, sl = O.preprocess_subject(raws[b], (32, 32, 32), [0, 2], [3, 2])
        assert np.array_equal(slices[b].cpu().numpy().view(np.uint32), sl.view(np.uint32))
    draws = [draw_passes(int(g["script/seeds"][b]), 2, 5, (32, 32), targs) for b in range(2)]
    for p in range(2):
        params = torch.from_numpy(params_bytes([draws[b][p] for b in range(2)])).cuda()
        noise = torch.from_numpy(np.stack([draws[b][p].noise for b in range(2)])).cuda()
        aug = pre.tta_augment(slices, params, noise)
        x = pre.resize_slices(aug)
        torch.cuda.synchronize()
        for b in range(2):
            want = g[f"script/aug/{b}/{p}"]
            assert np.array_equal(aug[b].cpu().numpy().view(np.uint32), want.view(np.uint32)), (b, p)
            ref_in = O.slices_to_input(want, 64)                       # [L,3,S,S]
            np.testing.assert_allclose(x[b].cpu().numpy().transpose(0, 3, 1, 2), ref_in, atol=1e-5, rtol=0)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 1e-2)])
def test_builders_with_tta_match_reference_script(golden, tmp_path, monkeypatch, precision, tol):
    """embed_manifest(tta=2) with the reference run's per-subject seeds reproduces the MIL script's embeddings."""
    g = golden("tta")

    def _bb(backbone, pretrained=True):
        torch.manual_seed(1234)
        m = ResNet2D("resnet18")
        dim = m.fc.in_features
        m.fc = torch.nn.Identity()
        return m, dim, None
    monkeypatch.setattr(of, "_build_resnet_backbone", _bb)
    monkeypatch.setenv("PD_FUSION_B200_PRECISION", precision)
    manifest = write_synthetic_manifest(tmp_path / "vols", 2, shape=(48, 40, 36))
    df = pd.read_csv(manifest)
    targs = json.loads(str(g["script/targs"]))
    emb, avg = of.embed_manifest(df, "resnet18", (32, 32, 32), [0, 2], [3, 2], 64, tta=2, tta_cfg=targs,
                                 tta_seeds=[int(s) for s in g["script/seeds"]])
    want = g["script/emb"]
    rel = np.linalg.norm(emb - want, axis=2) / np.linalg.norm(want, axis=2)
    assert emb.shape == want.shape and rel.max() < tol, rel
    rel_avg = np.linalg.norm(avg - want.mean(axis=1), axis=1) / np.linalg.norm(want.mean(axis=1), axis=1)
    assert rel_avg.max() < tol


def test_overlapped_pipeline_equals_serial():
    """Two-stream execution (preprocessing of batch i+1 next to the convolutions of batch i, alternating encoder instances)
    returns exactly what the single-stream path returns, batch after batch."""
    from pd_fusion_b200.pipeline import EmbeddingPipeline
    torch.manual_seed(1234)
    sd = {k: v for k, v in ResNet2D("resnet18").state_dict().items() if not k.startswith("fc.")}
    shape, target = (48, 40, 36), (32, 32, 32)
    pipe = EmbeddingPipeline(sd, shape, target, [0, 2], [3, 2], 64, precision="bf16", max_subjects=3)
    batches = [torch.from_numpy(np.stack([synthetic_volume(10 * k + b, shape) for b in range(3 if k != 2 else 2)])).cuda() for k in range(5)]
    want = []
    for raw in batches:
        r = pipe.embed(raw)
        torch.cuda.synchronize()
        want.append((r.embeddings.clone(), r.mean.clone(), r.indices.clone()))
    pipe.overlap_begin() if pipe._ov else pipe.enable_overlap()
    pipe.overlap_begin()
    got = []
    for raw in batches:
        r = pipe.embed_overlapped(raw)
        with torch.cuda.stream(pipe.conv_stream):
            got.append((r.embeddings.clone(), r.mean.clone(), r.indices.clone()))
    pipe.overlap_end()
    torch.cuda.synchronize()
    for (e0, m0, i0), (e1, m1, i1) in zip(want, got):
        assert torch.equal(e0, e1) and torch.equal(m0, m1) and torch.equal(i0, i1)
