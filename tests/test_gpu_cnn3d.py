"""f4, the `cnn3d` feature mode: the native 3-D conv auto-encoder (pd_fusion_b200/cnn3d.py + csrc/cnn3d.cu) against torch autograd
of the same module, and the drop-in CLI against the UNMODIFIED reference script's output (tests/golden/cnn3d.npz)."""
import json
import runpy
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from pd_fusion_b200.cnn3d import Cnn3dTrainer, Simple3DAE, standardize_volumes
from pd_fusion_b200.synthetic import write_synthetic_manifest

ROOT = Path(__file__).resolve().parents[1]


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("shape,B", [((16, 16, 16), 3), ((24, 16, 8), 2)])
def test_autoencoder_forward_backward_vs_autograd(shape, B):
    """Loss, embeddings and every parameter gradient of one batch against torch autograd of the same nn.Module in float64
    (depth-decomposed 3x3x3 convolutions, max pool winners, transposed convolutions as GEMM + shuffle, the re-ordered fc columns)."""
    torch.manual_seed(3)
    model = Simple3DAE(shape, 24).cuda()
    ref = Simple3DAE(shape, 24).cuda().double()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    x = torch.randn(B, *shape, device="cuda")
    tr = Cnn3dTrainer(model, 1e-3)
    loss = tr.forward_backward(x)
    emb = tr.embed(x)
    recon, remb = ref(x.double().unsqueeze(1))
    rl = torch.nn.functional.mse_loss(recon, x.double().unsqueeze(1))
    rl.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(rl)) < 1e-5 * float(rl), (float(loss), float(rl))
    assert _rel(emb, remb.detach()) < 1e-5
    for k, p in ref.named_parameters():
        assert _rel(tr.g[k], p.grad) < 2e-4, (k, _rel(tr.g[k], p.grad))


def test_training_steps_follow_torch_adam():
    """Three native optimisation steps against torch.optim.Adam on the float64 copy: loss trajectory and final weights."""
    torch.manual_seed(5)
    shape = (16, 16, 16)
    model = Simple3DAE(shape, 16).cuda()
    ref = Simple3DAE(shape, 16).cuda().double()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    tr = Cnn3dTrainer(model, 1e-3)
    for step in range(3):
        x = torch.randn(4, *shape, device="cuda")
        loss = tr.train_step(x)
        recon, _ = ref(x.double().unsqueeze(1))
        rl = torch.nn.functional.mse_loss(recon, x.double().unsqueeze(1))
        opt.zero_grad()
        rl.backward()
        opt.step()
        assert abs(float(loss) - float(rl)) < 1e-4 * float(rl), (step, float(loss), float(rl))
    for k, p in ref.named_parameters():
        assert _rel(dict(model.named_parameters())[k].data, p.data) < 1e-4, k


def test_standardize_matches_reference_rule():
    rng = np.random.default_rng(0)
    v = (rng.gamma(4.0, 100.0, size=(3, 20, 18, 16)) * (rng.random((3, 20, 18, 16)) > 0.5)).astype(np.float32)
    v[2] = -np.abs(v[2])                                    # no positive voxel: passes through
    got = standardize_volumes(torch.from_numpy(v).cuda()).cpu().numpy()
    for b in range(3):
        d = v[b].copy()
        m = d > 0
        if m.sum() > 0:
            d = (d - d[m].mean()) / (d[m].std() + 1e-6)    # scripts/build_cnn3d_embeddings.py:35-40
        np.testing.assert_allclose(got[b], d, rtol=2e-5, atol=2e-5)


def test_cli_matches_reference_script(golden, tmp_path, monkeypatch):
    """scripts/build_cnn3d_embeddings.py (this repo) with the arguments of the golden run: same file names and columns; embeddings of
    the trained auto-encoder within 2e-3 of the reference's CPU run (same seed -> same initial weights and batch order; four Adam
    steps of float32 arithmetic in a different summation order)."""
    g = golden("cnn3d")
    spec = json.loads(str(g["spec"]))
    manifest = write_synthetic_manifest(tmp_path / "vols", spec["n"], shape=tuple(spec["shape"]), start=spec["start"])
    od = tmp_path / "out"
    monkeypatch.setattr(sys, "argv", ["build_cnn3d_embeddings.py", "--manifest", str(manifest), "--out-dir", str(od)] + [str(a) for a in g["argv"]])
    monkeypatch.syspath_prepend(str(ROOT / "scripts"))
    runpy.run_path(str(ROOT / "scripts" / "build_cnn3d_embeddings.py"), run_name="__main__")
    import pandas as pd
    files = sorted(p.name for p in od.iterdir())
    assert len(files) == 2 and files[0].endswith(".json") and files[1].endswith(".parquet") and files[1].startswith(str(g["file_prefix"]) + "_")
    df = pd.read_parquet(od / files[1])
    assert list(df.columns) == [str(c) for c in g["columns"]]
    emb = df[[c for c in df.columns if c.startswith("mri_cnn_")]].values.astype(np.float32)
    want = g["emb"]
    assert emb.shape == want.shape
    err = np.abs(emb - want).max()
    assert err < 2e-3 * max(1.0, np.abs(want).max()), err
    meta = json.loads((od / files[0]).read_text())
    assert meta["config"]["target_shape"] == spec["target"] and meta["config"]["epochs"] == spec["epochs"]
