"""Size-independent properties of the hot path at BASELINE's FULL sizes (C3: 256x256x176 volumes -> 160^3 -> 48 slices -> ResNet50 ->
gated MIL head under the scenario masks), where the CPU oracle would take minutes per subject: results must not depend on a subject's
position in the device batch or on its batch mates, repeated runs must agree bit for bit, and the masked sweep must reduce to the
unmasked forward wherever the modality is present."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from pd_fusion_b200.backbone import ResNet2D
from pd_fusion_b200.heads import MilHead
from pd_fusion_b200.models.mil_attention import MILAttentionNet
from pd_fusion_b200.pipeline import EmbeddingPipeline
from pd_fusion_b200.synthetic import synthetic_volume

SHAPE, TARGET, L, SIZE = (256, 256, 176), (160, 160, 160), 48, 224


@pytest.fixture(scope="module")
def c3():
    torch.manual_seed(1234)
    sd = {k: v for k, v in ResNet2D("resnet50").state_dict().items() if not k.startswith("fc.")}
    pipe = EmbeddingPipeline(sd, SHAPE, TARGET, [2], [L], SIZE, precision="bf16", max_subjects=4)
    vols = torch.from_numpy(np.stack([synthetic_volume(300 + i, SHAPE, 1e-5) for i in range(4)])).cuda()
    return pipe, vols


def _embed(pipe, vols):
    r = pipe.embed(vols)
    torch.cuda.synchronize()
    return r.embeddings.clone(), r.mean.clone(), r.indices.clone()


def test_full_size_batch_position_independence_and_determinism(c3):
    pipe, vols = c3
    assert pipe.pre.run_slice_major                          # the C3 configuration runs the slice-major preprocessing
    e0, m0, i0 = _embed(pipe, vols)
    e1, m1, i1 = _embed(pipe, vols)
    assert torch.equal(e0, e1) and torch.equal(m0, m1) and torch.equal(i0, i1), "two runs on the same batch differ"
    perm = torch.tensor([2, 0, 3, 1], device="cuda")
    ep, mp, ip = _embed(pipe, vols[perm].contiguous())
    assert torch.equal(ip, i0[perm]), "slice indices depend on the batch position"
    assert torch.equal(ep, e0[perm]) and torch.equal(mp, m0[perm]), "embeddings depend on the batch position"
    es, ms, _ = _embed(pipe, vols[1:2].contiguous())         # a partial batch (unused tail zeroed)
    assert torch.equal(es[0], e0[1]) and torch.equal(ms[0], m0[1]), "embeddings depend on the batch mates"
    assert float(e0.abs().max()) > 0 and torch.isfinite(e0).all()
    # slice mean = mean over the 48 slices of the per-slice embeddings
    np.testing.assert_allclose(m0.cpu().numpy(), e0.mean(dim=1).cpu().numpy(), rtol=2e-6, atol=1e-6)


def test_full_size_mil_sweep_reduces_to_the_unmasked_forward(c3):
    pipe, vols = c3
    e0, _, _ = _embed(pipe, vols)
    torch.manual_seed(4321)
    head = MilHead(MILAttentionNet(pipe.D, 256, 128, 0.2, gated=True).state_dict(), True, 0.5, precision="fp32")
    lens = torch.full((4,), L, dtype=torch.int32, device="cuda")
    live = torch.tensor([[1, 1, 1, 1], [0, 1, 0, 1], [0, 0, 0, 0], [1, 0, 0, 0]], dtype=torch.uint8, device="cuda")
    full = head.forward(e0.contiguous(), lens).cpu().numpy()
    sweep = head.sweep(e0.contiguous(), lens, live).cpu().numpy()
    lv = live.cpu().numpy().astype(bool)
    assert np.array_equal(sweep[lv], np.broadcast_to(full, sweep.shape)[lv])
    assert np.all(sweep[~lv] == np.float32(0.5))
    # a bag is its own world: permuting the bags permutes the probabilities
    perm = [3, 1, 0, 2]
    fp = head.forward(e0[perm].contiguous(), lens).cpu().numpy()
    assert np.array_equal(fp, full[perm])
