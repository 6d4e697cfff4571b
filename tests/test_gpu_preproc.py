"""K1 parity on the GPU: CUDA preprocessing (through the C ABI) vs the CPU oracle and the golden fixtures.
Integer outputs (slice indices, counts) and the resampled/normalised volumes are bit-exact."""
import json

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from pd_fusion_b200 import _lib
from pd_fusion_b200.preprocess import VolumePreprocessor
from pd_fusion_b200.synthetic import synthetic_volume


def _run(raws, target, axes, counts, size, mode):
    pre = VolumePreprocessor(raws[0].shape, target, axes, counts, size, out_mode=mode, max_batch=len(raws))
    raw = torch.from_numpy(np.stack(raws)).cuda()
    res = pre.run(raw)
    norm = pre.normalized_volume(len(raws))
    torch.cuda.synchronize()
    return pre, res, norm


CASES = ["small_a", "small_b", "small_c", "full_c2", "full_c5"]


@pytest.mark.parametrize("case", CASES)
def test_preproc_vs_golden_and_oracle(golden, case):
    g = golden("preproc")
    sp = json.loads(str(g[f"{case}/spec"]))
    raw = synthetic_volume(sp["index"], tuple(sp["shape"]), bad_fraction=1e-4 if sp["shape"][0] < 100 else 1e-5)
    size = 56 if sp["target"][0] <= 64 else 224
    pre, res, norm = _run([raw], sp["target"], sp["axes"], sp["counts"], size, _lib.OUT_F32_NHWC3)
    zoomed = res.zoomed[0].cpu().numpy()
    ref_zoom = O.load_volume(raw, tuple(sp["target"]))
    assert np.array_equal(zoomed.view(np.uint32), ref_zoom.view(np.uint32)), "resample not bit-exact"
    lo, hi = O.percentile_bounds(ref_zoom)
    lohi = res.lohi[0].cpu().numpy()
    assert lohi[0] == lo == g[f"{case}/lo"] and lohi[1] == hi == g[f"{case}/hi"]
    ref_norm = O.normalize_volume_for_resnet(ref_zoom)
    assert np.array_equal(norm[0].cpu().numpy().view(np.uint32), ref_norm.view(np.uint32)), "normalised volume not bit-exact"
    idx = res.indices[0].cpu().numpy()
    ns = res.nslices[0].cpu().numpy()
    off = 0
    sl = []
    for a, (axis, c) in enumerate(zip(sp["axes"], sp["counts"])):
        want = g[f"{case}/idx{axis}"]
        assert ns[a] == len(want)
        assert np.array_equal(idx[off:off + len(want)], want), f"axis {axis} indices differ"
        assert np.all(idx[off + len(want):off + c] == -1)
        sl.append((off, len(want), O.select_slices(ref_norm, axis, c)))
        off += c
    x = res.net_input[0].cpu().numpy()            # [L, S, S, 3]
    for off, n, s in sl:
        ref = O.slices_to_input(s, size).transpose(0, 2, 3, 1)
        np.testing.assert_allclose(x[off:off + n], ref, atol=1e-5, rtol=0)
    # reference-generated sample of the network input (first two slices)
    np.testing.assert_allclose(x[:2].transpose(0, 3, 1, 2)[:, :, ::4, ::4], g[f"{case}/input_sample"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("case", ["zeros", "negative"])
def test_preproc_degenerate(golden, case):
    g = golden("preproc")
    raw = np.zeros((20, 20, 20), np.float32) if case == "zeros" else -synthetic_volume(5, (20, 22, 24), 0.0) - 1.0
    pre, res, norm = _run([raw], (16, 16, 16), [2], [4], 32, _lib.OUT_F32_NHWC3)
    lohi = res.lohi[0].cpu().numpy()
    assert lohi[0] == g[f"{case}/lo"] and lohi[1] == g[f"{case}/hi"] and lohi[3] == 0.0
    ref_norm = O.normalize_volume_for_resnet(O.load_volume(raw, (16, 16, 16)))
    assert np.array_equal(norm[0].cpu().numpy().view(np.uint32), ref_norm.view(np.uint32))
    want = g[f"{case}/idx2"]
    assert np.array_equal(res.indices[0].cpu().numpy()[:len(want)], want)


@pytest.mark.parametrize("kind", ["two_levels", "constant", "narrow"])
def test_percentiles_when_candidates_overflow_the_list(kind):
    """The refinement keeps a candidate list of the voxels that share a coarse bucket with a queried order statistic (a quarter
    of the volume at most).  Volumes whose positive voxels crowd into one or two buckets overflow it and must take the
    full-scan fallback with the same bit-exact percentiles."""
    rng = np.random.default_rng(7)
    shape = (40, 40, 40)
    if kind == "two_levels":
        raw = rng.choice(np.array([0.0, 100.0, 100.5], np.float32), size=shape, p=[0.2, 0.5, 0.3]).astype(np.float32)
    elif kind == "constant":
        raw = np.full(shape, 37.25, np.float32)
    else:
        raw = (1000.0 + rng.random(shape) * 1e-2).astype(np.float32)          # all positives inside one 13-bit bucket
    raws = [raw, synthetic_volume(3, shape, 1e-4)]                              # an ordinary subject next to it in the batch
    pre, res, norm = _run(raws, shape, [2], [6], 32, _lib.OUT_F32_NHWC3)       # identity zoom
    for b, r in enumerate(raws):
        zoom = O.load_volume(r, shape)
        lo, hi = O.percentile_bounds(zoom)
        got = res.lohi[b].cpu().numpy()
        assert got[0] == lo and got[1] == hi, (kind, b, got[:2], lo, hi)
        ref_norm = O.normalize_volume_for_resnet(zoom)
        assert np.array_equal(norm[b].cpu().numpy().view(np.uint32), ref_norm.view(np.uint32))


def test_resample_extreme_values_bit_exact():
    """The resample kernel widens float32 to float64 and narrows the sum back with integer instructions on the common path:
    subnormals, signed zeros, huge magnitudes, results that underflow to float32 subnormals and NaN/Inf must still match the
    float64 reference bit for bit."""
    rng = np.random.default_rng(11)
    shape, target = (24, 28, 32), (17, 19, 23)
    mag = np.float32(10.0) ** rng.uniform(-44, 38, size=shape).astype(np.float32)        # down into the subnormal range
    raw = (mag * rng.choice(np.array([-1.0, 1.0], np.float32), size=shape)).astype(np.float32)
    raw[rng.random(shape) < 0.15] = 0.0
    raw[rng.random(shape) < 0.05] = -0.0
    raw[rng.random(shape) < 0.01] = np.nan
    raw[rng.random(shape) < 0.01] = np.inf
    raw[0, 0, :4] = np.array([1e-45, 1.17549435e-38, 3.4028235e38, -3.4028235e38], np.float32)
    with np.errstate(all="ignore"):
        ref = O.load_volume(raw, target)
    pre = VolumePreprocessor(shape, target, [2], [4], 32, out_mode=_lib.OUT_F32_NHWC3, max_batch=1)
    got = pre.resample(torch.from_numpy(raw[None]).cuda())[0].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_preproc_batch_and_bf16_output():
    """Several subjects in one launch; bf16 one-channel output equals the rounded f32 output."""
    shape, target = (40, 36, 44), (32, 32, 32)
    raws = [synthetic_volume(20 + i, shape, 1e-4) for i in range(5)]
    _, res32, _ = _run(raws, target, [0, 2], [5, 3], 48, _lib.OUT_F32_NHWC3)
    _, res16, _ = _run(raws, target, [0, 2], [5, 3], 48, _lib.OUT_BF16_C1)
    for b, raw in enumerate(raws):
        vol, idx, sl = O.preprocess_subject(raw, target, [0, 2], [5, 3])
        got = res32.indices[b].cpu().numpy()
        assert np.array_equal(got[:len(idx[0])], idx[0]) and np.array_equal(got[5:5 + len(idx[1])], idx[1])
    a = res32.net_input[..., 0].to(torch.bfloat16)
    assert torch.equal(a, res16.net_input)
    # zero-padded layout read by the fused stem: same pixels at origin (5, 5), exact zeros everywhere else
    _, resp, _ = _run(raws, target, [0, 2], [5, 3], 48, _lib.OUT_BF16_C1_PAD)
    lo = _lib.STEM_PAD_LO
    pad = resp.net_input
    assert pad.shape[2:] == tuple(reversed(_lib.stem_padded_dims(48)))
    assert torch.equal(pad[:, :, lo:lo + 48, lo:lo + 48], res16.net_input)
    border = pad.clone()
    border[:, :, lo:lo + 48, lo:lo + 48] = 0
    assert not border.any()


def test_resample_property_full_size():
    """Size-independent properties at BASELINE's full size: constant volumes stay constant, output is
    bounded by the input range, NaN/Inf never survive."""
    raw = np.full((256, 256, 176), 3.25, np.float32)
    raw[5, 7, 9] = np.nan
    raw[100, 100, 100] = np.inf
    pre = VolumePreprocessor(raw.shape, (160, 160, 160), [2], [24], 224, max_batch=1)
    z = pre.resample(torch.from_numpy(raw[None]).cuda())
    torch.cuda.synchronize()
    z = z.cpu().numpy()
    assert np.isfinite(z).all() and z.max() <= 3.25 and z.min() >= 0.0
    assert (z == 3.25).mean() > 0.99


@pytest.mark.parametrize("shape,target,count,size,mode", [((64, 48, 44), (40, 40, 40), 6, 56, _lib.OUT_F32_NHWC3),
                                                           ((96, 80, 72), (64, 56, 48), 10, 96, _lib.OUT_BF16_C1_PAD),
                                                           ((256, 256, 176), (160, 160, 160), 48, 224, _lib.OUT_BF16_C1_PAD)])
def test_slice_major_run_equals_c_order_run(shape, target, count, size, mode):
    """The fused run() with the resampled volume kept slice-axis-major ([B, T2, T0, T1]: contiguous planes, no gather pass) must
    give the C-order run's results bit for bit: the volume (transposed), lo/hi, the indices and the network input."""
    raws = np.stack([synthetic_volume(20 + i, shape, 1e-4) for i in range(2)])
    raw = torch.from_numpy(raws).cuda()
    a = VolumePreprocessor(shape, target, [2], [count], size, out_mode=mode, max_batch=2)
    b = VolumePreprocessor(shape, target, [2], [count], size, out_mode=mode, max_batch=2, slice_major=True)
    assert b.run_slice_major and not a.run_slice_major
    ra, rb = a.run(raw), b.run(raw)
    torch.cuda.synchronize()
    assert tuple(rb.zoomed.shape) == (2, target[2], target[0], target[1])
    assert torch.equal(rb.zoomed.permute(0, 2, 3, 1), ra.zoomed), "slice-major volume differs"
    assert torch.equal(ra.lohi, rb.lohi) and torch.equal(ra.indices, rb.indices) and torch.equal(ra.nslices, rb.nslices)
    assert torch.equal(ra.net_input.view(torch.int16) if ra.net_input.dtype == torch.bfloat16 else ra.net_input,
                       rb.net_input.view(torch.int16) if rb.net_input.dtype == torch.bfloat16 else rb.net_input), "network input differs"
    # configurations the layout does not apply to fall back silently to C order
    c = VolumePreprocessor(shape, target, [0, 2], [3, 3], size, out_mode=mode, max_batch=2, slice_major=True)
    assert not c.run_slice_major
