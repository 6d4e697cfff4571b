"""Device-side voxel decode: stored NIfTI voxels -> the float32 array the reference's loader produces,
bit-exact against the oracle's restatement of nibabel's get_fdata().astype(float32), for every stored type, both storage orders,
scaled and unscaled, ragged sizes; and through the drop-in builder's batching."""
import gzip
import struct

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from pd_fusion_b200 import _lib
from pd_fusion_b200.data import openneuro_features as F

CODES = {"uint8": 2, "int16": 4, "int32": 8, "float32": 16, "float64": 64, "int8": 256, "uint16": 512, "uint32": 768}


def _voxels(dtype, n, rng):
    if dtype.startswith("float"):
        v = (rng.standard_normal(n) * 1e3).astype(dtype)
        v[::97] = np.nan
        v[5::211] = np.inf
        return v
    info = np.iinfo(dtype)
    v = rng.integers(info.min, info.max, size=n, dtype=np.int64 if dtype != "uint32" else np.uint64).astype(dtype)
    v[:2] = (info.min, info.max)
    return v


@pytest.mark.parametrize("dtype", sorted(CODES))
@pytest.mark.parametrize("fortran", [True, False])
@pytest.mark.parametrize("slope,inter", [(1.0, 0.0), (0.0, 5.0), (0.0123456789, -3.25), (float("nan"), 1.0)])
def test_decode_matches_oracle(dtype, fortran, slope, inter):
    slope, inter = float(np.float32(slope)), float(np.float32(inter))          # the header stores float32
    rng = np.random.default_rng(len(dtype) + int(fortran))
    B, shape = 3, (37, 21, 45)                                   # nothing a multiple of the 32 x 32 transpose tile
    vox = np.stack([_voxels(dtype, int(np.prod(shape)), rng) for _ in range(B)])
    lib = _lib.load()
    src = torch.from_numpy(vox.view(np.uint8)).cuda()
    out = torch.full((B,) + shape, -7.0, dtype=torch.float32, device="cuda")
    _lib.check(lib.pdf_decode_volume(B, CODES[dtype], *shape, int(fortran), slope, inter, src.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for b in range(B):
        ref = O.decode_stored_voxels(vox[b], shape, fortran, slope, inter)
        nan = np.isnan(ref)                                       # NaN payloads are not part of the contract (nan_to_num zeroes them next)
        assert np.array_equal(np.isnan(got[b]), nan), (dtype, fortran, slope, inter, b)
        assert np.array_equal(got[b][~nan].view(np.uint32), ref[~nan].view(np.uint32)), (dtype, fortran, slope, inter, b)


def test_decode_rejects_unknown_type():
    lib = _lib.load()
    x = torch.zeros(64, dtype=torch.uint8, device="cuda")
    o = torch.zeros(8, dtype=torch.float32, device="cuda")
    assert lib.pdf_decode_volume(1, 1024, 2, 2, 2, 1, 1.0, 0.0, x.data_ptr(), o.data_ptr(), _lib.stream_ptr()) != 0
    assert b"datatype" in lib.pdf_last_error()


def _write_nifti(path, vol, code, slope, inter):
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, *vol.shape, 1, 1, 1, 1)
    struct.pack_into("<h", hdr, 70, code)
    struct.pack_into("<h", hdr, 72, vol.dtype.itemsize * 8)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, slope, inter)
    with gzip.open(path, "wb") as f:
        f.write(bytes(hdr) + vol.tobytes(order="F"))


@pytest.fixture()
def seeded_backbone(monkeypatch):
    from pd_fusion_b200.backbone import ResNet2D

    def _bb(backbone, pretrained=True):
        torch.manual_seed(1234)
        m = ResNet2D("resnet50" if backbone == "resnet50" else "resnet18")
        dim = m.fc.in_features
        m.fc = torch.nn.Identity()
        return m, dim, None
    monkeypatch.setattr(F, "_build_resnet_backbone", _bb)


def test_builder_batches_stored_volumes(tmp_path, seeded_backbone):
    """The manifest builder uploads the stored voxels and decodes on the device: same float32 volumes as the host reader, and the
    embeddings of int16 NIfTI files equal those of the same volumes given as float32 .npy."""
    import pandas as pd
    from pd_fusion_b200.synthetic import synthetic_volume
    shape = (40, 36, 44)
    vols = [np.round(synthetic_volume(30 + i, shape, 0.0)).astype(np.int16) for i in range(3)]
    rows_n, rows_f = [], []
    for i, v in enumerate(vols):
        pn, pf = tmp_path / f"s{i}.nii.gz", tmp_path / f"s{i}.npy"
        _write_nifti(pn, v, 4, 0.5, 2.0)
        np.save(pf, (v.astype(np.float64) * 0.5 + 2.0).astype(np.float32))
        rows_n.append({"subject_id": f"sub-{i}", "session": 1, "label": i % 2, "t1wbrain_path": str(pn)})
        rows_f.append({"subject_id": f"sub-{i}", "session": 1, "label": i % 2, "t1wbrain_path": str(pf)})
        sv = F._read_volume_stored(pn)
        dev = F._decode_on_device([sv], torch.device("cuda", 0))[0].cpu().numpy()
        assert np.array_equal(dev.view(np.uint32), F._read_volume_host(pn).view(np.uint32))
    kw = dict(backbone="resnet18", target_shape=(32, 32, 32), axes=[2], counts=[4], input_size=64)
    en, an = F.embed_manifest(pd.DataFrame(rows_n), **kw)
    ef, af = F.embed_manifest(pd.DataFrame(rows_f), **kw)
    assert np.array_equal(en, ef) and np.array_equal(an, af)
