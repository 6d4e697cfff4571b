"""The C-ABI library loads on a machine without a GPU and exports every symbol include/pdfusion_b200.h declares;
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from pd_fusion_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "pdfusion_b200.h").read_text()


def declared_symbols():
    names = set(re.findall(r"\b(pdf_[a-z0-9_]+)\s*\(", HEADER))
    return sorted(n for n in names if not n.endswith("_t"))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.PROTOTYPES) == set(syms), "ctypes prototypes and header drifted apart"
    assert lib.pdf_version() >= 100


def test_struct_sizes_match_header_layout():
    # natural alignment, same field order as the header
    assert C.sizeof(_lib.PreprocCfg) == 4 * (3 + 3 + 1 + 3 + 3 + 1 + 3 + 3 + 1 + 1)        # (+ slice_major)
    assert C.sizeof(_lib.Op) == 15 * 4 + 4 + 12 * 8 + 8     # 15 ints, padding, 6 + 3 (fused 1x1) + 3 (chained 1x1) pointers, k3 + padding
    assert C.sizeof(_lib.Mlp) == 8 + 9 * 4 + 4 + 16 * 8 + 4 + 9 * 4 + 4 or C.sizeof(_lib.Mlp) % 8 == 0


def test_argument_validation_without_gpu():
    lib = _lib.load()
    cfg = _lib.PreprocCfg()
    assert lib.pdf_preproc_workspace_bytes(C.byref(cfg), 0) == 0
    rc = lib.pdf_resample_stats(C.byref(cfg), 1, None, None, None, None)      # all-zero cfg -> argument error, no crash
    assert rc == -1 and b"preproc" in lib.pdf_last_error()
    rc = lib.pdf_slice_mean(0, 1, 1, None, None, None, None)
    assert rc == -1


def test_stem_padded_dims_host_function():
    # pure host arithmetic: 224 -> 56 x 56 pooled pixels in 7 x 8 tiles of 8 x 7; last patch ends at row 230 / column 231
    assert _lib.stem_padded_dims(224) == (232, 231)
    for S in (64, 70, 96, 128):
        pitch, rows = _lib.stem_padded_dims(S)
        assert pitch % 8 == 0 and pitch >= S + 5 and rows >= S + 5


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PdfusionError):
        _lib.require_cuda()
    from pd_fusion_b200.preprocess import VolumePreprocessor
    with pytest.raises(_lib.PdfusionError):
        VolumePreprocessor((8, 8, 8), (4, 4, 4))
    from pd_fusion_b200.utils.torch_utils import get_torch_device
    with pytest.raises(RuntimeError):
        get_torch_device()
    lib = _lib.load()
    plan = C.c_void_p()
    ops = (_lib.Op * 1)()
    assert lib.pdf_plan_create(C.byref(plan), ops, 1) != 0        # no device -> error, never a host computation
