"""Tensor-core weight gradient (csrc/wgrad_tc.cu: tcgen05 with MN-major operands, im2col-mode TMA) against the FP32 CUDA-core
wgrad kernel on the same bf16-rounded operands, and against torch autograd of F.conv2d."""
import ctypes as C

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from pd_fusion_b200 import _lib

CASES = [  # n, h, c, k, r, stride, pad
    (2, 8, 64, 64, 1, 1, 0),        # one k-block pair, 1x1
    (3, 14, 64, 128, 3, 1, 1),      # 3 taps per CTA, N = 64
    (2, 16, 128, 128, 3, 2, 1),     # strided 3x3, N = 128 (T = 3)
    (2, 9, 256, 64, 1, 1, 0),       # Cout = 64: the second dY sub-tile is out of bounds (zero fill)
    (3, 7, 512, 256, 3, 1, 1),      # N = 256, T = 1, two Cin blocks, two Cout blocks
    (2, 15, 128, 256, 1, 2, 0),     # strided 1x1 (downsample)
    (40, 28, 64, 64, 3, 1, 1),      # many pixel slabs (split reduction through atomics); row-tiled X loads, 4 rows per k-block
    (3, 56, 64, 64, 3, 1, 1),       # row-tiled, 2 rows of 56 pixels per k-block
    (5, 12, 128, 64, 3, 1, 1),      # row-tiled, 4 rows of 12 pixels (48-pixel k-blocks), Cin = 128 (three taps share the dY tile)
    (2, 20, 64, 128, 5, 1, 2),      # 5x5 filter, row-tiled (4 rows of 20), taps of a filter row in separate CTAs
]


@pytest.mark.parametrize("n,h,c,k,r,stride,pad", CASES)
def test_wgrad_tc_vs_f32_and_autograd(n, h, c, k, r, stride, pad):
    lib = _lib.load()
    ho = (h + 2 * pad - r) // stride + 1
    g = torch.Generator().manual_seed(n * 100 + c + k + r)
    x = (torch.randn(n, h, h, c, generator=g) * 0.5).to(torch.bfloat16)
    dy = (torch.randn(n, ho, ho, k, generator=g) * 0.5).to(torch.bfloat16)
    xd, dyd = x.cuda(), dy.cuda()
    op = _lib.Op()
    op.kind, op.precision = _lib.OP_CONV, _lib.PREC_BF16
    op.n, op.h, op.w, op.c, op.k, op.r, op.s, op.stride, op.pad, op.ho, op.wo = n, h, h, c, k, r, r, stride, pad, ho, ho
    dw = torch.zeros(k, r, r, c, dtype=torch.float32, device="cuda")
    for _ in range(2):                                       # accumulates: two calls = twice the gradient
        _lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), _lib.stream_ptr()), "pdf_conv_wgrad_bf16")
    torch.cuda.synchronize()
    dw = dw / 2
    # FP32 kernel on the same (bf16-rounded) values: [R,S,C,K]
    dw32 = torch.zeros(r, r, c, k, dtype=torch.float32, device="cuda")
    x32, dy32 = xd.float().contiguous(), dyd.float().contiguous()        # (named: a temporary's memory would be reused before the launch)
    _lib.check(lib.pdf_conv_wgrad_f32(C.byref(op), x32.data_ptr(), dy32.data_ptr(), dw32.data_ptr(), _lib.stream_ptr()), "pdf_conv_wgrad_f32")
    torch.cuda.synchronize()
    want = dw32.permute(3, 0, 1, 2)
    rel = float((dw - want).norm() / want.norm())
    assert rel < 1e-5, rel                                   # same products, f32 accumulation on both sides
    # torch autograd
    w = torch.zeros(k, c, r, r, device="cuda", requires_grad=True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out = torch.nn.functional.conv2d(xd.float().permute(0, 3, 1, 2), w, stride=stride, padding=pad)
        out.backward(dyd.float().permute(0, 3, 1, 2))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    rel2 = float((dw - w.grad.permute(0, 2, 3, 1)).norm() / w.grad.norm())
    assert rel2 < 1e-4, rel2
