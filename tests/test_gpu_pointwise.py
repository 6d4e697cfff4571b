"""conv_pw_kernel (1x1 convolutions with the shared-memory epilogue, optionally chained with the next 1x1 convolution)
against an f32 torch computation on the same bf16 operands -- torchvision Bottleneck.forward semantics
(conv3 + bn3 + add + relu, then the next block's conv1 + bn1 + relu)."""
import ctypes as C

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from pd_fusion_b200 import _lib

CASES = [  # n, h, cin, cout, stride, residual, relu, k3
    (2, 7, 64, 256, 1, True, True, 0),        # M = 98: a single partial tile
    (3, 14, 64, 256, 1, True, True, 64),      # layer-1 shape, chained 256 -> 64
    (5, 28, 128, 512, 1, True, True, 128),    # layer-2 shape, several tiles per CTA
    (2, 14, 256, 1024, 1, True, True, 256),   # layer-3 shape: the chained weight tile takes a whole ring stage
    (3, 15, 256, 512, 2, False, False, 0),    # strided downsample (im2col-mode TMA)
    (2, 9, 64, 256, 1, False, False, 0),      # stride-1 downsample
    (4, 8, 512, 2048, 1, True, True, 0),      # layer-4 expansion, 16 chunks per tile
    (700, 7, 64, 256, 1, True, True, 64),     # > 148 M tiles: persistence, ring phases, buffer reuse
    (2, 6, 192, 64, 1, False, True, 0),       # 64-column variant
]


@pytest.mark.parametrize("n,h,cin,cout,stride,res,relu,k3", CASES)
def test_pointwise_vs_torch(n, h, cin, cout, stride, res, relu, k3):
    lib = _lib.load()
    lib.pdf_debug_set_pw(2)
    try:
        g = torch.Generator().manual_seed(n * 1000 + cin + cout + k3)
        ho = (h - 1) // stride + 1
        x = (torch.randn(n, h, h, cin, generator=g) * 0.5).to(torch.bfloat16)
        w = (torch.randn(cout, cin, generator=g) / cin ** 0.5).to(torch.bfloat16)
        bias = torch.randn(cout, generator=g)
        resid = (torch.randn(n, ho, ho, cout, generator=g) * 0.5).to(torch.bfloat16) if res else None
        w3 = (torch.randn(k3, cout, generator=g) / cout ** 0.5).to(torch.bfloat16) if k3 else None
        b3 = torch.randn(k3, generator=g) if k3 else None
        xd, wd, bd = x.cuda(), w.cuda(), bias.cuda()
        rd = resid.cuda() if res else None
        out = torch.full((n, ho, ho, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
        op = (_lib.Op * 1)()
        o = op[0]
        o.kind, o.precision = _lib.OP_CONV, _lib.PREC_BF16
        o.n, o.h, o.w, o.c, o.k, o.r, o.s, o.stride, o.pad, o.ho, o.wo, o.relu = n, h, h, cin, cout, 1, 1, stride, 0, ho, ho, int(relu)
        o.d_in, o.d_weight, o.d_bias, o.d_out = xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr()
        o.d_residual = rd.data_ptr() if res else None
        if k3:
            w3d, b3d = w3.cuda(), b3.cuda()
            out3 = torch.full((n, ho, ho, k3), float("nan"), dtype=torch.bfloat16, device="cuda")
            o.d_weight3, o.d_bias3, o.d_out3, o.k3 = w3d.data_ptr(), b3d.data_ptr(), out3.data_ptr(), k3
        plan = C.c_void_p()
        _lib.check(lib.pdf_plan_create(C.byref(plan), op, 1))
        for _ in range(2):                                  # twice: barriers / TMEM must come back clean
            _lib.check(lib.pdf_plan_run(plan, _lib.stream_ptr()))
        torch.cuda.synchronize()
        lib.pdf_plan_destroy(plan)
        xs = x[:, ::stride, ::stride].float()
        y = xs.reshape(-1, cin) @ w.float().T + bias
        if res:
            y = y + resid.float().reshape(-1, cout)
        if relu:
            y = y.clamp_min(0)
        got = out.float().cpu().reshape(-1, cout)
        assert torch.isfinite(got).all()
        err = (got - y).abs().max().item()
        assert err < 1.2e-2 * max(1.0, y.abs().max().item()), err          # one bf16 rounding of the output
        if k3:
            t = (got @ w3.float().T + b3).clamp_min(0)                      # chained conv reads the ROUNDED output, as the unfused path does
            got3 = out3.float().cpu().reshape(-1, k3)
            assert torch.isfinite(got3).all()
            err3 = (got3 - t).abs().max().item()
            assert err3 < 1.2e-2 * max(1.0, t.abs().max().item()), err3
    finally:
        lib.pdf_debug_set_pw(1)


def test_chained_encoder_equals_unchained():
    """ResNet50 with conv3 -> next conv1 chaining must reproduce the launch-per-conv network (same bf16 intermediates)."""
    import os
    from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder
    torch.manual_seed(1234)
    sd = {k: v for k, v in ResNet2D("resnet50").state_dict().items() if not k.startswith("fc.")}
    x = (torch.rand(5, 96, 96) * 2 - 1).to(torch.bfloat16).cuda()
    outs = []
    for chain in ("1,2,3", ""):
        os.environ["PDFUSION_B200_CHAIN"] = chain
        try:
            enc = ResNetEncoder(sd, 5, 96, precision="bf16")
        finally:
            del os.environ["PDFUSION_B200_CHAIN"]
        outs.append(enc.forward(x).clone())
        torch.cuda.synchronize()
        n_ops = enc.n_ops
        outs.append(n_ops)
    (a, na), (b, nb) = (outs[0], outs[1]), (outs[2], outs[3])
    assert na == nb - 12                                   # (3-1) + (4-1) + (6-1) conv1 launches folded into the preceding conv3,
                                                           # + layer2.0.conv1 and layer3.0.conv1 across the stage boundaries
    rel = (a - b).norm() / b.norm()
    assert rel < 2e-3, rel


@pytest.mark.parametrize("n,S", [(5, 96), (3, 64), (16, 128)])
def test_weight_multicast_pairs_equal_single_ctas(n, S):
    """conv_pw_kernel as weight-multicast CTA pairs (two CTAs of a cluster share every weight tile of the ring, each fetching half)
    must reproduce the single-CTA launches bit for bit -- odd tile counts (a pair's second tile past the end), chained and
    un-chained launches, strided downsamples."""
    from pd_fusion_b200 import _lib
    from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder
    lib = _lib.load()
    torch.manual_seed(1234)
    sd = {k: v for k, v in ResNet2D("resnet50").state_dict().items() if not k.startswith("fc.")}
    x = (torch.rand(n, S, S) * 2 - 1).to(torch.bfloat16).cuda()
    outs = []
    try:
        for mc in (0, 1):
            _lib.check(lib.pdf_debug_set_pw_multicast(mc))
            enc = ResNetEncoder(sd, n, S, precision="bf16")
            outs.append(enc.forward(x).clone())
            torch.cuda.synchronize()
    finally:
        lib.pdf_debug_set_pw_multicast(0)
    assert torch.isfinite(outs[1]).all() and float(outs[1].abs().max()) > 0
    assert torch.equal(outs[0], outs[1]), float((outs[0] - outs[1]).abs().max())
