"""GPU probe: wgrad_tc_kernel on the ResNet50 fine-tune shapes (256 slices), one launch each after a warm-up -- timing with CUDA
events, or the target of an ncu capture (-k regex:wgrad_tc).  usage: wgrad_shapes.py"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib

lib = _lib.load()
SHAPES = [("layer1 3x3 64->64", 256, 56, 64, 64, 3, 1, 1), ("layer1 1x1 64->256", 256, 56, 64, 256, 1, 1, 0), ("layer1 1x1 256->64", 256, 56, 256, 64, 1, 1, 0),
          ("layer2 3x3 128->128", 256, 28, 128, 128, 3, 1, 1), ("layer3 3x3 256->256", 256, 14, 256, 256, 3, 1, 1), ("layer3 1x1 256->1024", 256, 14, 256, 1024, 1, 1, 0),
          ("layer4 3x3 512->512", 256, 7, 512, 512, 3, 1, 1), ("layer4 1x1 512->2048", 256, 7, 512, 2048, 1, 1, 0)]
for name, n, h, c, k, r, stride, pad in SHAPES:
    ho = (h + 2 * pad - r) // stride + 1
    x = (torch.randn(n, h, h, c, device="cuda") * 0.5).to(torch.bfloat16)
    dy = (torch.randn(n, ho, ho, k, device="cuda") * 0.5).to(torch.bfloat16)
    dw = torch.zeros(k, r, r, c, dtype=torch.float32, device="cuda")
    op = _lib.Op()
    op.kind, op.precision = _lib.OP_CONV, _lib.PREC_BF16
    op.n, op.h, op.w, op.c, op.k, op.r, op.s, op.stride, op.pad, op.ho, op.wo = n, h, h, c, k, r, r, stride, pad, ho, ho
    run = lambda: _lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    flops = 2.0 * n * ho * ho * k * c * r * r
    byts = (x.numel() + dy.numel()) * 2
    print(f"{name:24s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  operands once {byts / 1e6:7.1f} MB = {byts / us / 1e3:6.2f} TB/s")
