"""GPU debug probe: row-tiled wgrad on a 14x14 map (8 rows per k-block: the second block of an image is partly outside it) --
which input ROW's contribution goes wrong?"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
lib = _lib.load()
n, h, c, k, r = 3, 14, 64, 128, 3
op = _lib.Op()
op.kind, op.precision = _lib.OP_CONV, _lib.PREC_BF16
op.n, op.h, op.w, op.c, op.k, op.r, op.s, op.stride, op.pad, op.ho, op.wo = n, h, h, c, k, r, r, 1, 1, h, h
torch.manual_seed(0)
dy = (torch.randn(n, h, h, k, device="cuda") * 0.5).to(torch.bfloat16)
xf = (torch.randn(n, h, h, c, device="cuda") * 0.5).to(torch.bfloat16)
for img in range(n):
    for row in range(h):
        x = torch.zeros_like(xf)
        x[img, row] = xf[img, row]
        outs = []
        for mode in (1, 0):
            lib.pdf_debug_set_wgrad_rowtile(mode)
            dw = torch.zeros(k, r, r, c, dtype=torch.float32, device="cuda")
            _lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
            torch.cuda.synchronize()
            outs.append(dw)
        err = [(float((outs[0][:, t] - outs[1][:, t]).norm() / outs[1][:, t].norm().clamp_min(1e-20))) for t in range(3)]
        if max(err) > 1e-4:
            print(f"image {img} input row {row}: rel err per filter row {['%.3f' % e for e in err]}")
print("done")
