"""GPU debug probe for wgrad_tc_kernel: one-hot operands show where each (pixel, cout, cin) product lands."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib

lib = _lib.load()
n, h, c, k = 2, 8, 64, 64
op = _lib.Op()
op.kind, op.precision = _lib.OP_CONV, _lib.PREC_BF16
op.n, op.h, op.w, op.c, op.k, op.r, op.s, op.stride, op.pad, op.ho, op.wo = n, h, h, c, k, 1, 1, 1, 0, h, h
M = n * h * h
for (m0, k0, c0) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 0), (0, 8, 0), (0, 0, 8), (9, 3, 5), (17, 40, 33), (70, 2, 3), (127, 63, 63)]:
    x = torch.zeros(M, c, dtype=torch.bfloat16, device="cuda")
    dy = torch.zeros(M, k, dtype=torch.bfloat16, device="cuda")
    x[m0, c0] = 1
    dy[m0, k0] = 1
    dw = torch.zeros(k, c, dtype=torch.float32, device="cuda")
    _lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    nz = dw.nonzero().tolist()
    print(f"pixel {m0} cout {k0} cin {c0}: nonzero at {nz[:6]} values {[float(dw[a, b]) for a, b in nz[:6]]}")
# all-pixels check: x = 1 everywhere in channel c0, dy = 1 everywhere in cout k0 -> dW[k0, c0] = M
x = torch.zeros(M, c, dtype=torch.bfloat16, device="cuda"); dy = torch.zeros(M, k, dtype=torch.bfloat16, device="cuda")
x[:, 5] = 1; dy[:, 3] = 1
dw = torch.zeros(k, c, dtype=torch.float32, device="cuda")
_lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize()
print("ones: nonzero", dw.nonzero().tolist()[:8], "values", [float(dw[a, b]) for a, b in dw.nonzero().tolist()[:8]], "expected", M, "at [3, 5]")
torch.manual_seed(0)
x = (torch.randn(M, c, device="cuda") * 0.5).to(torch.bfloat16)
dy = (torch.randn(M, k, device="cuda") * 0.5).to(torch.bfloat16)
for rep in range(2):
    dw = torch.zeros(k, c, dtype=torch.float32, device="cuda")
    _lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = dy.float().T @ x.float()
    err = (dw - want).abs()
    print("dense rep", rep, "rel", float((dw - want).norm() / want.norm()), "max err", float(err.max()), "at", divmod(int(err.argmax()), c),
          "dw[0,:4]", dw[0, :4].tolist(), "want[0,:4]", want[0, :4].tolist())
# only the first 64 pixels non-zero / only the second 64
for lo, hi in ((0, 64), (64, 128), (0, 16), (16, 32)):
    x2, dy2 = torch.zeros_like(x), torch.zeros_like(dy)
    x2[lo:hi], dy2[lo:hi] = x[lo:hi], dy[lo:hi]
    dw = torch.zeros(k, c, dtype=torch.float32, device="cuda")
    _lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x2.data_ptr(), dy2.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = dy2.float().T @ x2.float()
    print(f"pixels [{lo},{hi}): rel", float((dw - want).norm() / want.norm()))
# two pixels, cross terms
x2, dy2 = torch.zeros_like(x), torch.zeros_like(dy)
x2[3, 5] = 1; dy2[4, 7] = 1
dw = torch.zeros(k, c, dtype=torch.float32, device="cuda")
_lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x2.data_ptr(), dy2.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize()
print("cross (x at pixel 3, dy at pixel 4): nonzero", dw.nonzero().tolist())
for dpx in (1, 2, 8, 16):
    x2, dy2 = torch.zeros_like(x), torch.zeros_like(dy)
    x2[3, 5] = 1; dy2[3 + dpx, 7] = 1
    dw = torch.zeros(k, c, dtype=torch.float32, device="cuda")
    _lib.check(lib.pdf_conv_wgrad_bf16(C.byref(op), x2.data_ptr(), dy2.data_ptr(), dw.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    print(f"cross +{dpx}: nonzero", dw.nonzero().tolist())
