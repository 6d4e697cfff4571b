"""GPU debug: one 3x3 stride-1 conv through conv3x3_hs.cu vs torch, error pattern by output column / row / channel block."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib

lib = _lib.load()
n, h, w_, c, k = (int(v) for v in sys.argv[1:6]) if len(sys.argv) > 5 else (2, 6, 6, 64, 128)
g = torch.Generator().manual_seed(1)
x = (torch.randn(n, h, w_, c, generator=g) * 0.5).to(torch.bfloat16).cuda()
wt = (torch.randn(k, 3, 3, c, generator=g) / 24.0).to(torch.bfloat16).cuda()
bias = torch.randn(k, generator=g).cuda()
use_res = len(sys.argv) > 6 and int(sys.argv[6])
relu = len(sys.argv) > 7 and int(sys.argv[7])
resid = (torch.randn(n, h, w_, k, generator=g) * 0.5).to(torch.bfloat16).cuda() if use_res else None
ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, padding=1).permute(0, 2, 3, 1)
if use_res:
    ref = ref + resid.float()
if relu:
    ref = torch.relu(ref)
for mode in (0, 3):
    lib.pdf_debug_set_hs_mode(mode)
    out = torch.full((n, h, w_, k), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops = (_lib.Op * 1)()
    o = ops[0]
    o.kind, o.precision = _lib.OP_CONV, _lib.PREC_BF16
    o.n, o.h, o.w, o.c, o.k, o.r, o.s, o.stride, o.pad, o.ho, o.wo, o.relu = n, h, w_, c, k, 3, 3, 1, 1, h, w_, int(relu)
    o.d_residual = resid.data_ptr() if use_res else None
    o.d_in, o.d_weight, o.d_bias, o.d_out = x.data_ptr(), wt.data_ptr(), bias.data_ptr(), out.data_ptr()
    plan = C.c_void_p()
    _lib.check(lib.pdf_plan_create(C.byref(plan), ops, 1))
    _lib.check(lib.pdf_plan_run(plan, _lib.stream_ptr()))
    torch.cuda.synchronize()
    lib.pdf_plan_destroy(plan)
    err = (out.float() - ref).abs()
    print(f"mode {mode}: max err {err.max().item():.4f} nan {torch.isnan(out.float()).sum().item()}")
    if mode == 3:
        print(" err by output column q:", [round(v, 3) for v in err.amax(dim=(0, 1, 3)).tolist()])
        print(" err by output row p   :", [round(v, 3) for v in err.amax(dim=(0, 2, 3)).tolist()])
        print(" err by image          :", [round(v, 3) for v in err.amax(dim=(1, 2, 3)).tolist()])
        # which single tap explains the difference?  out_hs - ref vs contributions
        d = (out.float() - ref)
        xp = torch.nn.functional.pad(x.float().permute(0, 3, 1, 2), (2, 2, 2, 2))
        for r in range(3):
            for s in range(3):
                for ds in (-1, 1):
                    # hypothesis: tap (r,s) read the pixel one column off (ds)
                    good = torch.nn.functional.conv2d(xp[:, :, 1 + r:1 + r + h, 1 + s:1 + s + w_], wt.float()[:, r, s, :, None, None])
                    bad = torch.nn.functional.conv2d(xp[:, :, 1 + r:1 + r + h, 1 + s + ds:1 + s + ds + w_], wt.float()[:, r, s, :, None, None])
                    resid = (d - (bad - good).permute(0, 2, 3, 1)).abs().max().item()
                    if resid < 0.5 * d.abs().max().item():
                        print(f"  tap ({r},{s}) shifted by {ds} explains part: residual {resid:.3f}")
lib.pdf_debug_set_hs_mode(1)
