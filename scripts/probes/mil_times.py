"""GPU probe: device time of one pdf_mil_sweep over a table of bags (C3 head: D 2048, H 256, A 128 gated, L 48, 7 scenarios) with
128-row and 256-row tiles in the tf32 GEMMs (pdf_debug_set_mil_mt), A/B inside one process.  usage: mil_times.py [n_bags]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.heads import MilHead
from pd_fusion_b200.models.mil_attention import MILAttentionNet

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
L, D = 48, 2048
lib = _lib.load()
torch.manual_seed(4321)
sd = MILAttentionNet(D, 256, 128, 0.2, gated=True).state_dict()
bags = torch.randn((nb, L, D), device="cuda")
lens = torch.full((nb,), L, dtype=torch.int32, device="cuda")
live = (torch.rand((7, nb), device="cuda") < 0.6).to(torch.uint8)
out = torch.empty((7, nb), dtype=torch.float32, device="cuda")
for prec in ("tf32", "fp32"):
    head = MilHead(sd, True, 0.5, precision=prec)
    for mt in ((1, 2, 0) if prec == "tf32" else (0,)):
        _lib.check(lib.pdf_debug_set_mil_mt(mt))
        for _ in range(3):
            head.sweep(bags, lens, live, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                head.sweep(bags, lens, live, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        gbs = (4 * L * D + 4) * nb / (best / 1e3) / 1e9
        print(f"{prec} row tiles per weight pass {mt}: {best * 1e3:.1f} us per sweep of {nb} bags, {gbs:.0f} GB/s of bag bytes")
_lib.check(lib.pdf_debug_set_mil_mt(0))
