"""GPU debug: run resnet18 op by op with the horizontally-shared 3x3 kernel off / on and report the first op whose output differs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D("resnet18").state_dict().items() if not k.startswith("fc.")}
x = (torch.rand(n, 224, 224, generator=torch.Generator().manual_seed(9)) * 2 - 1).to(torch.bfloat16).cuda()
encs = []
for mode in (0, 1):
    lib.pdf_debug_set_hs_mode(mode)
    e = ResNetEncoder(sd, n, 224, precision="bf16")
    e.input.copy_(x)
    encs.append(e)
lib.pdf_debug_set_hs_mode(1)
a, b = encs
for i in range(a.n_ops):
    a.run_range(i, 1); b.run_range(i, 1)
    torch.cuda.synchronize()
    worst = 0.0
    for name in a.buffers:
        ta, tb = a.buffers[name], b.buffers[name]
        va = ta.view(torch.bfloat16).float() if name != "o4" else ta.view(torch.float32)
        vb = tb.view(torch.bfloat16).float() if name != "o4" else tb.view(torch.float32)
        d = (va - vb).abs()
        d = torch.nan_to_num(d, nan=0.0, posinf=0.0)
        worst = max(worst, d.max().item())
    print(f"op {i:2d} {a.op_names[i]:26s} max |diff| over all buffers {worst:.4f}")
print("final", (a.output - b.output).abs().max().item())
