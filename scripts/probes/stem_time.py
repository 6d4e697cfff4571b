"""GPU probe: device time of single encoder ops in isolation (CUDA events over many launches).
usage: stem_time.py [first_op count]   (default 0 1 = the fused stem); PDFUSION_B200_LIB selects the library build."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

a, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 1)
arch = sys.argv[3] if len(sys.argv) > 3 else "resnet18"
lib = _lib.load()
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D(arch).state_dict().items() if not k.startswith("fc.")}
n = 768
enc = ResNetEncoder(sd, n, 224, precision="bf16", arch=arch)
enc.input.copy_((torch.rand(n, 224, 224, device="cuda") * 2 - 1).to(torch.bfloat16))
enc.forward(None)
torch.cuda.synchronize()
best = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        enc.run_range(a, b)
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 50 * 1e3)
print(f"{arch} {enc.op_names[a]} ops [{a},+{b}): {min(best):.1f} us best, {sorted(best)[2]:.1f} us median of 5 x 50 launches")
