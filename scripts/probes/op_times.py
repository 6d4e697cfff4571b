"""GPU probe: device time of every encoder op in isolation (CUDA events, 30 launches each, best of 3).
PDFUSION_B200_LIB selects the library build.  usage: op_times.py [arch] [n_slices]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 768
lib = _lib.load()
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D(arch).state_dict().items() if not k.startswith("fc.")}
enc = ResNetEncoder(sd, n, 224, precision="bf16", arch=arch)
enc.input.copy_((torch.rand(n, 224, 224, device="cuda") * 2 - 1).to(torch.bfloat16))
enc.forward(None)
torch.cuda.synchronize()
total = 0.0
names = enc.op_names
for i in range(enc.n_ops):
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            enc.run_range(i, 1)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30 * 1e3)
    total += best
    print(f"op {i:2d} {names[i] if names else '':28s} {best:8.1f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    enc.forward(None)
e1.record()
torch.cuda.synchronize()
print(f"sum of ops {total:.1f} us; whole stack back to back {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
