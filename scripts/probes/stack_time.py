"""GPU probe: device time of the whole encoder stack back to back (CUDA events, 20 passes, best of 3) under the current environment
(PDFUSION_B200_CHUNKS, PDFUSION_B200_CHAIN, PDFUSION_B200_XCHAIN ...).  usage: stack_time.py [arch] [n_slices]"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 768
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D(arch).state_dict().items() if not k.startswith("fc.")}
enc = ResNetEncoder(sd, n, 224, precision="bf16", arch=arch)
enc.input.copy_((torch.rand(n, 224, 224, device="cuda") * 2 - 1).to(torch.bfloat16))
enc.forward(None)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        enc.forward(None)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
env = {k: v for k, v in os.environ.items() if k.startswith("PDFUSION_B200_")}
print(f"{arch} n={n} ops={enc.n_ops} {env}: {best:.1f} us per pass, {enc.algorithmic_flops() / best / 1e6:.1f} TFLOP/s")
