"""GPU probe: conv_pw_kernel launches of the ResNet50 stack under several (ring stages, staging buffers) configurations,
A/B inside ONE process (box-to-box variance is larger than the effects).  usage: pw_cfg.py [n_slices] [cfg ...] (cfg = s,r)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(3, 3), (4, 2), (3, 2), (2, 3), (2, 4)]
lib = _lib.load()
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D("resnet50").state_dict().items() if not k.startswith("fc.")}
enc = ResNetEncoder(sd, n, 224, precision="bf16", arch="resnet50")
enc.input.copy_((torch.rand(n, 224, 224, device="cuda") * 2 - 1).to(torch.bfloat16))
enc.forward(None)
torch.cuda.synchronize()
ops = [i for i, nm in enumerate(enc.op_names) if ("conv3" in nm or "downsample" in nm or nm.endswith("0.conv1"))]


def t_op(i, reps=30):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            enc.run_range(i, 1)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


print("op".ljust(26) + "".join(f"{str(c):>10s}" for c in cfgs))
tot = [0.0] * len(cfgs)
for i in ops:
    row = []
    for j, (s, r) in enumerate(cfgs):
        _lib.check(lib.pdf_debug_set_pw_config(s, r))
        t = t_op(i)
        row.append(t)
        tot[j] += t
    print(enc.op_names[i].ljust(26) + "".join(f"{t:10.1f}" for t in row))
print("sum".ljust(26) + "".join(f"{t:10.1f}" for t in tot))
for j, (s, r) in enumerate(cfgs):
    _lib.check(lib.pdf_debug_set_pw_config(s, r))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc.forward(None)
    e0.record()
    for _ in range(20):
        enc.forward(None)
    e1.record()
    torch.cuda.synchronize()
    print(f"cfg {(s, r)}: whole stack back to back {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
