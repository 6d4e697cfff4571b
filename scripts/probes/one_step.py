"""One hot-path step (preprocessing + conv stack + fusion head) for profiler captures: two warm-up steps, then ONE step between
cudaProfilerStart/Stop.  usage: one_step.py [c3|c2] [subjects]   (ncu --profile-from-start off ...)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "robust-multimodal-pd_b200"))
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from bench import IN_SHAPE, INPUT_SIZE, TARGET, WORKLOADS, host_pool
from pd_fusion_b200.backbone import ResNet2D
from pd_fusion_b200.heads import MilHead, ModDropSweep
from pd_fusion_b200.pipeline import EmbeddingPipeline

wl_name = sys.argv[1] if len(sys.argv) > 1 else "c3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
wl = WORKLOADS[wl_name]
L = sum(wl["counts"])
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D(wl["arch"]).state_dict().items() if not k.startswith("fc.")}
pipe = EmbeddingPipeline(sd, IN_SHAPE, TARGET, wl["axes"], wl["counts"], INPUT_SIZE, precision="bf16", max_subjects=B)
pool = host_pool(4)
raw = torch.stack([torch.from_numpy(pool[i % 4]) for i in range(B)]).cuda()
torch.manual_seed(4321)
S = 7
live = (torch.rand(S, B, device="cuda") < 0.6).to(torch.uint8)
if wl_name == "c3":
    from pd_fusion_b200.models.mil_attention import MILAttentionNet
    head = MilHead(MILAttentionNet(pipe.D, 256, 128, 0.2, gated=True).state_dict(), True, 0.5, precision="tf32")
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    fuse = lambda r: head.sweep(r.embeddings, lens, live)
else:
    from pd_fusion_b200.models.fusion_moddrop import ModalityDropoutNet
    dims = {"clinical": 0, "datspect": 0, "mri": pipe.D}
    head = ModDropSweep(ModalityDropoutNet(dims, [256, 128, 64], 0.3).state_dict(), dims)
    masks = torch.stack([torch.zeros_like(live), torch.zeros_like(live), live], dim=2).contiguous()
    fuse = lambda r: head.forward(r.mean, masks)
for _ in range(2):
    fuse(pipe.embed(raw))
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = fuse(pipe.embed(raw))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one step done", wl_name, B, float(out.float().mean()))
