"""GPU probe: ONE fine-tune training step of BASELINE config 5 (ResNet50, 4 bags x 64 slices) after two warm-up steps -- the
command profiled by ncu for the per-kernel time list of the training step.  PD_FUSION_B200_TRAIN_PRECISION selects bf16 | fp32."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import numpy as np
import torch
from pd_fusion_b200.models.mil_attention_finetune import MilAttentionFineTuneModel

bags_n, L = (int(sys.argv[1]) if len(sys.argv) > 1 else 4), (int(sys.argv[2]) if len(sys.argv) > 2 else 64)
params = {"backbone": "resnet50", "pretrained": False, "target_shape": [160, 160, 160], "slice_axis": 2, "slice_count": L, "input_size": 224,
          "slice_batch_size": 16, "batch_size": bags_n, "hidden_dim": 256, "attn_dim": 128, "dropout": 0.2, "gated": True, "loss_type": "focal",
          "focal_gamma": 2.0, "focal_alpha": 0.25, "lr": 3e-4, "lr_backbone": 1e-4, "weight_decay": 1e-3, "max_grad_norm": 1.0, "train_aug": False}
import os
if os.environ.get("PDFUSION_B200_WGRAD_WAVES"):
    from pd_fusion_b200 import _lib
    _lib.check(_lib.load().pdf_debug_set_wgrad_waves(int(os.environ["PDFUSION_B200_WGRAD_WAVES"])))
torch.manual_seed(1234)
model = MilAttentionFineTuneModel(params)
rng = np.random.default_rng(100)
bags = [torch.from_numpy(rng.random((L, 160, 160)).astype(np.float32)).cuda() for _ in range(bags_n)]
y = np.array([1.0, 0.0] * (bags_n // 2) + [1.0] * (bags_n % 2), dtype=np.float32)
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for _ in range(warm):
    model.train_step(bags, y, frozen=False, clip=1.0)
torch.cuda.synchronize()
reps = int(os.environ.get("FT_STEP_REPS", "1"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    loss, _ = model.train_step(bags, y, frozen=False, clip=1.0)
e1.record()
torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1) / reps:.2f} ms, loss {float(loss):.5f}")
