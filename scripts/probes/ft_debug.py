"""GPU debug probe: native fine-tune step vs torch float64 autograd, teacher-forced (the reference gets the native weights before
every step): per-step loss, gradient norm and the parameters with the largest gradient error."""
import copy
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import numpy as np
import torch
import torch.nn.functional as F
from pd_fusion_b200.models.mil_attention_finetune import MilAttentionFineTuneModel

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
lens = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [6, 5, 6]
CLIP = 50.0
params = {"backbone": "resnet18", "pretrained": False, "input_size": 64, "hidden_dim": 32, "attn_dim": 16, "dropout": 0.0, "gated": True,
          "batch_size": 3, "slice_batch_size": 4, "lr_backbone": 1e-3, "lr": 3e-3, "weight_decay": 1e-3, "loss_type": "focal",
          "focal_gamma": 2.0, "focal_alpha": 0.25, "train_aug": False, "max_grad_norm": CLIP, "slice_count": 6, "target_shape": [32, 32, 32]}
torch.manual_seed(7)
model = MilAttentionFineTuneModel(params)
rng = np.random.default_rng(0)
bags = [rng.random((L, 32, 32)).astype(np.float32) for L in lens]
y = np.array([1.0, 0.0, 1.0], dtype=np.float32)[:len(lens)]
dt = torch.float64


def ref_grads():
    rb, ra = copy.deepcopy(model.backbone).to(dt), copy.deepcopy(model.attn).to(dt)
    mean = torch.tensor(model.mean_vals, device="cuda", dtype=dt).view(1, 3, 1, 1)
    std = torch.tensor(model.std_vals, device="cuda", dtype=dt).view(1, 3, 1, 1)
    rb.train(); ra.train()
    feats = []
    for b in bags:
        sl = torch.from_numpy(b).cuda()
        xx = F.interpolate(sl.unsqueeze(1), size=(64, 64), mode="bilinear", align_corners=False).repeat(1, 3, 1, 1).to(dt)
        xx = (xx - mean) / std
        feats.append(torch.cat([rb(xx[i:i + 4]) for i in range(0, xx.shape[0], 4)], dim=0))
    lmax = max(f.shape[0] for f in feats)
    X = torch.zeros(len(bags), lmax, feats[0].shape[1], device="cuda", dtype=dt)
    M = torch.zeros(len(bags), lmax, device="cuda", dtype=dt)
    for i, f in enumerate(feats):
        X[i, :f.shape[0]] = f
        M[i, :f.shape[0]] = 1
    pred = ra(X, M)
    t = torch.from_numpy(y).cuda().to(dt)
    per = F.binary_cross_entropy(pred, t, reduction="none")
    pos = t >= 0.5
    w = (1.0 - torch.where(pos, pred, 1.0 - pred)) ** 2.0 * torch.where(pos, 0.25, 0.75)
    loss = (w * per).mean()
    loss.backward()
    g = {"backbone." + k: v.grad for k, v in rb.named_parameters()}
    g.update({"attn." + k: v.grad for k, v in ra.named_parameters()})
    return float(loss), g, pred.detach()


# free-running float64 reference with torch's own Adam (weights compared after every step)
frb, fra = copy.deepcopy(model.backbone).to(dt), copy.deepcopy(model.attn).to(dt)
fopt = torch.optim.Adam([{"params": frb.parameters(), "lr": 1e-3}, {"params": fra.parameters(), "lr": 3e-3}], weight_decay=1e-3)


def free_step():
    mean = torch.tensor(model.mean_vals, device="cuda", dtype=dt).view(1, 3, 1, 1)
    std = torch.tensor(model.std_vals, device="cuda", dtype=dt).view(1, 3, 1, 1)
    frb.train(); fra.train()
    feats = []
    for b in bags:
        sl = torch.from_numpy(b).cuda()
        xx = F.interpolate(sl.unsqueeze(1), size=(64, 64), mode="bilinear", align_corners=False).repeat(1, 3, 1, 1).to(dt)
        xx = (xx - mean) / std
        feats.append(torch.cat([frb(xx[i:i + 4]) for i in range(0, xx.shape[0], 4)], dim=0))
    lmax = max(f.shape[0] for f in feats)
    X = torch.zeros(len(bags), lmax, feats[0].shape[1], device="cuda", dtype=dt)
    M = torch.zeros(len(bags), lmax, device="cuda", dtype=dt)
    for i, f in enumerate(feats):
        X[i, :f.shape[0]] = f
        M[i, :f.shape[0]] = 1
    pred = fra(X, M)
    t = torch.from_numpy(y).cuda().to(dt)
    per = F.binary_cross_entropy(pred, t, reduction="none")
    pos = t >= 0.5
    w = (1.0 - torch.where(pos, pred, 1.0 - pred)) ** 2.0 * torch.where(pos, 0.25, 0.75)
    loss = (w * per).mean()
    fopt.zero_grad()
    loss.backward()
    total = torch.nn.utils.clip_grad_norm_(list(frb.parameters()) + list(fra.parameters()), CLIP)
    fopt.step()
    return float(loss.detach()), float(total)


for step in range(3):
    fl, fn = free_step()
    l64, g64, pred = ref_grads()                          # on the CURRENT native weights
    rt, ht, opt = model._trainers()
    w_before = {"backbone." + k: v.data.clone() for k, v in model.backbone.named_parameters()}
    loss, prob = model.train_step(bags, y, frozen=False, clip=CLIP)
    torch.cuda.synchronize()
    gn = {"backbone." + k: v for k, v in rt.grad.items()}
    gn.update({"attn." + k: v for k, v in ht.g.items()})
    tot64 = float(torch.sqrt(sum((v.double() ** 2).sum() for v in g64.values())))
    print(f"step {step}: loss native {float(loss):.6f} f64 {l64:.6f} | grad norm native {float(opt._scale[1]):.5f} f64 {tot64:.5f} | prob {prob.cpu().numpy()} ref {pred.cpu().numpy()}")
    errs = sorted(((float((gn[k].double() - g64[k]).norm() / (g64[k].norm() + 1e-30)), float(g64[k].norm()), k) for k in g64), reverse=True)
    for e, nrm, k in errs[:4]:
        print(f"    {k:40s} rel err {e:.3e}  |g64| {nrm:.3e}")
    if step == 0:
        for k in ("backbone.layer3.0.conv2.weight", "backbone.layer2.1.conv2.weight", "backbone.conv1.weight"):
            wcur = w_before[k].double()
            a, b = gn[k].double() + 1e-3 * wcur, g64[k] + 1e-3 * wcur
            flips = (torch.sign(a) != torch.sign(b))
            print(f"    {k}: sign flips of (g + wd*w): {int(flips.sum())} of {a.numel()}; |g+wd*w| of flipped (f64): "
                  f"median {float(b[flips].abs().median()) if flips.any() else 0:.2e} max {float(b[flips].abs().max()) if flips.any() else 0:.2e}; "
                  f"abs grad err: max {float((gn[k].double() - g64[k]).abs().max()):.2e} rms {float((gn[k].double() - g64[k]).pow(2).mean().sqrt()):.2e}; "
                  f"|g64| rms {float(g64[k].pow(2).mean().sqrt()):.2e}; wd*w rms {float((1e-3 * wcur).pow(2).mean().sqrt()):.2e}")
    print(f"  free-running f64: loss {fl:.6f} norm {fn:.5f}")
    wn = {"backbone." + k: v.data for k, v in model.backbone.named_parameters()}
    wn.update({"attn." + k: v.data for k, v in model.attn.named_parameters()})
    wf = {"backbone." + k: v.data for k, v in frb.named_parameters()}
    wf.update({"attn." + k: v.data for k, v in fra.named_parameters()})
    d = sorted(((float((wn[k].double() - wf[k]).norm()), float(wf[k].norm()), wn[k].numel(), k) for k in wf), reverse=True)
    for e, nrm, ne, k in d[:6]:
        print(f"    weights {k:40s} |diff| {e:.3e}  |w| {nrm:.3e}  numel {ne}  diff/sqrt(n) {e / ne ** 0.5:.2e}")
