"""K5 parity: the native training kernels (csrc/train.cu, pd_fusion_b200/training.py) against torch autograd of the SAME model --
the computation the reference's `loss.backward(); clip_grad_norm_; optimizer.step()` performs
(models/mil_attention_finetune.py:209-229, models/mil_attention.py:120-134).  FP32 on both sides."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import torch.nn.functional as F

from pd_fusion_b200.backbone import ResNet2D
from pd_fusion_b200.models.mil_attention import MILAttentionNet
from pd_fusion_b200.training import MilHeadTrainer, NativeAdam, ResNetTrainer


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    """torch's cuDNN convolutions default to TF32 on this GPU: the autograd REFERENCE must run true FP32 to be the yardstick."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def _ref_loss(pred, y, loss_type, pos_weight, gamma, alpha):
    per = F.binary_cross_entropy(pred, y, reduction="none")
    pos = y >= 0.5
    if loss_type == "focal":
        w = (1.0 - torch.where(pos, pred, 1.0 - pred)) ** gamma
        if alpha is not None:
            w = w * torch.where(pos, float(alpha), 1.0 - float(alpha))
        return (w * per).mean()
    if pos_weight is not None:
        return (per * torch.where(pos, float(pos_weight), 1.0)).mean()
    return per.mean()


@pytest.mark.parametrize("gated,loss_type,pos_weight,alpha", [(True, "bce", None, None), (False, "bce", 2.5, None), (True, "focal", None, 0.25),
                                                               (False, "focal", None, None)])
def test_mil_head_gradients_and_adam_trajectory(gated, loss_type, pos_weight, alpha):
    D, H, A, B, L = 96, 48, 24, 7, 13
    torch.manual_seed(3)
    net = MILAttentionNet(D, H, A, 0.0, gated=gated).cuda()
    ref = MILAttentionNet(D, H, A, 0.0, gated=gated).cuda()
    ref.load_state_dict(net.state_dict())
    g = torch.Generator().manual_seed(5)
    lens = torch.tensor([13, 1, 7, 13, 4, 0, 9], dtype=torch.int32)
    lens[5] = 2
    X = torch.randn(B, L, D, generator=g)
    M = torch.zeros(B, L)
    for i in range(B):
        X[i, lens[i]:] = 0
        M[i, :lens[i]] = 1
    y = torch.tensor([1, 0, 1, 1, 0, 0, 1], dtype=torch.float32)
    Xd, Md, yd = X.cuda(), M.cuda(), y.cuda()
    ht = MilHeadTrainer(net, gated)
    opt = NativeAdam([([p for p, _ in ht.param_grads()], 3e-3)], weight_decay=1e-3)
    ropt = torch.optim.Adam(ref.parameters(), lr=3e-3, weight_decay=1e-3)
    ref.train()
    for step in range(3):
        ht.zero_grad()
        loss, prob, dX = ht.forward_backward(Xd, lens, yd, loss_type, pos_weight, 2.0, alpha, need_dx=True)
        Xr = Xd.clone().requires_grad_(True)
        pred = ref(Xr, Md)
        rl = _ref_loss(pred, yd, loss_type, pos_weight, 2.0, alpha)
        ropt.zero_grad()
        rl.backward()
        torch.cuda.synchronize()
        assert abs(float(loss) - float(rl)) < 1e-5 * max(1.0, abs(float(rl))), (step, float(loss), float(rl))
        np.testing.assert_allclose(prob.cpu().numpy(), pred.detach().cpu().numpy(), atol=2e-6)
        names = [k for k, _ in ref.named_parameters()]
        for k, (p_, g_) in zip(names, ht.param_grads()):
            rg = dict(ref.named_parameters())[k].grad
            assert _rel(g_, rg) < 2e-4 or float((g_ - rg).abs().max()) < 1e-7, (step, k, _rel(g_, rg))
        assert _rel(dX, Xr.grad) < 2e-4
        # clip + Adam
        total = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
        pg = [(p_, g_, 3e-3) for p_, g_ in ht.param_grads()]
        sc = opt.clip([g_ for _, g_, _ in pg], 0.05)
        opt.step(pg, sc)
        ropt.step()
        torch.cuda.synchronize()
        assert abs(float(sc[1]) - float(total)) < 1e-4 * float(total)
        for k, (p_, _) in zip(names, ht.param_grads()):
            assert _rel(p_, dict(ref.named_parameters())[k].data) < 1e-5, (step, k)


@pytest.mark.parametrize("arch,n,S,groups", [("resnet18", 6, 64, [0, 4, 6]), ("resnet50", 5, 64, [0, 2, 5])])
def test_backbone_train_forward_backward_vs_autograd(arch, n, S, groups):
    """Train-mode forward (BatchNorm statistics per group of images, running statistics) and the full backward of the backbone:
    embeddings, every parameter gradient and the running statistics against torch autograd run group by group."""
    torch.manual_seed(1234)
    net = ResNet2D(arch)
    net.fc = torch.nn.Identity()
    ref = ResNet2D(arch)
    ref.fc = torch.nn.Identity()
    ref.load_state_dict(net.state_dict())
    net, ref = net.cuda().float(), ref.cuda().double()               # yardstick: the same network in float64
    net.train(); ref.train()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, 3, S, S, generator=g).cuda()
    Rw = torch.randn(n, 512 if arch == "resnet18" else 2048, generator=g).cuda()
    rt = ResNetTrainer(net, arch, S)
    emb = rt.forward(x.permute(0, 2, 3, 1).contiguous(), groups)
    rt.zero_grad()
    rt.backward(Rw)
    outs = [ref(x[groups[i]:groups[i + 1]].double()) for i in range(len(groups) - 1)]        # the reference forwards chunk by chunk
    remb = torch.cat(outs, dim=0)
    (remb * Rw.double()).sum().backward()
    torch.cuda.synchronize()
    assert _rel(emb.double(), remb.detach()) < 1e-4, _rel(emb.double(), remb.detach())
    rp = dict(ref.named_parameters())
    worst = 0.0
    for k, gk in rt.grad.items():
        e = _rel(gk.double(), rp[k].grad)
        worst = max(worst, e)
        # float32 rounding is amplified on the way down through train-mode BatchNorm over tiny groups (8-32 samples per channel in
        # the last stage of this toy): 1e-3 in stage 4, up to ~1e-2 at the stem -- an indexing error would show as O(1)
        assert e < (2e-3 if k.startswith("layer4") else 2e-2), (k, e)
    rb, nb = dict(ref.named_buffers()), dict(net.named_buffers())
    for k in rb:
        if k.endswith("num_batches_tracked"):
            assert int(rb[k]) == int(nb[k]), k
        else:
            assert _rel(nb[k].double(), rb[k]) < 1e-4, (k, _rel(nb[k].double(), rb[k]))
    print("worst gradient rel err", worst)


def test_finetune_model_training_steps_vs_autograd(tmp_path):
    """MilAttentionFineTuneModel.train_step (slices -> backbone in train mode, 16-slice chunks -> MIL head -> focal loss ->
    backward -> clip -> Adam with two learning-rate groups) against the same three steps done by torch autograd + torch.optim.Adam on a
    copy of the model: loss, gradient norm and weight trajectory."""
    from pd_fusion_b200.models.mil_attention_finetune import MilAttentionFineTuneModel
    params = {"backbone": "resnet18", "pretrained": False, "input_size": 64, "hidden_dim": 32, "attn_dim": 16, "dropout": 0.0, "gated": True,
              "batch_size": 3, "slice_batch_size": 4, "lr_backbone": 1e-3, "lr": 3e-3, "weight_decay": 1e-3, "loss_type": "focal",
              "focal_gamma": 2.0, "focal_alpha": 0.25, "train_aug": False, "max_grad_norm": 1.0, "slice_count": 6, "target_shape": [32, 32, 32]}
    torch.manual_seed(7)
    model = MilAttentionFineTuneModel(params)
    import copy
    rb, ra = copy.deepcopy(model.backbone).float(), copy.deepcopy(model.attn).float()
    ropt = torch.optim.Adam([{"params": rb.parameters(), "lr": 1e-3}, {"params": ra.parameters(), "lr": 3e-3}], weight_decay=1e-3)
    rng = np.random.default_rng(0)
    bags = [rng.random((L, 32, 32)).astype(np.float32) for L in (6, 5, 6)]
    y = np.array([1.0, 0.0, 1.0], dtype=np.float32)
    mean = torch.tensor(model.mean_vals, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(model.std_vals, device="cuda").view(1, 3, 1, 1)
    for step in range(3):
        loss, prob = model.train_step(bags, y, frozen=False, clip=1.0)
        rb.train(); ra.train()
        feats = []
        for b in bags:
            sl = torch.from_numpy(b).cuda()
            xx = F.interpolate(sl.unsqueeze(1), size=(64, 64), mode="bilinear", align_corners=False).repeat(1, 3, 1, 1)
            xx = (xx - mean) / std
            feats.append(torch.cat([rb(xx[i:i + 4]) for i in range(0, xx.shape[0], 4)], dim=0))
        lmax = max(f.shape[0] for f in feats)
        X = torch.zeros(3, lmax, feats[0].shape[1], device="cuda")
        M = torch.zeros(3, lmax, device="cuda")
        for i, f in enumerate(feats):
            X[i, :f.shape[0]] = f
            M[i, :f.shape[0]] = 1
        rl = _ref_loss(ra(X, M), torch.from_numpy(y).cuda(), "focal", None, 2.0, 0.25)
        ropt.zero_grad()
        rl.backward()
        total = torch.nn.utils.clip_grad_norm_(list(rb.parameters()) + list(ra.parameters()), 1.0)
        ropt.step()
        torch.cuda.synchronize()
        assert abs(float(loss) - float(rl)) < 1e-2 * max(abs(float(rl)), 1e-3), (step, float(loss), float(rl))
        norm = float(model._trainers()[2]._scale[1])
        assert abs(norm - float(total)) < 1e-2 * float(total), (step, norm, float(total))
        for (k, p_), (_, q_) in zip(model.backbone.named_parameters(), rb.named_parameters()):
            assert _rel(p_.data, q_.data) < 1e-2, (step, k, _rel(p_.data, q_.data))
        for (k, p_), (_, q_) in zip(model.attn.named_parameters(), ra.named_parameters()):
            assert _rel(p_.data, q_.data) < 1e-2, (step, k, _rel(p_.data, q_.data))
    # frozen step: only the head moves, BatchNorm running statistics still update
    before = {k: v.detach().clone() for k, v in model.backbone.state_dict().items()}
    model.train_step(bags, y, frozen=True, clip=1.0)
    after = model.backbone.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before if "running" not in k and "num_batches" not in k)
    assert any(not torch.equal(before[k], after[k]) for k in before if "running_mean" in k)


def test_mil_attention_model_train_is_native_and_learns():
    """MilAttentionModel.train runs on the native kernels (launch counter moves, no autograd graph) and fits separable bags."""
    from pd_fusion_b200 import _lib
    from pd_fusion_b200.models.mil_attention import MilAttentionModel
    rng = np.random.default_rng(1)
    y = (np.arange(40) % 2).astype(int)
    bags = [(rng.standard_normal((int(rng.integers(3, 9)), 32)) + (1.5 if y[i] else -1.5)).astype(np.float32) for i in range(40)]
    torch.manual_seed(0)
    m = MilAttentionModel(32, {"hidden_dim": 32, "attn_dim": 16, "gated": True, "dropout": 0.1, "epochs": 8, "batch_size": 8, "lr": 5e-3,
                               "max_grad_norm": 1.0})
    l0 = _lib.launch_count()
    m.train(bags, y)
    assert _lib.launch_count() - l0 > 8 * 5 * 10
    assert all(p.grad is None for p in m.model.parameters())                  # no autograd involved
    p = m.predict_proba(bags)
    assert ((p > 0.5).astype(int) == y).mean() > 0.95
