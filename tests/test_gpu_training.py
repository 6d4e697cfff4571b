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


@pytest.mark.parametrize("arch,n,S,groups", [("resnet18", 6, 64, [0, 4, 6]), ("resnet50", 5, 96, [0, 2, 5])])
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
    rt = ResNetTrainer(net, arch, S, precision="fp32")
    emb = rt.forward(x.permute(0, 2, 3, 1).contiguous(), groups)
    rt.zero_grad()
    rt.backward(Rw)
    outs = [ref(x[groups[i]:groups[i + 1]].double()) for i in range(len(groups) - 1)]        # the reference forwards chunk by chunk
    remb = torch.cat(outs, dim=0)
    (remb * Rw.double()).sum().backward()
    torch.cuda.synchronize()
    assert _rel(emb.double(), remb.detach()) < 1e-4, _rel(emb.double(), remb.detach())
    rp = dict(ref.named_parameters())
    # calibration: torch's own float32 autograd of the same network against the float64 yardstick.  Rounding is amplified on the way
    # down through train-mode BatchNorm over tiny groups (8-32 samples per channel in the last stage of this toy); the native
    # gradients must be as close to float64 as torch's float32 ones are (x5, floor 5e-3) -- an indexing error would show as O(1).
    ref32 = ResNet2D(arch)
    ref32.fc = torch.nn.Identity()
    ref32.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    ref32 = ref32.cuda().float().train()
    (torch.cat([ref32(x[groups[i]:groups[i + 1]]) for i in range(len(groups) - 1)], dim=0) * Rw).sum().backward()
    r32 = dict(ref32.named_parameters())
    worst = 0.0
    for k, gk in rt.grad.items():
        e = _rel(gk.double(), rp[k].grad)
        e32 = _rel(r32[k].grad.double(), rp[k].grad)
        worst = max(worst, e)
        assert e < max(5.0 * e32, 5e-3), (k, e, e32)
    rb, nb = dict(ref.named_buffers()), dict(net.named_buffers())
    for k in rb:
        if k.endswith("num_batches_tracked"):
            assert int(rb[k]) == int(nb[k]), k
        else:
            assert _rel(nb[k].double(), rb[k]) < 1e-4, (k, _rel(nb[k].double(), rb[k]))
    print("worst gradient rel err", worst)


@pytest.mark.parametrize("arch,n,S,groups", [("resnet18", 24, 64, [0, 16, 24]), ("resnet50", 20, 96, [0, 8, 20])])
def test_backbone_train_tensor_path_vs_autograd(arch, n, S, groups):
    """The bf16 tensor-core training path (forward / data gradient through the tcgen05 implicit-GEMM kernels, weight gradient through
    wgrad_tc.cu; BatchNorm, activations, gradients and master weights in f32) against torch autograd of the same network in float64.
    Calibration: torch's OWN bf16 mixed precision (autocast) of the same network -- the native path must be as close to float64 as
    that is (x3; floors 2e-2 on the embeddings, 5e-2 on a gradient).  An indexing / rotation / dilation error shows as O(1)."""
    torch.manual_seed(4321)
    net = ResNet2D(arch)
    net.fc = torch.nn.Identity()
    ref = ResNet2D(arch)
    ref.fc = torch.nn.Identity()
    ref.load_state_dict(net.state_dict())
    net, ref = net.cuda().float(), ref.cuda().double()
    net.train(); ref.train()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, 3, S, S, generator=g).cuda()
    Rw = torch.randn(n, 512 if arch == "resnet18" else 2048, generator=g).cuda()
    rt = ResNetTrainer(net, arch, S, precision="bf16")
    assert rt.bf16 and rt.wk16["conv1"].shape == (64, 192)
    emb = rt.forward(x.permute(0, 2, 3, 1).contiguous(), groups)
    rt.zero_grad()
    rt.backward(Rw)
    remb = torch.cat([ref(x[groups[i]:groups[i + 1]].double()) for i in range(len(groups) - 1)], dim=0)
    (remb * Rw.double()).sum().backward()
    rp = dict(ref.named_parameters())
    amp = ResNet2D(arch)
    amp.fc = torch.nn.Identity()
    amp.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    amp = amp.cuda().float().train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        aemb = torch.cat([amp(x[groups[i]:groups[i + 1]]) for i in range(len(groups) - 1)], dim=0)
    (aemb.float() * Rw).sum().backward()
    ap = dict(amp.named_parameters())
    torch.cuda.synchronize()
    e_emb, a_emb = _rel(emb.double(), remb.detach()), _rel(aemb.double(), remb.detach())
    assert e_emb < max(3 * a_emb, 2e-2), (e_emb, a_emb)
    worst = (0.0, 0.0, "")
    for k, gk in rt.grad.items():
        e, ea = _rel(gk.double(), rp[k].grad), _rel(ap[k].grad.double(), rp[k].grad)
        worst = max(worst, (e, ea, k))
        assert e < max(3.0 * ea, 5e-2), (k, e, ea)
    print("tensor path: embeddings rel err", e_emb, "(autocast", a_emb, ") worst gradient", worst)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_finetune_model_training_steps_vs_autograd(tmp_path, monkeypatch, precision):
    """MilAttentionFineTuneModel.train_step (slices -> backbone in train mode, 16-slice chunks -> MIL head -> focal loss ->
    backward -> clip -> Adam with two learning-rate groups), three free-running steps, against the same three steps done by torch
    autograd + torch.optim.Adam on copies of the model in float64 (the yardstick) and in float32 (the calibration: Adam's normalised
    updates amplify rounding noise of ill-conditioned gradients -- BatchNorm biases -- so two float32 trajectories drift apart at the
    percent level; the native one must stay as close to float64 as torch's own float32 run does, x3, floor 1e-2)."""
    import copy
    from pd_fusion_b200.models.mil_attention_finetune import MilAttentionFineTuneModel
    monkeypatch.setenv("PD_FUSION_B200_TRAIN_PRECISION", precision)
    # bf16 tensor path: the calibration run is torch's own bf16 autocast of the backbone; floors 5e-2 on the first step instead of
    # 1e-2, and 0.15 on the free-running steps after it -- two bf16 trajectories of this toy (BatchNorm over 4-image groups, Adam's
    # normalised updates) drift apart by several percent in the gradient norm within three steps, whichever arithmetic produces
    # them (atomics order alone moves it); the per-gradient check of the tensor path is test_backbone_train_tensor_path_vs_autograd
    floors = [1e-2] * 3 if precision == "fp32" else [5e-2, 0.15, 0.15]
    CLIP = 50.0
    params = {"backbone": "resnet18", "pretrained": False, "input_size": 64, "hidden_dim": 32, "attn_dim": 16, "dropout": 0.0, "gated": True,
              "batch_size": 3, "slice_batch_size": 4, "lr_backbone": 1e-3, "lr": 3e-3, "weight_decay": 1e-3, "loss_type": "focal",
              "focal_gamma": 2.0, "focal_alpha": 0.25, "train_aug": False, "max_grad_norm": CLIP, "slice_count": 6, "target_shape": [32, 32, 32]}
    torch.manual_seed(7)
    model = MilAttentionFineTuneModel(params)
    refs = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        rb, ra = copy.deepcopy(model.backbone).to(dt), copy.deepcopy(model.attn).to(dt)
        refs[name] = (rb, ra, torch.optim.Adam([{"params": rb.parameters(), "lr": 1e-3}, {"params": ra.parameters(), "lr": 3e-3}], weight_decay=1e-3), dt)
    rng = np.random.default_rng(0)
    bags = [rng.random((L, 32, 32)).astype(np.float32) for L in (6, 5, 6)]
    y = np.array([1.0, 0.0, 1.0], dtype=np.float32)

    def ref_step(rb, ra, ropt, dt):
        rb.train(); ra.train()
        feats = []
        for b in bags:
            # the SAME network input on both sides (the native resize kernel's output, itself within 1e-5 of F.interpolate --
            # tests/test_gpu_preproc.py): train-mode BatchNorm over 4-image groups amplifies a 1e-6 input difference a thousandfold
            sl = torch.from_numpy(b).cuda()
            L = int(sl.shape[0])
            xin = torch.empty((1, L, 64, 64, 3), dtype=torch.float32, device="cuda")
            model._train_resizer().resize_slices(sl.view(1, L, 32, 32).contiguous(), out=xin)
            xx = xin[0].permute(0, 3, 1, 2).to(dt)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(precision == "bf16" and dt == torch.float32)):
                f = torch.cat([rb(xx[i:i + 4]) for i in range(0, xx.shape[0], 4)], dim=0)
            feats.append(f.to(dt))
        lmax = max(f.shape[0] for f in feats)
        X = torch.zeros(3, lmax, feats[0].shape[1], device="cuda", dtype=dt)
        M = torch.zeros(3, lmax, device="cuda", dtype=dt)
        for i, f in enumerate(feats):
            X[i, :f.shape[0]] = f
            M[i, :f.shape[0]] = 1
        rl = _ref_loss(ra(X, M), torch.from_numpy(y).cuda().to(dt), "focal", None, 2.0, 0.25)
        ropt.zero_grad()
        rl.backward()
        total = torch.nn.utils.clip_grad_norm_(list(rb.parameters()) + list(ra.parameters()), CLIP)
        ropt.step()
        return float(rl.detach()), float(total)

    for step in range(3):
        floor = floors[step]
        loss, prob = model.train_step(bags, y, frozen=False, clip=CLIP)
        l64, n64 = ref_step(*refs["f64"])
        l32, n32 = ref_step(*refs["f32"])
        torch.cuda.synchronize()
        norm = float(model._trainers()[2]._scale[1])
        assert abs(float(loss) - l64) < max(3 * abs(l32 - l64), floor * abs(l64)), (step, float(loss), l32, l64)
        assert abs(norm - n64) < max(3 * abs(n32 - n64), floor * n64), (step, norm, n32, n64)
        p64 = dict(list(refs["f64"][0].named_parameters()) + list(refs["f64"][1].named_parameters()))
        p32 = dict(list(refs["f32"][0].named_parameters()) + list(refs["f32"][1].named_parameters()))
        for k, p_ in list(model.backbone.named_parameters()) + list(model.attn.named_parameters()):
            err = float((p_.data.double() - p64[k].data).norm())
            err32 = float((p32[k].data.double() - p64[k].data).norm())
            assert err < max(3 * err32, floor * float(p64[k].data.norm())), (step, k, err, err32, float(p64[k].data.norm()))
    # frozen step: only the head moves, BatchNorm running statistics still update
    before = {k: v.detach().clone() for k, v in model.backbone.state_dict().items()}
    model.train_step(bags, y, frozen=True, clip=CLIP)
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


def test_moddrop_native_training_matches_autograd():
    """ModalityDropoutModel.train (reference loop: torch.randperm batches, np.random.rand() modality dropout, BCELoss, Adam) on the
    native kernels against the same loop done by torch autograd -- identical RNG draws, dropout 0: weight trajectories agree."""
    import copy
    from pd_fusion_b200.models.fusion_moddrop import ModalityDropoutModel
    dims = {"clinical": 6, "datspect": 4, "mri": 22}
    params = {"hidden_dims": [48, 24], "dropout": 0.0, "lr": 3e-3, "batch_size": 16, "epochs": 3, "moddrop_rate": 0.3, "weight_decay": 1e-4}
    rng = np.random.default_rng(2)
    X = rng.standard_normal((90, 32)).astype(np.float32)
    y = (X[:, 0] - X[:, 7] > 0).astype(np.float32)
    torch.manual_seed(11)
    m = ModalityDropoutModel(dims, params)
    ref = copy.deepcopy(m.model).cuda()
    ropt = torch.optim.Adam(ref.parameters(), lr=3e-3, weight_decay=1e-4)
    torch.manual_seed(21); np.random.seed(21)
    m.train(X, y)
    torch.manual_seed(21); np.random.seed(21)
    Xt, yt = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda().view(-1, 1)
    for _ in range(3):
        ref.train()
        order = torch.randperm(len(Xt))
        for i in range(0, len(Xt), 16):
            sel = order[i:i + 16].cuda()
            ropt.zero_grad()
            F.binary_cross_entropy(ref(Xt[sel], training_dropout=True, drop_rate=0.3), yt[sel]).backward()
            ropt.step()
    for (k, a), (_, b) in zip(m.model.named_parameters(), ref.named_parameters()):
        assert _rel(a.data, b.data) < 2e-4, (k, _rel(a.data, b.data))
    p = m.predict_proba(X)
    assert p.shape == (90,) and np.isfinite(p).all()


def test_moe_native_training_matches_autograd():
    """MoEModel.train (full-batch steps, router softmax gating, BCELoss, Adam) on the native kernels against torch autograd."""
    import copy
    from pd_fusion_b200.models.moe import MoEModel
    dims = {"clinical": 10, "mri": 40}
    params = {"expert_hidden_dims": [32, 16], "router_hidden_dims": [16], "lr": 3e-3, "epochs": 5, "weight_decay": 1e-4}
    rng = np.random.default_rng(4)
    N = 300
    Xd = {"clinical": rng.standard_normal((N, 10)).astype(np.float32), "mri": rng.standard_normal((N, 40)).astype(np.float32)}
    mask = (rng.random((N, 2)) < 0.8).astype(np.float32)
    Xd = {m: Xd[m] * mask[:, i:i + 1] for i, m in enumerate(sorted(Xd))}
    y = (Xd["clinical"][:, 0] + Xd["mri"][:, 1] > 0).astype(np.float32)
    torch.manual_seed(5)
    m = MoEModel(dims, params)
    ref = copy.deepcopy(m.model).cuda()
    ropt = torch.optim.Adam(ref.parameters(), lr=3e-3, weight_decay=1e-4)
    m.train({k: torch.from_numpy(v) for k, v in Xd.items()}, y, torch.from_numpy(mask))
    Xt = {k: torch.from_numpy(v).cuda() for k, v in Xd.items()}
    mt, yt = torch.from_numpy(mask).cuda(), torch.from_numpy(y).cuda().view(-1, 1)
    losses = []
    for _ in range(5):
        ref.train()
        ropt.zero_grad()
        loss = F.binary_cross_entropy(ref(Xt, mt), yt)
        loss.backward()
        ropt.step()
        losses.append(float(loss))
    np.testing.assert_allclose([float(v) for v in m.last_losses], losses, rtol=2e-5)
    for (k, a), (_, b) in zip(m.model.named_parameters(), ref.named_parameters()):
        assert _rel(a.data, b.data) < 2e-4, (k, _rel(a.data, b.data))
