"""Host side of K2: ResNet18/50 (torchvision state_dict layout, fc = Identity, eval mode) lowered to a
flat list of NHWC ops executed by libpdfusion_b200.so.

Reference: `_build_resnet_backbone` + the chunked `model(batch)` loop, data/openneuro_features.py:153-164,
257-262.  The state_dict key layout is the weight-interchange format: any torchvision
ResNet18/50 state_dict loads unchanged.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from . import _lib

RESNET_SPECS = {"resnet18": ("basic", [2, 2, 2, 2], 512), "resnet50": ("bottleneck", [3, 4, 6, 3], 2048)}
_PLANES = [64, 128, 256, 512]


# ----------------------------------------------------------------------------------------------------
# state_dict-compatible module (used only to create / hold weights; the forward pass below is for
# completeness of the nn.Module contract, the product path never calls it)
# ----------------------------------------------------------------------------------------------------
class _Basic(nn.Module):
    expansion = 1

    def __init__(self, inp, planes, stride, down):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = down

    def forward(self, x):
        idn = x if self.downsample is None else self.downsample(x)
        o = self.relu(self.bn1(self.conv1(x)))
        return self.relu(self.bn2(self.conv2(o)) + idn)


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inp, planes, stride, down):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = down

    def forward(self, x):
        idn = x if self.downsample is None else self.downsample(x)
        o = self.relu(self.bn1(self.conv1(x)))
        o = self.relu(self.bn2(self.conv2(o)))
        return self.relu(self.bn3(self.conv3(o)) + idn)


class ResNet2D(nn.Module):
    """torchvision.models.resnet.ResNet with identical parameter names, construction order and
    initialisation, so `torch.manual_seed(s); ResNet2D(arch)` reproduces torchvision's random init."""

    def __init__(self, arch: str):
        super().__init__()
        kind, layers, _ = RESNET_SPECS[arch]
        block = _Basic if kind == "basic" else _Bottleneck
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        for i, (planes, n) in enumerate(zip(_PLANES, layers), start=1):
            setattr(self, f"layer{i}", self._make_layer(block, planes, n, 1 if i == 1 else 2))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, 1000)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                                 nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, 1, None))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def detect_arch(sd: Dict[str, torch.Tensor]) -> str:
    return "resnet50" if "layer1.0.conv3.weight" in sd else "resnet18"


# ----------------------------------------------------------------------------------------------------
# lowering
# ----------------------------------------------------------------------------------------------------
def _bn_affine(sd, prefix: str, eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """eval-mode BatchNorm as y = x*scale + shift (float64 maths)."""
    g, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    mu, var = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    scale = g / torch.sqrt(var + eps)
    return scale, b - mu * scale


def conv_list(arch: str) -> List[dict]:
    """Every conv of the network in execution order with its wiring."""
    kind, layers, _ = RESNET_SPECS[arch]
    out = [dict(name="conv1", bn="bn1", cin=3, cout=64, k=7, stride=2, pad=3, role="stem")]
    inp = 64
    for li, (planes, n) in enumerate(zip(_PLANES, layers), start=1):
        for bi in range(n):
            p = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            exp = 1 if kind == "basic" else 4
            has_down = bi == 0 and (stride != 1 or inp != planes * exp)
            if kind == "basic":
                out.append(dict(name=p + ".conv1", bn=p + ".bn1", cin=inp, cout=planes, k=3, stride=stride, pad=1, role="a", block=p, down=has_down))
                out.append(dict(name=p + ".conv2", bn=p + ".bn2", cin=planes, cout=planes, k=3, stride=1, pad=1, role="last", block=p))
            else:
                out.append(dict(name=p + ".conv1", bn=p + ".bn1", cin=inp, cout=planes, k=1, stride=1, pad=0, role="a", block=p, down=has_down))
                out.append(dict(name=p + ".conv2", bn=p + ".bn2", cin=planes, cout=planes, k=3, stride=stride, pad=1, role="b", block=p))
                out.append(dict(name=p + ".conv3", bn=p + ".bn3", cin=planes, cout=planes * 4, k=1, stride=1, pad=0, role="last", block=p))
            if has_down:
                out.append(dict(name=p + ".downsample.0", bn=p + ".downsample.1", cin=inp, cout=planes * exp, k=1, stride=stride, pad=0, role="down", block=p))
            inp = planes * exp
    return out


def flops_per_image(arch: str, input_size: int = 224, folded_stem: bool = False) -> float:
    """2*MAC over all convs at 3 x S x S (3 627 122 688 r18 / 8 174 272 512 r50 at 224)."""
    convs = conv_list(arch)
    ho = (input_size + 6 - 7) // 2 + 1
    total = 2.0 * ho * ho * 64 * 49 * (1 if folded_stem else 3)
    size = (ho + 2 - 3) // 2 + 1
    block_in = t = size
    for cv in convs[1:]:
        if cv["role"] == "a":
            block_in = t = size
        hin = block_in if cv["role"] in ("a", "down") else t
        ho = (hin + 2 * cv["pad"] - cv["k"]) // cv["stride"] + 1
        total += 2.0 * ho * ho * cv["cout"] * cv["k"] ** 2 * cv["cin"]
        if cv["role"] in ("a", "b"):
            t = ho
        elif cv["role"] == "last":
            t = size = ho
    return total


class ResNetEncoder:
    """Runs `n_images` slices [n, S, S] through the backbone and returns [n, D] f32 embeddings.

    precision "bf16": tcgen05/TMEM implicit-GEMM convs, BN folded into bf16 weights + f32 bias, stem as a
        one-channel 7x7 (the 3 identical input channels folded: needs channel-uniform mean/std);
    precision "fp32": CUDA-core FFMA convs on the exact 3-channel input, BN applied as scale/shift in the epilogue.
    """

    def __init__(self, state_dict: Dict[str, torch.Tensor], n_images: int, input_size: int = 224,
                 precision: str = "bf16", arch: str | None = None, device=None, fused_stem: bool = True):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.arch = arch or detect_arch(state_dict)
        self.kind, self.layers, self.emb_dim = RESNET_SPECS[self.arch]
        self.n = int(n_images)
        self.S = int(input_size)
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.bf16 = precision == "bf16"
        self.precision = precision
        self.fused_stem = bool(fused_stem)
        sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        self._keep: List[torch.Tensor] = []     # device tensors referenced by the plan
        self._build(sd)

    # -- weights ---------------------------------------------------------------------------------
    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        d = t.contiguous().to(self.device)
        self._keep.append(d)
        return d

    def _conv_weights(self, sd, cv):
        w = sd[cv["name"] + ".weight"].double()                      # [K, C, R, S]
        scale, shift = _bn_affine(sd, cv["bn"])
        if self.bf16:
            wf = w * scale.view(-1, 1, 1, 1)
            if cv["role"] == "stem":
                w1 = wf.sum(dim=1)                                      # [64, 7, 7]: 3 identical channels folded
                mat = torch.zeros(w1.shape[0], 8, 8, dtype=torch.float64)      # K index = r*8 + s (stem_tc.cu)
                mat[:, :7, :7] = w1
                mat = mat.reshape(w1.shape[0], 64)
                return self._dev(mat.to(torch.bfloat16)), None, self._dev(shift.float())
            return self._dev(wf.permute(0, 2, 3, 1).to(torch.bfloat16)), None, self._dev(shift.float())   # [K,R,S,C]
        return self._dev(w.permute(2, 3, 1, 0).float()), self._dev(scale.float()), self._dev(shift.float())  # [R,S,C,K]

    # -- op list ---------------------------------------------------------------------------------
    def _build(self, sd):
        n, S = self.n, self.S
        prec = _lib.PREC_BF16 if self.bf16 else _lib.PREC_F32
        esz = 2 if self.bf16 else 4
        dt = torch.bfloat16 if self.bf16 else torch.float32
        h1 = (S + 6 - 7) // 2 + 1
        h2 = (h1 + 2 - 3) // 2 + 1
        exp = 1 if self.kind == "basic" else 4
        max_elems = max(n * h1 * h1 * 64, n * h2 * h2 * 64 * exp)
        slot_bytes = max_elems * esz
        hf = h2
        for _ in range(3):
            hf = (hf + 2 - 3) // 2 + 1
        slot_bytes = max(slot_bytes, n * hf * hf * self.emb_dim * 4)
        self.slots = [torch.empty(slot_bytes, dtype=torch.uint8, device=self.device) for _ in range(5)]
        free = list(range(5))
        if self.bf16:
            self.input = torch.empty((n, S, S), dtype=torch.bfloat16, device=self.device)
        else:
            self.input = torch.empty((n, S, S, 3), dtype=torch.float32, device=self.device)
        self.output = torch.empty((n, self.emb_dim), dtype=torch.float32, device=self.device)
        ops: List[_lib.Op] = []
        self.op_names: List[str] = []

        def add(name, **kw):
            op = _lib.Op()
            for k, v in kw.items():
                setattr(op, k, v)
            ops.append(op)
            self.op_names.append(name)

        def sp(i):
            return self.slots[i].data_ptr()

        convs = conv_list(self.arch)
        # ---- stem
        stem = convs[0]
        w, sc, b = self._conv_weights(sd, stem)
        if self.bf16 and self.fused_stem:
            # conv1 + bn1 + relu + maxpool in one kernel (stem_tc.cu): the patch matrix never leaves shared memory
            s_pool = free.pop(0)
            add("stem.fused", kind=_lib.OP_STEM_FUSED, precision=prec, n=n, h=S, w=S, c=1, k=64, r=7, s=7, stride=2, pad=3,
                ho=h2, wo=h2, relu=1, d_in=self.input.data_ptr(), d_weight=w.data_ptr(), d_bias=b.data_ptr(), d_out=sp(s_pool))
        else:
            s_out = free.pop(0)
            if self.bf16:
                s_col = free.pop(0)
                add("stem.im2col", kind=_lib.OP_STEM_IM2COL, precision=prec, n=n, h=S, w=S, c=1, k=64, r=7, s=7, stride=2, pad=3,
                    ho=h1, wo=h1, d_in=self.input.data_ptr(), d_out=sp(s_col))
                add("conv1", kind=_lib.OP_CONV, precision=prec, n=n, h=h1, w=h1, c=64, k=64, r=1, s=1, stride=1, pad=0, ho=h1, wo=h1,
                    relu=1, d_in=sp(s_col), d_weight=w.data_ptr(), d_bias=b.data_ptr(), d_out=sp(s_out))
                free.append(s_col)
            else:
                add("conv1", kind=_lib.OP_CONV, precision=prec, n=n, h=S, w=S, c=3, k=64, r=7, s=7, stride=2, pad=3, ho=h1, wo=h1,
                    relu=1, d_in=self.input.data_ptr(), d_weight=w.data_ptr(), d_scale=sc.data_ptr(), d_bias=b.data_ptr(), d_out=sp(s_out))
            s_pool = free.pop(0)
            add("maxpool", kind=_lib.OP_MAXPOOL, precision=prec, n=n, h=h1, w=h1, c=64, ho=h2, wo=h2, d_in=sp(s_out), d_out=sp(s_pool))
            free.append(s_out)
        cur, cur_h, cur_c = s_pool, h2, 64
        # ---- residual stages
        blocks: Dict[str, List[dict]] = {}
        order: List[str] = []
        for cv in convs[1:]:
            if cv["block"] not in blocks:
                blocks[cv["block"]] = []
                order.append(cv["block"])
            blocks[cv["block"]].append(cv)
        last_block = order[-1]
        for bname in order:
            cvs = {cv["role"]: cv for cv in blocks[bname]}
            x_slot, x_h, x_c = cur, cur_h, cur_c
            idn_slot = x_slot
            if "down" in cvs:
                cv = cvs["down"]
                ho = (x_h + 2 * cv["pad"] - cv["k"]) // cv["stride"] + 1
                w, sc, b = self._conv_weights(sd, cv)
                idn_slot = free.pop(0)
                add(cv["name"], kind=_lib.OP_CONV, precision=prec, n=n, h=x_h, w=x_h, c=x_c, k=cv["cout"], r=1, s=1, stride=cv["stride"],
                    pad=0, ho=ho, wo=ho, relu=0, d_in=sp(x_slot), d_weight=w.data_ptr(), d_scale=_lib.ptr(sc), d_bias=b.data_ptr(),
                    d_out=sp(idn_slot))
            t_slot, t_h, t_c = x_slot, x_h, x_c
            seq = [cvs["a"]] + ([cvs["b"]] if "b" in cvs else []) + [cvs["last"]]
            for cv in seq:
                ho = (t_h + 2 * cv["pad"] - cv["k"]) // cv["stride"] + 1
                w, sc, b = self._conv_weights(sd, cv)
                o_slot = free.pop(0)
                is_last = cv["role"] == "last"
                final = is_last and bname == last_block
                add(cv["name"], kind=_lib.OP_CONV, precision=prec, n=n, h=t_h, w=t_h, c=t_c, k=cv["cout"], r=cv["k"], s=cv["k"],
                    stride=cv["stride"], pad=cv["pad"], ho=ho, wo=ho, relu=1, out_f32=1 if (final and self.bf16) else 0,
                    d_in=sp(t_slot), d_weight=w.data_ptr(), d_scale=_lib.ptr(sc), d_bias=b.data_ptr(),
                    d_residual=sp(idn_slot) if is_last else None, d_out=sp(o_slot))
                if t_slot != x_slot:
                    free.append(t_slot)
                t_slot, t_h, t_c = o_slot, ho, cv["cout"]
            if idn_slot != x_slot:
                free.append(idn_slot)
            free.append(x_slot)
            cur, cur_h, cur_c = t_slot, t_h, t_c
        add("avgpool", kind=_lib.OP_AVGPOOL, precision=prec, n=n, h=cur_h, w=cur_h, c=cur_c, out_f32=1 if self.bf16 else 0,
            d_in=sp(cur), d_out=self.output.data_ptr())
        self.final_hw = cur_h
        arr = (_lib.Op * len(ops))(*ops)
        plan = C.c_void_p()
        _lib.check(self.lib.pdf_plan_create(C.byref(plan), arr, len(ops)), "pdf_plan_create")
        self.plan = plan
        self.n_ops = len(ops)
        self.ops = ops

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.pdf_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    @property
    def flops(self) -> float:
        """FLOPs actually executed by one run (stem folded to one channel, K padded to 64, on the bf16 path)."""
        return float(self.lib.pdf_plan_flops(self.plan))

    def algorithmic_flops(self) -> float:
        """SURVEY.md 8(d) figure for the work done: reference FLOPs, minus 2/3 of the stem when it is folded."""
        return self.n * flops_per_image(self.arch, self.S, folded_stem=self.bf16)

    def forward(self, x: torch.Tensor | None = None) -> torch.Tensor:
        """x: [n,S,S] bf16 (bf16 path) or [n,S,S,3] f32 (fp32 path); None = data already in self.input.
        Enqueues on the current stream, returns the (reused) output buffer [n, D] f32."""
        if x is not None and x.data_ptr() != self.input.data_ptr():
            self.input.copy_(x.reshape(self.input.shape))
        _lib.check(self.lib.pdf_plan_run(self.plan, _lib.stream_ptr()), "pdf_plan_run")
        return self.output

    def run_range(self, first: int, count: int) -> None:
        _lib.check(self.lib.pdf_plan_run_range(self.plan, first, count, _lib.stream_ptr()), "pdf_plan_run_range")
