"""Host side of K3/K4: thin wrappers that marshal state_dict tensors into the C structs of
pdf_mil_forward / pdf_moddrop_sweep / pdf_moe_sweep.  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def _dev_f32(t, device) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class MilHead:
    """MILAttentionNet (models/mil_attention.py:10-51) in eval mode for a batch of padded bags.

    precision "fp32": FFMA GEMMs, <= 5e-6 against the reference; "tf32": both linear layers as tcgen05 kind::tf32 GEMMs on the
    f32 bags (the throughput path; needs H, (2)A in {64,128,256} and D, H multiples of 32 -- `tensor_path_ok`)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], gated: bool, missing_prob: float = 0.5, device=None,
                 precision: str = "fp32"):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        g = {k: _dev_f32(v, self.device) for k, v in state_dict.items()}
        self._keep = g
        w = _lib.MilWeights()
        w.H, w.D = g["instance.0.weight"].shape
        w.gated = 1 if gated else 0
        w.w_inst, w.b_inst = g["instance.0.weight"].data_ptr(), g["instance.0.bias"].data_ptr()
        if gated:
            w.A = g["attn_v.0.weight"].shape[0]
            # [W_v; W_u] and [b_v; b_u] packed: the tensor path runs both attention layers as ONE [2A, H] GEMM
            g["_attn_vu.weight"] = torch.cat([g["attn_v.0.weight"], g["attn_u.0.weight"]], dim=0).contiguous()
            g["_attn_vu.bias"] = torch.cat([g["attn_v.0.bias"], g["attn_u.0.bias"]], dim=0).contiguous()
            esz = 4
            w.w_v, w.b_v = g["_attn_vu.weight"].data_ptr(), g["_attn_vu.bias"].data_ptr()
            w.w_u, w.b_u = w.w_v + int(w.A) * int(w.H) * esz, w.b_v + int(w.A) * esz
            w.w_w, w.b_w = g["attn_w.weight"].data_ptr(), g["attn_w.bias"].data_ptr()
        else:
            w.A = g["attn.0.weight"].shape[0]
            w.w_v, w.b_v = g["attn.0.weight"].data_ptr(), g["attn.0.bias"].data_ptr()
            w.w_u, w.b_u = None, None
            w.w_w, w.b_w = g["attn.2.weight"].data_ptr(), g["attn.2.bias"].data_ptr()
        w.w_cls, w.b_cls = g["classifier.0.weight"].data_ptr(), g["classifier.0.bias"].data_ptr()
        w.missing_prob = float(missing_prob)
        self.w = w
        self.D = int(w.D)
        na = int(w.A) * (2 if gated else 1)
        self.tensor_path_ok = int(w.H) in (64, 128, 256) and na in (64, 128, 256) and self.D % 32 == 0 and int(w.H) % 32 == 0
        if precision not in ("fp32", "tf32"):
            raise ValueError("precision must be 'fp32' or 'tf32'")
        if precision == "tf32" and not self.tensor_path_ok:
            raise ValueError("the tf32 tensor path needs H and (2)A in {64,128,256}, D and H multiples of 32")
        self.precision = precision
        self._ws: Optional[torch.Tensor] = None

    def _workspace(self, n: int, lmax: int, device) -> torch.Tensor:
        need = int(self.lib.pdf_mil_workspace_bytes(C.byref(self.w), n, lmax))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def sweep(self, bags: torch.Tensor, lens: torch.Tensor, live: Optional[torch.Tensor] = None, n_scenarios: int = 1,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bags [n, Lmax, D] f32 cuda (zero padded), lens [n] i32 (0 = no bag), live [S, n] u8 (1 = the bag is present under
        scenario s; None = present everywhere) -> prob [S, n] f32: ONE projection + pooling pass for all scenarios."""
        if bags.dtype != torch.float32 or not bags.is_cuda or not bags.is_contiguous() or bags.shape[2] != self.D:
            raise ValueError("bags must be a contiguous float32 CUDA tensor [n, Lmax, D]")
        n, lmax = int(bags.shape[0]), int(bags.shape[1])
        lens = lens.to(device=bags.device, dtype=torch.int32).contiguous()
        if live is not None:
            live = live.to(device=bags.device, dtype=torch.uint8).contiguous()
            n_scenarios = int(live.shape[0])
            if tuple(live.shape) != (n_scenarios, n):
                raise ValueError("live must be [S, n]")
        ws = self._workspace(n, lmax, bags.device)
        prob = out if out is not None else torch.empty((n_scenarios, n), dtype=torch.float32, device=bags.device)
        prec = _lib.PREC_TF32 if self.precision == "tf32" else _lib.PREC_F32
        _lib.check(self.lib.pdf_mil_sweep(C.byref(self.w), n, lmax, bags.data_ptr(), lens.data_ptr(), n_scenarios, _lib.ptr(live), prec,
                                          ws.data_ptr(), prob.data_ptr(), _lib.stream_ptr()), "pdf_mil_sweep")
        return prob

    def forward(self, bags: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        """bags [n, Lmax, D] f32 cuda (zero padded), lens [n] i32 cuda (0 = missing) -> prob [n] f32."""
        return self.sweep(bags, lens)[0]


def _fill_mlp(m: "_lib.Mlp", weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor]) -> None:
    if len(weights) > _lib.PDF_MAX_LAYERS:
        raise ValueError(f"at most {_lib.PDF_MAX_LAYERS} linear layers are supported")
    m.n_layers = len(weights)
    m.dims[0] = int(weights[0].shape[1])
    for i, (w, b) in enumerate(zip(weights, biases)):
        m.dims[i + 1] = int(w.shape[0])
        m.w[i] = w.data_ptr()
        m.b[i] = b.data_ptr()


class ModDropSweep:
    """ModalityDropoutNet (models/fusion_moddrop.py:8-55) evaluated under S scenario masks at once."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], modality_dims: Dict[str, int], device=None):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        idx = sorted({int(k.split(".")[1]) for k in state_dict if k.startswith("net.") and k.endswith(".weight")})
        self.weights = [_dev_f32(state_dict[f"net.{i}.weight"], self.device) for i in idx]
        self.biases = [_dev_f32(state_dict[f"net.{i}.bias"], self.device) for i in idx]
        self.mods = sorted(modality_dims)                       # fusion_moddrop.py:12 -- sorted(modality) feature layout
        m = _lib.Mlp()
        _fill_mlp(m, self.weights, self.biases)
        m.n_mods = len(self.mods)
        off = 0
        for i, mod in enumerate(self.mods):
            m.mod_off[i] = off
            off += int(modality_dims[mod])
        m.mod_off[len(self.mods)] = off
        if off != int(m.dims[0]):
            raise ValueError("modality_dims do not add up to the network input width")
        self.net = m

    def forward(self, X: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """X [N,F] f32 cuda; masks [S,N,M] u8 cuda in self.mods order -> prob [S,N] f32."""
        N, S = int(X.shape[0]), int(masks.shape[0])
        X = X.to(dtype=torch.float32).contiguous()
        masks = masks.to(dtype=torch.uint8).contiguous()
        ws = torch.empty(self.lib.pdf_moddrop_workspace_bytes(C.byref(self.net), N), dtype=torch.uint8, device=X.device)
        prob = torch.empty((S, N), dtype=torch.float32, device=X.device)
        _lib.check(self.lib.pdf_moddrop_sweep(C.byref(self.net), N, S, X.data_ptr(), masks.data_ptr(), ws.data_ptr(),
                                              prob.data_ptr(), _lib.stream_ptr()), "pdf_moddrop_sweep")
        return prob


class MoeSweep:
    """MoENet (models/moe.py:23-47) evaluated under S scenario masks at once; experts in sorted(modality) order."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], modalities: Sequence[str], device=None):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.mods = sorted(modalities)
        if len(self.mods) > _lib.PDF_MAX_MODS:
            raise ValueError(f"at most {_lib.PDF_MAX_MODS} experts are supported")
        self._keep: List[torch.Tensor] = []
        net = _lib.Moe()
        net.n_experts = len(self.mods)
        for e, mod in enumerate(self.mods):
            idx = sorted({int(k.split(".")[3]) for k in state_dict if k.startswith(f"experts.{mod}.net.") and k.endswith(".weight")})
            ws = [_dev_f32(state_dict[f"experts.{mod}.net.{i}.weight"], self.device) for i in idx]
            bs = [_dev_f32(state_dict[f"experts.{mod}.net.{i}.bias"], self.device) for i in idx]
            self._keep += ws + bs
            _fill_mlp(net.expert[e], ws, bs)
        r = [_dev_f32(state_dict[k], self.device) for k in ("router.0.weight", "router.0.bias", "router.2.weight", "router.2.bias")]
        self._keep += r
        net.router_hidden = int(r[0].shape[0])
        net.w_r0, net.b_r0, net.w_r1, net.b_r1 = (t.data_ptr() for t in r)
        self.net = net

    def forward(self, X: Dict[str, torch.Tensor], masks: torch.Tensor) -> torch.Tensor:
        """X {mod: [N,d] f32 cuda, UNMASKED}; masks [S,N,M] u8 in self.mods order -> prob [S,N] f32."""
        xs = [X[m].to(dtype=torch.float32).contiguous() for m in self.mods]
        N, S = int(xs[0].shape[0]), int(masks.shape[0])
        masks = masks.to(dtype=torch.uint8).contiguous()
        ptrs = (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
        prob = torch.empty((S, N), dtype=torch.float32, device=xs[0].device)
        _lib.check(self.lib.pdf_moe_sweep(C.byref(self.net), N, S, ptrs, masks.data_ptr(), prob.data_ptr(), _lib.stream_ptr()),
                   "pdf_moe_sweep")
        return prob
