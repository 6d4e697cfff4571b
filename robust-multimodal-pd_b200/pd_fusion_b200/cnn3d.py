"""f4, the `cnn3d` feature mode: the 3-D convolutional auto-encoder of the reference's scripts/build_cnn3d_embeddings.py
(`Simple3DAE` :56-88, `load_volume` :28-41, training loop :122-142, embedding extraction :144-152), trained and evaluated on the
device by hand-written kernels -- no autograd, no cuDNN.

Layout is channels-last [N, D, H, W, C] f32.  A 3x3x3 convolution is DEPTH-DECOMPOSED into three 2-D 3x3 convolutions over
depth-shifted slice ranges, out[:, d] = sum_kd conv2d(x[:, d + kd - 1], W[:, :, kd]), each running on the FP32 implicit-GEMM kernels
the ResNet parity path already has (forward with in-place residual accumulation, `pdf_conv_dgrad_f32`, `pdf_conv_wgrad_f32`); a
2x2x2 stride-2 transposed convolution is `pdf_gemm_f32` on [voxels, Cin] x [Cin, 8*Cout] plus a pixel shuffle (csrc/cnn3d.cu).  The
`nn.Module` below exists for its parameter names and initialisation draws (the `state_dict` is the interchange format and the same
`torch.manual_seed` must give the reference's initial weights); its torch `forward` is never on the product path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .training import NativeAdam


class Simple3DAE(nn.Module):
    """Same layers, names and construction order as the reference's Simple3DAE (scripts/build_cnn3d_embeddings.py:56-88)."""

    def __init__(self, input_shape=(96, 96, 96), embedding_dim=128):
        super().__init__()
        self.encoder = nn.Sequential(
            nn.Conv3d(1, 8, 3, padding=1), nn.ReLU(), nn.MaxPool3d(2),
            nn.Conv3d(8, 16, 3, padding=1), nn.ReLU(), nn.MaxPool3d(2),
            nn.Conv3d(16, 32, 3, padding=1), nn.ReLU(), nn.MaxPool3d(2),
        )
        if any(int(v) % 8 for v in input_shape):
            raise ValueError("Simple3DAE: every extent of input_shape must be a multiple of 8 (three 2x poolings, mirrored by the decoder)")
        self._enc_shape = (32,) + tuple(int(v) // 8 for v in input_shape)
        enc_dim = int(np.prod(self._enc_shape))
        self.fc = nn.Linear(enc_dim, embedding_dim)
        self.fc_dec = nn.Linear(embedding_dim, enc_dim)
        self.decoder = nn.Sequential(
            nn.ConvTranspose3d(32, 16, 2, stride=2), nn.ReLU(),
            nn.ConvTranspose3d(16, 8, 2, stride=2), nn.ReLU(),
            nn.ConvTranspose3d(8, 1, 2, stride=2),
        )

    def forward(self, x):                                   # torch restatement (tests' autograd yardstick); not the product path
        z = self.encoder(x)
        emb = self.fc(z.view(z.size(0), -1))
        out = self.decoder(self.fc_dec(emb).view(z.size(0), *self._enc_shape))
        return out, emb


def standardize_volumes(vols: torch.Tensor) -> torch.Tensor:
    """`load_volume`'s per-volume normalisation (scripts/build_cnn3d_embeddings.py:35-40): (v - mean) / (std + 1e-6) with the mean /
    std of the positive voxels; volumes without one pass through.  vols [B, D, H, W] f32 on the device."""
    lib = _lib.load()
    B = int(vols.shape[0])
    vox = int(vols[0].numel())
    stride = int(lib.pdf_simple_stats_stride())
    stats = torch.empty((B, stride), dtype=torch.float64, device=vols.device)
    _lib.check(lib.pdf_simple_stats(B, vox, 1, vols.data_ptr(), stats.data_ptr(), _lib.stream_ptr()), "pdf_simple_stats")
    out = torch.empty_like(vols)
    _lib.check(lib.pdf_standardize_volume(B, vox, vols.data_ptr(), stats.data_ptr(), stride, out.data_ptr(), _lib.stream_ptr()),
               "pdf_standardize_volume")
    return out


class Cnn3dTrainer:
    """Forward, backward and Adam step of Simple3DAE on the native kernels; parameters live in (and are updated inside) the nn.Module."""

    def __init__(self, model: Simple3DAE, lr: float = 1e-3):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.m = model
        self.p = dict(model.named_parameters())
        self.dev = self.p["fc.weight"].device
        self.enc_c, self.enc_sp = model._enc_shape[0], tuple(model._enc_shape[1:])
        self.E = int(self.p["fc.weight"].shape[0])
        self.convs = [("encoder.0", 1, 8), ("encoder.3", 8, 16), ("encoder.6", 16, 32)]
        self.deconvs = [("decoder.0", 32, 16, True), ("decoder.2", 16, 8, True), ("decoder.4", 8, 1, False)]
        self.g = {k: torch.zeros_like(v.data) for k, v in self.p.items()}
        self.opt = NativeAdam([(list(v.data for v in self.p.values()), float(lr))])
        self.lr = float(lr)
        self.sync_weights()

    # -- re-laid-out weight copies (layout changes only) -------------------------------------------------------------------------
    def sync_weights(self):
        w: Dict[str, object] = {}
        for name, cin, cout in self.convs:
            W = self.p[name + ".weight"].data                                   # [Cout, Cin, 3, 3, 3]
            w[name] = [W[:, :, kd].permute(2, 3, 1, 0).contiguous() for kd in range(3)]          # [kh][kw][Cin][Cout] per depth tap
        S3 = int(np.prod(self.enc_sp))
        Cc = self.enc_c
        w["fc"] = self.p["fc.weight"].data.view(self.E, Cc, S3).permute(0, 2, 1).reshape(self.E, S3 * Cc).contiguous()
        w["fc_dec"] = self.p["fc_dec.weight"].data.view(Cc, S3, self.E).permute(1, 0, 2).reshape(S3 * Cc, self.E).contiguous()
        w["fc_dec.b"] = self.p["fc_dec.bias"].data.view(Cc, S3).t().reshape(-1).contiguous()
        for name, cin, cout, _ in self.deconvs:
            w[name] = self.p[name + ".weight"].data.permute(0, 2, 3, 4, 1).reshape(cin, 8 * cout).contiguous()
        self.w = w

    # -- primitives --------------------------------------------------------------------------------------------------------------
    def _gemm(self, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, Cm, c_rs, bias=None, act=0, acc=0):
        _lib.check(self.lib.pdf_gemm_f32(M, N, K, A.data_ptr(), a_rs, a_cs, B.data_ptr(), b_rs, b_cs, Cm.data_ptr(), c_rs,
                                         _lib.ptr(bias), act, acc, _lib.stream_ptr()), "pdf_gemm_f32")

    def _slabs(self, N: int, D: int):
        """(depth tap, images, input slice offset, output slice offset): kd = 1 covers every slice in one pass; kd = 0 / 2 cover D-1
        slices of each volume (slice d reads slice d-1 / d+1 of the SAME volume)."""
        yield 1, N * D, 0, 0
        for n in range(N):
            yield 0, D - 1, n * D, n * D + 1
            yield 2, D - 1, n * D + 1, n * D

    def _op(self, images, H, W, cin, cout) -> "_lib.Op":
        op = _lib.Op()
        op.kind, op.precision = _lib.OP_CONV, _lib.PREC_F32
        op.n, op.h, op.w, op.c, op.k, op.r, op.s, op.stride, op.pad, op.ho, op.wo, op.relu = images, H, W, cin, cout, 3, 3, 1, 1, H, W, 0
        return op

    def _conv3d(self, x: torch.Tensor, name: str, cin: int, cout: int) -> torch.Tensor:
        N, D, H, W, _ = x.shape
        y = torch.empty((N, D, H, W, cout), dtype=torch.float32, device=self.dev)
        bias = self.p[name + ".bias"].data
        for kd, images, so_in, so_out in self._slabs(N, D):
            if images <= 0:
                continue
            op = self._op(images, H, W, cin, cout)
            op.d_in = x.data_ptr() + so_in * H * W * cin * 4
            op.d_out = y.data_ptr() + so_out * H * W * cout * 4
            op.d_weight = self.w[name][kd].data_ptr()
            if kd == 1:
                op.d_bias = bias.data_ptr()
            else:
                op.d_residual = op.d_out                       # accumulate in place on the centre tap's result
            plan = C.c_void_p()
            _lib.check(self.lib.pdf_plan_create(C.byref(plan), (_lib.Op * 1)(op), 1), "pdf_plan_create")
            _lib.check(self.lib.pdf_plan_run(plan, _lib.stream_ptr()), "pdf_plan_run")
            self.lib.pdf_plan_destroy(plan)
        _lib.check(self.lib.pdf_relu_f32(y.data_ptr(), y.numel(), _lib.stream_ptr()), "pdf_relu_f32")
        return y

    def _conv3d_backward(self, x, y, dy, name, cin, cout, need_dx: bool):
        """dy: gradient w.r.t. the ReLU output (masked here, in place).  Accumulates weight / bias gradients; returns dx or None."""
        N, D, H, W, _ = x.shape
        s = _lib.stream_ptr()
        _lib.check(self.lib.pdf_relu_mask_backward(dy.data_ptr(), y.data_ptr(), None, dy.numel(), s), "pdf_relu_mask_backward")
        gb = torch.zeros(cout, dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.pdf_colsum_f32(N * D * H * W, cout, dy.data_ptr(), gb.data_ptr(), 0, s), "pdf_colsum_f32")
        self.g[name + ".bias"].copy_(gb)
        gw = [torch.zeros_like(self.w[name][kd]) for kd in range(3)]
        dx = torch.empty_like(x) if need_dx else None
        for kd, images, so_in, so_out in self._slabs(N, D):
            if images <= 0:
                continue
            op = self._op(images, H, W, cin, cout)
            xp = x.data_ptr() + so_in * H * W * cin * 4
            dyp = dy.data_ptr() + so_out * H * W * cout * 4
            _lib.check(self.lib.pdf_conv_wgrad_f32(C.byref(op), xp, dyp, gw[kd].data_ptr(), s), "pdf_conv_wgrad_f32")
            if need_dx:
                dxp = dx.data_ptr() + so_in * H * W * cin * 4
                _lib.check(self.lib.pdf_conv_dgrad_f32(C.byref(op), dyp, self.w[name][kd].data_ptr(), dxp, 0 if kd == 1 else 1, s),
                           "pdf_conv_dgrad_f32")
        G = self.g[name + ".weight"]
        for kd in range(3):
            G[:, :, kd].copy_(gw[kd].permute(3, 2, 0, 1))
        return dx

    def _pool(self, x):
        N, D, H, W, Cc = x.shape
        y = torch.empty((N, D // 2, H // 2, W // 2, Cc), dtype=torch.float32, device=self.dev)
        idx = torch.empty(y.shape, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.pdf_maxpool3d_forward(N, D, H, W, Cc, x.data_ptr(), y.data_ptr(), idx.data_ptr(), _lib.stream_ptr()),
                   "pdf_maxpool3d_forward")
        return y, idx

    def _pool_backward(self, shape, idx, dy):
        N, D, H, W, Cc = shape
        dx = torch.empty(shape, dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.pdf_maxpool3d_backward(N, D, H, W, Cc, idx.data_ptr(), dy.data_ptr(), dx.data_ptr(), _lib.stream_ptr()),
                   "pdf_maxpool3d_backward")
        return dx

    # -- network -----------------------------------------------------------------------------------------------------------------
    def encode(self, x: torch.Tensor, tape: List | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """x [B, D, H, W] f32 -> (z [B, S3*C] channels-last flattened, emb [B, E])."""
        t = x.contiguous().unsqueeze(-1)
        for name, cin, cout in self.convs:
            a = self._conv3d(t, name, cin, cout)
            p, idx = self._pool(a)
            if tape is not None:
                tape.append((name, cin, cout, t, a, idx))
            t = p
        B = int(x.shape[0])
        z = t.reshape(B, -1)
        enc = int(z.shape[1])
        emb = torch.empty((B, self.E), dtype=torch.float32, device=self.dev)
        self._gemm(B, self.E, enc, z, enc, 1, self.w["fc"], 1, enc, emb, self.E, bias=self.p["fc.bias"].data)
        return z, emb

    def embed(self, x: torch.Tensor) -> torch.Tensor:
        return self.encode(x)[1]

    def train_step(self, x: torch.Tensor) -> torch.Tensor:
        """One optimisation step on a batch of standardised volumes [B, D, H, W] (reference: scripts/build_cnn3d_embeddings.py:134-141);
        returns the MSE loss as a device tensor."""
        loss = self.forward_backward(x)
        self.opt.step([(self.p[k].data, self.g[k], self.lr) for k in self.p])
        self.sync_weights()
        return loss

    def forward_backward(self, x: torch.Tensor) -> torch.Tensor:
        """recon = decoder(fc_dec(fc(encoder(x)))), MSE(recon, x) and every parameter gradient (into self.g, the nn.Module's layouts)."""
        s = _lib.stream_ptr()
        lib = self.lib
        B = int(x.shape[0])
        x = x.contiguous()
        tape: List = []
        z, emb = self.encode(x, tape)
        enc = int(z.shape[1])
        r = torch.empty((B, enc), dtype=torch.float32, device=self.dev)
        self._gemm(B, enc, self.E, emb, self.E, 1, self.w["fc_dec"], 1, self.E, r, enc, bias=self.w["fc_dec.b"])
        t = r.view((B,) + self.enc_sp + (self.enc_c,))
        dtape = []
        for name, cin, cout, relu in self.deconvs:
            N, D, H, W, _ = t.shape
            V = N * D * H * W
            tm = torch.empty((V, 8 * cout), dtype=torch.float32, device=self.dev)
            self._gemm(V, 8 * cout, cin, t, cin, 1, self.w[name], 8 * cout, 1, tm, 8 * cout)
            y = torch.empty((N, 2 * D, 2 * H, 2 * W, cout), dtype=torch.float32, device=self.dev)
            _lib.check(lib.pdf_shuffle2_3d(N, D, H, W, cout, tm.data_ptr(), self.p[name + ".bias"].data.data_ptr(), 1 if relu else 0, y.data_ptr(), s),
                       "pdf_shuffle2_3d")
            dtape.append((name, cin, cout, relu, t, y))
            t = y
        recon = t                                              # [B, D, H, W, 1]
        loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        dy = torch.empty_like(recon)
        _lib.check(lib.pdf_mse_train(recon.numel(), recon.data_ptr(), x.data_ptr(), loss.data_ptr(), dy.data_ptr(), s), "pdf_mse_train")
        # ---- backward: decoder
        for name, cin, cout, relu, tin, y in reversed(dtape):
            N, D, H, W, _ = tin.shape
            V = N * D * H * W
            dt = torch.empty((V, 8 * cout), dtype=torch.float32, device=self.dev)
            _lib.check(lib.pdf_unshuffle2_3d(N, D, H, W, cout, dy.data_ptr(), y.data_ptr(), 1 if relu else 0, dt.data_ptr(), s), "pdf_unshuffle2_3d")
            gb = torch.empty(cout, dtype=torch.float32, device=self.dev)
            _lib.check(lib.pdf_colsum_f32(V * 8, cout, dt.data_ptr(), gb.data_ptr(), 0, s), "pdf_colsum_f32")      # rows = (voxel, ijk)
            self.g[name + ".bias"].copy_(gb)
            gwm = torch.empty((cin, 8 * cout), dtype=torch.float32, device=self.dev)
            self._gemm(cin, 8 * cout, V, tin, 1, cin, dt, 8 * cout, 1, gwm, 8 * cout)
            self.g[name + ".weight"].copy_(gwm.view(cin, 2, 2, 2, cout).permute(0, 4, 1, 2, 3))
            dx = torch.empty((N, D, H, W, cin), dtype=torch.float32, device=self.dev)
            self._gemm(V, cin, 8 * cout, dt, 8 * cout, 1, self.w[name], 1, 8 * cout, dx, cin)
            dy = dx
        # ---- the two linear layers (channels-last column order; gradients mapped back to the reference's order)
        dr = dy.reshape(B, enc)
        S3, Cc = int(np.prod(self.enc_sp)), self.enc_c
        gwd = torch.empty((enc, self.E), dtype=torch.float32, device=self.dev)
        self._gemm(enc, self.E, B, dr, 1, enc, emb, self.E, 1, gwd, self.E)
        self.g["fc_dec.weight"].copy_(gwd.view(S3, Cc, self.E).permute(1, 0, 2).reshape(enc, self.E))
        gbd = torch.empty(enc, dtype=torch.float32, device=self.dev)
        _lib.check(lib.pdf_colsum_f32(B, enc, dr.data_ptr(), gbd.data_ptr(), 0, s), "pdf_colsum_f32")
        self.g["fc_dec.bias"].copy_(gbd.view(S3, Cc).t().reshape(-1))
        demb = torch.empty((B, self.E), dtype=torch.float32, device=self.dev)
        self._gemm(B, self.E, enc, dr, enc, 1, self.w["fc_dec"], self.E, 1, demb, self.E)
        gwf = torch.empty((self.E, enc), dtype=torch.float32, device=self.dev)
        self._gemm(self.E, enc, B, demb, 1, self.E, z, enc, 1, gwf, enc)
        self.g["fc.weight"].copy_(gwf.view(self.E, S3, Cc).permute(0, 2, 1).reshape(self.E, enc))
        gbf = torch.empty(self.E, dtype=torch.float32, device=self.dev)
        _lib.check(lib.pdf_colsum_f32(B, self.E, demb.data_ptr(), gbf.data_ptr(), 0, s), "pdf_colsum_f32")
        self.g["fc.bias"].copy_(gbf)
        dz = torch.empty((B, enc), dtype=torch.float32, device=self.dev)
        self._gemm(B, enc, self.E, demb, self.E, 1, self.w["fc"], enc, 1, dz, enc)
        # ---- encoder
        dp = dz.view((B,) + self.enc_sp + (self.enc_c,))
        for li, (name, cin, cout, tin, a, idx) in enumerate(reversed(tape)):
            da = self._pool_backward(tuple(a.shape), idx, dp)
            dp = self._conv3d_backward(tin, a, da, name, cin, cout, need_dx=li < len(tape) - 1)
        return loss


def train_autoencoder(model: Simple3DAE, vols: torch.Tensor, epochs: int, batch_size: int, lr: float, log=print) -> Cnn3dTrainer:
    """The reference's loop (scripts/build_cnn3d_embeddings.py:122-142): shuffled mini-batches from a torch DataLoader (the batch
    order comes from the global torch RNG exactly as there -- the DataLoader here iterates over row INDICES, the standardised volumes
    stay resident on the device), MSE reconstruction loss, Adam."""
    from torch.utils.data import DataLoader
    tr = Cnn3dTrainer(model, lr)
    loader = DataLoader(range(int(vols.shape[0])), batch_size=int(batch_size), shuffle=True)
    loss = None
    for epoch in range(int(epochs)):
        for idx in loader:
            loss = tr.train_step(vols[idx.to(vols.device)])
        if loss is not None:
            log(f"epoch {epoch + 1}/{epochs} loss={float(loss.item()):.4f}")
    return tr
