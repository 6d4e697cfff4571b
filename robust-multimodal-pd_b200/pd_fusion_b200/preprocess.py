"""Host side of K1: batched volume -> network-input preprocessing on the GPU.

Mirrors, for a whole batch of subjects resident in HBM, the reference's per-subject CPU sequence
`_load_volume` -> `_normalize_volume_for_resnet` -> `_select_slices` -> F.interpolate/normalise
(data/openneuro_features.py:22-32, 121-151, 250-255).  All arithmetic runs in
libpdfusion_b200.so; this class only owns the device buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import torch

from . import _lib


@dataclass
class PreprocResult:
    net_input: torch.Tensor   # [B, L, S, S] bf16, [B, L, rows, pitch] bf16 (zero-padded) or [B, L, S, S, 3] f32
    zoomed: torch.Tensor      # [B, T0, T1, T2] f32 (resampled, un-normalised); [B, T2, T0, T1] from a slice-major run()
    lohi: torch.Tensor        # [B, 4] f32: lo, hi, denominator, has_positive
    indices: torch.Tensor     # [B, L] i32 (-1 in unused slots)
    nslices: torch.Tensor     # [B, n_axes] i32


class VolumePreprocessor:
    def __init__(self, in_shape: Sequence[int], target_shape: Sequence[int] = (160, 160, 160),
                 axes: Sequence[int] = (2,), counts: Sequence[int] = (24,), input_size: int = 224,
                 mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), out_mode: int = _lib.OUT_BF16_C1,
                 max_batch: int = 1, device=None, slice_major: bool = False):
        """slice_major: `run()` (the fused pipeline) keeps the resampled volume slice-axis-major, [B, T2, T0, T1], when the library
        supports it for this configuration (one axis group on axis 2; `pdf_preproc_slice_major_ok`): the selected planes are then
        contiguous and the plane-gather pass disappears.  `PreprocResult.zoomed` of such a run has that layout
        (`self.run_slice_major`); the stage-level entry points (resample / select / gather / gather_slices) always use the
        reference's C order."""
        _lib.require_cuda()
        self.lib = _lib.load()
        if len(axes) != len(counts) or not 1 <= len(axes) <= _lib.PDF_MAX_AXES:
            raise ValueError("slice-counts must match length of slice-axes")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.in_shape = tuple(int(v) for v in in_shape)
        self.target_shape = tuple(int(v) for v in target_shape)
        self.axes, self.counts = [int(a) for a in axes], [int(c) for c in counts]
        self.lmax = sum(self.counts)
        self.input_size = int(input_size)
        self.out_mode = out_mode
        self.max_batch = int(max_batch)
        cfg = _lib.PreprocCfg()
        cfg.in_shape[:] = self.in_shape
        cfg.out_shape[:] = self.target_shape
        cfg.n_axes = len(self.axes)
        for i in range(_lib.PDF_MAX_AXES):
            cfg.axes[i] = self.axes[i] if i < len(self.axes) else 0
            cfg.counts[i] = self.counts[i] if i < len(self.counts) else 0
        cfg.input_size = self.input_size
        cfg.mean[:] = [float(m) for m in mean]
        cfg.std[:] = [float(s) for s in std]
        self.cfg = cfg
        self.cfg_run = cfg                                    # configuration of the fused run(): slice-major where possible
        self.run_slice_major = False
        if slice_major and self.lib.pdf_preproc_slice_major_ok(C.byref(cfg)):
            self.cfg_run = _lib.PreprocCfg.from_buffer_copy(cfg)
            self.cfg_run.slice_major = 1
            self.run_slice_major = True
        ws = self.lib.pdf_preproc_workspace_bytes(C.byref(cfg), self.max_batch)
        if ws == 0:
            raise _lib.PdfusionError("pdf_preproc_workspace_bytes returned 0 (bad configuration)")
        B, S = self.max_batch, self.input_size
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(ws, dtype=torch.uint8, device=self.device)
            self.zoomed = torch.empty((B,) + self.target_shape, dtype=torch.float32, device=self.device)
            self.lohi = torch.empty((B, 4), dtype=torch.float32, device=self.device)
            self.indices = torch.empty((B, self.lmax), dtype=torch.int32, device=self.device)
            self.nslices = torch.empty((B, len(self.axes)), dtype=torch.int32, device=self.device)
            if out_mode == _lib.OUT_BF16_C1:
                self.net_input = torch.empty((B, self.lmax, S, S), dtype=torch.bfloat16, device=self.device)
            elif out_mode == _lib.OUT_BF16_C1_PAD:
                pitch, rows = _lib.stem_padded_dims(S)       # the border is zeroed here, once; the kernels only write the image
                self.net_input = torch.zeros((B, self.lmax, rows, pitch), dtype=torch.bfloat16, device=self.device)
            else:
                self.net_input = torch.empty((B, self.lmax, S, S, 3), dtype=torch.float32, device=self.device)

    def _check_raw(self, raw: torch.Tensor) -> int:
        if not (raw.is_cuda and raw.dtype == torch.float32 and raw.is_contiguous()):
            raise ValueError("raw volumes must be a contiguous float32 CUDA tensor [B, X, Y, Z]")
        if tuple(raw.shape[1:]) != self.in_shape or raw.shape[0] > self.max_batch:
            raise ValueError(f"raw volumes {tuple(raw.shape)} do not match in_shape {self.in_shape} / max_batch {self.max_batch}")
        return int(raw.shape[0])

    def run(self, raw: torch.Tensor, net_input: torch.Tensor | None = None) -> PreprocResult:
        """Enqueues resample+stats, select, gather/resize on the current stream (no sync)."""
        B = self._check_raw(raw)
        out = self.net_input if net_input is None else net_input
        rc = self.lib.pdf_preprocess(C.byref(self.cfg_run), B, raw.data_ptr(), self.zoomed.data_ptr(), self.workspace.data_ptr(),
                                     self.lohi.data_ptr(), self.indices.data_ptr(), self.nslices.data_ptr(), out.data_ptr(),
                                     self.out_mode, _lib.stream_ptr())
        _lib.check(rc, "pdf_preprocess")
        zoomed = self.zoomed[:B]
        if self.run_slice_major:                              # same bytes, [B, T2, T0, T1]
            T0, T1, T2 = self.target_shape
            zoomed = zoomed.view(B, T2, T0, T1)
        return PreprocResult(out[:B], zoomed, self.lohi[:B], self.indices[:B], self.nslices[:B])

    # stage-level entry points (parity tests, per-kernel timing)
    def resample(self, raw: torch.Tensor) -> torch.Tensor:
        B = self._check_raw(raw)
        _lib.check(self.lib.pdf_resample_stats(C.byref(self.cfg), B, raw.data_ptr(), self.zoomed.data_ptr(),
                                               self.workspace.data_ptr(), _lib.stream_ptr()), "pdf_resample_stats")
        return self.zoomed[:B]

    def select(self, B: int):
        _lib.check(self.lib.pdf_select_bounds_indices(C.byref(self.cfg), B, self.zoomed.data_ptr(), self.workspace.data_ptr(),
                                                      self.lohi.data_ptr(), self.indices.data_ptr(), self.nslices.data_ptr(),
                                                      _lib.stream_ptr()), "pdf_select_bounds_indices")
        return self.lohi[:B], self.indices[:B], self.nslices[:B]

    def select_raw(self, vol: torch.Tensor):
        """Indices of `_select_slices` on an already-normalised volume: extent test is voxel > 0 (no re-normalisation)."""
        B = self._check_raw(vol)
        self.resample(vol)                      # in_shape == target_shape: identity zoom, gathers plane maxima
        self.cfg.extent_raw = 1
        try:
            _, idx, ns = self.select(B)
        finally:
            self.cfg.extent_raw = 0
        return idx, ns

    def gather(self, B: int) -> torch.Tensor:
        _lib.check(self.lib.pdf_gather_resize_normalize(C.byref(self.cfg), B, self.zoomed.data_ptr(), self.workspace.data_ptr(),
                                                        self.lohi.data_ptr(), self.indices.data_ptr(), self.nslices.data_ptr(),
                                                        self.net_input.data_ptr(), self.out_mode, _lib.stream_ptr()),
                   "pdf_gather_resize_normalize")
        return self.net_input[:B]

    # ---- test-time augmentation (a6): slices -> affine/intensity/noise/clip -> network input
    def slice_shape(self):
        """(H, W) of the selected slices; every axis group must agree (the reference concatenates them)."""
        shapes = set()
        for a in self.axes:
            dims = list(self.target_shape)
            dims.pop(a)
            shapes.add(tuple(dims))
        if len(shapes) != 1:
            raise ValueError("slice groups of different shapes cannot be concatenated (the reference needs a cubic target_shape)")
        return shapes.pop()

    def gather_slices(self, B: int) -> torch.Tensor:
        """`_select_slices` of the normalised volume, all axis groups concatenated: [B, L, H, W] f32 (after select())."""
        H, W = self.slice_shape()
        if getattr(self, "_slices", None) is None:
            self._slices = torch.empty((self.max_batch, self.lmax, H, W), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.pdf_gather_slices(C.byref(self.cfg), B, self.zoomed.data_ptr(), self.lohi.data_ptr(), self.indices.data_ptr(),
                                              self.nslices.data_ptr(), self._slices.data_ptr(), _lib.stream_ptr()), "pdf_gather_slices")
        return self._slices[:B]

    def tta_augment(self, slices: torch.Tensor, params: torch.Tensor, noise: torch.Tensor | None, affine_only: bool = False) -> torch.Tensor:
        """One augmentation pass.  slices [B, L, H, W] f32; params = pdf_tta_params bytes of the B subjects on the device;
        noise [B, L, H, W] f64 on the device or None."""
        B, L, H, W = (int(v) for v in slices.shape)
        out = torch.empty_like(slices)
        _lib.check(self.lib.pdf_tta_augment(B, L, H, W, slices.data_ptr(), params.data_ptr(), _lib.ptr(noise), int(affine_only),
                                            out.data_ptr(), _lib.stream_ptr()), "pdf_tta_augment")
        return out

    def resize_slices(self, slices: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Bilinear resize + (x-mean)/std of ready-made [0,1] slices into the network-input layout of this preprocessor."""
        B, L, H, W = (int(v) for v in slices.shape)
        out = self.net_input if out is None else out
        mean = (C.c_float * 3)(*[float(v) for v in self.cfg.mean])
        std = (C.c_float * 3)(*[float(v) for v in self.cfg.std])
        _lib.check(self.lib.pdf_resize_slices(B, L, H, W, self.input_size, mean, std, slices.data_ptr(), out.data_ptr(), self.out_mode,
                                              _lib.stream_ptr()), "pdf_resize_slices")
        return out[:B]

    def normalized_volume(self, B: int) -> torch.Tensor:
        """`_normalize_volume_for_resnet` output itself (parity helper)."""
        out = torch.empty_like(self.zoomed[:B])
        vox = self.target_shape[0] * self.target_shape[1] * self.target_shape[2]
        _lib.check(self.lib.pdf_normalize_volume(B, vox, self.zoomed.data_ptr(), self.lohi.data_ptr(), out.data_ptr(),
                                                 _lib.stream_ptr()), "pdf_normalize_volume")
        return out

    def algorithmic_bytes(self) -> int:
        """4XYZ + 2*4*T^3 + 4*L*T^2 + b_out*L*I^2 per subject."""
        X, Y, Z = self.in_shape
        T0, T1, T2 = self.target_shape
        plane = 0
        for a, c in zip(self.axes, self.counts):
            dims = [T0, T1, T2]
            dims.pop(a)
            plane += 4 * c * dims[0] * dims[1]
        b_out = 12 if self.out_mode == _lib.OUT_F32_NHWC3 else 2
        return 4 * X * Y * Z + 8 * T0 * T1 * T2 + plane + b_out * self.lmax * self.input_size ** 2
