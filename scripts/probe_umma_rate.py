"""GPU probe: cycles per tcgen05.mma (M=128, N, K=16, both operands in shared memory) when issued back to back."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib

lib = _lib.load()
out = torch.zeros(296, dtype=torch.int64, device="cuda")
for grid in (1, 148):
    for N in (64, 128, 256):
        for mode in (0, 1, 2, 3):          # bit 0: descriptor walks the 9 halo windows; bit 1: CTA pairs (cta_group::2, M = 256)
            iters = 200
            _lib.check(lib.pdf_selftest_umma_rate(N, iters, mode, grid, out.data_ptr(), _lib.stream_ptr()))
            torch.cuda.synchronize()
            cyc = out[:max(1, grid // 2) if mode & 2 else grid].float()
            per = cyc / (iters * 36)
            print(f"grid {grid:3d} N {N:3d} mode {mode}: cycles/MMA mean {per.mean().item():6.1f} max {per.max().item():6.1f}  (tensor floor {128 * N // 256})")
