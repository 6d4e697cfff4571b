"""GPU probe: device time of the three preprocessing stages in isolation (C2 shapes, 32 subjects; CUDA events, best of 5 x 10).
PDFUSION_B200_LIB selects the library build."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import numpy as np
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.preprocess import VolumePreprocessor
from pd_fusion_b200.synthetic import synthetic_volume

B = 32
shape, target = (256, 256, 176), (160, 160, 160)
pre = VolumePreprocessor(shape, target, [2], [24], 224, out_mode=_lib.OUT_BF16_C1_PAD, max_batch=B)
pool = [torch.from_numpy(synthetic_volume(i, shape, 1e-5)) for i in range(4)]
raw = torch.stack([pool[i % 4] for i in range(B)]).cuda()
pre.run(raw); torch.cuda.synchronize()

def timeit(fn, reps=10):
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best

# the select stage accumulates into per-subject state that the resample stage resets: stages are timed cumulatively
t_res = timeit(lambda: pre.resample(raw))
t_rs = timeit(lambda: (pre.resample(raw), pre.select(B)))
t_all = timeit(lambda: pre.run(raw))
print(f"resample+stats {t_res:.1f} us | select (percentiles, extents, indices) {t_rs - t_res:.1f} us | gather+resize {t_all - t_rs:.1f} us | whole {t_all:.1f} us")
