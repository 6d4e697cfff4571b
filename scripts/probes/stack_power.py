"""GPU probe: SM clock / power while the conv stack runs back to back for ~1.5 s (NVML sampled every 5 ms)."""
import sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
import pynvml
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
lib = _lib.load()
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D("resnet18").state_dict().items() if not k.startswith("fc.")}
n = 768
enc = ResNetEncoder(sd, n, 224, precision="bf16")
enc.input.copy_((torch.rand(n, 224, 224, device="cuda") * 2 - 1).to(torch.bfloat16))
enc.forward(None)
torch.cuda.synchronize()
samples, stop = [], threading.Event()
def sampler():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.005)
t = threading.Thread(target=sampler); t.start()
time.sleep(0.05)
for first, count, label in [(0, enc.n_ops, "whole stack"), (12, 8, "layers 3-4 only"), (1, 4, "layer 1 only"), (0, 1, "stem only")]:
    torch.cuda.synchronize(); time.sleep(0.3)
    i0 = len(samples)
    evs = []
    reps = 400 if count == enc.n_ops else 1200
    for blk in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps // 4):
            enc.run_range(first, count)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    seg = samples[i0:]
    clk = sorted(s[1] for s in seg); pw = sorted(s[2] for s in seg)
    reasons = 0
    for s in seg: reasons |= s[3]
    print(f"{label:16s} per pass by quarter (us): " + " ".join(f"{a.elapsed_time(b) / (reps // 4) * 1e3:7.1f}" for a, b in evs) +
          f" | clk median {clk[len(clk)//2]} min {clk[0]} MHz, power median {pw[len(pw)//2]:.0f} max {pw[-1]:.0f} W, reasons 0x{reasons:x} ({len(seg)} samples)")
stop.set(); t.join()
