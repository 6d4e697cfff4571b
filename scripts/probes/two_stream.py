"""GPU probe: does running two half-batches on two streams, each kernel capped to half of the SMs, overlap the HBM-bound 1x1
launches of one half with the tensor-bound 3x3 launches of the other?  usage: two_stream.py [arch] [n_slices]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
lib = _lib.load()
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D(arch).state_dict().items() if not k.startswith("fc.")}
full = ResNetEncoder(sd, n, 224, precision="bf16", arch=arch)
halves = [ResNetEncoder(sd, n // 2, 224, precision="bf16", arch=arch) for _ in range(2)]
for e in [full] + halves:
    e.input.copy_((torch.rand(e.n, 224, 224, device="cuda") * 2 - 1).to(torch.bfloat16))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def t_full(reps=10):
    full.forward(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        full.forward(None)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def t_split(cap, reps=10, interleave=True):
    lib.pdf_debug_set_sm_cap(cap)
    cur = torch.cuda.current_stream()
    def go():
        for s in streams:
            s.wait_stream(cur)
        if interleave:   # op by op, alternating streams, stream 1 one op behind
            for i in range(halves[0].n_ops + 1):
                if i < halves[0].n_ops:
                    with torch.cuda.stream(streams[0]):
                        halves[0].run_range(i, 1)
                if i >= 1:
                    with torch.cuda.stream(streams[1]):
                        halves[1].run_range(i - 1, 1)
        else:
            for h, s in zip(halves, streams):
                with torch.cuda.stream(s):
                    h.forward(None)
        for s in streams:
            cur.wait_stream(s)
    go()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    lib.pdf_debug_set_sm_cap(0)
    return e0.elapsed_time(e1) / reps


print(f"{arch} n={n}: one stream, all SMs: {t_full():.3f} ms")
for pdl in (1, 0):
    lib.pdf_debug_enable_pdl(pdl)
    print(f"PDL={pdl}: one stream {t_full():.3f} ms")
    for cap in (74, 80, 96, 148):
        print(f"  two half-batches on two streams, grids capped to {cap} SMs: {t_split(cap):.3f} ms (interleaved issue), {t_split(cap, interleave=False):.3f} ms (stream by stream)")
