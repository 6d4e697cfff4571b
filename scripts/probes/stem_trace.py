"""GPU probe: per-role timeline of CTA 0 of the fused stem kernel (cycles relative to the first stamp).
Needs a library built with the stamps compiled in: make -C robust-multimodal-pd_b200/csrc -B build/stem_tc.o EXTRA=-DPDF_STEM_TRACE && make -C ..."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "robust-multimodal-pd_b200"))
import torch
from pd_fusion_b200 import _lib
from pd_fusion_b200.backbone import ResNet2D, ResNetEncoder

lib = _lib.load()
torch.manual_seed(1234)
sd = {k: v for k, v in ResNet2D("resnet18").state_dict().items() if not k.startswith("fc.")}
n = 768
enc = ResNetEncoder(sd, n, 224, precision="bf16")
enc.input.copy_((torch.rand(n, 224, 224, device="cuda") * 2 - 1).to(torch.bfloat16))
for _ in range(2):
    enc.run_range(0, 1)
torch.cuda.synchronize()
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.pdf_debug_set_trace(buf.data_ptr())
enc.run_range(0, 1)
torch.cuda.synchronize()
lib.pdf_debug_set_trace(None)
t = buf.cpu().view(64, 16)
t0 = int(t[t > 0].min())
names = ["b:pfull", "b:bempty", "b:done", "m:accempty", "m:bfull", "m:commit", "e:accfull", "e:released", "e:done"]
print("tile " + " ".join(f"{n_:>10s}" for n_ in names))
for it in range(4, 30):
    print(f"{it:4d} " + " ".join(f"{int(t[it, e]) - t0:10d}" for e in range(9)))
d = (t[5:40, :9] - t[4:39, :9]).float().mean(dim=0)
print("mean period per event:", [round(float(x)) for x in d])
print("builder work (bempty->done):", float((t[4:40, 2] - t[4:40, 1]).float().mean()), " mma (bfull->commit):", float((t[4:40, 5] - t[4:40, 4]).float().mean()),
      " epilogue (accfull->done):", float((t[4:40, 8] - t[4:40, 6]).float().mean()), " commit->accfull seen:", float((t[4:40, 6] - t[4:40, 5]).float().mean()))

# inside the epilogue (warp 0): accfull -> loads of row group i landed (ld0..ld3), end of group 0's arithmetic and stores (g0done)
print("tile   ld0-accfull  g0done-ld0  ld1-g0done  ld2-ld1  ld3-ld2  done-ld3")
for it in range(4, 20):
    r = [int(x) for x in t[it]]
    print(f"{it:4d} {r[9]-r[6]:10d} {r[13]-r[9]:10d} {r[10]-r[13]:10d} {r[11]-r[10]:10d} {r[12]-r[11]:10d} {r[8]-r[12]:10d}")
