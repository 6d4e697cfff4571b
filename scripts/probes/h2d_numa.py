"""8-rank probe (torchrun): aggregate pinned-host -> device bandwidth of all ranks copying at once, as allocated by default and
after binding each rank's CPUs / memory to the NUMA node of its GPU (NVML affinity, libnuma if present).  Prints the topology facts
the binding depends on.  usage: torchrun --nproc-per-node 8 scripts/probes/h2d_numa.py"""
import ctypes
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, lr, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return f"<{e}>"


if rank == 0:
    print("allowed cpus", sorted(os.sched_getaffinity(0)))
    print(sh("nvidia-smi topo -m"))
    print(sh("lscpu | grep -i -E 'numa|socket|^CPU\\(s\\)'"))
    print("cpuset.cpus", sh("cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null || cat /sys/fs/cgroup/cpuset/cpuset.cpus"))
    print("cpuset.mems", sh("cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null || cat /sys/fs/cgroup/cpuset/cpuset.mems"))
    print(sh("numactl -H 2>&1 | head -12"))


def measure(tag, buf):
    d = torch.empty(buf.shape, dtype=buf.dtype, device="cuda")
    for _ in range(2):
        d.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    if ws > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        d.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([10 * buf.numel() * buf.element_size() / dt / 1e9], device="cuda")
    if ws > 1:
        allg = [torch.zeros_like(gbs) for _ in range(ws)]
        dist.all_gather(allg, gbs)
        if rank == 0:
            v = [float(x) for x in allg]
            print(f"{tag}: per rank {[round(x, 1) for x in v]} GB/s, sum {sum(v):.1f} GB/s")
    elif rank == 0:
        print(f"{tag}: {float(gbs):.1f} GB/s")
    del d


N = 1 << 29
measure("default placement", torch.empty(N, dtype=torch.int16).pin_memory())

# NUMA node of this rank's GPU
node = -1
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    try:
        node = pynvml.nvmlDeviceGetNumaNodeId(h)
    except Exception:
        node = -1
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 4)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
    ok = sorted(set(cpus) & os.sched_getaffinity(0))
    print(f"rank {rank}: gpu numa node {node}, nvml affinity {cpus[:4]}..{cpus[-1:] if cpus else ''} ({len(cpus)} cpus), usable {len(ok)}", flush=True)
    if ok:
        os.sched_setaffinity(0, ok)
except Exception as e:
    print(f"rank {rank}: nvml affinity failed: {e}", flush=True)
measure("cpu affinity -> GPU's node (first touch)", torch.empty(N, dtype=torch.int16).pin_memory())

try:
    numa = ctypes.CDLL("libnuma.so.1")
    if numa.numa_available() >= 0 and node >= 0:
        numa.numa_set_preferred(node)
        measure("libnuma preferred node", torch.empty(N, dtype=torch.int16).pin_memory())
    elif rank == 0:
        print("libnuma: not available or node unknown")
except OSError as e:
    if rank == 0:
        print("libnuma missing:", e)
if ws > 1:
    dist.destroy_process_group()
