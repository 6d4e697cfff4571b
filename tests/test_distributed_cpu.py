"""N>1 path on the CPU: world_size-2 gloo processes shard rows and all-gather the table in manifest order."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, sys.argv[1])
    import torch
    from pd_fusion_b200.parallel import init_distributed, shard_range, all_gather_rows, barrier
    rank, local_rank, ws = init_distributed(backend="gloo")
    n = int(sys.argv[2])
    lo, hi = shard_range(n, rank, ws)
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1).repeat(1, 3) * 10.0
    full = all_gather_rows(local, n)
    want = torch.arange(n, dtype=torch.float32).view(-1, 1).repeat(1, 3) * 10.0
    assert full.shape == want.shape and torch.equal(full, want), (rank, full)
    barrier()
    if rank == 0:
        print("GATHER_OK", n, ws)
    torch.distributed.destroy_process_group()
""")


def _run(n_rows: int, ws: int, port: int, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(ws):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(ws), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT / "robust-multimodal-pd_b200"), str(n_rows)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert any("GATHER_OK" in o for o in outs)


def test_all_gather_rows_world2(tmp_path):
    _run(7, 2, 29611, tmp_path)      # ragged: 4 + 3 rows


def test_all_gather_rows_world2_fewer_rows_than_ranks(tmp_path):
    _run(1, 2, 29612, tmp_path)      # rank 1 holds an empty shard
