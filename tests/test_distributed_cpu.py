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


TRAIN_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, sys.argv[1])
    import torch
    from pd_fusion_b200.parallel import init_distributed, barrier
    from pd_fusion_b200.training import BucketedAllReduce, _flat_like, allreduce_mean, bucket_ranges, flatten_parameters
    rank, local_rank, ws = init_distributed(backend="gloo")
    torch.manual_seed(0)                                     # same initial weights on every rank, as under DDP
    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.ReLU(), torch.nn.Linear(3, 1))
    before = {k: v.clone() for k, v in net.state_dict().items()}
    params = dict(net.named_parameters())
    flat_param = flatten_parameters(params)                  # every parameter becomes a view into ONE buffer
    flat_grad, grads = _flat_like({k: v.data for k, v in params.items()})
    assert flat_param.numel() == flat_grad.numel()
    for k, v in net.state_dict().items():                    # the module still sees the same values ...
        assert torch.equal(v, before[k]), k
    flat_param.mul_(2.0)                                     # ... and an in-place update of the flat buffer IS the parameter update
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k] * 2.0), k
    assert net(torch.ones(2, 5)).shape == (2, 1)
    for i, g in enumerate(grads.values()):                   # rank-dependent gradients: the data-parallel step averages them
        g.fill_(float(rank + 1) * (i + 1))
    allreduce_mean([flat_grad])
    for i, g in enumerate(grads.values()):
        assert torch.allclose(g, torch.full_like(g, 1.5 * (i + 1))), (rank, i, g)
    # bucketed form (what the fine-tune backward does block by block, last block first): same result, several collectives
    names = {"conv1.weight": (4, 3), "bn1.weight": (4,), "layer1.0.conv1.weight": (6, 4), "layer1.0.bn1.bias": (6,), "layer1.1.conv1.weight": (5, 6)}
    flat2, g2 = _flat_like({k: torch.zeros(v) for k, v in names.items()})
    buckets = bucket_ranges(_flat_like.last_offsets, {k: int(torch.zeros(v).numel()) for k, v in names.items()})
    assert list(buckets) == ["stem", "layer1.0", "layer1.1"] and buckets["stem"][0] == 0 and buckets["layer1.1"][1] == flat2.numel()
    for i, g in enumerate(g2.values()):
        g.fill_(float(rank) * 2.0 + i)
    red = BucketedAllReduce(flat2)
    for b in reversed(list(buckets)):
        red.ready(*buckets[b])
    red.finish()
    assert red.calls == 3
    for i, g in enumerate(g2.values()):
        assert torch.allclose(g, torch.full_like(g, 1.0 + i)), (rank, i, g)
    barrier()
    if rank == 0:
        print("ALLREDUCE_OK", ws)
    torch.distributed.destroy_process_group()
""")


def test_flat_gradient_allreduce_world2(tmp_path):
    """Config 5's data-parallel step: parameters flattened into one buffer per module, one all-reduce (mean) per
    flat gradient buffer -- two gloo ranks with different gradients end up with their average."""
    script = tmp_path / "train_worker.py"
    script.write_text(TRAIN_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT / "robust-multimodal-pd_b200")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert any("ALLREDUCE_OK" in o for o in outs)
