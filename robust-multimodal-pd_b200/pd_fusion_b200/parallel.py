"""Subject sharding across the GPUs of one box: one process per GPU, contiguous blocks of
manifest rows per rank, no data-path collective until the embedding table / per-scenario probabilities are
assembled with an NCCL all-gather over NVLink.  The reference has no distributed code at all
(every builder is a serial Python loop: data/openneuro_features.py:226).
"""
from __future__ import annotations

import datetime
import os
from typing import Tuple

import torch


def world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process if absent)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend: str | None = None) -> Tuple[int, int, int]:
    rank, local_rank, ws = world()
    if ws > 1 and not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        # a mismatched collective should fail in minutes, not hold N GPUs for the default 10
        timeout = datetime.timedelta(seconds=int(os.environ.get("PD_FUSION_B200_DIST_TIMEOUT_S", "180")))
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            torch.distributed.init_process_group(backend=backend, rank=rank, world_size=ws, device_id=torch.device("cuda", local_rank),
                                                 timeout=timeout)
        else:
            torch.distributed.init_process_group(backend=backend, rank=rank, world_size=ws, timeout=timeout)
    return rank, local_rank, ws


def shard_range(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`: rows [r*ceil(n/G), min(n,(r+1)*ceil(n/G))).
    Concatenating the shards in rank order reproduces the reference's row order."""
    per = -(-n_rows // world_size) if n_rows > 0 else 0
    start = min(n_rows, rank * per)
    return start, min(n_rows, start + per)


_GATHER_BUFFERS: dict = {}


def all_gather_rows(local: torch.Tensor, n_rows_total: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Assembles the row-sharded table on every rank.  `local` holds this rank's shard_range rows (any trailing
    dims).  Shards are padded to ceil(n/G) rows (NCCL all-gather needs equal sizes) and the pad is trimmed.
    The staging shard and the gathered table are cached per (shape, dtype, device): a step loop that gathers the same
    table every step allocates nothing (the result is overwritten by the next call with the same signature; pass `out` or
    clone to keep it)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return local
    ws = torch.distributed.get_world_size()
    per = -(-n_rows_total // ws)
    tail = tuple(local.shape[1:])
    if local.shape[0] == per and local.is_contiguous():
        src = local                                      # equal shards: gather straight from the caller's tensor
    else:
        key = ("pad", per, tail, local.dtype, local.device)
        src = _GATHER_BUFFERS.get(key)
        if src is None:
            src = _GATHER_BUFFERS[key] = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
        src[: local.shape[0]].copy_(local)
        if local.shape[0] < per:
            src[local.shape[0]:].zero_()
    if out is None:
        key = ("out", ws * per, tail, local.dtype, local.device)
        out = _GATHER_BUFFERS.get(key)
        if out is None:
            out = _GATHER_BUFFERS[key] = torch.empty((ws * per,) + tail, dtype=local.dtype, device=local.device)
    torch.distributed.all_gather_into_tensor(out, src)
    return out[:n_rows_total]


def barrier() -> None:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()
