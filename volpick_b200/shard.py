"""Multi-GPU sharding of independent records (station-days) -- SURVEY.md section 8(e).

Records are independent: each rank (one process per GPU) annotates its own slice of the record
list with no collective on the data path.  The only exchange is the final gather of the
variable-length pick lists to rank 0: ``gather_triggers`` (two ``all_gather`` calls of byte tensors, what
``bench.py`` uses) for trigger tables, ``gather_picks`` (``gather_object``) for arbitrary payloads; both work with
NCCL and gloo.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np


def shard_indices(n_records: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment: record i goes to rank i % world_size (balances unequal record lengths)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_records, world_size))


def gather_picks(local: Sequence[Any], rank: int, world_size: int) -> Optional[List[Any]]:
    """Gather per-rank result lists [(record_index, payload), ...] on rank 0, ordered by record index."""
    if world_size == 1:
        return sorted(local, key=lambda r: r[0])
    import torch.distributed as dist

    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(list(local), gathered, dst=0)
    if rank != 0:
        return None
    merged = [item for part in gathered for item in part]
    return sorted(merged, key=lambda r: r[0])


def gather_triggers(local: Sequence[Any], rank: int, world_size: int, device=None) -> Optional[List[Any]]:
    """``gather_picks`` for payloads that are NumPy structured trigger arrays (``_lib.TRIGGER_DTYPE``, 32 bytes each):
    every rank sends ONE byte tensor (its triggers back to back) and one int64 index table, padded to the longest rank and
    exchanged with two ``all_gather`` calls (NCCL over NVLink on a GPU box, gloo on CPU) -- a thousand station-days carry
    2 M triggers = 69 MB, which pickling through ``gather_object`` moved in 1.2 - 1.7 s (8 GPUs) against 2.7 s of compute."""
    if world_size == 1:
        return sorted(local, key=lambda r: r[0])
    import torch
    import torch.distributed as dist

    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    dtype = local[0][1].dtype if len(local) else None
    itemsize = int(dtype.itemsize) if dtype is not None else 32
    table = np.array([[i, len(t)] for i, t in local], dtype=np.int64).reshape(-1, 2)
    blob = np.concatenate([np.ascontiguousarray(t).view(np.uint8).reshape(-1) for _, t in local]) if len(local) else np.zeros(0, np.uint8)
    sizes = torch.tensor([table.shape[0], blob.size, itemsize], dtype=torch.int64, device=device)
    all_sizes = [torch.empty_like(sizes) for _ in range(world_size)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    max_rec, max_bytes = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    tab_t = torch.zeros((max(max_rec, 1), 2), dtype=torch.int64, device=device)
    blob_t = torch.zeros(max(max_bytes, 1), dtype=torch.uint8, device=device)
    if table.shape[0]:
        tab_t[: table.shape[0]] = torch.from_numpy(table).to(device)
    if blob.size:
        blob_t[: blob.size] = torch.from_numpy(blob).to(device)
    tabs = [torch.empty_like(tab_t) for _ in range(world_size)]
    blobs = [torch.empty_like(blob_t) for _ in range(world_size)]
    dist.all_gather(tabs, tab_t)
    dist.all_gather(blobs, blob_t)
    if rank != 0:
        return None
    if dtype is None:  # this rank had no records: take the record type from the sizes the others announced
        from . import _lib

        dtype = _lib.TRIGGER_DTYPE
    merged = []
    for r in range(world_size):
        n_rec = int(all_sizes[r, 0])
        tab = tabs[r][:n_rec].cpu().numpy()
        raw = blobs[r][: int(all_sizes[r, 1])].cpu().numpy()
        off = 0
        for i, cnt in tab:
            nb = int(cnt) * dtype.itemsize
            merged.append((int(i), raw[off : off + nb].view(dtype).copy()))
            off += nb
    return sorted(merged, key=lambda r: r[0])
