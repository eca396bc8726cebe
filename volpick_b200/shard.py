"""Multi-GPU sharding of independent records (station-days) -- SURVEY.md section 8(e).

Records are independent: each rank (one process per GPU) annotates its own slice of the record
list with no collective on the data path.  The only exchange is the final gather of the
variable-length pick lists to rank 0 (``torch.distributed.gather_object``; works with NCCL and gloo).
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence


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
