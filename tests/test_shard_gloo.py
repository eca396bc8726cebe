"""world_size-2 gloo test of the multi-GPU host logic (sharding + final pick gather), on CPU."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from volpick_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard.shard_indices(7, rank, world)
    local = [(i, [("P", i * 10 + k) for k in range(i % 3)]) for i in idx]  # variable-length pick lists
    merged = shard.gather_picks(local, rank, world)
    # the tensor path used by bench.py: structured trigger arrays, variable length, one rank possibly empty-handed
    import numpy as np

    from volpick_b200 import _lib

    trig_local = []
    for i in (idx if rank == 0 else idx[:1]):
        t = np.zeros(i % 4, dtype=_lib.TRIGGER_DTYPE)
        t["s0"] = i * 100 + np.arange(len(t))
        t["value"] = 0.5 + i
        t["label"] = i % 3
        trig_local.append((i, t))
    merged_t = shard.gather_triggers(trig_local, rank, world)
    if rank == 0:
        q.put((merged, [(i, t.tolist()) for i, t in merged_t]))
    else:
        assert merged is None and merged_t is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    for n in (0, 1, 7, 1000):
        for w in (1, 2, 4, 8):
            parts = [shard.shard_indices(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def test_gather_picks_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, merged_t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [i for i, _ in merged] == list(range(7))
    assert merged[5][1] == [("P", 50), ("P", 51)]
    assert [i for i, _ in merged_t] == [0, 1, 2, 4, 6]  # rank 0 owns 0, 2, 4, 6; rank 1 sent only its first record (1)
    assert len(merged_t[2][1]) == 2 and merged_t[2][1][1][0] == 201 and merged_t[4][1][0][3] == 6.5
    single = shard.gather_picks([(1, "b"), (0, "a")], 0, 1)
    assert single == [(0, "a"), (1, "b")]
