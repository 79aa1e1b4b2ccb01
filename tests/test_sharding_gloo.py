"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (read-range sharding, image broadcast, counter reduction)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rapmap_b200.sharding import broadcast_bytes, reduce_counters, shard_range


def test_shard_ranges_tile_the_input():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import SynthTxome

    total = 5001
    tx = SynthTxome(777, 8)
    b, e = shard_range(total, rank, world)
    s1, s2 = tx.reads(e - b, rseed=4242, first=b)
    # the counter-based stream makes a rank's shard identical to the same slice of the unsharded stream
    f1, f2 = tx.reads(total, rseed=4242, first=0)
    ok = np.array_equal(s1, f1[b:e]) and np.array_equal(s2, f2[b:e])
    # image replication protocol: only rank 0 knows the size
    blob = torch.arange(1000, dtype=torch.int64).to(torch.uint8) if rank == 0 else None
    got = broadcast_bytes(blob, 0)
    ok = ok and got.numel() == 1000 and int(got[999]) == 999 % 256
    # counters: every rank contributes its own
    tot = reduce_counters([e - b, rank, 1, 2 * (e - b), 0])
    ok = ok and tot.tolist() == [total, sum(range(world)), world, 2 * total, 0]
    with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok{r}").read_text() == "1"
