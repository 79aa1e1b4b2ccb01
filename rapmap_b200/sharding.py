"""Multi-GPU plumbing: one process per GPU, index replicated, read pairs sharded by contiguous range.

The reference's only parallelism is data parallelism over reads inside one process (N consumer threads pulling
10k-read chunks, reference src/RapMapSAMapper.cpp:752-799); the index is shared read-only.  Across GPUs that maps
to: every rank holds a full index replica (one broadcast of the packed image) and maps its own range of pairs.
No collective sits on the data path; the HitCounters are summed once at the end.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) of rank's items; ranges of all ranks tile [0, n_items) in rank order."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_counters(counters, device="cpu"):
    """Sum of the five HitCounters (numReads, peHits, seHits, totHits, tooManyHits) over all ranks."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(np.asarray(counters, dtype=np.int64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def broadcast_bytes(blob, src: int = 0):
    """Broadcast of a uint8 tensor whose size only the source knows (index image replication)."""
    import torch
    import torch.distributed as dist

    dev = blob.device if blob is not None else torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([blob.numel() if dist.get_rank() == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if dist.get_rank() != src:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(blob, src)
    return blob


def replicate_index(index, rank: int, local_device: int):
    """Rank 0 passes its loaded Index, the others None; every rank returns an Index on its own GPU."""
    import torch
    import torch.distributed as dist

    from . import Index

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return index, None
    blob = None
    if rank == 0:
        ptr, nbytes = index.image()

        class _Ext:  # zero-copy torch view of the packed image
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

        blob = torch.as_tensor(_Ext(), device=torch.device("cuda", local_device))
    blob = broadcast_bytes(blob, 0)
    torch.cuda.synchronize()
    if rank != 0:
        index = Index.from_image(local_device, blob.data_ptr(), blob.numel())
    return index, blob  # keep `blob` alive as long as the index is used
