"""Multi-GPU sharding of the read stream.

Reads are independent units: each rank (one process per GPU) takes one CONTIGUOUS shard of the
batch, runs the whole path on its own device with its own replica of the (KB-sized) match tables,
and rank 0 gathers the write operations in rank order -- which is input order.  There is no
collective on the data path; torch.distributed (NCCL or gloo) only carries the final gather.
"""
from typing import List, Tuple


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of shard `rank` (the first n % world shards get one extra)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_ordered(obj, dist, dst: int = 0) -> List:
    """Gather one python object per rank onto `dst`, in rank order (None elsewhere)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def process_sequences_sharded(seq_records, parameters, specimens, args, prefilter=None, dist=None,
                              device: int = 0, _binding=None):
    """process_sequences over a contiguous shard per rank; rank 0 returns the reference-ordered
    (write_ops, total, matched) of the whole batch, other ranks return ([], 0, 0)."""
    from .demultiplex import process_sequences
    if dist is None or not dist.is_initialized():
        return process_sequences(seq_records, parameters, specimens, args, prefilter, None, 0, device, _binding)
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_bounds(len(seq_records), world, rank)
    part = process_sequences(seq_records[lo:hi], parameters, specimens, args, prefilter, None, lo, device, _binding)
    gathered = gather_ordered(part, dist)
    if rank != 0:
        return [], 0, 0
    ops, total, matched = [], 0, 0
    for o, t, m in gathered:
        ops.extend(o)
        total += t
        matched += m
    return ops, total, matched
