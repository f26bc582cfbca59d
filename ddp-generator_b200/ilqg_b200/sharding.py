"""Batch sharding across ranks (one process per GPU).  Problems are independent, so a shard is just a contiguous
block of problem ids; there is no exchange on the solve path and only counters are reduced at the end."""
from __future__ import annotations


def shard_range(batch, rank, world):
    """[first, first+count) of the problem ids owned by `rank`: equal blocks, the last rank takes the remainder."""
    per = batch // world
    first = rank * per
    count = per if rank < world - 1 else batch - first
    return first, count
