"""Sequence -> rank assignment for multi-GPU runs.

Sequences are independent (one emulator instance each), so N GPUs are N shards of the sequence
list with no collective on the data path.  Contiguous blocks keep a sequence's packs on one
device for its whole life (the emulator state never moves).
"""
from __future__ import annotations

from typing import List


def shard_sequences(n_sequences: int, rank: int, world_size: int) -> range:
    """Contiguous block of sequence indices owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_sequences, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(sequence: int, n_sequences: int, world_size: int) -> int:
    """Rank that owns `sequence` under shard_sequences."""
    for r in range(world_size):
        if sequence in shard_sequences(n_sequences, r, world_size):
            return r
    raise ValueError("sequence out of range")


def all_shards(n_sequences: int, world_size: int) -> List[range]:
    return [shard_sequences(n_sequences, r, world_size) for r in range(world_size)]
