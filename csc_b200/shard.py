"""Chunk -> rank partition for one box: contiguous chunk ranges per GPU, no
data-path collective; the host concatenates per-chunk streams in chunk order."""
from __future__ import annotations


def chunk_range(n_chunks: int, rank: int, world: int):
    """Chunks [lo, hi) owned by ``rank``: gpu = chunk_id * world // n_chunks, contiguous."""
    lo = (n_chunks * rank + world - 1) // world
    hi = (n_chunks * (rank + 1) + world - 1) // world
    return lo, hi


def owner(chunk_id: int, n_chunks: int, world: int) -> int:
    return chunk_id * world // n_chunks
