"""Host-side multi-GPU logic of the path.

The path shards by chunk range: every GPU encodes / decodes its own contiguous
range of chunks independently.  The only exchange is the sum of the sample
histograms (256*4 + 8192*64 = 525 312 u32 counters, 2.1 MB) when the sample is
histogrammed cooperatively; integer sums are order independent, so the tables
every rank builds afterwards are bit-identical to a single-GPU run.  The +1
prior (src/fse_sequence.cpp:149-150, src/fse_quality.cpp:75-76) is added after
the reduction, inside fq28_build_tables.

Backend agnostic: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

SEQ_COUNTERS = 256 * 4
QUAL_COUNTERS = 8192 * 64


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [a, b) of `n_items` for `rank` (records of the
    sample, or chunks of the archive)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return n_items * rank // world, n_items * (rank + 1) // world


def shard_by_bytes(offsets, rank: int, world: int) -> tuple[int, int]:
    """Chunk range [a, b) for `rank` balanced by bytes: offsets = chunk
    boundaries (n_chunks + 1 entries, as fq28_split returns them)."""
    n = len(offsets) - 1
    total = int(offsets[-1]) - int(offsets[0])
    lo = int(offsets[0]) + total * rank // world
    hi = int(offsets[0]) + total * (rank + 1) // world

    def first_at_or_after(x):
        a, b = 0, n
        while a < b:
            m = (a + b) // 2
            if int(offsets[m]) < x:
                a = m + 1
            else:
                b = m
        return a

    a = first_at_or_after(lo)
    b = n if rank == world - 1 else first_at_or_after(hi)
    return a, max(a, b)


def allreduce_counts(seq_counts, qual_counts, group=None):
    """C1: the path's only collective.  seq_counts / qual_counts are int32
    torch tensors (device tensors with NCCL, CPU tensors with gloo) holding the
    raw u32 counters WITHOUT the +1 prior; summed in place over all ranks."""
    import torch.distributed as dist

    if seq_counts.numel() != SEQ_COUNTERS or qual_counts.numel() != QUAL_COUNTERS:
        raise ValueError("unexpected histogram shape")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(seq_counts, group=group)
        dist.all_reduce(qual_counts, group=group)
    return seq_counts, qual_counts
