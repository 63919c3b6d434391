"""Host-side multi-GPU logic of the path.

The path shards by chunk range: every GPU encodes / decodes its own contiguous
range of chunks independently.  The only collective is the sum of the sample
histograms (256*4 + 8192*64 = 525 312 u32 counters, 2.1 MB) when the sample is
histogrammed cooperatively; integer sums are order independent, so the tables
every rank builds afterwards are bit-identical to a single-GPU run.  The +1
prior (src/fse_sequence.cpp:149-150, src/fse_quality.cpp:75-76) is added after
the reduction, inside fq28_build_tables.

Chunk boundaries are a sequential recurrence over the file (FastqReader::
readNextChunk, src/fastq_io.cpp:23-65).  With one rank per GPU, rank r holds its
record range plus reading_size - 1 bytes of lookahead (`slab_end`), pre-parses
it at once, and learns from rank r-1 one number -- the file offset at which that
rank's last chunk ends (`Baton`) -- from which its own boundary walk starts
(fq28_plan_cut_dev).  The chunks of all ranks are then exactly those of the
single-process walk.

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


def slab_end(own_end: int, reading_size: int, file_size: int, is_last: bool) -> int:
    """End (global offset, exclusive) of the bytes rank r must hold when its records end at
    `own_end`: reading_size - 1 bytes of lookahead.  With exactly that much, a chunk that
    starts before own_end has its whole window inside the slab (so the walk, which only emits
    chunks whose window fits, emits it), and a chunk that starts at or after own_end does not
    (it belongs to the next rank)."""
    if is_last:
        return file_size
    return min(file_size, own_end + reading_size - 1)


class Baton:
    """The one number that travels between ranks: the global file offset at which the previous
    rank's last chunk ends.  Control metadata, not a data-path collective.  Rank r blocks in
    recv() until rank r-1 has walked its boundaries, so the hop latency is on the critical path
    of every step: N - 1 hops in a row.  Measured over the process group's TCP store a hop cost
    ~0.55 ms (8 GPUs: 81 % of linear for an 18 ms compress step); with all ranks on one node the
    baton therefore lives in a shared-memory segment (a ring of (step, cut) slots per rank, the
    receiver spins; a sender may run SLOTS - 1 steps ahead of its receiver, then waits for its
    acknowledgement), and the store only carries the segment's name once.  Several nodes: the store, one key per
    (step, rank)."""

    TIMEOUT_S = 120.0
    SLOTS = 64

    def __init__(self, rank: int, world: int, store=None, shared_memory=None):
        self.rank, self.world, self.step = rank, world, 0
        self.store = store
        self._shm = None
        self._arr = None
        if world > 1 and store is None:
            from torch.distributed.distributed_c10d import _get_default_store

            self.store = _get_default_store()
        if shared_memory is None:
            import os

            lws = os.environ.get("LOCAL_WORLD_SIZE")
            shared_memory = world > 1 and (lws is None or int(lws) == world)
        if world > 1 and shared_memory:
            self._attach()

    def _attach(self) -> None:
        import numpy as np
        from multiprocessing import shared_memory as shm

        key = "fq28/baton_shm"
        if self.rank == 0:
            seg = shm.SharedMemory(create=True, size=8 * self.world * (2 * self.SLOTS + 1))
            arr = np.ndarray((self.world, 2 * self.SLOTS + 1), dtype=np.int64, buffer=seg.buf)
            arr[:] = 0
            self.store.set(key, seg.name)
        else:
            seg = shm.SharedMemory(name=self.store.get(key).decode())
            try:   # (the creator unlinks it; keep Python's resource tracker of this process out of it)
                from multiprocessing import resource_tracker

                resource_tracker.unregister(seg._name, "shared_memory")
            except Exception:
                pass
            arr = np.ndarray((self.world, 2 * self.SLOTS + 1), dtype=np.int64, buffer=seg.buf)
        self._shm, self._arr = seg, arr   # row r: SLOTS x (step, cut) written by rank r-1, then the last step rank r has read

    def close(self) -> None:
        if self._shm is not None:
            self._arr = None
            try:
                self._shm.close()
                if self.rank == 0:
                    self._shm.unlink()
            except Exception:
                pass
            self._shm = None

    def next_step(self) -> None:
        self.step += 1

    def _spin(self, done, what: str) -> None:
        import time

        t0, spins = time.monotonic(), 0
        while not done():
            spins += 1
            if (spins & 0x3FFF) == 0 and time.monotonic() - t0 > self.TIMEOUT_S:
                raise TimeoutError(f"baton, rank {self.rank}: {what}")

    def recv(self) -> int:
        if self.rank == 0:
            return 0
        if self._arr is not None:
            row, step = self._arr[self.rank], self.step
            i = 2 * (step % self.SLOTS)
            self._spin(lambda: int(row[i]) == step, f"no cut from rank {self.rank - 1} for step {step}")
            cut = int(row[i + 1])
            row[2 * self.SLOTS] = step   # acknowledged: the slot may be reused SLOTS steps from now
            return cut
        return int(self.store.get(f"fq28/cut/{self.step}/{self.rank}").decode())

    def send(self, cut: int) -> None:
        if self.rank + 1 >= self.world:
            return
        if self._arr is not None:
            row, step = self._arr[self.rank + 1], self.step
            self._spin(lambda: int(row[2 * self.SLOTS]) > step - self.SLOTS, f"rank {self.rank + 1} is {self.SLOTS} steps behind")
            i = 2 * (step % self.SLOTS)
            row[i + 1] = int(cut)   # the value first, then the step that publishes it (x86: stores stay in order)
            row[i] = step
            return
        self.store.set(f"fq28/cut/{self.step}/{self.rank + 1}", str(int(cut)))
