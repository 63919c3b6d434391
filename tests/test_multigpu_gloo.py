"""world_size-2 gloo test of the multi-GPU host logic (CPU, no GPU needed):
sample sharding + histogram all-reduce gives the single-process tables, and
chunk-range sharding reproduces the single-process streams.  The per-rank
compute is done by the oracle here (the product has no CPU path); on the GPU
box bench.py runs the same logic over NCCL with the CUDA kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_fixture


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fqcomp28_b200 import multigpu as M
    from oracle import oracle as O

    d = np.concatenate([load_fixture("SRR065390_sub_1"), load_fixture("SRR065390_sub_2")])
    S, R = 300_000, 40_000
    # sample = records wholly inside the first S bytes (src/prepare.cpp:42-47)
    sample = d[: int(O.split_chunks(d, S)[1])]
    recs, _ = O.parse_records(sample)
    a, b = M.shard_range(len(recs), rank, world)
    cs, cq = O.hist(sample, recs[a:b])
    ts, tq = torch.from_numpy(cs.view(np.int32).reshape(-1).copy()), torch.from_numpy(cq.view(np.int32).reshape(-1).copy())
    M.allreduce_counts(ts, tq)
    fs, fq = O.make_ft(ts.numpy().view(np.uint32), tq.numpy().view(np.uint32))
    # chunk-range sharding of the whole file
    offs = O.split_chunks(d, R)
    ca, cb = M.shard_by_bytes(offs, rank, world)
    cod = O.Codec(fs, fq)
    mine = []
    for k in range(ca, cb):
        sub = d[int(offs[k]) : int(offs[k + 1])]
        r, _ = O.parse_records(sub)
        e = cod.encode_chunk(sub, r)
        mine.append((k, e["seq"].tobytes(), e["qual"].tobytes()))
    gathered = [None] * world
    dist.all_gather_object(gathered, (fs.tobytes(), fq.tobytes(), mine, (ca, cb)))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_tables_and_streams_match_single_process(oracle):
    O = oracle
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process truth
    d = np.concatenate([load_fixture("SRR065390_sub_1"), load_fixture("SRR065390_sub_2")])
    S, R = 300_000, 40_000
    sample = d[: int(O.split_chunks(d, S)[1])]
    recs, _ = O.parse_records(sample)
    fs, fq = O.make_ft(*O.hist(sample, recs))
    assert all(g[0] == fs.tobytes() and g[1] == fq.tobytes() for g in gathered), "tables differ across ranks"
    offs = O.split_chunks(d, R)
    cod = O.Codec(fs, fq)
    seen = []
    for g in gathered:
        for k, s, qq in g[2]:
            sub = d[int(offs[k]) : int(offs[k + 1])]
            r, _ = O.parse_records(sub)
            e = cod.encode_chunk(sub, r)
            assert e["seq"].tobytes() == s and e["qual"].tobytes() == qq
            seen.append(k)
    assert sorted(seen) == list(range(len(offs) - 1)), "chunk ranges must tile the archive"
    assert gathered[0][3][1] == gathered[1][3][0]


def _chain_worker(rank, world, port, q, use_shm):
    """the N-rank chunking of bench.py / DESIGN.md section 7 with the oracle's walk standing in for
    fq28_plan_cut_dev: record ranges + lookahead (slab_end), one cut offset per rank (Baton)"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fqcomp28_b200 import multigpu as M
    from oracle import oracle as O

    d = np.concatenate([load_fixture("SRR065390_sub_1"), load_fixture("SRR065390_sub_2"), load_fixture("without_ns")])
    R = 50_000
    recs, _ = O.parse_records(d)
    ends = np.append(recs["hdr_off"].astype(np.int64), d.size)
    per = (len(recs) + world - 1) // world
    b0, b1 = int(ends[min(rank * per, len(recs))]), int(ends[min((rank + 1) * per, len(recs))])
    last = rank == world - 1
    slab = d[b0 : M.slab_end(b1, R, d.size, last)]
    baton = M.Baton(rank, world, shared_memory=use_shm)
    assert (baton._arr is not None) == use_shm
    out = []
    for step in range(70):                      # more steps than ring slots: a slot of one step must not leak into another
        baton.next_step()
        cut = baton.recv()
        first = cut - b0
        assert 0 <= first < R
        # walk from the cut; like the GPU walk with eof = 0, only chunks whose window fits are the rank's own
        offs = [int(o) + first for o in O.split_chunks(slab[first:], R)]
        if not last:
            offs = [o for i, o in enumerate(offs) if i == 0 or offs[i - 1] + R <= slab.size]
        baton.send(b0 + offs[-1])
        out.append([(b0 + offs[i], b0 + offs[i + 1]) for i in range(len(offs) - 1)])
    assert all(o == out[0] for o in out)
    baton.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, out[0])
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,use_shm", [(2, True), (3, True), (3, False)])
def test_cut_chain_tiles_the_file_like_one_walk(oracle, world, use_shm):
    port = 31500 + (os.getpid() % 2000) + world + 7 * int(use_shm)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chain_worker, args=(r, world, port, q, use_shm)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    d = np.concatenate([load_fixture("SRR065390_sub_1"), load_fixture("SRR065390_sub_2"), load_fixture("without_ns")])
    offs = [int(o) for o in oracle.split_chunks(d, 50_000)]
    chunks = [c for g in gathered for c in g]
    assert chunks == [(offs[i], offs[i + 1]) for i in range(len(offs) - 1)]
    assert all(len(g) >= 2 for g in gathered)


def test_slab_end():
    from fqcomp28_b200 import multigpu as M

    assert M.slab_end(1000, 100, 5000, False) == 1099
    assert M.slab_end(1000, 100, 1050, False) == 1050
    assert M.slab_end(1000, 100, 5000, True) == 5000


def test_shard_helpers():
    from fqcomp28_b200 import multigpu as M

    for n in (0, 1, 7, 1000):
        for w in (1, 2, 3, 8):
            rs = [M.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    offs = np.array([0, 10, 25, 27, 60, 61, 100], dtype=np.uint64)
    for w in (1, 2, 4, 8):
        rs = [M.shard_by_bytes(offs, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == 6
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        M.shard_range(5, 2, 2)
