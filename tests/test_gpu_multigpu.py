"""One file over several GPUs and the robustness fixes of round 2.

The chunk boundaries of a file are ONE sequential walk (FastqReader::readNextChunk,
src/fastq_io.cpp:23-65): slab i+1 starts where slab i stopped.  fq28_plan gives that
offset without encoding, fq28_stage starts the H2D copy before it is known.  These
tests check, through the C ABI and the CLI, that cutting a file into slabs handled by
different handles / GPUs yields exactly the blocks of the one-pass run, in file order
(src/process.cpp:32-105, src/archive.cpp:57-106), and that corrupt side information
or swapped tables are reported instead of decoded into garbage.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_cli_archive import CLI, build_cli, parse_archive

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import fqcomp28_b200

    return fqcomp28_b200


@pytest.fixture(scope="module")
def data():
    import synth

    return synth.illumina(0, 20000, seed=31).numpy()  # ~7 MB, ~7 chunks at R = 1 MB


def tables(oracle, d, sample):
    recs, used = oracle.parse_records(d[:sample])
    return oracle.make_ft(*oracle.hist(d[:used], recs))


def blocks_of(infos, summ, ar):
    out = []
    for k in range(int(summ.n_chunks)):
        ci = infos[k]
        out.append((int(ci.total), int(ci.n_records), ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len].tobytes(),
                    ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len].tobytes(),
                    ar["readlens"][ci.rec_off : ci.rec_off + ci.n_records].tobytes(),
                    ar["n_pos"][ci.n_pos_off : ci.n_pos_off + ci.n_pos_len].tobytes()))
    return out


def test_plan_then_compress_is_compress(P, oracle, data):
    R = 1 << 20
    fs, fq = tables(oracle, data, 2 << 20)
    h = P.Handle(0)
    h.load_tables(fs, fq)
    ref = blocks_of(*h.compress(data, R, eof=True))
    for eof in (True, False):
        consumed, n = h.plan(data, R, eof=eof)
        offs = h.split(data, R, eof=eof)
        assert consumed == int(offs[-1]) and n == len(offs) - 1
        consumed2, _ = h.plan(data, R, eof=eof)          # plan twice: same answer
        assert consumed2 == consumed
        infos, summ, ar = h.compress(data, R, eof=eof)   # reuses the plan
        assert int(summ.consumed) == consumed
        assert blocks_of(infos, summ, ar) == ref[: int(summ.n_chunks)]
    h.close()


def chained(P, fs, fq, d, R, B, n_workers, use_stage):
    """what the CLI's workers do: slab i = [i*B, (i+1)*B + R) on handle i mod N, started at the
    offset the previous slab's plan returned"""
    hs = [P.Handle(0) for _ in range(n_workers)]
    for h in hs:
        h.load_tables(fs, fq)
    out, start, i = [], 0, 0
    while True:
        base, end = i * B, min(d.size, (i + 1) * B + R)
        eof = end == d.size
        buf = np.ascontiguousarray(d[base:end])
        h = hs[i % n_workers]
        if use_stage:
            h.stage(buf)
        assert base <= start <= end
        sub = buf[start - base :]
        consumed, _ = h.plan(sub, R, eof=eof)
        infos, summ, ar = h.compress(sub, R, eof=eof)
        assert int(summ.consumed) == consumed
        out += blocks_of(infos, summ, ar)
        if use_stage:
            h.stage(None)
        if eof:
            break
        start += consumed
        i += 1
    for h in hs:
        h.close()
    return out


@pytest.mark.parametrize("use_stage", [False, True])
def test_slab_chain_equals_one_pass(P, oracle, data, use_stage):
    R = 1 << 20
    fs, fq = tables(oracle, data, 2 << 20)
    h = P.Handle(0)
    h.load_tables(fs, fq)
    ref = blocks_of(*h.compress(data, R, eof=True))
    h.close()
    offs = oracle.split_chunks(data, R)
    assert len(ref) == len(offs) - 1
    for B, n in ((2 << 20, 2), (3 << 20, 3), (2 << 20, 1)):
        assert chained(P, fs, fq, data, R, B, n, use_stage) == ref, (B, n)


def run_cli(args, **kw):
    return subprocess.run([CLI] + args, capture_output=True, text=True, check=True, **kw)


def test_cli_gpus_same_archive(tmp_path, data, P):
    """--gpus N does not change the archive (with one visible device the workers share it; with
    two or more they really are different GPUs), and `d --gpus N` restores the input."""
    build_cli()
    src = str(tmp_path / "s.fastq")
    data.tofile(src)
    a1, a2, a3, out = (str(tmp_path / n) for n in ("g1.fqz", "g2.fqz", "g3.fqz", "o.fastq"))
    common = ["c", "--i1", src, "-R", "1", "-S", "2", "--slab-mb", "2"]
    run_cli(common + ["-o", a1])
    run_cli(common + ["-o", a2, "--gpus", "2"])
    run_cli(common + ["-o", a3, "--gpus", "3", "--slab-mb", "3"])
    assert open(a1, "rb").read() == open(a2, "rb").read() == open(a3, "rb").read()
    for g in ("1", "2", "3"):
        run_cli(["d", "-i", a2, "--o1", out, "--gpus", g, "--slab-mb", "2"])
        assert np.array_equal(np.fromfile(out, dtype=np.uint8), data), g
    first, _, _, blocks = parse_archive(open(a1, "rb").read(), None)
    assert [b["idx"] for b in blocks] == list(range(len(blocks))) and len(blocks) >= 6


def test_cli_two_real_gpus(tmp_path, data, P):
    if P.load().fq28_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    build_cli()
    src = str(tmp_path / "s.fastq")
    data.tofile(src)
    a1, a2, out = str(tmp_path / "g1.fqz"), str(tmp_path / "g2.fqz"), str(tmp_path / "o.fastq")
    common = ["c", "--i1", src, "-R", "1", "-S", "2", "--slab-mb", "2"]
    run_cli(common + ["-o", a1])
    r = run_cli(common + ["-o", a2, "--gpus", "2"])
    assert "workers share devices" not in r.stderr
    assert open(a1, "rb").read() == open(a2, "rb").read()
    run_cli(["d", "-i", a2, "--o1", out, "--gpus", "2", "--slab-mb", "2"])
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), data)


def test_cli_ref_compat_accumulates_n_data(tmp_path, oracle):
    """SURVEY Q2 behind --ref-compat: block k carries the n_count / n_pos of blocks 0..k
    (src/compressed_buffers.h:58-68), and still decodes (consumed from the back)."""
    import synth

    build_cli()
    d = synth.random_fastq(6000, seed=5, min_len=60, max_len=120, n_rate=0.05)
    src, arc, out = str(tmp_path / "s.fastq"), str(tmp_path / "s.fqz"), str(tmp_path / "o.fastq")
    d.tofile(src)
    # reading size is in MB on the command line; a ~1.3 MB file at -R 1 gives two chunks
    run_cli(["c", "--i1", src, "-o", arc, "-R", "1", "-S", "1", "--ref-compat"])
    _, _, _, blocks = parse_archive(open(arc, "rb").read(), None)
    assert len(blocks) >= 2
    recs_seen = 0
    for b in blocks:
        recs_seen += b["n_rec"]
        assert b["side"][1][0] == 2 * recs_seen            # original n_count bytes: all records so far
    run_cli(["d", "-i", arc, "--o1", out])
    body = d[: int(oracle.split_chunks(d, 1 << 20)[-1])]
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), body)


# ---------------------------------------------------------------- corrupt side information
def test_corrupt_side_information_is_an_error_not_a_fault(P, oracle, data):
    R = 1 << 20
    fs, fq = tables(oracle, data, 2 << 20)
    h = P.Handle(0)
    h.load_tables(fs, fq)
    infos, summ, ar = h.compress(data, R, eof=True)
    n, nrec = int(summ.n_chunks), int(summ.n_records)
    body = data[: int(summ.consumed)]
    recs, _ = oracle.parse_records(body)
    hdr, _ = oracle.gather_headers(body, recs)

    def dec(arenas, headers=hdr, n_records=nrec):
        return h.decompress(arenas, infos, n, headers, n_records, n_pos_entries=int(summ.n_pos_entries))

    assert np.array_equal(dec(ar), body)
    cases = []
    a = dict(ar); a["readlens"] = ar["readlens"].copy(); a["readlens"][nrec // 2] = 0
    cases.append(("zero read length", a, {}, (P.capi.ERR_NAMES[-4],)))
    a = dict(ar); a["readlens"] = ar["readlens"].copy(); a["readlens"][5] += 400
    cases.append(("records larger than the chunk", a, {}, ("FQ28_ERR_FORMAT",)))
    a = dict(ar); a["hdr_lens"] = ar["hdr_lens"].copy(); a["hdr_lens"][7] += 3000
    cases.append(("header lengths beyond the header bytes", a, {}, ("FQ28_ERR_FORMAT",)))
    a = dict(ar); a["n_count"] = ar["n_count"].copy(); a["n_count"][3] += 9000
    cases.append(("more N positions than stored", a, {}, ("FQ28_ERR_STREAM",)))
    cases.append(("record count mismatch", ar, {"n_records": nrec - 1}, ("FQ28_ERR_ARG",)))
    for what, arenas, kw, codes in cases:
        with pytest.raises(P.capi.Fq28Error) as e:
            dec(arenas, **kw)
        assert any(c in str(e.value) for c in codes), (what, str(e.value))
        assert np.array_equal(dec(ar), body), f"handle unusable after: {what}"  # no sticky CUDA fault
    h.close()


# ---------------------------------------------------------------- tables: one source of truth
def test_rebuilt_tables_reach_both_halves_of_a_large_slab(P, oracle, data, monkeypatch):
    """A handle that loaded tables A and then built tables B must encode a slab that takes the
    pipelined path (parts alternate between the handle and a sibling handle) with B in ALL parts."""
    monkeypatch.setenv("FQ28_PIPE_MIN_MB", "1")  # read at fq28_create: slabs >= 1 MB are pipelined
    R = 1 << 20
    h = P.Handle(0)
    monkeypatch.delenv("FQ28_PIPE_MIN_MB")
    fa = tables(oracle, data, 1 << 20)
    h.load_tables(*fa)
    h.compress(data, R, eof=True)                      # pipelined with A: the sibling has encoded with A
    cs, cq = h.hist(data[: 3 << 20][: int(oracle.split_chunks(data[: 3 << 20], 3 << 20)[1])])
    fb = h.build_tables(cs, cq)                        # B, through the counts path
    assert not np.array_equal(fa[1], fb[1])
    infos, summ, ar = h.compress(data, R, eof=True)    # sample_bytes = 0: must use B everywhere
    cod = oracle.Codec(fb[0], fb[1])
    offs = oracle.split_chunks(data, R)
    assert int(summ.n_chunks) == len(offs) - 1
    for k in range(int(summ.n_chunks)):
        sub = data[int(offs[k]) : int(offs[k + 1])]
        r, _ = oracle.parse_records(sub)
        e = cod.encode_chunk(sub, r)
        ci = infos[k]
        assert ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len].tobytes() == e["seq"].tobytes(), k
        assert ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len].tobytes() == e["qual"].tobytes(), k
    h.close()


# ---------------------------------------------------------------- one rank per GPU: record ranges + first_cut
def test_record_ranges_with_first_cut_equal_one_pass(P, oracle, data):
    """What bench.py does with N ranks: rank r holds records [r*M, (r+1)*M) of the file plus
    reading_size bytes of lookahead, pre-parses them at once, and learns from rank r-1 only the
    offset at which that rank's last chunk ends (fq28_preparse_dev + fq28_plan_cut_dev).  The
    blocks of all ranks, minus each rank's dropped head, are the blocks of the one-pass run."""
    import torch

    R = 1 << 20
    fs, fq = tables(oracle, data, 2 << 20)
    h = P.Handle(0)
    h.load_tables(fs, fq)
    ref = blocks_of(*h.compress(data, R, eof=True))
    recs, _ = oracle.parse_records(data)
    ends = np.append(recs["hdr_off"].astype(np.int64), data.size)
    for n_ranks in (2, 3, 5):
        M = (len(recs) + n_ranks - 1) // n_ranks
        out, cut_global = [], 0
        for r in range(n_ranks):
            b0 = int(ends[min(r * M, len(recs))])                    # global offset of the rank's first record
            b1 = int(ends[min((r + 1) * M, len(recs))])
            last = r == n_ranks - 1
            stop = data.size if last else min(data.size, b1 + R - 1)  # exactly R - 1 bytes of lookahead:
            slab = torch.from_numpy(data[b0:stop].copy()).cuda()      # chunks starting at >= b1 cannot be emitted
            h.preparse_dev(slab.data_ptr(), slab.numel())
            first_cut = cut_global - b0
            assert 0 <= first_cut < R
            consumed, n = h.plan_cut_dev(slab.data_ptr(), slab.numel(), R, last, first_cut)
            infos, summ = h.compress_dev(slab.data_ptr(), slab.numel(), R, eof=last)
            assert int(summ.consumed) == consumed and int(summ.n_chunks) == n
            ar = h.compress_fetch_all(summ)
            blocks = blocks_of(infos, summ, ar)
            out += blocks[1:] if first_cut else blocks
            cut_global = b0 + consumed
        assert out == ref, n_ranks
    with pytest.raises(P.capi.Fq28Error):  # a cut that is not a record boundary
        slab = torch.from_numpy(data[: 3 << 20].copy()).cuda()
        h.plan_cut_dev(slab.data_ptr(), slab.numel(), R, False, 12345)
    h.close()
