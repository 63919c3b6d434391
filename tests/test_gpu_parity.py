"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Bit-exact bar: histograms (u32), FreqTable images (3076 / 1081348 B),
CTable/DTable cells, seq/qual streams, readlens / n_count / n_pos, decoded
FASTQ.  The reference's own tests are re-expressed here on the same fixtures:
test/fse_sequence_test.cpp, test/fse_quality_test.cpp,
test/workspace_test.cpp:45-69, test/fastq_io_test.cpp.
"""
import hashlib

import numpy as np
import pytest

from conftest import FIXTURES, load_fixture

pytestmark = pytest.mark.gpu


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def H():
    import fqcomp28_b200

    h = fqcomp28_b200.Handle(0)
    yield h
    h.close()


def oracle_tables(O, d):
    recs, _ = O.parse_records(d)
    cs, cq = O.hist(d, recs)
    fs, fq = O.make_ft(cs, cq)
    return recs, cs, cq, fs, fq


def check_chunks(O, h, d, R, fs, fq, eof=True):
    """GPU compress of slab d at reading size R == oracle chunk by chunk, and
    GPU decompress restores the bytes."""
    infos, summ, ar = h.compress(d, R, eof=eof)
    offs = O.split_chunks(d, R)
    if not eof:  # only whole windows
        offs = np.array([o for i, o in enumerate(offs) if i == 0 or int(offs[i - 1]) + R <= d.size], dtype=np.uint64)
    n = int(summ.n_chunks)
    assert n == len(offs) - 1
    cod = O.Codec(fs, fq)
    rec_base = 0
    for k in range(n):
        ci = infos[k]
        a, b = int(offs[k]), int(offs[k + 1])
        assert (ci.fastq_off, ci.total) == (a, b - a)
        sub = d[a:b]
        recs, used = O.parse_records(sub)
        assert used == sub.size
        enc = cod.encode_chunk(sub, recs)
        assert ci.n_records == len(recs) and ci.rec_off == rec_base
        assert np.array_equal(ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], enc["seq"]), f"seq stream chunk {k}"
        assert np.array_equal(ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len], enc["qual"]), f"qual stream chunk {k}"
        sl = slice(rec_base, rec_base + len(recs))
        assert np.array_equal(ar["readlens"][sl], enc["readlens"])
        assert np.array_equal(ar["n_count"][sl], enc["n_count"])
        assert np.array_equal(ar["hdr_lens"][sl], recs["hdr_len"].astype(np.uint16))
        assert np.array_equal(ar["n_pos"][ci.n_pos_off : ci.n_pos_off + ci.n_pos_len], enc["n_pos"])
        rec_base += len(recs)
    assert int(summ.consumed) == int(offs[-1])
    # decode everything in one batch
    body = d[: int(offs[-1])]
    recs_all, _ = O.parse_records(body)
    hdr, _ = O.gather_headers(body, recs_all)
    out = h.decompress(ar, infos, n, hdr, int(summ.n_records), n_pos_entries=int(summ.n_pos_entries))
    assert np.array_equal(out, body)
    return infos, summ, ar


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("name", FIXTURES)
def test_parse_records(oracle, H, name):
    d = load_fixture(name)
    recs, used = oracle.parse_records(d)
    g, cons = H.parse(d)
    assert cons == used
    for k in ("hdr_off", "seq_off", "qual_off", "hdr_len", "len"):
        assert np.array_equal(g[k].astype(np.uint64), recs[k].astype(np.uint64)), k


def test_parse_truncated_slabs(oracle, H):
    """parseRecords' return value for every kind of cut (src/fastq_io.cpp:78-90)."""
    d = load_fixture("SRR065390_1_first5")
    for cut in list(range(1, 40)) + list(range(250, 540, 7)) + [d.size - 1, d.size]:
        recs, used = oracle.parse_records(d[:cut])
        g, cons = H.parse(d[:cut])
        assert cons == used and len(g["len"]) == len(recs), cut


def test_split_every_block_size(oracle, H):
    """test/fastq_io_test.cpp:15-53: every block size 1000..filesize."""
    d = load_fixture("SRR065390_1_first5")
    for R in range(1000, d.size + 1):
        assert np.array_equal(H.split(d, R), oracle.split_chunks(d, R)), R


def test_split_many_chunks(oracle, H):
    d = load_fixture("SRR065390_sub_1")
    for R in (300, 1000, 4096, 65536, 1 << 20):
        assert np.array_equal(H.split(d, R), oracle.split_chunks(d, R)), R
    # trailing partial record is dropped at EOF
    cut = d[:-1]
    assert np.array_equal(H.split(cut, 4096), oracle.split_chunks(cut, 4096))


# ------------------------------------------------------------------ K3 / K4
@pytest.mark.parametrize("name", FIXTURES)
def test_hist_and_tables(oracle, H, name):
    d = load_fixture(name)
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    gcs, gcq = H.hist(d)
    assert np.array_equal(gcs, cs)
    assert np.array_equal(gcq, cq)
    gfs, gfq = H.build_tables(gcs, gcq)
    assert np.array_equal(gfs, fs), "ft_seq image"
    assert np.array_equal(gfq, fq), "ft_qual image"
    # partial histograms add up (the multi-GPU reduction relies on it)
    half = int(recs["hdr_off"][len(recs) // 2])
    a_s, a_q = H.hist(d[:half])
    a_s, a_q = H.hist(d[half:], a_s, a_q)
    assert np.array_equal(a_s, cs) and np.array_equal(a_q, cq)


def test_ctable_dtable_cells(oracle, H):
    d = load_fixture("SRR065390_sub_1")
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    H.load_tables(fs, fq)
    O = oracle
    for kind, ft, ctxs in ((0, fs, range(256)), (1, fq, list(range(0, 8192, 97)) + list(np.nonzero(cq.sum(1))[0][:200]))):
        norm, logs = O.ft_norm(ft), O.ft_logs(ft)
        for c in ctxs:
            c = int(c)
            st, dfs, dnb, lg = H.get_ctable(kind, c)
            ost, odfs, odnb = O.build_ctable(norm[c], int(logs[c]))
            assert lg == logs[c]
            assert np.array_equal(st, ost) and np.array_equal(dfs, odfs)
            live = norm[c] != 0
            assert np.array_equal(dnb[live], odnb[live])
            cells, _ = H.get_dtable(kind, c)
            assert np.array_equal(cells, O.build_dtable(norm[c], int(logs[c])))


# ------------------------------------------------------------------ K2 / K5 / K6 / K7
@pytest.mark.parametrize("name", FIXTURES)
def test_whole_file_chunk_golden(oracle, H, name):
    """Same digests as tests/test_oracle_golden.py."""
    from test_oracle_golden import GOLD

    d = load_fixture(name)
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    gcs, gcq = H.hist(d)
    H.build_tables(gcs, gcq)
    infos, summ, ar = check_chunks(oracle, H, d, 256 << 20, fs, fq)
    g = GOLD[name]
    ci = infos[0]
    assert (ci.seq_len, sha(ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len])) == (g[1], g[2])
    assert (ci.qual_len, sha(ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len])) == (g[4], g[5])


@pytest.mark.parametrize("R", [2000, 20000, 100000])
def test_multi_chunk(oracle, H, R):
    d = load_fixture("SRR065390_sub_1")
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    H.load_tables(fs, fq)
    check_chunks(oracle, H, d, R, fs, fq)
    check_chunks(oracle, H, d, R, fs, fq, eof=False)


def test_tables_from_other_sample(oracle, H):
    """Static tables come from a leading sample and are applied to all chunks
    (src/prepare.cpp:42-47): encode sub_2 with tables of sub_1."""
    s = load_fixture("SRR065390_sub_1")
    d = load_fixture("SRR065390_sub_2")
    _, _, _, fs, fq = oracle_tables(oracle, s)
    H.load_tables(fs, fq)
    check_chunks(oracle, H, d, 50000, fs, fq)


@pytest.mark.parametrize("seed,kw", [
    (1, dict(n_records=300, min_len=3, max_len=40)),
    (2, dict(n_records=200, min_len=3, max_len=700, n_rate=0.2)),
    (3, dict(n_records=40, min_len=2000, max_len=9000, qual_levels=64)),
    (4, dict(n_records=1, min_len=3, max_len=3)),
    (5, dict(n_records=3, min_len=65535, max_len=65535, n_rate=0.001)),
])
def test_random_fastq(oracle, H, seed, kw):
    import synth

    d = synth.random_fastq(seed=seed, **kw)
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    gcs, gcq = H.hist(d)
    assert np.array_equal(gcs, cs) and np.array_equal(gcq, cq)
    gfs, gfq = H.build_tables(gcs, gcq)
    assert np.array_equal(gfs, fs) and np.array_equal(gfq, fq)
    for R in (max(4096, d.size // 7), 256 << 20):
        if R < int((recs["hdr_len"] + 2 * recs["len"] + 6).max()):
            continue
        check_chunks(oracle, H, d, R, fs, fq)


def test_plus_line_with_text_leaves_nul_tail(oracle, H):
    """'+header' text is dropped; `total` keeps the original size so
    the decoded chunk ends with NUL bytes (src/workspace.h:130)."""
    d = np.frombuffer(b"@r1 a\nACGTN\n+r1 a\n!!#5I\n@r2 b\nGGGTT\n+\nIIIII\n", dtype=np.uint8)
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    H.load_tables(fs, fq)
    infos, summ, ar = H.compress(d, 1 << 20)
    hdr, _ = oracle.gather_headers(d, recs)
    out = H.decompress(ar, infos, 1, hdr, 2, n_pos_entries=int(summ.n_pos_entries))
    want = b"@r1 a\nACGTN\n+\n!!#5I\n@r2 b\nGGGTT\n+\nIIIII\n"
    assert bytes(out[: len(want)]) == want
    assert out.size == d.size and not out[len(want) :].any()


def test_decode_oracle_streams_at_odd_offsets(oracle, H):
    """The decoder accepts streams at arbitrary byte offsets of the arenas
    (archive blocks are not aligned, src/archive.cpp:57-106)."""
    d = load_fixture("without_ns")
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    H.load_tables(fs, fq)
    import fqcomp28_b200 as P

    enc = oracle.Codec(fs, fq).encode_chunk(d, recs)
    for pad_s, pad_q in ((1, 3), (2, 5), (7, 6)):
        ar = {
            "seq": np.concatenate([np.full(pad_s, 0xA5, np.uint8), enc["seq"], np.zeros(8, np.uint8)]),
            "qual": np.concatenate([np.full(pad_q, 0x5A, np.uint8), enc["qual"], np.zeros(8, np.uint8)]),
            "readlens": enc["readlens"], "n_count": enc["n_count"],
            "n_pos": np.zeros(1, np.uint16), "hdr_lens": recs["hdr_len"].astype(np.uint16),
        }
        infos = (P.ChunkInfo * 1)()
        ci = infos[0]
        ci.total, ci.n_records, ci.rec_off = d.size, len(recs), 0
        ci.seq_off, ci.seq_len, ci.qual_off, ci.qual_len = pad_s, enc["seq"].size, pad_q, enc["qual"].size
        ci.n_pos_off, ci.n_pos_len = 0, 0
        hdr, _ = oracle.gather_headers(d, recs)
        out = H.decompress(ar, infos, 1, hdr, len(recs), n_pos_entries=0)
        assert np.array_equal(out, d)


# ------------------------------------------------------------------ errors
def test_error_codes(oracle, H):
    import fqcomp28_b200 as P

    ok = load_fixture("SRR065390_1_first5")
    H.build_tables(*H.hist(ok))
    cases = [
        (b"@r1\nAC\n+\n!!\n", -4),                 # read shorter than 3
        (b"@r1\nACGX\n+\n!!!!\n", -3),             # base outside ACGTN
        (b"@r1\nACGT\n+\n!!!~\n", -3),             # quality above Q63
        (b"@r1\nACGT\n+\n!!!\n@r2\nACGT\n+\n!!!!\n", -2),  # qual length != seq length
        (b"r1\nACGT\n+\n!!!!\n", -2),              # header without '@'
        (b"@r1\nACGT\n-\n!!!!\n", -2),             # third line without '+'
    ]
    for raw, code in cases:
        with pytest.raises(P.Fq28Error) as e:
            H.compress(np.frombuffer(raw, dtype=np.uint8), 1 << 20)
        assert e.value.code == code, raw
    long_line = b"@r1\n" + b"A" * 70000 + b"\n+\n" + b"!" * 70000 + b"\n"
    with pytest.raises(P.Fq28Error) as e:
        H.compress(np.frombuffer(long_line, dtype=np.uint8), 1 << 20)
    assert e.value.code == -5  # narrow_cast<readlen_t> throws, src/fastq_io.cpp:95
    with pytest.raises(P.Fq28Error) as e:  # record larger than the reading size
        H.compress(ok, 100)
    assert e.value.code == -2


def test_corrupt_stream_detected(oracle, H):
    d = load_fixture("without_ns")
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    H.load_tables(fs, fq)
    import fqcomp28_b200 as P

    infos, summ, ar = H.compress(d, 1 << 20)
    hdr, _ = oracle.gather_headers(d, recs)
    infos[0].seq_len -= 1  # drop the end-mark byte
    with pytest.raises(P.Fq28Error) as e:
        H.decompress(ar, infos, 1, hdr, len(recs), n_pos_entries=0)
    assert e.value.code == -8


# ------------------------------------------------------------------ larger, size-independent properties
def test_synthetic_illumina_32mb_roundtrip_and_checksum(oracle, H):
    """BASELINE config-2 shape at a size the oracle finishes in seconds:
    identical stream checksum (FNV over all chunk streams) and exact round trip."""
    import synth
    import torch

    dev = "cuda" if torch.cuda.is_available() else "cpu"
    t, n_rec = synth.illumina_bytes(32 << 20, seed=30, device=dev)
    d = t.cpu().numpy()
    R, S = 1 << 20, 8 << 20
    res = oracle.bench(d, S, R, threads=8, do_decompress=True)
    assert res.err == 0 and res.roundtrip_ok == 1
    sample = d[: int(oracle.split_chunks(d, S)[1])]
    cs, cq = H.hist(sample)
    H.build_tables(cs, cq)
    infos, summ, ar = H.compress(d, R)
    assert int(summ.n_chunks) == res.n_chunks and int(summ.n_records) == res.n_records
    # FNV-1a over (seq stream, qual stream) of every chunk in order: the checksum fq28o_bench computes
    tot_s = tot_q = 0
    hh = 1469598103934665603
    for k in range(int(summ.n_chunks)):
        ci = infos[k]
        tot_s += ci.seq_len
        tot_q += ci.qual_len
        hh = oracle.fnv1a(ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], hh)
        hh = oracle.fnv1a(ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len], hh)
    assert (tot_s, tot_q) == (res.seq_bytes, res.qual_bytes)
    assert hh == res.checksum, "FNV over all chunk streams differs from the multi-threaded oracle run"
    recs_all, _ = oracle.parse_records(d)
    hdr, _ = oracle.gather_headers(d, recs_all)
    out = H.decompress(ar, infos, int(summ.n_chunks), hdr, int(summ.n_records), n_pos_entries=int(summ.n_pos_entries))
    assert np.array_equal(out, d)
    # per-chunk byte equality for a spread of chunks
    fs, fq = H.build_tables(cs, cq)
    cod = oracle.Codec(fs, fq)
    for k in (0, 1, int(summ.n_chunks) // 2, int(summ.n_chunks) - 1):
        ci = infos[k]
        sub = d[ci.fastq_off : ci.fastq_off + ci.total]
        recs, _ = oracle.parse_records(sub)
        enc = cod.encode_chunk(sub, recs)
        assert np.array_equal(ar["seq"][ci.seq_off : ci.seq_off + ci.seq_len], enc["seq"])
        assert np.array_equal(ar["qual"][ci.qual_off : ci.qual_off + ci.qual_len], enc["qual"])


@pytest.mark.parametrize("profile", ["hiseq", "novaseq"])
def test_synthetic_profiles_chunk_parity(oracle, H, profile):
    """BASELINE config-2 data shapes (4-bin NovaSeq / 41-level HiSeq qualities)
    at a size the oracle handles in seconds: every chunk's streams are
    byte-identical, with tables from a leading sample only."""
    import synth

    d = synth.illumina(0, 9000, seed=30, profile=profile).numpy()
    sample = d[: int(oracle.split_chunks(d, 1 << 20)[1])]
    recs, _ = oracle.parse_records(sample)
    fs, fq = oracle.make_ft(*oracle.hist(sample, recs))
    ft = (np.zeros(3076, np.uint8), np.zeros(1081348, np.uint8))
    infos, summ, ar = H.compress(d, 1 << 20, sample_bytes=1 << 20, ft_out=ft)
    assert np.array_equal(ft[0], fs) and np.array_equal(ft[1], fq)  # analyzeDataset inside fq28_compress
    check_chunks(oracle, H, d, 1 << 20, fs, fq)


@pytest.mark.parametrize("kind", ["hiseq", "ont", "fixtures", "novaseq", "outside"])
def test_windowed_quality_decoder(oracle, monkeypatch, kind):
    """The windowed layout of the quality decoder's cached cells (used when the dense layout cannot
    hold all streams of a batch at once) forced on small batches: the decode restores the input
    for many-valued qualities, long reads, the reference's fixtures, binned qualities, and for data
    whose contexts fall outside the windows of tables built from a narrow sample."""
    import synth
    import fqcomp28_b200 as P

    monkeypatch.setenv("FQ28_QUAL_WINDOWED", "1")   # read at fq28_create
    h = P.Handle(0)
    R = 1 << 18
    if kind == "fixtures":
        d = np.concatenate([load_fixture(n) for n in FIXTURES])
        sample = d
    elif kind == "ont":
        d = synth.ont(0, 120, seed=34).numpy()
        sample = d[: d.size // 2]
    elif kind == "outside":
        d = synth.random_fastq(3000, seed=11)
        sample = synth.illumina(0, 600, seed=5, profile="hiseq").numpy()
    else:
        d = synth.illumina(0, 6000, seed=36, profile=kind).numpy()
        sample = d[: d.size // 3]
    recs, used = oracle.parse_records(sample)
    fs, fq = oracle.make_ft(*oracle.hist(sample[:used], recs))
    h.load_tables(fs, fq)
    check_chunks(oracle, h, d, R, fs, fq)
    h.close()


@pytest.mark.parametrize("eof,parts,lanes,kind", [(True, 4, 3, "illumina"), (False, 4, 2, "illumina"), (True, 2, 2, "illumina"),
                                                  (False, 7, 4, "illumina"), (True, 5, 1, "illumina"), (True, 8, 4, "ont"),
                                                  (False, 3, 3, "ont")])
def test_pipelined_host_compress(oracle, monkeypatch, capfd, eof, parts, lanes, kind):
    """Large host slabs are compressed as overlapped parts (copies on a copy
    stream, parts dealt to `lanes` handles, one host thread each): same
    chunks, streams and side arrays as the one-pass walk, with the tables from
    the sample inside fq28_compress and with pre-loaded tables."""
    import synth

    monkeypatch.setenv("FQ28_PIPE_MIN_MB", "1")      # both read at fq28_create
    monkeypatch.setenv("FQ28_PIPE_PARTS", str(parts))
    monkeypatch.setenv("FQ28_PIPE_LANES", str(lanes))
    monkeypatch.setenv("FQ28_PIPE_TRACE", "1")       # the timeline on stderr proves the parts were taken
    import fqcomp28_b200 as P

    H = P.Handle(0)
    if kind == "ont":   # reads of 1-50 kb: records straddle the part boundaries
        d = synth.ont(0, 2000, seed=33).numpy()
    else:
        d = synth.illumina_bytes(40 << 20, seed=31)[0].numpy()
    S, R = 4 << 20, 1 << 20
    sample = d[: int(oracle.split_chunks(d, S)[1])]
    recs, _ = oracle.parse_records(sample)
    fs, fq = oracle.make_ft(*oracle.hist(sample, recs))
    ft = (np.zeros(3076, np.uint8), np.zeros(1081348, np.uint8))
    capfd.readouterr()
    infos, summ, ar = H.compress(d, R, eof=eof, sample_bytes=S, ft_out=ft)
    assert f"fq28 pipe part {parts - 1}/{parts}" in capfd.readouterr().err, "the slab was not cut into the parts asked for"
    assert np.array_equal(ft[0], fs) and np.array_equal(ft[1], fq)
    assert int(summ.n_chunks) >= d.size // R - 2
    seq_a = ar["seq"][: int(summ.seq_bytes)].copy()
    # pre-loaded tables (sample_bytes == 0), chunk by chunk against the oracle + decode
    H.load_tables(fs, fq)
    infos2, summ2, ar2 = check_chunks(oracle, H, d, R, fs, fq, eof=eof)
    assert int(summ2.n_chunks) == int(summ.n_chunks) and int(summ2.consumed) == int(summ.consumed)
    assert np.array_equal(ar2["seq"][: int(summ2.seq_bytes)], seq_a)
    # and identical to the one-pass path
    monkeypatch.setenv("FQ28_PIPE_PARTS", "1")
    H1 = P.Handle(0)
    H1.load_tables(fs, fq)
    infos1, summ1, ar1 = H1.compress(d, R, eof=eof)
    assert int(summ1.n_chunks) == int(summ.n_chunks)
    for k in range(int(summ1.n_chunks)):
        for f in ("fastq_off", "total", "n_records", "rec_off", "seq_off", "qual_off", "seq_len", "qual_len",
                  "n_pos_off", "n_pos_len", "hdr_bytes", "hdr_off"):
            assert getattr(infos1[k], f) == getattr(infos2[k], f), (k, f)
    for key in ("seq", "qual"):
        nb = int(getattr(summ1, key + "_bytes"))
        assert np.array_equal(ar1[key][:nb], ar2[key][:nb]), key
    for key, n in (("readlens", "n_records"), ("n_count", "n_records"), ("n_pos", "n_pos_entries")):
        nn = int(getattr(summ1, n))
        assert np.array_equal(ar1[key][:nn], ar2[key][:nn]), key
    H.close()
    H1.close()


def _runny_fastq(n_records, seed, levels, p_stay, min_len=40, max_len=260):
    """Qualities with long runs of several dominant levels (two-state chains per
    level) and a tail of rare levels: every dominant-chain / zero-bit-run path."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    rare = np.array([2, 7, 11, 14, 19, 23, 25, 29, 33], dtype=np.uint8)
    for i in range(n_records):
        L = int(rng.integers(min_len, max_len + 1))
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L)
        q = np.empty(L, np.uint8)
        cur = levels[int(rng.integers(0, len(levels)))]
        for j in range(L):
            u = rng.random()
            if u > p_stay:
                cur = levels[int(rng.integers(0, len(levels)))] if u > p_stay + (1 - p_stay) * 0.5 else cur
                q[j] = rare[int(rng.integers(0, len(rare)))] if u <= p_stay + (1 - p_stay) * 0.5 else cur
            else:
                q[j] = cur
        out += b"@x%d r:%d\n" % (i, int(rng.integers(0, 999)))
        out += seq.tobytes() + b"\n+\n" + (q + 33).tobytes() + b"\n"
    return np.frombuffer(bytes(out), dtype=np.uint8)


@pytest.mark.parametrize("levels,p_stay", [((37,), 0.93), ((37, 2), 0.95), ((37, 25, 11, 2, 30), 0.97)])
def test_dominant_runs_many_levels(oracle, H, levels, p_stay):
    """Long dominant-symbol chains whose other symbols go beyond the three with a
    step table of their own (generic ops of k_chain_dom), several self-loop
    contexts with zero-bit run tables, variable read lengths (runs cut by record
    ends): chunk streams identical to the oracle, exact round trip."""
    d = _runny_fastq(9000, 41 + len(levels), levels, p_stay)
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    gcs, gcq = H.hist(d)
    assert np.array_equal(gcs, cs) and np.array_equal(gcq, cq)
    H.build_tables(gcs, gcq)
    norm, logs = oracle.ft_norm(fq), oracle.ft_logs(fq)
    dom_ctx = [c for c in range(8192) if norm[c].max() > (1 << int(logs[c])) // 2]
    assert len(dom_ctx) >= len(levels)  # at least ctx(d,d,d) of every run level is dominant
    check_chunks(oracle, H, d, 1 << 20, fs, fq)
    check_chunks(oracle, H, d, 300000, fs, fq, eof=False)


def test_long_reads_ont_like(oracle, H):
    """BASELINE config 4: variable-length long reads (1-50 kb), ONT-like
    qualities (thousands of live quality contexts)."""
    import synth

    d = synth.ont(0, 120, seed=32).numpy()
    recs, cs, cq, fs, fq = oracle_tables(oracle, d)
    assert int(recs["len"].max()) > 20000 and int((cq.sum(1) > 0).sum()) > 1500
    gcs, gcq = H.hist(d)
    assert np.array_equal(gcs, cs) and np.array_equal(gcq, cq)
    gfs, gfq = H.build_tables(gcs, gcq)
    assert np.array_equal(gfs, fs) and np.array_equal(gfq, fq)
    check_chunks(oracle, H, d, 1 << 20, fs, fq)
    check_chunks(oracle, H, d, 256 << 20, fs, fq)
