"""Pins the CPU oracle on the reference's four fixtures.

Digests are SURVEY.md Appendix C: produced by an independent Python
restatement of the reference path whose FSE primitives were validated against
libzstd 1.5.5.  Whole file = one chunk, sample = whole file (what
`fqcomp28 c -S 128 -R 256` does on these inputs, src/prepare.cpp:42-47).
"""
import hashlib

import numpy as np
import pytest

from conftest import FIXTURES, load_fixture


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


GOLD = {
    # name: (n_rec, seq_len, seq_sha, ft_seq_sha, qual_len, qual_sha, ft_qual_sha)
    "SRR065390_1_first5": (5, 363, "31939fa15a068c7a", "567b9f6382adcba9", 7305, "722a51f4fe596f1f", "2ae4b41c7aa59b5b"),
    "SRR065390_sub_1": (1000, 23212, "021d231cebfdbe66", "b7fe2ecf7314511b", 37539, "ce1a01574452d9ae", "bf90d83a9ac479d7"),
    "SRR065390_sub_2": (1000, 23608, "25145605abc9208a", "70d89608519647ef", 36950, "5c8079cefb59b3cc", "01daea2da64d3a2a"),
    "without_ns": (851, 20237, "0b002ca7615b4bb3", "240d479eb3511240", 35055, "c2b011b0b93269bd", "aac69be4cbd94637"),
}
SIDE = {
    # name: (n_count bytes, sha, n_pos bytes, sha, readlens sha)
    "SRR065390_1_first5": (10, "1328b35937974524", 702, "070dffeb56370567", "829a0bd4c19bf8a9"),
    "SRR065390_sub_1": (2000, "0235eb3c9e8b3009", 12526, "06402edb69ab8f86", "f09a0227b1309fa1"),
    "SRR065390_sub_2": (2000, "4c18862b4598b6ac", 3200, "af1aa315fe9c2c05", None),
    "without_ns": (1702, "75022da29e2afddd", 0, None, "eb71cbd9d48cf3d2"),
}
LOGS = {
    "SRR065390_1_first5": ({5: 98, 11: 158}, {7: 8192}),
    "SRR065390_sub_1": ({5: 109, 6: 101, 7: 38, 8: 6, 9: 2}, {7: 8180, 8: 6, 9: 2, 10: 1, 11: 3}),
    "SRR065390_sub_2": ({5: 102, 6: 102, 7: 42, 8: 8, 9: 2}, {7: 8180, 8: 6, 9: 3, 11: 3}),
    "without_ns": ({5: 123, 6: 93, 7: 33, 8: 5, 9: 2}, {7: 8182, 8: 5, 9: 2, 11: 3}),
}


def encode_whole(O, name):
    d = load_fixture(name)
    recs, used = O.parse_records(d)
    assert used == d.size
    cs, cq = O.hist(d, recs)
    fs, fq = O.make_ft(cs, cq)
    cod = O.Codec(fs, fq)
    return d, recs, fs, fq, cod, cod.encode_chunk(d, recs)


@pytest.mark.parametrize("name", FIXTURES)
def test_golden_digests(oracle, name):
    O = oracle
    d, recs, fs, fq, cod, enc = encode_whole(O, name)
    g = GOLD[name]
    assert len(recs) == g[0]
    assert (enc["seq"].size, sha(enc["seq"])) == (g[1], g[2])
    assert sha(fs) == g[3]
    assert (enc["qual"].size, sha(enc["qual"])) == (g[4], g[5])
    assert sha(fq) == g[6]
    s = SIDE[name]
    assert enc["n_count"].nbytes == s[0] and sha(enc["n_count"]) == s[1]
    assert enc["n_pos"].nbytes == s[2]
    if s[3]:
        assert sha(enc["n_pos"]) == s[3]
    if s[4]:
        assert sha(enc["readlens"]) == s[4]
    ls, lq = LOGS[name]
    assert dict(zip(*[x.tolist() for x in np.unique(O.ft_logs(fs), return_counts=True)])) == ls
    assert dict(zip(*[x.tolist() for x in np.unique(O.ft_logs(fq), return_counts=True)])) == lq


@pytest.mark.parametrize("name", FIXTURES)
def test_roundtrip(oracle, name):
    """test/workspace_test.cpp:45-69 / fse_sequence_test.cpp / fse_quality_test.cpp."""
    O = oracle
    d, recs, fs, fq, cod, enc = encode_whole(O, name)
    hdr, hl = O.gather_headers(d, recs)
    out = cod.decode_chunk(enc, hdr, hl, d.size)
    assert np.array_equal(out, d)


def test_ft_struct_sizes(oracle):
    """sizeof(FreqTable<256,4>) / sizeof(FreqTable<8192,64>), src/fse_common.hpp:147-174."""
    assert oracle.FT_SEQ_BYTES == 256 * 4 * 2 + 256 * 4 + 4 == 3076
    assert oracle.FT_QUAL_BYTES == 8192 * 64 * 2 + 8192 * 4 + 4 == 1081348


def test_chunk_boundaries_every_block_size(oracle):
    """test/fastq_io_test.cpp:15-53: for every block size 1000..filesize the
    chunks tile the 5-record file exactly and each ends on a record end."""
    O = oracle
    d = load_fixture("SRR065390_1_first5")
    recs, _ = O.parse_records(d)
    ends = set((recs["qual_off"] + recs["len"] + 1).tolist())
    for R in range(1000, d.size + 1):
        offs = O.split_chunks(d, R)
        assert offs[0] == 0 and offs[-1] == d.size
        assert all(int(o) in ends for o in offs[1:])
        assert all(0 < int(b) - int(a) <= R for a, b in zip(offs[:-1], offs[1:]))
        # greedy rule: the next record would not have fitted in the window
        for a, b in zip(offs[:-1], offs[1:]):
            nxt = [e for e in ends if e > b]
            if nxt:
                assert min(nxt) - int(a) > R


def test_trailing_partial_record_dropped(oracle):
    """src/fastq_io.cpp:31-32,63: a final record without '\\n' is never emitted."""
    O = oracle
    d = load_fixture("SRR065390_1_first5")
    cut = d[:-1]
    offs = O.split_chunks(cut, 1 << 20)
    recs, _ = O.parse_records(d)
    assert int(offs[-1]) == int(recs["hdr_off"][-1])


def test_seq_ctx_formulation(oracle):
    """The oracle codes base i in context (b[i-1],b[i-2],b[i-3],b[i-4]) over the
    virtual prefix TCCT; src/fse_sequence.cpp:53-112 builds the same contexts
    with two loops.  Restate the reference loops literally and compare."""
    rng = np.random.default_rng(7)
    b2b = {"A": 0, "C": 1, "G": 2, "T": 3}
    REV = "TCCTCCCACCTC"
    for L in list(range(1, 12)) + [50]:
        s = "".join(rng.choice(list("ACGT"), L))
        # --- literal restatement of encodeRecord's context walk
        ctx = 0xD7
        add = lambda c, ch: ((c << 2) + b2b[ch]) & 0xFF
        to_add = min(3, L - 1)
        p = L - 2
        for _ in range(to_add):
            ctx = add(ctx, s[p]); p -= 1
        first_partial = 3 if L > 4 else L - 1
        ref = {}
        pos = L - 1
        while pos > first_partial:
            ctx = add(ctx, s[pos - 4]); ref[pos] = ctx; pos -= 1
        i = 0
        for i in range(3 - to_add):
            ctx = add(ctx, REV[i])
        i = 3 - to_add
        pos = first_partial
        while pos >= 0:
            ctx = add(ctx, REV[i]); ref[pos] = ctx; pos -= 1; i += 1
        # --- oracle formulation
        pre = "TCCT"  # b[-4..-1] = T,C,C,T
        ext = pre + s
        for pos in range(L):
            j = pos + 4
            mine = (b2b[ext[j - 1]] << 6) | (b2b[ext[j - 2]] << 4) | (b2b[ext[j - 3]] << 2) | b2b[ext[j - 4]]
            assert mine == ref[pos], (L, pos)
        assert ref[0] == 0xD7


def test_short_read_rejected(oracle):
    """Reads shorter than 3 are UB in the reference (src/fse_quality.cpp:42-52); rejected."""
    O = oracle
    d = np.frombuffer(b"@r1\nAC\n+\n!!\n", dtype=np.uint8)
    recs, _ = O.parse_records(d)
    fs, fq = O.make_ft(*O.hist(d, recs))
    with pytest.raises(O.OracleError) as e:
        O.Codec(fs, fq).encode_chunk(d, recs)
    assert e.value.code == -4


def test_bad_alphabet_rejected(oracle):
    O = oracle
    d = np.frombuffer(b"@r1\nACGX\n+\n!!!!\n", dtype=np.uint8)
    recs, _ = O.parse_records(d)
    with pytest.raises(O.OracleError):
        O.hist(d, recs)


def test_random_fastq_roundtrip_and_stream_invariants(oracle):
    """Property checks on adversarial random FASTQ (variable lengths from 3, N
    runs, every quality level): the oracle's chunk streams decode back to the
    chunk bytes at several chunk sizes, every stream ends with a non-zero byte
    (BIT_closeCStream's end mark), and side arrays have one entry per record."""
    import synth

    for seed, levels in ((3, 64), (4, 8), (5, 41)):
        d = synth.random_fastq(400, seed=seed, min_len=3, max_len=260, qual_levels=levels)
        recs, used = oracle.parse_records(d)
        assert used == d.size
        fs, fq = oracle.make_ft(*oracle.hist(d, recs))
        cod = oracle.Codec(fs, fq)
        for R in (d.size, 20000, 3000):
            offs = oracle.split_chunks(d, R)
            assert int(offs[0]) == 0 and int(offs[-1]) == d.size and np.all(np.diff(offs.astype(np.int64)) > 0)
            for a, b in zip(offs[:-1], offs[1:]):
                sub = d[int(a) : int(b)]
                r, u = oracle.parse_records(sub)
                assert u == sub.size
                enc = cod.encode_chunk(sub, r)
                assert enc["seq"][-1] != 0 and enc["qual"][-1] != 0
                assert enc["readlens"].size == len(r) and enc["n_count"].size == len(r)
                assert int(enc["n_count"].astype(np.int64).sum()) == enc["n_pos"].size
                hdr, hl = oracle.gather_headers(sub, r)
                out = cod.decode_chunk(enc, hdr, hl, sub.size)
                assert np.array_equal(out, sub)
