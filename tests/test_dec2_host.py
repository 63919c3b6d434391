"""The product's per-stream decoders (fqcomp28_b200/csrc/fq28_dec2.cuh) compiled
for the CPU (tests/cpp/dec2_host.cu) and checked against the oracle's encoder:
the algorithm -- cached cells, deferred refresh with the STALE protocol, inline
homopolymer / zero-bit-run contexts, the slow path for quality values the sample
never showed -- is verified here without a GPU; the `-m gpu` tests then check the
device build of the same source through the C ABI.

Reference behaviour decoded: SequenceDecoder::decodeRecord
(src/fse_sequence.cpp:114-143), QualityDecoder::decodeRecord
(src/fse_quality.cpp:55-67), FSE_Decoder::startChunk/endChunk
(src/fse_common.hpp:130-141).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import FIXTURES, ROOT, load_fixture

SRC = os.path.join(ROOT, "tests", "cpp", "dec2_host.cu")
HDR = os.path.join(ROOT, "fqcomp28_b200", "csrc", "fq28_dec2.cuh")
SO = os.path.join(ROOT, "tests", "cpp", "libdec2_host.so")


@pytest.fixture(scope="module")
def D(oracle):
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(
            ["g++", "-x", "c++", "-std=c++17", "-O2", "-g", "-Wno-unknown-pragmas", "-shared", "-fPIC", SRC, "-o", SO,
             "-L" + os.path.join(ROOT, "oracle"), "-lfq28_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    L = C.CDLL(SO)
    vp, u32 = C.c_void_p, C.c_uint32
    L.dec2h_seq.argtypes = [vp, vp, u32, C.c_uint, vp, vp, u32, vp]
    L.dec2h_seq.restype = C.c_int
    L.dec2h_qual.argtypes = [vp, vp, u32, C.c_uint, vp, vp, u32, vp, vp]
    L.dec2h_qual.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data


def roundtrip(O, D, d, fs, fq, misalign=0, expect_stats=None):
    """oracle-encode the slab as ONE chunk, decode both streams with the host build of
    the product decoders, compare with the input (N -> A in the sequence slots)."""
    recs, used = O.parse_records(d)
    body = d[:used]
    enc = O.Codec(fs, fq).encode_chunk(body, recs)
    n = len(recs)
    readlens = np.ascontiguousarray(enc["readlens"], dtype="<u2")
    hdr_lens = recs["hdr_len"].astype("<u2")
    out = np.zeros(body.size, dtype=np.uint8)
    seq = np.ascontiguousarray(enc["seq"])
    qual = np.ascontiguousarray(enc["qual"])
    assert D.dec2h_seq(_p(fs), _p(seq), seq.size, misalign, _p(readlens), _p(hdr_lens), n, _p(out)) == 0, "seq stream not consumed"
    stats = np.zeros(2, dtype=np.uint32)
    assert D.dec2h_qual(_p(fq), _p(qual), qual.size, misalign, _p(readlens), _p(hdr_lens), n, _p(out), _p(stats)) == 0, "qual stream not consumed"
    for r in recs:
        L = int(r["len"])
        s0, q0 = int(r["seq_off"]), int(r["qual_off"])
        want = body[s0 : s0 + L].copy()
        want[want == ord("N")] = ord("A")
        assert np.array_equal(out[s0 : s0 + L], want), "sequence differs"
        assert np.array_equal(out[q0 : q0 + L], body[q0 : q0 + L]), "quality differs"
    if expect_stats:
        expect_stats(stats)
    return stats


def tables_of(O, d):
    recs, _ = O.parse_records(d)
    cs, cq = O.hist(d, recs)
    return O.make_ft(cs, cq)


@pytest.mark.parametrize("name", FIXTURES)
def test_fixtures(oracle, D, name):
    d = load_fixture(name)
    fs, fq = tables_of(oracle, d)
    for mis in (0, 3):
        roundtrip(oracle, D, d, fs, fq, mis)


def test_every_misalignment(oracle, D):
    d = load_fixture("SRR065390_1_first5")
    fs, fq = tables_of(oracle, d)
    for mis in range(8):
        roundtrip(oracle, D, d, fs, fq, mis)


@pytest.mark.parametrize("profile", ["novaseq", "hiseq"])
def test_synthetic_illumina(oracle, D, profile):
    import synth

    d = synth.illumina(0, 3000, profile=profile).numpy()
    fs, fq = tables_of(oracle, d[: d.size // 3 * 2])  # tables from a leading sample
    stats = roundtrip(oracle, D, d, fs, fq)
    if profile == "novaseq":
        assert stats[1] >= 1, "binned qualities must get a zero-bit run slot"


def test_run_tables_ignored(oracle, D):
    """The tables carry zero-bit run slots but the decoder is told there are none (what the
    product does when the 16 KB run tables would cost a launch wave): the run contexts are
    then ordinary contexts, same bytes."""
    import synth

    D.dec2h_set_no_runs.argtypes = [C.c_int]
    D.dec2h_set_no_runs(1)
    try:
        d = synth.illumina(0, 3000, profile="novaseq").numpy()
        fs, fq = tables_of(oracle, d[: d.size // 3 * 2])
        stats = roundtrip(oracle, D, d, fs, fq)
        assert stats[1] >= 1, "the tables must have run slots for this test to mean anything"
        d = load_fixture("SRR065390_sub_1")
        fs, fq = tables_of(oracle, d)
        roundtrip(oracle, D, d, fs, fq, 5)
    finally:
        D.dec2h_set_no_runs(0)


def test_windowed_layout(oracle, D):
    """decode_qual_stream<WIN>: the cached cells of every row cover only the columns its touched
    contexts span; everything else (values outside V, contexts outside a row's window) takes the
    slow path.  On everything the dense layout is tested on, plus tables from a narrow sample
    against data that leaves the windows all the time."""
    import synth

    D.dec2h_set_win.argtypes = [C.c_int]
    D.dec2h_set_win(1)
    try:
        for name in FIXTURES:
            d = load_fixture(name)
            fs, fq = tables_of(oracle, d)
            for mis in (0, 3):
                roundtrip(oracle, D, d, fs, fq, mis)
        for profile in ("novaseq", "hiseq"):
            d = synth.illumina(0, 3000, profile=profile).numpy()
            fs, fq = tables_of(oracle, d[: d.size // 3 * 2])
            st = roundtrip(oracle, D, d, fs, fq)
            if profile == "hiseq":
                assert st[1] * 4 < st[0] * 512 // 2, "the windows must be much smaller than the dense layout"
        d = synth.ont(0, 40).numpy()
        fs, fq = tables_of(oracle, d)
        roundtrip(oracle, D, d, fs, fq, 1)
        for seed in range(4):
            d = synth.random_fastq(300, seed=seed)
            fs, fq = tables_of(oracle, d)
            roundtrip(oracle, D, d, fs, fq, seed)
        # windows from a narrow sample (HiSeq-like, values in runs), data = uniformly random qualities:
        # nearly every context is outside its row's window or outside V
        d = synth.random_fastq(400, seed=9)
        fs, fq = tables_of(oracle, synth.illumina(0, 300, profile="hiseq").numpy())
        roundtrip(oracle, D, d, fs, fq)
        fs, fq = tables_of(oracle, synth.illumina(0, 300, profile="novaseq").numpy())
        roundtrip(oracle, D, d, fs, fq)
        # ... and the other way round: wide tables, narrow data
        fs, fq = tables_of(oracle, synth.random_fastq(400, seed=9))
        roundtrip(oracle, D, synth.illumina(0, 800, profile="hiseq").numpy(), fs, fq)
    finally:
        D.dec2h_set_win(0)


def test_synthetic_ont(oracle, D):
    import synth

    d = synth.ont(0, 40).numpy()
    fs, fq = tables_of(oracle, d)
    roundtrip(oracle, D, d, fs, fq)


def test_random_fastq(oracle, D):
    import synth

    for seed in range(4):
        d = synth.random_fastq(300, seed=seed)
        fs, fq = tables_of(oracle, d)
        roundtrip(oracle, D, d, fs, fq, seed)


def test_quality_values_missing_from_the_sample(oracle, D):
    """Tables from records whose qualities use few values; the data then brings
    values (and whole contexts) the sample never showed: the slow path."""
    rng = np.random.default_rng(7)
    recs = []
    for i in range(400):
        L = int(rng.integers(3, 120))
        seq = rng.choice(list(b"ACGT"), L).astype(np.uint8).tobytes()
        if i < 200:
            q = rng.choice([35, 40, 60, 70], L, p=[0.1, 0.1, 0.2, 0.6])
        else:  # mostly the same values plus intruders, some in runs
            q = rng.choice([35, 40, 60, 70, 33, 50, 96], L, p=[0.1, 0.1, 0.2, 0.45, 0.05, 0.05, 0.05])
            if i % 7 == 0:
                q[L // 2 :] = 50
        recs.append(b"@r%d\n" % i + seq + b"\n+\n" + q.astype(np.uint8).tobytes() + b"\n")
    d = np.frombuffer(b"".join(recs), dtype=np.uint8)
    sample = np.frombuffer(b"".join(recs[:200]), dtype=np.uint8)
    fs, fq = tables_of(oracle, sample)
    stats = roundtrip(oracle, D, d, fs, fq)
    assert stats[0] == 5  # V = {0} + the four sampled values


def test_corrupt_streams_are_rejected_not_crashing(oracle, D):
    d = load_fixture("SRR065390_sub_1")
    fs, fq = tables_of(oracle, d)
    recs, used = oracle.parse_records(d)
    enc = oracle.Codec(fs, fq).encode_chunk(d, recs)
    n = len(recs)
    readlens = np.ascontiguousarray(enc["readlens"], dtype="<u2")
    hdr_lens = recs["hdr_len"].astype("<u2")
    out = np.zeros(d.size, dtype=np.uint8)
    rng = np.random.default_rng(1)
    bad = 0
    for trial in range(6):
        seq = np.array(enc["seq"], copy=True)
        qual = np.array(enc["qual"], copy=True)
        if trial == 0:
            seq, qual = seq[: seq.size // 2].copy(), qual[: qual.size // 2].copy()
        else:
            seq[rng.integers(0, seq.size, 20)] ^= 0x5A
            qual[rng.integers(0, qual.size, 20)] ^= 0x5A
        if seq[-1] == 0:
            seq[-1] = 1
        if qual[-1] == 0:
            qual[-1] = 1
        a = D.dec2h_seq(_p(fs), _p(seq), seq.size, 0, _p(readlens), _p(hdr_lens), n, _p(out))
        b = D.dec2h_qual(_p(fq), _p(qual), qual.size, 0, _p(readlens), _p(hdr_lens), n, _p(out), None)
        bad += (a != 0) + (b != 0)
    assert bad >= 2  # the truncated pair at least; flips inside symbol fields may go unnoticed


def test_corrupt_streams_under_asan(oracle, tmp_path):
    """The decoders on corrupted streams (bit flips, overwritten tails, truncations, random runs),
    compiled with AddressSanitizer + UBSan: whatever the bits say, no access outside the
    shared-memory image, the tables or the chunk's output.  (compute-sanitizer is not available
    on the GPU pool; fq28_dec2.cuh is the same source for the device.)"""
    import struct
    import synth

    exe = str(tmp_path / "dec2_fuzz")
    r = subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        "-Wno-unknown-pragmas", "-DDEC2_FUZZ_MAIN", SRC, "-o", exe, "-L" + os.path.join(ROOT, "oracle"),
                        "-lfq28_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")], capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("no AddressSanitizer runtime in this toolchain")
    assert r.returncode == 0, r.stderr[-2000:]
    cases = {"fixture": (load_fixture("SRR065390_sub_1"), 12), "hiseq": (synth.illumina(0, 250, profile="hiseq").numpy(), 60),
             "novaseq": (synth.illumina(0, 250, profile="novaseq").numpy(), 60)}
    for name, (d, iters) in cases.items():
        fs, fq = tables_of(oracle, d)
        recs, used = oracle.parse_records(d)
        body = d[:used]
        enc = oracle.Codec(fs, fq).encode_chunk(body, recs)
        readlens = np.ascontiguousarray(enc["readlens"], dtype="<u2")
        hdr_lens = recs["hdr_len"].astype("<u2")
        for kind, ft, stream in ((0, fs, enc["seq"]), (1, fq, enc["qual"])):
            blob = tmp_path / f"{name}_{kind}.bin"
            with open(blob, "wb") as f:
                f.write(struct.pack("<6I", 0x46513238, kind, ft.size, stream.size, len(recs), body.size))
                f.write(ft.tobytes())
                f.write(np.ascontiguousarray(stream).tobytes())
                f.write(readlens.tobytes())
                f.write(hdr_lens.tobytes())
            env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
            p = subprocess.run([exe, str(blob), str(iters)], capture_output=True, text=True, env=env, timeout=600)
            assert p.returncode == 0 and "Sanitizer" not in p.stderr and "runtime error" not in p.stderr, (name, kind, p.stderr[-3000:])
            assert "fuzz:" in p.stdout
