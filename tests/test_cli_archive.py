"""Row N1: the drop-in `fqcomp28 c|d` CLI (fqcomp28_b200/host/fqcomp28_cli.cpp).

* `c` then `d` restores the input (scripts/check_compression_integrity.sh:21-24);
* the archive follows the reference's framing (src/archive.cpp:22-163,
  src/archive.h:10-29): u32 n_blocks | meta (u16 hlen, first header, raw
  FreqTable images) | blocks | index of {i64 offset; u32 idx; pad}; the meta
  tables and every block's seq / qual streams equal the CPU oracle's.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import DATA, FIXTURES, ROOT, load_fixture

CLI = os.path.join(ROOT, "fqcomp28_b200", "fqcomp28")


def build_cli():
    pkg = os.path.join(ROOT, "fqcomp28_b200")
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", "-pthread", "-o", CLI, os.path.join(pkg, "host", "fqcomp28_cli.cpp"),
                           "-L", pkg, "-lfq28", "-Wl,-rpath,$ORIGIN"])


def test_cli_builds():
    build_cli()
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 106 and "fqcomp28 c" in r.stderr  # usage, like CLI11's RequiredError exit


# test/headers_test.cpp:12-33 ("HeaderFormatSpec"), the reference's only literal expectations
HEADER_FORMAT_CASES = [
    ("@SRR22543904.1 1 length=150", 5, "SNNSN", ". \x20="),
    ("@SRR065390.1000 HWUSI-EAS687_61DAJ:8:1:1174:9158 length=100", 11, "SNSSSNNNNSN", ". -_:::: ="),
]


@pytest.mark.parametrize("header,n,types,seps", HEADER_FORMAT_CASES)
def test_header_format_spec(header, n, types, seps):
    """Format::fromHeader of the CLI == the reference's REQUIREs (field types and separators)."""
    build_cli()
    seps = seps.replace("\x20", " ")
    r = subprocess.run([CLI, "header-format", header], capture_output=True, text=True, check=True)
    got = r.stdout.split("\n")
    assert (int(got[0]), got[1], got[2]) == (n, types, seps)
    bad = subprocess.run([CLI, "header-format", "@ends.in.separator."], capture_output=True, text=True)
    assert bad.returncode == 1 and "should end in alnum" in bad.stderr   # src/headers.cpp:64-66


def parse_archive(buf: bytes, n_fields_types):
    """-> (first_header, ft_seq, ft_qual, [blocks sorted by idx])"""
    (n_blocks,) = struct.unpack_from("<I", buf, 0)
    (hlen,) = struct.unpack_from("<H", buf, 4)
    pos = 6
    first = buf[pos : pos + hlen].decode()
    pos += hlen
    ft_seq = np.frombuffer(buf, np.uint8, 3076, pos); pos += 3076
    ft_qual = np.frombuffer(buf, np.uint8, 1081348, pos); pos += 1081348
    index = [struct.unpack_from("<qI4x", buf, len(buf) - 16 * n_blocks + 16 * i) for i in range(n_blocks)]
    blocks = []
    for off, idx in sorted(index, key=lambda t: t[1]):
        p = off
        total, n_rec = struct.unpack_from("<II", buf, p); p += 8
        side = []
        for _ in range(3):  # readlens, n_count, n_pos: u32 original, u32 size, bytes
            orig, sz = struct.unpack_from("<II", buf, p); p += 8
            side.append((orig, buf[p : p + sz])); p += sz
        (sz,) = struct.unpack_from("<I", buf, p); p += 4
        seq = buf[p : p + sz]; p += sz
        (sz,) = struct.unpack_from("<I", buf, p); p += 4
        qual = buf[p : p + sz]; p += sz
        blocks.append(dict(idx=idx, total=total, n_rec=n_rec, side=side, seq=seq, qual=qual))
    return first, ft_seq, ft_qual, blocks


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_cli_roundtrip_fixture(tmp_path, oracle, name):
    build_cli()
    src = os.path.join(DATA, name + ".fastq")
    arc, out = str(tmp_path / "a.fqz"), str(tmp_path / "o.fastq")
    subprocess.check_call([CLI, "c", "--input1", src, "-o", arc, "--threads", "4"])
    subprocess.check_call([CLI, "d", "--input", arc, "--o1", out, "--threads", "4"])
    assert open(out, "rb").read() == open(src, "rb").read()
    # archive framing + streams vs the oracle (default -S 128 -R 256: one chunk, sample = whole file)
    O = oracle
    d = load_fixture(name)
    recs, _ = O.parse_records(d)
    fs, fq = O.make_ft(*O.hist(d, recs))
    enc = O.Codec(fs, fq).encode_chunk(d, recs)
    first, a_fs, a_fq, blocks = parse_archive(open(arc, "rb").read(), None)
    assert first == bytes(d[: recs["hdr_len"][0]]).decode()
    assert np.array_equal(a_fs, fs) and np.array_equal(a_fq, fq)
    assert len(blocks) == 1 and blocks[0]["idx"] == 0
    b = blocks[0]
    assert (b["total"], b["n_rec"]) == (d.size, len(recs))
    assert b["seq"] == enc["seq"].tobytes() and b["qual"] == enc["qual"].tobytes()
    assert [s[0] for s in b["side"]] == [2 * len(recs), 2 * len(recs), 2 * enc["n_pos"].size]
    # STORED container: 28-byte header then the raw buffer
    assert b["side"][0][1][:8] == b"FQ28STOR" and b["side"][0][1][28:] == enc["readlens"].tobytes()
    assert b["side"][2][1][28:] == enc["n_pos"].tobytes()


@pytest.mark.gpu
def test_cli_multichunk_synthetic(tmp_path, oracle):
    import synth

    build_cli()
    d = synth.illumina(0, 12000, seed=30).numpy()  # ~4.2 MB
    src, arc, out = str(tmp_path / "s.fastq"), str(tmp_path / "s.fqz"), str(tmp_path / "s.out")
    d.tofile(src)
    subprocess.check_call([CLI, "c", "--i1", src, "--output", arc, "-R", "1", "-S", "2", "--slab-mb", "3"])
    subprocess.check_call([CLI, "d", "-i", arc, "--output1", out])
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), d)
    O = oracle
    offs = O.split_chunks(d, 1 << 20)
    sample = d[: int(O.split_chunks(d, 2 << 20)[1])]
    recs, _ = O.parse_records(sample)
    fs, fq = O.make_ft(*O.hist(sample, recs))
    first, a_fs, a_fq, blocks = parse_archive(open(arc, "rb").read(), None)
    assert np.array_equal(a_fs, fs) and np.array_equal(a_fq, fq)
    assert [b["idx"] for b in blocks] == list(range(len(offs) - 1))
    cod = O.Codec(fs, fq)
    for k, b in enumerate(blocks):
        sub = d[int(offs[k]) : int(offs[k + 1])]
        r, _ = O.parse_records(sub)
        e = cod.encode_chunk(sub, r)
        assert b["total"] == sub.size and b["seq"] == e["seq"].tobytes() and b["qual"] == e["qual"].tobytes()


@pytest.mark.gpu
def test_cli_report(tmp_path, oracle):
    """Row N4: the stderr table of `fqcomp28 c` (src/report.cpp:37-102) -- input
    sizes, stored stream sizes, ratios, block count -- and its invariant
    archive_size == file size (the reference's assert at :94-96)."""
    import synth

    build_cli()
    d = synth.illumina(0, 12000, seed=30).numpy()
    src, arc = str(tmp_path / "s.fastq"), str(tmp_path / "s.fqz")
    d.tofile(src)
    r = subprocess.run([CLI, "c", "--i1", src, "--output", arc, "-R", "1", "-S", "2", "--slab-mb", "3"],
                       capture_output=True, text=True, check=True)
    err = r.stderr
    assert "warning" not in err
    rows = {}
    for line in err.splitlines():
        f = line.split("\t")
        if len(f) >= 2 and f[0] not in rows:
            rows[f[0]] = f[1:]
    recs, _ = oracle.parse_records(d)
    n_sym = int(recs["len"].sum())
    assert int(rows["Sequence"][0]) == n_sym and int(rows["Headers"][0]) == int(recs["hdr_len"].sum())
    first, _, _, blocks = parse_archive(open(arc, "rb").read(), None)
    assert int(rows["# blocks: "][0]) == len(blocks)
    assert int(rows["seq"][0]) == sum(len(b["seq"]) for b in blocks)
    assert int(rows["qual"][0]) == sum(len(b["qual"]) for b in blocks)
    assert int(rows["readlens"][0]) == sum(len(b["side"][0][1]) for b in blocks)
    assert int(rows["meta_seq"][0]) == 3076 and int(rows["meta_qual"][0]) == 1081348
    size = os.path.getsize(arc)
    total_in = int(recs["hdr_len"].sum()) + 2 * n_sym + 5 * len(recs)
    # CR section comes after the stream table: the last "Total" row
    total_cr = [l.split("\t")[1] for l in err.splitlines() if l.startswith("Total\t")][-1]
    assert total_cr == f"{total_in / size:.3f}"


@pytest.mark.gpu
def test_cli_gpu_header_tokeniser_same_archive(tmp_path):
    """Row N2 (encode side): the archive written with the GPU header tokeniser is
    byte-identical to the one written with the host tokeniser."""
    import synth

    build_cli()
    d = synth.illumina(0, 12000, seed=33).numpy()
    src = str(tmp_path / "s.fastq")
    d.tofile(src)
    a1, a2, out = str(tmp_path / "gpu.fqz"), str(tmp_path / "host.fqz"), str(tmp_path / "o.fastq")
    subprocess.check_call([CLI, "c", "--i1", src, "-o", a1, "-R", "1", "-S", "2", "--slab-mb", "3"])
    subprocess.check_call([CLI, "c", "--i1", src, "-o", a2, "-R", "1", "-S", "2", "--slab-mb", "3", "--host-headers"])
    assert open(a1, "rb").read() == open(a2, "rb").read()
    subprocess.check_call([CLI, "d", "-i", a1, "--o1", out])
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), d)
    subprocess.check_call([CLI, "d", "-i", a1, "--o1", out, "--host-headers", "--slab-mb", "1"])
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), d)
    for name in FIXTURES:
        f = os.path.join(DATA, name + ".fastq")
        subprocess.check_call([CLI, "c", "--i1", f, "-o", a1])
        subprocess.check_call([CLI, "c", "--i1", f, "-o", a2, "--host-headers"])
        assert open(a1, "rb").read() == open(a2, "rb").read(), name
        subprocess.check_call([CLI, "d", "-i", a1, "--o1", out])
        assert open(out, "rb").read() == open(f, "rb").read(), name
