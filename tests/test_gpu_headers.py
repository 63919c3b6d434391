"""Row N2: the GPU header tokeniser / detokeniser (fq28_tokenize_headers,
fq28_detokenize_headers) against Python restatements of encodeHeader / decodeHeader
(src/workspace.cpp:95-157) with storeString / storeNumeric / loadNextString /
loadNextNumeric (src/headers.cpp:75-133)."""
import struct

import numpy as np
import pytest

from conftest import FIXTURES, load_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    import fqcomp28_b200

    h = fqcomp28_b200.Handle(0)
    yield h
    h.close()


def from_chars_i32(b: bytes) -> int:
    """std::from_chars<int32_t> into a value preset to 0."""
    i, neg = 0, False
    if i < len(b) and b[i : i + 1] == b"-":
        neg, i = True, 1
    j = i
    while j < len(b) and 48 <= b[j] <= 57:
        j += 1
    if j == i:
        return 0
    v = int(b[i:j])
    v = -v if neg else v
    return v if -(1 << 31) <= v < (1 << 31) else 0


def ref_tokenize(lines, chunk_rec, types, seps, first):
    """-> [[{flag, content, clen} per field] per chunk]"""
    F = len(types)
    first_vals = [f if types[i] else from_chars_i32(f) for i, f in enumerate(first)]
    out = []
    for k in range(len(chunk_rec) - 1):
        prev = list(first_vals)  # startNewChunk, src/workspace.cpp:90-93
        st = [{"flag": bytearray(), "content": bytearray(), "clen": bytearray()} for _ in range(F)]
        for h in lines[chunk_rec[k] : chunk_rec[k + 1]]:
            p, end = 1, len(h)
            for i in range(F):
                if i + 1 < F:
                    e = h.find(bytes([seps[i]]), min(p + 1, end), end)
                    e = end if e < 0 else e
                else:
                    e = end
                val = h[p:e]
                if types[i]:
                    if val == prev[i]:
                        st[i]["flag"].append(0)
                    else:
                        assert len(val) < 255
                        st[i]["flag"].append(1)
                        st[i]["content"] += val
                        st[i]["clen"].append(len(val))
                        prev[i] = val
                else:
                    v = from_chars_i32(val)
                    st[i]["content"] += struct.pack("<I", (v - prev[i]) & 0xFFFFFFFF)
                    prev[i] = v
                p = e + 1 if e < end else end
        out.append([{n: bytes(b) for n, b in f.items()} for f in st])
    return out


def ref_detokenize(streams, n_per_chunk, types, seps, first):
    """decodeHeader (src/workspace.cpp:127-157) -> list of header lines"""
    F = len(types)
    first_vals = [f if types[i] else from_chars_i32(f) for i, f in enumerate(first)]
    lines = []
    for k, n in enumerate(n_per_chunk):
        prev = list(first_vals)
        dpos, cpos, lpos = [0] * F, [0] * F, [0] * F
        for _ in range(n):
            line = b"@"
            for i in range(F):
                st = streams[k][i]
                if types[i]:
                    if st["flag"][dpos[i]]:
                        ln = st["clen"][lpos[i]]
                        lpos[i] += 1
                        prev[i] = st["content"][cpos[i] : cpos[i] + ln]
                        cpos[i] += ln
                    dpos[i] += 1
                    line += prev[i]
                else:
                    (delta,) = struct.unpack_from("<i", st["content"], cpos[i])
                    cpos[i] += 4
                    v = (prev[i] + delta + (1 << 31)) % (1 << 32) - (1 << 31)
                    prev[i] = v
                    line += str(v).encode()
                if i + 1 < F:
                    line += bytes([seps[i]])
            lines.append(line)
    return lines


def run_case(H, lines, chunk_rec, canonical=False):
    raw = np.frombuffer(b"".join(lines), dtype=np.uint8)
    lens = np.array([len(l) for l in lines], dtype=np.uint16)
    fmt, got = H.tokenize_headers(raw, lens, chunk_rec, lines[0])
    want = ref_tokenize(lines, chunk_rec, fmt["types"], fmt["separators"], fmt["first"])
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        for i, (gf, wf) in enumerate(zip(g, w)):
            for name in ("flag", "content", "clen"):
                assert gf[name] == wf[name], (k, i, name)
    # decode side: GPU detokeniser == restated decodeHeader on the same streams
    back, blens = H.detokenize_headers()
    n_per_chunk = [chunk_rec[k + 1] - chunk_rec[k] for k in range(len(chunk_rec) - 1)]
    ref_lines = ref_detokenize(want, n_per_chunk, fmt["types"], fmt["separators"], fmt["first"])
    assert [int(x) for x in blens] == [len(l) for l in ref_lines]
    assert back.tobytes() == b"".join(ref_lines)
    if canonical:  # well-formed headers survive the round trip
        assert ref_lines == lines
    return fmt


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_headers(oracle, H, name):
    d = load_fixture(name)
    recs, _ = oracle.parse_records(d)
    lines = [bytes(d[int(r["hdr_off"]) : int(r["hdr_off"]) + int(r["hdr_len"])]) for r in recs]
    n = len(lines)
    run_case(H, lines, [0, n])
    if n >= 4:
        run_case(H, lines, [0, 1, n // 2, n - 1, n])  # every chunk restarts from the first header


def test_reference_header_format_literals(H):
    """test/headers_test.cpp:12-33: field types / separators of the two literal headers, as the
    format handed to fq28_tokenize_headers, and a tokenise -> detokenise round trip on records
    of each shape."""
    cases = [
        (b"@SRR22543904.1 1 length=150", [1, 0, 0, 1, 0], b".  ="),
        (b"@SRR065390.1000 HWUSI-EAS687_61DAJ:8:1:1174:9158 length=100", [1, 0, 1, 1, 1, 0, 0, 0, 0, 1, 0], b". -_:::: ="),
    ]
    for first, types, seps in cases:
        lines = [first]
        for i in range(2, 40):
            if len(types) == 5:
                lines.append(b"@SRR22543904.%d %d length=%d" % (i, i, 150 - (i % 3)))
            else:
                lines.append(b"@SRR065390.%d HWUSI-EAS687_61DAJ:8:%d:%d:%d length=100" % (999 + i, 1 + i // 20, 1174 + 7 * i, 9158 - i))
        fmt = run_case(H, lines, [0, 10, len(lines)], canonical=True)
        assert list(fmt["types"]) == types and bytes(fmt["separators"]) == seps


def test_synthetic_illumina_headers(oracle, H):
    import synth

    d = synth.illumina(0, 30000, seed=30).numpy()
    recs, _ = oracle.parse_records(d)
    lines = [bytes(d[int(r["hdr_off"]) : int(r["hdr_off"]) + int(r["hdr_len"])]) for r in recs]
    offs = oracle.split_chunks(d, 1 << 20)
    ends = recs["hdr_off"].astype(np.int64)
    chunk_rec = [int(np.searchsorted(ends, int(o), side="left")) for o in offs[:-1]] + [len(lines)]
    fmt = run_case(H, lines, chunk_rec, canonical=True)
    assert len(fmt["types"]) >= 8 and 0 in fmt["types"] and 1 in fmt["types"]


def test_adversarial_headers(H):
    rng = np.random.default_rng(5)
    first = b"@ab12.77 x:0042:-7_q/1"
    lines = [first]
    alnum = b"abcXYZ0189"
    for _ in range(4000):
        parts = []
        for _f in range(int(rng.integers(1, 9))):  # fewer or more fields than the format
            kind = int(rng.integers(0, 6))
            if kind == 0:
                parts.append(str(int(rng.integers(-5, 5000))).encode())
            elif kind == 1:
                parts.append(bytes(rng.choice(np.frombuffer(alnum, np.uint8), int(rng.integers(0, 12)))))
            elif kind == 2:
                parts.append(str(int(rng.integers(2**31 - 3, 2**31 + 3))).encode())  # around INT32_MAX
            elif kind == 3:
                parts.append(b"-" + str(int(rng.integers(2**31 - 2, 2**31 + 3))).encode())
            elif kind == 4:
                parts.append(b"12ab" if rng.random() < 0.5 else b"-")
            else:
                parts.append(lines[-1].split(b":")[0][1:7])  # repeats of earlier values
        seps = [bytes([c]) for c in b". :_/:"]
        h = b"@"
        for j, p in enumerate(parts):
            h += p + (seps[j % len(seps)] if j + 1 < len(parts) else b"")
        if not h[-1:].isalnum():
            h += b"z"
        lines.append(h)
    lines.append(b"@" + b"q" * 254 + b".1 x:1:1_q/1")  # longest legal STRING value
    n = len(lines)
    run_case(H, lines, [0, n])
    run_case(H, lines, [0, 7, 8, 1000, 1001, n])


def test_string_value_too_long(H):
    import fqcomp28_b200

    lines = [b"@ab.1", b"@" + b"q" * 255 + b".2"]
    raw = np.frombuffer(b"".join(lines), dtype=np.uint8)
    lens = np.array([len(l) for l in lines], dtype=np.uint16)
    with pytest.raises(fqcomp28_b200.Fq28Error) as e:
        H.tokenize_headers(raw, lens, [0, 2], lines[0])
    assert e.value.code == -2  # FQ28_ERR_FORMAT


def test_detokenize_rejects_short_streams(H):
    import fqcomp28_b200

    lines = [b"@ab.1", b"@ab.2", b"@cd.3"]
    raw = np.frombuffer(b"".join(lines), dtype=np.uint8)
    lens = np.array([len(l) for l in lines], dtype=np.uint16)
    H.tokenize_headers(raw, lens, [0, 3], lines[0])
    fmt, strings, arena, infos, cr = H._hdr_raw
    infos[1].content_len -= 4  # numeric field one record short
    with pytest.raises(fqcomp28_b200.Fq28Error) as e:
        H.detokenize_headers((fmt, strings, arena, infos, cr))
    assert e.value.code == -8  # FQ28_ERR_STREAM
