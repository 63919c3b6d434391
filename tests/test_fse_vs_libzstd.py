"""Pins the oracle's FSE primitives against the container's real libzstd.

fqcomp28 takes FSE_optimalTableLog / FSE_normalizeCount / FSE_buildCTable /
FSE_buildDTable / FSE_encodeSymbol / FSE_decodeSymbol and the bit stream from a
zstd fork that is not in the reference tree (cmake/Dependencies.cmake:21-27).
The system libzstd (1.5.5) hides those symbols, but every zstd FRAME embeds
their results:

  * Huffman-weight headers: an NCount header + a 2-state FSE stream produced
    by FSE_optimalTableLog(6, n, max) + FSE_normalizeCount(useLowProbCount=0)
    + FSE_buildCTable + FSE_compress_usingCTable.  We parse them, decode them
    with the oracle's DTable, re-derive table log + normalised counts with the
    oracle and RE-ENCODE the stream with the oracle's CTable: bytes must match.
  * Sequence sections: up to three NCount headers (literal-length, offset,
    match-length codes) produced by FSE_optimalTableLog(9|8|9, nbSeq, max) +
    FSE_normalizeCount(useLowProbCount = nbSeq >= 2048) -- the same
    useLowProbCount=1 / -1 low-probability / FSE_normalizeM2 paths fqcomp28
    exercises -- and a 3-state interleaved stream.  We decode the stream with
    the oracle's DTables (it must be consumed exactly), histogram the codes and
    re-derive table logs + normalised counts: they must equal the headers.

SURVEY.md Appendix B / B.1 documents the method and the frame-format crib.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import FIXTURES, load_fixture


# --------------------------------------------------------------------------- libzstd
def _zstd():
    for name in ("libzstd.so.1", "/usr/lib/x86_64-linux-gnu/libzstd.so.1"):
        try:
            z = C.CDLL(name)
            z.ZSTD_compress.restype = C.c_size_t
            z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
            z.ZSTD_compressBound.restype = C.c_size_t
            z.ZSTD_compressBound.argtypes = [C.c_size_t]
            z.ZSTD_isError.restype = C.c_uint
            z.ZSTD_isError.argtypes = [C.c_size_t]
            return z
        except OSError:
            continue
    return None


def zstd_compress(z, data: bytes, level: int) -> bytes:
    cap = z.ZSTD_compressBound(len(data))
    dst = C.create_string_buffer(cap)
    n = z.ZSTD_compress(dst, cap, data, len(data), level)
    assert not z.ZSTD_isError(n)
    return dst.raw[:n]


# --------------------------------------------------------------------------- bit readers
class FwdBits:
    """LSB-first forward reader (FSE_readNCount)."""

    def __init__(self, buf: bytes, pos: int):
        self.v = int.from_bytes(buf[pos : pos + 600], "little")
        self.bit = 0

    def peek(self, n):
        return (self.v >> self.bit) & ((1 << n) - 1)

    def skip(self, n):
        self.bit += n


class BackBits:
    """BIT_DStream_t: reads from the end mark downward (Appendix A.6)."""

    def __init__(self, buf: bytes):
        assert buf and buf[-1] != 0, "no end mark"
        self.v = int.from_bytes(buf, "little")
        self.pos = (len(buf) - 1) * 8 + buf[-1].bit_length() - 1

    def read(self, n):
        if n == 0:
            return 0
        self.pos -= n
        if self.pos < 0:  # reading past the start yields zeros (zstd semantics)
            return (self.v << (-self.pos)) & ((1 << n) - 1)
        return (self.v >> self.pos) & ((1 << n) - 1)


def read_ncount(buf: bytes, pos: int, max_sv: int):
    """FSE_readNCount -> (norm list, table_log, header bytes)."""
    br = FwdBits(buf, pos)
    table_log = br.peek(4) + 5
    br.skip(4)
    remaining = (1 << table_log) + 1
    threshold = 1 << table_log
    nb_bits = table_log + 1
    norm = []
    previous0 = False
    while remaining > 1 and len(norm) <= max_sv:
        if previous0:
            n0 = len(norm)
            while br.peek(16) == 0xFFFF:
                n0 += 24
                br.skip(16)
            while br.peek(2) == 3:
                n0 += 3
                br.skip(2)
            n0 += br.peek(2)
            br.skip(2)
            norm += [0] * (n0 - len(norm))
        mx = (2 * threshold - 1) - remaining
        if br.peek(nb_bits - 1) < mx:
            count = br.peek(nb_bits - 1)
            br.skip(nb_bits - 1)
        else:
            count = br.peek(nb_bits)
            if count >= threshold:
                count -= mx
            br.skip(nb_bits)
        count -= 1
        remaining -= abs(count)
        norm.append(count)
        previous0 = count == 0
        while remaining < threshold:
            nb_bits -= 1
            threshold >>= 1
    assert remaining == 1, "corrupt NCount"
    return norm, table_log, (br.bit + 7) >> 3


# --------------------------------------------------------------------------- oracle-backed tables
def dtable(O, norm, table_log):
    cells = O.build_dtable(np.array(norm, dtype=np.int16), table_log)
    return [(int(c) & 0xFFFF, (int(c) >> 16) & 0xFF, int(c) >> 24) for c in cells]  # (newState, sym, nb)


class CTab:
    def __init__(self, O, norm, table_log):
        self.st, self.dfs, self.dnb = O.build_ctable(np.array(norm, dtype=np.int16), table_log)
        self.log = table_log

    def init2(self, sym):  # FSE_initCState2
        dnb = int(self.dnb[sym])
        nb = ((dnb + (1 << 15)) & 0xFFFFFFFF) >> 16
        v = ((nb << 16) - dnb) & 0xFFFFFFFF
        return int(self.st[(v >> nb) + int(self.dfs[sym])])

    def encode(self, w, state, sym):  # FSE_encodeSymbol
        nb = ((state + int(self.dnb[sym])) & 0xFFFFFFFF) >> 16
        w.add(state, nb)
        return int(self.st[(state >> nb) + int(self.dfs[sym])])


class BitW:
    def __init__(self):
        self.v = 0
        self.n = 0

    def add(self, value, nb):
        self.v |= (value & ((1 << nb) - 1)) << self.n
        self.n += nb

    def close(self) -> bytes:  # BIT_closeCStream
        self.add(1, 1)
        return self.v.to_bytes((self.n + 7) // 8, "little")


# --------------------------------------------------------------------------- Huffman weights
def check_huf_weights(O, buf: bytes, stats):
    """buf = the FSE-compressed weight description (header byte < 128 already stripped)."""
    norm, tl, hsz = read_ncount(buf, 0, 12)
    stream = buf[hsz:]
    D = dtable(O, norm, tl)
    br = BackBits(stream)
    s1 = br.read(tl)
    s2 = br.read(tl)
    out = []
    while True:
        ns, sym, nb = D[s1]
        out.append(sym)
        if nb > br.pos:
            out.append(D[s2][1])
            break
        s1 = ns + br.read(nb)
        ns, sym, nb = D[s2]
        out.append(sym)
        if nb > br.pos:
            out.append(D[s1][1])
            break
        s2 = ns + br.read(nb)
    n = len(out)
    mx = max(out)
    count = np.bincount(out, minlength=mx + 1).astype(np.uint32)
    # HUF_compressWeights: FSE_optimalTableLog(6, wtSize, max) + normalize(useLowProbCount = 0)
    assert O.optimal_table_log(6, n, mx) == tl
    mine = O.normalize_count(count, tl, 0).tolist()
    assert mine == norm[: mx + 1] and not any(norm[mx + 1 :]), (mine, norm)
    # FSE_compress_usingCTable, 2 states
    ct = CTab(O, mine, tl)
    w = BitW()
    ip = n
    if n & 1:
        ip -= 1; st1 = ct.init2(out[ip])
        ip -= 1; st2 = ct.init2(out[ip])
        ip -= 1; st1 = ct.encode(w, st1, out[ip])
    else:
        ip -= 1; st2 = ct.init2(out[ip])
        ip -= 1; st1 = ct.init2(out[ip])
    while ip > 0:
        ip -= 1; st2 = ct.encode(w, st2, out[ip])
        if ip == 0:
            break
        ip -= 1; st1 = ct.encode(w, st1, out[ip])
    w.add(st2, tl)
    w.add(st1, tl)
    assert w.close() == stream, "re-encoded Huffman weight stream differs from libzstd's"
    stats["huf"] += 1
    stats["logs"].add(tl)


# --------------------------------------------------------------------------- sequences
LL_BITS = [0] * 16 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
ML_BITS = [0] * 32 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
LL_DEF = ([4, 3] + [2] * 11 + [1, 1, 1] + [2] * 9 + [3, 2, 1, 1, 1, 1, 1] + [-1] * 4, 6)
ML_DEF = ([1, 4, 3] + [2] * 6 + [1] * 37 + [-1] * 7, 6)
OF_DEF = ([1] * 6 + [2, 2, 2] + [1] * 15 + [-1] * 5, 5)
assert len(LL_DEF[0]) == 36 and len(ML_DEF[0]) == 53 and len(OF_DEF[0]) == 29
MAXSV = {"LL": 35, "OF": 31, "ML": 52}
FSELOG = {"LL": 9, "OF": 8, "ML": 9}
DEFAULT = {"LL": LL_DEF, "OF": OF_DEF, "ML": ML_DEF}


def check_sequences(O, blk: bytes, pos: int, prev: dict, stats):
    """Sequences section starting at blk[pos]."""
    if pos >= len(blk):
        return
    b0 = blk[pos]
    if b0 == 0:
        return
    if b0 < 128:
        nb_seq, pos = b0, pos + 1
    elif b0 < 255:
        nb_seq, pos = ((b0 - 128) << 8) + blk[pos + 1], pos + 2
    else:
        nb_seq, pos = blk[pos + 1] + (blk[pos + 2] << 8) + 0x7F00, pos + 3
    modes = blk[pos]
    pos += 1
    mode = {"LL": modes >> 6, "OF": (modes >> 4) & 3, "ML": (modes >> 2) & 3}
    tabs = {}
    parsed = {}
    for kind in ("LL", "OF", "ML"):
        m = mode[kind]
        if m == 0:
            tabs[kind] = DEFAULT[kind]
        elif m == 1:
            tabs[kind] = ("rle", blk[pos])
            pos += 1
        elif m == 2:
            norm, tl, hsz = read_ncount(blk, pos, MAXSV[kind])
            pos += hsz
            tabs[kind] = (norm, tl)
            parsed[kind] = (norm, tl)
        else:
            assert kind in prev, "repeat mode without a previous table"
            tabs[kind] = prev[kind]
        prev[kind] = tabs[kind]
    stream = blk[pos:]
    if not parsed:
        return
    # decode all sequences with the oracle's DTables
    D, LOG = {}, {}
    for kind, tb in tabs.items():
        if tb[0] == "rle":
            D[kind], LOG[kind] = [(0, tb[1], 0)], 0
        else:
            D[kind], LOG[kind] = dtable(O, tb[0], tb[1]), tb[1]
    br = BackBits(stream)
    st = {}
    for kind in ("LL", "OF", "ML"):
        st[kind] = br.read(LOG[kind])
    codes = {"LL": [], "OF": [], "ML": []}
    for i in range(nb_seq):
        ll, of, ml = D["LL"][st["LL"]][1], D["OF"][st["OF"]][1], D["ML"][st["ML"]][1]
        codes["LL"].append(ll); codes["OF"].append(of); codes["ML"].append(ml)
        br.read(of)            # offset extra bits
        br.read(ML_BITS[ml])
        br.read(LL_BITS[ll])
        if i != nb_seq - 1:
            for kind in ("LL", "ML", "OF"):
                ns, _, nb = D[kind][st[kind]]
                st[kind] = ns + br.read(nb)
    assert br.pos == 0, f"sequence bitstream not exactly consumed (pos {br.pos})"
    stats["seq_streams"] += 1
    # ZSTD_buildCTable for the FSE-compressed tables
    for kind, (norm, tl) in parsed.items():
        cs = codes[kind]
        mx = max(cs)
        count = np.bincount(cs, minlength=mx + 1).astype(np.uint32)
        n1 = nb_seq
        assert O.optimal_table_log(FSELOG[kind], nb_seq, mx) == tl, (kind, tl)
        if count[cs[-1]] > 1:
            count[cs[-1]] -= 1
            n1 -= 1
        mine = O.normalize_count(count, tl, 1 if n1 >= 2048 else 0).tolist()
        assert mine == norm[: mx + 1] and not any(norm[mx + 1 :]), (kind, mine, norm)
        stats["seq_tables"] += 1
        stats["logs"].add(tl)
        stats["lowprob"] += int(-1 in mine)
        stats["n_ge_2048"] += int(n1 >= 2048)


# --------------------------------------------------------------------------- frame walk
def walk_frame(O, frame: bytes, stats):
    assert frame[:4] == b"\x28\xb5\x2f\xfd"
    fhd = frame[4]
    pos = 5
    single = (fhd >> 5) & 1
    if not single:
        pos += 1
    pos += [0, 1, 2, 4][fhd & 3]
    pos += [1 if single else 0, 2, 4, 8][fhd >> 6]
    prev = {}
    while True:
        bh = int.from_bytes(frame[pos : pos + 3], "little")
        pos += 3
        last, btype, bsize = bh & 1, (bh >> 1) & 3, bh >> 3
        if btype == 1:
            bsize = 1
        blk = frame[pos : pos + bsize]
        pos += bsize
        if btype == 2:
            p = 0
            lt, sf = blk[0] & 3, (blk[0] >> 2) & 3
            if lt in (0, 1):
                if sf in (0, 2):
                    size, hs = blk[0] >> 3, 1
                elif sf == 1:
                    size, hs = int.from_bytes(blk[:2], "little") >> 4, 2
                else:
                    size, hs = int.from_bytes(blk[:3], "little") >> 4, 3
                p = hs + (size if lt == 0 else 1)
            else:
                if sf in (0, 1):
                    h, hs = int.from_bytes(blk[:3], "little"), 3
                    csize = (h >> 14) & 0x3FF
                elif sf == 2:
                    h, hs = int.from_bytes(blk[:4], "little"), 4
                    csize = (h >> 18) & 0x3FFF
                else:
                    h, hs = int.from_bytes(blk[:5], "little"), 5
                    csize = (h >> 22) & 0x3FFFF
                if lt == 2:
                    hb = blk[hs]
                    if hb < 128:
                        check_huf_weights(O, blk[hs + 1 : hs + 1 + hb], stats)
                p = hs + csize
            check_sequences(O, blk, p, prev, stats)
        if last:
            break


def corpora():
    rng = np.random.default_rng(28)
    out = []
    for name in FIXTURES:
        out.append(load_fixture(name).tobytes())
    words = [bytes(rng.integers(97, 123, rng.integers(2, 9)).astype(np.uint8)) for _ in range(600)]
    out.append(b" ".join(words[i] for i in rng.zipf(1.3, 60000) % 600))
    out.append(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 200000)))
    out.append("\n".join(f"{i},{int(rng.integers(0, 1000))},{rng.random():.4f},row{i % 37}" for i in range(12000)).encode())
    base = bytearray(rng.integers(0, 256, 4000).astype(np.uint8).tobytes())
    rep = bytearray()
    for _ in range(60):  # mutated repeats: many short matches
        for j in rng.integers(0, 4000, 12):
            base[j] = int(rng.integers(0, 256))
        rep += base
    out.append(bytes(rep))
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "fq28_oracle.c"), "rb").read()
    out.append(src)
    return out


def test_fse_primitives_match_libzstd(oracle):
    z = _zstd()
    if z is None:
        pytest.skip("libzstd.so.1 not available")
    stats = {"huf": 0, "seq_tables": 0, "seq_streams": 0, "lowprob": 0, "n_ge_2048": 0, "logs": set()}
    for data in corpora():
        for level in (1, 3, 7, 12, 19):
            for cut in (len(data), len(data) // 3, 20000):
                if cut < 64:
                    continue
                walk_frame(oracle, zstd_compress(z, data[:cut], level), stats)
    print(stats)
    assert stats["huf"] >= 30, stats
    assert stats["seq_tables"] >= 300, stats
    assert stats["lowprob"] >= 50, stats        # tables with -1 (low-probability) symbols
    assert stats["n_ge_2048"] >= 20, stats      # useLowProbCount = 1, as fqcomp28 calls it
    assert max(stats["logs"]) >= 9 and min(stats["logs"]) <= 6, stats
