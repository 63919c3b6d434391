"""Independent Python model of the fqcomp28 codec path -- TEST INFRASTRUCTURE.

Written from the specification in SURVEY.md (Appendix A: zstd's FSE / bitstream
semantics; section 8(a) rows A2-A8: what the reference's classes compute), NOT
from oracle/fq28_oracle.c.  It exists so that the C oracle is pinned by a second
implementation that shares no code with it: tests/golden/make_golden.py runs this
model over the fixtures and writes tests/golden/golden.json; tests compare the
oracle (and through it the CUDA path) with those digests.

Reference behaviour restated (file:line in /root/reference):
  FSE_Sequence::calculateFreqTable  src/fse_sequence.cpp:145-169
  FSE_Quality::calculateFreqTable   src/fse_quality.cpp:69-97, calcContext src/fse_quality.h:40-44
  makeNormalizedFreqTable           src/fse_common.hpp:179-200 (FSE_optimalTableLog, FSE_normalizeCount)
  FSE_Encoder start/endChunk        src/fse_common.hpp:77-90
  SequenceEncoder::encodeRecord     src/fse_sequence.cpp:35-112
  QualityEncoder::encodeRecord      src/fse_quality.cpp:5-53
  FSE_Decoder / decodeRecord        src/fse_common.hpp:130-141, src/fse_sequence.cpp:114-143, src/fse_quality.cpp:55-67
Pure Python: slow (a few microseconds per symbol), meant for fixtures of a few MB.
"""
from __future__ import annotations

import struct

SEQ_N, SEQ_A, QUAL_N, QUAL_A = 256, 4, 8192, 64
SEQ_INIT = 0xD7  # bases T, C, C, T before the read, the closest one in the top two bits
BASE = {65: 0, 67: 1, 71: 2, 84: 3}  # A C G T


def hb(x: int) -> int:
    return x.bit_length() - 1


# ---------------------------------------------------------------- Appendix A.1 / A.2
def optimal_table_log(max_table_log: int, src_size: int, max_sv: int) -> int:
    max_bits_src = (hb((src_size - 1) & 0xFFFFFFFF) - 2) & 0xFFFFFFFF  # U32 arithmetic: srcSize = 4 wraps to 0xFFFFFFFF
    min_bits = min(hb(src_size & 0xFFFFFFFF) + 1, hb(max_sv) + 2)
    t = max_table_log if max_table_log else 11
    if max_bits_src < t:
        t = max_bits_src
    if min_bits > t:
        t = min_bits
    return max(5, min(12, t))


RTB = (0, 473195, 504333, 520860, 550000, 700000, 750000, 830000)


def normalize_m2(norm, t, count, total, max_sv, low):
    NYA = -2
    distributed = 0
    low_threshold = (total >> t) & 0xFFFFFFFF
    low_one = ((total * 3) >> (t + 1)) & 0xFFFFFFFF
    for s in range(max_sv + 1):
        if count[s] == 0:
            norm[s] = 0
        elif count[s] <= low_threshold:
            norm[s] = low
            distributed += 1
            total -= count[s]
        elif count[s] <= low_one:
            norm[s] = 1
            distributed += 1
            total -= count[s]
        else:
            norm[s] = NYA
    to_distribute = (1 << t) - distributed
    if to_distribute == 0:
        return
    if total // to_distribute > low_one:
        low_one = ((total * 3) // (to_distribute * 2)) & 0xFFFFFFFF
        for s in range(max_sv + 1):
            if norm[s] == NYA and count[s] <= low_one:
                norm[s] = 1
                distributed += 1
                total -= count[s]
        to_distribute = (1 << t) - distributed
    if distributed == max_sv + 1:
        best, best_c = 0, 0
        for s in range(max_sv + 1):
            if count[s] > best_c:
                best, best_c = s, count[s]
        norm[best] += to_distribute
        return
    if total == 0:
        s = 0
        while to_distribute > 0:
            if norm[s] > 0:
                to_distribute -= 1
                norm[s] += 1
            s = (s + 1) % (max_sv + 1)
        return
    v_step_log = 62 - t
    mid = (1 << (v_step_log - 1)) - 1
    r_step = (((1 << v_step_log) * to_distribute) + mid) // (total & 0xFFFFFFFF)
    tmp = mid
    for s in range(max_sv + 1):
        if norm[s] == NYA:
            end = tmp + count[s] * r_step
            w = ((end >> v_step_log) & 0xFFFFFFFF) - ((tmp >> v_step_log) & 0xFFFFFFFF)
            assert w >= 1
            norm[s] = w
            tmp = end


def normalize_count(t: int, count, total: int, max_sv: int, use_low_prob: int = 1):
    low = -1 if use_low_prob else 1
    scale = 62 - t
    step = (1 << 62) // (total & 0xFFFFFFFF)
    v_step = 1 << (scale - 20)
    still = 1 << t
    largest, largest_p = 0, 0
    low_threshold = (total >> t) & 0xFFFFFFFF
    norm = [0] * (max_sv + 1)
    for s in range(max_sv + 1):
        c = count[s]
        assert c != total
        if c == 0:
            norm[s] = 0
        elif c <= low_threshold:
            norm[s] = low
            still -= 1
        else:
            p = (c * step) >> scale
            if p < 8 and (c * step - (p << scale)) > v_step * RTB[p]:
                p += 1
            if p > largest_p:
                largest_p, largest = p, s
            norm[s] = p
            still -= p
    if -still >= (norm[largest] >> 1):
        normalize_m2(norm, t, count, total, max_sv, low)
    else:
        norm[largest] += still
    return norm


# ---------------------------------------------------------------- Appendix A.3 / A.4 / A.6
def spread(norm, t: int):
    T = 1 << t
    step = (T >> 1) + (T >> 3) + 3
    cell = [0] * T
    high = T - 1
    for s, n in enumerate(norm):
        if n == -1:
            cell[high] = s
            high -= 1
    pos = 0
    for s, n in enumerate(norm):
        for _ in range(max(n, 0)):
            cell[pos] = s
            pos = (pos + step) & (T - 1)
            while pos > high:
                pos = (pos + step) & (T - 1)
    assert pos == 0
    return cell


class CTable:
    def __init__(self, norm, t: int):
        T = 1 << t
        self.t = t
        cell = spread(norm, t)
        cumul = [0]
        for n in norm:
            cumul.append(cumul[-1] + (1 if n == -1 else n))
        nxt = cumul[:-1].copy()
        self.state_table = [0] * T
        for u in range(T):
            s = cell[u]
            self.state_table[nxt[s]] = T + u
            nxt[s] += 1
        self.delta_nb, self.delta_find = [], []
        total = 0
        for n in norm:
            if n == 0:
                self.delta_nb.append(((t + 1) << 16) - T)
                self.delta_find.append(0)
            elif n in (-1, 1):
                self.delta_nb.append((t << 16) - T)
                self.delta_find.append(total - 1)
                total += 1
            else:
                mbo = t - hb(n - 1)
                self.delta_nb.append((mbo << 16) - (n << mbo))
                self.delta_find.append(total - n)
                total += n


class DTable:
    def __init__(self, norm, t: int):
        T = 1 << t
        cell = spread(norm, t)
        nxt = [1 if n == -1 else n for n in norm]
        self.sym, self.nb, self.new = [0] * T, [0] * T, [0] * T
        for u in range(T):
            s = cell[u]
            x = nxt[s]
            nxt[s] += 1
            nb = t - hb(x)
            self.sym[u], self.nb[u], self.new[u] = s, nb, (x << nb) - T


# ---------------------------------------------------------------- bit streams (A.5 / A.6)
class BitWriter:
    def __init__(self):
        self.words, self.acc, self.n = [], 0, 0

    def add(self, value: int, nb: int):
        self.acc |= (value & ((1 << nb) - 1)) << self.n
        self.n += nb
        while self.n >= 64:
            self.words.append(self.acc & 0xFFFFFFFFFFFFFFFF)
            self.acc >>= 64
            self.n -= 64

    def close(self) -> bytes:
        self.add(1, 1)  # BIT_closeCStream: the end mark
        total_bits = 64 * len(self.words) + self.n
        raw = b"".join(struct.pack("<Q", w) for w in self.words) + self.acc.to_bytes(8, "little")
        return raw[: (total_bits + 7) // 8]


class BitReader:
    def __init__(self, data: bytes):
        self.v = int.from_bytes(data, "little")
        self.p = 8 * (len(data) - 1) + hb(data[-1])  # the end mark itself is not data

    def read(self, nb: int) -> int:
        self.p -= nb
        assert self.p >= 0
        return (self.v >> self.p) & ((1 << nb) - 1)


# ---------------------------------------------------------------- FASTQ
def parse(fastq: bytes):
    """-> list of (header, seq, qual) bytes; 4-line records (src/fastq_io.cpp:67-125)"""
    lines = fastq.split(b"\n")
    recs = []
    for i in range(0, len(lines) - 3, 4):
        recs.append((lines[i], lines[i + 1], lines[i + 3]))
    return recs


def qual_ctx(q: int, q1: int, q2: int) -> int:
    return ((((q1 if q1 > q2 else q2) << 6) + q) & 0xFFF) + ((1 << 12) if q1 == q2 else 0)


# ---------------------------------------------------------------- frequency tables
def freq_tables(recs):
    cs = [[1] * SEQ_A for _ in range(SEQ_N)]
    cq = [[1] * QUAL_A for _ in range(QUAL_N)]
    for _, seq, qual in recs:
        ctx = SEQ_INIT
        for c in seq:
            if c == 78:  # 'N': skipped, context unchanged
                continue
            s = BASE[c]
            cs[ctx][s] += 1
            ctx = (ctx >> 2) + (s << 6)
        ctx, q1, q2 = qual_ctx(0, 0, 0), 0, 0
        for c in qual:
            q = c - 33
            cq[ctx][q] += 1
            ctx = qual_ctx(q, q1, q2)
            q2, q1 = q1, q
    return make_ft(cs, SEQ_A), make_ft(cq, QUAL_A)


def make_ft(counts, alphabet):
    norms, logs = [], []
    for c in counts:
        total = sum(c)
        t = optimal_table_log(0, total, alphabet - 1)
        norms.append(normalize_count(t, c, total, alphabet - 1, 1))
        logs.append(t)
    return norms, logs


def ft_image(ft) -> bytes:
    """raw FreqTable<N, A> as dumped by src/prepare.cpp:18-20: short norm[N][A]; unsigned logs[N]; unsigned max_log"""
    norms, logs = ft
    out = bytearray()
    for n in norms:
        out += struct.pack("<%dh" % len(n), *n)
    out += struct.pack("<%dI" % len(logs), *logs)
    out += struct.pack("<I", max(logs))
    return bytes(out)


# ---------------------------------------------------------------- encoders
class Encoder:
    def __init__(self, ft):
        self.norms, self.logs = ft
        self.tabs = {}
        self.state = [1 << t for t in self.logs]  # FSE_initCState: value = 1 << log
        self.bw = BitWriter()

    def put(self, ctx: int, sym: int):
        tab = self.tabs.get(ctx)
        if tab is None:
            tab = self.tabs[ctx] = CTable(self.norms[ctx], self.logs[ctx])
        v = self.state[ctx]
        nb = (v + tab.delta_nb[sym]) >> 16
        self.bw.add(v, nb)
        self.state[ctx] = tab.state_table[(v >> nb) + tab.delta_find[sym]]

    def end(self) -> bytes:
        for ctx, t in enumerate(self.logs):  # FSE_flushCState, contexts ascending
            self.bw.add(self.state[ctx], t)
        return self.bw.close()


def encode_chunk(recs, ft_seq, ft_qual):
    """-> dict(seq, qual, readlens, n_count, n_pos) of bytes (src/workspace.cpp:14-45 minus headers / libbsc)"""
    es, eq = Encoder(ft_seq), Encoder(ft_qual)
    readlens, n_count, n_pos = bytearray(), bytearray(), bytearray()
    for _, seq, qual in recs:
        L = len(seq)
        readlens += struct.pack("<H", L)
        # replaceAndEncodeNs
        syms, cnt, prev = [], 0, 0
        for i, c in enumerate(seq):
            if c == 78:
                n_pos += struct.pack("<H", i - prev)
                prev = i
                cnt += 1
                syms.append(0)
            else:
                syms.append(BASE[c])
        n_count += struct.pack("<H", cnt)
        ctxs, ctx = [], SEQ_INIT
        for s in syms:
            ctxs.append(ctx)
            ctx = (ctx >> 2) + (s << 6)
        for i in range(L - 1, -1, -1):
            es.put(ctxs[i], syms[i])
        q = [c - 33 for c in qual]
        for i in range(L - 1, -1, -1):
            a = q[i - 1] if i >= 1 else 0
            b = q[i - 2] if i >= 2 else 0
            c = q[i - 3] if i >= 3 else 0
            eq.put(qual_ctx(a, b, c), q[i])
    return {"seq": es.end(), "qual": eq.end(), "readlens": bytes(readlens), "n_count": bytes(n_count), "n_pos": bytes(n_pos)}


# ---------------------------------------------------------------- decoders (round-trip check of the model itself)
def decode_chunk(enc, ft_seq, ft_qual, n_records):
    rl = struct.unpack("<%dH" % n_records, enc["readlens"])
    nc = struct.unpack("<%dH" % n_records, enc["n_count"])
    npos = struct.unpack("<%dH" % (len(enc["n_pos"]) // 2), enc["n_pos"])
    out_seq, out_qual = [None] * n_records, [None] * n_records

    def start(data, ft):
        br = BitReader(data)
        st = [0] * len(ft[1])
        for ctx in range(len(ft[1]) - 1, -1, -1):
            st[ctx] = br.read(ft[1][ctx])
        return br, st, {}

    bs, ss, ts = start(enc["seq"], ft_seq)
    bq, sq, tq = start(enc["qual"], ft_qual)

    def get(br, st, tabs, ft, ctx):
        tab = tabs.get(ctx)
        if tab is None:
            tab = tabs[ctx] = DTable(ft[0][ctx], ft[1][ctx])
        u = st[ctx]
        st[ctx] = tab.new[u] + br.read(tab.nb[u])
        return tab.sym[u]

    np_end = len(npos)
    for r in range(n_records - 1, -1, -1):
        L = rl[r]
        ctx, seq = SEQ_INIT, bytearray()
        for _ in range(L):
            s = get(bs, ss, ts, ft_seq, ctx)
            seq.append(b"ACGT"[s])
            ctx = (ctx >> 2) + (s << 6)
        mine = npos[np_end - nc[r] : np_end]
        np_end -= nc[r]
        p = 0
        for d in mine:
            p += d
            seq[p] = 78
        out_seq[r] = bytes(seq)
        ctx, q1, q2, qual = qual_ctx(0, 0, 0), 0, 0, bytearray()
        for _ in range(L):
            q = get(bq, sq, tq, ft_qual, ctx)
            qual.append(q + 33)
            ctx = qual_ctx(q, q1, q2)
            q2, q1 = q1, q
        out_qual[r] = bytes(qual)
    assert bs.p == 0 and bq.p == 0, "stream not consumed exactly"
    return out_seq, out_qual
