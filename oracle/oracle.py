"""ctypes binding of the CPU oracle (oracle/libfq28_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/fq28_oracle.h.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (fqcomp28_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfq28_oracle.so")

SEQ_MODELS, SEQ_ALPHABET = 256, 4
QUAL_MODELS, QUAL_ALPHABET = 8192, 64
FT_SEQ_BYTES = 3076
FT_QUAL_BYTES = 1081348

REC_DTYPE = np.dtype(
    [("hdr_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8"), ("hdr_len", "<u4"), ("len", "<u4")]
)


class BenchResult(C.Structure):
    _fields_ = [
        ("t_analyze_s", C.c_double),
        ("t_compress_s", C.c_double),
        ("t_decompress_s", C.c_double),
        ("fastq_bytes", C.c_uint64),
        ("seq_bytes", C.c_uint64),
        ("qual_bytes", C.c_uint64),
        ("n_records", C.c_uint64),
        ("n_chunks", C.c_uint64),
        ("checksum", C.c_uint64),
        ("roundtrip_ok", C.c_int),
        ("err", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "fq28_oracle.c")
    hdr = os.path.join(_HERE, "fq28_oracle.h")
    stale = (
        force
        or not os.path.exists(_SO)
        or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfq28_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint, C.c_int
    L.fq28o_optimal_table_log.argtypes = [u32, sz, u32]
    L.fq28o_optimal_table_log.restype = u32
    L.fq28o_normalize_count.argtypes = [vp, u32, vp, sz, u32, u32]
    L.fq28o_normalize_count.restype = sz
    L.fq28o_spread.argtypes = [vp, vp, u32, u32]
    L.fq28o_spread.restype = None
    L.fq28o_build_ctable.argtypes = [vp, vp, vp, vp, u32, u32]
    L.fq28o_build_ctable.restype = None
    L.fq28o_build_dtable.argtypes = [vp, vp, u32, u32]
    L.fq28o_build_dtable.restype = None
    L.fq28o_parse_records.argtypes = [vp, sz, vp, sz, C.POINTER(sz), C.POINTER(i32)]
    L.fq28o_parse_records.restype = sz
    L.fq28o_split_chunks.argtypes = [vp, sz, sz, vp, sz]
    L.fq28o_split_chunks.restype = C.c_long
    L.fq28o_hist_seq.argtypes = [vp, vp, sz, vp]
    L.fq28o_hist_seq.restype = i32
    L.fq28o_hist_qual.argtypes = [vp, vp, sz, vp]
    L.fq28o_hist_qual.restype = i32
    L.fq28o_make_ft_seq.argtypes = [vp, vp]
    L.fq28o_make_ft_seq.restype = None
    L.fq28o_make_ft_qual.argtypes = [vp, vp]
    L.fq28o_make_ft_qual.restype = None
    L.fq28o_codec_seq.argtypes = [vp]
    L.fq28o_codec_seq.restype = vp
    L.fq28o_codec_qual.argtypes = [vp]
    L.fq28o_codec_qual.restype = vp
    L.fq28o_codec_free.argtypes = [vp]
    L.fq28o_codec_free.restype = None
    L.fq28o_bound_seq.argtypes = [sz]
    L.fq28o_bound_seq.restype = sz
    L.fq28o_bound_qual.argtypes = [sz]
    L.fq28o_bound_qual.restype = sz
    L.fq28o_encode_seq.argtypes = [vp, vp, vp, sz, vp, sz, vp, vp, sz, C.POINTER(sz)]
    L.fq28o_encode_seq.restype = C.c_long
    L.fq28o_encode_qual.argtypes = [vp, vp, vp, sz, vp, sz]
    L.fq28o_encode_qual.restype = C.c_long
    L.fq28o_decode_seq.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp, sz]
    L.fq28o_decode_seq.restype = i32
    L.fq28o_decode_qual.argtypes = [vp, vp, sz, vp, vp, sz]
    L.fq28o_decode_qual.restype = i32
    L.fq28o_layout_chunk.argtypes = [vp, sz, vp, vp, vp, sz, vp]
    L.fq28o_layout_chunk.restype = sz
    L.fq28o_fnv1a.argtypes = [vp, sz, C.c_uint64]
    L.fq28o_fnv1a.restype = C.c_uint64
    L.fq28o_bench.argtypes = [vp, sz, sz, sz, i32, i32, C.POINTER(BenchResult)]
    L.fq28o_bench.restype = i32
    _lib = L
    return L


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"oracle: {what} failed with code {code}")
        self.code = code


# ---------------------------------------------------------------- primitives
def optimal_table_log(max_table_log: int, src_size: int, max_sv: int) -> int:
    return lib().fq28o_optimal_table_log(max_table_log, src_size, max_sv)


def normalize_count(count, table_log: int, use_low_prob: int = 1) -> np.ndarray:
    count = np.ascontiguousarray(count, dtype=np.uint32)
    norm = np.zeros(len(count), dtype=np.int16)
    ret = lib().fq28o_normalize_count(_p(norm), table_log, _p(count), int(count.sum()), len(count) - 1, use_low_prob)
    if ret == C.c_size_t(-1).value:
        raise OracleError(-1, "normalize_count")
    return norm


def spread(norm, table_log: int) -> np.ndarray:
    norm = np.ascontiguousarray(norm, dtype=np.int16)
    cell = np.zeros(1 << table_log, dtype=np.uint8)
    lib().fq28o_spread(_p(cell), _p(norm), len(norm) - 1, table_log)
    return cell


def build_ctable(norm, table_log: int):
    norm = np.ascontiguousarray(norm, dtype=np.int16)
    st = np.zeros(1 << table_log, dtype=np.uint16)
    dfs = np.zeros(len(norm), dtype=np.int32)
    dnb = np.zeros(len(norm), dtype=np.uint32)
    lib().fq28o_build_ctable(_p(st), _p(dfs), _p(dnb), _p(norm), len(norm) - 1, table_log)
    return st, dfs, dnb


def build_dtable(norm, table_log: int) -> np.ndarray:
    norm = np.ascontiguousarray(norm, dtype=np.int16)
    cells = np.zeros(1 << table_log, dtype=np.uint32)
    lib().fq28o_build_dtable(_p(cells), _p(norm), len(norm) - 1, table_log)
    return cells


# ---------------------------------------------------------------- parsing
def parse_records(data: np.ndarray):
    """-> (records structured array, consumed_bytes).  FastqReader::parseRecords."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = C.c_size_t(0)
    err = C.c_int(0)
    L = lib()
    L.fq28o_parse_records(_p(data), data.size, None, 0, C.byref(n), C.byref(err))
    if err.value:
        raise OracleError(err.value, "parse_records")
    recs = np.zeros(n.value, dtype=REC_DTYPE)
    used = L.fq28o_parse_records(_p(data), data.size, _p(recs), n.value, C.byref(n), C.byref(err))
    if err.value:
        raise OracleError(err.value, "parse_records")
    return recs, used


def split_chunks(data: np.ndarray, reading_size: int) -> np.ndarray:
    data = np.ascontiguousarray(data, dtype=np.uint8)
    cap = data.size // max(1, reading_size) * 2 + 16
    offs = np.zeros(cap, dtype=np.uint64)
    n = lib().fq28o_split_chunks(_p(data), data.size, reading_size, _p(offs), cap)
    if n < 0:
        raise OracleError(n, "split_chunks")
    return offs[: n + 1].copy()


# ---------------------------------------------------------------- tables
def hist(data: np.ndarray, recs: np.ndarray):
    """Raw context histograms WITHOUT the +1 prior: (seq[256,4], qual[8192,64])."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    cs = np.zeros((SEQ_MODELS, SEQ_ALPHABET), dtype=np.uint32)
    cq = np.zeros((QUAL_MODELS, QUAL_ALPHABET), dtype=np.uint32)
    e = lib().fq28o_hist_seq(_p(data), _p(recs), len(recs), _p(cs))
    if e:
        raise OracleError(e, "hist_seq")
    e = lib().fq28o_hist_qual(_p(data), _p(recs), len(recs), _p(cq))
    if e:
        raise OracleError(e, "hist_qual")
    return cs, cq


def make_ft(cs: np.ndarray, cq: np.ndarray):
    """-> (ft_seq bytes[3076], ft_qual bytes[1081348]): raw FreqTable images."""
    fs = np.zeros(FT_SEQ_BYTES, dtype=np.uint8)
    fq = np.zeros(FT_QUAL_BYTES, dtype=np.uint8)
    cs = np.ascontiguousarray(cs, dtype=np.uint32)
    cq = np.ascontiguousarray(cq, dtype=np.uint32)
    lib().fq28o_make_ft_seq(_p(cs), _p(fs))
    lib().fq28o_make_ft_qual(_p(cq), _p(fq))
    return fs, fq


def ft_logs(ft: np.ndarray) -> np.ndarray:
    n = SEQ_MODELS if ft.size == FT_SEQ_BYTES else QUAL_MODELS
    a = SEQ_ALPHABET if ft.size == FT_SEQ_BYTES else QUAL_ALPHABET
    return ft[n * a * 2 : n * a * 2 + 4 * n].view("<u4")


def ft_norm(ft: np.ndarray) -> np.ndarray:
    n = SEQ_MODELS if ft.size == FT_SEQ_BYTES else QUAL_MODELS
    a = SEQ_ALPHABET if ft.size == FT_SEQ_BYTES else QUAL_ALPHABET
    return ft[: n * a * 2].view("<i2").reshape(n, a)


class Codec:
    """SequenceEncoder/Decoder + QualityEncoder/Decoder over one DatasetMeta."""

    def __init__(self, ft_seq: np.ndarray, ft_qual: np.ndarray):
        self.ft_seq = np.ascontiguousarray(ft_seq, dtype=np.uint8)
        self.ft_qual = np.ascontiguousarray(ft_qual, dtype=np.uint8)
        assert self.ft_seq.size == FT_SEQ_BYTES and self.ft_qual.size == FT_QUAL_BYTES
        self._cs = lib().fq28o_codec_seq(_p(self.ft_seq))
        self._cq = lib().fq28o_codec_qual(_p(self.ft_qual))

    def __del__(self):
        try:
            lib().fq28o_codec_free(self._cs)
            lib().fq28o_codec_free(self._cq)
        except Exception:
            pass

    def encode_chunk(self, data: np.ndarray, recs: np.ndarray):
        """CompressionWorkspace::encodeChunk minus headers / libbsc.
        -> dict(seq, qual, readlens, n_count, n_pos) raw byte arrays; `data`
        is NOT modified (a private copy takes the N->A substitution)."""
        L = lib()
        buf = np.array(data, dtype=np.uint8, copy=True)
        n = len(recs)
        tot = int(recs["len"].sum())
        seq = np.zeros(L.fq28o_bound_seq(tot), dtype=np.uint8)
        qual = np.zeros(L.fq28o_bound_qual(tot), dtype=np.uint8)
        n_count = np.zeros(n, dtype="<u2")
        n_pos = np.zeros(tot + 1, dtype="<u2")
        nn = C.c_size_t(0)
        s = L.fq28o_encode_seq(self._cs, _p(buf), _p(recs), n, _p(seq), seq.size, _p(n_count), _p(n_pos), n_pos.size, C.byref(nn))
        if s <= 0:
            raise OracleError(s, "encode_seq")
        q = L.fq28o_encode_qual(self._cq, _p(buf), _p(recs), n, _p(qual), qual.size)
        if q <= 0:
            raise OracleError(q, "encode_qual")
        return {
            "seq": seq[:s].copy(),
            "qual": qual[:q].copy(),
            "readlens": recs["len"].astype("<u2"),
            "n_count": n_count,
            "n_pos": n_pos[: nn.value].copy(),
        }

    def decode_chunk(self, enc: dict, headers: np.ndarray, hdr_lens: np.ndarray, total: int) -> np.ndarray:
        """DecompressionWorkspace::decodeChunk with the header bytes given."""
        L = lib()
        n = len(enc["readlens"])
        out = np.zeros(total, dtype=np.uint8)
        recs = np.zeros(n, dtype=REC_DTYPE)
        headers = np.ascontiguousarray(headers, dtype=np.uint8)
        hdr_lens = np.ascontiguousarray(hdr_lens, dtype=np.uint32)
        readlens = np.ascontiguousarray(enc["readlens"], dtype="<u2")
        wrote = L.fq28o_layout_chunk(_p(out), total, _p(headers), _p(hdr_lens), _p(readlens), n, _p(recs))
        if wrote == 0 and n:
            raise OracleError(-6, "layout_chunk")
        seq = np.ascontiguousarray(enc["seq"], dtype=np.uint8)
        qual = np.ascontiguousarray(enc["qual"], dtype=np.uint8)
        nc = np.ascontiguousarray(enc["n_count"], dtype="<u2")
        npos = np.ascontiguousarray(enc["n_pos"], dtype="<u2")
        e = L.fq28o_decode_seq(self._cs, _p(seq), seq.size, _p(out), _p(recs), n, _p(nc), _p(npos), npos.size)
        if e:
            raise OracleError(e, "decode_seq")
        e = L.fq28o_decode_qual(self._cq, _p(qual), qual.size, _p(out), _p(recs), n)
        if e:
            raise OracleError(e, "decode_qual")
        return out


def gather_headers(data: np.ndarray, recs: np.ndarray):
    """Concatenated header lines (with '@', no newline) + their u32 lengths."""
    lens = recs["hdr_len"].astype(np.uint32)
    parts = [data[int(o) : int(o) + int(l)] for o, l in zip(recs["hdr_off"], lens)]
    hdr = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    return hdr, lens


def fnv1a(data: np.ndarray, h: int = 1469598103934665603) -> int:
    data = np.ascontiguousarray(data, dtype=np.uint8)
    return int(lib().fq28o_fnv1a(_p(data), data.size, h))


def bench(data: np.ndarray, sample_bytes: int, reading_size: int, threads: int, do_decompress: bool = True) -> BenchResult:
    data = np.ascontiguousarray(data, dtype=np.uint8)
    res = BenchResult()
    lib().fq28o_bench(_p(data), data.size, sample_bytes, reading_size, threads, int(do_decompress), C.byref(res))
    return res
