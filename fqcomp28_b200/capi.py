"""ctypes binding of libfq28.so (the C ABI in include/fq28.h).

This is plumbing for bench.py and the parity tests; the product is the shared
library.  There is no CPU fallback: if the library is missing, or no CUDA
device is usable, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfq28.so")

SEQ_MODELS, SEQ_ALPHABET = 256, 4
QUAL_MODELS, QUAL_ALPHABET = 8192, 64
FT_SEQ_BYTES = 3076
FT_QUAL_BYTES = 1081348

ERR_NAMES = {
    -1: "FQ28_ERR_CUDA",
    -2: "FQ28_ERR_FORMAT",
    -3: "FQ28_ERR_ALPHABET",
    -4: "FQ28_ERR_SHORT",
    -5: "FQ28_ERR_LONG",
    -6: "FQ28_ERR_CAP",
    -7: "FQ28_ERR_ARG",
    -8: "FQ28_ERR_STREAM",
}


class Fq28Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


HDR_MAX_FIELDS = 32


class HdrFormat(C.Structure):  # fq28_hdr_format
    _fields_ = [
        ("n_fields", C.c_uint32),
        ("is_string", C.c_uint8 * HDR_MAX_FIELDS),
        ("separators", C.c_uint8 * HDR_MAX_FIELDS),
        ("first_numeric", C.c_int32 * HDR_MAX_FIELDS),
        ("first_str_off", C.c_uint32 * (HDR_MAX_FIELDS + 1)),
        ("first_strings", C.c_char_p),
    ]


class HdrFieldInfo(C.Structure):  # fq28_hdr_field_info
    _fields_ = [(n, C.c_uint64) for n in ("flag_off", "flag_len", "content_off", "content_len", "clen_off", "clen_len")]


class ChunkInfo(C.Structure):
    _fields_ = [
        ("fastq_off", C.c_uint64),
        ("total", C.c_uint32),
        ("n_records", C.c_uint32),
        ("rec_off", C.c_uint64),
        ("seq_off", C.c_uint64),
        ("qual_off", C.c_uint64),
        ("seq_len", C.c_uint32),
        ("qual_len", C.c_uint32),
        ("n_pos_off", C.c_uint64),
        ("n_pos_len", C.c_uint32),
        ("hdr_bytes", C.c_uint32),
        ("hdr_off", C.c_uint64),
    ]


class EncArenas(C.Structure):
    _fields_ = [
        ("seq", C.c_void_p), ("seq_cap", C.c_size_t),
        ("qual", C.c_void_p), ("qual_cap", C.c_size_t),
        ("readlens", C.c_void_p), ("readlens_cap", C.c_size_t),
        ("n_count", C.c_void_p), ("n_count_cap", C.c_size_t),
        ("n_pos", C.c_void_p), ("n_pos_cap", C.c_size_t),
        ("hdr_lens", C.c_void_p), ("hdr_lens_cap", C.c_size_t),
        ("headers", C.c_void_p), ("headers_cap", C.c_size_t),
    ]


class EncSummary(C.Structure):
    _fields_ = [
        ("n_chunks", C.c_uint64), ("n_records", C.c_uint64), ("n_symbols", C.c_uint64),
        ("seq_bytes", C.c_uint64), ("qual_bytes", C.c_uint64), ("n_pos_entries", C.c_uint64),
        ("consumed", C.c_uint64), ("hdr_bytes", C.c_uint64),
    ]


class DecArenas(C.Structure):
    _fields_ = [
        ("seq", C.c_void_p), ("seq_bytes", C.c_size_t),
        ("qual", C.c_void_p), ("qual_bytes", C.c_size_t),
        ("readlens", C.c_void_p),
        ("n_count", C.c_void_p),
        ("n_pos", C.c_void_p), ("n_pos_entries", C.c_size_t),
        ("hdr_lens", C.c_void_p),
        ("headers", C.c_void_p), ("headers_bytes", C.c_size_t),
        ("n_records", C.c_size_t),
    ]


# every symbol include/fq28.h declares (tests check the library exports all)
SYMBOLS = [
    "fq28_create", "fq28_destroy", "fq28_last_error", "fq28_set_stream", "fq28_launch_count",
    "fq28_parse", "fq28_split", "fq28_hist", "fq28_hist_dev", "fq28_build_tables",
    "fq28_build_tables_dev", "fq28_load_tables", "fq28_compress", "fq28_compress_dev",
    "fq28_compress_fetch", "fq28_bound_seq", "fq28_bound_qual", "fq28_decompress",
    "fq28_decompress_dev", "fq28_get_ctable", "fq28_get_dtable", "fq28_compress_dev_arenas",
    "fq28_last_timings", "fq28_stage_name", "fq28_tokenize_headers", "fq28_detokenize_headers",
    "fq28_device_count", "fq28_stage", "fq28_plan", "fq28_plan_dev", "fq28_preparse_dev", "fq28_plan_cut_dev",
    "fq28_preparse", "fq28_plan_cut",
]

_lib = None


def load() -> C.CDLL:
    """dlopen libfq28.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C fqcomp28_b200/csrc).  fqcomp28_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    psz = C.POINTER(C.c_size_t)
    L.fq28_create.argtypes = [i32, C.POINTER(vp)]
    L.fq28_destroy.argtypes = [vp]
    L.fq28_destroy.restype = None
    L.fq28_last_error.argtypes = [vp]
    L.fq28_last_error.restype = C.c_char_p
    L.fq28_set_stream.argtypes = [vp, vp]
    L.fq28_launch_count.argtypes = [vp]
    L.fq28_launch_count.restype = C.c_uint64
    L.fq28_parse.argtypes = [vp, vp, sz, vp, vp, vp, vp, vp, sz, psz, psz]
    L.fq28_split.argtypes = [vp, vp, sz, sz, i32, vp, sz, psz]
    L.fq28_hist.argtypes = [vp, vp, sz, vp, vp]
    L.fq28_hist_dev.argtypes = [vp, vp, sz, vp, vp]
    L.fq28_build_tables.argtypes = [vp, vp, vp, vp, vp]
    L.fq28_build_tables_dev.argtypes = [vp, vp, vp, vp, vp]
    L.fq28_load_tables.argtypes = [vp, vp, vp]
    L.fq28_compress.argtypes = [vp, vp, sz, sz, sz, i32, vp, vp, C.POINTER(EncArenas), C.POINTER(ChunkInfo), sz, C.POINTER(EncSummary)]
    L.fq28_compress_dev.argtypes = [vp, vp, sz, sz, sz, i32, vp, vp, C.POINTER(ChunkInfo), sz, C.POINTER(EncSummary)]
    L.fq28_compress_fetch.argtypes = [vp, C.POINTER(EncArenas)]
    L.fq28_device_count.argtypes = []
    L.fq28_stage.argtypes = [vp, vp, sz]
    L.fq28_plan.argtypes = [vp, vp, sz, sz, i32, C.POINTER(C.c_uint64), psz]
    L.fq28_plan_dev.argtypes = [vp, vp, sz, sz, i32, C.POINTER(C.c_uint64), psz]
    L.fq28_preparse_dev.argtypes = [vp, vp, sz]
    L.fq28_preparse.argtypes = [vp, vp, sz]
    L.fq28_plan_cut.argtypes = [vp, vp, sz, sz, i32, C.c_uint64, C.POINTER(C.c_uint64), psz]
    L.fq28_plan_cut_dev.argtypes = [vp, vp, sz, sz, i32, C.c_uint64, C.POINTER(C.c_uint64), psz]
    L.fq28_bound_seq.argtypes = [sz]
    L.fq28_bound_seq.restype = sz
    L.fq28_bound_qual.argtypes = [sz]
    L.fq28_bound_qual.restype = sz
    L.fq28_decompress.argtypes = [vp, C.POINTER(DecArenas), C.POINTER(ChunkInfo), sz, vp, sz, psz]
    L.fq28_decompress_dev.argtypes = [vp, C.POINTER(DecArenas), C.POINTER(ChunkInfo), sz, vp, sz, psz]
    L.fq28_get_ctable.argtypes = [vp, i32, C.c_uint, vp, vp, vp, C.POINTER(C.c_uint)]
    L.fq28_get_dtable.argtypes = [vp, i32, C.c_uint, vp, C.POINTER(C.c_uint)]
    L.fq28_compress_dev_arenas.argtypes = [vp, C.POINTER(DecArenas)]
    L.fq28_last_timings.argtypes = [vp, vp, sz, psz]
    L.fq28_stage_name.argtypes = [sz]
    L.fq28_stage_name.restype = C.c_char_p
    L.fq28_tokenize_headers.argtypes = [vp, vp, sz, vp, sz, vp, sz, C.POINTER(HdrFormat), vp, sz, C.POINTER(HdrFieldInfo), psz]
    L.fq28_detokenize_headers.argtypes = [vp, vp, sz, C.POINTER(HdrFieldInfo), vp, sz, C.POINTER(HdrFormat), vp, sz, vp, psz]
    _lib = L
    return L


def _ptr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    return a.ctypes.data


class Handle:
    """One fq28 handle (= one GPU worker).  Mirrors the role of
    CompressionWorkspace / DecompressionWorkspace (src/workspace.h)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.fq28_create(device, C.byref(h))
        if rc != 0:
            raise Fq28Error(rc, f"fq28_create(device={device}) failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        if stream is not None:
            self._ck(self.L.fq28_set_stream(self.h, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "h", None):
            self.L.fq28_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise Fq28Error(rc, self.L.fq28_last_error(self.h).decode())

    @property
    def launches(self) -> int:
        return int(self.L.fq28_launch_count(self.h))

    def tokenize_headers(self, headers: np.ndarray, hdr_lens: np.ndarray, chunk_rec, first_header: bytes):
        """fq28_tokenize_headers: header lines back to back + lengths + first record of every chunk
        (n_chunks + 1 entries) -> (format dict, [[{flag, content, clen} per field] per chunk]).
        The format is derived from `first_header` like HeaderFormatSpeciciation::fromHeader."""
        headers = np.ascontiguousarray(headers, dtype=np.uint8)
        hdr_lens = np.ascontiguousarray(hdr_lens, dtype=np.uint16)
        cr = np.ascontiguousarray(chunk_rec, dtype=np.uint64)
        fmt, types, seps, first = HdrFormat(), [], [], []
        body, pos = first_header[1:], 0
        while True:  # src/headers.cpp:43-73
            end = pos
            while end < len(body) and chr(body[end]).isalnum() and body[end] < 128:
                end += 1
            fld = body[pos:end]
            types.append(0 if fld.isdigit() or len(fld) == 0 else 1)
            first.append(fld)
            if end == len(body):
                break
            seps.append(body[end])
            pos = end + 1
        F = len(types)
        fmt.n_fields = F
        strings, off = b"", [0]
        for i in range(F):
            fmt.is_string[i] = types[i]
            if i < len(seps):
                fmt.separators[i] = seps[i]
            if types[i]:
                strings += first[i]
            else:
                fmt.first_numeric[i] = int(first[i]) if first[i] else 0
            off.append(len(strings))
        for i, o in enumerate(off):
            fmt.first_str_off[i] = o
        fmt.first_strings = strings
        n_chunks = len(cr) - 1
        arena = np.zeros(headers.size + 6 * hdr_lens.size * F + 64, np.uint8)
        infos = (HdrFieldInfo * (n_chunks * F))()
        used = C.c_size_t(0)
        self._ck(self.L.fq28_tokenize_headers(self.h, _ptr(headers), headers.size, _ptr(hdr_lens), hdr_lens.size, _ptr(cr),
                                              n_chunks, C.byref(fmt), _ptr(arena), arena.size, infos, C.byref(used)))
        out = []
        for k in range(n_chunks):
            row = []
            for i in range(F):
                fi = infos[k * F + i]
                row.append({
                    "flag": arena[fi.flag_off : fi.flag_off + fi.flag_len].tobytes(),
                    "content": arena[fi.content_off : fi.content_off + fi.content_len].tobytes(),
                    "clen": arena[fi.clen_off : fi.clen_off + fi.clen_len].tobytes(),
                })
            out.append(row)
        self._hdr_raw = (fmt, strings, arena[: used.value].copy(), infos, cr)  # for detokenize_headers
        return {"types": types, "separators": bytes(seps), "first": first}, out

    def detokenize_headers(self, raw=None):
        """fq28_detokenize_headers on the streams of the last tokenize_headers call (or `raw` =
        (fmt, strings, arena, infos, chunk_rec)) -> (header bytes, hdr_lens)."""
        fmt, strings, arena, infos, cr = raw if raw is not None else self._hdr_raw
        n_rec = int(cr[-1])
        lens = np.zeros(n_rec, np.uint16)
        used = C.c_size_t(0)
        out = np.zeros(64 * n_rec + 4096, np.uint8)
        rc = self.L.fq28_detokenize_headers(self.h, _ptr(arena), arena.size, infos, _ptr(cr), len(cr) - 1, C.byref(fmt),
                                            _ptr(out), out.size, _ptr(lens), C.byref(used))
        if rc == -6 and used.value > out.size:  # FQ28_ERR_CAP: the call reports the size it needs
            out = np.zeros(used.value, np.uint8)
            rc = self.L.fq28_detokenize_headers(self.h, _ptr(arena), arena.size, infos, _ptr(cr), len(cr) - 1, C.byref(fmt),
                                                _ptr(out), out.size, _ptr(lens), C.byref(used))
        self._ck(rc)
        return out[: used.value].copy(), lens

    def timings(self) -> dict:
        ms = (C.c_float * 16)()
        n = C.c_size_t(0)
        self._ck(self.L.fq28_last_timings(self.h, ms, 16, C.byref(n)))
        return {self.L.fq28_stage_name(i).decode(): float(ms[i]) for i in range(n.value)}

    # ------------------------------------------------------------ parse/split
    def parse(self, data: np.ndarray):
        """FastqReader::parseRecords -> (dict of arrays, consumed)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n = C.c_size_t(0)
        cons = C.c_size_t(0)
        self._ck(self.L.fq28_parse(self.h, _ptr(data), data.size, 0, 0, 0, 0, 0, 0, C.byref(n), C.byref(cons)))
        k = n.value
        out = {
            "hdr_off": np.zeros(k, np.uint32), "seq_off": np.zeros(k, np.uint32), "qual_off": np.zeros(k, np.uint32),
            "hdr_len": np.zeros(k, np.uint16), "len": np.zeros(k, np.uint16),
        }
        self._ck(self.L.fq28_parse(self.h, _ptr(data), data.size, _ptr(out["hdr_off"]), _ptr(out["seq_off"]),
                                   _ptr(out["qual_off"]), _ptr(out["hdr_len"]), _ptr(out["len"]), k,
                                   C.byref(n), C.byref(cons)))
        return out, cons.value

    def split(self, data: np.ndarray, reading_size: int, eof: bool = True) -> np.ndarray:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cap = 2 * (data.size // max(1, reading_size)) + 8
        offs = np.zeros(cap, dtype=np.uint64)
        n = C.c_size_t(0)
        self._ck(self.L.fq28_split(self.h, _ptr(data), data.size, reading_size, int(eof), _ptr(offs), cap, C.byref(n)))
        return offs[: n.value + 1].copy()

    def stage(self, data: np.ndarray | None) -> None:
        """fq28_stage: start the H2D copy of a host range early (None forgets it)."""
        if data is None:
            self._ck(self.L.fq28_stage(self.h, None, 0))
        else:
            self._ck(self.L.fq28_stage(self.h, _ptr(data), data.size))

    def plan(self, data: np.ndarray, reading_size: int, eof: bool = True):
        """fq28_plan: parseRecords + chunk boundary walk only -> (consumed, n_chunks)."""
        cons, n = C.c_uint64(0), C.c_size_t(0)
        self._ck(self.L.fq28_plan(self.h, _ptr(data), data.size, reading_size, int(eof), C.byref(cons), C.byref(n)))
        return cons.value, n.value

    def plan_dev(self, d_fastq: int, n_bytes: int, reading_size: int, eof: bool = True):
        cons, n = C.c_uint64(0), C.c_size_t(0)
        self._ck(self.L.fq28_plan_dev(self.h, d_fastq, n_bytes, reading_size, int(eof), C.byref(cons), C.byref(n)))
        return cons.value, n.value

    def preparse(self, data: np.ndarray) -> None:
        self._ck(self.L.fq28_preparse(self.h, _ptr(data), data.size))

    def plan_cut(self, data: np.ndarray, reading_size: int, eof: bool, first_cut: int):
        cons, n = C.c_uint64(0), C.c_size_t(0)
        self._ck(self.L.fq28_plan_cut(self.h, _ptr(data), data.size, reading_size, int(eof), first_cut, C.byref(cons), C.byref(n)))
        return cons.value, n.value

    def preparse_dev(self, d_fastq: int, n_bytes: int) -> None:
        self._ck(self.L.fq28_preparse_dev(self.h, d_fastq, n_bytes))

    def plan_cut_dev(self, d_fastq: int, n_bytes: int, reading_size: int, eof: bool, first_cut: int):
        """boundary walk of a (pre)parsed slab from `first_cut` -> (consumed, n_chunks); chunk 0 is
        the dropped head when first_cut > 0"""
        cons, n = C.c_uint64(0), C.c_size_t(0)
        self._ck(self.L.fq28_plan_cut_dev(self.h, d_fastq, n_bytes, reading_size, int(eof), first_cut, C.byref(cons), C.byref(n)))
        return cons.value, n.value

    # ------------------------------------------------------------ tables
    def hist(self, data: np.ndarray, cs: np.ndarray | None = None, cq: np.ndarray | None = None):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cs = np.zeros((SEQ_MODELS, SEQ_ALPHABET), np.uint32) if cs is None else cs
        cq = np.zeros((QUAL_MODELS, QUAL_ALPHABET), np.uint32) if cq is None else cq
        self._ck(self.L.fq28_hist(self.h, _ptr(data), data.size, _ptr(cs), _ptr(cq)))
        return cs, cq

    def hist_dev(self, d_fastq: int, n_bytes: int, d_cs: int, d_cq: int):
        self._ck(self.L.fq28_hist_dev(self.h, d_fastq, n_bytes, d_cs, d_cq))

    def build_tables(self, cs: np.ndarray, cq: np.ndarray):
        cs = np.ascontiguousarray(cs, dtype=np.uint32)
        cq = np.ascontiguousarray(cq, dtype=np.uint32)
        fs = np.zeros(FT_SEQ_BYTES, np.uint8)
        fq = np.zeros(FT_QUAL_BYTES, np.uint8)
        self._ck(self.L.fq28_build_tables(self.h, _ptr(cs), _ptr(cq), _ptr(fs), _ptr(fq)))
        return fs, fq

    def build_tables_dev(self, d_cs: int, d_cq: int):
        fs = np.zeros(FT_SEQ_BYTES, np.uint8)
        fq = np.zeros(FT_QUAL_BYTES, np.uint8)
        self._ck(self.L.fq28_build_tables_dev(self.h, d_cs, d_cq, _ptr(fs), _ptr(fq)))
        return fs, fq

    def load_tables(self, fs: np.ndarray, fq: np.ndarray):
        fs = np.ascontiguousarray(fs, dtype=np.uint8)
        fq = np.ascontiguousarray(fq, dtype=np.uint8)
        assert fs.size == FT_SEQ_BYTES and fq.size == FT_QUAL_BYTES
        self._ck(self.L.fq28_load_tables(self.h, _ptr(fs), _ptr(fq)))

    def get_ctable(self, kind: int, ctx: int):
        a = SEQ_ALPHABET if kind == 0 else QUAL_ALPHABET
        st = np.zeros(4096, np.uint16)
        dfs = np.zeros(a, np.int32)
        dnb = np.zeros(a, np.uint32)
        lg = C.c_uint(0)
        self._ck(self.L.fq28_get_ctable(self.h, kind, ctx, _ptr(st), _ptr(dfs), _ptr(dnb), C.byref(lg)))
        return st[: 1 << lg.value].copy(), dfs, dnb, lg.value

    def get_dtable(self, kind: int, ctx: int):
        cells = np.zeros(4096, np.uint32)
        lg = C.c_uint(0)
        self._ck(self.L.fq28_get_dtable(self.h, kind, ctx, _ptr(cells), C.byref(lg)))
        return cells[: 1 << lg.value].copy(), lg.value

    # ------------------------------------------------------------ compress
    def compress(self, data: np.ndarray, reading_size: int, eof: bool = True, arenas: dict | None = None,
                 sample_bytes: int = 0, ft_out: tuple | None = None):
        """fq28_compress on a host slab -> (infos, summary, arenas dict of numpy arrays).
        sample_bytes > 0 runs analyzeDataset on the leading sample first; ft_out =
        (ft_seq, ft_qual) uint8 arrays then receive the FreqTable images."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n = data.size
        max_chunks = 2 * (n // max(1, reading_size)) + 8
        max_rec = n // 12 + 1
        if arenas is None:
            arenas = {
                "seq": np.zeros(n // 2 + 4096 * max_chunks + 4096, np.uint8),
                "qual": np.zeros(n + 16384 * max_chunks + 4096, np.uint8),
                "readlens": np.zeros(max_rec, np.uint16),
                "n_count": np.zeros(max_rec, np.uint16),
                "n_pos": np.zeros(n // 2 + 16, np.uint16),
                "hdr_lens": np.zeros(max_rec, np.uint16),
                "headers": np.zeros(n // 2 + 64, np.uint8),
            }
        ea = self._enc_arenas(arenas)
        infos = (ChunkInfo * max_chunks)()
        summ = EncSummary()
        fs, fq = ft_out if ft_out is not None else (None, None)
        self._ck(self.L.fq28_compress(self.h, _ptr(data), n, sample_bytes, reading_size, int(eof), _ptr(fs), _ptr(fq),
                                      C.byref(ea), infos, max_chunks, C.byref(summ)))
        return infos, summ, arenas

    @staticmethod
    def _enc_arenas(a: dict) -> EncArenas:
        ea = EncArenas()
        for k in ("seq", "qual", "readlens", "n_count", "n_pos", "hdr_lens", "headers"):
            arr = a.get(k)
            setattr(ea, k, _ptr(arr) if arr is not None else None)
            setattr(ea, k + "_cap", int(arr.size) if arr is not None else 0)
        return ea

    def compress_dev(self, d_fastq: int, n_bytes: int, reading_size: int, eof: bool = True, max_chunks: int | None = None,
                     sample_bytes: int = 0, ft_out: tuple | None = None, infos=None):
        if max_chunks is None:
            max_chunks = 2 * (n_bytes // max(1, reading_size)) + 8
        if infos is None:
            infos = (ChunkInfo * max_chunks)()
        summ = EncSummary()
        fs, fq = ft_out if ft_out is not None else (None, None)
        self._ck(self.L.fq28_compress_dev(self.h, d_fastq, n_bytes, sample_bytes, reading_size, int(eof), _ptr(fs), _ptr(fq),
                                          infos, max_chunks, C.byref(summ)))
        return infos, summ

    def compress_fetch(self, arenas: dict):
        ea = self._enc_arenas(arenas)
        self._ck(self.L.fq28_compress_fetch(self.h, C.byref(ea)))

    def compress_fetch_all(self, summ) -> dict:
        """arenas sized from the summary of the last fq28_compress_dev, filled by fq28_compress_fetch"""
        nr = int(summ.n_records)
        arenas = {
            "seq": np.zeros(int(summ.seq_bytes) + 64, np.uint8), "qual": np.zeros(int(summ.qual_bytes) + 64, np.uint8),
            "readlens": np.zeros(nr + 8, np.uint16), "n_count": np.zeros(nr + 8, np.uint16),
            "n_pos": np.zeros(int(summ.n_pos_entries) + 8, np.uint16), "hdr_lens": np.zeros(nr + 8, np.uint16),
            "headers": np.zeros(int(summ.hdr_bytes) + 64, np.uint8),
        }
        self.compress_fetch(arenas)
        return arenas

    def compress_dev_arenas(self) -> DecArenas:
        v = DecArenas()
        self._ck(self.L.fq28_compress_dev_arenas(self.h, C.byref(v)))
        return v

    # ------------------------------------------------------------ decompress
    def decompress(self, arenas: dict, infos, n_chunks: int, headers: np.ndarray, n_records: int,
                   out: np.ndarray | None = None, n_pos_entries: int | None = None) -> np.ndarray:
        total = sum(int(infos[k].total) for k in range(n_chunks))
        if out is None:
            out = np.zeros(total, np.uint8)
        da = DecArenas()
        da.seq, da.seq_bytes = _ptr(arenas["seq"]), arenas["seq"].size
        da.qual, da.qual_bytes = _ptr(arenas["qual"]), arenas["qual"].size
        da.readlens = _ptr(arenas["readlens"])
        da.n_count = _ptr(arenas["n_count"])
        da.n_pos = _ptr(arenas["n_pos"])
        da.n_pos_entries = arenas["n_pos"].size if n_pos_entries is None else n_pos_entries
        da.hdr_lens = _ptr(arenas["hdr_lens"])
        headers = np.ascontiguousarray(headers, dtype=np.uint8)
        da.headers, da.headers_bytes = _ptr(headers), headers.size
        da.n_records = n_records
        wrote = C.c_size_t(0)
        self._ck(self.L.fq28_decompress(self.h, C.byref(da), infos, n_chunks, _ptr(out), out.size, C.byref(wrote)))
        return out[: wrote.value]

    def decompress_dev(self, da: DecArenas, infos, n_chunks: int, d_out: int, out_cap: int) -> int:
        wrote = C.c_size_t(0)
        self._ck(self.L.fq28_decompress_dev(self.h, C.byref(da), infos, n_chunks, d_out, out_cap, C.byref(wrote)))
        return wrote.value
