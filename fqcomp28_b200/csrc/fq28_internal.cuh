// fq28_internal.cuh -- shared declarations of the sm_100a implementation behind
// include/fq28.h.  Host-side orchestration is plain C++; every data-path
// operation is a CUDA kernel (there is no CPU fallback anywhere in this
// library).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/fq28.h"

namespace fq28 {

// ---- FSE constants (zstd lib/common/fse.h) -----------
constexpr unsigned FSE_MIN_TABLELOG = 5;
constexpr unsigned FSE_MAX_TABLELOG = 12;
constexpr unsigned FSE_DEFAULT_TABLELOG = 11;
constexpr unsigned FIX_LOG = 11;  // every log here is <= 11 (maxTableLog = 0 -> default 11, minBits <= 7)
constexpr unsigned SEQ_INITIAL_CTX = 0xD7;  // src/fse_sequence.h:41-63
constexpr unsigned QUAL_OFFSET = 33;        // src/fse_quality.h:24
constexpr unsigned SEQ_N = FQ28_SEQ_MODELS, SEQ_A = FQ28_SEQ_ALPHABET;
constexpr unsigned QUAL_N = FQ28_QUAL_MODELS, QUAL_A = FQ28_QUAL_ALPHABET;

// ---- partition / pack tiling ------------------------------------------------
constexpr unsigned SEQ_TILE = 8192;     // symbols per context-partition tile (seq)
constexpr unsigned QUAL_TILE = 131072;  // symbols per context-partition tile (qual)
// slots per tile region: every non-empty context run is padded to a multiple of 16
constexpr unsigned SEQ_STRIDE = SEQ_TILE + 16 * SEQ_N;
constexpr unsigned QUAL_STRIDE = 2 * QUAL_TILE;   // >= QUAL_TILE + 15 * QUAL_N
constexpr unsigned PACK_THREADS = 256;
constexpr unsigned PACK_EPT = 8;        // entries per thread in the bit packer
constexpr unsigned PACK_TILE = PACK_THREADS * PACK_EPT;

// ---- compressed sequence DTables (shared-memory resident in the decoder) ----
// A DTable cell is (symbol, nbBits, newState) with newState = (x << nbBits) - T,
// nbBits = log - hb(x), x = symbolNext[sym] + rank, rank = number of cells
// below u holding the same symbol (Appendix A.6).  So 2 bits per cell (the
// symbol) plus a two-level rank directory reproduce the cell exactly:
// 840 B per context instead of 8 KB, all 256 contexts fit in one SM's smem.
struct SeqDecTables {
  uint32_t symtab[SEQ_N][128];    // 16 two-bit symbols per word, cell u at word u>>4
  uint8_t fine[SEQ_N][64][4];     // rank of each symbol at the start of 32-cell block b, relative to its coarse block
  uint16_t coarse[SEQ_N][8][4];   // rank of each symbol at the start of 256-cell block
  uint16_t snext[SEQ_N][4];       // symbolNext: norm count (-1 counts as 1) in bits 0..11;
                                  // bits 12..15 of snext[ctx][0] hold the context's table log
};

// ---- zero-bit runs in the quality decoder -----------------------------------
// In a context whose dominant symbol d has norm > T/2, most cells decode d with
// nbBits == 0: the step reads no bits and its next state is a function of the
// state alone.  When the context maps to itself under d (ctx(d,d,d)), such
// steps chain: run table Z[x] = (k << 11) | state after the k <= 15 zero-bit
// steps that start at x; J1[x] = state after one such step (0xFFFF if the cell
// at x is not a zero-bit d cell), used when fewer than k symbols are left in
// the record.
constexpr unsigned QZ_MAX = 4;
constexpr unsigned QZ_CAP = 15;

// ---- error record written by kernels ----------------------------------------
struct DevStatus {
  int code;          // first (lowest) FQ28_ERR_* seen, 0 if none
  unsigned where;    // record / chunk index that raised it
};

// growable device buffer
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Device copy of one FreqTable plus the tables built from it
// (FSE_Encoder / FSE_Decoder ctors, src/fse_common.hpp:46-71,107-127).
constexpr unsigned QWIN_MAX_ENTRIES = 128 * 65;   // every row all 64 columns + its out-of-window word

struct DevTables {
  unsigned n_models = 0, alphabet = 0;
  uint32_t *counts = nullptr;  // [N*A] scratch for histogram input
  int16_t *norm = nullptr;     // [N*A]
  uint32_t *logs = nullptr;    // [N]
  uint32_t *max_log = nullptr; // [1]
  uint32_t *toff = nullptr;    // [N+1] first cell of each context's tables
  uint16_t *ctab = nullptr;    // next-state cells, packed by toff
  int2 *symtt = nullptr;       // [N*A] {deltaFindState, deltaNbBits}
  int8_t *dom_sym = nullptr;   // [N] symbol with norm > T/2 (nbBits in {0,1}), or -1
  uint32_t *dtab = nullptr;    // DTable cells newState | sym<<16 | nbBits<<24
  uint32_t *dtab_fix = nullptr; // same cells at fixed stride: cell (ctx << FIX_LOG) + state
  // decoder side structures
  uint32_t *logsuf = nullptr;   // [N+1] logsuf[c] = sum of logs[c'] for c' > c (initial-state bit offsets)
  uint8_t *seqdec = nullptr;    // sequence only: compressed DTables (SeqDecTables), copied to smem by the decoder
  uint16_t *cid = nullptr;      // quality only: [N] compact id of every context whose table differs from the
                                // untouched (prior-only) pattern, 0xFFFF otherwise
  uint32_t *n_touched = nullptr;   // quality only: [1] number of compact ids
  uint32_t h_n_touched = 0;
  // quality only: zero-bit run tables of the self-loop contexts ctx(d,d,d) whose
  // symbol d is dominant (QualZrun), copied to shared memory by the decoder
  uint16_t *zrun = nullptr;        // [QZ_MAX][2][1 << FIX_LOG]
  uint32_t *zinfo = nullptr;       // [0] = number of slots, [1 + j] = context of slot j
  uint32_t h_n_z = 0;
  uint32_t h_zctx[4] = {0, 0, 0, 0};
  // decoder v2 (fq28_dec2.cuh): W tables = DTable cells repacked for the cached-cell decoder.
  // sequence: [256 << 11]; quality: rows of the dense contexts only, [2 * n_v * 64 << 11]
  uint32_t *wtab = nullptr;
  uint8_t *qrk = nullptr;          // quality only: rk[64] (q -> rank in V, 0xFF outside) then vq[64] (rank -> q)
  uint32_t *qdinfo = nullptr;      // quality only: [0] = |V|, [1] = entries of the windowed layout, [2] = its rows
  uint32_t h_n_v = 0;
  // windowed layout of the cached cells (fq28_dec2.cuh, WIN): per row (rank(max) * 2 + eq) the byte offset
  // of its first entry and lo * 4 | (width * 4) << 16; W table rows in the same compact order
  uint2 *qwin = nullptr;           // [128]
  uint32_t *wtabw = nullptr;       // [QWIN_MAX_ENTRIES << 11]
  uint32_t h_n_win = 0, h_n_rows = 0;
  size_t cells_cap = 0;
  bool ready = false;
};

enum Stage {
  ST_PARSE = 0,   // K1 newline scan + record table + chunk walk
  ST_EXTRACT,     // K2 symbols / contexts / N positions
  ST_PART_SEQ,    // context partition, sequence (fused hist + stable rank)
  ST_PART_QUAL,   // context partition, quality (tile hist + stable rank)
  ST_CHAIN_SEQ,   // K5 tANS state chains, sequence
  ST_CHAIN_QUAL,  // K5 tANS state chains, quality
  ST_PACK_SEQ,    // K5 bit offsets + bit packing, sequence
  ST_PACK_QUAL,   // K5 bit offsets + bit packing, quality
  ST_LAYOUT,      // K7 FASTQ re-layout
  ST_DECODE_SEQ,  // K6 tANS decode, sequence
  ST_DECODE_QUAL, // K6 tANS decode, quality
  ST_NINSERT,     // N re-insertion
  ST_HIST,        // K3
  ST_TABLES,      // K4
  ST_COUNT
};

}  // namespace fq28

struct fq28_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  uint64_t launches = 0;

  fq28::DevTables seq, qual;
  fq28::DevStatus *d_status = nullptr;   // device
  fq28::DevStatus *h_status = nullptr;   // pinned host mirror
  uint64_t *d_scalars = nullptr;         // small device scratch (64 x u64)
  uint64_t *h_scalars = nullptr;         // pinned host mirror

  // parse results (valid for the slab of the last parse)
  const char *d_fastq = nullptr;         // slab being processed (not owned unless == in_fastq.p)
  size_t n_bytes = 0, n_lines = 0, n_rec = 0;
  fq28::DevBuf in_fastq;                 // staging for host-buffer entry points
  fq28::DevBuf tile_cnt, nl, hdr_off, seq_off, qual_off, len, hdr_len, symoff;
  fq28::DevBuf chunk_rec;                // u32 [cap+1]
  // pinned host scratch for the small device->host reads of the per-call paths (a pageable
  // destination makes cudaMemcpyAsync wait for the stream inside the runtime, which stalls the
  // CUDA calls of other host threads -- the two lanes of the pipelined compress)
  void *h_pin = nullptr;
  size_t h_pin_cap = 0;
  std::vector<uint32_t> h_chunk_rec;     // host copy
  std::vector<uint32_t> h_chunk_sym;     // symoff at chunk boundaries
  std::vector<uint32_t> h_chunk_byte;    // hdr_off at chunk boundaries
  size_t n_chunks = 0;
  size_t chunk_stride = 0;               // entries per array inside chunk_rec

  // encode work buffers
  fq28::DevBuf n_count, npos_off, n_pos, hdrscan, hdr_arena;
  fq28::DevBuf key_seq, key_qual, perm_seq, perm_qual, ssym_seq, ssym_qual, out_seq, out_qual;
  fq28::DevBuf tile0_seq, tile0_qual, tbase_seq, tbase_qual, fstate_seq, fstate_qual;
  fq28::DevBuf ptile0_seq, ptile0_qual, pbits_seq, pbits_qual, pscan_seq, pscan_qual;
  fq28::DevBuf arena_seq, arena_qual, d_infos, scan_tmp, scan_tmp_side, dom_list, present;
  fq28_enc_summary last_summary{};
  bool have_result = false;

  // decode work buffers
  fq28::DevBuf dec_in[8], dec_out, dec_recout, dec_hdrin, dec_npos_off, dec_meta, dec_cold;

  // side stream: the sequence and quality pipelines are independent
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // host-buffer compress of a large slab runs as overlapped parts: all host->device copies are
  // queued on copy_stream (one event per part), the parts are encoded in turn by this handle and
  // sibling handles (own streams and buffers), each driven by its own host thread (a "lane")
  std::vector<fq28_handle *> siblings;
  bool borrowing = false;                // sibling only: seq / qual are shallow copies of the owner's tables,
  fq28::DevTables own_seq, own_qual;     // its own (allocated at create) are kept here
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> ev_part;
  fq28::DevBuf in_raw;                   // pipelined compress: the slab as copied from the host (before alignment);
                                         // fq28_stage: the speculatively copied host range
  // fq28_stage: host range [stage_host, stage_host + stage_bytes) already copied (or being copied,
  // same stream) to in_raw; host-buffer entry points whose input lies inside it skip their H2D
  const char *stage_host = nullptr;
  size_t stage_bytes = 0;
  // fq28_plan / fq28_plan_dev: parse + chunk walk done for exactly this slab; a following
  // fq28_compress(_dev) with the same arguments and sample_bytes == 0 goes straight to the encode
  struct Plan { bool valid = false; const char *d_fastq = nullptr; size_t n_bytes = 0, reading_size = 0; int eof = 0; } plan;
  // fq28_preparse_dev also starts the field separation (k_extract needs the record table, not the
  // chunk boundaries) on a stream of its own, so that it runs while the caller waits for the cut
  // that the previous rank owes it; the boundary walk then goes to the high-priority side stream
  cudaStream_t bulk = nullptr;
  cudaEvent_t ev_parsed = nullptr, ev_extract = nullptr;
  bool extracted = false;                // keys / n_count of all h->n_rec parsed records are (being) written
  // fq28_preparse_dev: the record table of exactly this slab is in place (consumed by the next plan)
  struct Parsed { bool valid = false; const char *d_fastq = nullptr; size_t n_bytes = 0; } parsed;
  // generation of the device tables (bumped whenever they are rebuilt; the C++ facade compares it)
  uint64_t tables_gen = 0;
  const char *plan_host = nullptr;       // host slab of the last fq28_plan
  cudaEvent_t ev_copy = nullptr;

  // policy knobs, read from the environment once at fq28_create (diagnostics)
  struct Cfg {
    bool seq_v1 = false;           // FQ28_SEQ_V1 / FQ28_DEC_V1: round-1 sequence decoder (rank-directory tables)
    bool qual_v2 = true;           // cached-cell quality decoder (fq28_dec2.cuh); FQ28_QUAL_V1 / FQ28_DEC_V1: the round-1 state-table one
    bool dec_serial = false;       // FQ28_DEC_SERIAL: the two decode kernels one after the other (per-kernel timing)
    bool share_sms = false;        // FQ28_DEC_SHARE_SMS: do not keep the sequence decoder on SMs of its own
    bool force_win = false;        // FQ28_QUAL_WINDOWED: ... the windowed layout whenever the tables have one (A/B, tests)
    bool no_eager = false;         // FQ28_NO_EAGER_EXTRACT: fq28_preparse(_dev) does not start the field separation
    bool no_win = false;           // FQ28_QUAL_DENSE: many-valued qualities keep the dense layout of the cached cells (A/B)
    bool dec_concurrent = false;   // FQ28_DEC_CONCURRENT: never serialise the two decode kernels
    unsigned seq_lanes = 0, seq_warps = 0, qual_lanes = 0, qual_warps = 0;  // 0 = automatic
    int qual_carveout = -2;        // -2 = automatic
    bool no_zrun = false, no_dom = false, no_rankc = false, serial = false, full_overlap = false;
    size_t pipe_min_bytes = (size_t)256 << 20;   // FQ28_PIPE_MIN_MB: host-buffer slabs from this size on are pipelined
    bool pipe_trace = false;                     // FQ28_PIPE_TRACE: host-clock timeline of the parts to stderr
    unsigned pipe_lanes = 4;                     // FQ28_PIPE_LANES: parts in flight (handles and host threads)
    unsigned pipe_parts = 8;                     // FQ28_PIPE_PARTS: ... in this many parts (1 = one piece)
  } cfg;
  int qualw_carve_set = -1000;     // ... and for the windowed cached-cell quality decoder
  int qual_carve_set = -1;         // last shared-memory carve-out set for k_decode_qual on this device

  // timings
  struct EvRec { int stage; cudaEvent_t a, b; };
  std::vector<EvRec> ev_pool;
  size_t ev_used = 0;
  bool timing = true;
  float stage_ms[fq28::ST_COUNT]{};
};

namespace fq28 {

// ---- host helpers (fq28_api.cu) ---------------------------------------------
int fail(fq28_handle *h, int code, const char *fmt, ...);
int cuda_fail(fq28_handle *h, cudaError_t e, const char *what);
int ensure(fq28_handle *h, DevBuf &b, size_t bytes);
int ensure_pinned(fq28_handle *h, size_t bytes);      // h->h_pin holds at least `bytes`
int check_status(fq28_handle *h, const char *what, cudaStream_t on = nullptr);   // syncs (h->stream unless told) + reads d_status
void stage_reset(fq28_handle *h);
void stage_begin(fq28_handle *h, Stage s);
void stage_end(fq28_handle *h, Stage s);
// side stream: fork = side waits for everything queued on the main stream,
// join = main waits for the side stream; side_stage_* time a stage on it
int side_fork(fq28_handle *h);
int side_join(fq28_handle *h);
int extract_eager(fq28_handle *h);   // fq28_encode.cu
int extract_eager(fq28_handle *h);   // fq28_encode.cu
void side_stage_begin(fq28_handle *h, Stage s);
void side_stage_end(fq28_handle *h, Stage s);
// re-entrant form for stages that overlap on two streams: open returns a slot
int stage_open(fq28_handle *h, Stage s, cudaStream_t strm);
void stage_close(fq28_handle *h, int slot, cudaStream_t strm);

#define FQ28_CUDA(h, call)                                         \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) return fq28::cuda_fail((h), e__, #call); \
  } while (0)
#define FQ28_TRY(expr)            \
  do {                            \
    int r__ = (expr);             \
    if (r__ != FQ28_OK) return r__; \
  } while (0)
#define FQ28_LAUNCH_CHECK(h)                                  \
  do {                                                        \
    (h)->launches++;                                          \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return fq28::cuda_fail((h), e__, "kernel launch"); \
  } while (0)

// ---- scans (fq28_scan.cu) ---------------------------------------------------
// out[i] = sum(in[0..i)), out[n] = total.  `out` needs n+1 entries.  in may
// alias out only if types match.  Uses h->scan_tmp.
// `side` = run on h->side with its own scratch (for the overlapped pipelines).
int scan_exclusive_u16_to_u32(fq28_handle *h, const uint16_t *in, uint32_t *out, size_t n, bool side = false);
int scan_exclusive_u32(fq28_handle *h, const uint32_t *in, uint32_t *out, size_t n, bool side = false);
int scan_exclusive_u32_to_u64(fq28_handle *h, const uint32_t *in, uint64_t *out, size_t n, bool side = false);

// ---- stages -----------------------------------------------------------------
// fq28_parse.cu: K1.  Fills h->nl .. h->symoff, h->n_lines, h->n_rec.
int parse_slab(fq28_handle *h, const char *d_fastq, size_t n_bytes, bool need_symoff);
// chunk walk (A9); fills h->chunk_rec, h_chunk_*, n_chunks
int split_slab(fq28_handle *h, size_t reading_size, bool eof, size_t max_chunks, size_t first_cut = 0);
// fq28_tables.cu: K3 / K4
int tokenize_headers(fq28_handle *h, const uint8_t *headers, size_t headers_bytes, const uint16_t *hdr_lens, size_t n_rec,
                     const uint64_t *chunk_rec, size_t n_chunks, const fq28_hdr_format *fmt, uint8_t *arena, size_t arena_cap,
                     fq28_hdr_field_info *infos, size_t *arena_bytes);
int detokenize_headers(fq28_handle *h, const uint8_t *arena, size_t arena_bytes, const fq28_hdr_field_info *infos,
                       const uint64_t *chunk_rec, size_t n_chunks, const fq28_hdr_format *fmt, uint8_t *headers_out,
                       size_t headers_cap, uint16_t *hdr_lens_out, size_t *headers_bytes);
int hist_slab(fq28_handle *h, uint32_t *d_seq_counts, uint32_t *d_qual_counts);
int tables_alloc(fq28_handle *h, DevTables &t, unsigned n_models, unsigned alphabet);
int tables_from_counts(fq28_handle *h, DevTables &t, const uint32_t *d_counts);
int tables_from_norm(fq28_handle *h, DevTables &t);
// fq28_encode.cu: K2 + K5
int encode_slab(fq28_handle *h, fq28_chunk_info *infos, size_t infos_cap, fq28_enc_summary *summary);
// per-device kernel attributes (dynamic shared memory opt-in); called by fq28_create
int encode_init_device(fq28_handle *h);
int decode_init_device(fq28_handle *h);
// fq28_decode.cu: K6 + K7
int decode_batch(fq28_handle *h, const fq28_dec_arenas *d_in, const fq28_chunk_info *infos,
                 size_t n_chunks, char *d_out, size_t out_cap, size_t *out_bytes);

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ void set_error(DevStatus *st, int code, unsigned where) {
  // keep the first error by (where) order for determinism: lowest record wins
  int old = atomicCAS(&st->code, 0, code);
  if (old == 0) st->where = where;
  else atomicMin(&st->where, where);
}

__device__ __forceinline__ unsigned qual_ctx(unsigned q, unsigned q1, unsigned q2) {
  // FSE_Quality::calcContext, src/fse_quality.h:40-44
  unsigned ctx = (((q1 > q2 ? q1 : q2) << 6) + q) & 0xFFFu;
  return ctx + ((unsigned)(q1 == q2) << 12);
}

__device__ __forceinline__ int base2bits(unsigned char c) {
  // src/fse_sequence.cpp:6-14; -1 for anything outside ACGT
  // A=0x41 C=0x43 G=0x47 T=0x54
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

}  // namespace fq28
