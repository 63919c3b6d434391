// fq28_decode.cu -- K7 (FASTQ re-layout) and K6 (tANS decode + N re-insertion)
// for a batch of chunks.  Replaces DecompressionWorkspace::decodeChunk
// (src/workspace.cpp:47-88), prepareFastqChunk (src/workspace.h:127-133),
// SequenceDecoder::decodeRecord (src/fse_sequence.cpp:114-143),
// QualityDecoder::decodeRecord (src/fse_quality.cpp:55-67) and
// FSE_Decoder::startChunk/endChunk (src/fse_common.hpp:130-141).
//
// A chunk stream is inherently serial (the next context and the next bit
// offset both depend on the symbol just decoded), so parallelism is
// (#chunks x 2 stream types): one thread per stream, the per-context decoder
// states of that stream in shared memory, DTables read through L1/L2 at a
// fixed stride (cell = ctx << 11 | state).
#include "fq28_internal.cuh"

namespace fq28 {

struct DecChunk {            // per-chunk decode descriptor (device)
  uint64_t seq_off, qual_off;   // byte offsets in the stream arenas
  uint64_t out_off;             // chunk start in the output
  uint32_t seq_len, qual_len;
  uint32_t rec0, n_rec;         // record range
  uint32_t npos0, npos_len;     // n_pos segment
  uint32_t total;               // cb_original_sizes_t::total
  uint32_t pad;
};

__device__ __forceinline__ unsigned find_chunk_rec(const DecChunk *__restrict__ ch, unsigned n, unsigned r) {
  unsigned lo = 0, hi = n;
  while (hi - lo > 1) {
    const unsigned mid = (lo + hi) >> 1;
    if (ch[mid].rec0 <= r) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_rec_bytes(const uint16_t *__restrict__ readlens, const uint16_t *__restrict__ hdr_lens, size_t n,
                            uint32_t *__restrict__ rec_bytes) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) rec_bytes[r] = (uint32_t)hdr_lens[r] + 2u * readlens[r] + 5u;  // '\n' '\n' '+' '\n' '\n'
}

// decodeChunk pass 1 (src/workspace.cpp:62-80): header '\n' seq '\n' '+' '\n'
// qual '\n'.  One warp per record copies the header bytes and writes the five
// separator bytes; seq/qual slots are filled by the decode kernels.
constexpr int LAY_WARPS = 8;
__global__ void __launch_bounds__(LAY_WARPS * 32)
k_layout(const DecChunk *__restrict__ ch, unsigned n_chunks, const uint32_t *__restrict__ recscan,
         const uint32_t *__restrict__ hdrscan, const uint16_t *__restrict__ readlens,
         const uint16_t *__restrict__ hdr_lens, const uint8_t *__restrict__ headers, size_t n_rec,
         char *__restrict__ out) {
  const unsigned lane = threadIdx.x & 31;
  const size_t r = (size_t)blockIdx.x * LAY_WARPS + (threadIdx.x >> 5);
  if (r >= n_rec) return;
  const unsigned k = find_chunk_rec(ch, n_chunks, (unsigned)r);
  char *dst = out + ch[k].out_off + (recscan[r] - recscan[ch[k].rec0]);
  const unsigned hl = hdr_lens[r], L = readlens[r];
  const uint8_t *src = headers + hdrscan[r];
  for (unsigned i = lane; i < hl; i += 32) dst[i] = (char)src[i];
  if (lane == 0) {
    dst[hl] = '\n';
    dst[hl + 1 + L] = '\n';
    dst[hl + 2 + L] = '+';
    dst[hl + 3 + L] = '\n';
    dst[hl + 4 + 2 * L] = '\n';
  }
}

// Q7: when the input had text after '+', `total` exceeds what is laid out and
// the reference emits the value-initialised (NUL) tail of raw_data
// (src/workspace.h:130).  One warp per chunk.
__global__ void k_chunk_tail(const DecChunk *__restrict__ ch, unsigned n_chunks, const uint32_t *__restrict__ recscan,
                             char *__restrict__ out, DevStatus *st) {
  const unsigned k = blockIdx.x;
  const uint32_t laid = recscan[ch[k].rec0 + ch[k].n_rec] - recscan[ch[k].rec0];
  if (laid > ch[k].total) {
    if (threadIdx.x == 0) set_error(st, FQ28_ERR_FORMAT, k);
    return;
  }
  for (uint32_t i = laid + threadIdx.x; i < ch[k].total; i += blockDim.x) out[ch[k].out_off + i] = 0;
}

// ---- backward bit reader (BIT_DStream_t, Appendix A.6) -----------------------
// Stream bit i lives at bit (mis + i) of the 32-bit word array w (the stream's
// address rounded down to 4 bytes).  buf caches bits [base, base+64).
struct BitReader {
  const uint32_t *w;
  unsigned long long buf;
  long long base, pos, floor_;
  bool bad;
  __device__ __forceinline__ void init(const uint8_t *p, uint32_t len) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    floor_ = (long long)(a & 3) * 8;
    bad = false;
    const unsigned last = len ? p[len - 1] : 0u;
    if (last == 0) { bad = true; pos = floor_; base = 0; buf = 0; return; }  // BIT_initDStream: no end mark
    pos = floor_ + (long long)(len - 1) * 8 + (31 - __clz(last));
    base = ((pos - 1) >> 5 << 5) - 32;
    if (base < 0) base = 0;
    buf = ((unsigned long long)w[(base >> 5) + 1] << 32) | w[base >> 5];
  }
  __device__ __forceinline__ unsigned read(unsigned nb) {
    if (nb == 0) return 0;
    if (pos - floor_ < (long long)nb) { bad = true; pos = floor_; return 0; }
    pos -= nb;
    const unsigned v = (unsigned)(buf >> (pos - base)) & ((1u << nb) - 1u);
    if (pos - base < 32 && base >= 32) {
      base -= 32;
      buf = (buf << 32) | w[base >> 5];
    }
    return v;
  }
  __device__ __forceinline__ bool finished() const { return !bad && pos == floor_; }
};

// One thread per chunk stream.  STREAMS threads per CTA; decoder states for
// context c of stream s at states[c * STREAMS + s].
template <unsigned N, unsigned STREAMS, bool IS_SEQ>
__global__ void __launch_bounds__(STREAMS)
k_decode(const DecChunk *__restrict__ ch, unsigned n_chunks, const uint8_t *__restrict__ arena,
         const uint32_t *__restrict__ logs, const uint32_t *__restrict__ dtab_fix,
         const uint32_t *__restrict__ recscan, const uint16_t *__restrict__ readlens,
         const uint16_t *__restrict__ hdr_lens, char *__restrict__ out, DevStatus *st) {
  extern __shared__ uint16_t states[];
  const unsigned s = threadIdx.x;
  const unsigned k = blockIdx.x * STREAMS + s;
  if (k >= n_chunks) return;
  const DecChunk c = ch[k];
  BitReader br;
  br.init(arena + (IS_SEQ ? c.seq_off : c.qual_off), IS_SEQ ? c.seq_len : c.qual_len);
  // FSE_Decoder::startChunk: states for ctx N-1 .. 0 (src/fse_common.hpp:134-138)
  for (unsigned i = N; i > 0; --i) states[(i - 1) * STREAMS + s] = (uint16_t)br.read(logs[i - 1]);
  const uint32_t scan0 = recscan[c.rec0];
  // records n-1 .. 0 (src/workspace.cpp:84-87)
  for (unsigned rr = c.n_rec; rr > 0; --rr) {
    const unsigned r = c.rec0 + rr - 1;
    const unsigned L = readlens[r], hl = hdr_lens[r];
    char *dst = out + c.out_off + (recscan[r] - scan0) + hl + 1 + (IS_SEQ ? 0u : L + 3u);
    if (IS_SEQ) {
      unsigned ctx = SEQ_INITIAL_CTX;
      for (unsigned i = 0; i < L; i++) {
        const unsigned sidx = ctx * STREAMS + s;
        const unsigned e = __ldg(&dtab_fix[(ctx << FIX_LOG) + states[sidx]]);
        const unsigned sym = (e >> 16) & 3u;
        states[sidx] = (uint16_t)((e & 0xFFFFu) + br.read(e >> 24));
        dst[i] = (char)((0x54474341u >> (8 * sym)) & 0xFFu);  // "ACGT"
        ctx = (ctx >> 2) + (sym << 6);  // addSymUpper
      }
    } else {
      unsigned ctx = qual_ctx(0, 0, 0), q1 = 0, q2 = 0;
      for (unsigned i = 0; i < L; i++) {
        const unsigned sidx = ctx * STREAMS + s;
        const unsigned e = __ldg(&dtab_fix[(ctx << FIX_LOG) + states[sidx]]);
        const unsigned q = (e >> 16) & 63u;
        states[sidx] = (uint16_t)((e & 0xFFFFu) + br.read(e >> 24));
        dst[i] = (char)(q + QUAL_OFFSET);
        ctx = qual_ctx(q, q1, q2);
        q2 = q1;
        q1 = q;
      }
    }
  }
  if (!br.finished()) set_error(st, FQ28_ERR_STREAM, k);  // BIT_endOfDStream, src/fse_common.hpp:141
}

// N re-insertion (src/fse_sequence.cpp:115-126,138-142): cumulative deltas.
__global__ void k_ninsert(const DecChunk *__restrict__ ch, unsigned n_chunks, const uint32_t *__restrict__ recscan,
                          const uint32_t *__restrict__ nscan, const uint16_t *__restrict__ readlens,
                          const uint16_t *__restrict__ hdr_lens, const uint16_t *__restrict__ n_count,
                          const uint16_t *__restrict__ n_pos, size_t n_rec, char *__restrict__ out, DevStatus *st) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const unsigned cnt = n_count[r];
  if (cnt == 0) return;
  const unsigned k = find_chunk_rec(ch, n_chunks, (unsigned)r);
  const uint32_t rel = nscan[r] - nscan[ch[k].rec0];
  if (rel + cnt > ch[k].npos_len) { set_error(st, FQ28_ERR_STREAM, k); return; }
  const uint16_t *np = n_pos + ch[k].npos0 + rel;
  char *dst = out + ch[k].out_off + (recscan[r] - recscan[ch[k].rec0]) + hdr_lens[r] + 1;
  const unsigned L = readlens[r];
  unsigned p = 0;
  for (unsigned i = 0; i < cnt; i++) {
    p = (p + np[i]) & 0xFFFFu;  // readlen_t arithmetic
    if (p >= L) { set_error(st, FQ28_ERR_STREAM, k); return; }
    dst[p] = 'N';
  }
}

constexpr unsigned SEQ_STREAMS = 32;
constexpr unsigned QUAL_STREAMS = 12;

int decode_batch(fq28_handle *h, const fq28_dec_arenas *in, const fq28_chunk_info *infos, size_t n_chunks,
                 char *d_out, size_t out_cap, size_t *out_bytes) {
  if (!h->seq.ready || !h->qual.ready) return fail(h, FQ28_ERR_ARG, "frequency tables not built/loaded");
  if (out_bytes) *out_bytes = 0;
  if (n_chunks == 0) return FQ28_OK;
  FQ28_CUDA(h, cudaMemsetAsync(h->d_status, 0, sizeof(DevStatus), h->stream));
  const size_t n_rec = in->n_records;
  std::vector<DecChunk> meta(n_chunks);
  uint64_t out_off = 0;
  for (size_t k = 0; k < n_chunks; k++) {
    const fq28_chunk_info &ci = infos[k];
    DecChunk &m = meta[k];
    m.seq_off = ci.seq_off; m.qual_off = ci.qual_off;
    m.seq_len = ci.seq_len; m.qual_len = ci.qual_len;
    m.out_off = out_off;
    m.rec0 = (uint32_t)ci.rec_off; m.n_rec = ci.n_records;
    m.npos0 = (uint32_t)ci.n_pos_off; m.npos_len = ci.n_pos_len;
    m.total = ci.total; m.pad = 0;
    if (ci.rec_off + ci.n_records > n_rec) return fail(h, FQ28_ERR_ARG, "chunk %zu: records out of range", k);
    if (k && ci.rec_off != infos[k - 1].rec_off + infos[k - 1].n_records)
      return fail(h, FQ28_ERR_ARG, "chunk %zu: record ranges must be contiguous", k);
    if (ci.seq_off + ci.seq_len > in->seq_bytes || ci.qual_off + ci.qual_len > in->qual_bytes)
      return fail(h, FQ28_ERR_ARG, "chunk %zu: stream out of arena", k);
    if (ci.n_pos_off + ci.n_pos_len > in->n_pos_entries) return fail(h, FQ28_ERR_ARG, "chunk %zu: n_pos out of range", k);
    out_off += ci.total;
  }
  if (infos[0].rec_off != 0) return fail(h, FQ28_ERR_ARG, "first chunk must start at record 0");
  if (out_off > out_cap) return fail(h, FQ28_ERR_CAP, "output needs %llu bytes, cap %zu", (unsigned long long)out_off, out_cap);
  if (out_off > FQ28_MAX_SLAB) return fail(h, FQ28_ERR_ARG, "batch output exceeds FQ28_MAX_SLAB");
  FQ28_TRY(ensure(h, h->dec_meta, n_chunks * sizeof(DecChunk)));
  FQ28_CUDA(h, cudaMemcpyAsync(h->dec_meta.p, meta.data(), n_chunks * sizeof(DecChunk), cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  const DecChunk *ch = h->dec_meta.as<DecChunk>();

  FQ28_TRY(ensure(h, h->dec_recout, (n_rec + 2) * 4));
  FQ28_TRY(ensure(h, h->dec_hdrin, (n_rec + 2) * 4));
  FQ28_TRY(ensure(h, h->dec_npos_off, (n_rec + 2) * 4));
  uint32_t *recscan = h->dec_recout.as<uint32_t>(), *hdrscan = h->dec_hdrin.as<uint32_t>(),
           *nscan = h->dec_npos_off.as<uint32_t>();

  stage_begin(h, ST_LAYOUT);
  if (n_rec) {
    k_rec_bytes<<<(unsigned)((n_rec + 255) / 256), 256, 0, h->stream>>>(in->readlens, in->hdr_lens, n_rec, recscan);
    FQ28_LAUNCH_CHECK(h);
  }
  FQ28_TRY(scan_exclusive_u32(h, recscan, recscan, n_rec));
  FQ28_TRY(scan_exclusive_u16_to_u32(h, in->hdr_lens, hdrscan, n_rec));
  FQ28_TRY(scan_exclusive_u16_to_u32(h, in->n_count, nscan, n_rec));
  if (n_rec) {
    k_layout<<<(unsigned)((n_rec + LAY_WARPS - 1) / LAY_WARPS), LAY_WARPS * 32, 0, h->stream>>>(
        ch, (unsigned)n_chunks, recscan, hdrscan, in->readlens, in->hdr_lens, in->headers, n_rec, d_out);
    FQ28_LAUNCH_CHECK(h);
  }
  k_chunk_tail<<<(unsigned)n_chunks, 128, 0, h->stream>>>(ch, (unsigned)n_chunks, recscan, d_out, h->d_status);
  FQ28_LAUNCH_CHECK(h);
  stage_end(h, ST_LAYOUT);

  stage_begin(h, ST_DECODE_SEQ);
  {
    const size_t smem = (size_t)SEQ_N * SEQ_STREAMS * sizeof(uint16_t);
    k_decode<SEQ_N, SEQ_STREAMS, true><<<(unsigned)((n_chunks + SEQ_STREAMS - 1) / SEQ_STREAMS), SEQ_STREAMS, smem, h->stream>>>(
        ch, (unsigned)n_chunks, in->seq, h->seq.logs, h->seq.dtab_fix, recscan, in->readlens, in->hdr_lens, d_out,
        h->d_status);
    FQ28_LAUNCH_CHECK(h);
  }
  stage_end(h, ST_DECODE_SEQ);
  stage_begin(h, ST_DECODE_QUAL);
  {
    const size_t smem = (size_t)QUAL_N * QUAL_STREAMS * sizeof(uint16_t);
    static bool attr_set = false;
    if (!attr_set) {
      FQ28_CUDA(h, cudaFuncSetAttribute(k_decode<QUAL_N, QUAL_STREAMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      attr_set = true;
    }
    k_decode<QUAL_N, QUAL_STREAMS, false><<<(unsigned)((n_chunks + QUAL_STREAMS - 1) / QUAL_STREAMS), QUAL_STREAMS, smem, h->stream>>>(
        ch, (unsigned)n_chunks, in->qual, h->qual.logs, h->qual.dtab_fix, recscan, in->readlens, in->hdr_lens, d_out,
        h->d_status);
    FQ28_LAUNCH_CHECK(h);
  }
  stage_end(h, ST_DECODE_QUAL);

  stage_begin(h, ST_NINSERT);
  if (n_rec && in->n_pos_entries) {
    k_ninsert<<<(unsigned)((n_rec + 255) / 256), 256, 0, h->stream>>>(ch, (unsigned)n_chunks, recscan, nscan, in->readlens,
                                                                     in->hdr_lens, in->n_count, in->n_pos, n_rec, d_out,
                                                                     h->d_status);
    FQ28_LAUNCH_CHECK(h);
  }
  stage_end(h, ST_NINSERT);
  FQ28_TRY(check_status(h, "decode"));
  if (out_bytes) *out_bytes = (size_t)out_off;
  return FQ28_OK;
}

}  // namespace fq28
