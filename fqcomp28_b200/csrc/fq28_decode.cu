// fq28_decode.cu -- K7 (FASTQ re-layout) and K6 (tANS decode + N re-insertion)
// for a batch of chunks.  Replaces DecompressionWorkspace::decodeChunk
// (src/workspace.cpp:47-88), prepareFastqChunk (src/workspace.h:127-133),
// SequenceDecoder::decodeRecord (src/fse_sequence.cpp:114-143),
// QualityDecoder::decodeRecord (src/fse_quality.cpp:55-67) and
// FSE_Decoder::startChunk/endChunk (src/fse_common.hpp:130-141).
//
// A chunk stream is inherently serial (the next context and the next bit
// offset both depend on the symbol just decoded), so parallelism is
// (#chunks x 2 stream types) and the time of a batch is
//     symbols per stream  x  latency of one symbol
// until the machine runs out of issue slots.  Design:
//  * one THREAD per stream, 32 streams per warp: a decoded symbol costs 1/32 of
//    a warp instruction, so the kernels stay latency-bound (not issue-bound)
//    up to tens of thousands of streams per GPU;
//  * lockstep across the 32 streams of a warp is only harmless if every step
//    has the same latency, so the per-symbol table lookups must not go to L2:
//      - sequence: all 256 DTables live in shared memory in compressed form
//        (SeqDecTables, 832 B per context) -- a cell is rebuilt from its 2-bit
//        symbol plus a two-level rank directory, in 32-bit integer ops;
//      - quality: decoder states sit in shared memory under compact ids of the
//        contexts that have a real table (a few hundred of 8192), which leaves
//        most of the SM's 256 KB as L1 for the hot DTable cells;
//  * the bit reader keeps the next <= 64 stream bits left-aligned in two 32-bit
//    registers: taking nbBits is ONE funnel shift on the critical path, and the
//    next stream word is always prefetched one refill ahead;
//  * the two stream types run concurrently on two CUDA streams.
#include <stdlib.h>

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
                            uint32_t *__restrict__ rec_bytes, DevStatus *st) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  rec_bytes[r] = (uint32_t)hdr_lens[r] + 2u * readlens[r] + 5u;  // '\n' '\n' '+' '\n' '\n'
  if (readlens[r] == 0) set_error(st, FQ28_ERR_SHORT, (unsigned)r);  // the encoder never emits empty reads
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
// (src/workspace.h:130).  One CTA per chunk.
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

// ---- backward bit reader (BIT_DStream_t, Appendix A.6), one per lane --------
// Stream bit i lives at bit (floor_ + i) of the 32-bit word array w (the
// stream's address rounded down to 8 bytes).  The unconsumed stream is
// [floor_, P).  {hi:lo} holds the next `avail` bits [P-avail, P), left-aligned
// (bit P-1 is bit 31 of hi), with P-avail always a multiple of 32 so that a
// refill appends exactly one word, `nxw`.
// Words reach `nxw` through a two-slot per-lane ring of word PAIRS in shared
// memory that is filled with cp.async (global -> shared, 8 bytes) two refills
// ahead.  The global prefetch must NOT land in a register: the 32 streams of a
// warp run in lockstep, and a load pending in a register on behalf of one lane
// would stall (scoreboard) the refill another lane executes in the very next
// iteration -- one L2 round trip per symbol.  cp.async has no destination
// register; the ring -> nxw move is a shared-memory load issued one refill
// (>= 1 iteration) before its use, so it is off the critical path too.
constexpr unsigned RING_WORDS = 2 * 32 * 2;  // [slot][lane][2 words]
struct LaneBitReader {
  const uint32_t *w;
  uint32_t hi, lo, nxw;
  int avail;
  int wq;                  // word moved from the ring into nxw by the next refill
  int last_word;
  long long remaining;     // unconsumed stream bits, < 0 after an underflow
  long long floor_, top_bit;
  volatile uint32_t *ring; // this lane's pair in slot 0; slot 1 is ring + 64

  __device__ __forceinline__ uint32_t load_word(int idx) const {
    return (idx >= 0 && idx <= last_word) ? __ldg(w + idx) : 0u;
  }
  __device__ __forceinline__ void prefetch_pair(int pidx) {  // words 2p, 2p+1 -> slot p & 1
    volatile uint32_t *slot = ring + ((pidx & 1) << 6);
    if (pidx >= 0 && 2 * pidx <= last_word) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(const_cast<uint32_t *>(slot));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(w + 2 * pidx) : "memory");
    } else {
      slot[0] = 0u;
      slot[1] = 0u;
    }
  }
  // position the reader so that the unconsumed stream is [floor_, P)
  __device__ __forceinline__ void seek(long long P) {
    remaining = P - floor_;
    const int tw = (int)((P - 1) >> 5);                 // word holding bit P-1
    const unsigned kbits = (unsigned)(P - ((long long)tw << 5));  // 1..32 valid bits in it
    const uint32_t wt = load_word(tw), w1 = load_word(tw - 1);
    hi = __funnelshift_l(w1, wt, 32 - kbits);
    lo = w1 << (32 - kbits);
    avail = 32 + (int)kbits;
    nxw = load_word(tw - 2);
    wq = tw - 3;
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    prefetch_pair(wq >> 1);
    prefetch_pair((wq >> 1) - 1);
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  }
  __device__ __forceinline__ bool init(const uint8_t *p, uint32_t len, uint32_t *ring_lane) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)7);
    floor_ = (long long)(a & 7) * 8;
    ring = ring_lane;
    hi = lo = nxw = 0; avail = 64; wq = -1; remaining = 0; last_word = 0; top_bit = floor_;
    ring[0] = ring[1] = ring[64] = ring[65] = 0u;
    const unsigned last = len ? p[len - 1] : 0u;
    if (last == 0) return false;  // BIT_initDStream: no end mark
    last_word = (int)((floor_ + (long long)(len - 1) * 8) >> 5);
    top_bit = floor_ + (long long)(len - 1) * 8 + (31 - __clz(last));
    if (top_bit > floor_) seek(top_bit);
    return true;
  }
  // BIT_readBits: nb <= 16.  Critical path: one funnel shift.
  __device__ __forceinline__ unsigned read(unsigned nb) {
    const unsigned v = __funnelshift_l(hi, 0u, nb);   // top nb bits of hi (0 when nb == 0)
    hi = __funnelshift_l(lo, hi, nb);
    lo <<= nb;
    avail -= (int)nb;
    remaining -= nb;
    if (avail <= 32) {
      const unsigned s = 32u - (unsigned)avail;       // 0..15
      hi |= __funnelshift_l(nxw, 0u, s);
      lo |= nxw << s;
      avail += 32;
      const int pidx = wq >> 1;
      if (wq & 1) {  // first touch of pair pidx (copy issued two refills ago); the other slot is free
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        nxw = ring[((pidx & 1) << 6) + 1];
        prefetch_pair(pidx - 1);
      } else {
        nxw = ring[(pidx & 1) << 6];
      }
      --wq;
    }
    return v;
  }
  __device__ __forceinline__ bool finished() const { return remaining == 0; }
};

// 16-bit field i (0..3) of a uint2
__device__ __forceinline__ unsigned pick_u16x4(uint2 v, unsigned i) {
  return __byte_perm(v.x, v.y, 0x4410u + 0x22u * i) & 0xFFFFu;  // bytes (2i, 2i+1)
}

constexpr size_t SEQ_TAB_BYTES = (sizeof(SeqDecTables) + 15) & ~(size_t)15;
constexpr size_t SEQ_DEC_SMEM = SEQ_TAB_BYTES + (size_t)SEQ_N * 32 * sizeof(uint16_t) + RING_WORDS * sizeof(uint32_t);

// record cursor: walks records n-1 .. 0 (src/workspace.cpp:84-87); the next
// record's metadata is requested a whole record ahead
struct RecMeta { unsigned L, hl; uint32_t scan; };
__device__ __forceinline__ RecMeta load_meta(const uint16_t *__restrict__ readlens, const uint16_t *__restrict__ hdr_lens,
                                            const uint32_t *__restrict__ recscan, unsigned r) {
  RecMeta m;
  m.L = readlens[r]; m.hl = hdr_lens[r]; m.scan = recscan[r];
  return m;
}

// ---------------------------------------------------------------------------
// sequence decoder: CTA = one warp = 32 streams, one CTA per SM
// (shared memory: compressed tables 210 KB + states u16[256][32] 16 KB + ring)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 1)
k_decode_seq(const DecChunk *__restrict__ ch, unsigned n_chunks, const uint8_t *__restrict__ arena,
             const SeqDecTables *__restrict__ gtab, const uint32_t *__restrict__ logsuf,
             const uint32_t *__restrict__ recscan, const uint16_t *__restrict__ readlens,
             const uint16_t *__restrict__ hdr_lens, char *__restrict__ out, DevStatus *st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned char *tb = smem_raw;
  uint16_t *S = reinterpret_cast<uint16_t *>(smem_raw + SEQ_TAB_BYTES);  // S[ctx * 32 + lane]
  uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + SEQ_TAB_BYTES + (size_t)SEQ_N * 32 * sizeof(uint16_t));
  {  // cooperative copy of the compressed tables (uint4 granularity)
    const uint4 *src = reinterpret_cast<const uint4 *>(gtab);
    uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
    for (unsigned i = threadIdx.x; i < sizeof(SeqDecTables) / 16; i += 32) dst[i] = __ldg(src + i);
  }
  __syncwarp();
  const unsigned lane = threadIdx.x;
  const unsigned k = blockIdx.x * 32 + lane;
  const bool live = k < n_chunks;
  DecChunk c;
  c.n_rec = 0; c.rec0 = 0; c.out_off = 0; c.seq_off = 0; c.seq_len = 0;
  if (live) c = ch[k];
  LaneBitReader br;
  bool ok = true;
  if (live) {
    ok = br.init(arena + c.seq_off, c.seq_len, ring + lane * 2);
    // FSE_Decoder::startChunk (src/fse_common.hpp:134-138): states for ctx N-1 .. 0
    if (!ok || br.top_bit - br.floor_ < (long long)logsuf[SEQ_N]) {
      ok = false;
      c.n_rec = 0;
    } else {
      for (unsigned cc = SEQ_N; cc > 0; --cc) {
        const unsigned lg = *reinterpret_cast<const uint16_t *>(tb + offsetof(SeqDecTables, snext) + (cc - 1) * 8) >> 12;
        S[(cc - 1) * 32 + lane] = (uint16_t)br.read(lg);
      }
    }
  }
  const uint32_t scan0 = live ? recscan[c.rec0] : 0u;
  unsigned rr = c.n_rec;  // records left, current one included
  RecMeta cur{0, 0, 0}, nxt{0, 0, 0};
  if (rr) cur = load_meta(readlens, hdr_lens, recscan, c.rec0 + rr - 1);
  if (rr > 1) nxt = load_meta(readlens, hdr_lens, recscan, c.rec0 + rr - 2);
  char *dst = out + c.out_off + (cur.scan - scan0) + cur.hl + 1;
  unsigned i = 0, ctx = SEQ_INITIAL_CTX;
  for (;;) {
    // uniform trip count: symbols until the first lane reaches a record end
    const unsigned n = __reduce_min_sync(0xffffffffu, rr > 0 ? cur.L - i : 0xFFFFFFFFu);
    if (n == 0xFFFFFFFFu) break;
    if (rr > 0) {
      for (unsigned t = 0; t < n; t++) {
        const unsigned s0 = S[ctx * 32 + lane];
        // independent of the state
        const uint2 nx = *reinterpret_cast<const uint2 *>(tb + offsetof(SeqDecTables, snext) + ctx * 8);
        // 32-cell block of the state
        const unsigned blk = s0 >> 5, p = s0 & 31;
        const uint2 wv = *reinterpret_cast<const uint2 *>(tb + offsetof(SeqDecTables, symtab) + ctx * 512 + blk * 8);
        const unsigned fr = *reinterpret_cast<const unsigned *>(tb + offsetof(SeqDecTables, fine) + ctx * 256 + blk * 4);
        const uint2 cr = *reinterpret_cast<const uint2 *>(tb + offsetof(SeqDecTables, coarse) + ctx * 64 + (s0 >> 8) * 8);
        const unsigned lg = (nx.x >> 12) & 15u;
        // masks of the cells below p in each word: depend on p only
        const unsigned m0 = p >= 16 ? 0xFFFFFFFFu : ((1u << (2 * p)) - 1u);
        const unsigned m1 = p > 16 ? ((1u << (2 * (p - 16))) - 1u) : 0u;
        const unsigned wsel = (p & 16) ? wv.y : wv.x;
        const unsigned sym = (wsel >> ((p & 15) * 2)) & 3u;
        const unsigned pat = sym * 0x55555555u;
        const unsigned x0 = wv.x ^ pat, x1 = wv.y ^ pat;
        const unsigned e0 = ~(x0 | (x0 >> 1)) & 0x55555555u & m0;
        const unsigned e1 = ~(x1 | (x1 >> 1)) & 0x55555555u & m1;
        const unsigned rank = __popc(e0) + __popc(e1) + ((fr >> (8 * sym)) & 0xFFu) + pick_u16x4(cr, sym);
        const unsigned xs = (pick_u16x4(nx, sym) & 0xFFFu) + rank;  // symbolNext + rank
        const unsigned nb = lg - (31u - (unsigned)__clz(xs));       // Appendix A.6
        const unsigned ns = (xs << nb) - (1u << lg);
        S[ctx * 32 + lane] = (uint16_t)(ns + br.read(nb));
        dst[i + t] = (char)((0x54474341u >> (8 * sym)) & 0xFFu);    // "ACGT"
        ctx = (ctx >> 2) + (sym << 6);                              // addSymUpper
      }
      i += n;
      if (i >= cur.L) {  // record done
        --rr;
        cur = nxt;
        i = 0;
        ctx = SEQ_INITIAL_CTX;
        dst = out + c.out_off + (cur.scan - scan0) + cur.hl + 1;
        if (rr > 1) nxt = load_meta(readlens, hdr_lens, recscan, c.rec0 + rr - 2);
      }
    }
  }
  if (live && (!ok || !br.finished())) set_error(st, FQ28_ERR_STREAM, k);  // BIT_endOfDStream, src/fse_common.hpp:141
}

// ---------------------------------------------------------------------------
// quality decoder: CTA = one warp = up to 32 streams
// smem: cid map u16[8192] + ring + compact states u16[n_touched][lanes].
// Contexts without a compact id (prior-only tables: never seen in the sample)
// keep their state in a global fallback array.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_decode_qual(const DecChunk *__restrict__ ch, unsigned n_chunks, unsigned lanes, const uint8_t *__restrict__ arena,
              const uint32_t *__restrict__ logs, const uint32_t *__restrict__ logsuf,
              const uint32_t *__restrict__ dtab_fix, const uint16_t *__restrict__ gcid, unsigned n_touched,
              uint16_t *cold_states /*[n_chunks][8192]*/, const uint32_t *__restrict__ recscan,
              const uint16_t *__restrict__ readlens, const uint16_t *__restrict__ hdr_lens, char *__restrict__ out,
              DevStatus *st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t *cid = reinterpret_cast<uint16_t *>(smem_raw);
  uint32_t *ring = reinterpret_cast<uint32_t *>(cid + QUAL_N);
  uint16_t *S = reinterpret_cast<uint16_t *>(ring + RING_WORDS);  // S[id * lanes + lane]
  for (unsigned i = threadIdx.x; i < QUAL_N; i += 32) cid[i] = gcid[i];
  __syncwarp();
  const unsigned lane = threadIdx.x;
  const unsigned k = blockIdx.x * lanes + lane;
  const bool live = lane < lanes && k < n_chunks;
  DecChunk c;
  c.n_rec = 0; c.rec0 = 0; c.out_off = 0; c.qual_off = 0; c.qual_len = 0;
  if (live) c = ch[k];
  uint16_t *cold = cold_states + (size_t)(live ? k : 0) * QUAL_N;
  LaneBitReader br;
  bool ok = true;
  if (live) {
    ok = br.init(arena + c.qual_off, c.qual_len, ring + lane * 2);
    if (!ok || br.top_bit - br.floor_ < (long long)logsuf[QUAL_N]) {
      ok = false;
      c.n_rec = 0;
    } else {
      for (unsigned cc = QUAL_N; cc > 0; --cc) {
        const uint16_t v = (uint16_t)br.read(logs[cc - 1]);
        const unsigned id = cid[cc - 1];
        if (id != 0xFFFFu) S[id * lanes + lane] = v; else cold[cc - 1] = v;
      }
    }
  }
  const uint32_t scan0 = live ? recscan[c.rec0] : 0u;
  unsigned rr = c.n_rec;
  RecMeta cur{0, 0, 0}, nxt{0, 0, 0};
  if (rr) cur = load_meta(readlens, hdr_lens, recscan, c.rec0 + rr - 1);
  if (rr > 1) nxt = load_meta(readlens, hdr_lens, recscan, c.rec0 + rr - 2);
  char *dst = out + c.out_off + (cur.scan - scan0) + cur.hl + 1 + cur.L + 3;
  unsigned i = 0, ctx = qual_ctx(0, 0, 0), q1 = 0, q2 = 0;
  for (;;) {
    const unsigned n = __reduce_min_sync(0xffffffffu, rr > 0 ? cur.L - i : 0xFFFFFFFFu);
    if (n == 0xFFFFFFFFu) break;
    if (rr > 0) {
      for (unsigned t = 0; t < n; t++) {
        const unsigned id = cid[ctx];
        const bool hot = id != 0xFFFFu;
        const unsigned s0 = hot ? S[id * lanes + lane] : cold[ctx];
        const unsigned e = __ldg(&dtab_fix[(ctx << FIX_LOG) + s0]);
        const unsigned q = (e >> 16) & 63u;
        const unsigned nv = (e & 0xFFFFu) + br.read(e >> 24);
        if (hot) S[id * lanes + lane] = (uint16_t)nv; else cold[ctx] = (uint16_t)nv;
        dst[i + t] = (char)(q + QUAL_OFFSET);
        // calcContext(q, q1, q2), src/fse_quality.h:40-44
        const unsigned mx = q1 > q2 ? q1 : q2;
        ctx = (((mx << 6) + q) & 0xFFFu) + ((unsigned)(q1 == q2) << 12);
        q2 = q1;
        q1 = q;
      }
      i += n;
      if (i >= cur.L) {
        --rr;
        cur = nxt;
        i = 0;
        ctx = qual_ctx(0, 0, 0);
        q1 = q2 = 0;
        dst = out + c.out_off + (cur.scan - scan0) + cur.hl + 1 + cur.L + 3;
        if (rr > 1) nxt = load_meta(readlens, hdr_lens, recscan, c.rec0 + rr - 2);
      }
    }
  }
  if (live && (!ok || !br.finished())) set_error(st, FQ28_ERR_STREAM, k);
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
  FQ28_TRY(ensure(h, h->dec_cold, n_chunks * (size_t)QUAL_N * sizeof(uint16_t)));
  uint32_t *recscan = h->dec_recout.as<uint32_t>(), *hdrscan = h->dec_hdrin.as<uint32_t>(),
           *nscan = h->dec_npos_off.as<uint32_t>();

  stage_begin(h, ST_LAYOUT);
  if (n_rec) {
    k_rec_bytes<<<(unsigned)((n_rec + 255) / 256), 256, 0, h->stream>>>(in->readlens, in->hdr_lens, n_rec, recscan, h->d_status);
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

  // the two stream types are independent: quality runs on the side stream
  FQ28_TRY(side_fork(h));
  stage_begin(h, ST_DECODE_SEQ);
  {
    static bool attr_set = false;
    if (!attr_set) {
      FQ28_CUDA(h, cudaFuncSetAttribute(k_decode_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQ_DEC_SMEM));
      attr_set = true;
    }
    k_decode_seq<<<(unsigned)((n_chunks + 31) / 32), 32, SEQ_DEC_SMEM, h->stream>>>(
        ch, (unsigned)n_chunks, in->seq, reinterpret_cast<const SeqDecTables *>(h->seq.seqdec), h->seq.logsuf, recscan,
        in->readlens, in->hdr_lens, d_out, h->d_status);
    FQ28_LAUNCH_CHECK(h);
  }
  stage_end(h, ST_DECODE_SEQ);
  {
    const unsigned nt = h->qual.h_n_touched;
    // Streams per warp.  The DTable cells come through L1/L2, so the streams of
    // a warp wait for the slowest lane: pack only as many streams per warp as
    // are needed to keep ~2 warps per scheduler busy (about 1000 warps per
    // GPU), up to 32 for big batches; the compact state arrays must fit ~160 KB.
    unsigned lanes = 1;
    while (lanes < 32 && n_chunks / lanes > 1024) lanes <<= 1;
    if (const char *e = getenv("FQ28_QUAL_LANES")) lanes = (unsigned)atoi(e) ? (unsigned)atoi(e) : lanes;
    if (lanes > 32) lanes = 32;
    while (lanes > 1 && (size_t)nt * lanes * sizeof(uint16_t) > 160 * 1024) lanes >>= 1;
    const size_t smem = (size_t)QUAL_N * sizeof(uint16_t) + RING_WORDS * sizeof(uint32_t) + (size_t)nt * lanes * sizeof(uint16_t) + 16;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
      FQ28_CUDA(h, cudaFuncSetAttribute(k_decode_qual, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    side_stage_begin(h, ST_DECODE_QUAL);
    k_decode_qual<<<(unsigned)((n_chunks + lanes - 1) / lanes), 32, smem, h->side>>>(
        ch, (unsigned)n_chunks, lanes, in->qual, h->qual.logs, h->qual.logsuf, h->qual.dtab_fix, h->qual.cid, nt,
        h->dec_cold.as<uint16_t>(), recscan, in->readlens, in->hdr_lens, d_out, h->d_status);
    FQ28_LAUNCH_CHECK(h);
    side_stage_end(h, ST_DECODE_QUAL);
  }
  FQ28_TRY(side_join(h));

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
