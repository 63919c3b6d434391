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
//  * one THREAD per stream.  A stream is one in-order dependent instruction
//    chain, so the kernels are latency-bound: few streams per warp for small
//    batches (a slow lane stalls its warp), up to 32 per warp for big ones;
//  * the per-symbol table lookups must not go to L2:
//      - sequence: all 256 DTables live in shared memory in compressed form
//        (SeqDecTables, 840 B per context) -- a cell is rebuilt from its 2-bit
//        symbol plus a two-level rank directory, in 32-bit integer ops; the
//        loop is software-pipelined (next symbol's loads before the current
//        symbol's rank / bit-read chain);
//      - quality: decoder states sit in shared memory under compact ids of the
//        contexts that have a real table (a few hundred of 8192), the maps are
//        shared by the streams of a CTA, which leaves most of the SM's 256 KB
//        as L1 for the hot DTable cells; runs of a dominant symbol in its
//        self-loop context are decoded up to 15 symbols per lookup (QZ_MAX);
//  * the bit reader keeps the next <= 64 stream bits left-aligned in two 32-bit
//    registers: taking nbBits is ONE funnel shift on the critical path, and the
//    next stream word is always prefetched one refill ahead;
//  * the two stream types run concurrently on two CUDA streams, on disjoint
//    sets of SMs (a sequence CTA takes a whole SM's shared memory).
#include <stdlib.h>

#include "fq28_internal.cuh"
#include "fq28_dec2.cuh"

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
// sequence decoder: one CTA per SM (shared memory: compressed tables 210 KB +
// states u16[256][32] 16 KB + ring); `warps` warps x `lanes` streams per warp,
// at most 32 streams per CTA.
//
// The loop is software-pipelined.  Per symbol there are two dependent chains:
//   A  symbol -> next context -> its state (shared) -> its table words (shared)
//   B  table words -> rank -> nbBits -> bit read -> new state -> store
// Only A is a recurrence between symbols (B feeds back only when the next
// context equals the current one, i.e. inside homopolymers), so the loads of
// chain A for symbol i+1 are issued before chain B of symbol i and both run
// in flight together.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
k_decode_seq(const DecChunk *__restrict__ ch, unsigned n_chunks, unsigned lanes, const uint8_t *__restrict__ arena,
             const SeqDecTables *__restrict__ gtab, const uint32_t *__restrict__ logsuf,
             const uint32_t *__restrict__ recscan, const uint16_t *__restrict__ readlens,
             const uint16_t *__restrict__ hdr_lens, char *__restrict__ out, DevStatus *st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned char *tb = smem_raw;
  uint16_t *S = reinterpret_cast<uint16_t *>(smem_raw + SEQ_TAB_BYTES);  // S[ctx * 32 + slot]
  uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + SEQ_TAB_BYTES + (size_t)SEQ_N * 32 * sizeof(uint16_t));
  {  // cooperative copy of the compressed tables (uint4 granularity)
    const uint4 *src = reinterpret_cast<const uint4 *>(gtab);
    uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
    for (unsigned i = threadIdx.x; i < sizeof(SeqDecTables) / 16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned per_cta = lanes * (blockDim.x >> 5);
  const unsigned slot = warp * lanes + lane;  // stream slot inside the CTA
  const unsigned k = blockIdx.x * per_cta + slot;
  const bool live = lane < lanes && k < n_chunks;
  DecChunk c;
  c.n_rec = 0; c.rec0 = 0; c.out_off = 0; c.seq_off = 0; c.seq_len = 0;
  if (live) c = ch[k];
  LaneBitReader br;
  bool ok = true;
  uint16_t *Sl = S + (live ? slot : 0);
  if (live) {
    ok = br.init(arena + c.seq_off, c.seq_len, ring + slot * 2);
    // FSE_Decoder::startChunk (src/fse_common.hpp:134-138): states for ctx N-1 .. 0
    if (!ok || br.top_bit - br.floor_ < (long long)logsuf[SEQ_N]) {
      ok = false;
      c.n_rec = 0;
    } else {
      for (unsigned cc = SEQ_N; cc > 0; --cc) {
        const unsigned lg = *reinterpret_cast<const uint16_t *>(tb + offsetof(SeqDecTables, snext) + (cc - 1) * 8) >> 12;
        Sl[(cc - 1) * 32] = (uint16_t)br.read(lg);
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
  // chain A results of the symbol about to be decoded
  unsigned s0, fr;
  uint2 nx, wv, cr;
  auto load_nx = [&](unsigned cx) {
    return *reinterpret_cast<const uint2 *>(tb + offsetof(SeqDecTables, snext) + cx * 8);
  };
  auto load_cell = [&](unsigned cx, unsigned sx, uint2 &w_, unsigned &f_, uint2 &c_) {
    const unsigned blk = sx >> 5;  // 32-cell block of the state
    w_ = *reinterpret_cast<const uint2 *>(tb + offsetof(SeqDecTables, symtab) + cx * 512 + blk * 8);
    f_ = *reinterpret_cast<const unsigned *>(tb + offsetof(SeqDecTables, fine) + cx * 256 + blk * 4);
    c_ = *reinterpret_cast<const uint2 *>(tb + offsetof(SeqDecTables, coarse) + cx * 64 + (sx >> 8) * 8);
  };
  uint16_t *scur = Sl + ctx * 32;  // state slot of the current context
  s0 = *scur;
  nx = load_nx(ctx);
  load_cell(ctx, s0, wv, fr, cr);
  // one symbol; `last` = this may be the last symbol of the lane's record
  auto step = [&](unsigned t, bool last) {
    const unsigned p = s0 & 31;
    const bool upper = (p & 16) != 0;
    const unsigned wsel = upper ? wv.y : wv.x;
    const unsigned sym = (wsel >> ((p & 15) * 2)) & 3u;
    // the first symbol of the next record starts from the initial context
    unsigned ctx1 = (ctx >> 2) + (sym << 6);  // addSymUpper
    if (last && i + t + 1 == cur.L) ctx1 = SEQ_INITIAL_CTX;
    // chain B up to the new-state base: (symbol, rank) -> cell (Appendix A.6)
    auto cell_base = [&](unsigned &nb) -> unsigned {
      const unsigned lg = (nx.x >> 12) & 15u;
      // masks of the cells below p in the two words of the block: one shift serves both
      const unsigned msk = (1u << ((2 * p) & 31)) - 1u;
      const unsigned m0 = upper ? 0xFFFFFFFFu : msk;
      const unsigned m1 = upper ? msk : 0u;
      const unsigned pat = sym * 0x55555555u;
      const unsigned x0 = wv.x ^ pat, x1 = wv.y ^ pat;
      const unsigned e0 = ~(x0 | (x0 >> 1)) & 0x55555555u & m0;
      const unsigned e1 = ~(x1 | (x1 >> 1)) & 0x55555555u & m1;
      const unsigned rank = __popc(e0) + __popc(e1) + ((fr >> (8 * sym)) & 0xFFu) + pick_u16x4(cr, sym);
      const unsigned xs = (pick_u16x4(nx, sym) & 0xFFFu) + rank;  // symbolNext + rank
      nb = lg - (31u - (unsigned)__clz(xs));
      return (xs << nb) - (1u << lg);
    };
    unsigned s1, fr1;
    uint2 nx1, wv1, cr1;
    uint16_t *snext = Sl + ctx1 * 32;
    if (ctx1 != ctx) {
      s1 = *snext;
      nx1 = load_nx(ctx1);
      load_cell(ctx1, s1, wv1, fr1, cr1);
      unsigned nb;
      const unsigned ns = cell_base(nb);
      *scur = (uint16_t)(ns + br.read(nb));
    } else {  // homopolymer: the next symbol needs the state written by this one
      unsigned nb;
      const unsigned ns = cell_base(nb);
      s1 = ns + br.read(nb);
      *scur = (uint16_t)s1;
      nx1 = nx;
      load_cell(ctx1, s1, wv1, fr1, cr1);
    }
    dst[i + t] = (char)((0x54474341u >> (8 * sym)) & 0xFFu);  // "ACGT"
    ctx = ctx1; scur = snext; s0 = s1; nx = nx1; wv = wv1; fr = fr1; cr = cr1;
  };
  for (;;) {
    // uniform trip count: symbols until the first lane reaches a record end
    const unsigned n = __reduce_min_sync(0xffffffffu, rr > 0 ? cur.L - i : 0xFFFFFFFFu);
    if (n == 0xFFFFFFFFu) break;
    if (rr > 0) {
      for (unsigned t = 0; t + 1 < n; t++) step(t, false);
      step(n - 1, true);
      i += n;
      if (i >= cur.L) {  // record done (ctx is already the initial context)
        --rr;
        cur = nxt;
        i = 0;
        dst = out + c.out_off + (cur.scan - scan0) + cur.hl + 1;
        if (rr > 1) nxt = load_meta(readlens, hdr_lens, recscan, c.rec0 + rr - 2);
      }
    }
  }
  if (live && (!ok || !br.finished())) set_error(st, FQ28_ERR_STREAM, k);  // BIT_endOfDStream, src/fse_common.hpp:141
}

// ---------------------------------------------------------------------------
// quality decoder: CTA = `warps` warps x `lanes` streams per warp (<= 32 streams)
// smem: cid map u16[8192] + ring + zero-bit run tables + compact states
// u16[n_touched][streams]; the maps are shared by the CTA's streams, which
// leaves most of the SM's L1 to the DTable cells.
// Contexts without a compact id (prior-only tables: never seen in the sample)
// keep their state in a global fallback array.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_decode_qual(const DecChunk *__restrict__ ch, unsigned n_chunks, unsigned lanes, const uint8_t *__restrict__ arena,
              const uint32_t *__restrict__ logs, const uint32_t *__restrict__ logsuf,
              const uint32_t *__restrict__ dtab_fix, const uint16_t *__restrict__ gcid, unsigned n_touched,
              const uint16_t *__restrict__ gzrun, unsigned n_z, uint4 zctx, uint16_t *cold_states /*[n_chunks][8192]*/, const uint32_t *__restrict__ recscan,
              const uint16_t *__restrict__ readlens, const uint16_t *__restrict__ hdr_lens, char *__restrict__ out,
              DevStatus *st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t *cid = reinterpret_cast<uint16_t *>(smem_raw);
  uint32_t *ring = reinterpret_cast<uint32_t *>(cid + QUAL_N);
  uint16_t *zt = reinterpret_cast<uint16_t *>(ring + RING_WORDS);  // [n_z][2][2048] zero-bit run tables
  uint16_t *S = zt + (size_t)n_z * 2 * (1u << FIX_LOG);             // S[id * per_cta + slot]
  for (unsigned i = threadIdx.x; i < QUAL_N; i += blockDim.x) cid[i] = gcid[i];
  for (unsigned i = threadIdx.x; i < n_z * (1u << FIX_LOG); i += blockDim.x)  // u32 copies of the u16 tables
    reinterpret_cast<uint32_t *>(zt)[i] = reinterpret_cast<const uint32_t *>(gzrun)[i];
  __syncthreads();
  const unsigned lane = threadIdx.x & 31;
  const unsigned per_cta = lanes * (blockDim.x >> 5);
  const unsigned slot = (threadIdx.x >> 5) * lanes + lane;  // stream slot inside the CTA
  const unsigned k = blockIdx.x * per_cta + slot;
  const bool live = lane < lanes && k < n_chunks;
  const unsigned nt_pad = (n_touched + 1u) & ~1u;
  uint16_t *Sl = S + (size_t)slot * nt_pad;  // this stream's states
  DecChunk c;
  c.n_rec = 0; c.rec0 = 0; c.out_off = 0; c.qual_off = 0; c.qual_len = 0;
  if (live) c = ch[k];
  uint16_t *cold = cold_states + (size_t)(live ? k : 0) * QUAL_N;
  LaneBitReader br;
  bool ok = true;
  if (live) {
    ok = br.init(arena + c.qual_off, c.qual_len, ring + slot * 2);
    if (!ok || br.top_bit - br.floor_ < (long long)logsuf[QUAL_N]) {
      ok = false;
      c.n_rec = 0;
    } else {
      for (unsigned cc = QUAL_N; cc > 0; --cc) {
        const uint16_t v = (uint16_t)br.read(logs[cc - 1]);
        const unsigned id = cid[cc - 1];
        if (id != 0xFFFFu) Sl[id & 0x1FFFu] = v; else cold[cc - 1] = v;
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
  // shared-window addresses and the table base pinned in registers: the hot
  // loop is bound by its instruction count
  unsigned cid_s, zt_s, sl_s;
  const uint32_t *dt;
  {
    const unsigned a0 = (unsigned)__cvta_generic_to_shared(cid), a1 = (unsigned)__cvta_generic_to_shared(zt),
                   a2 = (unsigned)__cvta_generic_to_shared(Sl);
    asm volatile("mov.u32 %0, %1;" : "=r"(cid_s) : "r"(a0));
    asm volatile("mov.u32 %0, %1;" : "=r"(zt_s) : "r"(a1));
    asm volatile("mov.u32 %0, %1;" : "=r"(sl_s) : "r"(a2));
    asm volatile("mov.u64 %0, %1;" : "=l"(dt) : "l"(dtab_fix));
  }
  auto lds16 = [](unsigned a) -> unsigned {  // read-only tables
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
  };
  auto lds16_state = [](unsigned a) -> unsigned {  // ordered with the state stores
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
  };
  auto sts16_state = [](unsigned a, unsigned v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v));
  };
  for (;;) {
    const unsigned n = __reduce_min_sync(0xffffffffu, rr > 0 ? cur.L - i : 0xFFFFFFFFu);
    if (n == 0xFFFFFFFFu) break;
    if (rr > 0) {
      char *o = dst + i;
      for (unsigned t = 0; t < n;) {
        // compact id in bits 0..12, zero-bit run slot + 1 in bits 13..15 (k_qual_zrun)
        const unsigned cv = lds16(cid_s + ctx * 2);
        if (cv == 0xFFFFu) {  // context without a compact id: state in global memory (rare)
          const unsigned e = __ldg(dt + (ctx << FIX_LOG) + cold[ctx]);
          const unsigned q = (e >> 16) & 63u;
          cold[ctx] = (uint16_t)((e & 0xFFFFu) + br.read(e >> 24));
          *o++ = (char)(q + QUAL_OFFSET);
          const unsigned mx = q1 > q2 ? q1 : q2;
          ctx = (((mx << 6) + q) & 0xFFFu) + ((unsigned)(q1 == q2) << 12);
          q2 = q1;
          q1 = q;
          t++;
          continue;
        }
        const unsigned sa = sl_s + (cv & 0x1FFFu) * 2;
        const unsigned s0 = lds16_state(sa);
        // zero-bit run (QZ_MAX in fq28_internal.cuh): the last three symbols equal d
        // and d is dominant in ctx(d,d,d) -- up to 15 symbols from one table lookup
        if (n_z && (cv >> 13)) {
          const unsigned zb = zt_s + ((cv >> 13) - 1u) * (4u << FIX_LOG);
          const unsigned z = lds16(zb + s0 * 2);
          unsigned kz = z >> 11;
          if (kz) {
            unsigned x = z & 0x7FFu;
            if (kz > n - t) {  // record ends inside the run: single steps
              kz = n - t;
              x = s0;
              for (unsigned u = 0; u < kz; u++) x = lds16(zb + (2u << FIX_LOG) + x * 2);
            }
            sts16_state(sa, x);
            const char dc = (char)((ctx & 63u) + QUAL_OFFSET);
            for (unsigned u = 0; u < kz; u++) o[u] = dc;
            o += kz;
            t += kz;
            continue;  // q1 == q2 == d, context unchanged
          }
        }
        const unsigned e = __ldg(dt + (ctx << FIX_LOG) + s0);
        const unsigned q = (e >> 16) & 63u;
        sts16_state(sa, (e & 0xFFFFu) + br.read(e >> 24));
        *o++ = (char)(q + QUAL_OFFSET);
        // calcContext(q, q1, q2), src/fse_quality.h:40-44
        const unsigned mx = q1 > q2 ? q1 : q2;
        ctx = (((mx << 6) + q) & 0xFFFu) + ((unsigned)(q1 == q2) << 12);
        q2 = q1;
        q1 = q;
        t++;
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

// ---------------------------------------------------------------------------
// Decoder v2 (fq28_dec2.cuh): every context caches the cell of its current state
// in shared memory, the table fetch of the next cell is an async copy into that slot.
// One thread per stream, `lanes` streams per warp in lockstep (few for small
// batches: one warp per SM sub-partition is the sweet spot of a latency chain;
// 32 for big ones).  Shared memory per CTA:
//   sequence: 4 homopolymer tables (32 KB) | ring | S (1 KB per stream)
//   quality : rk | zc | run tables (16 KB per slot) | ring | S (|V| * 512 B per stream)
// ---------------------------------------------------------------------------
constexpr unsigned D2_HT_BYTES = 4u * (4u << FIX_LOG);
__host__ __device__ inline size_t d2_seq_smem(unsigned per_cta) {
  return D2_HT_BYTES + (size_t)per_cta * 16 + 1024 /*alignment slack*/ + (size_t)per_cta * 1024;
}
__host__ __device__ inline size_t d2_qual_fixed(unsigned nz) { return 128 + (size_t)nz * dec2::ZQ_SLOT_BYTES; }
__host__ __device__ inline size_t d2_qual_smem(unsigned per_cta, unsigned nz, unsigned nv) {
  return d2_qual_fixed(nz) + (size_t)per_cta * 16 + 256 /*alignment slack*/ + (size_t)per_cta * nv * 2 * dec2::QROW_BYTES;
}

// windowed layout: rk | zc | row descriptors (128 x 8 B) | ring | S (n_win * 4 B per stream, 16-byte rounded)
__host__ __device__ inline size_t d2_qualw_stream(unsigned n_win) { return ((size_t)n_win * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t d2_qualw_smem(unsigned per_cta, unsigned n_win) {
  return 128 + 1024 + (size_t)per_cta * 16 + 256 /*alignment slack*/ + (size_t)per_cta * d2_qualw_stream(n_win);
}

__device__ __forceinline__ void d2_args(dec2::StreamArgs &a, const DecChunk &c, bool live, const uint8_t *stream, uint32_t len,
                                        const uint32_t *recscan, const uint16_t *readlens, const uint16_t *hdr_lens, char *out,
                                        const uint32_t *logs, const uint32_t *logsuf, const uint32_t *wtab, void *ring) {
  a.src = stream; a.len = len; a.rec0 = c.rec0; a.n_rec = c.n_rec;
  a.readlens = readlens; a.hdr_lens = hdr_lens; a.recscan = recscan;
  a.out = out + c.out_off; a.logs = logs; a.logsuf = logsuf; a.wtab = wtab; a.ring = ring; a.live = live;
}

__global__ void __launch_bounds__(256)
k_dec2_seq(const DecChunk *__restrict__ ch, unsigned n_chunks, unsigned lanes, const uint8_t *__restrict__ arena,
           const uint32_t *__restrict__ wtab, const uint32_t *__restrict__ logs, const uint32_t *__restrict__ logsuf,
           const uint32_t *__restrict__ recscan, const uint16_t *__restrict__ readlens,
           const uint16_t *__restrict__ hdr_lens, char *__restrict__ out, DevStatus *st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned per_cta = lanes * (blockDim.x >> 5);
  {  // homopolymer contexts AAAA, CCCC, GGGG, TTTT = 0x00, 0x55, 0xAA, 0xFF
    uint4 *dst = reinterpret_cast<uint4 *>(smem_raw);
    for (unsigned i = threadIdx.x; i < D2_HT_BYTES / 16; i += blockDim.x) {
      const unsigned j = i >> (FIX_LOG - 2), u = i & ((1u << (FIX_LOG - 2)) - 1u);
      dst[i] = __ldg(reinterpret_cast<const uint4 *>(wtab + ((size_t)(j * 0x55u) << FIX_LOG)) + u);
    }
  }
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t ring0 = base + D2_HT_BYTES;
  const uint32_t s0 = (ring0 + per_cta * 16 + 1023u) & ~1023u;
  const unsigned lane = threadIdx.x & 31;
  const unsigned slot = (threadIdx.x >> 5) * lanes + lane;
  const unsigned k = blockIdx.x * per_cta + slot;
  const bool live = lane < lanes && k < n_chunks;
  DecChunk c;
  c.n_rec = 0; c.rec0 = 0; c.out_off = 0; c.seq_off = 0; c.seq_len = 0;
  if (live) c = ch[k];
  const unsigned sl = live ? slot : 0;
  dec2::StreamArgs a;
  d2_args(a, c, live, arena + c.seq_off, c.seq_len, recscan, readlens, hdr_lens, out, logs, logsuf, wtab,
          smem_raw + (ring0 - base) + sl * 16);
  const bool ok = dec2::decode_seq_stream(a, s0 + sl * 1024, base);
  if (live && !ok) set_error(st, FQ28_ERR_STREAM, k);  // BIT_endOfDStream, src/fse_common.hpp:141
}

// WIN: windowed layout of the cached cells (fq28_dec2.cuh); then nv = entries per stream, gzrun = the
// row descriptors (uint2[128]), nz = number of rows, and there are no run tables.
template <bool WIN>
__global__ void __launch_bounds__(256)
k_dec2_qual(const DecChunk *__restrict__ ch, unsigned n_chunks, unsigned lanes, const uint8_t *__restrict__ arena,
            const uint32_t *__restrict__ wtab, const uint32_t *__restrict__ logs, const uint32_t *__restrict__ logsuf,
            const uint32_t *__restrict__ dtab_fix, const uint16_t *__restrict__ cid, const uint8_t *__restrict__ qrk,
            unsigned nv, const uint16_t *__restrict__ gzrun, unsigned nz, uint4 zctx, uint16_t *cold_states,
            const uint32_t *__restrict__ recscan, const uint16_t *__restrict__ readlens,
            const uint16_t *__restrict__ hdr_lens, char *__restrict__ out, DevStatus *st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned per_cta = lanes * (blockDim.x >> 5);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_raw);
  dec2::QualShared qs;
  qs.rk_a = base; qs.zc_a = base + 64; qs.zq_a = base + 128; qs.n_slots = WIN ? 0u : nz;
  qs.row_a = base + 128; qs.n_rows = WIN ? nz : 0u;
  if (WIN) {
    for (unsigned i = threadIdx.x; i < 64; i += blockDim.x) smem_raw[i] = qrk[i];
    const uint2 *rows = reinterpret_cast<const uint2 *>(gzrun);
    for (unsigned i = threadIdx.x; i < 128; i += blockDim.x) reinterpret_cast<uint2 *>(smem_raw + 128)[i] = rows[i];
  } else {
    for (unsigned i = threadIdx.x; i < 64; i += blockDim.x) smem_raw[i] = qrk[i];  // (a CTA may be one warp)
    const unsigned zc[4] = {zctx.x, zctx.y, zctx.z, zctx.w};
    for (unsigned j = 0; j < nz; j++) {
      const unsigned d = zc[j] & 63u;        // run context = ctx(d, d, d)
      const unsigned r = qrk[d];
      if (threadIdx.x == 0) reinterpret_cast<uint32_t *>(smem_raw + 64)[j] = d + QUAL_OFFSET;
      // run table entry of state x: the W cell of x | ZENT entry after the zero-bit run + its length
      const uint16_t *zsrc = gzrun + (size_t)j * 2 * (1u << FIX_LOG);   // (k << 11) | state after k zero-bit steps
      const uint32_t *hsrc = wtab + ((size_t)dec2::qual_dense_id(r, 1, r) << FIX_LOG);
      uint2 *zdst = reinterpret_cast<uint2 *>(smem_raw + 128 + j * dec2::ZQ_SLOT_BYTES);
      for (unsigned x = threadIdx.x; x < (1u << FIX_LOG); x += blockDim.x) {
        const unsigned z = zsrc[x];
        zdst[x] = make_uint2(hsrc[x], dec2::make_zq_hi(z >> 11, z & 0x7FFu, j));
      }
    }
  }
  __syncthreads();
  const uint32_t ring0 = WIN ? base + 128 + 1024 : qs.zq_a + nz * dec2::ZQ_SLOT_BYTES;
  const uint32_t s0 = (ring0 + per_cta * 16 + 255u) & ~255u;
  const unsigned s_bytes = WIN ? (unsigned)d2_qualw_stream(nv) : nv * 2 * dec2::QROW_BYTES;
  const unsigned lane = threadIdx.x & 31;
  const unsigned slot = (threadIdx.x >> 5) * lanes + lane;
  const unsigned k = blockIdx.x * per_cta + slot;
  const bool live = lane < lanes && k < n_chunks;
  DecChunk c;
  c.n_rec = 0; c.rec0 = 0; c.out_off = 0; c.qual_off = 0; c.qual_len = 0;
  if (live) c = ch[k];
  const unsigned sl = live ? slot : 0;
  dec2::StreamArgs a;
  d2_args(a, c, live, arena + c.qual_off, c.qual_len, recscan, readlens, hdr_lens, out, logs, logsuf, wtab,
          smem_raw + (ring0 - base) + sl * 16);
  const bool ok = dec2::decode_qual_stream<WIN>(a, qs, s0 + sl * s_bytes, dtab_fix, cid,
                                          cold_states + (size_t)(live ? k : 0) * QUAL_N);
  if (live && !ok) set_error(st, FQ28_ERR_STREAM, k);
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

// Side-information checks that must hold BEFORE any kernel writes through offsets derived from
// it (a corrupt archive must produce an error code, not an out-of-bounds write): per chunk, the
// records' laid-out size (u64, no wrap) fits `total`, every read length is >= 1, the chunk's
// header bytes stay inside `headers`.  One CTA per chunk.
__global__ void __launch_bounds__(256)
k_validate_chunks(const DecChunk *__restrict__ ch, unsigned n_chunks, const uint16_t *__restrict__ readlens,
                  const uint16_t *__restrict__ hdr_lens, const uint16_t *__restrict__ n_count, size_t headers_bytes,
                  const uint32_t *__restrict__ hdrscan, size_t n_rec_total, DevStatus *st) {
  __shared__ unsigned long long s_bytes[8], s_np[8];
  __shared__ unsigned s_bad;
  const unsigned k = blockIdx.x;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const DecChunk c = ch[k];
  unsigned long long bytes = 0, np = 0;
  unsigned bad = 0;
  for (unsigned r = threadIdx.x; r < c.n_rec; r += blockDim.x) {
    const unsigned L = readlens[c.rec0 + r];
    bytes += (unsigned long long)hdr_lens[c.rec0 + r] + 2ull * L + 5ull;
    np += n_count[c.rec0 + r];
    bad |= L == 0;
  }
  for (int d = 16; d >= 1; d >>= 1) {
    bytes += __shfl_xor_sync(0xffffffffu, bytes, d);
    np += __shfl_xor_sync(0xffffffffu, np, d);
  }
  if ((threadIdx.x & 31) == 0) { s_bytes[threadIdx.x >> 5] = bytes; s_np[threadIdx.x >> 5] = np; }
  if (bad) atomicOr(&s_bad, 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tb = 0, tn = 0;
    for (int w = 0; w < 8; w++) { tb += s_bytes[w]; tn += s_np[w]; }
    if (s_bad) set_error(st, FQ28_ERR_SHORT, k);
    else if (tb > c.total) set_error(st, FQ28_ERR_FORMAT, k);
    else if (tn > c.npos_len) set_error(st, FQ28_ERR_STREAM, k);
    if (k == n_chunks - 1 && (size_t)hdrscan[n_rec_total] > headers_bytes) set_error(st, FQ28_ERR_FORMAT, k);
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
  if (infos[n_chunks - 1].rec_off + infos[n_chunks - 1].n_records != n_rec)
    return fail(h, FQ28_ERR_ARG, "chunks cover %llu records, arenas hold %zu",
                (unsigned long long)(infos[n_chunks - 1].rec_off + infos[n_chunks - 1].n_records), n_rec);
  if (n_rec >= 0xFFFFFFF0u) return fail(h, FQ28_ERR_ARG, "too many records in one batch");
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
  // no kernel below may write before the side information has been validated
  k_validate_chunks<<<(unsigned)n_chunks, 256, 0, h->stream>>>(ch, (unsigned)n_chunks, in->readlens, in->hdr_lens, in->n_count,
                                                              in->headers_bytes, hdrscan, n_rec, h->d_status);
  FQ28_LAUNCH_CHECK(h);
  FQ28_TRY(check_status(h, "decodeChunk side information"));
  if (n_rec) {
    k_layout<<<(unsigned)((n_rec + LAY_WARPS - 1) / LAY_WARPS), LAY_WARPS * 32, 0, h->stream>>>(
        ch, (unsigned)n_chunks, recscan, hdrscan, in->readlens, in->hdr_lens, in->headers, n_rec, d_out);
    FQ28_LAUNCH_CHECK(h);
  }
  k_chunk_tail<<<(unsigned)n_chunks, 128, 0, h->stream>>>(ch, (unsigned)n_chunks, recscan, d_out, h->d_status);
  FQ28_LAUNCH_CHECK(h);
  stage_end(h, ST_LAYOUT);

  // the two stream types are independent: quality runs on the side stream (FQ28_DEC_SERIAL puts
  // both on the main stream, one after the other, to time each kernel on its own)
  // Many-valued qualities (HiSeq-like, ONT: dozens of values, thousands of live contexts) make the
  // quality decoder several times slower than the sequence decoder and sensitive to sharing its
  // SMs' issue slots and L1 (measured, 41-level qualities, 1 GB: 152 ms next to the sequence
  // kernel on 83 SMs, 55 ms alone on the whole GPU): then the two kernels run one after the other.
  const bool serial = h->cfg.dec_serial || (h->cfg.qual_v2 && !h->cfg.dec_concurrent && h->qual.h_n_v >= 16);
  cudaStream_t qstream = serial ? h->stream : h->side;
  if (!serial) FQ28_TRY(side_fork(h));
  // Lanes (streams per warp) of the v2 kernels: a stream is one latency chain; lockstep lanes
  // save issue slots but every lane waits for the slowest path taken in its warp
  unsigned lanes = (unsigned)((n_chunks + 1183) / 1184);
  lanes = lanes < 1 ? 1 : lanes > 32 ? 32 : lanes;
  auto shape = [&](unsigned want_lanes, unsigned want_warps, unsigned &l, unsigned &w) {
    l = want_lanes ? (want_lanes > 32 ? 32 : want_lanes) : lanes;
    w = want_warps ? (want_warps > 8 ? 8 : want_warps) : 4;
  };
  stage_begin(h, ST_DECODE_SEQ);
  if (!h->cfg.seq_v1) {
    unsigned l, w;
    size_t smem;
    if (serial || h->cfg.share_sms) {
      shape(h->cfg.seq_lanes, h->cfg.seq_warps, l, w);
      while (d2_seq_smem(l * w) > 200 * 1024 && w > 1) w >>= 1;
      smem = d2_seq_smem(l * w);
    } else {
      // The two kernels must not share SMs: measured, either quality kernel takes 2-3x longer when
      // sequence CTAs live on its SMs (76 ms instead of 37 for the state-table one, 95 instead of
      // 30 for the cached-cell one: the table cells they fetch are L1 hits only as long as nothing
      // else streams a 2 MB table through that L1).  So the sequence CTAs are made fat -- 8 warps
      // and a shared-memory request no other CTA fits beside -- and sized to cover about 65 of
      // the 148 SMs; quality gets the rest.
      // (16 warps x 1 stream per CTA -- no lockstep lanes, four warps per scheduler -- was measured
      // slower: 42.0 ms against 36.0 ms for 8 warps x 2 streams)
      w = h->cfg.seq_warps ? (h->cfg.seq_warps > 8 ? 8 : h->cfg.seq_warps) : 8;
      l = (unsigned)((n_chunks + 65 * w - 1) / (65 * w));
      if (h->cfg.seq_lanes) l = h->cfg.seq_lanes;
      l = l < 1 ? 1 : l > 32 ? 32 : l;
      while (d2_seq_smem(l * w) > 220 * 1024 && l > 1) --l;
      smem = d2_seq_smem(l * w);
      if (smem < 200 * 1024) smem = 200 * 1024;
    }
    const unsigned per_cta = l * w;
    k_dec2_seq<<<(unsigned)((n_chunks + per_cta - 1) / per_cta), w * 32, smem, h->stream>>>(
        ch, (unsigned)n_chunks, l, in->seq, h->seq.wtab, h->seq.logs, h->seq.logsuf, recscan, in->readlens, in->hdr_lens,
        d_out, h->d_status);
    FQ28_LAUNCH_CHECK(h);
  } else {
  {
    // streams per CTA (= per SM): few per warp keeps the homopolymer path from
    // stalling the other streams of a warp; big batches fill 32 slots per SM
    // One warp per SM sub-partition: a stream is a chain of dependent instructions that wants an
    // issue slot every ~4.5 cycles, two warps on one scheduler already slow each other down
    // (measured: 4 warps x 3 streams 62 ms, 8 x 2 71 ms, 16 x 1 104 ms).  Streams per warp are
    // chosen so that the sequence CTAs take about 86 SMs and leave the rest to the quality decoder.
    unsigned s_warps = 4;
    unsigned s_lanes = (unsigned)((n_chunks + 4 * 86 - 1) / (4 * 86));
    if (s_lanes < 1) s_lanes = 1;
    if (s_lanes > 8) s_lanes = 8;
    if (h->cfg.seq_lanes) s_lanes = h->cfg.seq_lanes;
    if (h->cfg.seq_warps) s_warps = h->cfg.seq_warps;
    if (s_lanes > 32) s_lanes = 32;
    while (s_lanes * s_warps > 32) s_warps >>= 1;
    const unsigned per_cta = s_lanes * s_warps;
    k_decode_seq<<<(unsigned)((n_chunks + per_cta - 1) / per_cta), s_warps * 32, SEQ_DEC_SMEM, h->stream>>>(
        ch, (unsigned)n_chunks, s_lanes, in->seq, reinterpret_cast<const SeqDecTables *>(h->seq.seqdec), h->seq.logsuf,
        recscan, in->readlens, in->hdr_lens, d_out, h->d_status);
    FQ28_LAUNCH_CHECK(h);
  }
  }
  stage_end(h, ST_DECODE_SEQ);
  if (h->cfg.qual_v2) {
    {
      unsigned l, w;
      shape(h->cfg.qual_lanes, h->cfg.qual_warps, l, w);
      if (!h->cfg.qual_lanes) {  // about 83 SMs = 332 sub-partitions are the quality decoder's (all 592 when it runs alone): up to ~4 warps on each
        const unsigned sub = serial ? 592u : 332u;
        l = (unsigned)((n_chunks + 4 * sub - 1) / (4 * sub));
        l = l < 1 ? 1 : l > 32 ? 32 : l;
      }
      unsigned nz = h->cfg.no_zrun ? 0u : h->qual.h_n_z;
      const unsigned nv = h->qual.h_n_v;
      // Many-valued qualities: the cached cells take |V| * 512 B per stream in the dense layout (20 KB
      // for 40 values), so shared memory, not warp slots, decides how many streams an SM holds.  The
      // windowed layout keeps only the columns every row's touched contexts span (a few KB).
      // It costs two ALU operations in the symbol chain and the run tables (measured at 1 028 streams:
      // 70.9 against 63.1 ms HiSeq-like, 73.9 against 63.8 ms ONT-like), so it is used when the dense
      // layout cannot hold all streams at once (3 857 / 11 571 streams: 19.2 / 26.5 GB/s against
      // 13.3 / 14.0 HiSeq-like, 17.0 / 21.3 against 12.0 / 15.5 ONT-like).
      const unsigned n_win = h->qual.h_n_win;
      bool win = h->cfg.force_win && n_win > 0;
      if (serial && !h->cfg.no_win && n_win > 0 && !h->cfg.qual_lanes && !h->cfg.qual_warps) {
        const size_t per_stream = d2_qual_smem(2, 0, nv) - d2_qual_smem(1, 0, nv);
        const size_t dense_cap = std::min<size_t>(8, (224 * 1024 - d2_qual_smem(0, 0, nv)) / per_stream);
        win = win || n_chunks > 148 * std::max<size_t>(dense_cap, 1);
      }
      auto smem_of = [&](unsigned per_cta, unsigned z) {
        return win ? d2_qualw_smem(per_cta, n_win) : d2_qual_smem(per_cta, z, nv);
      };
      if (serial && !h->cfg.qual_lanes && !h->cfg.qual_warps) {
        // One stream per warp (lanes in lockstep cost 1.4x per doubling here), CTAs sized so that the
        // streams an SM can hold are resident at once: if that covers all streams the launch is ONE
        // wave (measured, 1 028 streams of 41-valued qualities in the dense layout: 257 CTAs of
        // 150 KB made two waves, 116 ms instead of 58).
        const size_t budget = 224 * 1024;
        auto cap_of = [&](unsigned z) {   // streams per SM that shared memory allows (one or more CTAs)
          const size_t per_stream = smem_of(2, z) - smem_of(1, z);
          const size_t fixed = smem_of(0, z);
          if (win) {   // small CTAs pack: count whole CTAs of 4 streams
            const size_t cta4 = smem_of(4, z) + 1024;
            const size_t n = budget / cta4;
            return (unsigned)(n ? n * 4 : 1);
          }
          const unsigned c = fixed + per_stream > budget ? 1u : (unsigned)((budget - fixed) / per_stream);
          return c > 8 ? 8u : c;
        };
        auto waves_of = [&](unsigned z) { return (n_chunks + 148ull * cap_of(z) - 1) / (148ull * cap_of(z)); };
        // the zero-bit run tables (16 KB per slot) are worth less than a wave: without them the run
        // contexts are ordinary contexts (same bytes)
        if (!win && nz && waves_of(0) < waves_of(nz)) nz = 0;
        const unsigned cap = cap_of(nz);
        const unsigned need = (unsigned)((n_chunks + 147) / 148);   // streams per SM for one wave
        w = need < cap ? need : cap;
        if (w > 4 && win) w = 4;   // (several 4-warp CTAs per SM rather than one fat one)
        if (w > 8) w = 8;
        l = 1;
      }
      // the per-stream context arrays must fit: fewer warps first, then fewer lanes
      while (smem_of(l * w, nz) > 224 * 1024 && l * w > 1) {
        if (w > 1) w >>= 1; else l = (l + 1) / 2;
      }
      const unsigned per_cta = l * w;
      uint4 zctx = make_uint4(h->qual.h_zctx[0], h->qual.h_zctx[1], h->qual.h_zctx[2], h->qual.h_zctx[3]);
      const int qslot = stage_open(h, ST_DECODE_QUAL, qstream);
      if (win) {
        {
          // The refreshed cells come through L1: reserve only the shared memory the resident CTAs
          // need (the default carve-out sizes it for as many CTAs as could ever fit and leaves no L1).
          const size_t n_ctas = (n_chunks + per_cta - 1) / per_cta;
          const size_t cta_bytes = smem_of(per_cta, 0) + 1024;
          size_t per_sm = (n_ctas + 147) / 148;
          const size_t fit = (224 * 1024) / cta_bytes;
          if (per_sm > fit) per_sm = fit;
          int pct = (int)((per_sm * cta_bytes * 100 + 228 * 1024 - 1) / (228 * 1024));
          pct = pct < 1 ? 1 : pct > 100 ? 100 : pct;
          if (h->cfg.qual_carveout != -2) pct = h->cfg.qual_carveout < 0 ? cudaSharedmemCarveoutDefault : h->cfg.qual_carveout;
          if (pct != h->qualw_carve_set) {
            FQ28_CUDA(h, cudaFuncSetAttribute(k_dec2_qual<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            h->qualw_carve_set = pct;
          }
        }
        k_dec2_qual<true><<<(unsigned)((n_chunks + per_cta - 1) / per_cta), w * 32, smem_of(per_cta, 0), qstream>>>(
            ch, (unsigned)n_chunks, l, in->qual, h->qual.wtabw, h->qual.logs, h->qual.logsuf, h->qual.dtab_fix, h->qual.cid,
            h->qual.qrk, n_win, reinterpret_cast<const uint16_t *>(h->qual.qwin), h->qual.h_n_rows, zctx,
            h->dec_cold.as<uint16_t>(), recscan, in->readlens, in->hdr_lens, d_out, h->d_status);
      } else {
        k_dec2_qual<false><<<(unsigned)((n_chunks + per_cta - 1) / per_cta), w * 32, smem_of(per_cta, nz), qstream>>>(
            ch, (unsigned)n_chunks, l, in->qual, h->qual.wtab, h->qual.logs, h->qual.logsuf, h->qual.dtab_fix, h->qual.cid,
            h->qual.qrk, nv, h->qual.zrun, nz, zctx, h->dec_cold.as<uint16_t>(), recscan, in->readlens, in->hdr_lens, d_out,
            h->d_status);
      }
      FQ28_LAUNCH_CHECK(h);
      stage_close(h, qslot, qstream);
    }
  } else {
  {
    const unsigned nt = h->qual.h_n_touched;
    // Streams per warp.  The DTable cells come through L1/L2, so the streams of
    // a warp wait for the slowest lane: pack only as many streams per warp as
    // are needed to keep ~2 warps per scheduler busy (about 1000 warps per
    // GPU), up to 32 for big batches; the compact state arrays must fit ~160 KB.
    unsigned lanes = 1, q_warps = 8;
    while (lanes < 4 && n_chunks / (lanes * q_warps) > 148) lanes <<= 1;
    if (h->cfg.qual_lanes) lanes = h->cfg.qual_lanes;
    if (h->cfg.qual_warps) q_warps = h->cfg.qual_warps;
    if (lanes > 32) lanes = 32;
    while (lanes * q_warps > 32) q_warps >>= 1;
    while (lanes * q_warps > 1 && (size_t)nt * lanes * q_warps * sizeof(uint16_t) > 160 * 1024) {
      if (q_warps > 1) q_warps >>= 1; else lanes >>= 1;
    }
    const unsigned q_per_cta = lanes * q_warps;
    const unsigned nz = h->cfg.no_zrun ? 0u : h->qual.h_n_z;
    const size_t zbytes = (size_t)nz * 2 * (1u << FIX_LOG) * sizeof(uint16_t);
    const size_t smem = (size_t)QUAL_N * sizeof(uint16_t) + RING_WORDS * sizeof(uint32_t) + zbytes + (size_t)((nt + 1u) & ~1u) * q_per_cta * sizeof(uint16_t) + 16;
    uint4 zctx = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    if (nz > 0) zctx.x = h->qual.h_zctx[0];
    if (nz > 1) zctx.y = h->qual.h_zctx[1];
    if (nz > 2) zctx.z = h->qual.h_zctx[2];
    if (nz > 3) zctx.w = h->qual.h_zctx[3];
    {
      // Few CTAs land on an SM (the batch has ~n_chunks / q_per_cta of them), but the default
      // carve-out reserves shared memory for a full SM of them and shrinks the L1 that holds
      // the DTable cells: ask for just what the resident CTAs need.
      const unsigned n_ctas = (unsigned)((n_chunks + q_per_cta - 1) / q_per_cta);
      const unsigned per_sm = (n_ctas + 83) / 84 + 1;  // quality shares the GPU with the sequence decoder
      int pct = (int)((per_sm * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
      if (h->cfg.qual_carveout != -2) pct = h->cfg.qual_carveout;
      if (pct > 100) pct = 100;
      if (pct < 0) pct = cudaSharedmemCarveoutDefault;
      if (pct != h->qual_carve_set) {
        FQ28_CUDA(h, cudaFuncSetAttribute(k_decode_qual, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        h->qual_carve_set = pct;
      }
    }
    const int qslot = stage_open(h, ST_DECODE_QUAL, qstream);
    k_decode_qual<<<(unsigned)((n_chunks + q_per_cta - 1) / q_per_cta), q_warps * 32, smem, qstream>>>(
        ch, (unsigned)n_chunks, lanes, in->qual, h->qual.logs, h->qual.logsuf, h->qual.dtab_fix, h->qual.cid, nt,
        h->qual.zrun, nz, zctx, h->dec_cold.as<uint16_t>(), recscan, in->readlens, in->hdr_lens, d_out, h->d_status);
    FQ28_LAUNCH_CHECK(h);
    stage_close(h, qslot, qstream);
  }
    }
  if (!serial) FQ28_TRY(side_join(h));

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

// Dynamic shared memory opt-ins, per device (fq28_create).
int decode_init_device(fq28_handle *h) {
  const int big = 227 * 1024;
  FQ28_CUDA(h, cudaFuncSetAttribute(k_decode_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQ_DEC_SMEM));
  FQ28_CUDA(h, cudaFuncSetAttribute(k_decode_qual, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  FQ28_CUDA(h, cudaFuncSetAttribute(k_dec2_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  FQ28_CUDA(h, cudaFuncSetAttribute(k_dec2_qual<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  FQ28_CUDA(h, cudaFuncSetAttribute(k_dec2_qual<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  return FQ28_OK;
}

}  // namespace fq28
