// fq28_dec2.cuh -- per-stream tANS decoders (K6), "cached cell + deferred refresh".
//
// Replaces SequenceDecoder::decodeRecord (src/fse_sequence.cpp:114-143),
// QualityDecoder::decodeRecord (src/fse_quality.cpp:55-67) and
// FSE_Decoder::startChunk/endChunk (src/fse_common.hpp:130-141).
//
// A chunk stream is serial: the next context and the next bit offset both
// depend on the symbol just decoded.  The reference's step is
//     state[ctx] -> DTable[ctx][state] -> (symbol, nbBits, newState) -> ctx'
// i.e. two dependent loads per symbol.  Here every context keeps the *cell of
// its current state* (W) in shared memory instead of the state:
//     W = S[ctx]  ->  symbol  ->  ctx'                (ONE shared load + 1 ALU)
// and the remaining work (bit read, new state, fetching the cell of the new
// state from the L2-resident table) is off the recurrence: the fetch is a 4-byte
// cp.async from the table straight into S[ctx] -- no destination register, hence
// no scoreboard wait in the instruction stream (ptxas puts every in-flight LDG of
// an unrolled loop on ONE scoreboard, so a register-based software pipeline waits
// for the youngest load at every step; measured, 400 cycles per symbol).  A
// context whose refresh is in flight holds STALE in S; reading STALE waits for
// the thread's async copies and reads again.  Contexts that loop onto themselves (sequence homopolymers,
// quality ctx(d,d,d) with a dominant d) are refreshed inline from shared-memory
// copies of their tables; the latter also use the zero-bit run tables.
//
// The code is __host__ __device__: tests/cpp/dec2_host.cu compiles the very same
// functions for the CPU (shared memory = a byte array, one lane) so that the
// algorithm is checked against the oracle without a GPU.  The product only ever
// runs the device instantiation (fq28_decode.cu).
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FQ_HD __host__ __device__ __forceinline__
#else
#define FQ_HD inline
#endif

namespace fq28 {
namespace dec2 {

constexpr unsigned TAB_LOG = 11;            // rows of the W tables hold 1 << 11 cells (every log <= 11)
constexpr unsigned SEQ_INIT_CTX = 0xD7;     // src/fse_sequence.h:41-63
constexpr unsigned QUAL_CHAR0 = 33;         // src/fse_quality.h:24

// ---- sequence W (cell of the current state of one context) -------------------
//  [0,8) output char   [8,10) symbol   [10] INLINE (homopolymer context)
//  [11] STALE (S entries only)   [12,16) nbBits   [16,27) newState base
//  [27,29) inline table slot (= the context's base)
constexpr uint32_t SW_INLINE = 1u << 10, SW_STALE = 1u << 11, SW_SPECIAL = SW_INLINE | SW_STALE;
FQ_HD uint32_t make_w_seq(uint32_t cell /* newState | sym<<16 | nb<<24 */, unsigned ctx) {
  const unsigned sym = (cell >> 16) & 3u, nb = cell >> 24, ns = cell & 0x7FFu;
  uint32_t w = ((0x54474341u >> (8 * sym)) & 0xFFu) | (sym << 8) | (nb << 12) | (ns << 16);
  if (ctx == 0x00 || ctx == 0x55 || ctx == 0xAA || ctx == 0xFF) w |= SW_INLINE | ((ctx & 3u) << 27);
  return w;
}

// ---- quality W ------------------------------------------------------------------
//  [0] UNSEEN (symbol outside the dense alphabet V)   [1] ZENT (S entry of a
//  zero-bit-run context: holds a state, not a cell)   [2,8) rank of the symbol in V
//  [8,12) nbBits   [12,23) newState base   [23,30) output char   [31] STALE
//  ZENT entry: [8,19) state   [20,22) run slot
constexpr uint32_t QW_UNSEEN = 1u, QW_ZENT = 2u, QW_STALE = 1u << 31;
// windowed layout only: the word behind the last column of every row; reading it means "the
// context (row, column) is outside the row's window": its state is not in shared memory
constexpr uint32_t QW_OOB = 1u << 30;
constexpr uint32_t QW_SPECIAL = QW_UNSEEN | QW_ZENT | QW_STALE | QW_OOB;
constexpr unsigned QROW_BYTES = 256;  // 64 entries per (max, eq) row
FQ_HD uint32_t make_w_qual(uint32_t cell, const uint8_t *rk /*[64] rank in V or 0xFF*/) {
  const unsigned sym = (cell >> 16) & 63u, nb = cell >> 24, ns = cell & 0x7FFu;
  const unsigned r = rk[sym];
  return (r == 0xFFu ? QW_UNSEEN : (r << 2)) | (nb << 8) | (ns << 12) | ((sym + QUAL_CHAR0) << 23);
}
FQ_HD uint32_t make_zent(unsigned state, unsigned slot) { return QW_ZENT | (state << 8) | (slot << 20); }
// run table of a zero-bit-run context, one 8-byte entry per state x:
//   .x = the W cell of x (an ordinary step from x)
//   .y = ZENT entry of the state reached after the k <= 15 zero-bit steps that start at x, with k in
//        bits [24,28) (k = 0: the cell of x is not a zero-bit cell of the dominant symbol)
constexpr uint32_t ZQ_K_SHIFT = 24, ZQ_K_MASK = 15u << ZQ_K_SHIFT;
constexpr unsigned ZQ_SLOT_BYTES = 8u << TAB_LOG;
FQ_HD uint32_t make_zq_hi(unsigned k, unsigned state_after, unsigned slot) { return make_zent(state_after, slot) | (k << ZQ_K_SHIFT); }
// dense id of ctx(q, q1, q2) for q, max(q1, q2) in V: row = rank(max) * 2 + eq
FQ_HD unsigned qual_dense_id(unsigned rank_mx, unsigned eq, unsigned rank_q) { return (rank_mx * 2 + eq) * 64 + rank_q; }
FQ_HD unsigned qual_ctx13(unsigned q, unsigned q1, unsigned q2) {  // calcContext, src/fse_quality.h:40-44
  return ((((q1 > q2 ? q1 : q2) << 6) + q) & 0xFFFu) + ((unsigned)(q1 == q2) << 12);
}

// host-only event counters for tests (drains, inline refreshes, runs, slow-path entries)
#if !defined(__CUDA_ARCH__) && defined(DEC2_STATS)
extern thread_local unsigned long long g_stats[12];
#define DEC2_COUNT(i) (++g_stats[i])
#else
#define DEC2_COUNT(i) ((void)0)
#endif

// ---- memory access: shared-window addresses on the device, a byte array on the host
#ifndef __CUDA_ARCH__
extern thread_local uint8_t *g_host_smem;
#endif
FQ_HD uint32_t sm_ld32(uint32_t a) {
#ifdef __CUDA_ARCH__
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
#else
  uint32_t v;
  memcpy(&v, g_host_smem + a, 4);
  return v;
#endif
}
FQ_HD void sm_st32(uint32_t a, uint32_t v) {
#ifdef __CUDA_ARCH__
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v));
#else
  memcpy(g_host_smem + a, &v, 4);
#endif
}
FQ_HD void sm_ld64(uint32_t a, uint32_t &x, uint32_t &y) {
#ifdef __CUDA_ARCH__
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
#else
  memcpy(&x, g_host_smem + a, 4);
  memcpy(&y, g_host_smem + a + 4, 4);
#endif
}
FQ_HD uint32_t sm_ld16(uint32_t a) {
#ifdef __CUDA_ARCH__
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
#else
  uint16_t v;
  memcpy(&v, g_host_smem + a, 2);
  return v;
#endif
}
FQ_HD uint32_t sm_ld8(uint32_t a) {
#ifdef __CUDA_ARCH__
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
#else
  return g_host_smem[a];
#endif
}
FQ_HD uint32_t gl_ld32(const uint32_t *p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
// S[a] <- *p, asynchronously (device: cp.async, lands after the L1/L2 latency; host: a FIFO
// that delivers after a few steps, so that the STALE protocol is exercised by the CPU tests too)
#ifndef __CUDA_ARCH__
struct HostAsync { uint32_t addr[64], val[64], due[64]; unsigned n, now; };
extern thread_local HostAsync g_host_async;
constexpr unsigned HOST_ASYNC_LATENCY = 5;
inline void host_async_deliver(bool all) {
  HostAsync &q = g_host_async;
  unsigned k = 0;
  for (unsigned i = 0; i < q.n; i++) {
    if (all || q.due[i] <= q.now) memcpy(g_host_smem + q.addr[i], &q.val[i], 4);
    else { q.addr[k] = q.addr[i]; q.val[k] = q.val[i]; q.due[k] = q.due[i]; k++; }
  }
  q.n = k;
}
#endif
FQ_HD void sm_async_ld32(uint32_t a, const uint32_t *p) {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(p));
#else
  HostAsync &q = g_host_async;
  if (q.n == 64) host_async_deliver(true);
  q.addr[q.n] = a; q.val[q.n] = *p; q.due[q.n] = q.now + HOST_ASYNC_LATENCY; q.n++;
#endif
}
FQ_HD void sm_async_wait_all() {  // every async copy of this thread has landed
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_all;" ::: "memory");
#else
  host_async_deliver(true);
#endif
}
FQ_HD void sm_async_tick() {  // host only: one decoding step of time passes
#ifndef __CUDA_ARCH__
  g_host_async.now++;
  host_async_deliver(false);
#endif
}
// S[a] holds a STALE marker: its refresh is in flight.  Poll the slot -- the wait then ends when
// THIS copy lands (it was issued some steps ago), whereas cp.async.wait_all waits for the
// youngest copy of the whole warp, a full L2 round trip.  Bounded: falls back to the wait.
template <uint32_t STALE_BIT>
FQ_HD uint32_t sm_resolve_stale(uint32_t a) {
#ifdef __CUDA_ARCH__
  for (int k = 0; k < 64; k++) {
    const uint32_t w = sm_ld32(a);
    if (!(w & STALE_BIT)) return w;
  }
#endif
  sm_async_wait_all();
  return sm_ld32(a);
}
FQ_HD unsigned hb32(unsigned v) {
#ifdef __CUDA_ARCH__
  return 31u - (unsigned)__clz((int)v);
#else
  return 31u - (unsigned)__builtin_clz(v);
#endif
}
FQ_HD uint32_t fshl(uint32_t lo, uint32_t hi, unsigned s) {  // (hi:lo << s) >> 32, s in [0,32]
#ifdef __CUDA_ARCH__
  return __funnelshift_l(lo, hi, s);
#else
  s &= 63;
  if (s == 0) return hi;
  if (s >= 32) return lo;
  return (hi << s) | (lo >> (32 - s));
#endif
}
FQ_HD bool warp_any(bool p) {
#ifdef __CUDA_ARCH__
  return __any_sync(0xffffffffu, p) != 0;
#else
  return p;
#endif
}
FQ_HD uint32_t warp_min(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __reduce_min_sync(0xffffffffu, v);
#else
  return v;
#endif
}

// ---- backward bit reader (BIT_DStream_t) -----------------
// Stream bit i lives at bit (floor_ + i) of the 32-bit word array w (the stream's
// address rounded down to 8 bytes).  {hi:lo} holds the next `avail` unread bits
// [P - avail, P), left-aligned, with P - avail a multiple of 32, so a refill
// appends exactly one word (nxw).  On the device the words reach nxw through a
// two-slot per-lane ring of word pairs in shared memory filled by cp.async two
// refills ahead: a global load into a register would stall the whole warp's next
// refill on its scoreboard.
struct BitReader {
  const uint32_t *w;
  uint32_t hi, lo, nxw;
  int avail, wq, last_word;
  int floor_;
  long long top_bit;
#ifdef __CUDA_ARCH__
  volatile uint32_t *ring;  // this stream's two word pairs: slot 0 = ring[0..1], slot 1 = ring[2..3]
#endif

  FQ_HD uint32_t load_word(int idx) const { return (idx >= 0 && idx <= last_word) ? gl_ld32(w + idx) : 0u; }
  FQ_HD void prefetch_pair(int pidx) {
#ifdef __CUDA_ARCH__
    volatile uint32_t *slot = ring + ((pidx & 1) << 1);
    if (pidx >= 0 && 2 * pidx + 1 <= last_word) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(const_cast<uint32_t *>(slot));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(w + 2 * pidx) : "memory");
      // own group: the refill that consumes the pair waits for the committed groups only, not
      // for the table refreshes issued since (those stay uncommitted until the next prefetch)
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    } else {
      slot[0] = load_word(2 * pidx);
      slot[1] = load_word(2 * pidx + 1);
    }
#else
    (void)pidx;
#endif
  }
  FQ_HD void seek(long long P) {
    const int tw = (int)((P - 1) >> 5);
    const unsigned kbits = (unsigned)(P - ((long long)tw << 5));  // 1..32 valid bits in the top word
    const uint32_t wt = load_word(tw), w1 = load_word(tw - 1);
    hi = fshl(w1, wt, 32 - kbits);
    lo = w1 << (32 - kbits);
    avail = 32 + (int)kbits;
    nxw = load_word(tw - 2);
    wq = tw - 3;
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    prefetch_pair(wq >> 1);
    prefetch_pair((wq >> 1) - 1);
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
  }
  // p[0..len) = the stream; returns false when there is no end mark (BIT_initDStream)
  FQ_HD bool init(const uint8_t *p, uint32_t len, void *ring_lane) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)7);
    floor_ = (int)(a & 7) * 8;
    hi = lo = nxw = 0; avail = 64; wq = -1; last_word = 0; top_bit = floor_;
#ifdef __CUDA_ARCH__
    ring = static_cast<volatile uint32_t *>(ring_lane);
    ring[0] = ring[1] = ring[2] = ring[3] = 0u;
#else
    (void)ring_lane;
#endif
    const unsigned last = len ? p[len - 1] : 0u;
    if (last == 0) return false;
    last_word = (int)(((long long)floor_ + (long long)(len - 1) * 8) >> 5);
    top_bit = (long long)floor_ + (long long)(len - 1) * 8 + hb32(last);
    if (top_bit > floor_) seek(top_bit);
    return true;
  }
  FQ_HD void refill() {
    const unsigned s = 32u - (unsigned)avail;  // 0..15
    hi |= fshl(nxw, 0u, s);
    lo |= nxw << s;
    avail += 32;
#ifdef __CUDA_ARCH__
    const int pidx = wq >> 1;
    if (wq & 1) {  // first touch of pair pidx (copy issued two refills ago); the other slot is free
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      nxw = ring[((pidx & 1) << 1) + 1];
      prefetch_pair(pidx - 1);
    } else {
      nxw = ring[(pidx & 1) << 1];
    }
#else
    nxw = load_word(wq);
#endif
    --wq;
  }
  // BIT_readBits, nb <= 16
  FQ_HD unsigned read(unsigned nb) {
    const unsigned v = fshl(hi, 0u, nb);  // top nb bits of hi (0 when nb == 0)
    hi = fshl(lo, hi, nb);
    lo <<= nb;
    avail -= (int)nb;
    if (avail <= 32) refill();
    return v;
  }
  // the same without the refill: the caller refills before its next read when avail <= 32
  // (the hot loops fold that test into their single "anything unusual" branch)
  FQ_HD unsigned read_nr(unsigned nb) {
    const unsigned v = fshl(hi, 0u, nb);
    hi = fshl(lo, hi, nb);
    lo <<= nb;
    avail -= (int)nb;
    return v;
  }
  FQ_HD bool low() const { return avail <= 32; }
  FQ_HD long long position() const { return 32ll * (wq + 2) + avail; }  // P
  FQ_HD bool finished() const { return position() == floor_; }            // BIT_endOfDStream
};

// ---- record cursor ------------------------------------------------------------
struct RecMeta { unsigned L, hl; uint32_t scan; };
FQ_HD RecMeta load_meta(const uint16_t *readlens, const uint16_t *hdr_lens, const uint32_t *recscan, unsigned r) {
  RecMeta m;
  m.L = readlens[r]; m.hl = hdr_lens[r]; m.scan = recscan[r];
  return m;
}

struct StreamArgs {
  const uint8_t *src;        // the stream
  uint32_t len;
  uint32_t rec0, n_rec;      // record range of the chunk (global record indices)
  const uint16_t *readlens, *hdr_lens;
  const uint32_t *recscan;   // exclusive prefix of the records' byte sizes
  char *out;                 // chunk start in the output
  const uint32_t *logs, *logsuf;
  const uint32_t *wtab;      // W table, row stride 1 << TAB_LOG
  void *ring;                // device: this lane's ring pair
  bool live;                 // false: idle lane (takes part in the warp votes only)
};

constexpr int UNROLL = 4;  // symbols per loop trip

// ---------------------------------------------------------------------------
// Sequence.  Shared memory: S = 256 words at sb (1 KB aligned), the four
// homopolymer tables at ht (4 x 2048 words).
// Returns true when the stream was consumed exactly.
// ---------------------------------------------------------------------------
FQ_HD bool decode_seq_stream(const StreamArgs &c, uint32_t sb, uint32_t ht) {
  BitReader br;
  bool ok = true;
  unsigned rr = 0;
  if (c.live) {
    ok = br.init(c.src, c.len, c.ring);
    if (!ok || br.top_bit - br.floor_ < (long long)c.logsuf[256]) {
      ok = false;
    } else {
      rr = c.n_rec;
      // FSE_Decoder::startChunk (src/fse_common.hpp:134-138): states for ctx N-1 .. 0
      for (unsigned cc = 256; cc > 0; --cc) {
        const unsigned x = br.read(c.logs[cc - 1]);
        sm_st32(sb + (cc - 1) * 4, gl_ld32(c.wtab + ((cc - 1) << TAB_LOG) + x));
      }
    }
  } else {
    br.w = nullptr; br.hi = br.lo = br.nxw = 0; br.avail = 64; br.wq = -1; br.last_word = -1; br.floor_ = 0; br.top_bit = 0;
#ifdef __CUDA_ARCH__
    br.ring = static_cast<volatile uint32_t *>(c.ring);
#endif
  }
  const uint32_t scan0 = c.live ? c.recscan[c.rec0] : 0u;
  RecMeta cur{0, 0, 0}, nxt{0, 0, 0};
  if (rr) cur = load_meta(c.readlens, c.hdr_lens, c.recscan, c.rec0 + rr - 1);
  if (rr > 1) nxt = load_meta(c.readlens, c.hdr_lens, c.recscan, c.rec0 + rr - 2);
  char *dst = c.out + (cur.scan - scan0) + cur.hl + 1;
  unsigned i = 0;
  const uint32_t a_init = sb + SEQ_INIT_CTX * 4;
  uint32_t a = a_init, W = sm_ld32(a);
  uint32_t T = sb | ((a >> 2) & 0xFCu);  // next S address without the symbol bits
  // W table biased so that ((a << 9) + ns) indexes it directly (a = sb + ctx * 4, sb < 2^18)
  const uint32_t *wadj = c.wtab - ((size_t)sb << (TAB_LOG - 2));
  // One symbol.  A taken branch costs ~25 cycles of fixed latency here (predicate -> BRA ->
  // BSYNC), so the common case has exactly one: everything unusual -- refill due, context
  // STALE, homopolymer context -- hides behind a single test.
  auto step = [&](char *o) {
    sm_async_tick();
    // mark the context before the speculative load of the next one (the two may coincide); a
    // context that is already STALE keeps its marker: its refresh may land at any moment
    if (!(W & SW_STALE)) sm_st32(a, SW_STALE);
    uint32_t an = T | (W & 0x300u);      // ctx' = (ctx >> 2) + (sym << 6), as a word address
    uint32_t Wn = sm_ld32(an);           // speculative: W may still turn out to be STALE
    if (((W & SW_SPECIAL) != 0) | br.low()) {
      if (br.low()) br.refill();
      if (W & SW_STALE) {                // refresh in flight: wait for it and start over
        DEC2_COUNT(0);
        W = sm_resolve_stale<SW_STALE>(a);
        sm_st32(a, SW_STALE);
        an = T | (W & 0x300u);
        Wn = sm_ld32(an);
      }
      if (W & SW_INLINE) {               // homopolymer context: table in shared memory
        DEC2_COUNT(1);
        *o = (char)(W & 0xFFu);
        const unsigned ns = ((W >> 16) & 0x7FFu) + br.read_nr((W >> 12) & 15u);
        const uint32_t W2 = sm_ld32(ht + (((W >> 27) & 3u) << (TAB_LOG + 2)) + ns * 4);
        sm_st32(a, W2);
        if (an == a) Wn = W2;
        T = sb | ((an >> 2) & 0xFCu);
        a = an;
        W = Wn;
        return;
      }
    }
    *o = (char)(W & 0xFFu);
    const unsigned ns = ((W >> 16) & 0x7FFu) + br.read_nr((W >> 12) & 15u);
    sm_async_ld32(a, wadj + ((a << (TAB_LOG - 2)) + ns));  // S[a] <- cell of the new state, when it arrives
    T = sb | ((an >> 2) & 0xFCu);
    a = an;
    W = Wn;
  };
  // Four symbols at once.  The contexts and cells of the four symbols are chained through
  // shared-memory loads WITHOUT side effects, then validated in one go -- no special cell, no
  // context twice among them (a later load must not see an unmarked earlier context), enough bits
  // in the window -- and only then committed.  Straight-line code lets the in-order pipeline
  // overlap the four dependent loads with the bit arithmetic; an invalid block is redone by
  // step(), which takes every case.
  auto block = [&](char *o) -> unsigned {
    if (br.low()) br.refill();
    const uint32_t a0 = a, W0 = W;
    const uint32_t a1 = T | (W0 & 0x300u);
    const uint32_t W1 = sm_ld32(a1);
    const uint32_t a2 = sb | ((a1 >> 2) & 0xFCu) | (W1 & 0x300u);
    const uint32_t W2 = sm_ld32(a2);
    const uint32_t a3 = sb | ((a2 >> 2) & 0xFCu) | (W2 & 0x300u);
    const uint32_t W3 = sm_ld32(a3);
    const uint32_t a4 = sb | ((a3 >> 2) & 0xFCu) | (W3 & 0x300u);
    const uint32_t W4 = sm_ld32(a4);
    // bit reads on a copy of the window
    uint32_t hi = br.hi, lo = br.lo;
    const unsigned n0 = (W0 >> 12) & 15u, n1 = (W1 >> 12) & 15u, n2 = (W2 >> 12) & 15u, n3 = (W3 >> 12) & 15u;
    const unsigned v0 = fshl(hi, 0u, n0); hi = fshl(lo, hi, n0); lo <<= n0;
    const unsigned v1 = fshl(hi, 0u, n1); hi = fshl(lo, hi, n1); lo <<= n1;
    const unsigned v2 = fshl(hi, 0u, n2); hi = fshl(lo, hi, n2); lo <<= n2;
    const unsigned v3 = fshl(hi, 0u, n3); hi = fshl(lo, hi, n3); lo <<= n3;
    const int avail = br.avail - (int)(n0 + n1 + n2 + n3);
    const bool bad = (((W0 | W1 | W2 | W3) & SW_SPECIAL) != 0) | (avail < 0) | (a2 == a0) | (a3 == a0) | (a3 == a1) |
                     (a4 == a0) | (a4 == a1) | (a4 == a2);
    if (bad) {  // (one careful step and a fresh block attempt was measured slower: 43.5 vs 35.7 ms per GB)
      DEC2_COUNT(8);
#pragma unroll
      for (int u = 0; u < UNROLL; u++) step(o + u);
      return 4u;
    }
    DEC2_COUNT(9);
    sm_async_tick(); sm_async_tick(); sm_async_tick(); sm_async_tick();
    sm_st32(a0, SW_STALE);
    sm_st32(a1, SW_STALE);
    sm_st32(a2, SW_STALE);
    sm_st32(a3, SW_STALE);
    o[0] = (char)(W0 & 0xFFu);
    o[1] = (char)(W1 & 0xFFu);
    o[2] = (char)(W2 & 0xFFu);
    o[3] = (char)(W3 & 0xFFu);
    sm_async_ld32(a0, wadj + ((a0 << (TAB_LOG - 2)) + ((W0 >> 16) & 0x7FFu) + v0));
    sm_async_ld32(a1, wadj + ((a1 << (TAB_LOG - 2)) + ((W1 >> 16) & 0x7FFu) + v1));
    sm_async_ld32(a2, wadj + ((a2 << (TAB_LOG - 2)) + ((W2 >> 16) & 0x7FFu) + v2));
    sm_async_ld32(a3, wadj + ((a3 << (TAB_LOG - 2)) + ((W3 >> 16) & 0x7FFu) + v3));
    br.hi = hi; br.lo = lo; br.avail = avail;
    a = a4;
    W = W4;
    T = sb | ((a4 >> 2) & 0xFCu);
    return 4u;
  };
  static_assert(UNROLL == 4, "block() is written for four symbols");
  for (;;) {
    // uniform trip count: symbols until the first lane of the warp reaches a record end
    const unsigned n = warp_min(rr > 0 ? cur.L - i : 0xFFFFFFFFu);
    if (n == 0xFFFFFFFFu) break;
    if (rr > 0) {
      char *o = dst + i;
      unsigned t = n;
      while (t >= UNROLL) {
        const unsigned k = block(o);
        o += k;
        t -= k;
      }
      for (; t; --t, ++o) step(o);
      i += n;
      if (i >= cur.L) {  // record done: records n-1 .. 0 (src/workspace.cpp:84-87)
        --rr;
        cur = nxt;
        i = 0;
        dst = c.out + (cur.scan - scan0) + cur.hl + 1;
        if (rr > 1) nxt = load_meta(c.readlens, c.hdr_lens, c.recscan, c.rec0 + rr - 2);
        a = a_init;
        W = sm_ld32(a);
        T = sb | ((a >> 2) & 0xFCu);
      }
    }
  }
  return !c.live || (ok && br.finished());
}


// ---------------------------------------------------------------------------
// Quality.  Shared per CTA: rk[64] (q -> rank in V, 0xFF outside) at rk_a; the run
// tables (n_z x 2048 entries of 8 bytes, see make_zq_hi) at zq_a; zc (n_z words:
// output char of run slot j) at zc_a.  Per stream: S = 2 * |V| rows of 64 words at sb
// (256-byte aligned).  V = {0} + the quality values that occur
// in a context of the sample: every context made of values in V has a dense id
// (row = rank(max) * 2 + eq, column = rank(q)).  Contexts outside the dense set
// keep their state in `cold` (global, u16[8192]) and decode through dtab_fix.
// ---------------------------------------------------------------------------
struct QualShared {
  uint32_t rk_a, zq_a, zc_a;
  uint32_t n_slots;   // run tables present at zq_a; 0 = the run contexts are decoded as ordinary contexts
  uint32_t row_a;     // windowed layout: row descriptors, 8 bytes per row (row = rank(max) * 2 + eq)
  uint32_t n_rows;    //   .x = byte offset of the row's first entry in S, .y = lo * 4 | (width * 4) << 16
};

// WIN = windowed layout of S for many-valued qualities.  With |V| = 40 values the dense layout
// takes 2 * 40 rows of 64 columns = 20 KB per stream, of which a few hundred entries are contexts
// the tables know; shared memory then holds 7 streams per SM.  The windowed layout keeps, for every
// row, only the columns [lo, lo + width) that span the row's touched contexts, plus one QW_OOB
// word; a context outside its row's window keeps its state in `cold` and decodes through the
// slow path, like a quality value outside V.  Address of (row, rank r): base + min(4r - 4lo, 4width)
// -- two more ALU operations in the symbol chain than the dense layout's OR.  No run tables.
template <bool WIN>
FQ_HD bool decode_qual_stream(const StreamArgs &c, const QualShared &q, uint32_t sb,
                              const uint32_t *dtab_fix, const uint16_t *cid /*[8192]: run slot + 1 in bits 13..15*/,
                              uint16_t *cold) {
  BitReader br;
  bool ok = true;
  unsigned rr = 0;
  if (c.live) {
    ok = br.init(c.src, c.len, c.ring);
    if (!ok || br.top_bit - br.floor_ < (long long)c.logsuf[8192]) {
      ok = false;
    } else {
      rr = c.n_rec;
      if (WIN) {
        for (unsigned r = 0; r < q.n_rows; r++) {
          uint32_t rx, ry;
          sm_ld64(q.row_a + r * 8, rx, ry);
          sm_st32(sb + rx + (ry >> 16), QW_OOB);
        }
      }
      // FSE_Decoder::startChunk (src/fse_common.hpp:134-138): states for ctx N-1 .. 0
      for (unsigned cc = 8192; cc > 0; --cc) {
        const unsigned cx = cc - 1;
        const unsigned x = br.read(c.logs[cx]);
        const unsigned rq = sm_ld8(q.rk_a + (cx & 63u)), rm = sm_ld8(q.rk_a + ((cx >> 6) & 63u));
        if (WIN) {
          bool in_s = false;
          if (rq != 0xFFu && rm != 0xFFu) {
            uint32_t rx, ry;
            sm_ld64(q.row_a + (rm * 2 + (cx >> 12)) * 8, rx, ry);
            const uint32_t rel = rq * 4 - (ry & 0xFFFFu);
            if (rel < (ry >> 16)) {
              const uint32_t off = rx + rel;   // byte offset in S = compact id * 4
              sm_st32(sb + off, gl_ld32(c.wtab + ((size_t)(off >> 2) << TAB_LOG) + x));
              in_s = true;
            }
          }
          if (!in_s) cold[cx] = (uint16_t)x;
        } else if (rq != 0xFFu && rm != 0xFFu) {
          const unsigned d = qual_dense_id(rm, cx >> 12, rq);
          const unsigned cv = cid[cx];
          const unsigned zs = (cv == 0xFFFFu || q.n_slots == 0) ? 0u : (cv >> 13);
          sm_st32(sb + d * 4, zs ? make_zent(x, zs - 1) : gl_ld32(c.wtab + ((size_t)d << TAB_LOG) + x));
        } else {
          cold[cx] = (uint16_t)x;
        }
      }
    }
  } else {
    br.w = nullptr; br.hi = br.lo = br.nxw = 0; br.avail = 64; br.wq = -1; br.last_word = -1; br.floor_ = 0; br.top_bit = 0;
#ifdef __CUDA_ARCH__
    br.ring = static_cast<volatile uint32_t *>(c.ring);
#endif
  }
  const uint32_t scan0 = c.live ? c.recscan[c.rec0] : 0u;
  RecMeta cur{0, 0, 0}, nxt{0, 0, 0};
  if (rr) cur = load_meta(c.readlens, c.hdr_lens, c.recscan, c.rec0 + rr - 1);
  if (rr > 1) nxt = load_meta(c.readlens, c.hdr_lens, c.recscan, c.rec0 + rr - 2);
  char *dst = c.out + (cur.scan - scan0) + cur.hl + 1 + cur.L + 3;
  unsigned i = 0;
  // record start: the three previous symbols are 0 (rank 0): row (max 0, eq 1), column 0
  // Windowed layout: a row is (R1 = address of its first entry, RW = lo * 4 | (width * 4) << 16);
  // the table builder keeps column 0 of row 1 inside its window.
  uint32_t RW = 0, RW_init = 0;
  auto row_of = [&](uint32_t r4a, uint32_t r4b) -> uint32_t {  // row of (max, eq) of two symbols, ranks * 4
    const uint32_t mx4 = r4a > r4b ? r4a : r4b;
    if (WIN) {
      uint32_t rx, ry;
      sm_ld64(q.row_a + mx4 * 4 + (r4a == r4b ? 8u : 0u), rx, ry);
      RW = ry;
      return sb + rx;
    }
    return sb + mx4 * (QROW_BYTES / 2) + (r4a == r4b ? QROW_BYTES : 0u);
  };
  // entry of column rank r4 / 4 in the row (R1, RW): the OOB word when outside the window
  auto col_of = [&](uint32_t row, uint32_t rw, uint32_t r4) -> uint32_t {
    if (WIN) {
      const uint32_t rel = r4 - (rw & 0xFFFFu), w4 = rw >> 16;
      return row + (rel < w4 ? rel : w4);
    }
    return row | r4;
  };
  uint32_t row_init = sb + QROW_BYTES;
  if (WIN) {
    row_init = row_of(0, 0);
    RW_init = RW;
  }
  uint32_t a = col_of(row_init, RW_init, 0), W = 0;
  const uint32_t a_init = a;
  uint32_t R1 = row_init;   // row of the NEXT symbol's context (from the two symbols before this one)
  uint32_t r4p = 0;         // rank * 4 of the previous symbol
  const uint32_t *wadj = c.wtab - ((size_t)sb << (TAB_LOG - 2));  // ((a << 9) + ns) indexes it directly
  // Slow path, entered after a symbol outside V (a quality value the sample never showed):
  // explicit q values, blocking table loads, until the record ends or the last three
  // symbols are in V again.  o = the record's quality slot, t = next position, rem = record length.
  auto cold_run = [&](char *o, unsigned t, unsigned rem) -> unsigned {
    sm_async_wait_all();
    unsigned qa = (unsigned)(unsigned char)o[t - 1] - QUAL_CHAR0;
    unsigned qb = t >= 2 ? (unsigned)(unsigned char)o[t - 2] - QUAL_CHAR0 : 0u;
    unsigned qc = t >= 3 ? (unsigned)(unsigned char)o[t - 3] - QUAL_CHAR0 : 0u;
    for (;;) {
      const unsigned ra = sm_ld8(q.rk_a + qa), rb = sm_ld8(q.rk_a + qb), rc = sm_ld8(q.rk_a + qc);
      if (ra != 0xFFu && rb != 0xFFu && rc != 0xFFu) {  // back to the fast path
        const uint32_t row = row_of(rb * 4, rc * 4);
        const uint32_t an = col_of(row, RW, ra * 4);
        const uint32_t wn = sm_ld32(an);
        if (!WIN || !(wn & QW_OOB)) {   // (windowed: ... if this context is inside its row's window)
          a = an;
          R1 = row_of(ra * 4, rb * 4);
          r4p = ra * 4;
          W = wn;
          return t;
        }
      }
      if (t >= rem) return t;
      const unsigned mx = qb > qc ? qb : qc, eq = qb == qc;
      const unsigned rm = sm_ld8(q.rk_a + mx);
      unsigned sym;
      uint32_t aa = 0;
      bool in_s = ra != 0xFFu && rm != 0xFFu;
      if (in_s) {
        if (WIN) {
          uint32_t rx, ry;
          sm_ld64(q.row_a + (rm * 2 + eq) * 8, rx, ry);
          const uint32_t rel = ra * 4 - (ry & 0xFFFFu);
          in_s = rel < (ry >> 16);
          aa = sb + rx + rel;
        } else {
          aa = sb + qual_dense_id(rm, eq, ra) * 4;
        }
      }
      if (in_s) {  // dense context: its cell (or run state) is in S
        DEC2_COUNT(6);
        const unsigned d = (aa - sb) >> 2;
        uint32_t w = sm_ld32(aa);
        if (w & QW_ZENT) {
          const unsigned zs = (w >> 20) & 3u;
          w = sm_ld32(q.zq_a + zs * ZQ_SLOT_BYTES + ((w >> 8) & 0x7FFu) * 8);
          sm_st32(aa, make_zent(((w >> 12) & 0x7FFu) + br.read((w >> 8) & 15u), zs));
        } else {
          const unsigned ns = ((w >> 12) & 0x7FFu) + br.read((w >> 8) & 15u);
          sm_st32(aa, gl_ld32(c.wtab + ((size_t)d << TAB_LOG) + ns));
        }
        sym = ((w >> 23) & 0x7Fu) - QUAL_CHAR0;
      } else {
        DEC2_COUNT(7);
        const unsigned cx = qual_ctx13(qa, qb, qc);
        const uint32_t e = gl_ld32(dtab_fix + ((size_t)cx << TAB_LOG) + cold[cx]);
        sym = (e >> 16) & 63u;
        cold[cx] = (uint16_t)((e & 0xFFFFu) + br.read(e >> 24));
      }
      o[t++] = (char)(sym + QUAL_CHAR0);
      qc = qb; qb = qa; qa = sym;
    }
  };
  // Everything that is not "a plain cell, bits at hand": lane idle (never here), refill due, context
  // STALE, run context, symbol outside V.  Out of the hot loop on purpose -- the loop below is a
  // dozen instructions around one shared-memory load, and a taken branch costs ~25 cycles.
  // Returns the new position; W / a / R1 / r4p are left ready for the next symbol.
  auto special = [&](char *o, unsigned t, unsigned rem) -> unsigned {
    if (br.low()) br.refill();
    if (WIN && (W & QW_OOB)) {   // the context of this symbol is outside its row's window
      DEC2_COUNT(5);
      if (t == 0) return rem + 1;   // (cannot happen: column 0 of the record-start row is always kept) -> stream error
      return cold_run(o, t, rem);
    }
    if (W & QW_STALE) {
      DEC2_COUNT(2);
      W = sm_resolve_stale<QW_STALE>(a);
    }
    uint32_t an;
    if (W & QW_ZENT) {
      const unsigned zs = (W >> 20) & 3u, x = (W >> 8) & 0x7FFu;
      uint32_t cell, hi;
      sm_ld64(q.zq_a + zs * ZQ_SLOT_BYTES + x * 8, cell, hi);
      unsigned kz = (hi >> ZQ_K_SHIFT) & 15u;
      if (kz) {  // the next kz symbols are d, read no bits, and stay in this context
        DEC2_COUNT(3);
        unsigned xs = (hi >> 8) & 0x7FFu;
        if (kz > rem - t) {  // the record ends inside the run: single steps
          kz = rem - t;
          xs = x;
          for (unsigned j = 0; j < kz; j++) xs = (sm_ld32(q.zq_a + zs * ZQ_SLOT_BYTES + xs * 8) >> 12) & 0x7FFu;
        }
        W = make_zent(xs, zs);
        sm_st32(a, W);
        const char dc = (char)sm_ld32(q.zc_a + zs * 4);
        if (t + 15 <= rem) {  // the bytes behind the run are rewritten by the record's later symbols
#pragma unroll
          for (unsigned j = 0; j < 15; j++) o[t + j] = dc;
        } else {
          for (unsigned j = 0; j < kz; j++) o[t + j] = dc;
        }
        return t + kz;
      }
      // one ordinary step in the run context: cell from the shared table, refreshed inline
      DEC2_COUNT(4);
      const unsigned ns = ((cell >> 12) & 0x7FFu) + br.read_nr((cell >> 8) & 15u);
      sm_st32(a, make_zent(ns, zs));
      W = cell;
    } else {
      const unsigned ns = ((W >> 12) & 0x7FFu) + br.read_nr((W >> 8) & 15u);
      sm_st32(a, QW_STALE);
      sm_async_ld32(a, wadj + ((a << (TAB_LOG - 2)) + ns));
    }
    o[t] = (char)((W >> 23) & 0x7Fu);
    if (W & QW_UNSEEN) {  // a quality value the sample never showed: explicit q values from here on
      DEC2_COUNT(5);
      return cold_run(o, t + 1, rem);   // comes back with a / R1 / r4p / W set, or at the record end
    }
    const uint32_t r4 = W & 0xFCu;
    an = col_of(R1, RW, r4);
    R1 = row_of(r4, r4p);
    r4p = r4;
    a = an;
    W = sm_ld32(a);
    return t + 1;
  };
  for (;;) {
    // uniform trip: every lane decodes until it has passed n symbols (n = symbols left in the
    // warp's shortest current record); runs may overshoot n but never the lane's own record
    const unsigned n = warp_min(rr > 0 ? cur.L - i : 0xFFFFFFFFu);
    if (n == 0xFFFFFFFFu) break;
    if (rr > 0) {
      char *o = dst;                     // positions are absolute inside the record
      const unsigned rem = cur.L, stop = i + n;
      unsigned t = i;
      W = sm_ld32(a);
      while (t < stop) {
        sm_async_tick();
        if (W & QW_ZENT) {
          // run context (binned qualities: more than half of all events): one 8-byte shared load
          // yields either "k symbols d, no bits, state after them" or the cell of an ordinary step
          const unsigned zs = (W >> 20) & 3u;
          uint32_t cell, hi;
          sm_ld64(q.zq_a + zs * ZQ_SLOT_BYTES + ((W >> 8) & 0x7FFu) * 8, cell, hi);
          const unsigned kz = (hi >> ZQ_K_SHIFT) & 15u;
          if ((kz != 0) & (t + 15 <= rem)) {
            DEC2_COUNT(3);
            const char dc = (char)sm_ld32(q.zc_a + zs * 4);
#pragma unroll
            for (unsigned j = 0; j < 15; j++) o[t + j] = dc;   // the bytes behind the run are rewritten later
            W = hi & ~ZQ_K_MASK;
            sm_st32(a, W);
            t += kz;
            continue;
          }
          if ((kz == 0) & !br.low() & !(cell & QW_UNSEEN)) {
            DEC2_COUNT(4);
            o[t] = (char)((cell >> 23) & 0x7Fu);
            const unsigned ns = ((cell >> 12) & 0x7FFu) + br.read_nr((cell >> 8) & 15u);
            sm_st32(a, make_zent(ns, zs));
            const uint32_t r4 = cell & 0xFCu;
            a = col_of(R1, RW, r4);
            W = sm_ld32(a);
            R1 = row_of(r4, r4p);
            r4p = r4;
            ++t;
            continue;
          }
        } else if (!(((W & QW_SPECIAL) != 0) | br.low())) {
          // the hot path: a cell whose symbol is in V, bits at hand
          o[t] = (char)((W >> 23) & 0x7Fu);
          const unsigned ns = ((W >> 12) & 0x7FFu) + br.read_nr((W >> 8) & 15u);
          sm_st32(a, QW_STALE);
          sm_async_ld32(a, wadj + ((a << (TAB_LOG - 2)) + ns));   // S[a] <- cell of the new state, when it arrives
          const uint32_t r4 = W & 0xFCu;
          a = col_of(R1, RW, r4);
          W = sm_ld32(a);
          R1 = row_of(r4, r4p);
          r4p = r4;
          ++t;
          continue;
        }
        t = special(o, t, rem);   // (one call site: the slow paths are long, keep them out of the loop body)
      }
      i = t;
      if (i >= cur.L) {
        --rr;
        cur = nxt;
        i = 0;
        dst = c.out + (cur.scan - scan0) + cur.hl + 1 + cur.L + 3;
        if (rr > 1) nxt = load_meta(c.readlens, c.hdr_lens, c.recscan, c.rec0 + rr - 2);
        a = a_init;
        R1 = row_init;
        RW = RW_init;
        r4p = 0;
      }
    }
  }
  return !c.live || (ok && br.finished());
}

}  // namespace dec2
}  // namespace fq28
