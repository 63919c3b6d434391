// fq28_encode.cu -- K2 (field separation) and K5 (tANS encode) for every chunk
// of a slab at once.  Replaces CompressionWorkspace::encodeChunk
// (src/workspace.cpp:14-45), replaceAndEncodeNs + SequenceEncoder::encodeRecord
// (src/fse_sequence.cpp:35-112), QualityEncoder::encodeRecord
// (src/fse_quality.cpp:5-53) and FSE_Encoder::startChunk/endChunk
// (src/fse_common.hpp:77-90).
//
// A chunk stream is the LSB-first concatenation of one bit field per symbol in
// ENCODE ORDER (records forward, positions backward, src/workspace.cpp:25-31 +
// src/fse_sequence.cpp:83), followed by the final state of every context
// 0..N-1 and a 1 end-mark bit.  The state of context
// c only evolves over the symbols coded in c, so the work splits into
//   extract   : key[g] = (ctx, sym) for every symbol, g = encode-order index
//   partition : per tile, stable counting sort of the symbols by context
//               (k_tile_part8 for the 256 sequence contexts; histogram +
//               compact-cursor rank for the 8192 quality contexts)
//   chain     : one thread per (chunk, context) walks its symbols in order
//               through the context's CTable held in shared memory and emits
//               (nbBits, bits) per symbol; long chains of contexts with a
//               dominant symbol go through composed step tables (k_chain_dom)
//   pack      : gather the fields back into encode order, prefix-sum nbBits,
//               OR the fields into the output words
#include <stdlib.h>

#include <algorithm>

#include "fq28_internal.cuh"

namespace fq28 {

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
// largest k with arr[k] <= v, arr ascending with arr[0] <= v < arr[n]
__device__ __forceinline__ unsigned find_chunk(const uint32_t *__restrict__ arr, unsigned n, unsigned v) {
  unsigned lo = 0, hi = n;
  while (hi - lo > 1) {
    const unsigned mid = (lo + hi) >> 1;
    if (arr[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------
// K2: extract.  One warp per record.
//   key_seq[g]  = ctx(8) << 2 | sym(2)      (N coded as A, src/fse_sequence.cpp:45)
//   key_qual[g] = ctx(13) << 6 | sym(6)
//   n_count[r]  = number of N in the read    (src/fse_sequence.cpp:36-50)
// g = symoff[r] + (L-1-i): symbols of a record are coded last-to-first.
// ---------------------------------------------------------------------------
constexpr int EX_WARPS = 8;

__global__ void __launch_bounds__(EX_WARPS * 32)
k_extract(const char *__restrict__ d, const uint32_t *__restrict__ seq_off, const uint32_t *__restrict__ qual_off,
          const uint16_t *__restrict__ len, const uint32_t *__restrict__ symoff, size_t n_rec,
          uint16_t *__restrict__ key_seq, uint32_t *__restrict__ key_qual, uint16_t *__restrict__ n_count,
          DevStatus *st) {
  const unsigned lane = threadIdx.x & 31;
  const size_t r = (size_t)blockIdx.x * EX_WARPS + (threadIdx.x >> 5);
  if (r >= n_rec) return;
  const unsigned L = len[r];
  const unsigned char *sp = reinterpret_cast<const unsigned char *>(d) + seq_off[r];
  const unsigned char *qp = reinterpret_cast<const unsigned char *>(d) + qual_off[r];
  const uint32_t g_last = symoff[r] + L - 1;  // g of position 0
  if (L < 3) {  // src/fse_quality.cpp:42-52 mis-codes L < 3
    if (lane == 0) { set_error(st, FQ28_ERR_SHORT, (unsigned)r); n_count[r] = 0; }
    return;
  }
  unsigned n_cnt = 0;
  bool bad = false;
  // Every lane loads its own base and quality once; the neighbours come from
  // warp votes / shuffles.  Sequence: the two bit planes of the 32 bases of a
  // step are ballots, (p, m) = planes of the previous and the current step, so
  // the four bases before position i are four consecutive bits of (m:p).  The
  // virtual prefix b[-4..-1] = T,C,C,T (0xD7) seeds the top bits of p.
  unsigned p0 = 0xF0000000u, p1 = 0x90000000u, prevq = 0;
  auto spread4 = [](unsigned x) {  // bit j -> bit 2j
    unsigned t = (x | (x << 2)) & 0x33u;
    return (t | (t << 1)) & 0x55u;
  };
  // All loads of (up to) EX_BATCH steps are issued before the first vote: a warp then has
  // 2 * EX_BATCH independent loads in flight instead of one dependent pair per step (round 1:
  // SM busy 78 % at 22 % of DRAM bandwidth, one record per warp walking its 5 steps serially).
  constexpr unsigned EX_BATCH = 5;  // 160 symbols: a whole 150 bp read
  unsigned cb[EX_BATCH], qb[EX_BATCH];
  for (unsigned base0 = 0; base0 < L; base0 += 32 * EX_BATCH) {
#pragma unroll
  for (unsigned u = 0; u < EX_BATCH; u++) {
    const unsigned i = base0 + 32 * u + lane;
    cb[u] = i < L ? sp[i] : (unsigned)'A';
    qb[u] = i < L ? qp[i] : QUAL_OFFSET;
  }
#pragma unroll
  for (unsigned u = 0; u < EX_BATCH; u++) {
    const unsigned base = base0 + 32 * u;
    if (base >= L) break;
    const unsigned i = base + lane;
    const bool in = i < L;
    const unsigned c = cb[u];
    const unsigned qc = qb[u];
    const bool is_n = c == 'N';
    const unsigned dl = c - 'A';
    if (!(dl < 26u && ((0x82045u >> dl) & 1u))) bad = true;   // A C G N T
    const unsigned bits = is_n ? 0u : ((c >> 1) ^ (c >> 2)) & 3u;  // A C G T -> 0 1 2 3, N coded as A
    const unsigned m0 = __ballot_sync(0xffffffffu, bits & 1u), m1 = __ballot_sync(0xffffffffu, bits >> 1);
    const unsigned x0 = (lane >= 4 ? m0 >> (lane - 4) : __funnelshift_r(p0, m0, 28 + lane)) & 15u;
    const unsigned x1 = (lane >= 4 ? m1 >> (lane - 4) : __funnelshift_r(p1, m1, 28 + lane)) & 15u;
    const unsigned ctx = spread4(x0) | (spread4(x1) << 1);  // b[i-1]<<6 | b[i-2]<<4 | b[i-3]<<2 | b[i-4]
    p0 = m0;
    p1 = m1;
    // quality: ctx = calcContext(q[i-1], q[i-2], q[i-3]), q[<0] = 0
    const unsigned q = qc - QUAL_OFFSET;
    if (q > 63u) bad = true;
    const unsigned pk = (prevq << 8) | (q & 63u);
    const unsigned s1 = __shfl_sync(0xffffffffu, pk, (lane - 1) & 31);
    const unsigned s2 = __shfl_sync(0xffffffffu, pk, (lane - 2) & 31);
    const unsigned s3 = __shfl_sync(0xffffffffu, pk, (lane - 3) & 31);
    const unsigned q0 = (lane >= 1 ? s1 : s1 >> 8) & 63u;
    const unsigned q1 = (lane >= 2 ? s2 : s2 >> 8) & 63u;
    const unsigned q2 = (lane >= 3 ? s3 : s3 >> 8) & 63u;
    prevq = q & 63u;
    if (in) {
      key_seq[g_last - i] = (uint16_t)((ctx << 2) | bits);
      key_qual[g_last - i] = (qual_ctx(q0, q1, q2) << 6) | (q & 63u);
    }
    n_cnt += __popc(__ballot_sync(0xffffffffu, in && is_n));
  }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) set_error(st, FQ28_ERR_ALPHABET, (unsigned)r);
  if (lane == 0) n_count[r] = (uint16_t)n_cnt;
}

// n_pos deltas: delta = i - prev_n_pos, prev_n_pos starts at 0
// (src/fse_sequence.cpp:38-48).  One warp per record, skipped when N-free.
__global__ void __launch_bounds__(EX_WARPS * 32)
k_npos(const char *__restrict__ d, const uint32_t *__restrict__ seq_off, const uint16_t *__restrict__ len,
       const uint16_t *__restrict__ n_count, const uint32_t *__restrict__ npos_off, size_t n_rec,
       uint16_t *__restrict__ n_pos) {
  const unsigned lane = threadIdx.x & 31;
  const size_t r = (size_t)blockIdx.x * EX_WARPS + (threadIdx.x >> 5);
  if (r >= n_rec) return;
  if (n_count[r] == 0) return;
  const unsigned L = len[r];
  const unsigned char *sp = reinterpret_cast<const unsigned char *>(d) + seq_off[r];
  uint32_t o = npos_off[r];
  unsigned prev = 0;
  for (unsigned base = 0; base < L; base += 32) {
    const unsigned i = base + lane;
    const bool is_n = i < L && sp[i] == 'N';
    const unsigned m = __ballot_sync(0xffffffffu, is_n);
    if (is_n) {
      const unsigned below = m & ((1u << lane) - 1u);
      const unsigned p = below ? base + (31u - (unsigned)__clz(below)) : prev;
      n_pos[o + __popc(below)] = (uint16_t)(i - p);
    }
    if (m) {
      prev = base + (31u - (unsigned)__clz(m));
      o += __popc(m);
    }
  }
}

// Header lines ('@'.., without '\n') back to back: the input of the host-side
// header tokeniser (src/workspace.cpp:95-125, out of path).  Half a warp per
// record (Illumina headers are 40-70 bytes).
__global__ void __launch_bounds__(EX_WARPS * 32)
k_gather_headers(const char *__restrict__ d, const uint32_t *__restrict__ hdr_off, const uint16_t *__restrict__ hdr_len,
                 const uint32_t *__restrict__ hdrscan, size_t n_rec, uint8_t *__restrict__ out) {
  const unsigned sub = threadIdx.x & 15;
  const size_t r = ((size_t)blockIdx.x * EX_WARPS * 32 + threadIdx.x) >> 4;
  if (r >= n_rec) return;
  const unsigned hl = hdr_len[r];
  const char *src = d + hdr_off[r];
  uint8_t *dst = out + hdrscan[r];
  for (unsigned i = sub; i < hl; i += 16) dst[i] = (uint8_t)src[i];
}

// ---------------------------------------------------------------------------
// partition
// ---------------------------------------------------------------------------
template <typename KEY, unsigned N, unsigned SHIFT>
struct Kind {
  using key_t = KEY;
  static constexpr unsigned n_models = N;
  static constexpr unsigned shift = SHIFT;       // key >> shift = ctx
  static constexpr unsigned sym_mask = (1u << SHIFT) - 1u;
};
using SeqKind = Kind<uint16_t, SEQ_N, 2>;
using QualKind = Kind<uint32_t, QUAL_N, 6>;

// tile t -> (chunk, first symbol g0, symbol count)
struct TileRef { unsigned chunk, g0, cnt; };
__device__ __forceinline__ TileRef tile_ref(unsigned t, const uint32_t *__restrict__ tile0,
                                           const uint32_t *__restrict__ chunk_sym, unsigned n_chunks,
                                           unsigned tile_syms) {
  TileRef tr;
  tr.chunk = find_chunk(tile0, n_chunks, t);
  tr.g0 = chunk_sym[tr.chunk] + (t - tile0[tr.chunk]) * tile_syms;
  const unsigned end = chunk_sym[tr.chunk + 1];
  tr.cnt = end - tr.g0 < tile_syms ? end - tr.g0 : tile_syms;
  return tr;
}

// Eight consecutive u32 starting at p[idx] (idx arbitrary): three aligned
// 128-bit loads and a uniform-per-CTA rotation instead of eight strided 32-bit loads.
__device__ __forceinline__ void load8_u32(const uint32_t *__restrict__ p, size_t idx, unsigned (&out)[8]) {
  const size_t a = idx & ~(size_t)3;
  const unsigned off = (unsigned)(idx & 3);
  const uint4 *v = reinterpret_cast<const uint4 *>(p + a);
  const uint4 x0 = __ldg(v), x1 = __ldg(v + 1);
  uint4 x2 = make_uint4(0, 0, 0, 0);
  if (off) x2 = __ldg(v + 2);
  const unsigned w[12] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w};
  switch (off) {  // sym0 is a per-chunk constant: the branch is uniform
    case 0:
#pragma unroll
      for (int i = 0; i < 8; i++) out[i] = w[i];
      break;
    case 1:
#pragma unroll
      for (int i = 0; i < 8; i++) out[i] = w[i + 1];
      break;
    case 2:
#pragma unroll
      for (int i = 0; i < 8; i++) out[i] = w[i + 2];
      break;
    default:
#pragma unroll
      for (int i = 0; i < 8; i++) out[i] = w[i + 3];
      break;
  }
}

// Per tile: context histogram, then exclusive scan over contexts of the counts
// rounded up to 16 (every context's run starts 16-byte aligned so that the
// chain kernel can move it with 128-bit loads):
// tbase[t][c] = first slot of context c inside the tile's partitioned region
// (a multiple of 16) | (count & 15); tbase[t][N] = padded size of the region.
template <class K, unsigned TILE>
__global__ void __launch_bounds__(256)
k_tile_hist(const typename K::key_t *__restrict__ key, const uint32_t *__restrict__ tile0,
            const uint32_t *__restrict__ chunk_sym, unsigned n_chunks, uint32_t *__restrict__ tbase,
            uint32_t *__restrict__ present /*[N] flags of the contexts seen in the slab, or null*/) {
  constexpr unsigned N = K::n_models;
  __shared__ uint32_t hist[N];
  __shared__ uint32_t wsum[9];
  const TileRef tr = tile_ref(blockIdx.x, tile0, chunk_sym, n_chunks, TILE);
  for (unsigned i = threadIdx.x; i < N; i += 256) hist[i] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31;
  // every thread walks 8 consecutive keys and issues one shared-memory atomic
  // per RUN of equal contexts (quality contexts come in runs), the loads of a
  // step are independent and in flight together
  const typename K::key_t *kp = key + tr.g0;
  for (unsigned j0 = threadIdx.x * 8; j0 < tr.cnt; j0 += 256 * 8) {
    unsigned c[8];
    if constexpr (sizeof(typename K::key_t) == 4) {
      if (j0 + 8 <= tr.cnt) {
        load8_u32(reinterpret_cast<const uint32_t *>(key), (size_t)tr.g0 + j0, c);
#pragma unroll
        for (unsigned u = 0; u < 8; u++) c[u] >>= K::shift;
      } else {
#pragma unroll
        for (unsigned u = 0; u < 8; u++) c[u] = j0 + u < tr.cnt ? (unsigned)kp[j0 + u] >> K::shift : 0xFFFFFFFFu;
      }
    } else {
#pragma unroll
      for (unsigned u = 0; u < 8; u++) c[u] = j0 + u < tr.cnt ? (unsigned)kp[j0 + u] >> K::shift : 0xFFFFFFFFu;
    }
    unsigned cur = c[0], n = 1;
#pragma unroll
    for (unsigned u = 1; u < 8; u++) {
      if (c[u] == cur) { n++; }
      else {
        if (cur != 0xFFFFFFFFu) atomicAdd(&hist[cur], n);
        cur = c[u];
        n = 1;
      }
    }
    if (cur != 0xFFFFFFFFu) atomicAdd(&hist[cur], n);
  }
  __syncthreads();
  // exclusive scan of hist[N]; thread i owns N/256 consecutive entries
  constexpr unsigned PER = N / 256;
  unsigned local[PER];
  unsigned s = 0;
#pragma unroll
  for (unsigned i = 0; i < PER; i++) { local[i] = hist[threadIdx.x * PER + i]; s += (local[i] + 15u) & ~15u; }
  if (present) {
#pragma unroll
    for (unsigned i = 0; i < PER; i++)
      if (local[i]) present[threadIdx.x * PER + i] = 1u;  // same value from every tile: plain stores
  }
  unsigned inc = s;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    unsigned o = __shfl_up_sync(0xffffffffu, inc, dd);
    if (lane >= (unsigned)dd) inc += o;
  }
  if (lane == 31) wsum[threadIdx.x >> 5] = inc;
  __syncthreads();
  unsigned woff = 0;
#pragma unroll
  for (unsigned i = 0; i < 8; i++) woff += (i < (threadIdx.x >> 5)) ? wsum[i] : 0u;
  unsigned ex = woff + inc - s;
  uint32_t *out = tbase + (size_t)blockIdx.x * (N + 1);
#pragma unroll
  for (unsigned i = 0; i < PER; i++) { out[threadIdx.x * PER + i] = ex | (local[i] & 15u); ex += (local[i] + 15u) & ~15u; }
  if (threadIdx.x == 255) out[N] = ex;
}

constexpr unsigned RANKC_CAP = 1024, RANKC_WARPS = 8;  // k_tile_rank_compact

// Stable rank: one warp per tile walks the tile in encode order, 32 symbols a
// step; run[c] is the next free slot of context c.  Emits
//   ssym[t * STRIDE + slot] = symbol    (partitioned symbols, chain input)
//   perm[g]                 = t * STRIDE + slot   (where the symbol's field will be)
template <class K, unsigned TILE, unsigned STRIDE, unsigned WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_tile_rank(const typename K::key_t *__restrict__ key, const uint32_t *__restrict__ tile0,
            const uint32_t *__restrict__ chunk_sym, unsigned n_chunks, unsigned n_tiles,
            const uint32_t *__restrict__ tbase, uint8_t *__restrict__ ssym, uint32_t *__restrict__ perm,
            const uint32_t *__restrict__ n_present /*null, or: skip when k_tile_rank_compact covers the slab*/) {
  constexpr unsigned N = K::n_models;
  __shared__ uint32_t run_s[WARPS][N];
  if (n_present && *n_present <= RANKC_CAP) return;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned t = blockIdx.x * WARPS + warp;
  if (t >= n_tiles) return;
  uint32_t *run = run_s[warp];
  const TileRef tr = tile_ref(t, tile0, chunk_sym, n_chunks, TILE);
  const uint32_t *tb = tbase + (size_t)t * (N + 1);
  for (unsigned i = lane; i < N; i += 32) run[i] = tb[i] & ~15u;
  __syncwarp();
  const typename K::key_t *kp = key + tr.g0;
  // keys are prefetched four steps (128 symbols) ahead: a step takes a few
  // hundred cycles, DRAM latency is several steps
  unsigned q0 = lane < tr.cnt ? (unsigned)kp[lane] : 0u;
  unsigned q1 = 32 + lane < tr.cnt ? (unsigned)kp[32 + lane] : 0u;
  unsigned q2 = 64 + lane < tr.cnt ? (unsigned)kp[64 + lane] : 0u;
  unsigned q3 = 96 + lane < tr.cnt ? (unsigned)kp[96 + lane] : 0u;
  for (unsigned j = 0; j < tr.cnt; j += 32) {
    const unsigned kv = q0;
    q0 = q1; q1 = q2; q2 = q3;
    const unsigned jn = j + 128 + lane;
    q3 = jn < tr.cnt ? (unsigned)kp[jn] : 0u;
    const bool live = j + lane < tr.cnt;
    const unsigned ctx = live ? kv >> K::shift : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xffffffffu, ctx);
    unsigned slot = 0;
    if (live) slot = run[ctx];
    __syncwarp();
    const unsigned below = peers & ((1u << lane) - 1u);
    if (live && below == 0) run[ctx] = slot + (unsigned)__popc(peers);
    __syncwarp();
    if (live) {
      const unsigned dst = t * STRIDE + slot + (unsigned)__popc(below);
      ssym[dst] = (uint8_t)(kv & K::sym_mask);
      perm[tr.g0 + j + lane] = dst;
    }
  }
}

// Compact ids of the contexts present in the slab (flags from k_tile_hist):
// map[c] = id or 0xFFFF, inv[id] = c, *n_present = number of ids.  Single CTA.
template <unsigned N>
__global__ void __launch_bounds__(1024)
k_present_compact(const uint32_t *__restrict__ present, uint16_t *__restrict__ map, uint16_t *__restrict__ inv,
                  uint32_t *__restrict__ n_present) {
  constexpr unsigned PER = N / 1024;
  static_assert(N % 1024 == 0, "contexts per thread");
  __shared__ unsigned wsum[32];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned f[PER], s = 0;
#pragma unroll
  for (unsigned i = 0; i < PER; i++) { f[i] = present[threadIdx.x * PER + i] ? 1u : 0u; s += f[i]; }
  unsigned inc = s;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, inc, dd);
    if (lane >= (unsigned)dd) inc += o;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  unsigned woff = 0;
  for (unsigned i = 0; i < warp; i++) woff += wsum[i];
  unsigned id = woff + inc - s;
#pragma unroll
  for (unsigned i = 0; i < PER; i++) {
    const unsigned c = threadIdx.x * PER + i;
    if (f[i]) { map[c] = (uint16_t)id; inv[id] = (uint16_t)c; id++; } else map[c] = 0xFFFFu;
  }
  if (threadIdx.x == 1023) *n_present = id;
}

// Stable rank with compact cursors: as k_tile_rank, but the next-free-slot
// cursors are indexed by the compact id of the context, so a warp needs
// 4 B x RANKC_CAP of shared memory instead of 4 B x N and every SM holds
// 32 tiles in flight instead of 7 (the pass is bound by the per-step latency).
// Runs when at most RANKC_CAP contexts are present; k_tile_rank covers the rest.
template <class K>
constexpr size_t rankc_smem() { return (size_t)K::n_models * 2 + (size_t)RANKC_WARPS * RANKC_CAP * 4; }

template <class K, unsigned TILE, unsigned STRIDE>
__global__ void __launch_bounds__(RANKC_WARPS * 32)
k_tile_rank_compact(const typename K::key_t *__restrict__ key, const uint32_t *__restrict__ tile0,
                    const uint32_t *__restrict__ chunk_sym, unsigned n_chunks, unsigned n_tiles,
                    const uint32_t *__restrict__ tbase, const uint16_t *__restrict__ gmap,
                    const uint16_t *__restrict__ inv, const uint32_t *__restrict__ n_present,
                    uint8_t *__restrict__ ssym, uint32_t *__restrict__ perm) {
  constexpr unsigned N = K::n_models;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t *map = reinterpret_cast<uint16_t *>(smem_raw);                 // [N]
  uint32_t *run_all = reinterpret_cast<uint32_t *>(smem_raw + N * 2);     // [RANKC_WARPS][RANKC_CAP]
  const unsigned ne = *n_present;
  if (ne > RANKC_CAP) return;
  for (unsigned i = threadIdx.x; i < N / 2; i += RANKC_WARPS * 32)
    reinterpret_cast<uint32_t *>(map)[i] = reinterpret_cast<const uint32_t *>(gmap)[i];
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned t = blockIdx.x * RANKC_WARPS + warp;
  if (t >= n_tiles) return;
  uint32_t *run = run_all + warp * RANKC_CAP;
  const TileRef tr = tile_ref(t, tile0, chunk_sym, n_chunks, TILE);
  const uint32_t *tb = tbase + (size_t)t * (N + 1);
  for (unsigned i = lane; i < ne; i += 32) run[i] = tb[inv[i]] & ~15u;
  __syncwarp();
  const typename K::key_t *kp = key + tr.g0;
  // keys are prefetched four steps (128 symbols) ahead, the compact id of a
  // key is looked up one step ahead
  unsigned q0 = lane < tr.cnt ? (unsigned)kp[lane] : 0u;
  unsigned q1 = 32 + lane < tr.cnt ? (unsigned)kp[32 + lane] : 0u;
  unsigned q2 = 64 + lane < tr.cnt ? (unsigned)kp[64 + lane] : 0u;
  unsigned q3 = 96 + lane < tr.cnt ? (unsigned)kp[96 + lane] : 0u;
  unsigned idn = lane < tr.cnt ? map[q0 >> K::shift] : 0xFFFFu;
  for (unsigned j = 0; j < tr.cnt; j += 32) {
    const unsigned kv = q0, id = idn;
    q0 = q1; q1 = q2; q2 = q3;
    const unsigned jn = j + 128 + lane;
    q3 = jn < tr.cnt ? (unsigned)kp[jn] : 0u;
    idn = j + 32 + lane < tr.cnt ? map[q0 >> K::shift] : 0xFFFFu;
    const bool live = j + lane < tr.cnt;
    const unsigned peers = __match_any_sync(0xffffffffu, id);
    unsigned slot = 0;
    if (live) slot = run[id];
    __syncwarp();
    const unsigned below = peers & ((1u << lane) - 1u);
    if (live && below == 0) run[id] = slot + (unsigned)__popc(peers);
    __syncwarp();
    if (live) {
      const unsigned dst = t * STRIDE + slot + (unsigned)__popc(below);
      ssym[dst] = (uint8_t)(kv & K::sym_mask);
      perm[tr.g0 + j + lane] = dst;
    }
  }
}

// Cooperative fused partition for a small context space (sequence): CTA = 16
// warps = one tile; warp w ranks the tile's w-th sixteenth in steps of 32
// consecutive symbols.  A step ranks its symbols with match_any against the
// warp's private context counters (u16[N] in shared memory); afterwards a
// scan over (context, warp) turns the counters into bases, and every symbol's
// slot is base + rank.  Same outputs as k_tile_hist + k_tile_rank, stable in
// encode order: (warp, step, lane) is ascending g.  The symbols of the tile
// are staged in shared memory and stored with 128-bit writes.
constexpr unsigned PART8_WARPS = 16;
template <class K, unsigned TILE, unsigned STRIDE>
struct Part8 {
  static constexpr unsigned N = K::n_models;
  static constexpr unsigned PER_WARP = TILE / PART8_WARPS;
  static constexpr unsigned STEPS = PER_WARP / 32;
  static_assert(N <= PART8_WARPS * 32 && N % 32 == 0 && TILE % (PART8_WARPS * 32) == 0 && STRIDE % 16 == 0 && STRIDE < 65536, "layout");
  static constexpr size_t SMEM = (size_t)PART8_WARPS * N * 2 + (size_t)N * 4 + STRIDE + 64;
};

template <class K, unsigned TILE, unsigned STRIDE>
__global__ void __launch_bounds__(PART8_WARPS * 32, 3)
k_tile_part8(const typename K::key_t *__restrict__ key, const uint32_t *__restrict__ tile0,
             const uint32_t *__restrict__ chunk_sym, unsigned n_chunks, unsigned n_tiles,
             uint32_t *__restrict__ tbase, uint8_t *__restrict__ ssym, uint32_t *__restrict__ perm) {
  using P = Part8<K, TILE, STRIDE>;
  constexpr unsigned N = P::N, STEPS = P::STEPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t *cnt = reinterpret_cast<uint16_t *>(smem_raw);                          // [8][N]
  uint32_t *cbase = reinterpret_cast<uint32_t *>(smem_raw + PART8_WARPS * N * 2);  // [N]
  uint8_t *region = smem_raw + PART8_WARPS * N * 2 + N * 4;                        // [STRIDE]
  __shared__ unsigned wsum[PART8_WARPS + 1];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned t = blockIdx.x;
  const TileRef tr = tile_ref(t, tile0, chunk_sym, n_chunks, TILE);
  for (unsigned i = threadIdx.x; i < PART8_WARPS * N / 2; i += PART8_WARPS * 32) reinterpret_cast<uint32_t *>(cnt)[i] = 0u;
  // the warp's keys, one per step and lane, all loads in flight together
  const typename K::key_t *kp = key + tr.g0;
  const unsigned j0 = warp * P::PER_WARP + lane;
  unsigned kr[STEPS];  // key | rank << 16
#pragma unroll
  for (unsigned i = 0; i < STEPS; i++) {
    const unsigned j = j0 + i * 32;
    kr[i] = j < tr.cnt ? (unsigned)kp[j] : 0xFFFFu;
  }
  __syncthreads();
  uint16_t *mycnt = cnt + warp * N;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (unsigned i = 0; i < STEPS; i++) {
    const unsigned kv = kr[i];
    const bool live = kv != 0xFFFFu;
    const unsigned ctx = live ? kv >> K::shift : 0u;
    // lanes with the same context: one vote per context bit (cheaper than match_any)
    unsigned peers = __ballot_sync(0xffffffffu, live);
#pragma unroll
    for (unsigned bit = 0; (1u << bit) < N; bit++) {
      const unsigned vote = __ballot_sync(0xffffffffu, (ctx >> bit) & 1u);
      peers &= ((ctx >> bit) & 1u) ? vote : ~vote;
    }
    unsigned old = 0;
    if (live) old = mycnt[ctx];
    __syncwarp();
    const unsigned below = peers & lt;
    if (live && below == 0) mycnt[ctx] = (uint16_t)(old + (unsigned)__popc(peers));
    __syncwarp();
    kr[i] = kv | ((old + (unsigned)__popc(below)) << 16);
  }
  __syncthreads();
  {  // thread c < N: exclusive scan of context c's counts over the warps, then of the padded totals over the contexts
    const unsigned c = threadIdx.x;
    unsigned run = 0;
    if (c < N) {
#pragma unroll
      for (unsigned w = 0; w < PART8_WARPS; w++) {
        const unsigned v = cnt[w * N + c];
        cnt[w * N + c] = (uint16_t)run;
        run += v;
      }
    }
    const unsigned padded = (run + 15u) & ~15u;
    unsigned inc = padded;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, inc, dd);
      if (lane >= (unsigned)dd) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned woff = 0;
#pragma unroll
    for (unsigned w = 0; w < N / 32; w++) woff += (w < warp) ? wsum[w] : 0u;
    const unsigned base = woff + inc - padded;
    if (c < N) {
      cbase[c] = base;
      uint32_t *tb = tbase + (size_t)t * (N + 1);
      tb[c] = base | (run & 15u);
      if (c == N - 1) { tb[N] = base + padded; wsum[PART8_WARPS] = base + padded; }
    }
  }
  __syncthreads();
  const unsigned t_slot0 = t * STRIDE;
  uint32_t *pp = perm + tr.g0;
#pragma unroll
  for (unsigned i = 0; i < STEPS; i++) {
    const unsigned kv = kr[i] & 0xFFFFu;
    if (kv != 0xFFFFu) {
      const unsigned ctx = kv >> K::shift;
      const unsigned slot = cbase[ctx] + mycnt[ctx] + (kr[i] >> 16);
      region[slot] = (uint8_t)(kv & K::sym_mask);
      pp[j0 + i * 32] = t_slot0 + slot;
    }
  }
  __syncthreads();
  const unsigned carry = wsum[PART8_WARPS];  // padded size of the tile's region (multiple of 16)
  const uint4 *rsrc = reinterpret_cast<const uint4 *>(region);
  uint4 *rdst = reinterpret_cast<uint4 *>(ssym + (size_t)t_slot0);
  for (unsigned i = threadIdx.x; i < (carry >> 4); i += PART8_WARPS * 32) rdst[i] = rsrc[i];
}

// symbols of a context inside a tile, from the packed tbase entries
__device__ __forceinline__ unsigned run_count(const uint32_t *__restrict__ tb) {
  const unsigned a0 = tb[0], a1 = tb[1];
  const unsigned padded = (a1 & ~15u) - (a0 & ~15u);
  return (a0 & 15u) ? padded - 16u + (a0 & 15u) : padded;
}
constexpr unsigned DOM_MIN = 16384;  // chain length from which a (chunk, context) pair goes to k_chain_dom

// ---------------------------------------------------------------------------
// chain: FSE_encodeSymbol (Appendix A.5) along the symbols of one context.
// CTA = one warp = one context x 32 consecutive chunks; the context's CTable
// (next-state cells + symbol transforms) sits in shared memory, so all 32
// lanes walk the same table and chain lengths within a warp are similar.
// field[slot] = nbBits << 12 | low bits.
//
// The lanes advance in ROUNDS of one 16-symbol group each: every lane loads its
// next group with one 128-bit load at the start of a round -- all lanes in the
// same instruction, one full round (16 dependent table steps) before the data
// is used -- and ends the round with two 128-bit stores of the 16 fields.  The
// only latency on the critical path is the shared-memory next-state lookup.
// ---------------------------------------------------------------------------
template <class K, unsigned A, unsigned TILE, unsigned STRIDE>
__global__ void __launch_bounds__(32)
k_chain(const uint8_t *__restrict__ ssym, const uint32_t *__restrict__ tile0, unsigned n_chunks,
        const uint32_t *__restrict__ tbase, const uint32_t *__restrict__ logs, const uint32_t *__restrict__ toff,
        const uint16_t *__restrict__ ctab, const int2 *__restrict__ symtt, uint16_t *__restrict__ field,
        uint16_t *__restrict__ fstate, const int8_t *__restrict__ dom_sym) {
  constexpr unsigned N = K::n_models;
  __shared__ uint16_t st[1u << FSE_MAX_TABLELOG];
  __shared__ int2 tt[A];
  const unsigned c = blockIdx.x, lane = threadIdx.x;
  const unsigned k = blockIdx.y * 32 + lane;
  bool live = k < n_chunks;
  const unsigned t_log = logs[c], T = 1u << t_log;
  unsigned t = 0, t_end = 0;
  if (live) { t = tile0[k]; t_end = tile0[k + 1]; }
  if (dom_sym && dom_sym[c] >= 0 && live) {  // long chains of dominant-symbol contexts go to k_chain_dom
    unsigned total = 0;
    for (unsigned tt_ = t; tt_ < t_end; tt_++) total += run_count(tbase + (size_t)tt_ * (N + 1) + c);
    if (total >= DOM_MIN) { live = false; t = t_end; }
  }
  // run cursor
  unsigned grp_left = 0, last_valid = 16;
  size_t slot = 0;                       // slot of the next group to load
  uint4 nxtv = make_uint4(0, 0, 0, 0);   // the group loaded in the previous round
  unsigned nxt_valid = 0;
  size_t nxt_slot = 0;
  // tbase entries of the next tile to open, requested one tile ahead: the lanes of a warp
  // change tiles at different rounds, an unprefetched pair of loads would stall every round
  unsigned na0 = 0, na1 = 0;
  if (t < t_end) {
    const uint32_t *tb = tbase + (size_t)t * (N + 1) + c;
    na0 = tb[0];
    na1 = tb[1];
  }
  auto fetch = [&]() {                   // advance to the next non-empty run if needed, load one group
    while (grp_left == 0 && t < t_end) {
      const unsigned a0 = na0, a1 = na1;
      const unsigned tcur = t;
      ++t;
      if (t < t_end) {
        const uint32_t *tb = tbase + (size_t)t * (N + 1) + c;
        na0 = tb[0];
        na1 = tb[1];
      }
      const unsigned b0 = a0 & ~15u, padded = (a1 & ~15u) - b0;
      if (padded) {
        grp_left = padded >> 4;
        last_valid = (a0 & 15u) ? (a0 & 15u) : 16u;
        slot = (size_t)tcur * STRIDE + b0;
      }
    }
    nxt_valid = 0;
    if (grp_left) {
      nxtv = __ldg(reinterpret_cast<const uint4 *>(ssym + slot));
      nxt_valid = grp_left == 1 ? last_valid : 16u;
      nxt_slot = slot;
      slot += 16;
      --grp_left;
    }
  };
  fetch();
  unsigned x = T;  // FSE_initCState, src/fse_common.hpp:82
  if (__any_sync(0xffffffffu, nxt_valid != 0)) {
    const uint16_t *gs = ctab + toff[c];
    for (unsigned i = lane; i < T; i += 32) st[i] = gs[i];
    for (unsigned i = lane; i < A; i += 32) tt[i] = symtt[(size_t)c * A + i];
    __syncwarp();
    while (__any_sync(0xffffffffu, nxt_valid != 0)) {
      const uint4 cur = nxtv;
      const unsigned valid = nxt_valid;
      const size_t cur_slot = nxt_slot;
      fetch();  // next round's group: issued now, consumed after 16 table steps
      if (valid) {
        const unsigned w[4] = {cur.x, cur.y, cur.z, cur.w};
        unsigned f[8];
#pragma unroll
        for (unsigned i = 0; i < 16; i++) {
          const unsigned s = (w[i >> 2] >> (8 * (i & 3))) & (A - 1);  // padding bytes are masked into range
          const int2 a = tt[s];
          const unsigned nb = (x + (unsigned)a.y) >> 16;
          const unsigned fv = (nb << 12) | (x & ((1u << nb) - 1u));
          const unsigned xn = st[(int)(x >> nb) + a.x];
          if (i < valid) x = xn;
          if (i & 1) f[i >> 1] |= fv << 16; else f[i >> 1] = fv;
        }
        uint4 *fp = reinterpret_cast<uint4 *>(field + cur_slot);
        fp[0] = make_uint4(f[0], f[1], f[2], f[3]);
        fp[1] = make_uint4(f[4], f[5], f[6], f[7]);
      }
    }
  }
  if (live) fstate[(size_t)k * N + c] = (uint16_t)x;
}

// ---------------------------------------------------------------------------
// Long chains in contexts with a DOMINANT symbol (norm > T/2): binned qualities
// put > 80 % of a chunk's symbols into one such chain, and its length bounds
// the whole encoder.  The state recurrence x' = f_s(x) is split in two:
//   * a serial WALK that only tracks the state.  Composed step tables in shared
//     memory -- R_r = r consecutive dominant symbols (r = 1..16), S_j = one step
//     with the j-th most frequent other symbol -- turn a 16-symbol group into
//     ~2 table lookups (one per run of dominant symbols, one per other symbol);
//   * a parallel EXPANSION: given the state at the start of its group, every
//     lane replays its 16 symbols with the ordinary FSE_encodeSymbol step and
//     stores the 16 fields.
// CTA = 8 warps = one context x 8 consecutive chunks (the tables depend on the
// context only); a warp processes its chain in blocks of 32 groups.
// ---------------------------------------------------------------------------
constexpr unsigned DOM_WARPS = 8;
constexpr unsigned DOM_NS = 3;                     // other symbols with a step table of their own
constexpr unsigned DOM_IDENT = 16 + DOM_NS;          // op code of the identity table (pads the op list)
constexpr unsigned DOM_TABLES = 16 + DOM_NS + 1;
constexpr unsigned DOM_GENERIC = 0x100;            // op code: ordinary step with symbol (op & 0xFF)
constexpr unsigned DOM_OPS = 512 + 16;             // ops per block + padding to 8 + look-ahead
constexpr size_t DOM_SMEM = (size_t)DOM_TABLES * 4096 + 4096 + 64 * 8 + (size_t)DOM_WARPS * DOM_OPS * 2 * 2;

__device__ __forceinline__ unsigned lds_u16(unsigned saddr) {  // tables are read-only while they are walked
  unsigned short v;
  asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
  return v;
}

template <class K>
__global__ void k_dom_list(const uint32_t *__restrict__ tile0, unsigned n_chunks, const uint32_t *__restrict__ tbase,
                           const int8_t *__restrict__ dom_sym, uint32_t *__restrict__ list, unsigned cap,
                           unsigned long long *__restrict__ count) {
  constexpr unsigned N = K::n_models;
  const unsigned c = blockIdx.x;
  if (dom_sym[c] < 0) return;
  const unsigned k = blockIdx.y * blockDim.x + threadIdx.x;
  unsigned total = 0;
  if (k < n_chunks)
    for (unsigned t = tile0[k]; t < tile0[k + 1]; t++) total += run_count(tbase + (size_t)t * (N + 1) + c);
  // one entry per block of DOM_WARPS consecutive chunks with at least one long chain
  const unsigned any = __ballot_sync(0xffffffffu, total >= DOM_MIN);
  const unsigned lane = threadIdx.x & 31;
  if ((lane % DOM_WARPS) == 0 && ((any >> lane) & ((1u << DOM_WARPS) - 1u))) {
    const unsigned long long i = atomicAdd(count, 1ULL);
    if (i < cap) list[i] = c | ((k / DOM_WARPS) << 13);
  }
}

template <class K, unsigned A, unsigned STRIDE>
__global__ void __launch_bounds__(DOM_WARPS * 32)
k_chain_dom(const uint32_t *__restrict__ list, const unsigned long long *__restrict__ count,
            const uint8_t *__restrict__ ssym, const uint32_t *__restrict__ tile0, unsigned n_chunks,
            const uint32_t *__restrict__ tbase, const uint32_t *__restrict__ logs, const uint32_t *__restrict__ toff,
            const uint16_t *__restrict__ ctab, const int2 *__restrict__ symtt, const int16_t *__restrict__ norm,
            const int8_t *__restrict__ dom_sym, uint16_t *__restrict__ field, uint16_t *__restrict__ fstate) {
  constexpr unsigned N = K::n_models;
  static_assert(A <= 64, "symbol transform table sized for the quality alphabet");
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned char *tbl = dsm;                                                  // [DOM_TABLES][2048] u16: 2 * (x' - T)
  uint16_t *st = reinterpret_cast<uint16_t *>(dsm + (size_t)DOM_TABLES * 4096);  // [2048]
  int2 *tt = reinterpret_cast<int2 *>(st + 2048);                            // [64]
  uint16_t *ops_all = reinterpret_cast<uint16_t *>(tt + 64);                 // [DOM_WARPS][DOM_OPS]
  uint16_t *xs_all = ops_all + DOM_WARPS * DOM_OPS;                          // [DOM_WARPS][DOM_OPS]
  __shared__ unsigned nsym[DOM_NS];
  if (blockIdx.x >= *count) return;
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned ent = list[blockIdx.x];
  const unsigned c = ent & 0x1FFFu, k = (ent >> 13) * DOM_WARPS + warp;
  const unsigned t_log = logs[c], T = 1u << t_log;
  const unsigned sdom = (unsigned)dom_sym[c];
  {
    const uint16_t *gs = ctab + toff[c];
    for (unsigned i = tid; i < T; i += DOM_WARPS * 32) st[i] = gs[i];
    for (unsigned i = tid; i < A; i += DOM_WARPS * 32) tt[i] = symtt[(size_t)c * A + i];
    if (tid == 0) {  // the DOM_NS most frequent symbols besides the dominant one
      unsigned best[DOM_NS];
      int bn[DOM_NS];
      for (unsigned j = 0; j < DOM_NS; j++) { best[j] = 0xFFu; bn[j] = 0; }
      for (unsigned s = 0; s < A; s++) {
        int v = norm[(size_t)c * A + s];
        if (v < 0) v = 1;
        if (s == sdom || v == 0) continue;
        unsigned cs = s;
        for (unsigned j = 0; j < DOM_NS; j++)
          if (v > bn[j]) { const int tv = bn[j]; const unsigned ts = best[j]; bn[j] = v; best[j] = cs; v = tv; cs = ts; }
      }
      for (unsigned j = 0; j < DOM_NS; j++) nsym[j] = best[j];
    }
  }
  __syncthreads();
  auto step_from = [&](unsigned x, unsigned s) -> unsigned {  // FSE_encodeSymbol state update
    const int2 a = tt[s];
    const unsigned nb = (x + (unsigned)a.y) >> 16;
    return st[(int)(x >> nb) + a.x];
  };
  {
    uint16_t *R1 = reinterpret_cast<uint16_t *>(tbl);
    for (unsigned i = tid; i < T; i += DOM_WARPS * 32) {
      R1[i] = (uint16_t)((step_from(T + i, sdom) - T) * 2);
      reinterpret_cast<uint16_t *>(tbl + (size_t)DOM_IDENT * 4096)[i] = (uint16_t)(i * 2);
      for (unsigned j = 0; j < DOM_NS; j++)
        if (nsym[j] != 0xFFu)
          reinterpret_cast<uint16_t *>(tbl + (size_t)(16 + j) * 4096)[i] = (uint16_t)((step_from(T + i, nsym[j]) - T) * 2);
    }
    __syncthreads();
    for (unsigned r = 1; r < 16; r++) {  // R_{r+1} = R_1 o R_r
      const uint16_t *prev = reinterpret_cast<const uint16_t *>(tbl + (size_t)(r - 1) * 4096);
      uint16_t *cur = reinterpret_cast<uint16_t *>(tbl + (size_t)r * 4096);
      for (unsigned i = tid; i < T; i += DOM_WARPS * 32) cur[i] = R1[prev[i] >> 1];
      __syncthreads();
    }
  }
  // only chains k_chain left out are walked here (same criterion)
  if (k >= n_chunks) return;
  const unsigned t_begin = tile0[k], t_end = tile0[k + 1];
  {
    unsigned total = 0;
    for (unsigned t = t_begin; t < t_end; t++) total += run_count(tbase + (size_t)t * (N + 1) + c);
    if (total < DOM_MIN) return;
  }
  uint16_t *ops = ops_all + warp * DOM_OPS;
  uint16_t *xs = xs_all + warp * DOM_OPS;
  unsigned tbl_s;  // shared-window address of the tables, pinned in a register (not rematerialised per use)
  {
    const unsigned t0 = (unsigned)__cvta_generic_to_shared(tbl);
    asm volatile("mov.u32 %0, %1;" : "=r"(tbl_s) : "r"(t0));
  }
  const unsigned pat = sdom * 0x01010101u;
  const unsigned ns0 = nsym[0], ns1 = nsym[1], ns2 = nsym[2];
  unsigned xo = 0;  // 2 * (x - T); FSE_initCState: x = T
  for (unsigned t = t_begin; t < t_end; t++) {
    const uint32_t *tb = tbase + (size_t)t * (N + 1) + c;
    const unsigned a0 = tb[0], a1 = tb[1];
    const unsigned b0 = a0 & ~15u, padded = (a1 & ~15u) - b0;
    if (!padded) continue;
    const unsigned groups = padded >> 4;
    const unsigned last_valid = (a0 & 15u) ? (a0 & 15u) : 16u;
    const size_t slot0 = (size_t)t * STRIDE + b0;
    uint4 wn = make_uint4(0, 0, 0, 0);
    if (lane < groups) wn = __ldg(reinterpret_cast<const uint4 *>(ssym + slot0 + (size_t)lane * 16));
    for (unsigned blk0 = 0; blk0 < groups; blk0 += 32) {
      const unsigned g = blk0 + lane;
      const bool have = g < groups;
      const uint4 w = wn;
      if (g + 32 < groups)  // next block's group: in flight during this block's walk
        wn = __ldg(reinterpret_cast<const uint4 *>(ssym + slot0 + (size_t)(g + 32) * 16));
      const unsigned valid = have ? (g == groups - 1 ? last_valid : 16u) : 0u;
      // dominance mask of the group's valid symbols
      unsigned m = 0;
      {
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (unsigned q = 0; q < 4; q++)
          m |= (((__vcmpeq4(ww[q], pat) & 0x08040201u) * 0x01010101u) >> 24) << (4 * q);
        m &= (1u << valid) - 1u;
      }
      // ops of the group: one per run of dominant symbols, one per other symbol
      const unsigned n_ops = (valid - __popc(m)) + __popc(m & ~(m << 1));
      unsigned ofs = n_ops;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, ofs, d);
        if (lane >= (unsigned)d) ofs += o;
      }
      const unsigned total_ops = __shfl_sync(0xffffffffu, ofs, 31);
      ofs -= n_ops;
      {
        unsigned o = ofs;
        // op starts: every other symbol, and the first symbol of every dominant run
        for (unsigned starts = (~m & ((1u << valid) - 1u)) | (m & ~(m << 1)); starts; starts &= starts - 1) {
          const unsigned i = __ffs(starts) - 1;
          unsigned code;
          if ((m >> i) & 1u) {
            code = __ffs(~(m >> i)) - 2;  // run length - 1; the run ends at `valid` at the latest
          } else {
            const unsigned word = i < 8 ? (i < 4 ? w.x : w.y) : (i < 12 ? w.z : w.w);
            const unsigned sy = (word >> (8 * (i & 3))) & (A - 1);
            code = sy == ns0 ? 16u : sy == ns1 ? 17u : sy == ns2 ? 18u : (DOM_GENERIC | sy);
          }
          ops[o++] = (uint16_t)code;
        }
        if (lane < 8) ops[total_ops + lane] = (uint16_t)DOM_IDENT;  // pad to a multiple of 8
      }
      __syncwarp();
      // serial walk over the block's ops (all lanes in lockstep), state after op i -> xs[i]
      const unsigned xo_in = xo;
      {
        // eight ops per iteration; the op codes of the next half are loaded
        // before the current lookups, so the only dependent latency per op is
        // one address add + one shared-memory load
        auto quad = [&](const uint2 cq, unsigned i) {
          if (((cq.x | cq.y) & (DOM_GENERIC * 0x00010001u)) == 0) {
            const unsigned o0 = tbl_s + ((cq.x & 0xFFFFu) << 12), o1 = tbl_s + ((cq.x >> 16) << 12);
            const unsigned o2 = tbl_s + ((cq.y & 0xFFFFu) << 12), o3 = tbl_s + ((cq.y >> 16) << 12);
            const unsigned x0 = lds_u16(o0 + xo);
            const unsigned x1 = lds_u16(o1 + x0);
            const unsigned x2 = lds_u16(o2 + x1);
            const unsigned x3 = lds_u16(o3 + x2);
            if (lane == 0) *reinterpret_cast<uint2 *>(xs + i) = make_uint2(x0 | (x1 << 16), x2 | (x3 << 16));
            xo = x3;
          } else {
#pragma unroll 1
            for (unsigned j = 0; j < 4; j++) {
              const unsigned op = ((j < 2 ? cq.x : cq.y) >> (16 * (j & 1))) & 0xFFFFu;
              if (op < DOM_GENERIC) xo = lds_u16(tbl_s + op * 4096u + xo);
              else xo = (step_from(T + (xo >> 1), op & 0xFFu) - T) * 2;
              if (lane == 0) xs[i + j] = (uint16_t)xo;
            }
          }
        };
        uint2 qa = *reinterpret_cast<const uint2 *>(ops);
        for (unsigned i = 0; i < total_ops; i += 8) {
          const uint2 qb = *reinterpret_cast<const uint2 *>(ops + i + 4);
          quad(qa, i);
          qa = *reinterpret_cast<const uint2 *>(ops + i + 8);  // may be past the block's ops: never used then
          quad(qb, i + 4);
        }
      }
      __syncwarp();
      if (have) {
        unsigned x = T + ((ofs ? (unsigned)xs[ofs - 1] : xo_in) >> 1);
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
        unsigned f[8];
#pragma unroll
        for (unsigned i = 0; i < 16; i++) {
          const unsigned sy = (ww[i >> 2] >> (8 * (i & 3))) & (A - 1);
          const int2 a = tt[sy];
          const unsigned nb = (x + (unsigned)a.y) >> 16;
          const unsigned fv = (nb << 12) | (x & ((1u << nb) - 1u));
          const unsigned xn = st[(int)(x >> nb) + a.x];
          if (i < valid) x = xn;
          if (i & 1) f[i >> 1] |= fv << 16; else f[i >> 1] = fv;
        }
        uint4 *fp = reinterpret_cast<uint4 *>(field + slot0 + (size_t)g * 16);
        fp[0] = make_uint4(f[0], f[1], f[2], f[3]);
        fp[1] = make_uint4(f[4], f[5], f[6], f[7]);
      }
      __syncwarp();
    }
  }
  if (lane == 0) fstate[(size_t)k * N + c] = (uint16_t)(T + (xo >> 1));
}

// ---------------------------------------------------------------------------
// pack
// ---------------------------------------------------------------------------
// Entry e of chunk k's stream: e < n_sym -> the symbol's field; then N state
// flushes for ctx 0..N-1 (FSE_flushCState: low `log` bits of the state,
// src/fse_common.hpp:87-88); then the end mark (BIT_closeCStream).
template <unsigned N>
__device__ __forceinline__ unsigned entry_value(unsigned e, unsigned n_sym, unsigned sym0, unsigned k,
                                               const uint32_t *__restrict__ perm, const uint16_t *__restrict__ field,
                                               const uint32_t *__restrict__ logs, const uint16_t *__restrict__ fstate) {
  if (e < n_sym) return field[perm[sym0 + e]];
  const unsigned c = e - n_sym;
  if (c < N) {
    const unsigned lg = logs[c];
    return (lg << 12) | ((unsigned)fstate[(size_t)k * N + c] & ((1u << lg) - 1u));
  }
  if (c == N) return (1u << 12) | 1u;
  return 0;
}

// Values of PACK_EPT consecutive stream entries starting at e0 (a multiple of PACK_EPT).
template <unsigned N>
__device__ __forceinline__ void entry_values(unsigned e0, unsigned n_sym, unsigned sym0, unsigned k,
                                             const uint32_t *__restrict__ perm, const uint16_t *__restrict__ field,
                                             const uint32_t *__restrict__ logs, const uint16_t *__restrict__ fstate,
                                             unsigned (&v)[PACK_EPT]) {
  static_assert(PACK_EPT == 8, "load8_u32");
  if (e0 + PACK_EPT <= n_sym) {  // all symbols: vector perm load, then the field gather
    unsigned slot[8];
    load8_u32(perm, (size_t)sym0 + e0, slot);
#pragma unroll
    for (unsigned i = 0; i < PACK_EPT; i++) v[i] = field[slot[i]];
  } else {
#pragma unroll
    for (unsigned i = 0; i < PACK_EPT; i++) v[i] = entry_value<N>(e0 + i, n_sym, sym0, k, perm, field, logs, fstate);
  }
}

template <unsigned N>
__global__ void __launch_bounds__(PACK_THREADS)
k_pack_count(const uint32_t *__restrict__ ptile0, const uint32_t *__restrict__ chunk_sym, unsigned n_chunks,
             const uint32_t *__restrict__ perm, const uint16_t *__restrict__ field, const uint32_t *__restrict__ logs,
             const uint16_t *__restrict__ fstate, uint32_t *__restrict__ pbits) {
  __shared__ unsigned wsum[PACK_THREADS / 32];
  const unsigned k = find_chunk(ptile0, n_chunks, blockIdx.x);
  const unsigned sym0 = chunk_sym[k], n_sym = chunk_sym[k + 1] - sym0;
  const unsigned e0 = (blockIdx.x - ptile0[k]) * PACK_TILE + threadIdx.x * PACK_EPT;
  unsigned bits = 0;
  unsigned ev[PACK_EPT];
  entry_values<N>(e0, n_sym, sym0, k, perm, field, logs, fstate, ev);
#pragma unroll
  for (unsigned i = 0; i < PACK_EPT; i++) bits += ev[i] >> 12;
  bits = __reduce_add_sync(0xffffffffu, bits);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = bits;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
#pragma unroll
    for (unsigned i = 0; i < PACK_THREADS / 32; i++) t += wsum[i];
    pbits[blockIdx.x] = t;
  }
}

// Per chunk: stream sizes from the bit prefix sums, 16-byte aligned arena
// offsets, and the rest of fq28_chunk_info.  Single CTA; chunks are few.
__global__ void __launch_bounds__(1024)
k_chunk_finish(unsigned n_chunks, const uint32_t *__restrict__ chunk_rec, const uint32_t *__restrict__ chunk_sym,
               const uint32_t *__restrict__ chunk_byte, const uint32_t *__restrict__ npos_off,
               const uint32_t *__restrict__ hdrscan, const uint32_t *__restrict__ ptile0_seq, const unsigned long long *__restrict__ pscan_seq,
               const uint32_t *__restrict__ ptile0_qual, const unsigned long long *__restrict__ pscan_qual,
               fq28_chunk_info *__restrict__ infos, uint64_t *__restrict__ scalars) {
  __shared__ unsigned long long carry[2];
  __shared__ unsigned long long wsum[2][33];
  if (threadIdx.x == 0) carry[0] = carry[1] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned base = 0; base < n_chunks; base += blockDim.x) {
    const unsigned k = base + threadIdx.x;
    unsigned long long len[2] = {0, 0}, al[2] = {0, 0}, inc[2];
    if (k < n_chunks) {
      const unsigned long long bs = pscan_seq[ptile0_seq[k + 1]] - pscan_seq[ptile0_seq[k]];
      const unsigned long long bq = pscan_qual[ptile0_qual[k + 1]] - pscan_qual[ptile0_qual[k]];
      len[0] = (bs + 7) >> 3;
      len[1] = (bq + 7) >> 3;
      al[0] = (len[0] + 15) & ~15ULL;
      al[1] = (len[1] + 15) & ~15ULL;
    }
#pragma unroll
    for (int w = 0; w < 2; w++) {
      unsigned long long v = al[w];
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        unsigned long long o = __shfl_up_sync(0xffffffffu, v, dd);
        if (lane >= (unsigned)dd) v += o;
      }
      inc[w] = v;
      if (lane == 31) wsum[w][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int w = 0; w < 2; w++) {
        unsigned long long x = wsum[w][lane], v = x;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
          unsigned long long o = __shfl_up_sync(0xffffffffu, v, dd);
          if (lane >= (unsigned)dd) v += o;
        }
        wsum[w][lane] = v - x;
        if (lane == 31) wsum[w][32] = v;
      }
    }
    __syncthreads();
    if (k < n_chunks) {
      fq28_chunk_info ci;
      ci.fastq_off = chunk_byte[k];
      ci.total = chunk_byte[k + 1] - chunk_byte[k];
      ci.n_records = chunk_rec[k + 1] - chunk_rec[k];
      ci.rec_off = chunk_rec[k];
      ci.seq_off = carry[0] + wsum[0][warp] + inc[0] - al[0];
      ci.qual_off = carry[1] + wsum[1][warp] + inc[1] - al[1];
      ci.seq_len = (uint32_t)len[0];
      ci.qual_len = (uint32_t)len[1];
      ci.n_pos_off = npos_off[chunk_rec[k]];
      ci.n_pos_len = npos_off[chunk_rec[k + 1]] - npos_off[chunk_rec[k]];
      ci.hdr_off = hdrscan[chunk_rec[k]];
      ci.hdr_bytes = hdrscan[chunk_rec[k + 1]] - hdrscan[chunk_rec[k]];
      infos[k] = ci;
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry[0] += wsum[0][32]; carry[1] += wsum[1][32]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { scalars[0] = carry[0]; scalars[1] = carry[1]; }
}

__global__ void k_zero_words(uint32_t *__restrict__ p, const uint64_t *__restrict__ n_bytes_ptr) {
  const size_t n = (size_t)((*n_bytes_ptr + 3) >> 2);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0;
}

// Bit packer.  CTA per pack tile; each thread assembles PACK_EPT consecutive
// fields in registers, a CTA-wide prefix sum of the bit counts places them,
// the tile's words are built in shared memory and written with 32-bit stores
// (atomicOr only on the two words shared with neighbouring tiles).
template <unsigned N>
__global__ void __launch_bounds__(PACK_THREADS)
k_pack_write(const uint32_t *__restrict__ ptile0, const uint32_t *__restrict__ chunk_sym, unsigned n_chunks,
             const uint32_t *__restrict__ perm, const uint16_t *__restrict__ field, const uint32_t *__restrict__ logs,
             const uint16_t *__restrict__ fstate, const unsigned long long *__restrict__ pscan,
             const fq28_chunk_info *__restrict__ infos, int which, uint8_t *__restrict__ arena) {
  constexpr unsigned SW = PACK_TILE * 12 / 32 + 4;
  __shared__ uint32_t sw[SW];
  __shared__ unsigned wsum[PACK_THREADS / 32 + 1];
  const unsigned k = find_chunk(ptile0, n_chunks, blockIdx.x);
  const unsigned sym0 = chunk_sym[k], n_sym = chunk_sym[k + 1] - sym0;
  const unsigned e0 = (blockIdx.x - ptile0[k]) * PACK_TILE + threadIdx.x * PACK_EPT;
  const unsigned long long tile_bit0 = pscan[blockIdx.x] - pscan[ptile0[k]];
  const unsigned tile_bits = (unsigned)(pscan[blockIdx.x + 1] - pscan[blockIdx.x]);
  const unsigned shift = (unsigned)(tile_bit0 & 31);
  for (unsigned i = threadIdx.x; i < SW; i += PACK_THREADS) sw[i] = 0;
  unsigned v[PACK_EPT];
  unsigned bits = 0;
  entry_values<N>(e0, n_sym, sym0, k, perm, field, logs, fstate, v);
#pragma unroll
  for (unsigned i = 0; i < PACK_EPT; i++) bits += v[i] >> 12;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = bits;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    unsigned o = __shfl_up_sync(0xffffffffu, inc, dd);
    if (lane >= (unsigned)dd) inc += o;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();  // also orders the sw[] zeroing
  unsigned woff = 0;
#pragma unroll
  for (unsigned i = 0; i < PACK_THREADS / 32; i++) woff += (i < warp) ? wsum[i] : 0u;
  const unsigned start = shift + woff + inc - bits;  // bit position inside sw[]
  if (bits) {
    unsigned long long lo = 0, hi = 0;
    unsigned p = start & 31;
#pragma unroll
    for (unsigned i = 0; i < PACK_EPT; i++) {
      const unsigned nb = v[i] >> 12;
      const unsigned long long val = v[i] & 0xFFFu;
      if (p < 64) {
        lo |= val << p;
        if (p + nb > 64) hi |= val >> (64 - p);
      } else {
        hi |= val << (p - 64);
      }
      p += nb;
    }
    const unsigned w0 = start >> 5;
    const unsigned a = (unsigned)lo, b = (unsigned)(lo >> 32), c = (unsigned)hi, dd = (unsigned)(hi >> 32);
    if (a) atomicOr(&sw[w0], a);
    if (b) atomicOr(&sw[w0 + 1], b);
    if (c) atomicOr(&sw[w0 + 2], c);
    if (dd) atomicOr(&sw[w0 + 3], dd);
  }
  __syncthreads();
  if (tile_bits == 0) return;
  const unsigned long long arena_off = which == 0 ? infos[k].seq_off : infos[k].qual_off;
  uint32_t *gw = reinterpret_cast<uint32_t *>(arena + arena_off) + (tile_bit0 >> 5);
  const unsigned n_words = (shift + tile_bits + 31) >> 5;
  for (unsigned i = threadIdx.x; i < n_words; i += PACK_THREADS) {
    const uint32_t w = sw[i];
    if (i == 0 || i == n_words - 1) { if (w) atomicOr(&gw[i], w); }
    else gw[i] = w;
  }
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
// Per stream type: tile tables + buffers (prep, main stream), then partition ->
// chain -> bit counts (run) on `strm`.  The sequence and quality pipelines are
// independent, so they run concurrently on the main and the side stream: the
// quality chain is latency-bound on a handful of SMs and leaves the rest of
// the machine to the sequence kernels.
struct KindBufs {
  DevBuf &tile0, &tbase, &ssym, &perm, &field, &fstate, &ptile0, &pbits, &pscan;
  unsigned n_tiles = 0, n_ptiles = 0, dom_cap = 0;
};

template <class K, unsigned TILE, unsigned STRIDE>
static int prep_kind(fq28_handle *h, KindBufs &b, size_t G) {
  constexpr unsigned N = K::n_models;
  const unsigned n_chunks = (unsigned)h->n_chunks;
  std::vector<uint32_t> tile0(n_chunks + 1), ptile0(n_chunks + 1);
  tile0[0] = ptile0[0] = 0;
  for (unsigned k = 0; k < n_chunks; k++) {
    const uint32_t ns = h->h_chunk_sym[k + 1] - h->h_chunk_sym[k];
    tile0[k + 1] = tile0[k] + (ns + TILE - 1) / TILE;
    ptile0[k + 1] = ptile0[k] + (ns + N + 1 + PACK_TILE - 1) / PACK_TILE;
  }
  b.n_tiles = tile0[n_chunks];
  b.n_ptiles = ptile0[n_chunks];
  FQ28_TRY(ensure(h, b.tile0, (n_chunks + 1) * 4));
  FQ28_TRY(ensure(h, b.ptile0, (n_chunks + 1) * 4));
  FQ28_CUDA(h, cudaMemcpyAsync(b.tile0.p, tile0.data(), (n_chunks + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(b.ptile0.p, ptile0.data(), (n_chunks + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));  // host vectors go out of scope
  FQ28_TRY(ensure(h, b.tbase, (size_t)(b.n_tiles + 1) * (N + 1) * 4));
  FQ28_TRY(ensure(h, b.ssym, (size_t)b.n_tiles * STRIDE + 64));
  FQ28_TRY(ensure(h, b.perm, (G + 16) * 4));
  FQ28_TRY(ensure(h, b.field, ((size_t)b.n_tiles * STRIDE + 64) * 2));
  FQ28_TRY(ensure(h, b.fstate, (size_t)n_chunks * N * 2 + 16));
  FQ28_TRY(ensure(h, b.pbits, (size_t)(b.n_ptiles + 1) * 4));
  FQ28_TRY(ensure(h, b.pscan, (size_t)(b.n_ptiles + 2) * 8));
  if (N == QUAL_N) {  // every (chunk, context) pair with >= DOM_MIN symbols fits
    uint32_t max_syms = 0;
    for (unsigned k = 0; k < n_chunks; k++) max_syms = std::max(max_syms, h->h_chunk_sym[k + 1] - h->h_chunk_sym[k]);
    b.dom_cap = n_chunks * (max_syms / DOM_MIN + 1);
    if (h->cfg.no_dom) b.dom_cap = 0;
    FQ28_TRY(ensure(h, h->dom_list, (size_t)b.dom_cap * 4 + 16));
    FQ28_TRY(ensure(h, h->present, (size_t)N * 8 + 64));
  }
  // scan scratch of the side stream must exist before the pipelines fork
  FQ28_TRY(ensure(h, h->scan_tmp, ((size_t)b.n_ptiles / 4096 + 8) * 8));
  FQ28_TRY(ensure(h, h->scan_tmp_side, ((size_t)b.n_ptiles / 4096 + 8) * 8));
  return FQ28_OK;
}

template <class K, unsigned A, unsigned TILE, unsigned STRIDE, unsigned RANK_WARPS>
static int run_kind(fq28_handle *h, const DevTables &tab, const typename K::key_t *key, KindBufs &b, bool side,
                    cudaEvent_t after_partition = nullptr) {
  constexpr unsigned N = K::n_models;
  cudaStream_t strm = side ? h->side : h->stream;
  const unsigned n_chunks = (unsigned)h->n_chunks;
  const uint32_t *chunk_sym = h->chunk_rec.as<uint32_t>() + h->chunk_stride;
  const unsigned n_tiles = b.n_tiles, n_ptiles = b.n_ptiles;

  int slot = stage_open(h, N == SEQ_N ? ST_PART_SEQ : ST_PART_QUAL, strm);
  if constexpr (N == SEQ_N) {
    if (n_tiles) {
    using P = Part8<SeqKind, SEQ_TILE, SEQ_STRIDE>;
    k_tile_part8<SeqKind, SEQ_TILE, SEQ_STRIDE><<<n_tiles, PART8_WARPS * 32, P::SMEM, strm>>>(
        reinterpret_cast<const uint16_t *>(key), b.tile0.as<uint32_t>(), chunk_sym, n_chunks, n_tiles,
        b.tbase.as<uint32_t>(), b.ssym.as<uint8_t>(), b.perm.as<uint32_t>());
    FQ28_LAUNCH_CHECK(h);
    }
  } else if (n_tiles) {
    // presence flags [N] u32 | map [N] u16 | inv [N] u16 | n_present
    uint32_t *present = h->present.as<uint32_t>();
    uint16_t *cmap = reinterpret_cast<uint16_t *>(present + N), *cinv = cmap + N;
    uint32_t *n_present = reinterpret_cast<uint32_t *>(cinv + N);
    FQ28_CUDA(h, cudaMemsetAsync(present, 0, N * sizeof(uint32_t), strm));
    k_tile_hist<K, TILE><<<n_tiles, 256, 0, strm>>>(key, b.tile0.as<uint32_t>(), chunk_sym, n_chunks,
                                                   b.tbase.as<uint32_t>(), present);
    FQ28_LAUNCH_CHECK(h);
    k_present_compact<N><<<1, 1024, 0, strm>>>(present, cmap, cinv, n_present);
    FQ28_LAUNCH_CHECK(h);
    k_tile_rank_compact<K, TILE, STRIDE><<<(n_tiles + RANKC_WARPS - 1) / RANKC_WARPS, RANKC_WARPS * 32, rankc_smem<K>(), strm>>>(
        key, b.tile0.as<uint32_t>(), chunk_sym, n_chunks, n_tiles, b.tbase.as<uint32_t>(), cmap, cinv, n_present,
        b.ssym.as<uint8_t>(), b.perm.as<uint32_t>());
    FQ28_LAUNCH_CHECK(h);
    k_tile_rank<K, TILE, STRIDE, RANK_WARPS><<<(n_tiles + RANK_WARPS - 1) / RANK_WARPS, RANK_WARPS * 32, 0, strm>>>(
        key, b.tile0.as<uint32_t>(), chunk_sym, n_chunks, n_tiles, b.tbase.as<uint32_t>(), b.ssym.as<uint8_t>(),
        b.perm.as<uint32_t>(), h->cfg.no_rankc ? nullptr : n_present);
    FQ28_LAUNCH_CHECK(h);
  }
  stage_close(h, slot, strm);
  if (after_partition) FQ28_CUDA(h, cudaEventRecord(after_partition, strm));

  slot = stage_open(h, N == SEQ_N ? ST_CHAIN_SEQ : ST_CHAIN_QUAL, strm);
  {
    const bool use_dom = (N == QUAL_N) && b.dom_cap > 0;
    unsigned long long *dom_count = reinterpret_cast<unsigned long long *>(h->d_scalars + 8);
    if (use_dom) {
      FQ28_CUDA(h, cudaMemsetAsync(dom_count, 0, sizeof(unsigned long long), strm));
      dim3 lgrid(N, (n_chunks + 127) / 128);
      k_dom_list<K><<<lgrid, 128, 0, strm>>>(b.tile0.as<uint32_t>(), n_chunks, b.tbase.as<uint32_t>(), tab.dom_sym,
                                            h->dom_list.as<uint32_t>(), b.dom_cap, dom_count);
      FQ28_LAUNCH_CHECK(h);
    }
    dim3 grid(N, (n_chunks + 31) / 32);
    k_chain<K, A, TILE, STRIDE><<<grid, 32, 0, strm>>>(b.ssym.as<uint8_t>(), b.tile0.as<uint32_t>(), n_chunks,
                                                      b.tbase.as<uint32_t>(), tab.logs, tab.toff, tab.ctab, tab.symtt,
                                                      b.field.as<uint16_t>(), b.fstate.as<uint16_t>(),
                                                      use_dom ? tab.dom_sym : nullptr);
    FQ28_LAUNCH_CHECK(h);
    if (use_dom) {
      k_chain_dom<K, A, STRIDE><<<b.dom_cap, DOM_WARPS * 32, DOM_SMEM, strm>>>(
          h->dom_list.as<uint32_t>(), dom_count, b.ssym.as<uint8_t>(), b.tile0.as<uint32_t>(), n_chunks,
          b.tbase.as<uint32_t>(), tab.logs, tab.toff, tab.ctab, tab.symtt, tab.norm, tab.dom_sym,
          b.field.as<uint16_t>(), b.fstate.as<uint16_t>());
      FQ28_LAUNCH_CHECK(h);
    }
  }
  stage_close(h, slot, strm);

  slot = stage_open(h, N == SEQ_N ? ST_PACK_SEQ : ST_PACK_QUAL, strm);
  k_pack_count<N><<<n_ptiles, PACK_THREADS, 0, strm>>>(b.ptile0.as<uint32_t>(), chunk_sym, n_chunks,
                                                      b.perm.as<uint32_t>(), b.field.as<uint16_t>(), tab.logs,
                                                      b.fstate.as<uint16_t>(), b.pbits.as<uint32_t>());
  FQ28_LAUNCH_CHECK(h);
  FQ28_TRY(scan_exclusive_u32_to_u64(h, b.pbits.as<uint32_t>(), b.pscan.as<uint64_t>(), n_ptiles, side));
  stage_close(h, slot, strm);
  return FQ28_OK;
}

template <unsigned N>
static int pack_kind(fq28_handle *h, const DevTables &tab, KindBufs &b, int which, DevBuf &arena, int scalar_idx,
                     bool side) {
  cudaStream_t strm = side ? h->side : h->stream;
  const unsigned n_chunks = (unsigned)h->n_chunks;
  const uint32_t *chunk_sym = h->chunk_rec.as<uint32_t>() + h->chunk_stride;
  const int slot = stage_open(h, N == SEQ_N ? ST_PACK_SEQ : ST_PACK_QUAL, strm);
  k_zero_words<<<148 * 4, 256, 0, strm>>>(arena.as<uint32_t>(), h->d_scalars + scalar_idx);
  FQ28_LAUNCH_CHECK(h);
  k_pack_write<N><<<b.n_ptiles, PACK_THREADS, 0, strm>>>(b.ptile0.as<uint32_t>(), chunk_sym, n_chunks,
                                                        b.perm.as<uint32_t>(), b.field.as<uint16_t>(), tab.logs,
                                                        b.fstate.as<uint16_t>(), b.pscan.as<unsigned long long>(),
                                                        h->d_infos.as<fq28_chunk_info>(), which, arena.as<uint8_t>());
  FQ28_LAUNCH_CHECK(h);
  stage_close(h, slot, strm);
  return FQ28_OK;
}

// Field separation of ALL parsed records, started by fq28_preparse(_dev) on the bulk stream: it
// needs the record table only, so it runs while the caller waits for its cut (several GPUs, one
// file: rank r waits r walks).  Buffers are sized by the bound "a symbol takes two bytes".
int extract_eager(fq28_handle *h) {
  if (h->cfg.no_eager || h->n_rec == 0) return FQ28_OK;
  if (!h->bulk) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    FQ28_CUDA(h, cudaStreamCreateWithPriority(&h->bulk, cudaStreamNonBlocking, lo));
    FQ28_CUDA(h, cudaEventCreateWithFlags(&h->ev_parsed, cudaEventDisableTiming));
    FQ28_CUDA(h, cudaEventCreateWithFlags(&h->ev_extract, cudaEventDisableTiming));
  }
  const size_t n_rec = h->n_rec, g_bound = h->n_bytes / 2 + 16;
  FQ28_TRY(ensure(h, h->key_seq, (g_bound + 8) * 2));
  FQ28_TRY(ensure(h, h->key_qual, (g_bound + 16) * 4));
  FQ28_TRY(ensure(h, h->n_count, (n_rec + 2) * 2));
  FQ28_CUDA(h, cudaEventRecord(h->ev_parsed, h->stream));
  FQ28_CUDA(h, cudaStreamWaitEvent(h->bulk, h->ev_parsed, 0));
  const unsigned blocks = (unsigned)((n_rec + EX_WARPS - 1) / EX_WARPS);
  k_extract<<<blocks, EX_WARPS * 32, 0, h->bulk>>>(h->d_fastq, h->seq_off.as<uint32_t>(), h->qual_off.as<uint32_t>(),
                                                  h->len.as<uint16_t>(), h->symoff.as<uint32_t>(), n_rec,
                                                  h->key_seq.as<uint16_t>(), h->key_qual.as<uint32_t>(),
                                                  h->n_count.as<uint16_t>(), h->d_status);
  FQ28_LAUNCH_CHECK(h);
  FQ28_CUDA(h, cudaEventRecord(h->ev_extract, h->bulk));
  h->extracted = true;
  return FQ28_OK;
}

int encode_slab(fq28_handle *h, fq28_chunk_info *infos, size_t infos_cap, fq28_enc_summary *summary) {
  if (!h->seq.ready || !h->qual.ready) return fail(h, FQ28_ERR_ARG, "frequency tables not built/loaded");
  const unsigned n_chunks = (unsigned)h->n_chunks;
  memset(&h->last_summary, 0, sizeof(h->last_summary));
  h->have_result = false;
  if (n_chunks > infos_cap) return fail(h, FQ28_ERR_CAP, "infos_cap %zu < %u chunks", infos_cap, n_chunks);
  if (n_chunks == 0) {
    if (summary) *summary = h->last_summary;
    h->have_result = true;
    return FQ28_OK;
  }
  const size_t n_rec = h->h_chunk_rec[n_chunks];   // records inside emitted chunks
  const size_t G = h->h_chunk_sym[n_chunks];       // symbols inside emitted chunks
  // chunk_rec buffer holds 3 arrays of (cap+1) u32; recover the stride
  const size_t stride = h->chunk_stride;
  const uint32_t *chunk_rec = h->chunk_rec.as<uint32_t>();
  const uint32_t *chunk_sym = chunk_rec + stride, *chunk_byte = chunk_rec + 2 * stride;

  const bool eager = h->extracted;   // fq28_preparse(_dev) has started k_extract over all parsed records
  h->extracted = false;
  if (!eager) {
    FQ28_TRY(ensure(h, h->key_seq, (G + 8) * 2));
    FQ28_TRY(ensure(h, h->key_qual, (G + 16) * 4));
    FQ28_TRY(ensure(h, h->n_count, (n_rec + 2) * 2));
  }
  FQ28_TRY(ensure(h, h->npos_off, (n_rec + 2) * 4));

  stage_begin(h, ST_EXTRACT);
  {
    const unsigned blocks = (unsigned)((n_rec + EX_WARPS - 1) / EX_WARPS);
    if (eager) {
      FQ28_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_extract, 0));
    } else {
      k_extract<<<blocks, EX_WARPS * 32, 0, h->stream>>>(h->d_fastq, h->seq_off.as<uint32_t>(), h->qual_off.as<uint32_t>(),
                                                        h->len.as<uint16_t>(), h->symoff.as<uint32_t>(), n_rec,
                                                        h->key_seq.as<uint16_t>(), h->key_qual.as<uint32_t>(),
                                                        h->n_count.as<uint16_t>(), h->d_status);
      FQ28_LAUNCH_CHECK(h);
    }
    FQ28_TRY(scan_exclusive_u16_to_u32(h, h->n_count.as<uint16_t>(), h->npos_off.as<uint32_t>(), n_rec));
    FQ28_TRY(ensure(h, h->hdrscan, (n_rec + 2) * 4));
    FQ28_TRY(scan_exclusive_u16_to_u32(h, h->hdr_len.as<uint16_t>(), h->hdrscan.as<uint32_t>(), n_rec));
    uint32_t *totals = reinterpret_cast<uint32_t *>(h->h_scalars + 58);   // (pinned)
    FQ28_CUDA(h, cudaMemcpyAsync(totals, h->npos_off.as<uint32_t>() + n_rec, 4, cudaMemcpyDeviceToHost, h->stream));
    FQ28_CUDA(h, cudaMemcpyAsync(totals + 1, h->hdrscan.as<uint32_t>() + n_rec, 4, cudaMemcpyDeviceToHost, h->stream));
    FQ28_TRY(check_status(h, "field separation"));
    const uint32_t total_n = totals[0], total_h = totals[1];
    FQ28_TRY(ensure(h, h->n_pos, ((size_t)total_n + 8) * 2));
    h->last_summary.n_pos_entries = total_n;
    h->last_summary.hdr_bytes = total_h;
    FQ28_TRY(ensure(h, h->hdr_arena, (size_t)total_h + 64));
    k_gather_headers<<<(unsigned)((n_rec + EX_WARPS * 2 - 1) / (EX_WARPS * 2)), EX_WARPS * 32, 0, h->stream>>>(h->d_fastq, h->hdr_off.as<uint32_t>(),
                                                             h->hdr_len.as<uint16_t>(), h->hdrscan.as<uint32_t>(), n_rec,
                                                             h->hdr_arena.as<uint8_t>());
    FQ28_LAUNCH_CHECK(h);
    if (total_n) {
      k_npos<<<blocks, EX_WARPS * 32, 0, h->stream>>>(h->d_fastq, h->seq_off.as<uint32_t>(), h->len.as<uint16_t>(),
                                                     h->n_count.as<uint16_t>(), h->npos_off.as<uint32_t>(), n_rec,
                                                     h->n_pos.as<uint16_t>());
      FQ28_LAUNCH_CHECK(h);
    }
  }
  stage_end(h, ST_EXTRACT);

  KindBufs bs{h->tile0_seq, h->tbase_seq, h->ssym_seq, h->perm_seq, h->out_seq, h->fstate_seq, h->ptile0_seq, h->pbits_seq, h->pscan_seq};
  KindBufs bq{h->tile0_qual, h->tbase_qual, h->ssym_qual, h->perm_qual, h->out_qual, h->fstate_qual, h->ptile0_qual, h->pbits_qual, h->pscan_qual};
  FQ28_TRY((prep_kind<SeqKind, SEQ_TILE, SEQ_STRIDE>(h, bs, G)));
  FQ28_TRY((prep_kind<QualKind, QUAL_TILE, QUAL_STRIDE>(h, bq, G)));
  FQ28_TRY(ensure(h, h->d_infos, (size_t)(n_chunks + 1) * sizeof(fq28_chunk_info)));
  const bool overlap = !h->cfg.serial;
  if (overlap) {
    FQ28_TRY(side_fork(h));
    // quality partition first with the whole GPU; its chain is latency-bound on a
    // few SMs, so the complete sequence pipeline runs underneath it
    FQ28_TRY((run_kind<QualKind, QUAL_A, QUAL_TILE, QUAL_STRIDE, 1>(h, h->qual, h->key_qual.as<uint32_t>(), bq, true, h->ev_join)));
    if (!h->cfg.full_overlap) FQ28_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    FQ28_TRY((run_kind<SeqKind, SEQ_A, SEQ_TILE, SEQ_STRIDE, 4>(h, h->seq, h->key_seq.as<uint16_t>(), bs, false)));
    FQ28_TRY(side_join(h));
  } else {
    FQ28_TRY((run_kind<QualKind, QUAL_A, QUAL_TILE, QUAL_STRIDE, 1>(h, h->qual, h->key_qual.as<uint32_t>(), bq, false)));
    FQ28_TRY((run_kind<SeqKind, SEQ_A, SEQ_TILE, SEQ_STRIDE, 4>(h, h->seq, h->key_seq.as<uint16_t>(), bs, false)));
  }

  k_chunk_finish<<<1, 1024, 0, h->stream>>>(n_chunks, chunk_rec, chunk_sym, chunk_byte, h->npos_off.as<uint32_t>(),
                                           h->hdrscan.as<uint32_t>(), h->ptile0_seq.as<uint32_t>(), h->pscan_seq.as<unsigned long long>(),
                                           h->ptile0_qual.as<uint32_t>(), h->pscan_qual.as<unsigned long long>(),
                                           h->d_infos.as<fq28_chunk_info>(), h->d_scalars);
  FQ28_LAUNCH_CHECK(h);
  FQ28_TRY(ensure_pinned(h, (size_t)n_chunks * sizeof(fq28_chunk_info)));
  FQ28_CUDA(h, cudaMemcpyAsync(h->h_scalars, h->d_scalars, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(h->h_pin, h->d_infos.p, (size_t)n_chunks * sizeof(fq28_chunk_info),
                               cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  const size_t seq_bytes = (size_t)h->h_scalars[0], qual_bytes = (size_t)h->h_scalars[1];
  FQ28_TRY(ensure(h, h->arena_seq, seq_bytes + 64));
  FQ28_TRY(ensure(h, h->arena_qual, qual_bytes + 64));
  FQ28_TRY(side_fork(h));
  FQ28_TRY((pack_kind<QUAL_N>(h, h->qual, bq, 1, h->arena_qual, 1, true)));
  FQ28_TRY((pack_kind<SEQ_N>(h, h->seq, bs, 0, h->arena_seq, 0, false)));
  FQ28_TRY(side_join(h));

  h->last_summary.n_chunks = n_chunks;
  h->last_summary.n_records = n_rec;
  h->last_summary.n_symbols = G;
  h->last_summary.seq_bytes = seq_bytes;
  h->last_summary.qual_bytes = qual_bytes;
  h->last_summary.consumed = h->h_chunk_byte[n_chunks];
  memcpy(infos, h->h_pin, (size_t)n_chunks * sizeof(fq28_chunk_info));
  if (summary) *summary = h->last_summary;
  h->have_result = true;
  return FQ28_OK;
}

// Dynamic shared memory opt-ins.  The attribute is per device, so this runs for every handle
// (fq28_create, after cudaSetDevice) instead of once per process.
int encode_init_device(fq28_handle *h) {
  using P = Part8<SeqKind, SEQ_TILE, SEQ_STRIDE>;
  FQ28_CUDA(h, cudaFuncSetAttribute(k_tile_part8<SeqKind, SEQ_TILE, SEQ_STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)P::SMEM));
  FQ28_CUDA(h, cudaFuncSetAttribute(k_tile_rank_compact<QualKind, QUAL_TILE, QUAL_STRIDE>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rankc_smem<QualKind>()));
  FQ28_CUDA(h, cudaFuncSetAttribute(k_chain_dom<QualKind, QUAL_A, QUAL_STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)DOM_SMEM));
  FQ28_CUDA(h, cudaFuncSetAttribute(k_chain_dom<SeqKind, SEQ_A, SEQ_STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)DOM_SMEM));
  return FQ28_OK;
}

}  // namespace fq28
