// fq28_tables.cu -- K3 context histograms and K4 FSE normalisation + CTable /
// DTable construction.
//
// K3 replaces FSE_Sequence::calculateFreqTable (src/fse_sequence.cpp:145-169)
// and FSE_Quality::calculateFreqTable (src/fse_quality.cpp:69-97); counts are
// produced WITHOUT the +1 prior so that per-GPU partial histograms can be
// summed with one NCCL allreduce; the prior is added in K4.
// K4 replaces makeNormalizedFreqTable (src/fse_common.hpp:179-200) and the
// FSE_Encoder / FSE_Decoder constructors (:46-71, :107-127).  The arithmetic is
// zstd's FSE_optimalTableLog / FSE_normalizeCount / FSE_buildCTable_wksp /
// FSE_buildDTable_wksp as specified in SURVEY.md Appendix A.
#include "fq28_internal.cuh"
#include "fq28_dec2.cuh"

namespace fq28 {

// ---------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------
constexpr int HIST_WARPS = 8;
constexpr unsigned HQ_SLOTS = 4096;          // per-CTA cache of hot quality (ctx, sym) counters
constexpr uint32_t HQ_EMPTY = 0xFFFFFFFFu;

// One warp per record, and inside the warp one CONTIGUOUS segment of the read per lane: a lane
// walks its symbols with the context in registers (no votes, no shuffles per symbol) and folds
// runs of equal (context, symbol) pairs into one update.  Round 1 spent ~4.5 warp instructions
// per symbol here (a vote loop per 32 symbols); this is ~0.3.
//   sequence: N is skipped and leaves the context unchanged (src/fse_sequence.cpp:157-158), so a
//             lane first finds the four non-N bases before its segment (the virtual prefix
//             T,C,C,T of SEQ_INITIAL_CTX beyond the read start).  Counters: 4 KB in shared memory.
//   quality : ctx = calcContext(q[i-1], q[i-2], q[i-3]) (src/fse_quality.h:40-44).  The 2 MB table
//             does not fit shared memory; a direct-mapped cache of 4096 (key, count) slots holds the
//             hot pairs (binned qualities: a few dozen pairs carry everything) and is flushed with
//             one RED per used slot; pairs that lose their slot go to L2 with a RED right away.
__global__ void __launch_bounds__(HIST_WARPS * 32)
k_hist(const char *__restrict__ d, const uint32_t *__restrict__ seq_off, const uint32_t *__restrict__ qual_off,
       const uint16_t *__restrict__ len, size_t n_rec, uint32_t *__restrict__ g_seq,
       uint32_t *__restrict__ g_qual, DevStatus *st) {
  __shared__ uint32_t s_seq[SEQ_N * SEQ_A];
  __shared__ uint32_t s_key[HQ_SLOTS], s_cnt[HQ_SLOTS];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned i = threadIdx.x; i < SEQ_N * SEQ_A; i += blockDim.x) s_seq[i] = 0;
  for (unsigned i = threadIdx.x; i < HQ_SLOTS; i += blockDim.x) { s_key[i] = HQ_EMPTY; s_cnt[i] = 0; }
  __syncthreads();
  auto add_qual = [&](uint32_t key, uint32_t n) {
    const unsigned slot = (key ^ (key >> 12)) & (HQ_SLOTS - 1);
    uint32_t cur = s_key[slot];
    if (cur == HQ_EMPTY) cur = atomicCAS(&s_key[slot], HQ_EMPTY, key), cur = cur == HQ_EMPTY ? key : cur;
    if (cur == key) atomicAdd(&s_cnt[slot], n);
    else atomicAdd(&g_qual[key], n);
  };
  const size_t warps_total = (size_t)gridDim.x * HIST_WARPS;
  for (size_t r = (size_t)blockIdx.x * HIST_WARPS + warp; r < n_rec; r += warps_total) {
    const unsigned L = len[r];
    const unsigned char *sp = reinterpret_cast<const unsigned char *>(d) + seq_off[r];
    const unsigned char *qp = reinterpret_cast<const unsigned char *>(d) + qual_off[r];
    unsigned K = (L + 31) / 32;          // symbols per lane
    if (K < 4) K = 4;
    const unsigned s = lane * K;
    if (s >= L) continue;
    const unsigned e = s + K < L ? s + K : L;
    // ---- sequence
    {
      // the four bases before position s, closest first: real non-N bases, then the prefix T,C,C,T
      unsigned ctx = 0, found = 0;
      for (unsigned j = s; j > 0 && found < 4; --j) {
        const int b = base2bits(sp[j - 1]);
        if (b >= 0) { ctx |= (unsigned)b << (6 - 2 * found); ++found; }
      }
      for (unsigned v = 0; found < 4; ++v, ++found) ctx |= ((SEQ_INITIAL_CTX >> (6 - 2 * v)) & 3u) << (6 - 2 * found);
      unsigned run_key = 0xFFFFFFFFu, run_n = 0;
      for (unsigned i = s; i < e; ++i) {
        const unsigned char c = sp[i];
        if (c == 'N') continue;
        const int sym = base2bits(c);
        if (sym < 0) { set_error(st, FQ28_ERR_ALPHABET, (unsigned)r); break; }
        const unsigned key = ctx * 4 + (unsigned)sym;
        if (key == run_key) ++run_n;
        else {
          if (run_n) atomicAdd(&s_seq[run_key], run_n);
          run_key = key;
          run_n = 1;
        }
        ctx = (ctx >> 2) + ((unsigned)sym << 6);
      }
      if (run_n) atomicAdd(&s_seq[run_key], run_n);
    }
    // ---- quality
    {
      unsigned q1 = s >= 1 ? (unsigned)qp[s - 1] - QUAL_OFFSET : 0u;
      unsigned q2 = s >= 2 ? (unsigned)qp[s - 2] - QUAL_OFFSET : 0u;
      unsigned q3 = s >= 3 ? (unsigned)qp[s - 3] - QUAL_OFFSET : 0u;
      unsigned run_key = 0xFFFFFFFFu, run_n = 0;
      for (unsigned i = s; i < e; ++i) {
        const unsigned q = (unsigned)qp[i] - QUAL_OFFSET;
        if (q > 63u) { set_error(st, FQ28_ERR_ALPHABET, (unsigned)r); break; }
        const unsigned key = qual_ctx(q1 & 63u, q2 & 63u, q3 & 63u) * QUAL_A + q;
        if (key == run_key) ++run_n;
        else {
          if (run_n) add_qual(run_key, run_n);
          run_key = key;
          run_n = 1;
        }
        q3 = q2; q2 = q1; q1 = q;
      }
      if (run_n) add_qual(run_key, run_n);
    }
  }
  __syncthreads();
  for (unsigned i = threadIdx.x; i < SEQ_N * SEQ_A; i += blockDim.x) {
    const uint32_t v = s_seq[i];
    if (v) atomicAdd(&g_seq[i], v);
  }
  for (unsigned i = threadIdx.x; i < HQ_SLOTS; i += blockDim.x) {
    const uint32_t v = s_cnt[i];
    if (v) atomicAdd(&g_qual[s_key[i]], v);
  }
}

int hist_slab(fq28_handle *h, uint32_t *d_seq_counts, uint32_t *d_qual_counts) {
  if (h->n_rec == 0) return FQ28_OK;
  const unsigned blocks = 148 * 4;
  k_hist<<<blocks, HIST_WARPS * 32, 0, h->stream>>>(h->d_fastq, h->seq_off.as<uint32_t>(), h->qual_off.as<uint32_t>(),
                                                    h->len.as<uint16_t>(), h->n_rec, d_seq_counts, d_qual_counts,
                                                    h->d_status);
  FQ28_LAUNCH_CHECK(h);
  return FQ28_OK;
}

// ---------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned hb32(unsigned v) { return 31u - (unsigned)__clz(v); }

// zstd FSE_optimalTableLog(maxTableLog = 0, srcSize, maxSV), Appendix A.1
__device__ unsigned fse_optimal_table_log(unsigned long long src_size, unsigned max_sv) {
  const unsigned max_bits_src = hb32((unsigned)(src_size - 1)) - 2u;  // U32 wrap allowed
  unsigned min_bits_src = hb32((unsigned)src_size) + 1u;
  const unsigned min_bits_sym = hb32(max_sv) + 2u;
  const unsigned min_bits = min_bits_src < min_bits_sym ? min_bits_src : min_bits_sym;
  unsigned t = FSE_DEFAULT_TABLELOG;
  if (max_bits_src < t) t = max_bits_src;
  if (min_bits > t) t = min_bits;
  if (t < FSE_MIN_TABLELOG) t = FSE_MIN_TABLELOG;
  if (t > FSE_MAX_TABLELOG) t = FSE_MAX_TABLELOG;
  return t;
}

// zstd FSE_normalizeM2, Appendix A.2
template <int A>
__device__ void fse_normalize_m2(short *norm, unsigned t, const unsigned *count, unsigned long long total) {
  const short NYA = -2;
  unsigned distributed = 0, to_distribute;
  const unsigned low_threshold = (unsigned)(total >> t);
  unsigned low_one = (unsigned)((total * 3) >> (t + 1));
  for (int s = 0; s < A; s++) {
    if (count[s] == 0) { norm[s] = 0; continue; }
    if (count[s] <= low_threshold) { norm[s] = -1; distributed++; total -= count[s]; continue; }
    if (count[s] <= low_one) { norm[s] = 1; distributed++; total -= count[s]; continue; }
    norm[s] = NYA;
  }
  to_distribute = (1u << t) - distributed;
  if (to_distribute == 0) return;
  if ((total / to_distribute) > low_one) {
    low_one = (unsigned)((total * 3) / (to_distribute * 2));
    for (int s = 0; s < A; s++)
      if (norm[s] == NYA && count[s] <= low_one) { norm[s] = 1; distributed++; total -= count[s]; }
    to_distribute = (1u << t) - distributed;
  }
  if (distributed == (unsigned)A) {
    unsigned max_v = 0, max_c = 0;
    for (int s = 0; s < A; s++)
      if (count[s] > max_c) { max_v = s; max_c = count[s]; }
    norm[max_v] = (short)(norm[max_v] + (short)to_distribute);
    return;
  }
  if (total == 0) {
    for (unsigned s = 0; to_distribute > 0; s = (s + 1) % A)
      if (norm[s] > 0) { to_distribute--; norm[s]++; }
    return;
  }
  const unsigned long long v_step_log = 62 - t;
  const unsigned long long mid = (1ULL << (v_step_log - 1)) - 1;
  const unsigned long long r_step = (((1ULL << v_step_log) * to_distribute) + mid) / (unsigned)total;
  unsigned long long tmp_total = mid;
  for (int s = 0; s < A; s++) {
    if (norm[s] == NYA) {
      const unsigned long long end = tmp_total + (count[s] * r_step);
      const unsigned s_start = (unsigned)(tmp_total >> v_step_log);
      const unsigned s_end = (unsigned)(end >> v_step_log);
      norm[s] = (short)(s_end - s_start);  // weight >= 1 whenever every count >= 1
      tmp_total = end;
    }
  }
}

// zstd FSE_normalizeCount(norm, t, count, total, maxSV, useLowProbCount = 1)
template <int A>
__device__ void fse_normalize(short *norm, unsigned t, const unsigned *count, unsigned long long total) {
  const unsigned rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
  const unsigned long long scale = 62 - t;
  const unsigned long long step = (1ULL << 62) / (unsigned)total;
  const unsigned long long v_step = 1ULL << (scale - 20);
  int still = 1 << t;
  unsigned largest = 0;
  short largest_p = 0;
  const unsigned low_threshold = (unsigned)(total >> t);
  for (int s = 0; s < A; s++) {
    // count[s] == total (rle) cannot happen: every count >= 1 and A >= 2
    if (count[s] == 0) { norm[s] = 0; continue; }
    if (count[s] <= low_threshold) {
      norm[s] = -1;
      still--;
    } else {
      short proba = (short)((count[s] * step) >> scale);
      if (proba < 8) {
        const unsigned long long rest_to_beat = v_step * rtb[proba];
        proba = (short)(proba + (((count[s] * step) - ((unsigned long long)proba << scale)) > rest_to_beat ? 1 : 0));
      }
      if (proba > largest_p) { largest_p = proba; largest = s; }
      norm[s] = proba;
      still -= proba;
    }
  }
  if (-still >= (norm[largest] >> 1)) fse_normalize_m2<A>(norm, t, count, total);
  else norm[largest] = (short)(norm[largest] + (short)still);
}

// One thread per context: +1 prior (src/fse_sequence.cpp:149-150,
// src/fse_quality.cpp:75-76), table log, normalised counts.
template <int A>
__global__ void k_normalize(const uint32_t *__restrict__ counts, unsigned n_models, short *__restrict__ norm_out,
                            uint32_t *__restrict__ logs, uint32_t *__restrict__ max_log) {
  const unsigned ctx = blockIdx.x * blockDim.x + threadIdx.x;
  if (ctx >= n_models) return;
  unsigned cnt[A];
  short norm[A];
  unsigned long long total = 0;
  for (int s = 0; s < A; s++) {
    cnt[s] = counts[(size_t)ctx * A + s] + 1u;
    total += cnt[s];
    norm[s] = 0;
  }
  const unsigned t = fse_optimal_table_log(total, A - 1);
  fse_normalize<A>(norm, t, cnt, total);
  for (int s = 0; s < A; s++) norm_out[(size_t)ctx * A + s] = norm[s];
  logs[ctx] = t;
  atomicMax(max_log, t);
}

// toff[ctx] = first table cell of ctx (exclusive scan of 1<<log); single CTA
__global__ void __launch_bounds__(1024)
k_table_offsets(const uint32_t *__restrict__ logs, unsigned n_models, uint32_t *__restrict__ toff) {
  __shared__ unsigned wsum[33];
  __shared__ unsigned carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned base = 0; base < n_models; base += blockDim.x) {
    const unsigned i = base + threadIdx.x;
    const unsigned v = i < n_models ? (1u << logs[i]) : 0u;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned w = wsum[lane], winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        unsigned o = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= (unsigned)d) winc += o;
      }
      wsum[lane] = winc - w;
      if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    const unsigned carry = carry_s;
    if (i < n_models) toff[i] = carry + wsum[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + wsum[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) toff[n_models] = carry_s;
}

// One thread per context: symbol spread (A.3), CTable (A.4), DTable (A.6).
// The DTable cells double as the spread scratch (cell[u] = symbol first).
template <int A>
__global__ void k_build_tables(const short *__restrict__ norm_in, const uint32_t *__restrict__ logs,
                               const uint32_t *__restrict__ toff, unsigned n_models,
                               uint16_t *__restrict__ ctab, int2 *__restrict__ symtt,
                               uint32_t *__restrict__ dtab, uint32_t *__restrict__ dtab_fix, int8_t *__restrict__ dom_sym) {
  const unsigned ctx = blockIdx.x * blockDim.x + threadIdx.x;
  if (ctx >= n_models) return;
  short norm[A];
  unsigned cumul[A + 1];
  unsigned next[A];
  for (int s = 0; s < A; s++) norm[s] = norm_in[(size_t)ctx * A + s];
  const unsigned t = logs[ctx], T = 1u << t, mask = T - 1;
  const unsigned step = (T >> 1) + (T >> 3) + 3;
  uint32_t *cells = dtab + toff[ctx];
  uint16_t *st = ctab + toff[ctx];
  unsigned high = T - 1;
  for (int s = 0; s < A; s++)
    if (norm[s] == -1) cells[high--] = (unsigned)s;
  unsigned pos = 0;
  for (int s = 0; s < A; s++) {
    for (int i = 0; i < norm[s]; i++) {
      cells[pos] = (unsigned)s;
      pos = (pos + step) & mask;
      while (pos > high) pos = (pos + step) & mask;
    }
  }
  cumul[0] = 0;
  for (int s = 0; s < A; s++) {
    const unsigned w = norm[s] == -1 ? 1u : (unsigned)norm[s];
    cumul[s + 1] = cumul[s] + w;
    next[s] = w;
  }
  // symbol transforms
  unsigned total = 0;
  for (int s = 0; s < A; s++) {
    int2 tt;
    if (norm[s] == 0) {
      tt.x = 0;
      tt.y = (int)(((t + 1) << 16) - T);
    } else if (norm[s] == -1 || norm[s] == 1) {
      tt.y = (int)((t << 16) - T);
      tt.x = (int)(total - 1);
      total++;
    } else {
      const unsigned max_bits_out = t - hb32((unsigned)norm[s] - 1);
      const unsigned min_state_plus = (unsigned)norm[s] << max_bits_out;
      tt.y = (int)((max_bits_out << 16) - min_state_plus);
      tt.x = (int)(total - (unsigned)norm[s]);
      total += (unsigned)norm[s];
    }
    symtt[(size_t)ctx * A + s] = tt;
  }
  {  // dominant symbol: norm > T/2 <=> every step with it emits 0 or 1 bit (maxBitsOut == 1)
    int d = -1;
    for (int s = 0; s < A; s++)
      if (norm[s] > (int)(T >> 1)) d = s;
    dom_sym[ctx] = (int8_t)d;
  }
  // next-state table + decode cells, ascending u
  for (unsigned u = 0; u < T; u++) {
    const unsigned s = cells[u];
    st[cumul[s]++] = (uint16_t)(T + u);
    const unsigned x = next[s]++;
    const unsigned nb = t - hb32(x);
    const unsigned ns = (x << nb) - T;
    const unsigned cell = (ns & 0xFFFFu) | (s << 16) | (nb << 24);
    cells[u] = cell;
    if (t <= FIX_LOG) dtab_fix[((size_t)ctx << FIX_LOG) + u] = cell;
  }
}

// Sequence tables (4 symbols): one WARP per context.  Without low-probability
// symbols (norm == -1) the spread visits cell (v * step) & mask at visit v and
// symbol s owns visits [cumul[s], cumul[s+1]), so the cells are filled in
// parallel; the ascending-u pass ranks the 32 cells of a step with one vote
// per symbol.  Contexts with a -1 entry take the serial path of
// k_build_tables (lane 0).  Also fills the compressed decoder tables
// (SeqDecTables), replacing k_build_seqdec.
__global__ void __launch_bounds__(128)
k_build_tables_seq(const short *__restrict__ norm_in, const uint32_t *__restrict__ logs, const uint32_t *__restrict__ toff,
                   uint16_t *__restrict__ ctab, int2 *__restrict__ symtt, uint32_t *__restrict__ dtab,
                   uint32_t *__restrict__ dtab_fix, int8_t *__restrict__ dom_sym, SeqDecTables *__restrict__ dec) {
  constexpr int A = SEQ_A;
  const unsigned lane = threadIdx.x & 31;
  const unsigned ctx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ctx >= SEQ_N) return;
  short norm[A];
#pragma unroll
  for (int s = 0; s < A; s++) norm[s] = norm_in[(size_t)ctx * A + s];
  const unsigned t = logs[ctx], T = 1u << t, mask = T - 1;
  const unsigned step = (T >> 1) + (T >> 3) + 3;
  uint32_t *cells = dtab + toff[ctx];
  uint16_t *st = ctab + toff[ctx];
  bool lowprob = false;
  unsigned cumul[A + 1];
  cumul[0] = 0;
#pragma unroll
  for (int s = 0; s < A; s++) {
    lowprob |= norm[s] == -1;
    cumul[s + 1] = cumul[s] + (norm[s] == -1 ? 1u : (unsigned)norm[s]);
  }
  // ---- spread (A.3): cells[u] = symbol
  if (!lowprob) {
    for (unsigned v = lane; v < T; v += 32) {
      const unsigned s = (v >= cumul[1]) + (v >= cumul[2]) + (v >= cumul[3]);
      cells[(v * step) & mask] = s;
    }
  } else if (lane == 0) {
    unsigned high = T - 1;
    for (int s = 0; s < A; s++)
      if (norm[s] == -1) cells[high--] = (unsigned)s;
    unsigned pos = 0;
    for (int s = 0; s < A; s++)
      for (int i = 0; i < norm[s]; i++) {
        cells[pos] = (unsigned)s;
        pos = (pos + step) & mask;
        while (pos > high) pos = (pos + step) & mask;
      }
  }
  // ---- symbol transforms (A.4), dominant symbol
  if (lane == 0) {
    unsigned total = 0;
    int d = -1;
    for (int s = 0; s < A; s++) {
      int2 tt;
      if (norm[s] == 0) {
        tt.x = 0;
        tt.y = (int)(((t + 1) << 16) - T);
      } else if (norm[s] == -1 || norm[s] == 1) {
        tt.y = (int)((t << 16) - T);
        tt.x = (int)(total - 1);
        total++;
      } else {
        const unsigned max_bits_out = t - hb32((unsigned)norm[s] - 1);
        const unsigned min_state_plus = (unsigned)norm[s] << max_bits_out;
        tt.y = (int)((max_bits_out << 16) - min_state_plus);
        tt.x = (int)(total - (unsigned)norm[s]);
        total += (unsigned)norm[s];
      }
      symtt[(size_t)ctx * A + s] = tt;
      if (norm[s] > (int)(T >> 1)) d = s;
      dec->snext[ctx][s] = (uint16_t)((norm[s] == -1 ? 1 : norm[s]) | (s == 0 ? (t << 12) : 0u));
    }
    dom_sym[ctx] = (int8_t)d;
  }
  __syncwarp();
  // ---- next-state table + decode cells + compressed decoder tables, ascending u, 32 cells a step
  unsigned run[A];  // cells of symbol s below the current step (uniform across the warp)
#pragma unroll
  for (int s = 0; s < A; s++) run[s] = 0;
  unsigned coarse_base[A] = {0, 0, 0, 0};
  const unsigned lt = (1u << lane) - 1u;
  for (unsigned u0 = 0; u0 < (1u << FIX_LOG); u0 += 32) {
    const unsigned u = u0 + lane;
    const bool in = u < T;
    const unsigned s = in ? cells[u] & 3u : 0u;
    unsigned votes[A];
#pragma unroll
    for (int q = 0; q < A; q++) votes[q] = __ballot_sync(0xffffffffu, in && s == (unsigned)q);
    if ((u0 & 255) == 0) {
#pragma unroll
      for (int q = 0; q < A; q++) {
        coarse_base[q] = run[q];
        if (lane == 0) dec->coarse[ctx][u0 >> 8][q] = (uint16_t)run[q];
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < A; q++) dec->fine[ctx][u0 >> 5][q] = (uint8_t)(run[q] - coarse_base[q]);
    }
    // 16 two-bit symbols per word: lanes 0 and 16 assemble the two words of the step
    {
      unsigned w = s << (2 * (lane & 15));
#pragma unroll
      for (int d2 = 8; d2 >= 1; d2 >>= 1) w |= __shfl_xor_sync(0xffffffffu, w, d2);
      if ((lane & 15) == 0) dec->symtab[ctx][(u0 >> 4) + (lane >> 4)] = w;
    }
    if (in) {
      const unsigned myrun = s == 0 ? run[0] : s == 1 ? run[1] : s == 2 ? run[2] : run[3];
      const unsigned myvote = s == 0 ? votes[0] : s == 1 ? votes[1] : s == 2 ? votes[2] : votes[3];
      const unsigned r = myrun + (unsigned)__popc(myvote & lt);  // rank of the cell among the symbol's cells
      const unsigned cs = s == 0 ? cumul[0] : s == 1 ? cumul[1] : s == 2 ? cumul[2] : cumul[3];
      const unsigned w_ = s == 0 ? (norm[0] == -1 ? 1u : (unsigned)norm[0]) : s == 1 ? (norm[1] == -1 ? 1u : (unsigned)norm[1])
                        : s == 2 ? (norm[2] == -1 ? 1u : (unsigned)norm[2]) : (norm[3] == -1 ? 1u : (unsigned)norm[3]);
      st[cs + r] = (uint16_t)(T + u);
      const unsigned x = w_ + r;
      const unsigned nb = t - hb32(x);
      const unsigned nst = (x << nb) - T;
      const unsigned cell = (nst & 0xFFFFu) | (s << 16) | (nb << 24);
      cells[u] = cell;
      if (t <= FIX_LOG) dtab_fix[((size_t)ctx << FIX_LOG) + u] = cell;
    }
#pragma unroll
    for (int q = 0; q < A; q++) run[q] += (unsigned)__popc(votes[q]);
  }
}

// ---- decoder-side structures ----------------------------------------------
// logsuf[c] = sum of logs[c'] for c' > c: FSE_Decoder::startChunk reads the
// initial states for ctx N-1 .. 0 (src/fse_common.hpp:134-138), so the state
// of context c starts logsuf[c] bits below the end mark.  Single thread.
__global__ void __launch_bounds__(1024)
k_logsuf(const uint32_t *__restrict__ logs, unsigned n_models, uint32_t *__restrict__ logsuf) {
  // single CTA: thread i owns a block of consecutive contexts; suffix sums via
  // an inclusive prefix scan of the block totals
  __shared__ unsigned wsum[33];
  const unsigned per = (n_models + 1023) / 1024;
  const unsigned c0 = threadIdx.x * per, c1 = min(c0 + per, n_models);
  unsigned s = 0;
  for (unsigned c = c0; c < c1; c++) s += logs[c];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (unsigned)d) inc += o;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned v = wsum[lane], w = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= (unsigned)d) w += o;
    }
    wsum[lane] = w - v;
    if (lane == 31) wsum[32] = w;
  }
  __syncthreads();
  const unsigned total = wsum[32];
  unsigned before = wsum[warp] + inc - s;  // sum of logs[c] for c < c0
  for (unsigned c = c0; c < c1; c++) {
    before += logs[c];
    logsuf[c] = total - before;             // sum of logs[c'] for c' > c
  }
  if (threadIdx.x == 0) logsuf[n_models] = total;  // total bits of the state block
}

// Quality: compact ids for the contexts whose table is not the untouched
// pattern (64 symbols with norm 2 at log 7 = only the +1 prior was seen).
// Single CTA: flag per context, then an in-order prefix.
__global__ void __launch_bounds__(1024)
k_qual_cid(const short *__restrict__ norm, const uint32_t *__restrict__ logs, uint16_t *__restrict__ cid,
           uint32_t *__restrict__ n_touched) {
  __shared__ unsigned flags[QUAL_N];
  for (unsigned c = threadIdx.x; c < QUAL_N; c += blockDim.x) {
    bool touched = logs[c] != 7;
    for (int s = 0; s < (int)QUAL_A && !touched; s++) touched = norm[(size_t)c * QUAL_A + s] != 2;
    flags[c] = touched ? 1u : 0u;
  }
  __syncthreads();
  // in-order prefix: thread i owns QUAL_N / 1024 consecutive contexts
  __shared__ unsigned wsum[33];
  constexpr unsigned PER = QUAL_N / 1024;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned s = 0;
#pragma unroll
  for (unsigned i = 0; i < PER; i++) s += flags[threadIdx.x * PER + i];
  unsigned inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (unsigned)d) inc += o;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned v = wsum[lane], w = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= (unsigned)d) w += o;
    }
    wsum[lane] = w - v;
    if (lane == 31) wsum[32] = w;
  }
  __syncthreads();
  unsigned n = wsum[warp] + inc - s;
#pragma unroll
  for (unsigned i = 0; i < PER; i++) {
    const unsigned c = threadIdx.x * PER + i;
    cid[c] = flags[c] ? (uint16_t)(n++) : (uint16_t)0xFFFF;
  }
  if (threadIdx.x == 0) *n_touched = wsum[32];
}

// Zero-bit run tables (see QZ_MAX in fq28_internal.cuh).  Single CTA.
__global__ void __launch_bounds__(1024)
k_qual_zrun(const uint32_t *__restrict__ logs, const uint32_t *__restrict__ dtab_fix, const int8_t *__restrict__ dom_sym,
            uint16_t *__restrict__ cid, uint16_t *__restrict__ zrun, uint32_t *__restrict__ zinfo) {
  __shared__ unsigned slot_sym[QZ_MAX];
  __shared__ unsigned n_slots;
  if (threadIdx.x == 0) {
    unsigned n = 0;
    for (int d = (int)QUAL_A - 1; d >= 0 && n < QZ_MAX; --d) {  // high qualities first
      const unsigned cx = qual_ctx((unsigned)d, (unsigned)d, (unsigned)d);
      if (dom_sym[cx] == d && cid[cx] != 0xFFFFu) {
        cid[cx] = (uint16_t)(cid[cx] | ((n + 1) << 13));  // run slot + 1 above the 13-bit compact id
        slot_sym[n++] = (unsigned)d;
      }
    }
    n_slots = n;
    zinfo[0] = n;
    for (unsigned j = 0; j < QZ_MAX; j++)
      zinfo[1 + j] = j < n ? qual_ctx(slot_sym[j], slot_sym[j], slot_sym[j]) : 0xFFFFFFFFu;
  }
  __syncthreads();
  for (unsigned j = 0; j < n_slots; j++) {
    const unsigned d = slot_sym[j], cx = qual_ctx(d, d, d);
    const unsigned T = 1u << logs[cx];
    uint16_t *Z = zrun + (size_t)j * 2 * (1u << FIX_LOG), *J1 = Z + (1u << FIX_LOG);
    for (unsigned x = threadIdx.x; x < (1u << FIX_LOG); x += blockDim.x) {
      unsigned k = 0, y = x, j1 = 0xFFFFu;
      while (x < T && k < QZ_CAP) {
        const unsigned e = dtab_fix[((size_t)cx << FIX_LOG) + y];
        if (((e >> 16) & 63u) != d || (e >> 24) != 0) break;
        y = e & 0xFFFFu;
        if (k == 0) j1 = y;
        k++;
      }
      Z[x] = (uint16_t)((k << 11) | y);
      J1[x] = (uint16_t)j1;
    }
  }
}


// ---- decoder v2 tables (fq28_dec2.cuh) ---------------------------------------
// Sequence W table: the DTable cells repacked (char, symbol, nbBits, base, inline flag).
__global__ void k_seq_wtab(const uint32_t *__restrict__ logs, const uint32_t *__restrict__ dtab_fix,
                           uint32_t *__restrict__ wtab) {
  const unsigned ctx = blockIdx.x;
  const unsigned T = 1u << logs[ctx];
  for (unsigned u = threadIdx.x; u < (1u << FIX_LOG); u += blockDim.x) {
    const size_t i = ((size_t)ctx << FIX_LOG) + u;
    wtab[i] = u < T ? dec2::make_w_seq(dtab_fix[i], ctx) : 0u;
  }
}

// Quality: dense alphabet V = {0} + every value that is the last symbol of a touched
// context (cid != 0xFFFF).  rk[q] = rank of q in V (0xFF outside), vq[rank] = q.  Single CTA.
__global__ void __launch_bounds__(64)
k_qual_dense(const uint16_t *__restrict__ cid, uint8_t *__restrict__ qrk, uint32_t *__restrict__ qdinfo) {
  __shared__ unsigned inv[64];
  const unsigned q = threadIdx.x;
  unsigned seen = q == 0;
  for (unsigned c = q; c < QUAL_N && !seen; c += 64) seen = cid[c] != 0xFFFFu;  // contexts with (c & 63) == q
  inv[q] = seen;
  __syncthreads();
  unsigned rank = 0, nv = 0;
  for (unsigned j = 0; j < 64; j++) {
    if (j < q) rank += inv[j];
    nv += inv[j];
  }
  qrk[q] = seen ? (uint8_t)rank : (uint8_t)0xFF;
  qrk[64 + q] = 0;
  __syncthreads();
  if (seen) qrk[64 + rank] = (uint8_t)q;
  if (q == 0) qdinfo[0] = nv;
}

// Windowed layout of the dense contexts: for every row (rank(max) * 2 + eq) the columns [lo, hi]
// that span its touched contexts (cid != 0xFFFF); a row takes width + 1 entries (the last one is
// the out-of-window word).  Row 1, column 0 is the record-start context and is always kept.
// Single CTA of 128 threads, one per row.
__global__ void __launch_bounds__(128)
k_qual_win(const uint16_t *__restrict__ cid, const uint8_t *__restrict__ qrk, uint32_t *__restrict__ qdinfo,
           uint2 *__restrict__ qwin) {
  __shared__ unsigned wd[128], st[128];
  const unsigned nv = qdinfo[0];
  const unsigned row = threadIdx.x, rm = row >> 1, eq = row & 1u;
  unsigned lo = 64, hi = 0;
  bool any = false;
  if (rm < nv) {
    for (unsigned rq = 0; rq < nv; rq++) {
      const unsigned cx = ((unsigned)qrk[64 + rm] << 6) + qrk[64 + rq] + (eq << 12);
      if (cid[cx] != 0xFFFFu) {
        lo = rq < lo ? rq : lo;
        hi = rq > hi ? rq : hi;
        any = true;
      }
    }
    if (row == 1) { lo = 0; any = true; }
  }
  const unsigned width = any ? hi - lo + 1 : 0;
  wd[row] = row < 2 * nv ? width + 1 : 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    for (unsigned r = 0; r < 128; r++) { st[r] = acc; acc += wd[r]; }
    qdinfo[1] = acc;
    qdinfo[2] = 2 * nv;
  }
  __syncthreads();
  qwin[row] = make_uint2(st[row] * 4, (any ? lo * 4 : 0u) | ((width * 4) << 16));
}

// W table of the windowed layout: row = compact id of the (row, column) entry.
__global__ void k_qual_wtabw(const uint32_t *__restrict__ logs, const uint32_t *__restrict__ dtab_fix,
                             const uint8_t *__restrict__ qrk, const uint32_t *__restrict__ qdinfo,
                             const uint2 *__restrict__ qwin, uint32_t *__restrict__ wtabw) {
  __shared__ uint8_t rk[64];
  const unsigned nv = qdinfo[0];
  const unsigned rq = blockIdx.x & 63u, eq = (blockIdx.x >> 6) & 1u, rm = blockIdx.x >> 7;
  if (rq >= nv || rm >= nv) return;
  const uint2 w = qwin[rm * 2 + eq];
  const unsigned rel = rq * 4 - (w.y & 0xFFFFu);
  if (rel >= (w.y >> 16)) return;
  if (threadIdx.x < 64) rk[threadIdx.x] = qrk[threadIdx.x];
  __syncthreads();
  const unsigned cx = ((unsigned)qrk[64 + rm] << 6) + qrk[64 + rq] + (eq << 12);
  const unsigned d = (w.x + rel) >> 2;
  const unsigned T = 1u << logs[cx];
  for (unsigned u = threadIdx.x; u < (1u << FIX_LOG); u += blockDim.x)
    wtabw[((size_t)d << FIX_LOG) + u] = u < T ? dec2::make_w_qual(dtab_fix[((size_t)cx << FIX_LOG) + u], rk) : 0u;
}

// Quality W table: one CTA per dense context (row = rank(max) * 2 + eq, column = rank(q)).
__global__ void k_qual_wtab(const uint32_t *__restrict__ logs, const uint32_t *__restrict__ dtab_fix,
                            const uint8_t *__restrict__ qrk, const uint32_t *__restrict__ qdinfo,
                            uint32_t *__restrict__ wtab) {
  __shared__ uint8_t rk[64];
  const unsigned nv = qdinfo[0];
  const unsigned rq = blockIdx.x & 63u, eq = (blockIdx.x >> 6) & 1u, rm = blockIdx.x >> 7;
  if (rq >= nv || rm >= nv) return;
  if (threadIdx.x < 64) rk[threadIdx.x] = qrk[threadIdx.x];
  __syncthreads();
  const unsigned cx = ((unsigned)qrk[64 + rm] << 6) + qrk[64 + rq] + (eq << 12);
  const unsigned d = dec2::qual_dense_id(rm, eq, rq);
  const unsigned T = 1u << logs[cx];
  for (unsigned u = threadIdx.x; u < (1u << FIX_LOG); u += blockDim.x)
    wtab[((size_t)d << FIX_LOG) + u] = u < T ? dec2::make_w_qual(dtab_fix[((size_t)cx << FIX_LOG) + u], rk) : 0u;
}

int tables_alloc(fq28_handle *h, DevTables &t, unsigned n_models, unsigned alphabet) {
  if (t.norm) return FQ28_OK;
  t.n_models = n_models;
  t.alphabet = alphabet;
  const size_t na = (size_t)n_models * alphabet;
  FQ28_CUDA(h, cudaMalloc(&t.counts, na * sizeof(uint32_t)));
  FQ28_CUDA(h, cudaMalloc(&t.norm, na * sizeof(int16_t)));
  FQ28_CUDA(h, cudaMalloc(&t.logs, n_models * sizeof(uint32_t)));
  FQ28_CUDA(h, cudaMalloc(&t.max_log, sizeof(uint32_t)));
  FQ28_CUDA(h, cudaMalloc(&t.toff, (n_models + 1) * sizeof(uint32_t)));
  FQ28_CUDA(h, cudaMalloc(&t.symtt, na * sizeof(int2)));
  FQ28_CUDA(h, cudaMalloc(&t.dom_sym, n_models));
  // every table log is <= 11 here (FSE_DEFAULT_TABLELOG with maxTableLog = 0
  // and minBits <= 7), but size for FSE_MAX_TABLELOG to be safe
  t.cells_cap = (size_t)n_models << FSE_MAX_TABLELOG;
  FQ28_CUDA(h, cudaMalloc(&t.ctab, t.cells_cap * sizeof(uint16_t)));
  FQ28_CUDA(h, cudaMalloc(&t.dtab, t.cells_cap * sizeof(uint32_t)));
  FQ28_CUDA(h, cudaMalloc(&t.dtab_fix, ((size_t)n_models << FIX_LOG) * sizeof(uint32_t)));
  FQ28_CUDA(h, cudaMalloc(&t.logsuf, (n_models + 1) * sizeof(uint32_t)));
  // W tables of the v2 decoder: all contexts (sequence) / the largest dense set (quality: 2 * 64 * 64 rows)
  FQ28_CUDA(h, cudaMalloc(&t.wtab, ((size_t)n_models << FIX_LOG) * sizeof(uint32_t)));
  if (alphabet == SEQ_A) {
    FQ28_CUDA(h, cudaMalloc(&t.seqdec, sizeof(SeqDecTables)));
  } else {
    FQ28_CUDA(h, cudaMalloc(&t.qrk, 128));
    FQ28_CUDA(h, cudaMalloc(&t.qdinfo, 4 * sizeof(uint32_t)));
    FQ28_CUDA(h, cudaMalloc(&t.qwin, 128 * sizeof(uint2)));
    FQ28_CUDA(h, cudaMalloc(&t.wtabw, ((size_t)QWIN_MAX_ENTRIES << FIX_LOG) * sizeof(uint32_t)));
    FQ28_CUDA(h, cudaMalloc(&t.cid, n_models * sizeof(uint16_t)));
    FQ28_CUDA(h, cudaMalloc(&t.n_touched, sizeof(uint32_t)));
    FQ28_CUDA(h, cudaMalloc(&t.zrun, (size_t)QZ_MAX * 2 * (1u << FIX_LOG) * sizeof(uint16_t)));
    FQ28_CUDA(h, cudaMalloc(&t.zinfo, (QZ_MAX + 1) * sizeof(uint32_t)));
  }
  return FQ28_OK;
}

int tables_from_norm(fq28_handle *h, DevTables &t) {
  k_table_offsets<<<1, 1024, 0, h->stream>>>(t.logs, t.n_models, t.toff);
  FQ28_LAUNCH_CHECK(h);
  const unsigned threads = 64, blocks = (t.n_models + threads - 1) / threads;
  if (t.alphabet == SEQ_A)
    k_build_tables_seq<<<SEQ_N / 4, 128, 0, h->stream>>>(t.norm, t.logs, t.toff, t.ctab, t.symtt, t.dtab, t.dtab_fix, t.dom_sym,
                                                        reinterpret_cast<SeqDecTables *>(t.seqdec));
  else
    k_build_tables<QUAL_A><<<blocks, threads, 0, h->stream>>>(t.norm, t.logs, t.toff, t.n_models, t.ctab, t.symtt, t.dtab, t.dtab_fix, t.dom_sym);
  FQ28_LAUNCH_CHECK(h);
  k_logsuf<<<1, 1024, 0, h->stream>>>(t.logs, t.n_models, t.logsuf);
  FQ28_LAUNCH_CHECK(h);
  if (t.alphabet == SEQ_A) {
    k_seq_wtab<<<SEQ_N, 256, 0, h->stream>>>(t.logs, t.dtab_fix, t.wtab);
    FQ28_LAUNCH_CHECK(h);
  } else {
    k_qual_cid<<<1, 1024, 0, h->stream>>>(t.norm, t.logs, t.cid, t.n_touched);
    FQ28_LAUNCH_CHECK(h);
    k_qual_dense<<<1, 64, 0, h->stream>>>(t.cid, t.qrk, t.qdinfo);  // before k_qual_zrun tags cid with run slots
    FQ28_LAUNCH_CHECK(h);
    k_qual_wtab<<<2 * 64 * 64, 256, 0, h->stream>>>(t.logs, t.dtab_fix, t.qrk, t.qdinfo, t.wtab);
    FQ28_LAUNCH_CHECK(h);
    k_qual_win<<<1, 128, 0, h->stream>>>(t.cid, t.qrk, t.qdinfo, t.qwin);
    FQ28_LAUNCH_CHECK(h);
    k_qual_wtabw<<<2 * 64 * 64, 256, 0, h->stream>>>(t.logs, t.dtab_fix, t.qrk, t.qdinfo, t.qwin, t.wtabw);
    FQ28_LAUNCH_CHECK(h);
    k_qual_zrun<<<1, 1024, 0, h->stream>>>(t.logs, t.dtab_fix, t.dom_sym, t.cid, t.zrun, t.zinfo);
    FQ28_LAUNCH_CHECK(h);
    uint32_t zi[QZ_MAX + 1];
    FQ28_CUDA(h, cudaMemcpyAsync(&t.h_n_touched, t.n_touched, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    uint32_t qd[3] = {0, 0, 0};
    FQ28_CUDA(h, cudaMemcpyAsync(qd, t.qdinfo, sizeof(qd), cudaMemcpyDeviceToHost, h->stream));
    FQ28_CUDA(h, cudaMemcpyAsync(zi, t.zinfo, sizeof(zi), cudaMemcpyDeviceToHost, h->stream));
    FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
    t.h_n_v = qd[0];
    t.h_n_win = qd[1];
    t.h_n_rows = qd[2];
    t.h_n_z = zi[0];
    for (unsigned j = 0; j < QZ_MAX; j++) t.h_zctx[j] = zi[1 + j];
  }
  t.ready = true;
  h->tables_gen++;  // anything derived from older tables is stale now
  return FQ28_OK;
}

int tables_from_counts(fq28_handle *h, DevTables &t, const uint32_t *d_counts) {
  FQ28_CUDA(h, cudaMemsetAsync(t.max_log, 0, sizeof(uint32_t), h->stream));
  const unsigned threads = 64, blocks = (t.n_models + threads - 1) / threads;
  if (t.alphabet == SEQ_A)
    k_normalize<SEQ_A><<<blocks, threads, 0, h->stream>>>(d_counts, t.n_models, t.norm, t.logs, t.max_log);
  else
    k_normalize<QUAL_A><<<blocks, threads, 0, h->stream>>>(d_counts, t.n_models, t.norm, t.logs, t.max_log);
  FQ28_LAUNCH_CHECK(h);
  return tables_from_norm(h, t);
}

}  // namespace fq28
