// fq28_parse.cu -- K1: newline scan, line index, record table and the chunk
// boundary walk.  Replaces FastqReader::parseRecords (src/fastq_io.cpp:67-125)
// and the boundary rule of FastqReader::readNextChunk (src/fastq_io.cpp:23-65).
//
// HBM-bound: the slab is read twice with 128-bit loads (count pass, fill pass);
// output is one u32 per newline plus ~18 B per record.
#include "fq28_internal.cuh"

namespace fq28 {

constexpr int NL_THREADS = 256;
constexpr int NL_BYTES_PER_THREAD = 64;
constexpr int NL_TILE = NL_THREADS * NL_BYTES_PER_THREAD;  // 16 KB per CTA

__device__ __forceinline__ unsigned nl_mask4(unsigned w) {
  // 0xFF in every byte lane equal to '\n'
  return __vcmpeq4(w, 0x0A0A0A0Au);
}

// loads the 64 bytes [pos, pos+64) of this thread as 16 words; bytes past
// n_bytes read as 0
__device__ __forceinline__ void load64(const char *__restrict__ d, size_t pos, size_t n_bytes,
                                       unsigned w[16]) {
  if (pos + 64 <= n_bytes) {
    const uint4 *p = reinterpret_cast<const uint4 *>(d + pos);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint4 v = __ldg(p + i);
      w[4 * i + 0] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      unsigned v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        size_t q = pos + 4 * i + b;
        if (q < n_bytes) v |= (unsigned)(unsigned char)d[q] << (8 * b);
      }
      w[i] = v;
    }
  }
}

__global__ void __launch_bounds__(NL_THREADS)
k_count_nl(const char *__restrict__ d, size_t n_bytes, uint32_t *__restrict__ tile_cnt) {
  __shared__ unsigned wsum[NL_THREADS / 32];
  const size_t pos = (size_t)blockIdx.x * NL_TILE + (size_t)threadIdx.x * NL_BYTES_PER_THREAD;
  unsigned cnt = 0;
  if (pos < n_bytes) {
    unsigned w[16];
    load64(d, pos, n_bytes, w);
#pragma unroll
    for (int i = 0; i < 16; i++) cnt += __popc(nl_mask4(w[i])) >> 3;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < NL_THREADS / 32; i++) t += wsum[i];
    tile_cnt[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(NL_THREADS)
k_fill_nl(const char *__restrict__ d, size_t n_bytes, const uint32_t *__restrict__ tile_base,
          uint32_t *__restrict__ nl) {
  __shared__ unsigned wsum[NL_THREADS / 32];
  const size_t pos = (size_t)blockIdx.x * NL_TILE + (size_t)threadIdx.x * NL_BYTES_PER_THREAD;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned w[16];
  unsigned cnt = 0;
  if (pos < n_bytes) {
    load64(d, pos, n_bytes, w);
#pragma unroll
    for (int i = 0; i < 16; i++) cnt += __popc(nl_mask4(w[i])) >> 3;
  }
  // exclusive scan of cnt over the CTA
  unsigned inc = cnt;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    unsigned o = __shfl_up_sync(0xffffffffu, inc, dd);
    if (lane >= (unsigned)dd) inc += o;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  unsigned woff = 0;
#pragma unroll
  for (int i = 0; i < NL_THREADS / 32; i++) woff += (i < (int)warp) ? wsum[i] : 0u;
  unsigned rank = tile_base[blockIdx.x] + woff + inc - cnt;
  if (cnt) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      unsigned m = nl_mask4(w[i]) & 0x01010101u;
      while (m) {
        int b = (__ffs(m) - 1) >> 3;
        nl[rank++] = (uint32_t)(pos + 4 * i + b);
        m &= m - 1;
      }
    }
  }
}

// One thread per record r: lines 4r..4r+3.  hdr_off has n_rec+1 entries, the
// last being the end of the last complete record (parseRecords' return value).
__global__ void k_records(const char *__restrict__ d, const uint32_t *__restrict__ nl, size_t n_rec,
                          uint32_t *__restrict__ hdr_off, uint32_t *__restrict__ seq_off,
                          uint32_t *__restrict__ qual_off, uint16_t *__restrict__ len,
                          uint16_t *__restrict__ hdr_len, DevStatus *st) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rec) return;
  const uint32_t start = (r == 0) ? 0u : nl[4 * r - 1] + 1u;
  hdr_off[r] = start;
  if (r == n_rec) return;
  const uint32_t e0 = nl[4 * r], e1 = nl[4 * r + 1], e2 = nl[4 * r + 2], e3 = nl[4 * r + 3];
  const uint32_t hl = e0 - start, so = e0 + 1, L = e1 - so, po = e1 + 1, qo = e2 + 1, ql = e3 - qo;
  seq_off[r] = so;
  qual_off[r] = qo;
  len[r] = (uint16_t)L;
  hdr_len[r] = (uint16_t)hl;
  // narrow_cast<readlen_t>(line length) throws for every line, src/fastq_io.cpp:95
  if (hl > 65535u || L > 65535u || ql > 65535u || (e2 - po) > 65535u) {
    set_error(st, FQ28_ERR_LONG, (unsigned)r);
    return;
  }
  // asserts at src/fastq_io.cpp:74,102,107
  if (hl == 0 || d[start] != '@' || e2 == po || d[po] != '+' || ql != L)
    set_error(st, FQ28_ERR_FORMAT, (unsigned)r);
}

// Chunk boundary walk (src/fastq_io.cpp:23-65), one warp, 32-ary search per
// chunk over the monotone record-end array hdr_off[1..n_rec].
// out: chunk_rec[k] / chunk_sym[k] / chunk_byte[k] for k = 0..n_chunks.
__global__ void __launch_bounds__(32)
k_chunk_walk(const uint32_t *__restrict__ hdr_off, const uint32_t *__restrict__ symoff, size_t n_rec,
             size_t n_bytes, size_t R, int eof, size_t first_cut, size_t cap, uint32_t *__restrict__ chunk_rec,
             uint32_t *__restrict__ chunk_sym, uint32_t *__restrict__ chunk_byte,
             uint64_t *__restrict__ n_chunks_out, DevStatus *st) {
  // Speculation: the walk is a chain of dependent loads (one per chunk).  The
  // record index where chunk j ends is predictable from the average record
  // size, so the 32-record windows around the predicted ends of the next 32
  // chunks are fetched together (one memory latency for 32 chunks) and the
  // chunks are then resolved from shared memory; a window that does not
  // bracket the answer falls back to the search below.
  __shared__ uint32_t win[32][32];
  __shared__ unsigned long long wp0[32];
  unsigned spec_i = 32;
  const unsigned lane = threadIdx.x;
  size_t s_rec = 0, k = 0;
  if (lane == 0) { chunk_rec[0] = 0; chunk_sym[0] = 0; chunk_byte[0] = 0; }
  // average record size: the first probe of every chunk brackets the expected
  // answer, which settles fixed-length data in ONE dependent load per chunk
  const size_t avg = n_rec ? ((size_t)hdr_off[n_rec] - hdr_off[0] + n_rec - 1) / n_rec : 1;
  size_t s = n_rec ? hdr_off[0] : 0;  // byte start of the current chunk (carried, never re-loaded)
  size_t last_recs = n_rec ? n_rec : 1, last_bytes = n_rec ? (size_t)hdr_off[n_rec] - hdr_off[0] : 1;
  if (first_cut > s && n_rec) {
    // The slab starts inside a chunk that belongs to the previous slab (several GPUs, one file):
    // that chunk ends at byte `first_cut`, which must be a record boundary.  Emit the head
    // [0, first_cut) as chunk 0 (the caller drops it) and walk on from there.
    size_t lo = 0, hi = n_rec;       // largest r with hdr_off[r] <= first_cut
    while (lo < hi) {
      const size_t step = (hi - lo + 31) / 32;
      size_t p = lo + (size_t)(lane + 1) * step;
      if (p > hi) p = hi;
      const bool ok = (size_t)hdr_off[p] <= first_cut;
      const unsigned c = __popc(__ballot_sync(0xffffffffu, ok));
      size_t nlo = lo + (size_t)c * step, nhi = lo + (size_t)(c + 1) * step;
      if (nlo > hi) nlo = hi;
      if (nhi > hi) nhi = hi;        // probe c was clipped to hi and does not fit
      lo = nlo;
      if (c == 32) break;            // every probe fits, the last one is hi itself
      hi = nhi - 1;
    }
    if ((size_t)hdr_off[lo] != first_cut || lo == 0) {
      if (lane == 0) { set_error(st, FQ28_ERR_FORMAT, (unsigned)lo); *n_chunks_out = 0; }
      return;
    }
    k = 1;
    if (lane == 0) { chunk_rec[1] = (uint32_t)lo; chunk_byte[1] = (uint32_t)first_cut; }
    s_rec = lo;
    s = first_cut;
  }
  for (;;) {
    if (s_rec >= n_rec) {
      // bytes left but no complete record: in the reference this is a
      // zero-record chunk (malformed tail); reported only at EOF
      break;
    }
    const size_t limit = s + R;
    const bool reaches_eof = limit >= n_bytes;
    if (!eof && limit > n_bytes) break;
    const size_t L = reaches_eof ? n_bytes : limit;
    size_t lo = s_rec;
    size_t hi = s_rec + R / 12 + 1;  // a record is at least 12 bytes
    if (hi > n_rec) hi = n_rec;
    size_t lo_val = s;               // hdr_off[lo]
    bool resolved = false;
    if (spec_i >= 32 && avg) {       // fetch the windows of the next 32 chunks
      // records per chunk, 8 fractional bits, from the record size of the last chunk (record
      // sizes drift along a file: read ids grow), of the whole slab before the first chunk
      const size_t rpc256 = (size_t)((double)R * 256.0 * (double)last_recs / (double)last_bytes);
      // (fully unrolled: all 32 loads are in flight before the first one is stored -- one memory
      // latency for the 32 windows instead of one per batch of 8)
#pragma unroll
      for (unsigned j = 0; j < 32; j++) {
        // chunk j ends about (j+1) chunks of records ahead, minus half a record of slack per chunk
        size_t est = s_rec + (((size_t)(j + 1) * rpc256 - (size_t)j * 128) >> 8);
        if (est > n_rec) est = n_rec;
        const size_t p0 = est >= 16 ? est - 16 : 0;
        size_t p = p0 + lane;
        if (p > n_rec) p = n_rec;
        win[j][lane] = hdr_off[p];
        if (lane == 0) wp0[j] = p0;
      }
      spec_i = 0;
      __syncwarp();
    }
    if (spec_i < 32 && !reaches_eof) {
      const size_t p0 = (size_t)wp0[spec_i];
      const size_t v = win[spec_i][lane];
      ++spec_i;
      const unsigned m = __ballot_sync(0xffffffffu, v <= L);
      const unsigned c = __popc(m);
      // valid when the window starts at or below the answer and ends above it
      if ((m & 1u) && c < 32 && p0 + c - 1 <= hi && p0 + c - 1 > s_rec) {
        lo = p0 + c - 1;
        lo_val = __shfl_sync(0xffffffffu, v, c - 1);
        resolved = true;
      } else {
        spec_i = 32;                 // mis-predicted: search, then predict again from the next chunk
      }
    }
    if (!resolved) {  // bracket probe: 32 consecutive records around the estimate
      size_t est = s_rec + (L - s) / (avg ? avg : 1);
      if (est > hi) est = hi;
      size_t p0 = est >= s_rec + 16 ? est - 16 : s_rec;
      if (p0 + 31 > hi) p0 = hi >= 31 && hi - 31 >= s_rec ? hi - 31 : s_rec;
      size_t p = p0 + lane;
      if (p > hi) p = hi;
      const size_t v = hdr_off[p];
      const bool ok = v <= L;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (m & 1u) {                  // p0 is a valid lower bound
        const unsigned c = __popc(m);  // ok is monotone: a prefix of the lanes
        lo = p0 + c - 1 > hi ? hi : p0 + c - 1;
        lo_val = __shfl_sync(0xffffffffu, v, c - 1);
        if (c < 32) hi = lo;         // lane c is the first record that does not fit
      } else {
        hi = p0 - 1;                 // answer is below the bracket (p0 > s_rec here)
      }
    }
    while (!resolved && lo < hi) {
      const size_t step = (hi - lo + 31) / 32;
      size_t p = lo + (size_t)(lane + 1) * step;
      if (p > hi) p = hi;
      const size_t v = hdr_off[p];
      const bool ok = v <= L;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      const unsigned c = __popc(m);
      if (c == 32) { lo = hi; lo_val = __shfl_sync(0xffffffffu, v, 31); break; }
      size_t nlo = lo + (size_t)c * step;          // p_c (== lo when c == 0)
      if (nlo > hi) nlo = hi;
      size_t nhi = lo + (size_t)(c + 1) * step;    // p_{c+1}, not ok
      if (nhi > hi) nhi = hi;
      if (c) lo_val = __shfl_sync(0xffffffffu, v, c - 1);
      lo = nlo;
      hi = nhi - 1;
    }
    if (lo == s_rec) {  // record longer than the window: reference UB
      if (lane == 0) set_error(st, FQ28_ERR_FORMAT, (unsigned)s_rec);
      break;
    }
    ++k;
    if (k > cap) {
      if (lane == 0) set_error(st, FQ28_ERR_CAP, (unsigned)k);
      --k;
      break;
    }
    if (lane == 0) {
      chunk_rec[k] = (uint32_t)lo;
      chunk_byte[k] = (uint32_t)lo_val;
    }
    last_recs = lo - s_rec;
    last_bytes = lo_val - s;
    s_rec = lo;
    s = lo_val;
    if (reaches_eof) break;
  }
  if (lane == 0) *n_chunks_out = k;
  // (the symbol offsets of the chunk starts are gathered by k_chunk_sym afterwards: in here the
  // one warp would pay two dependent loads per 32 chunks, ~50 us per 1 000 chunks)
}

// chunk_sym[k] = symoff[chunk_rec[k]] for the chunks the walk emitted
__global__ void k_chunk_sym(const uint32_t *__restrict__ symoff, const uint32_t *__restrict__ chunk_rec,
                            const uint64_t *__restrict__ n_chunks, size_t cap, uint32_t *__restrict__ chunk_sym) {
  const size_t kk = 1 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk <= *n_chunks && kk <= cap) chunk_sym[kk] = symoff ? symoff[chunk_rec[kk]] : 0u;
}

int parse_slab(fq28_handle *h, const char *d_fastq, size_t n_bytes, bool need_symoff) {
  if (n_bytes > FQ28_MAX_SLAB) return fail(h, FQ28_ERR_ARG, "slab of %zu bytes exceeds FQ28_MAX_SLAB", n_bytes);
  if ((reinterpret_cast<uintptr_t>(d_fastq) & 15) != 0)
    return fail(h, FQ28_ERR_ARG, "device FASTQ pointer must be 16-byte aligned");
  h->d_fastq = d_fastq;
  h->n_bytes = n_bytes;
  h->n_lines = h->n_rec = 0;
  h->n_chunks = 0;
  h->plan.valid = false;  // the record table of any earlier plan is gone
  h->parsed.valid = false;
  if (h->extracted) {     // an eager field separation nobody consumed still reads the old record table
    FQ28_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_extract, 0));
    h->extracted = false;
  }
  FQ28_CUDA(h, cudaMemsetAsync(h->d_status, 0, sizeof(DevStatus), h->stream));
  const size_t n_tiles = (n_bytes + NL_TILE - 1) / NL_TILE;
  FQ28_TRY(ensure(h, h->tile_cnt, (n_tiles + 1) * sizeof(uint32_t)));
  uint32_t *tile_cnt = h->tile_cnt.as<uint32_t>();
  if (n_tiles) {
    k_count_nl<<<(unsigned)n_tiles, NL_THREADS, 0, h->stream>>>(d_fastq, n_bytes, tile_cnt);
    FQ28_LAUNCH_CHECK(h);
  }
  FQ28_TRY(scan_exclusive_u32(h, tile_cnt, tile_cnt, n_tiles));
  uint32_t *n_lines32 = reinterpret_cast<uint32_t *>(h->h_scalars + 60);   // (pinned)
  FQ28_CUDA(h, cudaMemcpyAsync(n_lines32, tile_cnt + n_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_lines = *n_lines32;
  h->n_rec = h->n_lines / 4;
  const size_t n_rec = h->n_rec;
  FQ28_TRY(ensure(h, h->nl, (h->n_lines + 4) * sizeof(uint32_t)));
  FQ28_TRY(ensure(h, h->hdr_off, (n_rec + 1) * sizeof(uint32_t)));
  FQ28_TRY(ensure(h, h->seq_off, (n_rec + 1) * sizeof(uint32_t)));
  FQ28_TRY(ensure(h, h->qual_off, (n_rec + 1) * sizeof(uint32_t)));
  FQ28_TRY(ensure(h, h->len, (n_rec + 1) * sizeof(uint16_t)));
  FQ28_TRY(ensure(h, h->hdr_len, (n_rec + 1) * sizeof(uint16_t)));
  FQ28_TRY(ensure(h, h->symoff, (n_rec + 2) * sizeof(uint32_t)));
  if (n_tiles && h->n_lines) {
    k_fill_nl<<<(unsigned)n_tiles, NL_THREADS, 0, h->stream>>>(d_fastq, n_bytes, tile_cnt, h->nl.as<uint32_t>());
    FQ28_LAUNCH_CHECK(h);
  }
  {
    const unsigned threads = 256;
    const unsigned blocks = (unsigned)((n_rec + 1 + threads - 1) / threads);
    k_records<<<blocks, threads, 0, h->stream>>>(d_fastq, h->nl.as<uint32_t>(), n_rec, h->hdr_off.as<uint32_t>(),
                                                 h->seq_off.as<uint32_t>(), h->qual_off.as<uint32_t>(),
                                                 h->len.as<uint16_t>(), h->hdr_len.as<uint16_t>(), h->d_status);
    FQ28_LAUNCH_CHECK(h);
  }
  if (need_symoff) FQ28_TRY(scan_exclusive_u16_to_u32(h, h->len.as<uint16_t>(), h->symoff.as<uint32_t>(), n_rec));
  return FQ28_OK;
}

int split_slab(fq28_handle *h, size_t reading_size, bool eof, size_t max_chunks, size_t first_cut) {
  if (reading_size == 0) return fail(h, FQ28_ERR_ARG, "reading_size must be >= 1 (SURVEY Q6)");
  if (first_cut >= h->n_bytes && first_cut) return fail(h, FQ28_ERR_ARG, "first_cut beyond the slab");
  size_t cap = 2 * (h->n_bytes / reading_size) + 5;
  if (max_chunks && cap > max_chunks) cap = max_chunks;
  FQ28_TRY(ensure(h, h->chunk_rec, 3 * (cap + 1) * sizeof(uint32_t)));
  h->chunk_stride = cap + 1;
  uint32_t *cr = h->chunk_rec.as<uint32_t>();
  // While an eager field separation fills the GPU (same priority as the main stream), the one-warp
  // walk would queue behind its CTAs: it goes to the high-priority side stream then.  The record
  // table it reads was complete before the separation started (ev_parsed).
  cudaStream_t ws = h->stream;
  if (h->extracted) {
    ws = h->side;
    FQ28_CUDA(h, cudaStreamWaitEvent(ws, h->ev_parsed, 0));
  }
  k_chunk_walk<<<1, 32, 0, ws>>>(h->hdr_off.as<uint32_t>(), h->symoff.as<uint32_t>(), h->n_rec, h->n_bytes,
                                       reading_size, eof ? 1 : 0, first_cut, cap, cr, cr + (cap + 1), cr + 2 * (cap + 1),
                                       h->d_scalars, h->d_status);
  FQ28_LAUNCH_CHECK(h);
  k_chunk_sym<<<(unsigned)((cap + 255) / 256), 256, 0, ws>>>(h->symoff.as<uint32_t>(), cr, h->d_scalars, cap, cr + (cap + 1));
  FQ28_LAUNCH_CHECK(h);
  FQ28_CUDA(h, cudaMemcpyAsync(h->h_scalars, h->d_scalars, sizeof(uint64_t), cudaMemcpyDeviceToHost, ws));
  FQ28_TRY(check_status(h, "record splitting", ws));
  const size_t n = (size_t)h->h_scalars[0];
  h->n_chunks = n;
  h->h_chunk_rec.resize(n + 1);
  h->h_chunk_sym.resize(n + 1);
  h->h_chunk_byte.resize(n + 1);
  FQ28_TRY(ensure_pinned(h, 3 * (n + 1) * 4));
  uint32_t *pin = static_cast<uint32_t *>(h->h_pin);
  FQ28_CUDA(h, cudaMemcpyAsync(pin, cr, (n + 1) * 4, cudaMemcpyDeviceToHost, ws));
  FQ28_CUDA(h, cudaMemcpyAsync(pin + (n + 1), cr + (cap + 1), (n + 1) * 4, cudaMemcpyDeviceToHost, ws));
  FQ28_CUDA(h, cudaMemcpyAsync(pin + 2 * (n + 1), cr + 2 * (cap + 1), (n + 1) * 4, cudaMemcpyDeviceToHost, ws));
  FQ28_CUDA(h, cudaStreamSynchronize(ws));
  memcpy(h->h_chunk_rec.data(), pin, (n + 1) * 4);
  memcpy(h->h_chunk_sym.data(), pin + (n + 1), (n + 1) * 4);
  memcpy(h->h_chunk_byte.data(), pin + 2 * (n + 1), (n + 1) * 4);
  return FQ28_OK;
}

}  // namespace fq28
