// fq28_headers.cu -- header tokeniser on the GPU (SURVEY section 8(f) row N2,
// encode side): CompressionWorkspace::encodeHeader (src/workspace.cpp:95-125)
// with storeString / storeNumeric (src/headers.cpp:75-89,108-118) for every
// header of a batch of chunks.  Field splitting and std::from_chars are per
// record; "differs from the previous record" looks one record back (the first
// record of a chunk looks at the archive's first header); the positions inside
// the content / contentLength streams are prefix sums over the records.
//
//   k_hdr_split  : thread per record -> field boundaries, numeric values
//   k_hdr_diff   : thread per record -> per STRING field: differs?, length if so
//   scans        : exclusive prefix sums of the flags and lengths, per field
//   k_hdr_layout : single CTA -> stream sizes and arena offsets per (chunk, field)
//   k_hdr_emit   : thread per record -> writes deltas, flags, values, lengths
#include "fq28_internal.cuh"

namespace fq28 {

struct HdrFmtDev {
  unsigned n_fields;
  unsigned char is_string[FQ28_HDR_MAX_FIELDS];
  char sep[FQ28_HDR_MAX_FIELDS];
  int first_num[FQ28_HDR_MAX_FIELDS];
  unsigned first_off[FQ28_HDR_MAX_FIELDS + 1];
};

// std::from_chars<int32_t>(p, e, v) with v preset to 0: optional '-', then digits;
// no digit or out of range -> v stays 0.
__device__ __forceinline__ int parse_i32(const unsigned char *p, const unsigned char *e) {
  bool neg = false;
  if (p < e && *p == '-') { neg = true; ++p; }
  long long acc = 0;
  bool any = false, over = false;
  for (; p < e; ++p) {
    const unsigned d = (unsigned)*p - '0';
    if (d > 9u) break;
    any = true;
    if (!over) {
      acc = acc * 10 + d;
      if (acc > 2147483648LL) over = true;  // keeps consuming digits like from_chars
    }
  }
  if (!any || over) return 0;
  if (neg) return (int)(-acc);               // -2147483648 is representable
  if (acc > 2147483647LL) return 0;
  return (int)acc;
}

__global__ void k_hdr_split(const unsigned char *__restrict__ hdr, const uint32_t *__restrict__ hoff,
                            const uint16_t *__restrict__ hlen, size_t n_rec, HdrFmtDev fmt,
                            uint16_t *__restrict__ fbeg, uint16_t *__restrict__ fend, int *__restrict__ val) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const unsigned char *h0 = hdr + hoff[r];
  const unsigned hl = hlen[r];
  unsigned p = 1;
  for (unsigned i = 0; i < fmt.n_fields; i++) {
    unsigned e = hl;
    if (i + 1 < fmt.n_fields) {  // std::find(std::min(p + 1, end), end, separators[i])
      unsigned q = p + 1 < hl ? p + 1 : hl;
      while (q < hl && h0[q] != (unsigned char)fmt.sep[i]) ++q;
      e = q;
    }
    const unsigned pb = p < hl ? p : hl;     // an exhausted header leaves empty fields at its end
    fbeg[r * fmt.n_fields + i] = (uint16_t)pb;
    fend[r * fmt.n_fields + i] = (uint16_t)(e > pb ? e : pb);
    if (!fmt.is_string[i]) val[r * fmt.n_fields + i] = parse_i32(h0 + pb, h0 + (e > pb ? e : pb));
    p = e < hl ? e + 1 : hl;
  }
}

__device__ __forceinline__ unsigned chunk_of(const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, size_t r) {
  unsigned lo = 0, hi = n_chunks;
  while (hi - lo > 1) {
    const unsigned mid = (lo + hi) >> 1;
    if (chunk_rec[mid] <= r) lo = mid; else hi = mid;
  }
  return lo;
}

// flags[f][r] (u16 0/1) and dlen[f][r] (u16) for the STRING fields, f = index among all fields
__global__ void k_hdr_diff(const unsigned char *__restrict__ hdr, const uint32_t *__restrict__ hoff, size_t n_rec,
                           const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, HdrFmtDev fmt,
                           const unsigned char *__restrict__ first_str, const uint16_t *__restrict__ fbeg,
                           const uint16_t *__restrict__ fend, uint16_t *__restrict__ flags, uint16_t *__restrict__ dlen,
                           DevStatus *st) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const unsigned k = chunk_of(chunk_rec, n_chunks, r);
  const bool first = chunk_rec[k] == r;
  const unsigned F = fmt.n_fields;
  for (unsigned i = 0; i < F; i++) {
    if (!fmt.is_string[i]) continue;
    const unsigned char *a = hdr + hoff[r] + fbeg[r * F + i];
    const unsigned la = fend[r * F + i] - fbeg[r * F + i];
    const unsigned char *b;
    unsigned lb;
    if (first) { b = first_str + fmt.first_off[i]; lb = fmt.first_off[i + 1] - fmt.first_off[i]; }
    else { b = hdr + hoff[r - 1] + fbeg[(r - 1) * F + i]; lb = fend[(r - 1) * F + i] - fbeg[(r - 1) * F + i]; }
    bool diff = la != lb;
    for (unsigned j = 0; !diff && j < la; j++) diff = a[j] != b[j];
    if (diff && la >= 255u) set_error(st, FQ28_ERR_FORMAT, (unsigned)r);  // FIELDLEN_MAX, src/headers.cpp:80
    flags[(size_t)i * n_rec + r] = diff ? 1u : 0u;
    dlen[(size_t)i * n_rec + r] = diff ? (uint16_t)la : (uint16_t)0;
  }
}

// sizes and arena offsets of the streams of every (chunk, field); single CTA, serial over the few entries
__global__ void k_hdr_layout(const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, HdrFmtDev fmt, size_t n_rec,
                             const uint32_t *__restrict__ sflag, const uint32_t *__restrict__ sdlen,
                             fq28_hdr_field_info *__restrict__ infos, uint64_t *__restrict__ total) {
  const unsigned F = fmt.n_fields;
  // sizes in parallel, offsets by one thread
  for (unsigned e = threadIdx.x; e < n_chunks * F; e += blockDim.x) {
    const unsigned k = e / F, i = e % F;
    const size_t r0 = chunk_rec[k], r1 = chunk_rec[k + 1];
    fq28_hdr_field_info fi{};
    if (fmt.is_string[i]) {
      fi.flag_len = r1 - r0;
      fi.clen_len = sflag[(size_t)i * (n_rec + 1) + r1] - sflag[(size_t)i * (n_rec + 1) + r0];
      fi.content_len = sdlen[(size_t)i * (n_rec + 1) + r1] - sdlen[(size_t)i * (n_rec + 1) + r0];
    } else {
      fi.content_len = (r1 - r0) * 4;
    }
    infos[e] = fi;
  }
  __syncthreads();
  // offsets: thread t owns a contiguous block of entries; CTA scan of the block totals
  __shared__ unsigned long long part[256];
  const unsigned E = n_chunks * F, per = (E + blockDim.x - 1) / blockDim.x;
  const unsigned e0 = threadIdx.x * per, e1 = e0 + per < E ? e0 + per : E;
  auto padded = [](const fq28_hdr_field_info &fi) {
    return (fi.flag_len + fi.content_len + fi.clen_len + 3) & ~(uint64_t)3;  // numeric content stays 4-byte aligned
  };
  unsigned long long sum = 0;
  for (unsigned e = e0; e < e1; e++) sum += padded(infos[e]);
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (unsigned t = 0; t < blockDim.x; t++) { const unsigned long long v = part[t]; part[t] = run; run += v; }
    *total = run;
  }
  __syncthreads();
  uint64_t off = part[threadIdx.x];
  for (unsigned e = e0; e < e1; e++) {
    fq28_hdr_field_info fi = infos[e];
    const uint64_t sz = padded(fi);
    fi.flag_off = off;
    fi.content_off = off + fi.flag_len;
    fi.clen_off = fi.content_off + fi.content_len;
    infos[e] = fi;
    off += sz;
  }
}

__global__ void k_hdr_emit(const unsigned char *__restrict__ hdr, const uint32_t *__restrict__ hoff, size_t n_rec,
                           const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, HdrFmtDev fmt,
                           const uint16_t *__restrict__ fbeg, const uint16_t *__restrict__ fend,
                           const int *__restrict__ val, const uint16_t *__restrict__ flags,
                           const uint32_t *__restrict__ sflag, const uint32_t *__restrict__ sdlen,
                           const fq28_hdr_field_info *__restrict__ infos, uint8_t *__restrict__ arena) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const unsigned k = chunk_of(chunk_rec, n_chunks, r);
  const size_t r0 = chunk_rec[k];
  const unsigned F = fmt.n_fields;
  for (unsigned i = 0; i < F; i++) {
    const fq28_hdr_field_info fi = infos[k * F + i];
    if (!fmt.is_string[i]) {  // storeNumeric: delta to the previous record (int32 wrap-around)
      const unsigned v = (unsigned)val[r * F + i];
      const unsigned pv = r == r0 ? (unsigned)fmt.first_num[i] : (unsigned)val[(r - 1) * F + i];
      *reinterpret_cast<unsigned *>(arena + fi.content_off + (r - r0) * 4) = v - pv;
    } else {
      const unsigned f = flags[(size_t)i * n_rec + r];
      arena[fi.flag_off + (r - r0)] = (uint8_t)f;
      if (f) {
        const unsigned la = fend[r * F + i] - fbeg[r * F + i];
        const unsigned char *a = hdr + hoff[r] + fbeg[r * F + i];
        uint8_t *dst = arena + fi.content_off + (sdlen[(size_t)i * (n_rec + 1) + r] - sdlen[(size_t)i * (n_rec + 1) + r0]);
        for (unsigned j = 0; j < la; j++) dst[j] = a[j];
        arena[fi.clen_off + (sflag[(size_t)i * (n_rec + 1) + r] - sflag[(size_t)i * (n_rec + 1) + r0])] = (uint8_t)la;
      }
    }
  }
}

int tokenize_headers(fq28_handle *h, const uint8_t *headers, size_t headers_bytes, const uint16_t *hdr_lens, size_t n_rec,
                     const uint64_t *chunk_rec, size_t n_chunks, const fq28_hdr_format *fmt, uint8_t *arena, size_t arena_cap,
                     fq28_hdr_field_info *infos, size_t *arena_bytes) {
  if (arena_bytes) *arena_bytes = 0;
  const unsigned F = fmt->n_fields;
  if (F == 0 || F > FQ28_HDR_MAX_FIELDS) return fail(h, FQ28_ERR_ARG, "header format with %u fields (1..%d supported)", F, FQ28_HDR_MAX_FIELDS);
  if (n_chunks == 0 || n_rec == 0) return FQ28_OK;
  if (chunk_rec[0] != 0 || chunk_rec[n_chunks] != n_rec) return fail(h, FQ28_ERR_ARG, "chunk_rec must cover records 0..n_records");
  for (size_t k = 0; k < n_chunks; k++)
    if (chunk_rec[k + 1] <= chunk_rec[k]) return fail(h, FQ28_ERR_ARG, "chunk %zu holds no record", k);
  HdrFmtDev df{};
  df.n_fields = F;
  for (unsigned i = 0; i < F; i++) {
    df.is_string[i] = fmt->is_string[i];
    df.sep[i] = fmt->separators[i];
    df.first_num[i] = fmt->first_numeric[i];
    df.first_off[i] = fmt->first_str_off[i];
  }
  df.first_off[F] = fmt->first_str_off[F];
  FQ28_CUDA(h, cudaMemsetAsync(h->d_status, 0, sizeof(DevStatus), h->stream));
  // staging: [headers][hdr_lens u16][chunk_rec u64][first strings]
  const size_t fs_bytes = fmt->first_str_off[F];
  DevBuf &in = h->dec_in[0];
  const size_t o_len = (headers_bytes + 15) & ~(size_t)15, o_cr = (o_len + n_rec * 2 + 15) & ~(size_t)15;
  const size_t o_fs = o_cr + (n_chunks + 1) * 8;
  FQ28_TRY(ensure(h, in, o_fs + fs_bytes + 64));
  uint8_t *d_in = in.as<uint8_t>();
  FQ28_CUDA(h, cudaMemcpyAsync(d_in, headers, headers_bytes, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(d_in + o_len, hdr_lens, n_rec * 2, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(d_in + o_cr, chunk_rec, (n_chunks + 1) * 8, cudaMemcpyHostToDevice, h->stream));
  if (fs_bytes) FQ28_CUDA(h, cudaMemcpyAsync(d_in + o_fs, fmt->first_strings, fs_bytes, cudaMemcpyHostToDevice, h->stream));
  const uint16_t *d_len = reinterpret_cast<const uint16_t *>(d_in + o_len);
  const uint64_t *d_cr = reinterpret_cast<const uint64_t *>(d_in + o_cr);
  // work arrays
  DevBuf &w_off = h->dec_in[1], &w_be = h->dec_in[2], &w_val = h->dec_in[3], &w_fl = h->dec_in[4], &w_sc = h->dec_in[5],
         &w_info = h->dec_in[6];
  FQ28_TRY(ensure(h, w_off, (n_rec + 2) * 4));
  FQ28_TRY(ensure(h, w_be, n_rec * F * 2 * 2 + 64));
  FQ28_TRY(ensure(h, w_val, n_rec * F * 4 + 64));
  FQ28_TRY(ensure(h, w_fl, n_rec * F * 2 * 2 + 64));
  FQ28_TRY(ensure(h, w_sc, (n_rec + 1) * F * 4 * 2 + 64));
  FQ28_TRY(ensure(h, w_info, n_chunks * F * sizeof(fq28_hdr_field_info) + 64));
  uint32_t *hoff = w_off.as<uint32_t>();
  uint16_t *fbeg = w_be.as<uint16_t>(), *fend = fbeg + n_rec * F;
  int *val = w_val.as<int>();
  uint16_t *flags = w_fl.as<uint16_t>(), *dlen = flags + n_rec * F;
  uint32_t *sflag = w_sc.as<uint32_t>(), *sdlen = sflag + (n_rec + 1) * F;
  fq28_hdr_field_info *d_infos = w_info.as<fq28_hdr_field_info>();
  FQ28_TRY(scan_exclusive_u16_to_u32(h, d_len, hoff, n_rec));
  const unsigned threads = 128, blocks = (unsigned)((n_rec + threads - 1) / threads);
  k_hdr_split<<<blocks, threads, 0, h->stream>>>(d_in, hoff, d_len, n_rec, df, fbeg, fend, val);
  FQ28_LAUNCH_CHECK(h);
  k_hdr_diff<<<blocks, threads, 0, h->stream>>>(d_in, hoff, n_rec, d_cr, (unsigned)n_chunks, df, d_in + o_fs, fbeg, fend, flags,
                                               dlen, h->d_status);
  FQ28_LAUNCH_CHECK(h);
  for (unsigned i = 0; i < F; i++) {
    if (!df.is_string[i]) continue;
    FQ28_TRY(scan_exclusive_u16_to_u32(h, flags + (size_t)i * n_rec, sflag + (size_t)i * (n_rec + 1), n_rec));
    FQ28_TRY(scan_exclusive_u16_to_u32(h, dlen + (size_t)i * n_rec, sdlen + (size_t)i * (n_rec + 1), n_rec));
  }
  k_hdr_layout<<<1, 256, 0, h->stream>>>(d_cr, (unsigned)n_chunks, df, n_rec, sflag, sdlen, d_infos, h->d_scalars + 32);
  FQ28_LAUNCH_CHECK(h);
  FQ28_CUDA(h, cudaMemcpyAsync(h->h_scalars + 32, h->d_scalars + 32, 8, cudaMemcpyDeviceToHost, h->stream));
  FQ28_TRY(check_status(h, "header tokeniser"));
  const size_t total = (size_t)h->h_scalars[32];
  if (total > arena_cap) return fail(h, FQ28_ERR_CAP, "header streams need %zu bytes, arena has %zu", total, arena_cap);
  FQ28_TRY(ensure(h, h->dec_out, total + 64));
  FQ28_CUDA(h, cudaMemsetAsync(h->dec_out.p, 0, total + 8, h->stream));  // alignment gaps
  k_hdr_emit<<<blocks, threads, 0, h->stream>>>(d_in, hoff, n_rec, d_cr, (unsigned)n_chunks, df, fbeg, fend, val, flags, sflag,
                                               sdlen, d_infos, h->dec_out.as<uint8_t>());
  FQ28_LAUNCH_CHECK(h);
  FQ28_CUDA(h, cudaMemcpyAsync(arena, h->dec_out.p, total, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(infos, d_infos, n_chunks * F * sizeof(fq28_hdr_field_info), cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  if (arena_bytes) *arena_bytes = total;
  return FQ28_OK;
}

// ---------------------------------------------------------------------------
// decode side: DecompressionWorkspace::decodeHeader (src/workspace.cpp:127-157)
//   k_dtk_gather : thread per record -> numeric deltas and string flags, record-major per field
//   k_dtk_clen   : thread per stored string value -> its length, values of all chunks back to back
//   scans        : deltas (mod 2^32) -> values; flags -> index of the value in force; lengths -> content offsets
//   k_dtk_len    : thread per record -> header length
//   k_dtk_write  : thread per record -> '@' field sep field ...
// ---------------------------------------------------------------------------
__global__ void k_dtk_check(const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, HdrFmtDev fmt,
                            const fq28_hdr_field_info *__restrict__ infos, size_t arena_bytes, DevStatus *st) {
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_chunks * fmt.n_fields) return;
  const unsigned k = e / fmt.n_fields, i = e % fmt.n_fields;
  const uint64_t n = chunk_rec[k + 1] - chunk_rec[k];
  const fq28_hdr_field_info fi = infos[e];
  bool bad = fi.flag_off + fi.flag_len > arena_bytes || fi.content_off + fi.content_len > arena_bytes ||
             fi.clen_off + fi.clen_len > arena_bytes;
  if (fmt.is_string[i]) bad |= fi.flag_len != n || fi.clen_len > n;
  else bad |= fi.content_len != n * 4;
  if (bad) set_error(st, FQ28_ERR_STREAM, k);
}

__global__ void k_dtk_gather(const uint8_t *__restrict__ arena, const fq28_hdr_field_info *__restrict__ infos, size_t n_rec,
                             const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, HdrFmtDev fmt,
                             uint32_t *__restrict__ delta, uint16_t *__restrict__ flags) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const unsigned k = chunk_of(chunk_rec, n_chunks, r);
  const size_t j = r - chunk_rec[k];
  const unsigned F = fmt.n_fields;
  for (unsigned i = 0; i < F; i++) {
    const fq28_hdr_field_info fi = infos[k * F + i];
    if (fmt.is_string[i]) {
      flags[(size_t)i * n_rec + r] = arena[fi.flag_off + j] != 0 ? 1u : 0u;
    } else {
      const uint8_t *p = arena + fi.content_off + j * 4;
      delta[(size_t)i * n_rec + r] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
  }
}

// lengths of the stored values of STRING field i, all chunks back to back: value v of chunk k sits at sflag[r0(k)] + v
__global__ void k_dtk_clen(const uint8_t *__restrict__ arena, const fq28_hdr_field_info *__restrict__ infos,
                           const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, HdrFmtDev fmt, unsigned field,
                           size_t n_rec, const uint32_t *__restrict__ sflag_f, uint16_t *__restrict__ clen_f, DevStatus *st) {
  const unsigned k = blockIdx.y;
  const fq28_hdr_field_info fi = infos[k * fmt.n_fields + field];
  const size_t r0 = chunk_rec[k], r1 = chunk_rec[k + 1];
  const uint32_t base = sflag_f[r0], n_set = sflag_f[r1] - base;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_set != fi.clen_len) set_error(st, FQ28_ERR_STREAM, k);
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_set && v < fi.clen_len; v += (size_t)gridDim.x * blockDim.x)
    clen_f[base + v] = arena[fi.clen_off + v];
}

struct DtkField { unsigned len; const unsigned char *src; int num; bool is_num; };

__device__ __forceinline__ unsigned dec_len(int v) {  // characters std::to_chars writes
  unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v, n = v < 0 ? 1u : 0u;
  do { ++n; u /= 10u; } while (u);
  return n;
}
__device__ __forceinline__ void dec_write(unsigned char *dst, int v, unsigned n) {
  unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
  for (unsigned j = n; j-- > (v < 0 ? 1u : 0u);) { dst[j] = (unsigned char)('0' + u % 10u); u /= 10u; }
  if (v < 0) dst[0] = '-';
}

// field i of record r (the value in force): numeric value, or pointer + length of the string
__device__ __forceinline__ DtkField dtk_field(size_t r, size_t r0, unsigned k, unsigned i, size_t n_rec, const HdrFmtDev &fmt,
                                              const uint8_t *__restrict__ arena, const fq28_hdr_field_info *__restrict__ infos,
                                              const unsigned char *__restrict__ first_str, const uint32_t *__restrict__ sdelta,
                                              const uint32_t *__restrict__ sflag, const uint32_t *__restrict__ sclen,
                                              const uint16_t *__restrict__ clen, DevStatus *st) {
  DtkField f{};
  if (!fmt.is_string[i]) {
    const uint32_t *sd = sdelta + (size_t)i * (n_rec + 1);
    f.is_num = true;
    f.num = (int)((unsigned)fmt.first_num[i] + (sd[r + 1] - sd[r0]));  // prev += delta, int32 wrap-around
    f.len = dec_len(f.num);
    return f;
  }
  const uint32_t *sf = sflag + (size_t)i * (n_rec + 1);
  const uint32_t n_set = sf[r + 1] - sf[r0];  // values stored in the chunk up to and including this record
  if (n_set == 0) {
    f.src = first_str + fmt.first_off[i];
    f.len = fmt.first_off[i + 1] - fmt.first_off[i];
    return f;
  }
  const fq28_hdr_field_info fi = infos[k * fmt.n_fields + i];
  const uint32_t *sc = sclen + (size_t)i * (n_rec + 1);
  const uint32_t v = sf[r0] + n_set - 1;  // global index of the value in force
  const uint32_t off = sc[v] - sc[sf[r0]];
  f.len = clen[(size_t)i * n_rec + v];
  if ((uint64_t)off + f.len > fi.content_len) { set_error(st, FQ28_ERR_STREAM, k); f.len = 0; }
  f.src = arena + fi.content_off + off;
  return f;
}

template <bool WRITE>
__global__ void k_dtk_emit(size_t n_rec, const uint64_t *__restrict__ chunk_rec, unsigned n_chunks, HdrFmtDev fmt,
                           const uint8_t *__restrict__ arena, const fq28_hdr_field_info *__restrict__ infos,
                           const unsigned char *__restrict__ first_str, const uint32_t *__restrict__ sdelta,
                           const uint32_t *__restrict__ sflag, const uint32_t *__restrict__ sclen,
                           const uint16_t *__restrict__ clen, uint32_t *__restrict__ len32 /*WRITE: offsets*/,
                           uint16_t *__restrict__ len16, unsigned char *__restrict__ out, DevStatus *st) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const unsigned k = chunk_of(chunk_rec, n_chunks, r);
  const size_t r0 = chunk_rec[k];
  unsigned char *dst = WRITE ? out + len32[r] : nullptr;
  unsigned pos = 1;
  if (WRITE) dst[0] = '@';
  for (unsigned i = 0; i < fmt.n_fields; i++) {
    const DtkField f = dtk_field(r, r0, k, i, n_rec, fmt, arena, infos, first_str, sdelta, sflag, sclen, clen, st);
    if (WRITE) {
      if (f.is_num) dec_write(dst + pos, f.num, f.len);
      else for (unsigned j = 0; j < f.len; j++) dst[pos + j] = f.src[j];
    }
    pos += f.len;
    if (i + 1 < fmt.n_fields) { if (WRITE) dst[pos] = (unsigned char)fmt.sep[i]; ++pos; }
  }
  if (!WRITE) {
    if (pos > 65535u) { set_error(st, FQ28_ERR_LONG, (unsigned)r); pos = 65535u; }  // narrow_cast<readlen_t>
    len32[r] = pos;
    len16[r] = (uint16_t)pos;
  }
}

int detokenize_headers(fq28_handle *h, const uint8_t *arena, size_t arena_bytes, const fq28_hdr_field_info *infos,
                       const uint64_t *chunk_rec, size_t n_chunks, const fq28_hdr_format *fmt, uint8_t *headers_out,
                       size_t headers_cap, uint16_t *hdr_lens_out, size_t *headers_bytes) {
  if (headers_bytes) *headers_bytes = 0;
  const unsigned F = fmt->n_fields;
  if (F == 0 || F > FQ28_HDR_MAX_FIELDS) return fail(h, FQ28_ERR_ARG, "header format with %u fields (1..%d supported)", F, FQ28_HDR_MAX_FIELDS);
  if (n_chunks == 0) return FQ28_OK;
  if (chunk_rec[0] != 0) return fail(h, FQ28_ERR_ARG, "chunk_rec must start at record 0");
  for (size_t k = 0; k < n_chunks; k++)
    if (chunk_rec[k + 1] <= chunk_rec[k]) return fail(h, FQ28_ERR_ARG, "chunk %zu holds no record", k);
  const size_t n_rec = (size_t)chunk_rec[n_chunks];
  HdrFmtDev df{};
  df.n_fields = F;
  for (unsigned i = 0; i < F; i++) {
    df.is_string[i] = fmt->is_string[i];
    df.sep[i] = fmt->separators[i];
    df.first_num[i] = fmt->first_numeric[i];
    df.first_off[i] = fmt->first_str_off[i];
  }
  df.first_off[F] = fmt->first_str_off[F];
  FQ28_CUDA(h, cudaMemsetAsync(h->d_status, 0, sizeof(DevStatus), h->stream));
  const size_t fs_bytes = fmt->first_str_off[F];
  DevBuf &in = h->dec_in[0], &w_info = h->dec_in[6], &w_a = h->dec_in[1], &w_b = h->dec_in[2], &w_c = h->dec_in[3],
         &w_d = h->dec_in[4], &w_e = h->dec_in[5];
  const size_t o_cr = (arena_bytes + 15) & ~(size_t)15, o_fs = o_cr + (n_chunks + 1) * 8;
  FQ28_TRY(ensure(h, in, o_fs + fs_bytes + 64));
  uint8_t *d_in = in.as<uint8_t>();
  FQ28_CUDA(h, cudaMemcpyAsync(d_in, arena, arena_bytes, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(d_in + o_cr, chunk_rec, (n_chunks + 1) * 8, cudaMemcpyHostToDevice, h->stream));
  if (fs_bytes) FQ28_CUDA(h, cudaMemcpyAsync(d_in + o_fs, fmt->first_strings, fs_bytes, cudaMemcpyHostToDevice, h->stream));
  const uint64_t *d_cr = reinterpret_cast<const uint64_t *>(d_in + o_cr);
  FQ28_TRY(ensure(h, w_info, n_chunks * F * sizeof(fq28_hdr_field_info) + 64));
  fq28_hdr_field_info *d_infos = w_info.as<fq28_hdr_field_info>();
  FQ28_CUDA(h, cudaMemcpyAsync(d_infos, infos, n_chunks * F * sizeof(fq28_hdr_field_info), cudaMemcpyHostToDevice, h->stream));
  FQ28_TRY(ensure(h, w_a, n_rec * F * 4 + 64));            // deltas u32 [F][n_rec]
  FQ28_TRY(ensure(h, w_b, n_rec * F * 2 * 2 + 64));        // flags u16 [F][n_rec] | clen u16 [F][n_rec]
  FQ28_TRY(ensure(h, w_c, (n_rec + 1) * F * 4 * 3 + 64));  // scans: deltas | flags | clen, [F][n_rec + 1] each
  FQ28_TRY(ensure(h, w_d, (n_rec + 2) * 4));               // header lengths / offsets
  FQ28_TRY(ensure(h, w_e, (n_rec + 2) * 2));               // header lengths u16
  uint32_t *delta = w_a.as<uint32_t>();
  uint16_t *flags = w_b.as<uint16_t>(), *clen = flags + n_rec * F;
  uint32_t *sdelta = w_c.as<uint32_t>(), *sflag = sdelta + (n_rec + 1) * F, *sclen = sflag + (n_rec + 1) * F;
  uint32_t *len32 = w_d.as<uint32_t>();
  uint16_t *len16 = w_e.as<uint16_t>();
  const unsigned threads = 128, blocks = (unsigned)((n_rec + threads - 1) / threads);
  k_dtk_check<<<(unsigned)((n_chunks * F + 127) / 128), 128, 0, h->stream>>>(d_cr, (unsigned)n_chunks, df, d_infos, arena_bytes, h->d_status);
  FQ28_LAUNCH_CHECK(h);
  FQ28_TRY(check_status(h, "header streams"));  // sizes are trusted from here on
  FQ28_CUDA(h, cudaMemsetAsync(clen, 0, n_rec * F * 2, h->stream));
  k_dtk_gather<<<blocks, threads, 0, h->stream>>>(d_in, d_infos, n_rec, d_cr, (unsigned)n_chunks, df, delta, flags);
  FQ28_LAUNCH_CHECK(h);
  for (unsigned i = 0; i < F; i++) {
    if (df.is_string[i]) {
      FQ28_TRY(scan_exclusive_u16_to_u32(h, flags + (size_t)i * n_rec, sflag + (size_t)i * (n_rec + 1), n_rec));
      dim3 grid(8, (unsigned)n_chunks);
      k_dtk_clen<<<grid, 128, 0, h->stream>>>(d_in, d_infos, d_cr, (unsigned)n_chunks, df, i, n_rec, sflag + (size_t)i * (n_rec + 1),
                                             clen + (size_t)i * n_rec, h->d_status);
      FQ28_LAUNCH_CHECK(h);
      FQ28_TRY(scan_exclusive_u16_to_u32(h, clen + (size_t)i * n_rec, sclen + (size_t)i * (n_rec + 1), n_rec));
    } else {
      FQ28_TRY(scan_exclusive_u32(h, delta + (size_t)i * n_rec, sdelta + (size_t)i * (n_rec + 1), n_rec));
    }
  }
  k_dtk_emit<false><<<blocks, threads, 0, h->stream>>>(n_rec, d_cr, (unsigned)n_chunks, df, d_in, d_infos, d_in + o_fs, sdelta, sflag,
                                                     sclen, clen, len32, len16, nullptr, h->d_status);
  FQ28_LAUNCH_CHECK(h);
  FQ28_TRY(scan_exclusive_u32(h, len32, len32, n_rec));
  FQ28_CUDA(h, cudaMemcpyAsync(h->h_scalars + 33, len32 + n_rec, 4, cudaMemcpyDeviceToHost, h->stream));
  FQ28_TRY(check_status(h, "header detokeniser"));
  const size_t total = (size_t)(uint32_t)h->h_scalars[33];
  if (total > headers_cap) {
    if (headers_bytes) *headers_bytes = total;  // the size to come back with
    return fail(h, FQ28_ERR_CAP, "headers need %zu bytes, buffer has %zu", total, headers_cap);
  }
  FQ28_TRY(ensure(h, h->dec_out, total + 64));
  k_dtk_emit<true><<<blocks, threads, 0, h->stream>>>(n_rec, d_cr, (unsigned)n_chunks, df, d_in, d_infos, d_in + o_fs, sdelta, sflag,
                                                    sclen, clen, len32, len16, h->dec_out.as<unsigned char>(), h->d_status);
  FQ28_LAUNCH_CHECK(h);
  FQ28_CUDA(h, cudaMemcpyAsync(headers_out, h->dec_out.p, total, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(hdr_lens_out, len16, n_rec * 2, cudaMemcpyDeviceToHost, h->stream));
  FQ28_TRY(check_status(h, "header detokeniser"));
  if (headers_bytes) *headers_bytes = total;
  return FQ28_OK;
}

}  // namespace fq28
