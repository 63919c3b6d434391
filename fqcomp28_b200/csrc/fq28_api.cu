// fq28_api.cu -- the C ABI of include/fq28.h: handle lifetime, error plumbing,
// host<->device staging for the host-buffer entry points, stage timers.
#include <stdarg.h>
#include <stdlib.h>

#include <array>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "fq28_internal.cuh"

namespace fq28 {

int fail(fq28_handle *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

int cuda_fail(fq28_handle *h, cudaError_t e, const char *what) {
  return fail(h, FQ28_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

int ensure(fq28_handle *h, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap && b.p) return FQ28_OK;
  // grow with headroom so that repeated slabs of similar size do not realloc
  size_t want = bytes + bytes / 8 + 256;
  want = (want + 255) & ~(size_t)255;
  if (b.p) {
    FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
    FQ28_CUDA(h, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  FQ28_CUDA(h, cudaMalloc(&b.p, want));
  b.cap = want;
  return FQ28_OK;
}

int ensure_pinned(fq28_handle *h, size_t bytes) {
  if (bytes <= h->h_pin_cap && h->h_pin) return FQ28_OK;
  if (h->h_pin) {
    FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
    FQ28_CUDA(h, cudaFreeHost(h->h_pin));
    h->h_pin = nullptr;
    h->h_pin_cap = 0;
  }
  const size_t want = (bytes + bytes / 4 + 4095) & ~(size_t)4095;
  FQ28_CUDA(h, cudaMallocHost(&h->h_pin, want));
  h->h_pin_cap = want;
  return FQ28_OK;
}

static const char *err_name(int code) {
  switch (code) {
    case FQ28_ERR_FORMAT: return "malformed FASTQ (4-line structure, '@'/'+' markers, qual length, or record larger than the reading size)";
    case FQ28_ERR_ALPHABET: return "symbol outside the codec alphabet (bases ACGTN, qualities '!'..'`')";
    case FQ28_ERR_SHORT: return "read shorter than 3 bases (undefined behaviour in the reference)";
    case FQ28_ERR_LONG: return "narrow_cast<>() failed: line longer than 65535";
    case FQ28_ERR_CAP: return "capacity exceeded";
    case FQ28_ERR_STREAM: return "corrupt stream (not exactly consumed / bad N positions)";
    default: return "error";
  }
}

int check_status(fq28_handle *h, const char *what, cudaStream_t on) {
  if (!on) on = h->stream;
  FQ28_CUDA(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, on));
  FQ28_CUDA(h, cudaStreamSynchronize(on));
  if (h->h_status->code != 0)
    return fail(h, h->h_status->code, "%s: %s (at index %u)", what, err_name(h->h_status->code), h->h_status->where);
  return FQ28_OK;
}

void stage_reset(fq28_handle *h) { h->ev_used = 0; }

void stage_begin(fq28_handle *h, Stage s) {
  if (!h->timing) return;
  if (h->ev_used == h->ev_pool.size()) {
    fq28_handle::EvRec r;
    r.stage = s;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    h->ev_pool.push_back(r);
  }
  h->ev_pool[h->ev_used].stage = s;
  cudaEventRecord(h->ev_pool[h->ev_used].a, h->stream);
}

void stage_end(fq28_handle *h, Stage s) {
  if (!h->timing) return;
  (void)s;
  cudaEventRecord(h->ev_pool[h->ev_used].b, h->stream);
  h->ev_used++;
}

static void stage_begin_on(fq28_handle *h, Stage s, cudaStream_t strm) {
  if (!h->timing) return;
  if (h->ev_used == h->ev_pool.size()) {
    fq28_handle::EvRec r;
    r.stage = s;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    h->ev_pool.push_back(r);
  }
  h->ev_pool[h->ev_used].stage = s;
  cudaEventRecord(h->ev_pool[h->ev_used].a, strm);
}
void side_stage_begin(fq28_handle *h, Stage s) { stage_begin_on(h, s, h->side); }
void side_stage_end(fq28_handle *h, Stage s) {
  if (!h->timing) return;
  (void)s;
  cudaEventRecord(h->ev_pool[h->ev_used].b, h->side);
  h->ev_used++;
}
int stage_open(fq28_handle *h, Stage s, cudaStream_t strm) {
  if (!h->timing) return -1;
  if (h->ev_used == h->ev_pool.size()) {
    fq28_handle::EvRec r;
    r.stage = s;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    h->ev_pool.push_back(r);
  }
  const int slot = (int)h->ev_used++;
  h->ev_pool[slot].stage = s;
  cudaEventRecord(h->ev_pool[slot].a, strm);
  return slot;
}
void stage_close(fq28_handle *h, int slot, cudaStream_t strm) {
  if (slot >= 0) cudaEventRecord(h->ev_pool[slot].b, strm);
}
int side_fork(fq28_handle *h) {
  FQ28_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
  FQ28_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  return FQ28_OK;
}
int side_join(fq28_handle *h) {
  FQ28_CUDA(h, cudaEventRecord(h->ev_join, h->side));
  FQ28_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  return FQ28_OK;
}

static void free_buf(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

static void free_tables(DevTables &t) {
  cudaFree(t.counts); cudaFree(t.norm); cudaFree(t.logs); cudaFree(t.max_log); cudaFree(t.toff);
  cudaFree(t.ctab); cudaFree(t.symtt); cudaFree(t.dtab); cudaFree(t.dtab_fix);
  cudaFree(t.wtab); cudaFree(t.qrk); cudaFree(t.qdinfo); cudaFree(t.qwin); cudaFree(t.wtabw);
  cudaFree(t.logsuf); cudaFree(t.seqdec); cudaFree(t.cid); cudaFree(t.n_touched); cudaFree(t.dom_sym); cudaFree(t.zrun); cudaFree(t.zinfo);
  t = DevTables();
}

static int bind(fq28_handle *h) {
  FQ28_CUDA(h, cudaSetDevice(h->device));
  return FQ28_OK;
}

// copies a FreqTable image out of device norm/logs/max_log
static int ft_image_out(fq28_handle *h, const DevTables &t, void *image) {
  if (!image) return FQ28_OK;
  uint8_t *p = static_cast<uint8_t *>(image);
  const size_t na = (size_t)t.n_models * t.alphabet;
  FQ28_CUDA(h, cudaMemcpyAsync(p, t.norm, na * 2, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(p + na * 2, t.logs, (size_t)t.n_models * 4, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(p + na * 2 + (size_t)t.n_models * 4, t.max_log, 4, cudaMemcpyDeviceToHost, h->stream));
  return FQ28_OK;
}

static int ft_image_in(fq28_handle *h, DevTables &t, const void *image) {
  const uint8_t *p = static_cast<const uint8_t *>(image);
  const size_t na = (size_t)t.n_models * t.alphabet;
  // validate logs on the host before any kernel indexes with them
  const uint32_t *logs = reinterpret_cast<const uint32_t *>(p + na * 2);
  for (unsigned i = 0; i < t.n_models; i++)
    if (logs[i] < FSE_MIN_TABLELOG || logs[i] > FIX_LOG) return fail(h, FQ28_ERR_ARG, "FreqTable image: bad table log %u in context %u", logs[i], i);
  FQ28_CUDA(h, cudaMemcpyAsync(t.norm, p, na * 2, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(t.logs, p + na * 2, (size_t)t.n_models * 4, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(t.max_log, p + na * 2 + (size_t)t.n_models * 4, 4, cudaMemcpyHostToDevice, h->stream));
  return FQ28_OK;
}

static int stage_in(fq28_handle *h, const char *fastq, size_t n_bytes) {
  FQ28_TRY(ensure(h, h->in_fastq, n_bytes + 64));
  if (h->stage_host && fastq >= h->stage_host && fastq + n_bytes <= h->stage_host + h->stage_bytes) {
    // already on the device (fq28_stage): move it to the aligned work buffer instead of a second H2D
    FQ28_CUDA(h, cudaMemcpyAsync(h->in_fastq.p, h->in_raw.as<char>() + (fastq - h->stage_host), n_bytes,
                                 cudaMemcpyDeviceToDevice, h->stream));
    return FQ28_OK;
  }
  FQ28_CUDA(h, cudaMemcpyAsync(h->in_fastq.p, fastq, n_bytes, cudaMemcpyHostToDevice, h->stream));
  return FQ28_OK;
}

}  // namespace fq28

using namespace fq28;

extern "C" {

int fq28_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int fq28_create(int device, fq28_handle **out) {
  if (!out) return FQ28_ERR_ARG;
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev) return FQ28_ERR_CUDA;  // no CPU fallback
  fq28_handle *h = new fq28_handle();
  h->device = device;
  int rc = FQ28_OK;
  do {
    if (cudaSetDevice(device) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    h->own_stream = true;
    {  // the side stream carries the latency-bound chains: its CTAs must not queue behind bulk kernels
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if (getenv("FQ28_SIDE_PRIO_NORMAL")) hi = 0;
      if (cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    }
    if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    if (cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    if (cudaMalloc(&h->d_status, sizeof(DevStatus)) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    if (cudaMallocHost(&h->h_status, sizeof(DevStatus)) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    if (cudaMalloc(&h->d_scalars, 64 * sizeof(uint64_t)) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    if (cudaMallocHost(&h->h_scalars, 64 * sizeof(uint64_t)) != cudaSuccess) { rc = FQ28_ERR_CUDA; break; }
    cudaMemset(h->d_status, 0, sizeof(DevStatus));
    {  // diagnostic knobs: read once here, never on the per-call paths
      auto env_u = [](const char *n) -> unsigned { const char *e = getenv(n); return e ? (unsigned)atoi(e) : 0u; };
      h->cfg.seq_v1 = getenv("FQ28_DEC_V1") != nullptr || getenv("FQ28_SEQ_V1") != nullptr;
      h->cfg.qual_v2 = getenv("FQ28_QUAL_V1") == nullptr && getenv("FQ28_DEC_V1") == nullptr;
      h->cfg.dec_serial = getenv("FQ28_DEC_SERIAL") != nullptr;
      h->cfg.share_sms = getenv("FQ28_DEC_SHARE_SMS") != nullptr;
      h->cfg.dec_concurrent = getenv("FQ28_DEC_CONCURRENT") != nullptr;
      h->cfg.no_win = getenv("FQ28_QUAL_DENSE") != nullptr;
      h->cfg.no_eager = getenv("FQ28_NO_EAGER_EXTRACT") != nullptr;
      h->cfg.force_win = getenv("FQ28_QUAL_WINDOWED") != nullptr;
      h->cfg.seq_lanes = env_u("FQ28_SEQ_LANES"); h->cfg.seq_warps = env_u("FQ28_SEQ_WARPS");
      h->cfg.qual_lanes = env_u("FQ28_QUAL_LANES"); h->cfg.qual_warps = env_u("FQ28_QUAL_WARPS");
      if (const char *e = getenv("FQ28_QUAL_CARVEOUT")) h->cfg.qual_carveout = atoi(e);
      h->cfg.no_zrun = getenv("FQ28_NO_ZRUN") != nullptr;
      h->cfg.no_dom = getenv("FQ28_NO_DOM") != nullptr;
      h->cfg.no_rankc = getenv("FQ28_NO_RANKC") != nullptr;
      h->cfg.serial = getenv("FQ28_SERIAL") != nullptr;
      h->cfg.full_overlap = getenv("FQ28_FULL_OVERLAP") != nullptr;
      if (const char *e = getenv("FQ28_PIPE_MIN_MB")) h->cfg.pipe_min_bytes = (size_t)atoll(e) << 20;
      h->cfg.pipe_trace = getenv("FQ28_PIPE_TRACE") != nullptr;
      if (const char *e = getenv("FQ28_PIPE_LANES")) h->cfg.pipe_lanes = (unsigned)std::max(1, std::min(8, atoi(e)));
      if (const char *e = getenv("FQ28_PIPE_PARTS")) h->cfg.pipe_parts = (unsigned)std::max(1, std::min(64, atoi(e)));
    }
    // kernel attributes are per device: set them for this handle's device (not once per process)
    rc = encode_init_device(h);
    if (rc) break;
    rc = decode_init_device(h);
    if (rc) break;
    rc = tables_alloc(h, h->seq, SEQ_N, SEQ_A);
    if (rc) break;
    rc = tables_alloc(h, h->qual, QUAL_N, QUAL_A);
  } while (0);
  if (rc != FQ28_OK) { fq28_destroy(h); return rc; }
  *out = h;
  return FQ28_OK;
}

void fq28_destroy(fq28_handle *h) {
  if (!h) return;
  for (fq28_handle *s : h->siblings) {
    if (s->borrowing) { s->seq = s->own_seq; s->qual = s->own_qual; s->borrowing = false; }
    fq28_destroy(s);
  }
  h->siblings.clear();
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  DevBuf *bufs[] = {&h->in_fastq, &h->tile_cnt, &h->nl, &h->hdr_off, &h->seq_off, &h->qual_off, &h->len, &h->hdr_len,
                    &h->symoff, &h->chunk_rec, &h->n_count, &h->npos_off, &h->n_pos, &h->key_seq, &h->key_qual,
                    &h->perm_seq, &h->perm_qual, &h->ssym_seq, &h->ssym_qual, &h->out_seq, &h->out_qual, &h->tile0_seq,
                    &h->tile0_qual, &h->tbase_seq, &h->tbase_qual, &h->fstate_seq, &h->fstate_qual, &h->ptile0_seq,
                    &h->ptile0_qual, &h->pbits_seq, &h->pbits_qual, &h->pscan_seq, &h->pscan_qual, &h->arena_seq,
                    &h->arena_qual, &h->d_infos, &h->scan_tmp, &h->scan_tmp_side, &h->dom_list, &h->present, &h->in_raw, &h->hdrscan, &h->hdr_arena, &h->dec_out, &h->dec_recout, &h->dec_hdrin,
                    &h->dec_npos_off, &h->dec_meta, &h->dec_cold};
  for (DevBuf *b : bufs) free_buf(*b);
  for (DevBuf &b : h->dec_in) free_buf(b);
  free_tables(h->seq);
  free_tables(h->qual);
  if (h->d_status) cudaFree(h->d_status);
  if (h->h_status) cudaFreeHost(h->h_status);
  if (h->d_scalars) cudaFree(h->d_scalars);
  if (h->h_scalars) cudaFreeHost(h->h_scalars);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  for (auto &r : h->ev_pool) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->bulk) { cudaStreamSynchronize(h->bulk); cudaStreamDestroy(h->bulk); }
  if (h->ev_parsed) cudaEventDestroy(h->ev_parsed);
  if (h->ev_extract) cudaEventDestroy(h->ev_extract);
  for (cudaEvent_t e : h->ev_part) cudaEventDestroy(e);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char *fq28_last_error(const fq28_handle *h) { return h ? h->err.c_str() : "null handle"; }

int fq28_set_stream(fq28_handle *h, void *cuda_stream) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  h->own_stream = false;
  return FQ28_OK;
}

uint64_t fq28_launch_count(const fq28_handle *h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------- parse/split
int fq28_parse(fq28_handle *h, const char *fastq, size_t n_bytes, uint32_t *hdr_off, uint32_t *seq_off,
               uint32_t *qual_off, uint16_t *hdr_len, uint16_t *len, size_t cap, size_t *n_records,
               size_t *consumed) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  FQ28_TRY(stage_in(h, fastq, n_bytes));
  stage_begin(h, ST_PARSE);
  FQ28_TRY(parse_slab(h, h->in_fastq.as<char>(), n_bytes, false));
  stage_end(h, ST_PARSE);
  FQ28_TRY(check_status(h, "parseRecords"));
  const size_t n = h->n_rec;
  if (n_records) *n_records = n;
  uint32_t end = 0;
  FQ28_CUDA(h, cudaMemcpyAsync(&end, h->hdr_off.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, h->stream));
  if ((hdr_off || seq_off || qual_off || hdr_len || len) && n > cap)
    return fail(h, FQ28_ERR_CAP, "record table needs %zu entries, cap %zu", n, cap);
  if (hdr_off) FQ28_CUDA(h, cudaMemcpyAsync(hdr_off, h->hdr_off.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
  if (seq_off) FQ28_CUDA(h, cudaMemcpyAsync(seq_off, h->seq_off.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
  if (qual_off) FQ28_CUDA(h, cudaMemcpyAsync(qual_off, h->qual_off.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
  if (hdr_len) FQ28_CUDA(h, cudaMemcpyAsync(hdr_len, h->hdr_len.p, n * 2, cudaMemcpyDeviceToHost, h->stream));
  if (len) FQ28_CUDA(h, cudaMemcpyAsync(len, h->len.p, n * 2, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  if (consumed) *consumed = end;
  return FQ28_OK;
}

int fq28_split(fq28_handle *h, const char *fastq, size_t n_bytes, size_t reading_size, int eof, uint64_t *offs,
               size_t cap, size_t *n_chunks) {
  if (!h || !offs || cap < 1) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  FQ28_TRY(stage_in(h, fastq, n_bytes));
  stage_begin(h, ST_PARSE);
  FQ28_TRY(parse_slab(h, h->in_fastq.as<char>(), n_bytes, true));
  FQ28_TRY(split_slab(h, reading_size, eof != 0, 0));
  stage_end(h, ST_PARSE);
  if (h->n_chunks + 1 > cap) return fail(h, FQ28_ERR_CAP, "offs needs %zu entries, cap %zu", h->n_chunks + 1, cap);
  for (size_t k = 0; k <= h->n_chunks; k++) offs[k] = h->h_chunk_byte[k];
  if (n_chunks) *n_chunks = h->n_chunks;
  return FQ28_OK;
}

// ---------------------------------------------------------------- tables
int fq28_hist_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes, uint32_t *d_seq_counts,
                  uint32_t *d_qual_counts) {
  if (!h || !d_seq_counts || !d_qual_counts) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  stage_begin(h, ST_PARSE);
  FQ28_TRY(parse_slab(h, d_fastq, n_bytes, false));
  stage_end(h, ST_PARSE);
  stage_begin(h, ST_HIST);
  FQ28_TRY(hist_slab(h, d_seq_counts, d_qual_counts));
  stage_end(h, ST_HIST);
  return check_status(h, "calculateFreqTable");
}

int fq28_hist(fq28_handle *h, const char *fastq, size_t n_bytes, uint32_t *seq_counts, uint32_t *qual_counts) {
  if (!h || !seq_counts || !qual_counts) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  FQ28_TRY(stage_in(h, fastq, n_bytes));
  const size_t ns = (size_t)SEQ_N * SEQ_A * 4, nq = (size_t)QUAL_N * QUAL_A * 4;
  FQ28_CUDA(h, cudaMemcpyAsync(h->seq.counts, seq_counts, ns, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(h->qual.counts, qual_counts, nq, cudaMemcpyHostToDevice, h->stream));
  FQ28_TRY(fq28_hist_dev(h, h->in_fastq.as<char>(), n_bytes, h->seq.counts, h->qual.counts));
  FQ28_CUDA(h, cudaMemcpyAsync(seq_counts, h->seq.counts, ns, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(qual_counts, h->qual.counts, nq, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  return FQ28_OK;
}

int fq28_build_tables_dev(fq28_handle *h, const uint32_t *d_seq_counts, const uint32_t *d_qual_counts, void *ft_seq_out,
                          void *ft_qual_out) {
  if (!h || !d_seq_counts || !d_qual_counts) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  stage_begin(h, ST_TABLES);
  FQ28_TRY(tables_from_counts(h, h->seq, d_seq_counts));
  FQ28_TRY(tables_from_counts(h, h->qual, d_qual_counts));
  stage_end(h, ST_TABLES);
  FQ28_TRY(ft_image_out(h, h->seq, ft_seq_out));
  FQ28_TRY(ft_image_out(h, h->qual, ft_qual_out));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  return FQ28_OK;
}

int fq28_build_tables(fq28_handle *h, const uint32_t *seq_counts, const uint32_t *qual_counts, void *ft_seq_out,
                      void *ft_qual_out) {
  if (!h || !seq_counts || !qual_counts) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  FQ28_CUDA(h, cudaMemcpyAsync(h->seq.counts, seq_counts, (size_t)SEQ_N * SEQ_A * 4, cudaMemcpyHostToDevice, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(h->qual.counts, qual_counts, (size_t)QUAL_N * QUAL_A * 4, cudaMemcpyHostToDevice, h->stream));
  return fq28_build_tables_dev(h, h->seq.counts, h->qual.counts, ft_seq_out, ft_qual_out);
}

int fq28_load_tables(fq28_handle *h, const void *ft_seq, const void *ft_qual) {
  if (!h || !ft_seq || !ft_qual) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  FQ28_TRY(ft_image_in(h, h->seq, ft_seq));
  FQ28_TRY(ft_image_in(h, h->qual, ft_qual));
  stage_begin(h, ST_TABLES);
  FQ28_TRY(tables_from_norm(h, h->seq));
  FQ28_TRY(tables_from_norm(h, h->qual));
  stage_end(h, ST_TABLES);
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  return FQ28_OK;
}

// ---------------------------------------------------------------- compress
size_t fq28_bound_seq(size_t n) {  // src/workspace.h:21-29
  if (n < 1024) return (size_t)1024 * FQ28_SEQ_MODELS;
  return n / 4 + 1024;
}
size_t fq28_bound_qual(size_t n) {  // src/workspace.h:31-35
  const size_t a = (size_t)1024 * FQ28_QUAL_MODELS, b = n * 7 / 8 + 1024;
  return a > b ? a : b;
}

// analyzeDataset (src/prepare.cpp:42-47): first chunk of a reader whose reading size is the
// sample size = records wholly inside the window; histograms, tables, FreqTable images
static int analyze_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes, size_t sample_bytes, void *ft_seq_out,
                       void *ft_qual_out) {
  const size_t win = sample_bytes < n_bytes ? sample_bytes : n_bytes;
  stage_begin(h, ST_PARSE);
  FQ28_TRY(parse_slab(h, d_fastq, win, false));
  stage_end(h, ST_PARSE);
  if (h->n_rec == 0) return fail(h, FQ28_ERR_FORMAT, "sample window of %zu bytes holds no complete record", win);
  stage_begin(h, ST_HIST);
  FQ28_CUDA(h, cudaMemsetAsync(h->seq.counts, 0, (size_t)SEQ_N * SEQ_A * 4, h->stream));
  FQ28_CUDA(h, cudaMemsetAsync(h->qual.counts, 0, (size_t)QUAL_N * QUAL_A * 4, h->stream));
  FQ28_TRY(hist_slab(h, h->seq.counts, h->qual.counts));
  stage_end(h, ST_HIST);
  stage_begin(h, ST_TABLES);
  FQ28_TRY(tables_from_counts(h, h->seq, h->seq.counts));
  FQ28_TRY(tables_from_counts(h, h->qual, h->qual.counts));
  stage_end(h, ST_TABLES);
  FQ28_TRY(check_status(h, "analyzeDataset"));
  FQ28_TRY(ft_image_out(h, h->seq, ft_seq_out));
  FQ28_TRY(ft_image_out(h, h->qual, ft_qual_out));
  return FQ28_OK;
}

int fq28_compress_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes, size_t sample_bytes, size_t reading_size,
                      int eof, void *ft_seq_out, void *ft_qual_out, fq28_chunk_info *infos, size_t infos_cap,
                      fq28_enc_summary *summary) {
  if (!h || !infos) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  if (sample_bytes > 0) FQ28_TRY(analyze_dev(h, d_fastq, n_bytes, sample_bytes, ft_seq_out, ft_qual_out));
  const bool planned = sample_bytes == 0 && h->plan.valid && h->plan.d_fastq == d_fastq && h->plan.n_bytes == n_bytes &&
                       h->plan.reading_size == reading_size && h->plan.eof == (eof != 0);
  if (!planned) {
    stage_begin(h, ST_PARSE);
    FQ28_TRY(parse_slab(h, d_fastq, n_bytes, true));
    FQ28_TRY(split_slab(h, reading_size, eof != 0, 0));
    stage_end(h, ST_PARSE);
  }
  h->plan.valid = false;
  return encode_slab(h, infos, infos_cap, summary);
}

// parseRecords + the chunk boundary walk of one slab, without encoding (src/fastq_io.cpp:23-125).
// *consumed = where the next slab must start.  The plan is kept: fq28_compress_dev with the same
// slab / reading size / eof and sample_bytes == 0 then goes straight to the encode.  This is what
// lets several GPUs share one file: slab i+1 starts at the `consumed` of slab i, which is known
// after this cheap step, long before slab i is encoded.
int fq28_preparse_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  stage_begin(h, ST_PARSE);
  FQ28_TRY(parse_slab(h, d_fastq, n_bytes, true));
  stage_end(h, ST_PARSE);
  h->parsed.valid = true;
  h->parsed.d_fastq = d_fastq;
  h->parsed.n_bytes = n_bytes;
  return extract_eager(h);   // the field separation does not need the chunk boundaries: it runs under the wait for the cut
}

int fq28_plan_cut_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes, size_t reading_size, int eof, uint64_t first_cut,
                      uint64_t *consumed, size_t *n_chunks) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  const bool have = h->parsed.valid && h->parsed.d_fastq == d_fastq && h->parsed.n_bytes == n_bytes;
  if (!have) stage_reset(h);
  stage_begin(h, ST_PARSE);
  if (!have) FQ28_TRY(parse_slab(h, d_fastq, n_bytes, true));
  h->parsed.valid = false;
  FQ28_TRY(split_slab(h, reading_size, eof != 0, 0, (size_t)first_cut));
  stage_end(h, ST_PARSE);
  h->plan.valid = true;
  h->plan.d_fastq = d_fastq; h->plan.n_bytes = n_bytes; h->plan.reading_size = reading_size; h->plan.eof = eof != 0;
  if (consumed) *consumed = h->h_chunk_byte[h->n_chunks];
  if (n_chunks) *n_chunks = h->n_chunks;
  return FQ28_OK;
}

int fq28_plan_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes, size_t reading_size, int eof, uint64_t *consumed,
                  size_t *n_chunks) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  stage_begin(h, ST_PARSE);
  FQ28_TRY(parse_slab(h, d_fastq, n_bytes, true));
  FQ28_TRY(split_slab(h, reading_size, eof != 0, 0));
  stage_end(h, ST_PARSE);
  h->plan.valid = true;
  h->plan.d_fastq = d_fastq; h->plan.n_bytes = n_bytes; h->plan.reading_size = reading_size; h->plan.eof = eof != 0;
  if (consumed) *consumed = h->h_chunk_byte[h->n_chunks];
  if (n_chunks) *n_chunks = h->n_chunks;
  return FQ28_OK;
}

int fq28_stage(fq28_handle *h, const char *fastq, size_t n_bytes) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  h->stage_host = nullptr;
  h->stage_bytes = 0;
  if (!fastq || n_bytes == 0) return FQ28_OK;  // forget the staged range
  FQ28_TRY(ensure(h, h->in_raw, n_bytes + 64));
  FQ28_CUDA(h, cudaMemcpyAsync(h->in_raw.p, fastq, n_bytes, cudaMemcpyHostToDevice, h->stream));
  h->stage_host = fastq;
  h->stage_bytes = n_bytes;
  return FQ28_OK;
}

int fq28_preparse(fq28_handle *h, const char *fastq, size_t n_bytes) {
  if (!h || !fastq) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  FQ28_TRY(stage_in(h, fastq, n_bytes));
  FQ28_TRY(fq28_preparse_dev(h, h->in_fastq.as<char>(), n_bytes));
  h->plan_host = fastq;
  return FQ28_OK;
}

int fq28_plan_cut(fq28_handle *h, const char *fastq, size_t n_bytes, size_t reading_size, int eof, uint64_t first_cut,
                  uint64_t *consumed, size_t *n_chunks) {
  if (!h || !fastq) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  const bool have = h->parsed.valid && h->parsed.d_fastq == h->in_fastq.as<char>() && h->parsed.n_bytes == n_bytes &&
                    h->plan_host == fastq;
  if (!have) FQ28_TRY(stage_in(h, fastq, n_bytes));
  FQ28_TRY(fq28_plan_cut_dev(h, h->in_fastq.as<char>(), n_bytes, reading_size, eof, first_cut, consumed, n_chunks));
  h->plan_host = fastq;
  return FQ28_OK;
}

int fq28_plan(fq28_handle *h, const char *fastq, size_t n_bytes, size_t reading_size, int eof, uint64_t *consumed,
              size_t *n_chunks) {
  if (!h || !fastq) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  FQ28_TRY(stage_in(h, fastq, n_bytes));
  FQ28_TRY(fq28_plan_dev(h, h->in_fastq.as<char>(), n_bytes, reading_size, eof, consumed, n_chunks));
  h->plan_host = fastq;
  return FQ28_OK;
}

static int fetch_async(fq28_handle *h, const fq28_enc_arenas *out) {
  if (!h->have_result) return fail(h, FQ28_ERR_ARG, "no compress result to fetch");
  FQ28_TRY(bind(h));
  const fq28_enc_summary &s = h->last_summary;
  if (s.seq_bytes > out->seq_cap || s.qual_bytes > out->qual_cap || s.n_records > out->readlens_cap ||
      s.n_records > out->n_count_cap || s.n_pos_entries > out->n_pos_cap ||
      (out->hdr_lens && s.n_records > out->hdr_lens_cap) || (out->headers && s.hdr_bytes > out->headers_cap))
    return fail(h, FQ28_ERR_CAP, "arena too small: need seq %llu qual %llu records %llu n_pos %llu",
                (unsigned long long)s.seq_bytes, (unsigned long long)s.qual_bytes, (unsigned long long)s.n_records,
                (unsigned long long)s.n_pos_entries);
  if (s.n_chunks == 0) return FQ28_OK;
  FQ28_CUDA(h, cudaMemcpyAsync(out->seq, h->arena_seq.p, s.seq_bytes, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(out->qual, h->arena_qual.p, s.qual_bytes, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(out->readlens, h->len.p, s.n_records * 2, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaMemcpyAsync(out->n_count, h->n_count.p, s.n_records * 2, cudaMemcpyDeviceToHost, h->stream));
  if (s.n_pos_entries)
    FQ28_CUDA(h, cudaMemcpyAsync(out->n_pos, h->n_pos.p, s.n_pos_entries * 2, cudaMemcpyDeviceToHost, h->stream));
  if (out->hdr_lens)
    FQ28_CUDA(h, cudaMemcpyAsync(out->hdr_lens, h->hdr_len.p, s.n_records * 2, cudaMemcpyDeviceToHost, h->stream));
  if (out->headers && s.hdr_bytes)
    FQ28_CUDA(h, cudaMemcpyAsync(out->headers, h->hdr_arena.p, s.hdr_bytes, cudaMemcpyDeviceToHost, h->stream));
  return FQ28_OK;
}

int fq28_compress_fetch(fq28_handle *h, const fq28_enc_arenas *out) {
  if (!h || !out) return FQ28_ERR_ARG;
  FQ28_TRY(fetch_async(h, out));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  return FQ28_OK;
}

// Host-buffer compress of a large slab as K overlapped parts (cfg.pipe_parts, default 8) on L lanes
// (cfg.pipe_lanes, default 4).
// All host->device copies are queued at once on a copy stream, one event per part; part k is
// the slab [cut_k, p_{k+1}) where cut_k is where the chunk walk of part k-1 stopped (eof = 0:
// only chunks whose whole window lies inside the part), so the chunks are exactly those of the
// one-pass walk (deterministic from its start offset).  L handles with their own streams and
// buffers take the parts in turn, each driven by its own host thread, so the kernels of L
// parts overlap (the chains are latency-bound on a few SMs) and run under the copies of the
// later parts; the device->host copy of a part's result runs under the kernels of the next.
// What crosses between the threads: where a part starts and how many chunks precede it (known
// after the previous part's chunk walk), and the arena offsets of its result (known after the
// previous part is encoded).
namespace {
struct PartSizes { uint64_t n_records = 0, seq_bytes = 0, qual_bytes = 0, n_pos_entries = 0, hdr_bytes = 0, n_symbols = 0; };
struct PartsShared {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<size_t> cut, cbase;      // part k starts at slab offset cut[k]; cbase[k] chunks precede it
  std::vector<PartSizes> pre;          // pre[k] = totals of parts < k
  unsigned planned = 0, sized = 0;     // cut / cbase valid up to index `planned`, pre up to `sized`
  int rc = FQ28_OK;
  std::string why;
  void abort(int code, const std::string &w) {
    std::lock_guard<std::mutex> lk(mu);
    if (rc == FQ28_OK) { rc = code; why = w; }
    cv.notify_all();
  }
};
}  // namespace

static int compress_parts(fq28_handle *h, const char *fastq, size_t n_bytes, unsigned n_parts, size_t p1, size_t sample_bytes,
                          size_t reading_size, int eof, void *ft_seq_out, void *ft_qual_out,
                          const fq28_enc_arenas *out, fq28_chunk_info *infos, size_t infos_cap,
                          fq28_enc_summary *summary) {
  const unsigned n_lanes = std::max(1u, std::min(h->cfg.pipe_lanes, n_parts));
  while (h->siblings.size() + 1 < n_lanes) {
    fq28_handle *s = nullptr;
    FQ28_TRY(fq28_create(h->device, &s));
    h->siblings.push_back(s);
    FQ28_TRY(bind(h));
  }
  auto lane_handle = [&](unsigned k) -> fq28_handle * { return (k % n_lanes) ? h->siblings[k % n_lanes - 1] : h; };
  if (!h->copy_stream) FQ28_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  while (h->ev_part.size() < n_parts) {
    cudaEvent_t e;
    FQ28_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->ev_part.push_back(e);
  }
  // part ends: p[0] = 0, p[1] = p1 (holds the sample window), the rest evenly, 16-byte aligned
  std::vector<size_t> p(n_parts + 1, 0);
  p[1] = p1;
  for (unsigned k = 2; k <= n_parts; k++) p[k] = (p1 + (n_bytes - p1) / (n_parts - 1) * (k - 1)) & ~(size_t)15;
  p[n_parts] = n_bytes;
  size_t longest = 0;
  for (unsigned k = 0; k < n_parts; k++) longest = std::max(longest, p[k + 1] - p[k]);
  // a part starts at most one window before its nominal start
  for (unsigned j = 0; j < n_lanes; j++) FQ28_TRY(ensure(lane_handle(j), lane_handle(j)->in_fastq, longest + reading_size + 64));
  FQ28_TRY(ensure(h, h->in_raw, n_bytes + 64));
  // the copies must not overtake work still queued on the buffers they overwrite
  if (!h->ev_copy) FQ28_CUDA(h, cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  FQ28_CUDA(h, cudaEventRecord(h->ev_copy, h->stream));
  FQ28_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_copy, 0));
  for (unsigned k = 0; k < n_parts; k++) {
    // part 0 starts at offset 0: straight into the aligned parse buffer
    char *dst = k == 0 ? h->in_fastq.as<char>() : h->in_raw.as<char>() + p[k];
    FQ28_CUDA(h, cudaMemcpyAsync(dst, fastq + p[k], p[k + 1] - p[k], cudaMemcpyHostToDevice, h->copy_stream));
    FQ28_CUDA(h, cudaEventRecord(h->ev_part[k], h->copy_stream));
  }
  // same device: the siblings encode with this handle's tables (complete before the first part's
  // walk is published); their own stay allocated and come back in fq28_destroy
  for (unsigned j = 1; j < n_lanes; j++) {
    fq28_handle *s = lane_handle(j);
    if (!s->borrowing) { s->own_seq = s->seq; s->own_qual = s->qual; s->borrowing = true; }
  }

  // FQ28_PIPE_TRACE: host clock (ms since the call started) at the sync points of every part, to stderr
  const bool trace = h->cfg.pipe_trace;
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::array<double, 5>> tr(n_parts);
  auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  PartsShared sh;
  sh.cut.assign(n_parts + 1, 0);
  sh.cbase.assign(n_parts + 1, 0);
  sh.pre.assign(n_parts + 1, PartSizes());
  const bool tables_first = sample_bytes > 0;

  // one part, on the handle of its lane (the caller's thread is lane 0 and runs on h)
  auto run_part = [&](unsigned k) -> int {
    fq28_handle *g = lane_handle(k);
    const bool last = k + 1 == n_parts;
    size_t cut, cbase;
    {
      std::unique_lock<std::mutex> lk(sh.mu);
      sh.cv.wait(lk, [&] { return sh.rc != FQ28_OK || sh.planned >= k; });
      if (sh.rc != FQ28_OK) return sh.rc;
      cut = sh.cut[k];
      cbase = sh.cbase[k];
    }
    FQ28_TRY(bind(g));
    const size_t len = p[k + 1] - cut;
    tr[k][0] = now_ms();
    if (trace) { cudaEventSynchronize(h->ev_part[k]); tr[k][1] = now_ms(); }
    FQ28_CUDA(g, cudaStreamWaitEvent(g->stream, h->ev_part[k], 0));
    if (k > 0) {
      // [cut, p_k) came with part k-1: in h->in_fastq if that was part 0 (still intact: h's next
      // part waits for the walks of all parts before it), else in in_raw
      const char *head = k == 1 ? h->in_fastq.as<char>() + cut : h->in_raw.as<char>() + cut;
      if (p[k] > cut) FQ28_CUDA(g, cudaMemcpyAsync(g->in_fastq.p, head, p[k] - cut, cudaMemcpyDeviceToDevice, g->stream));
      FQ28_CUDA(g, cudaMemcpyAsync(g->in_fastq.as<char>() + (p[k] - cut), h->in_raw.as<char>() + p[k], p[k + 1] - p[k],
                                   cudaMemcpyDeviceToDevice, g->stream));
    } else if (tables_first) {
      stage_reset(g);
      FQ28_TRY(analyze_dev(g, g->in_fastq.as<char>(), len, sample_bytes, ft_seq_out, ft_qual_out));
    }
    if (g != h) { g->seq = h->seq; g->qual = h->qual; }   // (pointers: cheap, and h may have rebuilt its tables)
    uint64_t consumed = 0;
    size_t nc = 0;
    FQ28_TRY(fq28_plan_dev(g, g->in_fastq.as<char>(), len, reading_size, last ? eof : 0, &consumed, &nc));
    if (cbase + nc > infos_cap) return fail(g, FQ28_ERR_CAP, "infos_cap %zu < %zu chunks", infos_cap, cbase + nc);
    tr[k][2] = now_ms();
    {
      std::lock_guard<std::mutex> lk(sh.mu);
      sh.cut[k + 1] = cut + (size_t)consumed;
      sh.cbase[k + 1] = cbase + nc;
      sh.planned = k + 1;
    }
    sh.cv.notify_all();
    fq28_enc_summary sk;
    FQ28_TRY(fq28_compress_dev(g, g->in_fastq.as<char>(), len, 0, reading_size, last ? eof : 0, nullptr, nullptr,
                               infos + cbase, infos_cap - cbase, &sk));
    tr[k][3] = now_ms();
    PartSizes base;
    {
      std::unique_lock<std::mutex> lk(sh.mu);
      sh.cv.wait(lk, [&] { return sh.rc != FQ28_OK || sh.sized >= k; });
      if (sh.rc != FQ28_OK) return sh.rc;
      base = sh.pre[k];
      PartSizes &nx = sh.pre[k + 1];
      nx.n_records = base.n_records + sk.n_records; nx.seq_bytes = base.seq_bytes + sk.seq_bytes;
      nx.qual_bytes = base.qual_bytes + sk.qual_bytes; nx.n_pos_entries = base.n_pos_entries + sk.n_pos_entries;
      nx.hdr_bytes = base.hdr_bytes + sk.hdr_bytes; nx.n_symbols = base.n_symbols + sk.n_symbols;
      sh.sized = k + 1;
    }
    sh.cv.notify_all();
    fq28_enc_arenas ob = *out;
    ob.seq += base.seq_bytes; ob.seq_cap -= std::min<uint64_t>(ob.seq_cap, base.seq_bytes);
    ob.qual += base.qual_bytes; ob.qual_cap -= std::min<uint64_t>(ob.qual_cap, base.qual_bytes);
    ob.readlens += base.n_records; ob.readlens_cap -= std::min<uint64_t>(ob.readlens_cap, base.n_records);
    ob.n_count += base.n_records; ob.n_count_cap -= std::min<uint64_t>(ob.n_count_cap, base.n_records);
    ob.n_pos += base.n_pos_entries; ob.n_pos_cap -= std::min<uint64_t>(ob.n_pos_cap, base.n_pos_entries);
    if (ob.hdr_lens) { ob.hdr_lens += base.n_records; ob.hdr_lens_cap -= std::min<uint64_t>(ob.hdr_lens_cap, base.n_records); }
    if (ob.headers) { ob.headers += base.hdr_bytes; ob.headers_cap -= std::min<uint64_t>(ob.headers_cap, base.hdr_bytes); }
    FQ28_TRY(fetch_async(g, &ob));  // device->host of this part overlaps the next parts' kernels
    tr[k][4] = now_ms();
    for (uint64_t c = 0; c < sk.n_chunks; c++) {
      fq28_chunk_info &ci = infos[cbase + c];
      ci.fastq_off += cut;
      ci.rec_off += base.n_records;
      ci.seq_off += base.seq_bytes;
      ci.qual_off += base.qual_bytes;
      ci.n_pos_off += base.n_pos_entries;
      ci.hdr_off += base.hdr_bytes;
    }
    return FQ28_OK;
  };
  auto run_lane = [&](unsigned first) {
    for (unsigned k = first; k < n_parts; k += n_lanes) {
      const int rc = run_part(k);
      if (rc != FQ28_OK) {
        sh.abort(rc, lane_handle(k)->err);
        return;
      }
    }
  };
  std::vector<std::thread> lanes;
  bool threaded = true;
  try {
    for (unsigned j = 1; j < n_lanes; j++) lanes.emplace_back([&, j] { run_lane(j); });
  } catch (...) {
    threaded = false;  // no more host threads: stop the ones that started, then the parts in order on this one
    sh.abort(FQ28_ERR_CUDA, "");
    for (std::thread &t : lanes) t.join();
    lanes.clear();
    std::lock_guard<std::mutex> lk(sh.mu);
    if (sh.planned || sh.sized) return fail(h, FQ28_ERR_CUDA, "pipelined compress: could not start the lane threads");
    sh.rc = FQ28_OK;
  }
  if (threaded) {
    run_lane(0);
    for (std::thread &t : lanes) t.join();
  } else {
    for (unsigned k = 0; k < n_parts && sh.rc == FQ28_OK; k++) {   // (each part only waits for earlier parts)
      const int rc = run_part(k);
      if (rc != FQ28_OK) sh.abort(rc, lane_handle(k)->err);
    }
  }
  // nothing may still be writing into the caller's buffers (or reading them) when this returns
  bind(h);
  cudaStreamSynchronize(h->copy_stream);
  for (unsigned j = 0; j < n_lanes; j++) {
    cudaStreamSynchronize(lane_handle(j)->stream);
    lane_handle(j)->have_result = false;  // the device-resident result is split over the parts: not fetchable again
  }
  if (sh.rc != FQ28_OK) return fail(h, sh.rc, "pipelined compress: %s", sh.why.c_str());
  if (trace) {
    for (unsigned k = 0; k < n_parts; k++)
      fprintf(stderr, "fq28 pipe part %u/%u [%zu MB]: start %.2f  copied %.2f  walked %.2f  chains done %.2f  fetch queued %.2f ms\n",
              k, n_parts, (p[k + 1] - p[k]) >> 20, tr[k][0], tr[k][1], tr[k][2], tr[k][3], tr[k][4]);
    fprintf(stderr, "fq28 pipe end %.2f ms\n", now_ms());
  }
  const PartSizes &t = sh.pre[n_parts];
  fq28_enc_summary m;
  memset(&m, 0, sizeof(m));
  m.n_chunks = sh.cbase[n_parts]; m.n_records = t.n_records; m.n_symbols = t.n_symbols;
  m.seq_bytes = t.seq_bytes; m.qual_bytes = t.qual_bytes; m.n_pos_entries = t.n_pos_entries; m.hdr_bytes = t.hdr_bytes;
  m.consumed = sh.cut[n_parts];
  if (summary) *summary = m;
  return FQ28_OK;
}

int fq28_compress(fq28_handle *h, const char *fastq, size_t n_bytes, size_t sample_bytes, size_t reading_size, int eof,
                  void *ft_seq_out, void *ft_qual_out, const fq28_enc_arenas *out, fq28_chunk_info *infos,
                  size_t infos_cap, fq28_enc_summary *summary) {
  if (!h || !infos) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  if (out) {
    // overlapped parts for large slabs: every part holds at least 4 windows, the first one the sample window
    const size_t win = sample_bytes < n_bytes ? sample_bytes : n_bytes;
    unsigned parts = h->cfg.pipe_parts;
    if (reading_size) parts = (unsigned)std::min<size_t>(parts, n_bytes / (4 * reading_size));
    // a part costs ~6 ms of latency whatever its size: at least a quarter of the threshold each (64 MB)
    parts = (unsigned)std::min<size_t>(parts, n_bytes / std::max<size_t>(h->cfg.pipe_min_bytes / 4, 1));
    const bool tables_ok = sample_bytes > 0 || (h->seq.ready && h->qual.ready);
    const bool planned = h->plan.valid && sample_bytes == 0;  // a planned slab is encoded as planned, in one piece
    if (parts >= 2 && n_bytes >= h->cfg.pipe_min_bytes && tables_ok && !planned) {
      size_t p1 = (n_bytes / parts) & ~(size_t)15;
      if (win > p1) p1 = (win + 15) & ~(size_t)15;
      // the other parts share what is left; fewer of them if the sample window took most of it
      while (parts > 2 && (n_bytes - p1) / (parts - 1) < 4 * reading_size) parts--;
      if (p1 + 4 * reading_size <= n_bytes && p1 + ((size_t)16 << 20) <= n_bytes)
        return compress_parts(h, fastq, n_bytes, parts, p1, sample_bytes, reading_size, eof, ft_seq_out, ft_qual_out, out,
                              infos, infos_cap, summary);
    }
  }
  const bool planned = sample_bytes == 0 && h->plan.valid && h->plan.d_fastq == h->in_fastq.as<char>() &&
                       h->plan.n_bytes == n_bytes && h->plan.reading_size == reading_size && h->plan.eof == (eof != 0) &&
                       h->plan_host == fastq;
  if (!planned) FQ28_TRY(stage_in(h, fastq, n_bytes));
  FQ28_TRY(fq28_compress_dev(h, h->in_fastq.as<char>(), n_bytes, sample_bytes, reading_size, eof, ft_seq_out, ft_qual_out,
                             infos, infos_cap, summary));
  // out == NULL: the result stays on the device; the caller sizes its arenas from *summary and
  // calls fq28_compress_fetch (or decodes / reads it in place through fq28_compress_dev_arenas)
  return out ? fq28_compress_fetch(h, out) : FQ28_OK;
}

int fq28_compress_dev_arenas(fq28_handle *h, fq28_dec_arenas *v) {
  if (!h || !v) return FQ28_ERR_ARG;
  if (!h->have_result) return fail(h, FQ28_ERR_ARG, "no compress result");
  memset(v, 0, sizeof(*v));
  v->seq = h->arena_seq.as<uint8_t>(); v->seq_bytes = h->last_summary.seq_bytes;
  v->qual = h->arena_qual.as<uint8_t>(); v->qual_bytes = h->last_summary.qual_bytes;
  v->readlens = h->len.as<uint16_t>();
  v->n_count = h->n_count.as<uint16_t>();
  v->n_pos = h->n_pos.as<uint16_t>(); v->n_pos_entries = h->last_summary.n_pos_entries;
  v->hdr_lens = h->hdr_len.as<uint16_t>();
  v->headers = h->hdr_arena.as<uint8_t>(); v->headers_bytes = h->last_summary.hdr_bytes;
  v->n_records = h->last_summary.n_records;
  return FQ28_OK;
}

// ---------------------------------------------------------------- decompress
int fq28_decompress_dev(fq28_handle *h, const fq28_dec_arenas *in, const fq28_chunk_info *infos, size_t n_chunks,
                        char *d_fastq_out, size_t out_cap, size_t *out_bytes) {
  if (!h || !in || !infos) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  return decode_batch(h, in, infos, n_chunks, d_fastq_out, out_cap, out_bytes);
}

int fq28_decompress(fq28_handle *h, const fq28_dec_arenas *in, const fq28_chunk_info *infos, size_t n_chunks,
                    char *fastq_out, size_t out_cap, size_t *out_bytes) {
  if (!h || !in || !infos) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  size_t total = 0;
  for (size_t k = 0; k < n_chunks; k++) total += infos[k].total;
  if (total > out_cap) return fail(h, FQ28_ERR_CAP, "output needs %zu bytes, cap %zu", total, out_cap);
  fq28_dec_arenas d = *in;
  const size_t nr = in->n_records;
  struct { const void *src; size_t bytes; const void **dst; } cp[] = {
      {in->seq, in->seq_bytes, (const void **)&d.seq},
      {in->qual, in->qual_bytes, (const void **)&d.qual},
      {in->readlens, nr * 2, (const void **)&d.readlens},
      {in->n_count, nr * 2, (const void **)&d.n_count},
      {in->n_pos, in->n_pos_entries * 2, (const void **)&d.n_pos},
      {in->hdr_lens, nr * 2, (const void **)&d.hdr_lens},
      {in->headers, in->headers_bytes, (const void **)&d.headers},
  };
  for (int i = 0; i < 7; i++) {
    FQ28_TRY(ensure(h, h->dec_in[i], cp[i].bytes + 64));
    if (cp[i].bytes)
      FQ28_CUDA(h, cudaMemcpyAsync(h->dec_in[i].p, cp[i].src, cp[i].bytes, cudaMemcpyHostToDevice, h->stream));
    *cp[i].dst = h->dec_in[i].p;
  }
  FQ28_TRY(ensure(h, h->dec_out, total + 64));
  size_t wrote = 0;
  FQ28_TRY(fq28_decompress_dev(h, &d, infos, n_chunks, h->dec_out.as<char>(), total, &wrote));
  FQ28_CUDA(h, cudaMemcpyAsync(fastq_out, h->dec_out.p, wrote, cudaMemcpyDeviceToHost, h->stream));
  FQ28_CUDA(h, cudaStreamSynchronize(h->stream));
  if (out_bytes) *out_bytes = wrote;
  return FQ28_OK;
}

// ---------------------------------------------------------------- headers
int fq28_tokenize_headers(fq28_handle *h, const uint8_t *headers, size_t headers_bytes, const uint16_t *hdr_lens,
                          size_t n_records, const uint64_t *chunk_rec, size_t n_chunks, const fq28_hdr_format *fmt,
                          uint8_t *arena, size_t arena_cap, fq28_hdr_field_info *infos, size_t *arena_bytes) {
  if (!h || !headers || !hdr_lens || !chunk_rec || !fmt || !arena || !infos) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  return tokenize_headers(h, headers, headers_bytes, hdr_lens, n_records, chunk_rec, n_chunks, fmt, arena, arena_cap, infos,
                          arena_bytes);
}

int fq28_detokenize_headers(fq28_handle *h, const uint8_t *arena, size_t arena_bytes, const fq28_hdr_field_info *infos,
                            const uint64_t *chunk_rec, size_t n_chunks, const fq28_hdr_format *fmt, uint8_t *headers_out,
                            size_t headers_cap, uint16_t *hdr_lens_out, size_t *headers_bytes) {
  if (!h || !arena || !infos || !chunk_rec || !fmt || !headers_out || !hdr_lens_out) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  stage_reset(h);
  return detokenize_headers(h, arena, arena_bytes, infos, chunk_rec, n_chunks, fmt, headers_out, headers_cap, hdr_lens_out,
                            headers_bytes);
}

// ---------------------------------------------------------------- introspection
int fq28_get_ctable(fq28_handle *h, int kind, unsigned ctx, uint16_t *state_table, int32_t *dfs, uint32_t *dnb,
                    unsigned *table_log) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  DevTables &t = kind == 0 ? h->seq : h->qual;
  if (!t.ready || ctx >= t.n_models) return fail(h, FQ28_ERR_ARG, "tables not ready / bad context");
  uint32_t lg = 0, off = 0;
  FQ28_CUDA(h, cudaMemcpy(&lg, t.logs + ctx, 4, cudaMemcpyDeviceToHost));
  FQ28_CUDA(h, cudaMemcpy(&off, t.toff + ctx, 4, cudaMemcpyDeviceToHost));
  if (table_log) *table_log = lg;
  if (state_table) FQ28_CUDA(h, cudaMemcpy(state_table, t.ctab + off, (size_t)(1u << lg) * 2, cudaMemcpyDeviceToHost));
  if (dfs || dnb) {
    std::vector<int2> tt(t.alphabet);
    FQ28_CUDA(h, cudaMemcpy(tt.data(), t.symtt + (size_t)ctx * t.alphabet, t.alphabet * sizeof(int2), cudaMemcpyDeviceToHost));
    for (unsigned s = 0; s < t.alphabet; s++) {
      if (dfs) dfs[s] = tt[s].x;
      if (dnb) dnb[s] = (uint32_t)tt[s].y;
    }
  }
  return FQ28_OK;
}

int fq28_get_dtable(fq28_handle *h, int kind, unsigned ctx, uint32_t *cells, unsigned *table_log) {
  if (!h) return FQ28_ERR_ARG;
  FQ28_TRY(bind(h));
  DevTables &t = kind == 0 ? h->seq : h->qual;
  if (!t.ready || ctx >= t.n_models) return fail(h, FQ28_ERR_ARG, "tables not ready / bad context");
  uint32_t lg = 0, off = 0;
  FQ28_CUDA(h, cudaMemcpy(&lg, t.logs + ctx, 4, cudaMemcpyDeviceToHost));
  FQ28_CUDA(h, cudaMemcpy(&off, t.toff + ctx, 4, cudaMemcpyDeviceToHost));
  if (table_log) *table_log = lg;
  if (cells) FQ28_CUDA(h, cudaMemcpy(cells, t.dtab + off, (size_t)(1u << lg) * 4, cudaMemcpyDeviceToHost));
  return FQ28_OK;
}

static const char *const k_stage_names[ST_COUNT] = {"parse", "extract", "part_seq", "part_qual", "chain_seq", "chain_qual", "pack_seq", "pack_qual",
                                                    "layout", "decode_seq", "decode_qual", "ninsert", "hist", "tables"};
const char *fq28_stage_name(size_t i) { return i < ST_COUNT ? k_stage_names[i] : ""; }

int fq28_last_timings(const fq28_handle *hc, float *ms, size_t cap, size_t *n) {
  fq28_handle *h = const_cast<fq28_handle *>(hc);
  if (!h || !ms) return FQ28_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (int s = 0; s < ST_COUNT; s++) h->stage_ms[s] = 0.f;
  for (size_t i = 0; i < h->ev_used; i++) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->ev_pool[i].a, h->ev_pool[i].b) == cudaSuccess) h->stage_ms[h->ev_pool[i].stage] += t;
  }
  const size_t m = cap < (size_t)ST_COUNT ? cap : (size_t)ST_COUNT;
  for (size_t i = 0; i < m; i++) ms[i] = h->stage_ms[i];
  if (n) *n = m;
  return FQ28_OK;
}

}  // extern "C"
