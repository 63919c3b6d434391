// fqcomp28_gpu.hpp -- host-side mirror of the reference's codec interface on top
// of the fq28 C ABI (include/fq28.h).  Same namespace, class names, argument
// meaning and error behaviour as the reference, so that src/process.cpp and the
// reference's unit tests compile against it unchanged:
//
//   FastqRecord / FastqChunk                      src/defs.h:22-52
//   FreqTable<N,A>                                src/fse_common.hpp:147-174
//   FSE_Sequence / SequenceEncoder / Decoder      src/fse_sequence.h:9-111
//   FSE_Quality  / QualityEncoder  / Decoder      src/fse_quality.h:12-84
//   CompressedBuffersDst / Src                    src/compressed_buffers.h:10-101
//   FastqReader::parseRecords                     src/fastq_io.h:26
//   Workspace::compressBoundSequence/Quality      src/workspace.h:21-35
//   CompressionWorkspace::encodeChunk             src/workspace.h:69
//   DecompressionWorkspace::decodeChunk           src/workspace.h:112
//
// Everything that touches bases or qualities runs on the GPU through
// libfq28.so.  Header tokenisation and libbsc stay host-side and out of this
// path (north_star); encodeChunk/decodeChunk therefore carry the raw header
// lines in CompressedBuffers::raw_headers instead of tokenised field streams.
//
// The per-record methods (encodeRecord / decodeRecord) only queue the record;
// the kernels run once per chunk, in endChunk().  The chunk-level methods and
// their batched forms (encodeChunks / decodeChunks) are the throughput path.
#pragma once

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/fq28.h"

namespace fqcomp28 {

using readlen_t = uint16_t;  // src/defs.h:14
using FastqData = std::vector<char>;

/** Non-owning - holds pointers into outside allocated data (src/defs.h:22-32) */
struct FastqRecord {
  char *seqp = nullptr, *qualp = nullptr, *headerp = nullptr;
  readlen_t length = 0, header_length = 0;
  [[nodiscard]] auto header() const { return std::string_view(headerp, header_length); }
  [[nodiscard]] auto seq() const { return std::string_view(seqp, length); }
  [[nodiscard]] auto qual() const { return std::string_view(qualp, length); }
};

struct FastqChunk {  // src/defs.h:34-52
  FastqData raw_data;
  std::vector<FastqRecord> records;
  std::size_t tot_reads_length = 0;
  std::size_t headers_length = 0;
  unsigned idx = 0;
  void clear() {
    idx = 0;
    tot_reads_length = 0;
    headers_length = 0;
    raw_data.clear();
    records.clear();
  }
};

/** (c) The C++ Programming Language, section 11.5 -- src/utils.h:17-23 */
template <class Target, class Source> Target narrow_cast(Source v) {
  auto r = static_cast<Target>(v);
  if (static_cast<Source>(r) != v) throw std::runtime_error("narrow_cast<>() failed");
  return r;
}

/** FreqTable POD, byte-identical to src/fse_common.hpp:147-174 */
template <unsigned N_MODELS_, unsigned ALPHABET_SIZE_> struct FreqTable {
  const static unsigned N_MODELS = N_MODELS_;
  const static unsigned ALPHABET_SIZE = ALPHABET_SIZE_;
  const static unsigned MAX_SYMBOL = ALPHABET_SIZE - 1;
  template <typename T> using fse_array = std::array<T, N_MODELS>;
  fse_array<std::array<short, ALPHABET_SIZE>> norm_counts;
  fse_array<unsigned> logs;
  unsigned max_log;
  bool operator==(const FreqTable &other) const = default;
};
using SeqFreqTable = FreqTable<FQ28_SEQ_MODELS, FQ28_SEQ_ALPHABET>;
using QualFreqTable = FreqTable<FQ28_QUAL_MODELS, FQ28_QUAL_ALPHABET>;
static_assert(sizeof(SeqFreqTable) == FQ28_FT_SEQ_BYTES);
static_assert(sizeof(QualFreqTable) == FQ28_FT_QUAL_BYTES);

/** src/compressed_buffers.h:10-33 (header-field sizes dropped: out of path) */
struct cb_original_sizes_t {
  uint32_t total = 0, readlens = 0, n_records = 0, n_count = 0, n_pos = 0;
  bool operator==(const cb_original_sizes_t &) const = default;
  void clear() { total = readlens = n_records = n_count = n_pos = 0; }
};

/** src/compressed_buffers.h:35-69 */
struct CompressedBuffers {
protected:
  CompressedBuffers() = default;
  virtual ~CompressedBuffers() = default;

public:
  std::vector<std::byte> seq, qual;
  std::vector<std::byte> readlens;
  std::vector<std::byte> n_count;
  std::vector<std::byte> n_pos;
  /** raw header lines ('@'.., no '\n') + their u16 lengths: what the host
   * header tokeniser (out of path) consumes / produces */
  std::vector<std::byte> raw_headers;
  std::vector<readlen_t> header_lengths;
  cb_original_sizes_t original_size;
  uint32_t chunk_idx = 0;

  /** NB: like the reference (src/compressed_buffers.h:58-68) this does NOT
   * clear n_count / n_pos: they accumulate across chunks. */
  virtual void clear() {
    seq.clear();
    qual.clear();
    readlens.clear();
    raw_headers.clear();
    header_lengths.clear();
    original_size.clear();
  }
};
class CompressedBuffersDst : public CompressedBuffers {};
class CompressedBuffersSrc : public CompressedBuffers {
public:
  struct {
    std::size_t n_count = 0;
    std::size_t n_pos = 0;
  } index;  // back-cursors, src/compressed_buffers.h:90-93
  void clear() override {
    CompressedBuffers::clear();
    index = {};
  }
};

// ---------------------------------------------------------------------------
// GPU context: one fq28 handle, shared by the encoder/decoder objects of one
// worker (the reference builds one set of tables per worker thread,
// src/workspace.h:62-66).
// ---------------------------------------------------------------------------
class GpuContext {
public:
  explicit GpuContext(int device = 0) {
    if (fq28_create(device, &h_) != FQ28_OK)
      throw std::runtime_error("fq28_create failed: no usable CUDA device (fqcomp28 GPU path has no CPU fallback)");
  }
  ~GpuContext() { fq28_destroy(h_); }
  GpuContext(const GpuContext &) = delete;
  GpuContext &operator=(const GpuContext &) = delete;
  fq28_handle *handle() const { return h_; }

  /** maps C-ABI status onto the reference's exception types */
  void check(int rc) const {
    if (rc == FQ28_OK) return;
    const std::string msg = fq28_last_error(h_);
    if (rc == FQ28_ERR_LONG) throw std::runtime_error("narrow_cast<>() failed");  // src/utils.h:21
    if (rc == FQ28_ERR_FORMAT || rc == FQ28_ERR_ALPHABET || rc == FQ28_ERR_SHORT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }

  /** process-wide context of one device (device 0 unless told otherwise) */
  static std::shared_ptr<GpuContext> shared(int device = 0) {
    static std::mutex mu;
    static std::map<int, std::shared_ptr<GpuContext>> ctxs;
    std::lock_guard<std::mutex> lk(mu);
    auto &c = ctxs[device];
    if (!c) c = std::make_shared<GpuContext>(device);
    return c;
  }

  /** The handle holds ONE set of tables and several facade objects share it (encoders, decoders,
   * workspaces, calculateFreqTable).  Whoever changes the tables names itself as the owner; a
   * workspace re-loads its DatasetMeta's tables before a batch whenever somebody else did. */
  void loadTables(const void *owner, const void *ft_seq, const void *ft_qual) {
    if (owner != nullptr && owner == tables_owner_) return;
    tables_owner_ = nullptr;
    check(fq28_load_tables(h_, ft_seq, ft_qual));
    tables_owner_ = owner;
  }
  void tablesChanged() { tables_owner_ = nullptr; }

  /** Host staging reused from batch to batch (a context serves one worker at a time).  Fresh
   * gigabyte-sized vectors per batch cost more in zero-fill and page faults than the codec. */
  struct Scratch {
    std::vector<uint8_t> seq, qual, headers;
    std::vector<uint16_t> readlens, n_count, n_pos, hdr_lens;
    std::vector<fq28_chunk_info> infos;
    std::vector<char> out;
  };
  Scratch &scratch() { return scratch_; }

private:
  fq28_handle *h_ = nullptr;
  const void *tables_owner_ = nullptr;
  Scratch scratch_;
};

namespace detail {
inline std::size_t alignUp(std::size_t v, std::size_t a) { return (v + a - 1) / a * a; }

/** a valid placeholder table for the stream type an encoder does not own:
 * every context uniform over the alphabet */
template <class FT> inline void fillUniform(FT &ft) {
  unsigned log = 0;
  while ((1u << log) < FT::ALPHABET_SIZE) ++log;
  if (log < 5) log = 5;  // FSE_MIN_TABLELOG
  for (unsigned c = 0; c < FT::N_MODELS; ++c) {
    ft.norm_counts[c].fill(static_cast<short>((1u << log) / FT::ALPHABET_SIZE));
    ft.logs[c] = log;
  }
  ft.max_log = log;
}

template <typename T> inline void appendBytes(std::vector<std::byte> &dst, const T *src, std::size_t n) {
  const auto *p = reinterpret_cast<const std::byte *>(src);
  dst.insert(dst.end(), p, p + n * sizeof(T));
}
}  // namespace detail

// ---------------------------------------------------------------------------
// FastqReader::parseRecords (src/fastq_io.cpp:67-125) on the GPU
// ---------------------------------------------------------------------------
class FastqReader {
public:
  /** sets pointers in chunk.records; @return position at which a partially
   * loaded read starts */
  static std::size_t parseRecords(FastqChunk &chunk, GpuContext &ctx) {
    const std::size_t n = chunk.raw_data.size();
    std::size_t n_rec = 0, consumed = 0;
    ctx.check(fq28_parse(ctx.handle(), chunk.raw_data.data(), n, nullptr, nullptr, nullptr, nullptr, nullptr, 0, &n_rec, &consumed));
    std::vector<uint32_t> ho(n_rec), so(n_rec), qo(n_rec);
    std::vector<uint16_t> hl(n_rec), ln(n_rec);
    ctx.check(fq28_parse(ctx.handle(), chunk.raw_data.data(), n, ho.data(), so.data(), qo.data(), hl.data(), ln.data(), n_rec, &n_rec, &consumed));
    chunk.records.resize(n_rec);
    char *base = chunk.raw_data.data();
    for (std::size_t i = 0; i < n_rec; ++i) {
      auto &r = chunk.records[i];
      r.headerp = base + ho[i];
      r.seqp = base + so[i];
      r.qualp = base + qo[i];
      r.header_length = hl[i];
      r.length = ln[i];
      chunk.headers_length += hl[i];
      chunk.tot_reads_length += ln[i];
    }
    return consumed;
  }
  static std::size_t parseRecords(FastqChunk &chunk) { return parseRecords(chunk, *GpuContext::shared()); }
};

// ---------------------------------------------------------------------------
// frequency tables
// ---------------------------------------------------------------------------
namespace detail {
inline void histChunk(GpuContext &ctx, const FastqChunk &chunk, std::vector<uint32_t> &cs, std::vector<uint32_t> &cq) {
  cs.assign(FQ28_SEQ_MODELS * FQ28_SEQ_ALPHABET, 0);
  cq.assign(static_cast<std::size_t>(FQ28_QUAL_MODELS) * FQ28_QUAL_ALPHABET, 0);
  // records of a parsed chunk tile raw_data; histogram over the slab
  std::size_t used = chunk.raw_data.size();
  if (!chunk.records.empty()) {
    const auto &last = chunk.records.back();
    used = static_cast<std::size_t>(last.qualp + last.length + 1 - chunk.raw_data.data());
  }
  ctx.check(fq28_hist(ctx.handle(), chunk.raw_data.data(), used, cs.data(), cq.data()));
}
}  // namespace detail

class FSE_Sequence {
public:
  constexpr static unsigned CONTEXT_SIZE = 4;
  constexpr static unsigned N_MODELS = FQ28_SEQ_MODELS;
  constexpr static unsigned ALPHABET_SIZE = FQ28_SEQ_ALPHABET;
  using FreqTableT = SeqFreqTable;
  /** src/fse_sequence.cpp:145-169 */
  static std::unique_ptr<FreqTableT> calculateFreqTable(const FastqChunk &chunk, GpuContext &ctx) {
    std::vector<uint32_t> cs, cq;
    detail::histChunk(ctx, chunk, cs, cq);
    auto ft = std::make_unique<FreqTableT>();
    auto fq = std::make_unique<QualFreqTable>();
    ctx.tablesChanged();
    ctx.check(fq28_build_tables(ctx.handle(), cs.data(), cq.data(), ft.get(), fq.get()));
    return ft;
  }
  static std::unique_ptr<FreqTableT> calculateFreqTable(const FastqChunk &chunk) {
    return calculateFreqTable(chunk, *GpuContext::shared());
  }
};

class FSE_Quality {
public:
  constexpr static int CONTEXT_SIZE = 3;
  constexpr static int N_MODELS = FQ28_QUAL_MODELS;
  constexpr static int ALPHABET_SIZE = FQ28_QUAL_ALPHABET;
  using FreqTableT = QualFreqTable;
  /** src/fse_quality.cpp:69-97 */
  static std::unique_ptr<FreqTableT> calculateFreqTable(const FastqChunk &chunk, GpuContext &ctx) {
    std::vector<uint32_t> cs, cq;
    detail::histChunk(ctx, chunk, cs, cq);
    auto fs = std::make_unique<SeqFreqTable>();
    auto ft = std::make_unique<FreqTableT>();
    ctx.tablesChanged();
    ctx.check(fq28_build_tables(ctx.handle(), cs.data(), cq.data(), fs.get(), ft.get()));
    return ft;
  }
  static std::unique_ptr<FreqTableT> calculateFreqTable(const FastqChunk &chunk) {
    return calculateFreqTable(chunk, *GpuContext::shared());
  }
};

// ---------------------------------------------------------------------------
// DatasetMeta (src/prepare.h:14-53), minus the header format (out of path)
// ---------------------------------------------------------------------------
struct DatasetMeta {
  DatasetMeta() = default;
  explicit DatasetMeta(const FastqChunk &chunk, GpuContext &ctx = *GpuContext::shared())
      : first_header(chunk.records.front().header()) {
    std::vector<uint32_t> cs, cq;
    detail::histChunk(ctx, chunk, cs, cq);
    ft_seq = std::make_unique<SeqFreqTable>();
    ft_qual = std::make_unique<QualFreqTable>();
    ctx.tablesChanged();
    ctx.check(fq28_build_tables(ctx.handle(), cs.data(), cq.data(), ft_seq.get(), ft_qual.get()));
  }
  std::string first_header;
  std::unique_ptr<SeqFreqTable> ft_seq;
  std::unique_ptr<QualFreqTable> ft_qual;

  /** src/prepare.cpp:12-21: u16 hlen | header | raw ft_seq | raw ft_qual */
  static void storeToBytes(const DatasetMeta &m, std::vector<char> &out) {
    const auto hlen = narrow_cast<readlen_t>(m.first_header.size());
    out.insert(out.end(), reinterpret_cast<const char *>(&hlen), reinterpret_cast<const char *>(&hlen) + sizeof(hlen));
    out.insert(out.end(), m.first_header.begin(), m.first_header.end());
    out.insert(out.end(), reinterpret_cast<const char *>(m.ft_seq.get()), reinterpret_cast<const char *>(m.ft_seq.get()) + sizeof(SeqFreqTable));
    out.insert(out.end(), reinterpret_cast<const char *>(m.ft_qual.get()), reinterpret_cast<const char *>(m.ft_qual.get()) + sizeof(QualFreqTable));
  }
  /** src/prepare.cpp:23-40 */
  static DatasetMeta loadFromBytes(const char *p, std::size_t n) {
    DatasetMeta m;
    readlen_t hlen;
    if (n < sizeof(hlen)) throw std::runtime_error("meta truncated");
    std::memcpy(&hlen, p, sizeof(hlen));
    if (n < sizeof(hlen) + hlen + sizeof(SeqFreqTable) + sizeof(QualFreqTable)) throw std::runtime_error("meta truncated");
    m.first_header.assign(p + sizeof(hlen), hlen);
    m.ft_seq = std::make_unique<SeqFreqTable>();
    m.ft_qual = std::make_unique<QualFreqTable>();
    std::memcpy(m.ft_seq.get(), p + sizeof(hlen) + hlen, sizeof(SeqFreqTable));
    std::memcpy(m.ft_qual.get(), p + sizeof(hlen) + hlen + sizeof(SeqFreqTable), sizeof(QualFreqTable));
    return m;
  }
};
inline bool operator==(const DatasetMeta &l, const DatasetMeta &r) {
  return l.first_header == r.first_header && *l.ft_seq == *r.ft_seq && *l.ft_qual == *r.ft_qual;
}

// ---------------------------------------------------------------------------
// Workspace bounds, src/workspace.h:21-35
// ---------------------------------------------------------------------------
class Workspace {
public:
  static std::size_t compressBoundSequence(std::size_t original_size) { return fq28_bound_seq(original_size); }
  static std::size_t compressBoundQuality(std::size_t original_size) { return fq28_bound_qual(original_size); }
};

// ---------------------------------------------------------------------------
// chunk-level codec: the throughput path
// ---------------------------------------------------------------------------
namespace detail {
/** splits `fastq` by the reading-size rule and encodes every chunk; appends one
 * CompressedBuffersDst per chunk.  accumulate_ns replicates SURVEY Q2. */
inline void encodeSlab(GpuContext &ctx, const char *fastq, std::size_t n, std::size_t reading_size, bool eof,
                       std::vector<CompressedBuffersDst> &out, std::size_t *consumed, uint32_t first_idx = 0) {
  const std::size_t max_chunks = 2 * (n / std::max<std::size_t>(1, reading_size)) + 8;
  GpuContext::Scratch &ar = ctx.scratch();
  auto &infos = ar.infos;
  if (infos.size() < max_chunks) infos.resize(max_chunks);
  fq28_enc_summary summ{};
  // encode first (the result stays on the device), then fetch into arenas of exactly the sizes
  // the summary names, kept from slab to slab
  ctx.check(fq28_compress(ctx.handle(), fastq, n, 0, reading_size, eof ? 1 : 0, nullptr, nullptr, nullptr, infos.data(), infos.size(), &summ));
  auto grow = [](auto &v, std::size_t need) { if (v.size() < need) v.resize(need + need / 8 + 64); };
  grow(ar.seq, summ.seq_bytes);
  grow(ar.qual, summ.qual_bytes);
  grow(ar.readlens, summ.n_records);
  grow(ar.n_count, summ.n_records);
  grow(ar.hdr_lens, summ.n_records);
  grow(ar.n_pos, summ.n_pos_entries);
  grow(ar.headers, summ.hdr_bytes);
  if (summ.n_chunks) {
    fq28_enc_arenas view{};
    view.seq = ar.seq.data(); view.seq_cap = ar.seq.size();
    view.qual = ar.qual.data(); view.qual_cap = ar.qual.size();
    view.readlens = ar.readlens.data(); view.readlens_cap = ar.readlens.size();
    view.n_count = ar.n_count.data(); view.n_count_cap = ar.n_count.size();
    view.n_pos = ar.n_pos.data(); view.n_pos_cap = ar.n_pos.size();
    view.hdr_lens = ar.hdr_lens.data(); view.hdr_lens_cap = ar.hdr_lens.size();
    view.headers = ar.headers.data(); view.headers_cap = ar.headers.size();
    ctx.check(fq28_compress_fetch(ctx.handle(), &view));
  }
  for (std::size_t k = 0; k < summ.n_chunks; ++k) {
    const auto &ci = infos[k];
    out.emplace_back();
    auto &cb = out.back();
    cb.chunk_idx = first_idx + static_cast<uint32_t>(k);
    appendBytes(cb.seq, ar.seq.data() + ci.seq_off, ci.seq_len);
    appendBytes(cb.qual, ar.qual.data() + ci.qual_off, ci.qual_len);
    appendBytes(cb.readlens, ar.readlens.data() + ci.rec_off, ci.n_records);
    appendBytes(cb.n_count, ar.n_count.data() + ci.rec_off, ci.n_records);
    appendBytes(cb.n_pos, ar.n_pos.data() + ci.n_pos_off, ci.n_pos_len);
    appendBytes(cb.raw_headers, ar.headers.data() + ci.hdr_off, ci.hdr_bytes);
    cb.header_lengths.assign(ar.hdr_lens.begin() + static_cast<std::ptrdiff_t>(ci.rec_off),
                             ar.hdr_lens.begin() + static_cast<std::ptrdiff_t>(ci.rec_off + ci.n_records));
    cb.original_size.n_records = narrow_cast<uint32_t>(static_cast<std::size_t>(ci.n_records));
    cb.original_size.total = ci.total;
    cb.original_size.readlens = narrow_cast<uint32_t>(cb.readlens.size());
    cb.original_size.n_count = narrow_cast<uint32_t>(cb.n_count.size());
    cb.original_size.n_pos = narrow_cast<uint32_t>(cb.n_pos.size());
  }
  if (consumed) *consumed = summ.consumed;
}
}  // namespace detail

class CompressionWorkspace : public Workspace {
public:
  explicit CompressionWorkspace(const DatasetMeta *meta, std::shared_ptr<GpuContext> ctx = GpuContext::shared())
      : meta_(meta), ctx_(std::move(ctx)) {
    useTables();
  }
  void useTables() { ctx_->loadTables(this, meta_->ft_seq.get(), meta_->ft_qual.get()); }

  /** src/workspace.cpp:14-45.  chunk.raw_data must be the chunk's FASTQ bytes
   * (what FastqReader::readNextChunk leaves there).  Like the reference, the
   * N data appended to cbs.n_count / cbs.n_pos accumulates across calls on the
   * same cbs, and N bases in chunk.raw_data are replaced by 'A'
   * (src/fse_sequence.cpp:45). */
  void encodeChunk(FastqChunk &chunk, CompressedBuffersDst &cbs) {
    useTables();
    std::vector<std::byte> keep_nc = std::move(cbs.n_count), keep_np = std::move(cbs.n_pos);
    std::vector<CompressedBuffersDst> one;
    detail::encodeSlab(*ctx_, chunk.raw_data.data(), chunk.raw_data.size(), std::max<std::size_t>(chunk.raw_data.size(), 1), true,
                       one, nullptr, chunk.idx);
    if (one.size() != 1) throw std::invalid_argument("encodeChunk: chunk does not hold whole records");
    cbs.clear();
    cbs.chunk_idx = chunk.idx;
    cbs.seq = std::move(one[0].seq);
    cbs.qual = std::move(one[0].qual);
    cbs.readlens = std::move(one[0].readlens);
    cbs.raw_headers = std::move(one[0].raw_headers);
    cbs.header_lengths = std::move(one[0].header_lengths);
    keep_nc.insert(keep_nc.end(), one[0].n_count.begin(), one[0].n_count.end());
    keep_np.insert(keep_np.end(), one[0].n_pos.begin(), one[0].n_pos.end());
    cbs.n_count = std::move(keep_nc);
    cbs.n_pos = std::move(keep_np);
    cbs.original_size = one[0].original_size;
    cbs.original_size.n_count = narrow_cast<uint32_t>(cbs.n_count.size());
    cbs.original_size.n_pos = narrow_cast<uint32_t>(cbs.n_pos.size());
    for (auto &r : chunk.records)
      std::replace(r.seqp, r.seqp + r.length, 'N', 'A');
  }

  /** batched form: all chunks of a slab in one GPU pass (thousands in flight) */
  void encodeChunks(const char *fastq, std::size_t n, std::size_t reading_size, bool eof,
                    std::vector<CompressedBuffersDst> &out, std::size_t *consumed = nullptr) {
    useTables();
    detail::encodeSlab(*ctx_, fastq, n, reading_size, eof, out, consumed);
  }

private:
  const DatasetMeta *const meta_;
  std::shared_ptr<GpuContext> ctx_;
};

class DecompressionWorkspace : public Workspace {
public:
  explicit DecompressionWorkspace(const DatasetMeta *meta, std::shared_ptr<GpuContext> ctx = GpuContext::shared())
      : meta_(meta), ctx_(std::move(ctx)) {
    useTables();
  }
  void useTables() { ctx_->loadTables(this, meta_->ft_seq.get(), meta_->ft_qual.get()); }

  /** src/workspace.cpp:47-88: resizes chunk, lays out records, decodes */
  void decodeChunk(FastqChunk &chunk, CompressedBuffersSrc &cbs) {
    std::vector<CompressedBuffersSrc *> one{&cbs};
    std::vector<FastqChunk *> outs{&chunk};
    decodeChunks(one, outs);
  }

  /** batched form */
  void decodeChunks(const std::vector<CompressedBuffersSrc *> &cbs, const std::vector<FastqChunk *> &chunks) {
    GpuContext::Scratch &sc = ctx_->scratch();
    const std::size_t n = cbs.size();
    const std::size_t total = stageBatch(cbs);
    if (sc.out.size() < total) sc.out.resize(total + total / 8);
    decodeStaged(n, sc.out.data(), total);
    const auto &infos = sc.infos;
    std::size_t off = 0, rec = 0;
    for (std::size_t k = 0; k < n; ++k) {
      auto &chunk = *chunks[k];
      chunk.clear();  // prepareFastqChunk, src/workspace.h:127-133
      chunk.idx = cbs[k]->chunk_idx;
      chunk.raw_data.assign(sc.out.begin() + static_cast<std::ptrdiff_t>(off), sc.out.begin() + static_cast<std::ptrdiff_t>(off + infos[k].total));
      chunk.records.resize(infos[k].n_records);
      char *dst = chunk.raw_data.data();
      for (std::size_t i = 0; i < infos[k].n_records; ++i, ++rec) {
        auto &r = chunk.records[i];
        r.headerp = dst;
        r.header_length = sc.hdr_lens[rec];
        dst += r.header_length + 1;
        r.seqp = dst;
        r.length = sc.readlens[rec];
        dst += r.length + 3;
        r.qualp = dst;
        dst += r.length + 1;
        chunk.tot_reads_length += r.length;
        chunk.headers_length += r.header_length;
      }
      off += infos[k].total;
    }
  }

  /** batched form without FastqChunk objects: the FASTQ bytes of all chunks, in order, straight
   * into `out` (resized to the total).  What a writer needs (FastqWriter::writeChunk,
   * src/fastq_io.cpp:131-143, writes raw_data and nothing else). */
  void decodeChunksRaw(const std::vector<CompressedBuffersSrc *> &cbs, std::vector<char> &out) {
    const std::size_t total = stageBatch(cbs);
    if (out.size() < total) out.resize(total);   // (a pooled buffer only ever grows: no zero-fill after its first use)
    decodeStaged(cbs.size(), out.data(), total);
    out.resize(total);
  }

private:
  /** gathers the blocks' buffers into the context's staging arenas + chunk infos; @return FASTQ bytes */
  std::size_t stageBatch(const std::vector<CompressedBuffersSrc *> &cbs) {
    useTables();
    GpuContext::Scratch &sc = ctx_->scratch();
    const std::size_t n = cbs.size();
    auto &seq = sc.seq; auto &qual = sc.qual; auto &headers = sc.headers;
    auto &readlens = sc.readlens; auto &n_count = sc.n_count; auto &n_pos = sc.n_pos; auto &hdr_lens = sc.hdr_lens;
    seq.clear(); qual.clear(); headers.clear(); readlens.clear(); n_count.clear(); n_pos.clear(); hdr_lens.clear();
    auto &infos = sc.infos;
    infos.assign(n, fq28_chunk_info{});
    std::size_t total = 0;
    for (std::size_t k = 0; k < n; ++k) {
      auto &c = *cbs[k];
      auto &ci = infos[k];
      const std::size_t nrec = c.original_size.n_records;
      ci.total = c.original_size.total;
      ci.n_records = c.original_size.n_records;
      ci.rec_off = readlens.size();
      ci.seq_off = seq.size();
      ci.seq_len = narrow_cast<uint32_t>(c.seq.size());
      ci.qual_off = qual.size();
      ci.qual_len = narrow_cast<uint32_t>(c.qual.size());
      auto app8 = [](std::vector<uint8_t> &d, const std::vector<std::byte> &s) {
        const auto *p = reinterpret_cast<const uint8_t *>(s.data());
        d.insert(d.end(), p, p + s.size());
      };
      app8(seq, c.seq);
      app8(qual, c.qual);
      app8(headers, c.raw_headers);
      // side information comes from an archive: check the sizes before anything indexes with them
      if (c.readlens.size() < nrec * sizeof(uint16_t)) throw std::runtime_error("readlens shorter than n_records");
      if (c.header_lengths.size() != nrec) throw std::runtime_error("header_lengths does not match n_records");
      const auto *rl = reinterpret_cast<const uint16_t *>(c.readlens.data());
      readlens.insert(readlens.end(), rl, rl + nrec);
      hdr_lens.insert(hdr_lens.end(), c.header_lengths.begin(), c.header_lengths.end());
      // the reference consumes n_count / n_pos from the BACK of the (possibly
      // accumulated) buffers: this chunk owns the last nrec counts
      // (src/fse_sequence.cpp:115-126, src/workspace.cpp:221,226)
      if (c.n_count.size() < nrec * sizeof(uint16_t)) throw std::runtime_error("n_count shorter than n_records");
      const auto *nc_all = reinterpret_cast<const uint16_t *>(c.n_count.data());
      const std::size_t nc_total = c.n_count.size() / sizeof(uint16_t);
      const uint16_t *nc = nc_all + (nc_total - nrec);
      std::size_t my_n = 0;
      for (std::size_t i = 0; i < nrec; ++i) my_n += nc[i];
      const auto *np_all = reinterpret_cast<const uint16_t *>(c.n_pos.data());
      const std::size_t np_total = c.n_pos.size() / sizeof(uint16_t);
      if (np_total < my_n) throw std::runtime_error("n_pos shorter than the sum of n_count");
      ci.n_pos_off = n_pos.size();
      ci.n_pos_len = narrow_cast<uint32_t>(my_n);
      n_count.insert(n_count.end(), nc, nc + nrec);
      n_pos.insert(n_pos.end(), np_all + (np_total - my_n), np_all + np_total);
      c.index.n_count = (nc_total - nrec) * sizeof(uint16_t);
      c.index.n_pos = (np_total - my_n) * sizeof(uint16_t);
      total += ci.total;
    }
    seq.resize(seq.size() + 16);
    qual.resize(qual.size() + 16);
    return total;
  }

  void decodeStaged(std::size_t n, char *out, std::size_t total) {
    GpuContext::Scratch &sc = ctx_->scratch();
    fq28_dec_arenas in{};
    in.seq = sc.seq.data(); in.seq_bytes = sc.seq.size();
    in.qual = sc.qual.data(); in.qual_bytes = sc.qual.size();
    in.readlens = sc.readlens.data();
    in.n_count = sc.n_count.data();
    in.n_pos = sc.n_pos.data(); in.n_pos_entries = sc.n_pos.size();
    in.hdr_lens = sc.hdr_lens.data();
    in.headers = sc.headers.data(); in.headers_bytes = sc.headers.size();
    in.n_records = sc.readlens.size();
    std::size_t wrote = 0;
    ctx_->check(fq28_decompress(ctx_->handle(), &in, sc.infos.data(), n, out, total, &wrote));
    if (wrote != total) throw std::runtime_error("decoded size differs from the blocks' original sizes");
  }

private:
  const DatasetMeta *const meta_;
  std::shared_ptr<GpuContext> ctx_;
};

// ---------------------------------------------------------------------------
// per-record encoders / decoders (reference unit-test API).  Records are
// queued; the GPU runs once per chunk in endChunk() / at the first
// decodeRecord().
// ---------------------------------------------------------------------------
namespace detail {
/** builds a FASTQ slab "@\n<seq>\n+\n<qual>\n" from queued records */
struct RecordQueue {
  std::vector<FastqRecord *> recs;
  std::vector<char> slab;
  void build(bool need_seq, bool need_qual) {
    slab.clear();
    for (auto *r : recs) {
      slab.push_back('@');
      slab.push_back('\n');
      if (need_seq) slab.insert(slab.end(), r->seqp, r->seqp + r->length);
      else slab.insert(slab.end(), r->length, 'A');
      slab.push_back('\n');
      slab.push_back('+');
      slab.push_back('\n');
      if (need_qual) slab.insert(slab.end(), r->qualp, r->qualp + r->length);
      else slab.insert(slab.end(), r->length, '!');
      slab.push_back('\n');
    }
  }
};
}  // namespace detail

template <bool IS_SEQ> class GpuRecordEncoder {
protected:
  using OwnFT = std::conditional_t<IS_SEQ, SeqFreqTable, QualFreqTable>;
  explicit GpuRecordEncoder(const OwnFT *ft, std::shared_ptr<GpuContext> ctx) : ft_(ft), ctx_(std::move(ctx)) {}
  const OwnFT *ft_;
  std::shared_ptr<GpuContext> ctx_;
  std::vector<std::byte> *dst_ = nullptr;
  detail::RecordQueue q_;
  CompressedBuffersDst *cbs_ = nullptr;

  void loadTables() {
    if constexpr (IS_SEQ) {
      auto other = std::make_unique<QualFreqTable>();
      detail::fillUniform(*other);
      ctx_->tablesChanged();
      ctx_->check(fq28_load_tables(ctx_->handle(), ft_, other.get()));
    } else {
      auto other = std::make_unique<SeqFreqTable>();
      detail::fillUniform(*other);
      ctx_->tablesChanged();
      ctx_->check(fq28_load_tables(ctx_->handle(), other.get(), ft_));
    }
  }

public:
  /** src/fse_common.hpp:77-83: ties the stream to dst (pre-sized by the caller) */
  void startChunk(std::vector<std::byte> &dst) {
    dst_ = &dst;
    q_.recs.clear();
  }
  /** src/fse_common.hpp:86-90. @return resulting compressed size, 0 if it
   * does not fit into dst (BIT_closeCStream's overflow convention) */
  std::size_t endChunk() {
    if (q_.recs.empty()) {
      // only the state flush + end mark: encode an empty chunk is not
      // expressible as FASTQ; the reference never does it either
      return 0;
    }
    loadTables();
    q_.build(IS_SEQ, !IS_SEQ);
    std::vector<CompressedBuffersDst> one;
    detail::encodeSlab(*ctx_, q_.slab.data(), q_.slab.size(), q_.slab.size(), true, one, nullptr);
    auto &stream = IS_SEQ ? one.at(0).seq : one.at(0).qual;
    if constexpr (IS_SEQ) {
      if (cbs_) {  // replaceAndEncodeNs side effects, src/fse_sequence.cpp:35-51
        cbs_->n_count.insert(cbs_->n_count.end(), one[0].n_count.begin(), one[0].n_count.end());
        cbs_->n_pos.insert(cbs_->n_pos.end(), one[0].n_pos.begin(), one[0].n_pos.end());
      }
      for (auto *r : q_.recs) std::replace(r->seqp, r->seqp + r->length, 'N', 'A');
    }
    if (stream.size() > dst_->size()) return 0;
    std::copy(stream.begin(), stream.end(), dst_->begin());
    return stream.size();
  }
};

class SequenceEncoder : public GpuRecordEncoder<true> {
public:
  using FreqTableT = SeqFreqTable;
  explicit SequenceEncoder(const FreqTableT *ft, std::shared_ptr<GpuContext> ctx = GpuContext::shared())
      : GpuRecordEncoder<true>(ft, std::move(ctx)) {}
  /** src/fse_sequence.h:93: N counts / positions go to cbs, N -> A in rec */
  void encodeRecord(FastqRecord &rec, CompressedBuffersDst &cbs) {
    cbs_ = &cbs;
    q_.recs.push_back(&rec);
  }
};

class QualityEncoder : public GpuRecordEncoder<false> {
public:
  using FreqTableT = QualFreqTable;
  explicit QualityEncoder(const FreqTableT *ft, std::shared_ptr<GpuContext> ctx = GpuContext::shared())
      : GpuRecordEncoder<false>(ft, std::move(ctx)) {}
  void encodeRecord(const FastqRecord &rec) { q_.recs.push_back(const_cast<FastqRecord *>(&rec)); }
};

template <bool IS_SEQ> class GpuRecordDecoder {
protected:
  using OwnFT = std::conditional_t<IS_SEQ, SeqFreqTable, QualFreqTable>;
  explicit GpuRecordDecoder(const OwnFT *ft, std::shared_ptr<GpuContext> ctx) : ft_(ft), ctx_(std::move(ctx)) {}
  const OwnFT *ft_;
  std::shared_ptr<GpuContext> ctx_;
  std::vector<std::byte> *src_ = nullptr;
  std::vector<FastqRecord *> recs_;  // in CALL order = reverse record order
  CompressedBuffersSrc *cbs_ = nullptr;

  /** runs the chunk: records were queued last-to-first (src/workspace.cpp:84-87) */
  void flush() {
    if (recs_.empty() || !src_) return;
    const std::size_t n = recs_.size();
    std::vector<FastqRecord *> fwd(recs_.rbegin(), recs_.rend());
    // encode the other stream of a dummy chunk with the placeholder table so
    // that one decode pass can run; only our stream's bytes are real
    std::unique_ptr<QualFreqTable> uq;
    std::unique_ptr<SeqFreqTable> us;
    if constexpr (IS_SEQ) {
      uq = std::make_unique<QualFreqTable>();
      detail::fillUniform(*uq);
      ctx_->tablesChanged();
      ctx_->check(fq28_load_tables(ctx_->handle(), ft_, uq.get()));
    } else {
      us = std::make_unique<SeqFreqTable>();
      detail::fillUniform(*us);
      ctx_->tablesChanged();
      ctx_->check(fq28_load_tables(ctx_->handle(), us.get(), ft_));
    }
    detail::RecordQueue q;
    q.recs = fwd;
    q.build(false, false);  // placeholder content for both streams
    std::vector<CompressedBuffersDst> dummy;
    detail::encodeSlab(*ctx_, q.slab.data(), q.slab.size(), q.slab.size(), true, dummy, nullptr);
    std::vector<uint8_t> seq, qual;
    auto as8 = [](const std::vector<std::byte> &s) {
      const auto *p = reinterpret_cast<const uint8_t *>(s.data());
      return std::vector<uint8_t>(p, p + s.size());
    };
    seq = IS_SEQ ? as8(*src_) : as8(dummy.at(0).seq);
    qual = IS_SEQ ? as8(dummy.at(0).qual) : as8(*src_);
    const uint32_t seq_len = static_cast<uint32_t>(seq.size()), qual_len = static_cast<uint32_t>(qual.size());
    seq.resize(seq.size() + 16);
    qual.resize(qual.size() + 16);
    std::vector<uint16_t> readlens(n), hdr_lens(n, 1), n_count(n, 0), n_pos;
    std::vector<uint8_t> headers(n, '@');
    std::size_t total = 0;
    for (std::size_t i = 0; i < n; ++i) {
      readlens[i] = fwd[i]->length;
      total += 2 * static_cast<std::size_t>(fwd[i]->length) + 6;
    }
    if constexpr (IS_SEQ) {
      if (cbs_) {  // pop from the back: src/fse_sequence.cpp:115-126
        const auto *nc_all = reinterpret_cast<const uint16_t *>(cbs_->n_count.data());
        const std::size_t nc_end = cbs_->index.n_count / sizeof(uint16_t);
        if (nc_end < n) throw std::runtime_error("n_count underflow");
        std::size_t my_n = 0;
        for (std::size_t i = 0; i < n; ++i) { n_count[i] = nc_all[nc_end - n + i]; my_n += n_count[i]; }
        const auto *np_all = reinterpret_cast<const uint16_t *>(cbs_->n_pos.data());
        const std::size_t np_end = cbs_->index.n_pos / sizeof(uint16_t);
        if (np_end < my_n) throw std::runtime_error("n_pos underflow");
        n_pos.assign(np_all + (np_end - my_n), np_all + np_end);
        cbs_->index.n_count -= n * sizeof(uint16_t);
        cbs_->index.n_pos -= my_n * sizeof(uint16_t);
      }
    }
    fq28_chunk_info ci{};
    ci.total = static_cast<uint32_t>(total);
    ci.n_records = static_cast<uint32_t>(n);
    ci.seq_len = seq_len;
    ci.qual_len = qual_len;
    ci.n_pos_len = static_cast<uint32_t>(n_pos.size());
    n_pos.resize(n_pos.size() + 1);
    fq28_dec_arenas in{};
    in.seq = seq.data(); in.seq_bytes = seq.size();
    in.qual = qual.data(); in.qual_bytes = qual.size();
    in.readlens = readlens.data();
    in.n_count = n_count.data();
    in.n_pos = n_pos.data(); in.n_pos_entries = ci.n_pos_len;
    in.hdr_lens = hdr_lens.data();
    in.headers = headers.data(); in.headers_bytes = headers.size();
    in.n_records = n;
    std::vector<char> out(total);
    std::size_t wrote = 0;
    ctx_->check(fq28_decompress(ctx_->handle(), &in, &ci, 1, out.data(), out.size(), &wrote));
    const char *p = out.data();
    for (std::size_t i = 0; i < n; ++i) {
      const std::size_t L = fwd[i]->length;
      p += 2;  // "@\n"
      if constexpr (IS_SEQ) std::memcpy(fwd[i]->seqp, p, L);
      p += L + 3;
      if constexpr (!IS_SEQ) std::memcpy(fwd[i]->qualp, p, L);
      p += L + 1;
    }
    recs_.clear();
    src_ = nullptr;
  }

public:
  /** src/fse_common.hpp:130-139 */
  void startChunk(std::vector<std::byte> &src) {
    src_ = &src;
    recs_.clear();
  }
  /** src/fse_common.hpp:141: the stream must be exactly consumed (checked on
   * the GPU; a violation throws from here) */
  void endChunk() { flush(); }
};

class SequenceDecoder : public GpuRecordDecoder<true> {
public:
  using FreqTableT = SeqFreqTable;
  explicit SequenceDecoder(const FreqTableT *ft, std::shared_ptr<GpuContext> ctx = GpuContext::shared())
      : GpuRecordDecoder<true>(ft, std::move(ctx)) {}
  /** src/fse_sequence.h:106; call for records n-1 .. 0; the bases appear in
   * r.seqp when endChunk() returns */
  void decodeRecord(FastqRecord &r, CompressedBuffersSrc &cbs) {
    cbs_ = &cbs;
    recs_.push_back(&r);
  }
};

class QualityDecoder : public GpuRecordDecoder<false> {
public:
  using FreqTableT = QualFreqTable;
  explicit QualityDecoder(const FreqTableT *ft, std::shared_ptr<GpuContext> ctx = GpuContext::shared())
      : GpuRecordDecoder<false>(ft, std::move(ctx)) {}
  void decodeRecord(FastqRecord &r) { recs_.push_back(&r); }
};

}  // namespace fqcomp28
