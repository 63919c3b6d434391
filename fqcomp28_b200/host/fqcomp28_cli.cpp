// fqcomp28_cli.cpp -- `fqcomp28 c|d` on top of the GPU codec path (row N1 of
// SURVEY.md section 8(f)): the reference's command line (src/app.cpp:27-80),
// worker loops (src/process.cpp:32-105) and archive framing
// (src/archive.cpp:22-163, src/archive.h:10-29) around
// CompressionWorkspace::encodeChunks / DecompressionWorkspace::decodeChunks.
//
// Everything that touches bases or qualities runs on the GPU (libfq28.so).
// Host-side, out of the hot path (north_star):
//   * header tokenisation -- same field model as src/headers.cpp:43-133
//     (alnum fields, NUMERIC = int32 delta, STRING = is-different flag +
//     content + 1-byte length), context reset to the dataset's first header at
//     every chunk (src/workspace.cpp:90-93);
//   * the generic coder of the misc buffers.  The reference uses libbsc
//     (src/memcompress.cpp), which is not available offline, so this build
//     writes a STORED container (28-byte header like LIBBSC_HEADER_SIZE, then
//     the raw bytes).  The block framing, the meta section, the seq/qual
//     streams and every integer field are the reference's; only the payload
//     of the bsc-coded buffers differs, so archives are self-consistent but
//     not exchangeable with a libbsc build.  Magic "FQ28STOR" marks them.
// Divergence by default: n_count / n_pos are written per block, not accumulated
// across blocks (SURVEY Q2 makes archives quadratic in size); --ref-compat writes
// them the reference's way (what `--threads 1` produces).
#include <atomic>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <thread>
#include <variant>

#include "fqcomp28_gpu.hpp"

namespace fqcomp28 {

// ---------------------------------------------------------------- settings
struct Settings {  // src/settings.h:15-39
  std::string mates1, archive;
  unsigned sample_mb = 128, reading_mb = 256, n_threads = 1;
  bool verbose = false;
  std::size_t slab_mb = 1024;  // FASTQ bytes handed to a GPU per pass
  bool host_headers = false;   // tokenise headers on the host instead of the GPU (row N2)
  unsigned n_gpus = 1, first_gpu = 0;  // --gpus N [--device D]: one worker per GPU, chunk ranges in file order
  bool ref_compat = false;     // --ref-compat: accumulate n_count / n_pos over blocks like `--threads 1`
};

// ---------------------------------------------------------------- headers
namespace headers {
enum class FieldType { NUMERIC = 0, STRING };
using numeric_t = int32_t;
constexpr std::size_t FIELDLEN_MAX = 255;

struct Format {  // HeaderFormatSpeciciation, src/headers.h:29-42
  std::vector<FieldType> field_types;
  std::vector<char> separators;
  std::size_t n_fields() const { return field_types.size(); }
  static Format fromHeader(std::string_view header) {  // src/headers.cpp:43-73
    if (header.empty() || header[0] != '@') throw std::invalid_argument("header must start with '@'");
    Format fmt;
    const char *p = header.data() + 1, *end = header.data() + header.size();
    for (;;) {
      const char *sep = std::find_if_not(p, end, [](char c) { return std::isalnum(static_cast<unsigned char>(c)); });
      const bool numeric = std::all_of(p, sep, [](char c) { return std::isdigit(static_cast<unsigned char>(c)); });
      fmt.field_types.push_back(numeric ? FieldType::NUMERIC : FieldType::STRING);
      if (sep == end) break;
      if (sep == end - 1) throw std::invalid_argument(std::string(header) + ": header should end in alnum char");
      fmt.separators.push_back(*sep);
      p = sep + 1;
    }
    return fmt;
  }
};

struct FieldStorage {  // src/headers.h:50-82
  std::vector<std::byte> isDifferentFlag, content, contentLength;
  void clear() { isDifferentFlag.clear(); content.clear(); contentLength.clear(); }
};

using field_t = std::variant<numeric_t, std::string>;

inline std::vector<field_t> fieldsOf(std::string_view header, const Format &fmt) {  // src/headers.cpp:25-41
  std::vector<field_t> f(fmt.n_fields());
  const char *p = header.data() + 1, *end = header.data() + header.size();
  for (std::size_t i = 0; i < fmt.n_fields(); ++i) {
    const char *e = i + 1 < fmt.n_fields() ? std::find(p + 1, end, fmt.separators[i]) : end;
    if (fmt.field_types[i] == FieldType::NUMERIC) {
      numeric_t v = 0;
      std::from_chars(p, e, v);
      f[i] = v;
    } else {
      f[i] = std::string(p, e);
    }
    p = e + 1;
  }
  return f;
}

/** encodeHeader for every header of a chunk (src/workspace.cpp:95-125) */
inline void tokenize(const std::vector<std::byte> &raw, const std::vector<readlen_t> &lens, const Format &fmt,
                     const std::vector<field_t> &first, std::vector<FieldStorage> &out) {
  out.assign(fmt.n_fields(), {});
  std::vector<field_t> prev = first;  // startNewChunk, src/workspace.cpp:90-93
  const char *h = reinterpret_cast<const char *>(raw.data());
  for (readlen_t hl : lens) {
    const char *p = h + 1, *end = h + hl;
    for (std::size_t i = 0; i < fmt.n_fields(); ++i) {
      const char *e = i + 1 < fmt.n_fields() ? std::find(std::min(p + 1, end), end, fmt.separators[i]) : end;
      auto &st = out[i];
      if (fmt.field_types[i] == FieldType::STRING) {  // storeString, src/headers.cpp:75-89
        auto &pv = std::get<std::string>(prev[i]);
        const std::string_view val(p, static_cast<std::size_t>(e - p));
        if (val == pv) {
          st.isDifferentFlag.push_back(std::byte{0});
        } else {
          if (val.size() >= FIELDLEN_MAX) throw std::invalid_argument("header field longer than 254 characters");
          st.isDifferentFlag.push_back(std::byte{1});
          st.content.insert(st.content.end(), reinterpret_cast<const std::byte *>(val.data()),
                            reinterpret_cast<const std::byte *>(val.data()) + val.size());
          st.contentLength.push_back(static_cast<std::byte>(val.size()));
          pv.assign(val);
        }
      } else {  // storeNumeric, src/headers.cpp:108-118
        numeric_t v = 0;
        std::from_chars(p, e, v);
        auto &pv = std::get<numeric_t>(prev[i]);
        const numeric_t delta = v - pv;
        st.content.insert(st.content.end(), reinterpret_cast<const std::byte *>(&delta),
                          reinterpret_cast<const std::byte *>(&delta) + sizeof(delta));
        pv = v;
      }
      p = e < end ? e + 1 : end;
    }
    h += hl;
  }
}

/** decodeHeader for n_records headers (src/workspace.cpp:127-157) */
inline void detokenize(const std::vector<FieldStorage> &in, const Format &fmt, const std::vector<field_t> &first,
                       std::size_t n_records, std::vector<std::byte> &raw, std::vector<readlen_t> &lens) {
  std::vector<field_t> prev = first;
  std::vector<std::size_t> dpos(fmt.n_fields(), 0), cpos(fmt.n_fields(), 0), lpos(fmt.n_fields(), 0);
  raw.clear();
  lens.clear();
  std::string line;
  char num[16];
  for (std::size_t r = 0; r < n_records; ++r) {
    line.assign("@");
    for (std::size_t i = 0; i < fmt.n_fields(); ++i) {
      const auto &st = in[i];
      if (fmt.field_types[i] == FieldType::STRING) {  // loadNextString, src/headers.cpp:91-106
        auto &pv = std::get<std::string>(prev[i]);
        if (st.isDifferentFlag.at(dpos[i]++) != std::byte{0}) {
          const std::size_t len = static_cast<unsigned char>(st.contentLength.at(lpos[i]++));
          if (cpos[i] + len > st.content.size()) throw std::runtime_error("header content underflow");
          pv.assign(reinterpret_cast<const char *>(st.content.data()) + cpos[i], len);
          cpos[i] += len;
        }
        line += pv;
      } else {  // loadNextNumeric, src/headers.cpp:120-133
        numeric_t delta;
        if (cpos[i] + sizeof(delta) > st.content.size()) throw std::runtime_error("header content underflow");
        std::memcpy(&delta, st.content.data() + cpos[i], sizeof(delta));
        cpos[i] += sizeof(delta);
        auto &pv = std::get<numeric_t>(prev[i]);
        pv += delta;
        const auto res = std::to_chars(num, num + sizeof(num), pv);
        line.append(num, res.ptr);
      }
      if (i + 1 < fmt.n_fields()) line.push_back(fmt.separators[i]);
    }
    raw.insert(raw.end(), reinterpret_cast<const std::byte *>(line.data()),
               reinterpret_cast<const std::byte *>(line.data()) + line.size());
    lens.push_back(narrow_cast<readlen_t>(line.size()));
  }
}

/** encodeHeader for every header of a batch of blocks on the GPU (fq28_tokenize_headers);
 *  same field streams as tokenize() above, block by block. */
inline void tokenizeOnGpu(GpuContext &ctx, const std::vector<CompressedBuffersDst> &blocks, const Format &fmt,
                          const std::vector<field_t> &first, std::vector<std::vector<FieldStorage>> &out) {
  out.assign(blocks.size(), std::vector<FieldStorage>(fmt.n_fields()));
  if (blocks.empty()) return;
  if (fmt.n_fields() > FQ28_HDR_MAX_FIELDS) throw std::invalid_argument("more header fields than the GPU tokeniser supports");
  std::vector<uint8_t> raw;
  std::vector<uint16_t> lens;
  std::vector<uint64_t> chunk_rec{0};
  for (const auto &cb : blocks) {
    const auto *p = reinterpret_cast<const uint8_t *>(cb.raw_headers.data());
    raw.insert(raw.end(), p, p + cb.raw_headers.size());
    lens.insert(lens.end(), cb.header_lengths.begin(), cb.header_lengths.end());
    chunk_rec.push_back(lens.size());
  }
  fq28_hdr_format f{};
  f.n_fields = static_cast<uint32_t>(fmt.n_fields());
  std::string strings;
  for (std::size_t i = 0; i < fmt.n_fields(); ++i) {
    f.is_string[i] = fmt.field_types[i] == FieldType::STRING;
    if (i + 1 < fmt.n_fields()) f.separators[i] = fmt.separators[i];
    f.first_str_off[i] = static_cast<uint32_t>(strings.size());
    if (f.is_string[i]) strings += std::get<std::string>(first[i]);
    else f.first_numeric[i] = std::get<numeric_t>(first[i]);
  }
  f.first_str_off[fmt.n_fields()] = static_cast<uint32_t>(strings.size());
  f.first_strings = strings.data();
  std::vector<uint8_t> arena(raw.size() + 6 * lens.size() * fmt.n_fields() + 64);
  std::vector<fq28_hdr_field_info> infos(blocks.size() * fmt.n_fields());
  std::size_t used = 0;
  ctx.check(fq28_tokenize_headers(ctx.handle(), raw.data(), raw.size(), lens.data(), lens.size(), chunk_rec.data(),
                                  blocks.size(), &f, arena.data(), arena.size(), infos.data(), &used));
  auto take = [&arena](std::vector<std::byte> &dst, uint64_t off, uint64_t len) {
    const auto *p = reinterpret_cast<const std::byte *>(arena.data() + off);
    dst.assign(p, p + len);
  };
  for (std::size_t k = 0; k < blocks.size(); ++k)
    for (std::size_t i = 0; i < fmt.n_fields(); ++i) {
      const auto &fi = infos[k * fmt.n_fields() + i];
      take(out[k][i].isDifferentFlag, fi.flag_off, fi.flag_len);
      take(out[k][i].content, fi.content_off, fi.content_len);
      take(out[k][i].contentLength, fi.clen_off, fi.clen_len);
    }
}

/** decodeHeader for every record of a batch of blocks on the GPU (fq28_detokenize_headers) */
inline void detokenizeOnGpu(GpuContext &ctx, const std::vector<std::vector<FieldStorage>> &fields,
                            const std::vector<std::size_t> &n_records, const Format &fmt, const std::vector<field_t> &first,
                            std::vector<CompressedBuffersSrc> &blocks) {
  if (fields.empty()) return;
  if (fmt.n_fields() > FQ28_HDR_MAX_FIELDS) throw std::invalid_argument("more header fields than the GPU tokeniser supports");
  const std::size_t F = fmt.n_fields();
  std::vector<uint8_t> arena;
  std::vector<fq28_hdr_field_info> infos(fields.size() * F);
  std::vector<uint64_t> chunk_rec{0};
  auto put = [&arena](const std::vector<std::byte> &b, uint64_t &off, uint64_t &len) {
    off = arena.size();
    len = b.size();
    const auto *p = reinterpret_cast<const uint8_t *>(b.data());
    arena.insert(arena.end(), p, p + b.size());
  };
  for (std::size_t k = 0; k < fields.size(); ++k) {
    for (std::size_t i = 0; i < F; ++i) {
      auto &fi = infos[k * F + i];
      put(fields[k][i].isDifferentFlag, fi.flag_off, fi.flag_len);
      put(fields[k][i].content, fi.content_off, fi.content_len);
      put(fields[k][i].contentLength, fi.clen_off, fi.clen_len);
    }
    chunk_rec.push_back(chunk_rec.back() + n_records[k]);
  }
  fq28_hdr_format f{};
  f.n_fields = static_cast<uint32_t>(F);
  std::string strings;
  for (std::size_t i = 0; i < F; ++i) {
    f.is_string[i] = fmt.field_types[i] == FieldType::STRING;
    if (i + 1 < F) f.separators[i] = fmt.separators[i];
    f.first_str_off[i] = static_cast<uint32_t>(strings.size());
    if (f.is_string[i]) strings += std::get<std::string>(first[i]);
    else f.first_numeric[i] = std::get<numeric_t>(first[i]);
  }
  f.first_str_off[F] = static_cast<uint32_t>(strings.size());
  f.first_strings = strings.data();
  const std::size_t n_rec = static_cast<std::size_t>(chunk_rec.back());
  std::vector<uint8_t> out(64 * n_rec + 4096);
  std::vector<uint16_t> lens(n_rec);
  std::size_t used = 0;
  if (arena.empty()) arena.push_back(0);
  int rc = fq28_detokenize_headers(ctx.handle(), arena.data(), arena.size(), infos.data(), chunk_rec.data(), fields.size(), &f,
                                   out.data(), out.size(), lens.data(), &used);
  if (rc == FQ28_ERR_CAP && used > out.size()) {  // the call reports the size it needs
    out.resize(used);
    rc = fq28_detokenize_headers(ctx.handle(), arena.data(), arena.size(), infos.data(), chunk_rec.data(), fields.size(), &f,
                                 out.data(), out.size(), lens.data(), &used);
  }
  ctx.check(rc);
  std::size_t pos = 0;
  for (std::size_t k = 0; k < fields.size(); ++k) {
    auto &cb = blocks[k];
    cb.header_lengths.assign(lens.begin() + static_cast<std::ptrdiff_t>(chunk_rec[k]),
                             lens.begin() + static_cast<std::ptrdiff_t>(chunk_rec[k + 1]));
    std::size_t bytes = 0;
    for (auto l : cb.header_lengths) bytes += l;
    const auto *p = reinterpret_cast<const std::byte *>(out.data() + pos);
    cb.raw_headers.assign(p, p + bytes);
    pos += bytes;
  }
}
}  // namespace headers

// ---------------------------------------------------------------- misc buffer container
// Stand-in for memcompress/memdecompress (src/memcompress.h:14-28): STORED.
constexpr std::size_t STORED_HEADER = 28;  // = LIBBSC_HEADER_SIZE, src/workspace.h:18
inline std::vector<std::byte> memcompress(const std::vector<std::byte> &src) {
  std::vector<std::byte> dst(STORED_HEADER + src.size(), std::byte{0});
  std::memcpy(dst.data(), "FQ28STOR", 8);
  const uint32_t n = narrow_cast<uint32_t>(src.size());
  std::memcpy(dst.data() + 8, &n, 4);
  if (!src.empty()) std::memcpy(dst.data() + STORED_HEADER, src.data(), src.size());
  return dst;
}
inline std::vector<std::byte> memdecompress(const std::vector<std::byte> &src, std::size_t original) {
  if (src.size() != STORED_HEADER + original || std::memcmp(src.data(), "FQ28STOR", 8) != 0)
    throw std::runtime_error("misc buffer is not a FQ28STOR container (archive written by a libbsc build?)");
  return std::vector<std::byte>(src.begin() + STORED_HEADER, src.end());
}

// ---------------------------------------------------------------- report (row N4)
// InputStats / CompressedStats of src/report.h:10-69: what the stderr table of
// `fqcomp28 c` is computed from.
struct InputStats {
  std::size_t seq = 0, header = 0, n_records = 0;
  std::size_t total() const { return header + 2 * seq + n_records * 5; }  // 4 newlines and a '+' per record
};
struct CompressedStats {
  std::size_t readlens = 0, qual = 0, seq = 0, n_count = 0, n_pos = 0, n_blocks = 0;
  std::vector<std::size_t> header_fields;
  std::size_t sequence() const { return seq + readlens + n_count + n_pos; }
  std::size_t quality() const { return qual; }
  std::size_t headers() const { return std::accumulate(header_fields.begin(), header_fields.end(), std::size_t{0}); }
  /** src/report.cpp:104-125: payload + the u32 size words of every block */
  std::size_t data_section_size(std::size_t n_string_fields) const {
    std::size_t block_meta = 4 + 4 + 2 * 4 + 2 * 4 + 2 * 4 + 2 * 4;
    block_meta += n_string_fields * 3 * 2 * 4 + (header_fields.size() - n_string_fields) * 2 * 4;
    return sequence() + quality() + headers() + n_blocks * block_meta;
  }
};

// ---------------------------------------------------------------- archive
struct BlockInfo {  // src/archive.h:17-26
  int64_t offset;
  uint32_t idx;
};
static_assert(sizeof(BlockInfo) == 16);

class Archive {
public:
  static constexpr std::streamoff OFFSET_META = 4;  // src/archive.h:29
  std::fstream fs;
  DatasetMeta meta;
  headers::Format fmt;
  std::vector<headers::field_t> first_fields;
  std::vector<BlockInfo> index;
  CompressedStats cstats;  // sizes of what writeBlock stores (the reference counts them in the workspace)

  template <typename T> void writeInteger(T v) { fs.write(reinterpret_cast<const char *>(&v), sizeof(v)); }
  template <typename T> T readInteger() { T v{}; fs.read(reinterpret_cast<char *>(&v), sizeof(v)); return v; }
  std::size_t writeBytes(const std::vector<std::byte> &b) {  // src/archive.cpp:165-169
    writeInteger(narrow_cast<uint32_t>(b.size()));
    fs.write(reinterpret_cast<const char *>(b.data()), static_cast<std::streamsize>(b.size()));
    return b.size();
  }
  void readBytes(std::vector<std::byte> &b) {  // src/archive.cpp:171-175
    b.resize(readInteger<uint32_t>());
    fs.read(reinterpret_cast<char *>(b.data()), static_cast<std::streamsize>(b.size()));
  }
  void setMeta(DatasetMeta &&m) {
    meta = std::move(m);
    fmt = headers::Format::fromHeader(meta.first_header);
    first_fields = headers::fieldsOf(meta.first_header, fmt);
  }
  void writeMeta() {  // src/archive.cpp:22-25, src/prepare.cpp:12-21
    std::vector<char> bytes;
    DatasetMeta::storeToBytes(meta, bytes);
    fs.seekp(OFFSET_META);
    fs.write(bytes.data(), static_cast<std::streamsize>(bytes.size()));
  }
  /** src/archive.cpp:57-106 */
  void writeBlock(const CompressedBuffersDst &cb, const std::vector<headers::FieldStorage> &fields) {
    BlockInfo bi{};
    bi.idx = cb.chunk_idx;
    bi.offset = static_cast<int64_t>(fs.tellp());
    writeInteger(cb.original_size.total);
    writeInteger(cb.original_size.n_records);
    writeInteger(cb.original_size.readlens);
    cstats.readlens += writeBytes(memcompress(cb.readlens));
    writeInteger(cb.original_size.n_count);
    cstats.n_count += writeBytes(memcompress(cb.n_count));
    writeInteger(cb.original_size.n_pos);
    cstats.n_pos += writeBytes(memcompress(cb.n_pos));
    cstats.seq += writeBytes(cb.seq);
    cstats.qual += writeBytes(cb.qual);
    cstats.header_fields.resize(fmt.n_fields());
    cstats.n_blocks++;
    for (std::size_t i = 0; i < fmt.n_fields(); ++i) {
      const auto &f = fields[i];
      if (fmt.field_types[i] == headers::FieldType::STRING) {
        writeInteger(narrow_cast<uint32_t>(f.isDifferentFlag.size()));
        cstats.header_fields[i] += writeBytes(memcompress(f.isDifferentFlag));
        writeInteger(narrow_cast<uint32_t>(f.content.size()));
        cstats.header_fields[i] += writeBytes(memcompress(f.content));
        writeInteger(narrow_cast<uint32_t>(f.contentLength.size()));
        cstats.header_fields[i] += writeBytes(memcompress(f.contentLength));
      } else {
        writeInteger(narrow_cast<uint32_t>(f.content.size()));
        cstats.header_fields[i] += writeBytes(memcompress(f.content));
      }
    }
    index.push_back(bi);
  }
  /** src/archive.cpp:45-55 */
  void writeIndex() {
    const std::streamoff end = fs.tellp();
    fs.seekp(0);
    writeInteger(narrow_cast<uint32_t>(index.size()));
    fs.seekp(end);
    fs.write(reinterpret_cast<const char *>(index.data()), static_cast<std::streamsize>(index.size() * sizeof(BlockInfo)));
  }
  /** src/archive.cpp:27-43 */
  void readHeader() {
    const uint32_t n_blocks = readInteger<uint32_t>();
    readlen_t hlen = readInteger<readlen_t>();
    std::vector<char> m(sizeof(hlen) + hlen + sizeof(SeqFreqTable) + sizeof(QualFreqTable));
    std::memcpy(m.data(), &hlen, sizeof(hlen));
    fs.read(m.data() + sizeof(hlen), static_cast<std::streamsize>(m.size() - sizeof(hlen)));
    setMeta(DatasetMeta::loadFromBytes(m.data(), m.size()));
    const auto data_start = fs.tellg();
    fs.seekg(-static_cast<std::streamoff>(n_blocks * sizeof(BlockInfo)), std::ios_base::end);
    index.resize(n_blocks);
    fs.read(reinterpret_cast<char *>(index.data()), static_cast<std::streamsize>(n_blocks * sizeof(BlockInfo)));
    fs.seekg(data_start);
    std::sort(index.begin(), index.end(), [](const BlockInfo &a, const BlockInfo &b) { return a.idx < b.idx; });
    if (!fs) throw std::runtime_error("truncated archive");
  }
  /** src/archive.cpp:108-163 */
  void readBlock(const BlockInfo &bi, CompressedBuffersSrc &cb, std::vector<headers::FieldStorage> &fields) {
    cb.clear();
    fs.seekg(bi.offset);
    cb.chunk_idx = bi.idx;
    cb.original_size.total = readInteger<uint32_t>();
    cb.original_size.n_records = readInteger<uint32_t>();
    std::vector<std::byte> tmp;
    cb.original_size.readlens = readInteger<uint32_t>();
    readBytes(tmp);
    cb.readlens = memdecompress(tmp, cb.original_size.readlens);
    cb.original_size.n_count = readInteger<uint32_t>();
    readBytes(tmp);
    cb.n_count = memdecompress(tmp, cb.original_size.n_count);
    cb.original_size.n_pos = readInteger<uint32_t>();
    readBytes(tmp);
    cb.n_pos = memdecompress(tmp, cb.original_size.n_pos);
    readBytes(cb.seq);
    readBytes(cb.qual);
    fields.assign(fmt.n_fields(), {});
    for (std::size_t i = 0; i < fmt.n_fields(); ++i) {
      auto &f = fields[i];
      if (fmt.field_types[i] == headers::FieldType::STRING) {
        uint32_t o = readInteger<uint32_t>();
        readBytes(tmp);
        f.isDifferentFlag = memdecompress(tmp, o);
        o = readInteger<uint32_t>();
        readBytes(tmp);
        f.content = memdecompress(tmp, o);
        o = readInteger<uint32_t>();
        readBytes(tmp);
        f.contentLength = memdecompress(tmp, o);
      } else {
        const uint32_t o = readInteger<uint32_t>();
        readBytes(tmp);
        f.content = memdecompress(tmp, o);
      }
    }
    if (!fs) throw std::runtime_error("truncated block");
  }
};

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- c
/** processReads, src/process.cpp:32-82 */
/** The stderr table of the reference (src/report.cpp:37-102): input sizes, stored
 *  stream sizes with their share of the archive, compression ratios, block count.
 *  Returns the archive size the table is based on (index + data section + meta),
 *  which the reference asserts to equal the file size. */
static std::size_t printReport(const InputStats &inp, const Archive &ar, std::FILE *os) {
  const CompressedStats &comp = ar.cstats;
  const char *sep = "*****************************\n";
  const std::size_t n_string = static_cast<std::size_t>(
      std::count(ar.fmt.field_types.begin(), ar.fmt.field_types.end(), headers::FieldType::STRING));
  const std::size_t meta_headers = sizeof(readlen_t) + ar.meta.first_header.size();
  const std::size_t meta_seq = sizeof(SeqFreqTable), meta_qual = sizeof(QualFreqTable);
  const std::size_t index_bytes = sizeof(uint32_t) + ar.index.size() * sizeof(BlockInfo);
  const std::size_t archive_size = index_bytes + comp.data_section_size(n_string) + meta_headers + meta_seq + meta_qual;
  auto ratio = [](std::size_t a, std::size_t b) { return static_cast<double>(a) / static_cast<double>(b); };
  std::fputs(sep, os);
  std::fputs("Input sizes:\n", os);
  std::fprintf(os, "Sequence\t%zu\t\nQuality\t%zu\t\nHeaders\t%zu\t\n", inp.seq, inp.seq, inp.header);
  std::fputs(sep, os);
  std::fputs("Compressed stream sizes:\n", os);
  auto cstream = [&](const std::string &name, std::size_t bytes) {
    std::fprintf(os, "%s\t%zu\t%.3f\t\n", name.c_str(), bytes, ratio(bytes, archive_size));
  };
  cstream("seq", comp.seq);
  cstream("readlens", comp.readlens);
  cstream("n_count", comp.n_count);
  cstream("n_pos", comp.n_pos);
  cstream("qual", comp.qual);
  for (std::size_t i = 0; i < comp.header_fields.size(); ++i) cstream("header_field_" + std::to_string(i + 1), comp.header_fields[i]);
  cstream("meta_seq", meta_seq);
  cstream("meta_qual", meta_qual);
  std::fputs(sep, os);
  std::fputs("CR\n", os);
  std::fprintf(os, "Sequence\t%.3f\t\n", ratio(inp.seq, comp.sequence() + meta_seq));
  std::fprintf(os, "Quality\t%.3f\t\n", ratio(inp.seq, comp.quality() + meta_qual));
  std::fprintf(os, "Headers\t%.3f\t\n", ratio(inp.header, comp.headers() + meta_headers));
  std::fprintf(os, "Total\t%.3f\t\n", ratio(inp.total(), archive_size));
  std::fputs(sep, os);
  std::fprintf(os, "# blocks: \t%zu\t\n", comp.n_blocks);
  return archive_size;
}

// ---------------------------------------------------------------- pipeline plumbing
/** results come back from the workers in any order and leave in index order */
template <class T> class OrderedQueue {
public:
  void put(std::size_t index, T &&v) {
    std::lock_guard<std::mutex> lk(mu_);
    items_.emplace(index, std::move(v));
    cv_.notify_all();
  }
  /** blocks until item `index` is there (or a worker failed) */
  bool take(std::size_t index, T &out) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return failed_ || items_.count(index); });
    if (failed_) return false;
    out = std::move(items_.at(index));
    items_.erase(index);
    return true;
  }
  void fail() {
    std::lock_guard<std::mutex> lk(mu_);
    failed_ = true;
    cv_.notify_all();
  }

private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::map<std::size_t, T> items_;
  bool failed_ = false;
};

/** counting semaphore: bounds the slabs / batches in flight (host memory) */
class Tokens {
public:
  explicit Tokens(std::size_t n) : n_(n) {}
  void acquire() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return n_ > 0; });
    --n_;
  }
  void release() {
    std::lock_guard<std::mutex> lk(mu_);
    ++n_;
    cv_.notify_one();
  }
  void abort() {  // a stage failed: nobody may stay blocked here
    std::lock_guard<std::mutex> lk(mu_);
    n_ = static_cast<std::size_t>(1) << 40;
    cv_.notify_all();
  }

private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::size_t n_;
};

static void warnFewChunks(std::size_t n_chunks, std::size_t per_gpu_mb, unsigned reading_mb) {
  static std::atomic<bool> warned{false};
  if (n_chunks >= 64 || warned.exchange(true)) return;
  std::fprintf(stderr,
               "fqcomp28: note: only %zu chunks per %zu MB GPU pass at -R %u; a chunk is one serial tANS stream per\n"
               "          stream type (src/fse_common.hpp:77-141), so the GPU decodes chunks in parallel, not symbols.\n"
               "          Archives written with -R 1..16 decompress one to two orders of magnitude faster here.\n",
               n_chunks, per_gpu_mb, reading_mb);
}

/** one context (= one fq28 handle, own streams and buffers) per worker; worker g runs on device
 *  first_gpu + g, wrapping around the visible devices (two workers on one GPU are legal: they
 *  just share it, which is also how the pipeline is tested on a single-GPU box) */
static std::vector<std::shared_ptr<GpuContext>> workerContexts(const Settings &set) {
  const int visible = fq28_device_count();
  if (visible <= 0) throw std::runtime_error("no CUDA device (the fqcomp28 GPU path has no CPU fallback)");
  const unsigned n = std::max(1u, set.n_gpus);
  if (static_cast<int>(n) > visible)
    std::fprintf(stderr, "fqcomp28: note: --gpus %u with %d visible device(s): workers share devices\n", n, visible);
  std::vector<std::shared_ptr<GpuContext>> ctxs;
  for (unsigned g = 0; g < n; ++g) ctxs.push_back(std::make_shared<GpuContext>(static_cast<int>((set.first_gpu + g) % visible)));
  return ctxs;
}

// ---------------------------------------------------------------- c
/** Byte buffers reused from slab to slab: a fresh gigabyte-sized vector costs its zero-fill and
 *  page faults every time, which is more than the codec takes. */
class BufferPool {
public:
  std::vector<char> get() {
    std::lock_guard<std::mutex> lk(mu_);
    if (free_.empty()) return {};
    std::vector<char> b = std::move(free_.back());
    free_.pop_back();
    return b;
  }
  void put(std::vector<char> &&b) {
    std::lock_guard<std::mutex> lk(mu_);
    free_.push_back(std::move(b));
  }

private:
  std::mutex mu_;
  std::vector<std::vector<char>> free_;
};

/** processReads, src/process.cpp:32-82, as a pipeline over one or several GPUs:
 *    reader thread  : file -> slab buffers (slab i = [i*B, (i+1)*B + R), so that it holds the start of
 *                     its first chunk whatever the previous slab consumed)
 *    worker per GPU : slab i goes to GPU i mod N.  H2D copy (fq28_stage) right away; then, in slab
 *                     order, the cheap boundary step (fq28_plan: parseRecords + chunk walk) that tells
 *                     where slab i+1 starts -- the only sequential dependency between GPUs, ~2 ms per
 *                     GB -- then encodeChunks + header tokenisation, overlapped between GPUs
 *    this thread    : writes blocks in slab order = chunk order (idx), like `--threads 1`
 *  The chunk boundaries are those of a single sequential walk, so the archive is byte-identical
 *  for any number of GPUs. */
static int compress(const Settings &set) {
  std::ifstream in(set.mates1, std::ios::binary);
  if (!in) throw std::system_error(errno, std::generic_category(), set.mates1);
  in.seekg(0, std::ios::end);
  const std::size_t file_size = static_cast<std::size_t>(in.tellg());
  in.seekg(0);
  const std::size_t R = static_cast<std::size_t>(set.reading_mb) << 20, S = static_cast<std::size_t>(set.sample_mb) << 20;
  const std::size_t B = std::max(set.slab_mb << 20, 2 * R);   // nominal slab
  const unsigned n_gpus = std::max(1u, set.n_gpus);
  std::vector<std::shared_ptr<GpuContext>> ctxs = workerContexts(set);
  Archive ar;
  ar.fs.open(set.archive, std::ios::binary | std::ios::out | std::ios::trunc);
  if (!ar.fs) throw std::system_error(errno, std::generic_category(), set.archive);

  const double t_start = now_s();
  {  // analyzeDataset (src/prepare.cpp:42-47): first chunk of a reader with reading size S, on the first GPU
    FastqChunk sample;
    sample.raw_data.resize(std::min(S, file_size));
    in.read(sample.raw_data.data(), static_cast<std::streamsize>(sample.raw_data.size()));
    const std::size_t used = FastqReader::parseRecords(sample, *ctxs[0]);
    sample.raw_data.resize(used);
    if (sample.records.empty()) throw std::invalid_argument("no complete FASTQ record in the sample window");
    ar.setMeta(DatasetMeta(sample, *ctxs[0]));
    ar.writeMeta();
  }
  const std::size_t n_slabs = file_size == 0 ? 0 : std::max<std::size_t>(1, (file_size + B - 1) / B);
  // slab i is the last one when its buffer reaches the end of the file
  auto slab_end = [&](std::size_t i) { return std::min(file_size, (i + 1) * B + R); };
  std::size_t last_slab = 0;
  while (last_slab + 1 < n_slabs && slab_end(last_slab) < file_size) ++last_slab;

  struct Slab { std::size_t base = 0, size = 0; std::vector<char> buf; };   // buf.size() >= size (pooled)
  BufferPool slab_pool;
  struct Result {
    std::vector<CompressedBuffersDst> blocks;
    std::vector<std::vector<headers::FieldStorage>> fields;
    double t_gpu = 0, t_hdr = 0;
  };
  OrderedQueue<Slab> slabs;
  OrderedQueue<Result> results;
  Tokens in_flight(2 * n_gpus + 1);
  std::exception_ptr error;
  std::mutex error_mu;
  std::mutex baton_mu;
  std::condition_variable baton_cv;
  bool aborted = false;
  auto fail = [&](std::exception_ptr e) {
    { std::lock_guard<std::mutex> lk(error_mu); if (!error) error = e; }
    slabs.fail();
    results.fail();
    in_flight.abort();
    { std::lock_guard<std::mutex> lk(baton_mu); aborted = true; baton_cv.notify_all(); }
  };

  std::thread reader([&] {
    try {
      std::ifstream f(set.mates1, std::ios::binary);
      for (std::size_t i = 0; i <= last_slab; ++i) {
        in_flight.acquire();
        Slab s;
        s.base = i * B;
        s.size = slab_end(i) - s.base;
        s.buf = slab_pool.get();
        if (s.buf.size() < s.size) { s.buf.clear(); s.buf.resize(s.size); }
        f.seekg(static_cast<std::streamoff>(s.base));
        f.read(s.buf.data(), static_cast<std::streamsize>(s.size));
        if (!f) throw std::runtime_error("short read from " + set.mates1);
        slabs.put(i, std::move(s));
      }
    } catch (...) { fail(std::current_exception()); }
  });

  // the baton: start offset of slab i, published by the worker of slab i-1 after its boundary step
  std::vector<std::size_t> start(last_slab + 2, 0);
  std::size_t known = 0;  // start[0..known] are valid

  std::vector<std::thread> workers;
  for (unsigned g = 0; g < n_gpus; ++g) {
    workers.emplace_back([&, g] {
      try {
        GpuContext &ctx = *ctxs[g];
        CompressionWorkspace wksp(&ar.meta, ctxs[g]);
        for (std::size_t i = g; i <= last_slab; i += n_gpus) {
          Slab s;
          if (!slabs.take(i, s)) return;
          ctx.check(fq28_stage(ctx.handle(), s.buf.data(), s.size));  // async H2D, before the start is known
          std::size_t s_i;
          {
            std::unique_lock<std::mutex> lk(baton_mu);
            baton_cv.wait(lk, [&] { return aborted || known >= i; });
            if (aborted) return;
            s_i = start[i];
          }
          const bool eof = i == last_slab;
          if (s_i < s.base || s_i > s.base + s.size) throw std::runtime_error("slab does not hold its first chunk; raise --slab-mb");
          const char *p = s.buf.data() + (s_i - s.base);
          const std::size_t n = s.size - (s_i - s.base);
          double t0 = now_s();
          uint64_t consumed = 0;
          std::size_t n_chunks = 0;
          wksp.useTables();
          ctx.check(fq28_plan(ctx.handle(), p, n, R, eof ? 1 : 0, &consumed, &n_chunks));
          if (!eof && consumed == 0) throw std::runtime_error("no whole chunk fits the slab; raise --slab-mb");
          {
            std::lock_guard<std::mutex> lk(baton_mu);
            start[i + 1] = s_i + consumed;
            known = i + 1;
            baton_cv.notify_all();
          }
          Result r;
          std::size_t consumed2 = 0;
          wksp.encodeChunks(p, n, R, eof, r.blocks, &consumed2);  // reuses the plan: straight to the encode
          if (consumed2 != consumed) throw std::runtime_error("boundary step and encode disagree");
          r.t_gpu = now_s() - t0;
          t0 = now_s();
          if (!set.host_headers) headers::tokenizeOnGpu(ctx, r.blocks, ar.fmt, ar.first_fields, r.fields);
          r.t_hdr = now_s() - t0;
          ctx.check(fq28_stage(ctx.handle(), nullptr, 0));
          warnFewChunks(r.blocks.size(), s.size >> 20, set.reading_mb);
          slab_pool.put(std::move(s.buf));
          results.put(i, std::move(r));
        }
      } catch (...) { fail(std::current_exception()); }
    });
  }

  double t_gpu = 0, t_host = 0;
  uint32_t next_idx = 0;
  std::size_t tot_seq = 0, tot_qual = 0, n_records = 0;
  InputStats istats;
  std::vector<std::byte> acc_n_count, acc_n_pos;  // --ref-compat: SURVEY Q2
  try {
  for (std::size_t i = 0; i <= last_slab && n_slabs; ++i) {
    Result r;
    if (!results.take(i, r)) break;
    in_flight.release();
    t_gpu += r.t_gpu;
    const double t0 = now_s();
    std::vector<headers::FieldStorage> fields;
    std::size_t bi = 0;
    for (auto &cb : r.blocks) {
      cb.chunk_idx = next_idx++;
      if (set.host_headers) headers::tokenize(cb.raw_headers, cb.header_lengths, ar.fmt, ar.first_fields, fields);
      tot_seq += cb.seq.size();
      tot_qual += cb.qual.size();
      n_records += cb.original_size.n_records;
      istats.n_records += cb.original_size.n_records;
      istats.header += cb.raw_headers.size();
      for (std::size_t q = 0; q + 1 < cb.readlens.size(); q += 2)  // readlens: u16 LE per record
        istats.seq += std::to_integer<std::size_t>(cb.readlens[q]) | (std::to_integer<std::size_t>(cb.readlens[q + 1]) << 8);
      if (set.ref_compat) {
        // src/compressed_buffers.h:58-68: the reference never clears n_count / n_pos between the
        // chunks of a worker thread, so with --threads 1 block k carries the N data of blocks 0..k
        acc_n_count.insert(acc_n_count.end(), cb.n_count.begin(), cb.n_count.end());
        acc_n_pos.insert(acc_n_pos.end(), cb.n_pos.begin(), cb.n_pos.end());
        cb.n_count = acc_n_count;
        cb.n_pos = acc_n_pos;
        cb.original_size.n_count = narrow_cast<uint32_t>(cb.n_count.size());
        cb.original_size.n_pos = narrow_cast<uint32_t>(cb.n_pos.size());
      }
      ar.writeBlock(cb, set.host_headers ? fields : r.fields[bi]);
      ++bi;
    }
    t_host += now_s() - t0 + r.t_hdr;
  }
  } catch (...) { fail(std::current_exception()); }
  reader.join();
  for (auto &w : workers) w.join();
  if (error) std::rethrow_exception(error);
  ar.writeIndex();
  ar.fs.flush();
  {  // src/process.cpp:80-81: flush, then the report; the reference asserts archive_size == file size
    const std::size_t archive_size = printReport(istats, ar, stderr);
    ar.fs.seekp(0, std::ios::end);
    const std::size_t on_disk = static_cast<std::size_t>(ar.fs.tellp());
    if (archive_size != on_disk)
      std::fprintf(stderr, "warning: report assumes an archive of %zu bytes, the file has %zu\n", archive_size, on_disk);
  }
  const double wall = now_s() - t_start;
  std::fprintf(stderr,
               "fqcomp28 (B200 path): %zu records, %u blocks, seq %zu B, qual %zu B; %u GPU(s); wall %.3f s = %.0f MB/s "
               "(codec passes %.3f s summed over GPUs, headers + archive writes %.3f s)\n",
               n_records, next_idx, tot_seq, tot_qual, n_gpus, wall, file_size / 1e6 / std::max(wall, 1e-9), t_gpu, t_host);
  return 0;
}

// ---------------------------------------------------------------- d
/** processArchiveParts, src/process.cpp:84-105: this thread reads blocks (sorted index) into
 *  batches, batch j is decoded on GPU j mod N, a writer thread emits the chunks in idx order
 *  (FastqWriter::writeChunk, src/fastq_io.cpp:131-143). */
static int decompress(const Settings &set) {
  Archive ar;
  ar.fs.open(set.archive, std::ios::binary | std::ios::in);
  if (!ar.fs) throw std::system_error(errno, std::generic_category(), set.archive);
  ar.readHeader();
  std::ofstream out(set.mates1, std::ios::binary | std::ios::trunc);
  if (!out) throw std::system_error(errno, std::generic_category(), set.mates1);
  const unsigned n_gpus = std::max(1u, set.n_gpus);
  std::vector<std::shared_ptr<GpuContext>> ctxs = workerContexts(set);
  const std::size_t batch_bytes = set.slab_mb << 20;
  const double t_start = now_s();

  struct Batch {
    std::vector<CompressedBuffersSrc> srcs;
    std::vector<std::vector<headers::FieldStorage>> fields;
    std::vector<std::size_t> records;
  };
  using Bytes = std::vector<char>;   // the FASTQ bytes of a batch, chunks in idx order
  BufferPool out_pool;
  OrderedQueue<Batch> batches;
  OrderedQueue<Bytes> decoded;
  Tokens in_flight(2 * n_gpus + 1);
  std::exception_ptr error;
  std::mutex error_mu;
  std::size_t n_batches = 0;
  std::mutex nb_mu;
  std::condition_variable nb_cv;
  bool all_read = false;
  auto fail = [&](std::exception_ptr e) {
    { std::lock_guard<std::mutex> lk(error_mu); if (!error) error = e; }
    batches.fail();
    decoded.fail();
    in_flight.abort();
    { std::lock_guard<std::mutex> lk(nb_mu); all_read = true; nb_cv.notify_all(); }
  };
  auto batch_exists = [&](std::size_t j) {  // waits until batch j has been read or the archive is exhausted
    std::unique_lock<std::mutex> lk(nb_mu);
    nb_cv.wait(lk, [&] { return all_read || n_batches > j; });
    return n_batches > j;
  };

  std::vector<std::thread> workers;
  for (unsigned g = 0; g < n_gpus; ++g) {
    workers.emplace_back([&, g] {
      try {
        GpuContext &ctx = *ctxs[g];
        DecompressionWorkspace wksp(&ar.meta, ctxs[g]);
        for (std::size_t j = g; batch_exists(j); j += n_gpus) {
          Batch b;
          if (!batches.take(j, b)) return;
          if (!set.host_headers) headers::detokenizeOnGpu(ctx, b.fields, b.records, ar.fmt, ar.first_fields, b.srcs);
          std::vector<CompressedBuffersSrc *> ps;
          for (std::size_t k = 0; k < b.srcs.size(); ++k) ps.push_back(&b.srcs[k]);
          Bytes bytes = out_pool.get();
          wksp.decodeChunksRaw(ps, bytes);
          decoded.put(j, std::move(bytes));
        }
      } catch (...) { fail(std::current_exception()); }
    });
  }
  std::thread writer([&] {
    try {
      for (std::size_t j = 0; batch_exists(j); ++j) {
        Bytes bytes;
        if (!decoded.take(j, bytes)) return;
        out.write(bytes.data(), static_cast<std::streamsize>(bytes.size()));  // FastqWriter::writeChunk, idx order
        out_pool.put(std::move(bytes));
        in_flight.release();
      }
    } catch (...) { fail(std::current_exception()); }
  });

  try {
    std::size_t i = 0;
    while (i < ar.index.size()) {
      in_flight.acquire();
      Batch b;
      std::vector<headers::FieldStorage> fields;
      std::size_t bytes = 0;
      while (i < ar.index.size() && (b.srcs.empty() || bytes < batch_bytes)) {
        b.srcs.emplace_back();
        ar.readBlock(ar.index[i++], b.srcs.back(), fields);
        auto &cb = b.srcs.back();
        if (set.host_headers) {
          headers::detokenize(fields, ar.fmt, ar.first_fields, cb.original_size.n_records, cb.raw_headers, cb.header_lengths);
        } else {
          b.fields.push_back(fields);
          b.records.push_back(cb.original_size.n_records);
        }
        bytes += cb.original_size.total;
      }
      warnFewChunks(b.srcs.size(), bytes >> 20, static_cast<unsigned>((bytes / std::max<std::size_t>(1, b.srcs.size())) >> 20));
      std::size_t j;
      { std::lock_guard<std::mutex> lk(nb_mu); j = n_batches; }
      batches.put(j, std::move(b));
      { std::lock_guard<std::mutex> lk(nb_mu); ++n_batches; nb_cv.notify_all(); }
    }
  } catch (...) { fail(std::current_exception()); }
  { std::lock_guard<std::mutex> lk(nb_mu); all_read = true; nb_cv.notify_all(); }
  for (auto &w : workers) w.join();
  writer.join();
  if (error) std::rethrow_exception(error);
  out.flush();
  {
    const double wall = now_s() - t_start;
    const std::size_t bytes = static_cast<std::size_t>(out.tellp());
    std::fprintf(stderr, "fqcomp28 d (B200 path): %zu blocks, %zu bytes on %u GPU(s); wall %.3f s = %.0f MB/s\n", ar.index.size(), bytes,
                 n_gpus, wall, bytes / 1e6 / std::max(wall, 1e-9));
  }
  return out ? 0 : 1;
}

}  // namespace fqcomp28

// ---------------------------------------------------------------- CLI (src/app.cpp:27-80, without CLI11)
static void usage() {
  std::fprintf(stderr,
               "fqcomp28 (B200 codec path)\n"
               "  fqcomp28 c --i1|--input1 FILE -o|--output ARCHIVE [-S|--sample-size-Mb N=128] [-R|--reading-size-Mb N=256]\n"
               "             [-t|--threads N] [--verbose] [--slab-mb N=1024] [--host-headers] [--gpus N] [--device D] [--ref-compat]\n"
               "  fqcomp28 d -i|--input ARCHIVE --o1|--output1 FILE [-t|--threads N] [--verbose] [--slab-mb N=1024] [--host-headers]\n"
               "             [--gpus N] [--device D]\n"
               "  --gpus N      one worker per GPU; chunk boundaries come from one sequential walk, so the archive does not\n"
               "                depend on N (blocks in file order, like --threads 1)\n"
               "  --ref-compat  n_count / n_pos accumulate over blocks as in the reference (src/compressed_buffers.h:58-68)\n");
}

int main(int argc, char **argv) {
  using namespace fqcomp28;
  if (argc < 2) { usage(); return 106; }
  const std::string cmd = argv[1];
  if (cmd == "header-format" && argc == 3) {
    // HeaderFormatSpeciciation::fromHeader (src/headers.cpp:43-73) on one header line: field count,
    // field types (S = STRING, N = NUMERIC), separators.  Test hook for test/headers_test.cpp:12-33.
    try {
      const auto fmt = headers::Format::fromHeader(argv[2]);
      std::string types, seps(fmt.separators.begin(), fmt.separators.end());
      for (auto t : fmt.field_types) types += t == headers::FieldType::STRING ? 'S' : 'N';
      std::printf("%zu\n%s\n%s\n", fmt.n_fields(), types.c_str(), seps.c_str());
      return 0;
    } catch (const std::exception &e) {
      std::fprintf(stderr, "fqcomp28: %s\n", e.what());
      return 1;
    }
  }
  Settings set;
  auto need = [&](int &i) -> std::string {
    if (i + 1 >= argc) { usage(); std::exit(106); }
    return argv[++i];
  };
  auto mb = [&](const std::string &v) {  // NonNegativeNumber in the reference; 0 is UB there
    const long x = std::strtol(v.c_str(), nullptr, 10);
    if (x < 1) { std::fprintf(stderr, "size options must be >= 1 MB\n"); std::exit(105); }
    return static_cast<unsigned>(x);
  };
  for (int i = 2; i < argc; ++i) {
    const std::string a = argv[i];
    if (cmd == "c" && (a == "--i1" || a == "--input1")) set.mates1 = need(i);
    else if (cmd == "c" && (a == "-o" || a == "--output")) set.archive = need(i);
    else if (cmd == "c" && (a == "-S" || a == "--sample-size-Mb")) set.sample_mb = mb(need(i));
    else if (cmd == "c" && (a == "-R" || a == "--reading-size-Mb")) set.reading_mb = mb(need(i));
    else if (cmd == "d" && (a == "-i" || a == "--input")) set.archive = need(i);
    else if (cmd == "d" && (a == "--o1" || a == "--output1")) set.mates1 = need(i);
    else if (a == "-t" || a == "--threads") set.n_threads = mb(need(i));  // accepted; the GPU path has one worker per GPU
    else if (a == "--slab-mb") set.slab_mb = mb(need(i));
    else if (a == "--verbose") set.verbose = true;
    else if (a == "--host-headers") set.host_headers = true;
    else if (a == "--gpus") set.n_gpus = mb(need(i));
    else if (a == "--device") set.first_gpu = static_cast<unsigned>(std::strtol(need(i).c_str(), nullptr, 10));
    else if (cmd == "c" && a == "--ref-compat") set.ref_compat = true;
    else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); usage(); return 109; }
  }
  if ((cmd != "c" && cmd != "d") || set.mates1.empty() || set.archive.empty()) { usage(); return 106; }
  try {
    return cmd == "c" ? compress(set) : decompress(set);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "fqcomp28: %s\n", e.what());
    return 1;
  }
}
