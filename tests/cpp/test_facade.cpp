// Transliteration of the reference's own unit tests onto the GPU facade
// (fqcomp28_b200/host/fqcomp28_gpu.hpp):
//   test/fse_sequence_test.cpp:17-50   "FSE Sequence"
//   test/fse_quality_test.cpp:17-49    "FSE Quality"
//   test/workspace_test.cpp:45-69      "Workspace::encodeChunk"
//   test/archive_test.cpp:23-50        meta round trip (store/load part)
// Catch2 is not available offline; CHECK is a counting assert.
#include <cstdio>
#include <fstream>
#include <ranges>

#include "../../fqcomp28_b200/host/fqcomp28_gpu.hpp"

using namespace fqcomp28;

static int g_failed = 0, g_checked = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    ++g_checked;                                                             \
    if (!(cond)) { ++g_failed; std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

static std::vector<char> loadFileContents(const std::string &path) {
  std::ifstream ifs(path, std::ios::binary);
  if (!ifs) throw std::runtime_error("cannot open " + path);
  return std::vector<char>((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
}
static FastqChunk loadFastqFileContents(const std::string &path) {  // test/test_utils.cpp:23-28
  FastqChunk chunk;
  chunk.raw_data = loadFileContents(path);
  FastqReader::parseRecords(chunk);
  return chunk;
}
static CompressedBuffersSrc convertToSrcBuffers(CompressedBuffersDst &&in) {  // test/test_utils.h:25-49
  CompressedBuffersSrc src;
  src.original_size = in.original_size;
  src.seq = std::move(in.seq);
  src.qual = std::move(in.qual);
  src.n_count = std::move(in.n_count);
  src.index.n_count = src.n_count.size();
  src.n_pos = std::move(in.n_pos);
  src.index.n_pos = src.n_pos.size();
  src.readlens = std::move(in.readlens);
  src.raw_headers = std::move(in.raw_headers);
  src.header_lengths = std::move(in.header_lengths);
  return src;
}
static std::vector<char> concatSeq(const FastqChunk &c) {
  std::vector<char> ret;
  for (const auto &r : c.records) ret.insert(ret.end(), r.seqp, r.seqp + r.length);
  return ret;
}
static std::vector<char> concatQual(const FastqChunk &c) {
  std::vector<char> ret;
  for (const auto &r : c.records) ret.insert(ret.end(), r.qualp, r.qualp + r.length);
  return ret;
}

static void testFseSequence(const std::string &dir) {
  FastqChunk chunk = loadFastqFileContents(dir + "/SRR065390_sub_1.fastq");
  CHECK(chunk.records.size() == 1000);
  const auto ft = FSE_Sequence::calculateFreqTable(chunk);
  const std::vector<char> sequences = concatSeq(chunk);
  CHECK(chunk.tot_reads_length == sequences.size());
  std::vector<std::byte> output_buf(Workspace::compressBoundSequence(chunk.tot_reads_length));
  SequenceEncoder encoder(ft.get());
  SequenceDecoder decoder(ft.get());
  CompressedBuffersDst cbs_dst;
  encoder.startChunk(output_buf);
  for (auto &r : chunk.records) encoder.encodeRecord(r, cbs_dst);
  const std::size_t compressed_size = encoder.endChunk();
  CHECK(compressed_size == 23212);
  output_buf.resize(compressed_size);
  CHECK(cbs_dst.n_count.size() == 2000);
  CHECK(cbs_dst.n_pos.size() == 12526);
  CHECK(concatSeq(chunk) != sequences);  // N -> A happened in place (src/fse_sequence.cpp:45)
  CompressedBuffersSrc cbs_src = convertToSrcBuffers(std::move(cbs_dst));
  decoder.startChunk(output_buf);
  for (auto &r : chunk.records | std::views::reverse) decoder.decodeRecord(r, cbs_src);
  decoder.endChunk();  // batched facade: bases land here
  CHECK(sequences == concatSeq(chunk));
  CHECK(cbs_src.index.n_count == 0 && cbs_src.index.n_pos == 0);
}

static void testFseQuality(const std::string &dir) {
  FastqChunk chunk = loadFastqFileContents(dir + "/without_ns.fastq");
  const auto ft = FSE_Quality::calculateFreqTable(chunk);
  const std::vector<char> qualities = concatQual(chunk);
  std::vector<std::byte> output_buf(Workspace::compressBoundQuality(chunk.tot_reads_length));
  QualityEncoder encoder(ft.get());
  QualityDecoder decoder(ft.get());
  encoder.startChunk(output_buf);
  for (const auto &r : chunk.records) encoder.encodeRecord(r);
  const std::size_t compressed_size = encoder.endChunk();
  CHECK(compressed_size == 35055);
  output_buf.resize(compressed_size);
  for (auto &r : chunk.records) std::fill(r.qualp, r.qualp + r.length, '?');
  decoder.startChunk(output_buf);
  for (auto &r : chunk.records | std::views::reverse) decoder.decodeRecord(r);
  decoder.endChunk();
  CHECK(qualities == concatQual(chunk));
}

static void testEncodeChunk(const std::string &dir) {  // test/workspace_test.cpp:45-69
  FastqChunk chunk_in = loadFastqFileContents(dir + "/without_ns.fastq");
  const FastqChunk original = [&] { FastqChunk c; c.raw_data = chunk_in.raw_data; FastqReader::parseRecords(c); return c; }();
  const DatasetMeta meta(chunk_in);
  CompressionWorkspace cwksp(&meta);
  DecompressionWorkspace dwksp(&meta);
  CompressedBuffersDst cbs;
  cwksp.encodeChunk(chunk_in, cbs);
  CHECK(cbs.seq.size() == 20237 && cbs.qual.size() == 35055);
  FastqChunk chunk_out;
  CompressedBuffersSrc src = convertToSrcBuffers(std::move(cbs));
  dwksp.decodeChunk(chunk_out, src);
  CHECK(chunk_out.records.size() == original.records.size());
  for (std::size_t i = 0; i < original.records.size() && i < chunk_out.records.size(); ++i) {
    CHECK(original.records[i].header() == chunk_out.records[i].header());
    CHECK(original.records[i].seq() == chunk_out.records[i].seq());
    CHECK(original.records[i].qual() == chunk_out.records[i].qual());
  }
  CHECK(chunk_out.raw_data == original.raw_data);
}

static void testQ2Accumulation(const std::string &dir) {
  // the same cbs reused across chunks keeps growing n_count/n_pos
  // and the decoder consumes the tail (src/compressed_buffers.h:58-68)
  FastqChunk a = loadFastqFileContents(dir + "/SRR065390_sub_1.fastq");
  FastqChunk b = loadFastqFileContents(dir + "/SRR065390_sub_2.fastq");
  const std::vector<char> b_orig = b.raw_data;
  const DatasetMeta meta(a);
  CompressionWorkspace cw(&meta);
  DecompressionWorkspace dw(&meta);
  CompressedBuffersDst cbs;
  cw.encodeChunk(a, cbs);
  CHECK(cbs.n_count.size() == 2000 && cbs.n_pos.size() == 12526);
  cw.encodeChunk(b, cbs);
  CHECK(cbs.n_count.size() == 4000 && cbs.n_pos.size() == 12526 + 3200);
  CHECK(cbs.original_size.n_count == 4000);
  FastqChunk out;
  CompressedBuffersSrc src = convertToSrcBuffers(std::move(cbs));
  dw.decodeChunk(out, src);
  CHECK(out.raw_data == b_orig);
  CHECK(src.index.n_count == 2000 && src.index.n_pos == 12526);
}

static void testMeta(const std::string &dir) {  // test/archive_test.cpp:23-50 (meta part)
  FastqChunk chunk = loadFastqFileContents(dir + "/SRR065390_sub_1.fastq");
  const DatasetMeta meta(chunk);
  std::vector<char> bytes;
  DatasetMeta::storeToBytes(meta, bytes);
  CHECK(bytes.size() == 2 + meta.first_header.size() + 3076 + 1081348);
  const DatasetMeta back = DatasetMeta::loadFromBytes(bytes.data(), bytes.size());
  CHECK(back == meta);
  CHECK(meta.first_header == "@SRR065390.1 HWUSI-EAS687_61DAJ:8:1:1055:3384 length=100");
}

static void testBatched(const std::string &dir) {
  const std::vector<char> data = loadFileContents(dir + "/SRR065390_sub_1.fastq");
  FastqChunk whole;
  whole.raw_data = data;
  FastqReader::parseRecords(whole);
  const DatasetMeta meta(whole);
  CompressionWorkspace cw(&meta);
  DecompressionWorkspace dw(&meta);
  std::vector<CompressedBuffersDst> blocks;
  std::size_t consumed = 0;
  cw.encodeChunks(data.data(), data.size(), 30000, true, blocks, &consumed);
  CHECK(consumed == data.size());
  CHECK(blocks.size() >= 8);
  std::vector<CompressedBuffersSrc> srcs;
  for (auto &b : blocks) srcs.push_back(convertToSrcBuffers(std::move(b)));
  std::vector<CompressedBuffersSrc *> ps;
  std::vector<FastqChunk> outs(srcs.size());
  std::vector<FastqChunk *> po;
  for (std::size_t i = 0; i < srcs.size(); ++i) { ps.push_back(&srcs[i]); po.push_back(&outs[i]); }
  dw.decodeChunks(ps, po);
  std::vector<char> joined;
  for (auto &c : outs) joined.insert(joined.end(), c.raw_data.begin(), c.raw_data.end());
  CHECK(joined == data);
}

static void testErrors() {
  FastqChunk c;
  const std::string s = "@r\n" + std::string(70000, 'A') + "\n+\n" + std::string(70000, '!') + "\n";
  c.raw_data.assign(s.begin(), s.end());
  bool threw = false;
  try { FastqReader::parseRecords(c); } catch (const std::runtime_error &e) { threw = std::string(e.what()) == "narrow_cast<>() failed"; }
  CHECK(threw);  // src/utils.h:17-23 via src/fastq_io.cpp:95
}

int main(int argc, char **argv) {
  const std::string dir = argc > 1 ? argv[1] : "tests/data";
  try {
    testFseSequence(dir);
    testFseQuality(dir);
    testEncodeChunk(dir);
    testQ2Accumulation(dir);
    testMeta(dir);
    testBatched(dir);
    testErrors();
  } catch (const std::exception &e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 2;
  }
  std::printf("%d checks, %d failed\n", g_checked, g_failed);
  return g_failed ? 1 : 0;
}
