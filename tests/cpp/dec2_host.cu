// dec2_host.cu -- TEST INFRASTRUCTURE.  Compiles the product's per-stream decoders
// (fqcomp28_b200/csrc/fq28_dec2.cuh, __host__ __device__) for the CPU so that
// tests/test_dec2_host.py can check the algorithm -- cached cells, deferred
// refresh, STALE protocol, zero-bit runs, slow path -- against the oracle without
// a GPU.  Shared memory is a byte array and there is one lane.  The derived
// tables (W tables, dense alphabet, run tables) are rebuilt here on the host with
// the same rules as the device kernels in fq28_tables.cu.  Never linked into
// libfq28.so.
#define DEC2_STATS 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../fqcomp28_b200/csrc/fq28_dec2.cuh"
#include "../../oracle/fq28_oracle.h"

namespace fq28 { namespace dec2 {
thread_local uint8_t *g_host_smem = nullptr;
thread_local unsigned long long g_stats[12];
thread_local HostAsync g_host_async;
} }
using namespace fq28::dec2;

namespace {

struct Tables {
  std::vector<uint32_t> logs, logsuf, dtab_fix;
  unsigned n = 0;
};

void build_dtabs(Tables &t, const int16_t *norm, const uint32_t *logs, unsigned n_models, unsigned alphabet) {
  t.n = n_models;
  t.logs.assign(logs, logs + n_models);
  t.logsuf.assign(n_models + 1, 0);
  uint32_t suf = 0;
  for (unsigned c = n_models; c > 0; --c) { t.logsuf[c - 1] = suf; suf += logs[c - 1]; }
  t.logsuf[n_models] = suf;
  t.dtab_fix.assign((size_t)n_models << TAB_LOG, 0);
  for (unsigned c = 0; c < n_models; c++)
    fq28o_build_dtable(&t.dtab_fix[(size_t)c << TAB_LOG], norm + (size_t)c * alphabet, alphabet - 1, logs[c]);
}

void fill_args(StreamArgs &a, const uint8_t *src, uint32_t len, const uint16_t *readlens, const uint16_t *hdr_lens,
               uint32_t n_rec, std::vector<uint32_t> &recscan, char *out, const Tables &t, const uint32_t *wtab) {
  recscan.assign(n_rec + 1, 0);
  for (uint32_t r = 0; r < n_rec; r++) recscan[r + 1] = recscan[r] + hdr_lens[r] + 2u * readlens[r] + 5u;
  a.src = src; a.len = len; a.rec0 = 0; a.n_rec = n_rec;
  a.readlens = readlens; a.hdr_lens = hdr_lens; a.recscan = recscan.data();
  a.out = out; a.logs = t.logs.data(); a.logsuf = t.logsuf.data(); a.wtab = wtab;
  a.ring = nullptr; a.live = true;
}

}  // namespace

// test switch: the tables have run slots (cid carries them) but the decoder is told to ignore them,
// which is what the product does when the run tables would not leave room for the streams
static bool g_no_runs = false;
// test switch: the windowed layout of the cached cells (what the product uses for many-valued qualities)
static bool g_win = false;

extern "C" {

void dec2h_set_no_runs(int on) { g_no_runs = on != 0; }
void dec2h_set_win(int on) { g_win = on != 0; }

// event counters since the last call: [0] seq drains, [1] seq inline refreshes, [2] qual drains,
// [3] zero-bit runs, [4] single steps in a run context, [5] slow-path entries, [6] slow-path
// symbols in dense contexts, [7] slow-path symbols in contexts outside the dense set,
// [8] sequence blocks redone step by step, [9] sequence blocks committed at once
void dec2h_stats(unsigned long long *out) {
  for (int i = 0; i < 12; i++) { out[i] = g_stats[i]; g_stats[i] = 0; }
}

// ft = FreqTable<256,4> image.  The stream is copied to `misalign` bytes past an 8-byte boundary.
int dec2h_seq(const uint8_t *ft, const uint8_t *stream, uint32_t len, unsigned misalign, const uint16_t *readlens,
              const uint16_t *hdr_lens, uint32_t n_rec, char *out) {
  const int16_t *norm = reinterpret_cast<const int16_t *>(ft);
  const uint32_t *logs = reinterpret_cast<const uint32_t *>(ft + 256 * 4 * 2);
  Tables t;
  build_dtabs(t, norm, logs, 256, 4);
  std::vector<uint32_t> wtab((size_t)256 << TAB_LOG, 0);
  for (unsigned c = 0; c < 256; c++)
    for (unsigned u = 0; u < (1u << logs[c]); u++)
      wtab[((size_t)c << TAB_LOG) + u] = make_w_seq(t.dtab_fix[((size_t)c << TAB_LOG) + u], c);
  // shared memory image: 4 homopolymer tables | S | scratch
  const uint32_t ht = 0, sb = 4 * (4u << TAB_LOG);
  std::vector<uint8_t> smem(sb + 1024 + 64, 0);
  for (unsigned j = 0; j < 4; j++)
    memcpy(&smem[ht + j * (4u << TAB_LOG)], &wtab[(size_t)(j * 0x55u) << TAB_LOG], 4u << TAB_LOG);
  g_host_smem = smem.data();
  std::vector<uint64_t> buf((len + 64) / 8 + 4, 0);
  uint8_t *src = reinterpret_cast<uint8_t *>(buf.data()) + 8 + (misalign & 7);
  memcpy(src, stream, len);
  StreamArgs a;
  std::vector<uint32_t> recscan;
  fill_args(a, src, len, readlens, hdr_lens, n_rec, recscan, out, t, wtab.data());
  g_host_async = HostAsync();
  const bool ok = decode_seq_stream(a, sb, ht);
  g_host_smem = nullptr;
  return ok ? 0 : -8;
}

// ft = FreqTable<8192,64> image.  stats (may be NULL): [0] = |V|, [1] = run slots
int dec2h_qual(const uint8_t *ft, const uint8_t *stream, uint32_t len, unsigned misalign, const uint16_t *readlens,
               const uint16_t *hdr_lens, uint32_t n_rec, char *out, uint32_t *stats) {
  const int16_t *norm = reinterpret_cast<const int16_t *>(ft);
  const uint32_t *logs = reinterpret_cast<const uint32_t *>(ft + (size_t)8192 * 64 * 2);
  Tables t;
  build_dtabs(t, norm, logs, 8192, 64);
  // touched contexts (k_qual_cid): table differs from the prior-only pattern
  std::vector<uint16_t> cid(8192, 0xFFFF);
  bool inV[64] = {false};
  inV[0] = true;
  unsigned nt = 0;
  for (unsigned c = 0; c < 8192; c++) {
    bool touched = logs[c] != 7;
    for (unsigned s = 0; s < 64 && !touched; s++) touched = norm[(size_t)c * 64 + s] != 2;
    if (touched) { cid[c] = (uint16_t)nt++; inV[c & 63] = true; }
  }
  uint8_t rk[64];
  unsigned nv = 0;
  for (unsigned qv = 0; qv < 64; qv++) rk[qv] = inV[qv] ? (uint8_t)nv++ : (uint8_t)0xFF;
  uint8_t vq[64] = {0};
  for (unsigned qv = 0; qv < 64; qv++) if (rk[qv] != 0xFF) vq[rk[qv]] = (uint8_t)qv;
  // dense W table
  const unsigned n_dense = 2 * nv * 64;
  std::vector<uint32_t> wtab((size_t)n_dense << TAB_LOG, 0);
  for (unsigned rm = 0; rm < nv; rm++)
    for (unsigned eq = 0; eq < 2; eq++)
      for (unsigned rq = 0; rq < nv; rq++) {
        const unsigned cx = ((unsigned)vq[rm] << 6) + vq[rq] + (eq << 12);
        const unsigned d = qual_dense_id(rm, eq, rq);
        for (unsigned u = 0; u < (1u << logs[cx]); u++)
          wtab[((size_t)d << TAB_LOG) + u] = make_w_qual(t.dtab_fix[((size_t)cx << TAB_LOG) + u], rk);
      }
  if (g_win) {
    // k_qual_win / k_qual_wtabw: per row the columns spanning its touched contexts + one out-of-window word
    uint32_t rowx[128] = {0}, rowy[128] = {0};
    unsigned total = 0;
    for (unsigned row = 0; row < 2 * nv; row++) {
      const unsigned rm = row >> 1, eq = row & 1u;
      unsigned lo = 64, hi = 0;
      bool any = false;
      for (unsigned rq = 0; rq < nv; rq++) {
        const unsigned cx = ((unsigned)vq[rm] << 6) + vq[rq] + (eq << 12);
        if (cid[cx] != 0xFFFF) { lo = rq < lo ? rq : lo; hi = rq > hi ? rq : hi; any = true; }
      }
      if (row == 1) { lo = 0; any = true; }
      const unsigned width = any ? hi - lo + 1 : 0;
      rowx[row] = total * 4;
      rowy[row] = (any ? lo * 4 : 0u) | ((width * 4) << 16);
      total += width + 1;
    }
    std::vector<uint32_t> wtabw((size_t)total << TAB_LOG, 0);
    for (unsigned rm = 0; rm < nv; rm++)
      for (unsigned eq = 0; eq < 2; eq++)
        for (unsigned rq = 0; rq < nv; rq++) {
          const unsigned row = rm * 2 + eq;
          const uint32_t rel = rq * 4 - (rowy[row] & 0xFFFFu);
          if (rel >= (rowy[row] >> 16)) continue;
          const unsigned cx = ((unsigned)vq[rm] << 6) + vq[rq] + (eq << 12);
          const unsigned d = (rowx[row] + rel) >> 2;
          for (unsigned u = 0; u < (1u << logs[cx]); u++)
            wtabw[((size_t)d << TAB_LOG) + u] = make_w_qual(t.dtab_fix[((size_t)cx << TAB_LOG) + u], rk);
        }
    // shared memory image: rk | (zc) | row descriptors | S
    const uint32_t rk_a = 0, row_a = 128, sbw = 128 + 1024;
    std::vector<uint8_t> smemw(sbw + total * 4 + 64, 0);
    memcpy(&smemw[rk_a], rk, 64);
    for (unsigned r = 0; r < 128; r++) {
      memcpy(&smemw[row_a + r * 8], &rowx[r], 4);
      memcpy(&smemw[row_a + r * 8 + 4], &rowy[r], 4);
    }
    g_host_smem = smemw.data();
    std::vector<uint64_t> bufw((len + 64) / 8 + 4, 0);
    uint8_t *srcw = reinterpret_cast<uint8_t *>(bufw.data()) + 8 + (misalign & 7);
    memcpy(srcw, stream, len);
    std::vector<uint16_t> coldw(8192, 0);
    StreamArgs aw;
    std::vector<uint32_t> recscanw;
    fill_args(aw, srcw, len, readlens, hdr_lens, n_rec, recscanw, out, t, wtabw.data());
    QualShared qsw{rk_a, 128, 64, 0u, row_a, 2 * nv};
    g_host_async = HostAsync();
    const bool okw = decode_qual_stream<true>(aw, qsw, sbw, t.dtab_fix.data(), cid.data(), coldw.data());
    g_host_smem = nullptr;
    if (stats) { stats[0] = nv; stats[1] = total; }
    return okw ? 0 : -8;
  }
  // zero-bit run slots (k_qual_zrun): ctx(d,d,d) with a dominant d, high qualities first
  unsigned nz = 0, zsym[4];
  for (int d = 63; d >= 0 && nz < 4; --d) {
    const unsigned cx = qual_ctx13((unsigned)d, (unsigned)d, (unsigned)d);
    const int T = 1 << logs[cx];
    if (cid[cx] != 0xFFFF && norm[(size_t)cx * 64 + d] > T / 2) {
      cid[cx] = (uint16_t)(cid[cx] | ((nz + 1) << 13));
      zsym[nz++] = (unsigned)d;
    }
  }
  // shared memory image: rk | zc | run tables (8 bytes per state) | S
  const uint32_t rk_a = 0, zc_a = 64, zq_a = 128;
  const uint32_t sb = (zq_a + nz * ZQ_SLOT_BYTES + 255) & ~255u;
  std::vector<uint8_t> smem(sb + n_dense * 4 + 64, 0);
  memcpy(&smem[rk_a], rk, 64);
  for (unsigned j = 0; j < nz; j++) {
    const unsigned d = zsym[j], cx = qual_ctx13(d, d, d), T = 1u << logs[cx];
    const uint32_t ch = d + QUAL_CHAR0;
    memcpy(&smem[zc_a + j * 4], &ch, 4);
    const unsigned dd = qual_dense_id(rk[d], 1, rk[d]);
    for (unsigned x = 0; x < (1u << TAB_LOG); x++) {
      unsigned k = 0, y = x;
      while (x < T && k < 15) {
        const uint32_t e = t.dtab_fix[((size_t)cx << TAB_LOG) + y];
        if (((e >> 16) & 63u) != d || (e >> 24) != 0) break;
        y = e & 0xFFFFu;
        k++;
      }
      const uint32_t e2[2] = {wtab[((size_t)dd << TAB_LOG) + x], make_zq_hi(k, y, j)};
      memcpy(&smem[zq_a + j * ZQ_SLOT_BYTES + x * 8], e2, 8);
    }
  }
  g_host_smem = smem.data();
  std::vector<uint64_t> buf((len + 64) / 8 + 4, 0);
  uint8_t *src = reinterpret_cast<uint8_t *>(buf.data()) + 8 + (misalign & 7);
  memcpy(src, stream, len);
  std::vector<uint16_t> cold(8192, 0);
  StreamArgs a;
  std::vector<uint32_t> recscan;
  fill_args(a, src, len, readlens, hdr_lens, n_rec, recscan, out, t, wtab.data());
  QualShared qs{rk_a, zq_a, zc_a, g_no_runs ? 0u : nz, 0u, 0u};
  g_host_async = HostAsync();
  const bool ok = decode_qual_stream<false>(a, qs, sb, t.dtab_fix.data(), cid.data(), cold.data());
  g_host_smem = nullptr;
  if (stats) { stats[0] = nv; stats[1] = nz; }
  return ok ? 0 : -8;
}

}  // extern "C"

#ifdef DEC2_FUZZ_MAIN
// Corrupt-stream fuzzer, built with -fsanitize=address,undefined by tests/test_dec2_host.py:
// the decoders must never touch memory outside their shared-memory image, their tables or the
// chunk's output, whatever the stream bits say (compute-sanitizer is not available on the GPU
// pool; this is the same source).  Input: a blob written by the test (FreqTable image, one
// chunk's stream, its record lengths); the stream is decoded intact, then with random bit
// flips, byte overwrites and truncations.  Exit code 0 unless the intact stream fails.
#include <stdio.h>
int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  uint32_t hd[6];
  if (fread(hd, 4, 6, f) != 6 || hd[0] != 0x46513238u) return 2;
  const uint32_t kind = hd[1], ft_bytes = hd[2], len = hd[3], n_rec = hd[4], out_bytes = hd[5];
  std::vector<uint8_t> ft(ft_bytes), stream(len);
  std::vector<uint16_t> readlens(n_rec), hdr_lens(n_rec);
  if (fread(ft.data(), 1, ft_bytes, f) != ft_bytes || fread(stream.data(), 1, len, f) != len ||
      fread(readlens.data(), 2, n_rec, f) != n_rec || fread(hdr_lens.data(), 2, n_rec, f) != n_rec)
    return 2;
  fclose(f);
  const unsigned iters = (unsigned)atoi(argv[2]);
  auto run = [&](const std::vector<uint8_t> &s, uint32_t l) -> int {
    std::vector<char> out(out_bytes, 0);   // exactly the chunk's bytes: an overrun is a heap overflow
    if (kind == 0) return dec2h_seq(ft.data(), s.data(), l, 0, readlens.data(), hdr_lens.data(), n_rec, out.data());
    return dec2h_qual(ft.data(), s.data(), l, 0, readlens.data(), hdr_lens.data(), n_rec, out.data(), nullptr);
  };
  unsigned long long rng = 0x9E3779B97F4A7C15ull ^ len;
  auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  unsigned ok = 0, bad = 0;
  for (int layout = 0; layout < (kind == 1 ? 2 : 1); layout++) {
    dec2h_set_win(layout);
    if (run(stream, len) != 0) { fprintf(stderr, "intact stream failed (layout %d)\n", layout); return 1; }
    for (unsigned it = 0; it < iters; it++) {
      std::vector<uint8_t> s = stream;
      uint32_t l = len;
      const unsigned mode = (unsigned)(next() % 4);
      if (mode == 0) {                       // a few bit flips anywhere
        for (unsigned k = 0, n = 1 + (unsigned)(next() % 8); k < n; k++) s[next() % len] ^= (uint8_t)(1u << (next() % 8));
      } else if (mode == 1) {                // the tail (initial states, end mark) overwritten
        for (unsigned k = 0, n = 1 + (unsigned)(next() % 16); k < n && k < len; k++) s[len - 1 - k] = (uint8_t)next();
      } else if (mode == 2) {                // truncated
        l = (uint32_t)(next() % len);
        if (l == 0) l = 1;
      } else {                               // a run of random bytes
        const size_t a = next() % len, n = 1 + next() % 64;
        for (size_t k = a; k < len && k < a + n; k++) s[k] = (uint8_t)next();
      }
      (run(s, l) == 0 ? ok : bad)++;
    }
  }
  dec2h_set_win(0);
  printf("fuzz: %u corrupted streams decoded to the end mark, %u rejected\n", ok, bad);
  return 0;
}
#endif
