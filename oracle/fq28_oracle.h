/*
 * fq28_oracle.h -- CPU oracle for the fqcomp28 codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * as the timed CPU baseline -- never behind the fq28_* C ABI.
 *
 * It restates, in plain C, the algorithm of the reference
 * (/root/reference/src, cited per function as file:line) plus the FSE / bit
 * stream primitives of zstd (fork iam28th/zstd @ b010526d, upstream >= 1.5.0),
 * which the reference pulls in at configure time and which are NOT in the
 * reference tree.  Those primitives follow upstream zstd's published
 * algorithm (lib/common/{fse.h,bitstream.h}, lib/compress/fse_compress.c,
 * lib/common/fse_decompress.c) as specified in SURVEY.md Appendix A.
 *
 * Parity pin: the reference holds no golden vectors for this path (its tests
 * are round trips only) and cannot be built here (un-vendored zstd fork,
 * libbsc, CLI11).  The oracle is pinned instead by
 *   (1) tests/test_fse_vs_libzstd.py -- the FSE primitives below re-derive,
 *       byte for byte, the normalised tables and bitstreams that the
 *       container's real libzstd 1.5.5 embeds in zstd frames;
 *   (2) tests/test_oracle_golden.py -- SURVEY.md Appendix C digests (an
 *       independent Python restatement) on the reference's four fixtures;
 *   (3) the reference's own round-trip tests re-expressed on those fixtures.
 * Residual, unpinnable offline: whatever the zstd *fork* changed vs upstream.
 */
#ifndef FQ28_ORACLE_H
#define FQ28_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants (zstd lib/common/fse.h) ------------- */
#define FQ28O_FSE_MIN_TABLELOG 5
#define FQ28O_FSE_MAX_TABLELOG 12
#define FQ28O_FSE_DEFAULT_TABLELOG 11

#define FQ28O_SEQ_MODELS 256    /* src/fse_sequence.h:66 */
#define FQ28O_SEQ_ALPHABET 4    /* src/fse_sequence.h:34-35 */
#define FQ28O_SEQ_INITIAL_CTX 0xD7u /* src/fse_sequence.h:41-63 */
#define FQ28O_QUAL_MODELS 8192  /* src/fse_quality.h:31 */
#define FQ28O_QUAL_ALPHABET 64  /* src/fse_quality.h:21-22 */
#define FQ28O_QUAL_OFFSET 33    /* src/fse_quality.h:24 */

/* error codes shared with the product ABI (include/fq28.h) */
#define FQ28O_OK 0
#define FQ28O_ERR_FORMAT (-2)   /* not 4-line FASTQ / qual len != seq len  */
#define FQ28O_ERR_ALPHABET (-3) /* base not in ACGTN or quality > Q63      */
#define FQ28O_ERR_SHORT (-4)    /* read shorter than 3 (reference UB, Q4)   */
#define FQ28O_ERR_LONG (-5)     /* line longer than 65535 (narrow_cast)     */
#define FQ28O_ERR_CAP (-6)      /* output capacity too small                */

/* FreqTable<256,4> / FreqTable<8192,64> raw images, src/fse_common.hpp:147-174
 * (sizeof 3076 / 1081348, dumped raw by src/prepare.cpp:18-20) */
typedef struct {
  int16_t norm[FQ28O_SEQ_MODELS][FQ28O_SEQ_ALPHABET];
  uint32_t logs[FQ28O_SEQ_MODELS];
  uint32_t max_log;
} fq28o_ft_seq;
typedef struct {
  int16_t norm[FQ28O_QUAL_MODELS][FQ28O_QUAL_ALPHABET];
  uint32_t logs[FQ28O_QUAL_MODELS];
  uint32_t max_log;
} fq28o_ft_qual;

/* FastqRecord (src/defs.h:22-32) with offsets instead of pointers */
typedef struct {
  uint64_t hdr_off, seq_off, qual_off;
  uint32_t hdr_len, len;
} fq28o_rec;

/* ---- FSE primitives (zstd) -------------------- */
unsigned fq28o_optimal_table_log(unsigned max_table_log, size_t src_size,
                                 unsigned max_sv);
/* returns table_log, 0 for the rle case, (size_t)-1 on error */
size_t fq28o_normalize_count(int16_t *norm, unsigned table_log,
                             const unsigned *count, size_t total,
                             unsigned max_sv, unsigned use_low_prob);
/* cell[u] = symbol occupying table cell u (the spread), T = 1<<table_log */
void fq28o_spread(uint8_t *cell, const int16_t *norm, unsigned max_sv,
                  unsigned table_log);
/* CTable: state_table[T] (u16 next-state values), per-symbol transforms */
void fq28o_build_ctable(uint16_t *state_table, int32_t *delta_find_state,
                        uint32_t *delta_nb_bits, const int16_t *norm,
                        unsigned max_sv, unsigned table_log);
/* DTable cell = newState | symbol<<16 | nbBits<<24 (FSE_decode_t, LE) */
void fq28o_build_dtable(uint32_t *cells, const int16_t *norm, unsigned max_sv,
                        unsigned table_log);

/* ---- parsing / chunking --------------------------------------------------- */
/* FastqReader::parseRecords src/fastq_io.cpp:67-125.  Returns the offset at
 * which the first incomplete record starts (== size when none).  recs may be
 * NULL (count only).  *err is set to FQ28O_ERR_LONG / FQ28O_ERR_FORMAT /
 * FQ28O_ERR_CAP (never by the reference, which asserts or throws instead). */
size_t fq28o_parse_records(const char *data, size_t size, fq28o_rec *recs,
                           size_t cap, size_t *n_recs, int *err);
/* chunk boundary rule of FastqReader::readNextChunk src/fastq_io.cpp:23-65:
 * offs[0]=0, offs[k+1] = end of last complete record in [offs[k], offs[k]+R).
 * Returns number of chunks (offs has n+1 entries), or <0 on error. */
long fq28o_split_chunks(const char *data, size_t size, size_t reading_size,
                        uint64_t *offs, size_t cap);

/* ---- frequency tables ----------------------------------------------------- */
/* raw counts WITHOUT the +1 prior (so partial histograms can be summed);
 * src/fse_sequence.cpp:145-169 and src/fse_quality.cpp:69-97 */
int fq28o_hist_seq(const char *data, const fq28o_rec *recs, size_t n,
                   uint32_t *counts /*[256*4]*/);
int fq28o_hist_qual(const char *data, const fq28o_rec *recs, size_t n,
                    uint32_t *counts /*[8192*64]*/);
/* +1 prior, then makeNormalizedFreqTable src/fse_common.hpp:179-200 */
void fq28o_make_ft_seq(const uint32_t *counts, fq28o_ft_seq *ft);
void fq28o_make_ft_qual(const uint32_t *counts, fq28o_ft_qual *ft);

/* ---- codecs --------------------------------------------------------------- */
typedef struct fq28o_codec fq28o_codec; /* CTables+DTables for one FreqTable */
fq28o_codec *fq28o_codec_seq(const fq28o_ft_seq *ft);
fq28o_codec *fq28o_codec_qual(const fq28o_ft_qual *ft);
void fq28o_codec_free(fq28o_codec *);

/* worst-case bounds, src/workspace.h:21-35 */
size_t fq28o_bound_seq(size_t tot_reads_length);
size_t fq28o_bound_qual(size_t tot_reads_length);

/* startChunk + encodeRecord over all records + endChunk.
 * seq: src/fse_sequence.cpp:35-112, src/fse_common.hpp:77-90.  `data` is
 * mutated (N -> A).  n_count gets n u16, n_pos gets *n_npos u16 deltas.
 * Returns stream size, 0 on overflow (BIT_closeCStream), <0 on error. */
long fq28o_encode_seq(const fq28o_codec *c, char *data, const fq28o_rec *recs,
                      size_t n, uint8_t *dst, size_t cap, uint16_t *n_count,
                      uint16_t *n_pos, size_t n_pos_cap, size_t *n_npos);
/* src/fse_quality.cpp:5-53 */
long fq28o_encode_qual(const fq28o_codec *c, const char *data,
                       const fq28o_rec *recs, size_t n, uint8_t *dst,
                       size_t cap);
/* startChunk + decodeRecord for records n-1..0 + endChunk.
 * src/fse_sequence.cpp:114-143 / src/fse_quality.cpp:55-67,
 * src/fse_common.hpp:130-141.  n_count/n_pos are the chunk's own entries
 * (the reference consumes them from the back).  Returns 0, or <0 if the
 * stream is not exactly consumed. */
int fq28o_decode_seq(const fq28o_codec *c, const uint8_t *src, size_t size,
                     char *data, const fq28o_rec *recs, size_t n,
                     const uint16_t *n_count, const uint16_t *n_pos,
                     size_t n_npos);
int fq28o_decode_qual(const fq28o_codec *c, const uint8_t *src, size_t size,
                      char *data, const fq28o_rec *recs, size_t n);

/* decodeChunk pass 1, src/workspace.cpp:62-80: lays out
 * header '\n' seq-slot '\n' '+' '\n' qual-slot '\n' per record.  headers =
 * concatenated header lines (with '@', no '\n'), hdr_lens[i] their lengths.
 * Returns bytes written. */
size_t fq28o_layout_chunk(char *out, size_t cap, const char *headers,
                          const uint32_t *hdr_lens, const uint16_t *readlens,
                          size_t n, fq28o_rec *recs);

/* ---- multi-threaded CPU baseline (reference threading model:
 * one chunk per worker thread, src/process.cpp:40-70,84-105) ---------------- */
typedef struct {
  double t_analyze_s;    /* parse sample + hist + normalise (src/prepare.cpp:42-47) */
  double t_compress_s;   /* parse + encode all chunks, wall */
  double t_decompress_s; /* layout + decode all chunks, wall */
  uint64_t fastq_bytes, seq_bytes, qual_bytes, n_records, n_chunks;
  uint64_t checksum;     /* FNV-1a over all seq+qual streams in chunk order */
  int roundtrip_ok;      /* decoded FASTQ == input */
  int err;
} fq28o_bench_result;
/* FNV-1a, chained: the checksum fq28o_bench computes over the streams */
uint64_t fq28o_fnv1a(const uint8_t *p, size_t n, uint64_t h);
int fq28o_bench(const char *fastq, size_t size, size_t sample_bytes,
                size_t reading_size, int threads, int do_decompress,
                fq28o_bench_result *res);

#ifdef __cplusplus
}
#endif
#endif
