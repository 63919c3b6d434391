/*
 * fq28.h -- C ABI of the B200-native fqcomp28 codec hot path (libfq28.so).
 *
 * The reference (iam28th/fqcomp28) has no FFI layer: its hot path sits behind
 * ordinary C++ classes called by src/process.cpp.  This header is the boundary
 * a drop-in replacement of that path binds to; the C++ facade that keeps the
 * reference's own class names on top of it is fqcomp28_b200/host/fqcomp28_gpu.hpp,
 * and INTEGRATION.md shows the reference-side patch.
 *
 * Conventions: opaque handle, int status (0 = OK, <0 = error, text via
 * fq28_last_error), no exceptions, caller-owned buffers, plain pointers and
 * sizes.  Every entry point runs on the CUDA device selected at fq28_create;
 * there is NO CPU fallback -- without a usable device fq28_create fails.
 *
 * Entry points come in pairs where it matters for measurement:
 *   *_dev  : inputs/outputs already resident in device memory (HBM)
 *   (none) : host buffers; the call performs the H2D / D2H copies itself.
 */
#ifndef FQ28_H
#define FQ28_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FQ28_OK 0
#define FQ28_ERR_CUDA (-1)     /* CUDA runtime error (see fq28_last_error)      */
#define FQ28_ERR_FORMAT (-2)   /* not 4-line FASTQ, qual len != seq len, record > R */
#define FQ28_ERR_ALPHABET (-3) /* base not in ACGTN / quality above Q63 */
#define FQ28_ERR_SHORT (-4)    /* read shorter than 3: UB in the reference (Q4)  */
#define FQ28_ERR_LONG (-5)     /* line > 65535: narrow_cast throws, src/fastq_io.cpp:95 */
#define FQ28_ERR_CAP (-6)      /* caller buffer too small                        */
#define FQ28_ERR_ARG (-7)      /* bad argument / tables not loaded               */
#define FQ28_ERR_STREAM (-8)   /* corrupt stream: not exactly consumed (src/fse_common.hpp:141) */

#define FQ28_SEQ_MODELS 256     /* src/fse_sequence.h:66 */
#define FQ28_SEQ_ALPHABET 4
#define FQ28_QUAL_MODELS 8192   /* src/fse_quality.h:31  */
#define FQ28_QUAL_ALPHABET 64
#define FQ28_FT_SEQ_BYTES 3076     /* sizeof(FreqTable<256,4>),   src/fse_common.hpp:147-174 */
#define FQ28_FT_QUAL_BYTES 1081348 /* sizeof(FreqTable<8192,64>) */
#define FQ28_MAX_SLAB ((size_t)0xFFFFFF00u) /* one call handles < 4 GiB of FASTQ */

typedef struct fq28_handle fq28_handle;

/* -- lifetime ---------------------------------------------------------------
 * One handle per worker (= per GPU); replaces the per-thread
 * CompressionWorkspace / DecompressionWorkspace of src/process.cpp:49-67,94-103.
 * Not thread-safe; use one handle per host thread. */
int fq28_create(int device, fq28_handle **out);
int fq28_device_count(void);   /* visible CUDA devices (0 when there is none) */
void fq28_destroy(fq28_handle *h);
const char *fq28_last_error(const fq28_handle *h);
/* Launch on an existing cudaStream_t (e.g. the caller's current stream). */
int fq28_set_stream(fq28_handle *h, void *cuda_stream);
/* Number of kernels this handle has launched so far (bench `gpu_launches`). */
uint64_t fq28_launch_count(const fq28_handle *h);

/* -- record splitting -------------------------------------------------------
 * FastqReader::parseRecords, src/fastq_io.cpp:67-125.  Splits a slab that
 * starts at a record start ('@') into records.  Outputs (each may be NULL):
 * hdr_off/seq_off/qual_off [u32 offsets into the slab], hdr_len/len [u16].
 * *consumed = offset where the first incomplete record starts (== n_bytes if
 * none), which is parseRecords' return value. */
int fq28_parse(fq28_handle *h, const char *fastq, size_t n_bytes,
               uint32_t *hdr_off, uint32_t *seq_off, uint32_t *qual_off,
               uint16_t *hdr_len, uint16_t *len, size_t cap,
               size_t *n_records, size_t *consumed);

/* Chunk boundary rule of FastqReader::readNextChunk, src/fastq_io.cpp:23-65:
 * offs[0] = 0, offs[k+1] = end of the last complete record in
 * [offs[k], min(offs[k]+reading_size, n_bytes)).  `eof` != 0 means the slab
 * ends the file (the final short window is emitted and a trailing partial
 * record dropped); with eof == 0 only chunks whose whole window lies inside
 * the slab are emitted.  offs needs *n_chunks+1 entries. */
int fq28_split(fq28_handle *h, const char *fastq, size_t n_bytes,
               size_t reading_size, int eof, uint64_t *offs, size_t cap,
               size_t *n_chunks);

/* -- frequency tables -------------------------------------------------------
 * FSE_Sequence::calculateFreqTable src/fse_sequence.cpp:145-169 and
 * FSE_Quality::calculateFreqTable src/fse_quality.cpp:69-97, split in two so
 * that partial histograms can be summed (NCCL allreduce) before normalising:
 *   fq28_hist          : raw u32 counts WITHOUT the +1 prior, accumulated into
 *                        seq_counts[256*4] / qual_counts[8192*64] (host)
 *   fq28_hist_dev      : same, slab and count buffers in device memory
 *                        (counts are accumulated, zero them first)
 *   fq28_build_tables  : +1 prior, makeNormalizedFreqTable
 *                        (src/fse_common.hpp:179-200), FSE_Encoder/FSE_Decoder
 *                        ctors (:46-71,:107-127); returns the raw FreqTable
 *                        images that src/prepare.cpp:18-20 dumps in the archive
 *   fq28_load_tables   : decompression side, from the archive's raw images
 *                        (src/prepare.cpp:23-40). */
int fq28_hist(fq28_handle *h, const char *fastq, size_t n_bytes,
              uint32_t *seq_counts, uint32_t *qual_counts);
int fq28_hist_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes,
                  uint32_t *d_seq_counts, uint32_t *d_qual_counts);
int fq28_build_tables(fq28_handle *h, const uint32_t *seq_counts,
                      const uint32_t *qual_counts, void *ft_seq_out,
                      void *ft_qual_out);
int fq28_build_tables_dev(fq28_handle *h, const uint32_t *d_seq_counts,
                          const uint32_t *d_qual_counts, void *ft_seq_out,
                          void *ft_qual_out);
int fq28_load_tables(fq28_handle *h, const void *ft_seq, const void *ft_qual);

/* -- compression ------------------------------------------------------------
 * CompressionWorkspace::encodeChunk src/workspace.cpp:14-45 for every chunk of
 * a slab at once (headers and libbsc excluded: host-side, out of path).
 * Per chunk k the result is described by fq28_chunk_info; the payload lives in
 * the arenas. */
typedef struct {
  uint64_t fastq_off;  /* chunk start in the slab                               */
  uint32_t total;      /* cb_original_sizes_t::total  (chunk bytes)             */
  uint32_t n_records;  /* cb_original_sizes_t::n_records                        */
  uint64_t rec_off;    /* first record: index into readlens / n_count / hdr_lens */
  uint64_t seq_off;    /* byte offset of the seq stream in the seq arena        */
  uint64_t qual_off;   /* byte offset of the qual stream in the qual arena      */
  uint32_t seq_len;    /* CompressedBuffers::seq.size()                         */
  uint32_t qual_len;   /* CompressedBuffers::qual.size()                        */
  uint64_t n_pos_off;  /* index of the chunk's first n_pos entry                */
  uint32_t n_pos_len;  /* number of u16 n_pos entries of this chunk             */
  uint32_t hdr_bytes;  /* sum of the chunk's header line lengths                */
  uint64_t hdr_off;    /* byte offset of the chunk's first header in `headers`  */
} fq28_chunk_info;

typedef struct {
  uint8_t *seq;       size_t seq_cap;      /* bytes */
  uint8_t *qual;      size_t qual_cap;     /* bytes */
  uint16_t *readlens; size_t readlens_cap; /* entries; CompressedBuffers::readlens */
  uint16_t *n_count;  size_t n_count_cap;  /* entries; this chunk's own counts (Q2: the
                                              facade replicates the accumulation) */
  uint16_t *n_pos;    size_t n_pos_cap;    /* entries */
  uint16_t *hdr_lens; size_t hdr_lens_cap; /* entries; FastqRecord::header_length, may be NULL */
  uint8_t *headers;   size_t headers_cap;  /* bytes; header lines ('@'.., no '\n') back to back:
                                              the input of the host header tokeniser, may be NULL */
} fq28_enc_arenas;

typedef struct {
  uint64_t n_chunks, n_records, n_symbols;
  uint64_t seq_bytes, qual_bytes, n_pos_entries;
  uint64_t consumed;   /* bytes of the slab covered by the emitted chunks */
  uint64_t hdr_bytes;  /* total header bytes of the emitted chunks         */
} fq28_enc_summary;

/* Host-buffer entry point: splits `fastq` with the rule of fq28_split and
 * encodes every emitted chunk.  Arena payloads are written back to the host;
 * stream offsets inside the arenas are 16-byte aligned.
 * sample_bytes > 0: first run analyzeDataset (src/prepare.cpp:42-47) on the
 * leading sample_bytes of this slab -- records wholly inside
 * [0, min(sample_bytes, n_bytes)) -- and build the tables from it, which is
 * what `fqcomp28 c` does when it opens the archive (src/archive.cpp:13-20);
 * the raw FreqTable images are returned in ft_seq_out / ft_qual_out (may be
 * NULL).  sample_bytes == 0: use the tables already built / loaded.
 * out == NULL: nothing is copied back; the caller reads the sizes from *summary,
 * sizes its arenas exactly and calls fq28_compress_fetch (compressBound-sized
 * arenas are ~2.8x the slab: allocating them per call costs more than the codec). */
int fq28_compress(fq28_handle *h, const char *fastq, size_t n_bytes,
                  size_t sample_bytes, size_t reading_size, int eof,
                  void *ft_seq_out, void *ft_qual_out,
                  const fq28_enc_arenas *out, fq28_chunk_info *infos,
                  size_t infos_cap, fq28_enc_summary *summary);
/* Device-resident entry point: slab already in HBM, results stay in HBM
 * (owned by the handle, valid until the next call); only infos/summary come
 * back to the host. */
int fq28_compress_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes,
                      size_t sample_bytes, size_t reading_size, int eof,
                      void *ft_seq_out, void *ft_qual_out,
                      fq28_chunk_info *infos, size_t infos_cap,
                      fq28_enc_summary *summary);
/* -- one file, several GPUs --------------------------------------------------
 * The chunk boundaries of a file are a sequential recurrence (src/fastq_io.cpp:23-65):
 * a slab can only be cut where the previous one stopped.  fq28_plan runs just
 * parseRecords + the boundary walk of a slab and returns *consumed, the offset at
 * which the next slab starts -- known a millisecond after the data is on the GPU, so
 * the next GPU can start while this one encodes.  The plan is kept: fq28_compress
 * (or _dev) with the same slab, reading size and eof, and sample_bytes == 0, reuses it.
 * fq28_stage starts the host-to-device copy of a host range early (before the
 * slab's exact start is known); host-buffer calls whose input lies inside the staged
 * range take it from there instead of copying again (a snapshot: it stays in use until
 * the next fq28_stage; fq28_stage(h, NULL, 0) forgets it). */
int fq28_stage(fq28_handle *h, const char *fastq, size_t n_bytes);
int fq28_plan(fq28_handle *h, const char *fastq, size_t n_bytes, size_t reading_size,
              int eof, uint64_t *consumed, size_t *n_chunks);
int fq28_plan_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes, size_t reading_size,
                  int eof, uint64_t *consumed, size_t *n_chunks);
/* A slab that starts at a record boundary INSIDE a chunk owned by the previous slab (one rank
 * per GPU, each holding its own record range of the file plus reading_size bytes of lookahead):
 * fq28_preparse_dev builds the record table as soon as the data is there -- it does not depend on
 * the other ranks -- and fq28_plan_cut_dev then only walks the boundaries, starting at
 * `first_cut`, the slab-relative offset at which the previous rank's last chunk ends (0: the slab
 * starts a chunk).  With first_cut > 0 the head [0, first_cut) is emitted as chunk 0 and must be
 * dropped by the caller: it is the tail of the previous rank's last chunk.  first_cut must be a
 * record boundary (else FQ28_ERR_FORMAT).  Between ranks only this one offset travels.
 * fq28_preparse_dev also starts the field separation of the slab (keys, N counts: it needs the
 * record table, not the chunk boundaries) on a stream of its own, so that it runs while the
 * caller waits for first_cut; a malformed record anywhere in the slab -- the lookahead included --
 * is therefore reported by the following fq28_plan_cut / fq28_compress of this slab. */
int fq28_preparse_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes);
/* host-buffer forms: fq28_preparse copies the slab to the device and builds the record table;
 * fq28_plan_cut / fq28_compress on the same buffer then reuse both */
int fq28_preparse(fq28_handle *h, const char *fastq, size_t n_bytes);
int fq28_plan_cut(fq28_handle *h, const char *fastq, size_t n_bytes, size_t reading_size,
                  int eof, uint64_t first_cut, uint64_t *consumed, size_t *n_chunks);
int fq28_plan_cut_dev(fq28_handle *h, const char *d_fastq, size_t n_bytes, size_t reading_size,
                      int eof, uint64_t first_cut, uint64_t *consumed, size_t *n_chunks);
/* Copies the device-resident result of the last fq28_compress_dev out. */
int fq28_compress_fetch(fq28_handle *h, const fq28_enc_arenas *out);
/* Upper bounds for sizing arenas: Workspace::compressBoundSequence/Quality,
 * src/workspace.h:21-35 (per chunk). */
size_t fq28_bound_seq(size_t tot_reads_length);
size_t fq28_bound_qual(size_t tot_reads_length);

/* -- decompression ----------------------------------------------------------
 * DecompressionWorkspace::decodeChunk src/workspace.cpp:47-88 for a batch of
 * chunks.  Inputs mirror the outputs above; `headers` are the decoded header
 * lines concatenated over all records of all chunks (with '@', without '\n';
 * produced by the host tokeniser) and hdr_lens their u16 lengths.  Output:
 * chunk k is written at out_off[k] = sum of total[0..k) in `fastq_out`. */
typedef struct {
  const uint8_t *seq;       size_t seq_bytes;
  const uint8_t *qual;      size_t qual_bytes;
  const uint16_t *readlens; /* all records of all chunks */
  const uint16_t *n_count;
  const uint16_t *n_pos;    size_t n_pos_entries;
  const uint16_t *hdr_lens;
  const uint8_t *headers;   size_t headers_bytes;
  size_t n_records;
} fq28_dec_arenas;

int fq28_decompress(fq28_handle *h, const fq28_dec_arenas *in,
                    const fq28_chunk_info *infos, size_t n_chunks,
                    char *fastq_out, size_t out_cap, size_t *out_bytes);
/* Device-resident variant: every pointer in `in` and `d_fastq_out` is device
 * memory; infos stay on the host. */
int fq28_decompress_dev(fq28_handle *h, const fq28_dec_arenas *in,
                        const fq28_chunk_info *infos, size_t n_chunks,
                        char *d_fastq_out, size_t out_cap, size_t *out_bytes);

/* -- header tokeniser (row N2, encode side) ----------------------------------
 * encodeHeader for every header of a batch of chunks: src/workspace.cpp:95-125
 * with storeString / storeNumeric of src/headers.cpp:75-89,108-118.  A header is
 * '@' field sep field ... field; the format (field types, separators) and the
 * field values of the archive's first header come from the host
 * (HeaderFormatSpeciciation::fromHeader, src/headers.cpp:43-73).  Every chunk
 * starts from the first header's fields (startNewChunk, src/workspace.cpp:90-93).
 * NUMERIC field: content = int32 delta to the previous record's value, 4 bytes per
 * record.  STRING field: isDifferentFlag = one byte per record; when the value
 * differs from the previous record's, its bytes go to content and its length
 * (one byte, < 255) to contentLength. */
#define FQ28_HDR_MAX_FIELDS 32
typedef struct {
  uint32_t n_fields;
  uint8_t is_string[FQ28_HDR_MAX_FIELDS];      /* headers::FieldType of field i          */
  char separators[FQ28_HDR_MAX_FIELDS];        /* separators[i] follows field i          */
  int32_t first_numeric[FQ28_HDR_MAX_FIELDS];  /* first header's value (NUMERIC fields)  */
  uint32_t first_str_off[FQ28_HDR_MAX_FIELDS + 1]; /* first_strings[off[i], off[i+1]) = first
                                                  header's value of STRING field i        */
  const char *first_strings;                   /* host pointer                           */
} fq28_hdr_format;
typedef struct {                               /* one per (chunk, field), offsets into the arena */
  uint64_t flag_off, flag_len;                 /* isDifferentFlag (STRING only)          */
  uint64_t content_off, content_len;           /* content                                */
  uint64_t clen_off, clen_len;                 /* contentLength (STRING only)            */
} fq28_hdr_field_info;
/* headers: header lines back to back (as in fq28_enc_arenas.headers), hdr_lens
 * per record, chunk_rec[k] = first record of chunk k (n_chunks + 1 entries).
 * infos: n_chunks * n_fields entries, chunk-major.  arena_cap >= headers_bytes +
 * 6 * n_records * n_fields is always enough.  A STRING value of 255 bytes or more
 * -> FQ28_ERR_FORMAT (std::invalid_argument in the reference). */
int fq28_tokenize_headers(fq28_handle *h, const uint8_t *headers, size_t headers_bytes,
                          const uint16_t *hdr_lens, size_t n_records,
                          const uint64_t *chunk_rec, size_t n_chunks,
                          const fq28_hdr_format *fmt, uint8_t *arena, size_t arena_cap,
                          fq28_hdr_field_info *infos, size_t *arena_bytes);

/* decodeHeader for every record of a batch of chunks (row N2, decode side):
 * src/workspace.cpp:127-157 with loadNextString / loadNextNumeric of
 * src/headers.cpp:91-106,120-133.  Input = the field streams as produced by
 * fq28_tokenize_headers (arena + infos); output = header lines back to back
 * ('@' field sep ... field, numeric fields printed with std::to_chars) and their
 * lengths.  Streams that do not fit their chunk's record count -> FQ28_ERR_STREAM;
 * headers_cap too small -> FQ28_ERR_CAP with the needed size in *headers_bytes. */
int fq28_detokenize_headers(fq28_handle *h, const uint8_t *arena, size_t arena_bytes,
                            const fq28_hdr_field_info *infos, const uint64_t *chunk_rec,
                            size_t n_chunks, const fq28_hdr_format *fmt,
                            uint8_t *headers_out, size_t headers_cap, uint16_t *hdr_lens_out,
                            size_t *headers_bytes);

/* -- introspection for tests (device tables copied out) --------------------- */
/* CTable next-state cells / DTable cells of one context (T = 1<<log entries);
 * kind 0 = seq, 1 = qual. */
int fq28_get_ctable(fq28_handle *h, int kind, unsigned ctx, uint16_t *state_table,
                    int32_t *delta_find_state, uint32_t *delta_nb_bits,
                    unsigned *table_log);
int fq28_get_dtable(fq28_handle *h, int kind, unsigned ctx, uint32_t *cells,
                    unsigned *table_log);
/* Device pointers of the last fq28_compress_dev result (for device-resident
 * decode benchmarks); all owned by the handle. */
int fq28_compress_dev_arenas(fq28_handle *h, fq28_dec_arenas *d_view);
/* Seconds spent inside the kernels of the last compress/decompress call,
 * measured with CUDA events on the launching stream, per stage
 * (stage names via fq28_stage_name; n = number of stages filled). */
int fq28_last_timings(const fq28_handle *h, float *ms, size_t cap, size_t *n);
const char *fq28_stage_name(size_t i);

#ifdef __cplusplus
}
#endif
#endif
