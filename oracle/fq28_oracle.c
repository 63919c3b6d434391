/*
 * fq28_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see fq28_oracle.h).
 *
 * Every function cites the reference file:line it follows
 * (/root/reference/src/...).  The FSE / bitstream primitives follow upstream
 * zstd >= 1.5.0 (un-vendored dependency of the reference: iam28th/zstd @
 * b010526d, cmake/Dependencies.cmake:21-27) per SURVEY.md Appendix A.
 */
#define _GNU_SOURCE
#include "fq28_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef uint8_t U8;
typedef uint16_t U16;
typedef uint32_t U32;
typedef uint64_t U64;

static inline U32 hb32(U32 v) { return 31u - (U32)__builtin_clz(v); }

/* ===================================================================== */
/* FSE primitives                                                        */
/* ===================================================================== */

/* zstd FSE_minTableLog (fse_compress.c) */
static U32 fse_min_table_log(size_t src_size, U32 max_sv) {
  U32 min_bits_src = hb32((U32)src_size) + 1;
  U32 min_bits_sym = hb32(max_sv) + 2;
  return min_bits_src < min_bits_sym ? min_bits_src : min_bits_sym;
}

/* zstd FSE_optimalTableLog == FSE_optimalTableLog_internal(..., minus=2);
 * call site src/fse_common.hpp:191.  Appendix A.1. */
unsigned fq28o_optimal_table_log(unsigned max_table_log, size_t src_size,
                                 unsigned max_sv) {
  U32 max_bits_src = hb32((U32)(src_size - 1)) - 2; /* U32 wrap allowed */
  U32 table_log = max_table_log;
  U32 min_bits = fse_min_table_log(src_size, max_sv);
  if (table_log == 0) table_log = FQ28O_FSE_DEFAULT_TABLELOG;
  if (max_bits_src < table_log) table_log = max_bits_src;
  if (min_bits > table_log) table_log = min_bits;
  if (table_log < FQ28O_FSE_MIN_TABLELOG) table_log = FQ28O_FSE_MIN_TABLELOG;
  if (table_log > FQ28O_FSE_MAX_TABLELOG) table_log = FQ28O_FSE_MAX_TABLELOG;
  return table_log;
}

/* zstd FSE_normalizeM2 (fse_compress.c).  Appendix A.2. */
static size_t fse_normalize_m2(int16_t *norm, U32 table_log, const unsigned *count,
                               size_t total, U32 max_sv, int16_t low_prob) {
  const int16_t NYA = -2;
  U32 s, distributed = 0, to_distribute;
  const U32 low_threshold = (U32)(total >> table_log);
  U32 low_one = (U32)((total * 3) >> (table_log + 1));

  for (s = 0; s <= max_sv; s++) {
    if (count[s] == 0) { norm[s] = 0; continue; }
    if (count[s] <= low_threshold) {
      norm[s] = low_prob; distributed++; total -= count[s]; continue;
    }
    if (count[s] <= low_one) {
      norm[s] = 1; distributed++; total -= count[s]; continue;
    }
    norm[s] = NYA;
  }
  to_distribute = (1u << table_log) - distributed;
  if (to_distribute == 0) return 0;

  if ((total / to_distribute) > low_one) {
    low_one = (U32)((total * 3) / (to_distribute * 2));
    for (s = 0; s <= max_sv; s++) {
      if (norm[s] == NYA && count[s] <= low_one) {
        norm[s] = 1; distributed++; total -= count[s];
      }
    }
    to_distribute = (1u << table_log) - distributed;
  }

  if (distributed == max_sv + 1) {
    U32 max_v = 0, max_c = 0;
    for (s = 0; s <= max_sv; s++)
      if (count[s] > max_c) { max_v = s; max_c = count[s]; }
    norm[max_v] = (int16_t)(norm[max_v] + (int16_t)to_distribute);
    return 0;
  }

  if (total == 0) {
    for (s = 0; to_distribute > 0; s = (s + 1) % (max_sv + 1))
      if (norm[s] > 0) { to_distribute--; norm[s]++; }
    return 0;
  }

  {
    const U64 v_step_log = 62 - table_log;
    const U64 mid = (1ULL << (v_step_log - 1)) - 1;
    const U64 r_step = (((1ULL << v_step_log) * to_distribute) + mid) / (U32)total;
    U64 tmp_total = mid;
    for (s = 0; s <= max_sv; s++) {
      if (norm[s] == NYA) {
        const U64 end = tmp_total + (count[s] * r_step);
        const U32 s_start = (U32)(tmp_total >> v_step_log);
        const U32 s_end = (U32)(end >> v_step_log);
        const U32 weight = s_end - s_start;
        if (weight < 1) return (size_t)-1;
        norm[s] = (int16_t)weight;
        tmp_total = end;
      }
    }
  }
  return 0;
}

/* zstd FSE_normalizeCount (fse_compress.c); call site src/fse_common.hpp:192-194
 * with useLowProbCount = 1.  Appendix A.2. */
size_t fq28o_normalize_count(int16_t *norm, unsigned table_log,
                             const unsigned *count, size_t total,
                             unsigned max_sv, unsigned use_low_prob) {
  static const U32 rtb[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
  if (table_log == 0) table_log = FQ28O_FSE_DEFAULT_TABLELOG;
  if (table_log < FQ28O_FSE_MIN_TABLELOG) return (size_t)-1;
  if (table_log > FQ28O_FSE_MAX_TABLELOG) return (size_t)-1;
  if (table_log < fse_min_table_log(total, max_sv)) return (size_t)-1;
  {
    const int16_t low_prob = use_low_prob ? -1 : 1;
    const U64 scale = 62 - table_log;
    const U64 step = (1ULL << 62) / (U32)total;
    const U64 v_step = 1ULL << (scale - 20);
    int still = 1 << table_log;
    unsigned s, largest = 0;
    int16_t largest_p = 0;
    const U32 low_threshold = (U32)(total >> table_log);

    for (s = 0; s <= max_sv; s++) {
      if (count[s] == total) return 0; /* rle */
      if (count[s] == 0) { norm[s] = 0; continue; }
      if (count[s] <= low_threshold) {
        norm[s] = low_prob;
        still--;
      } else {
        int16_t proba = (int16_t)((count[s] * step) >> scale);
        if (proba < 8) {
          U64 rest_to_beat = v_step * rtb[proba];
          proba = (int16_t)(proba + ((count[s] * step) - ((U64)proba << scale) > rest_to_beat));
        }
        if (proba > largest_p) { largest_p = proba; largest = s; }
        norm[s] = proba;
        still -= proba;
      }
    }
    if (-still >= (norm[largest] >> 1)) {
      size_t e = fse_normalize_m2(norm, table_log, count, total, max_sv, low_prob);
      if (e) return e;
    } else {
      norm[largest] = (int16_t)(norm[largest] + (int16_t)still);
    }
  }
  return table_log;
}

/* Symbol spread shared by FSE_buildCTable_wksp / FSE_buildDTable_wksp.
 * Appendix A.3.  (Upstream's fast path for tables without -1 symbols yields
 * the same layout.) */
void fq28o_spread(uint8_t *cell, const int16_t *norm, unsigned max_sv,
                  unsigned table_log) {
  const U32 T = 1u << table_log, mask = T - 1;
  const U32 step = (T >> 1) + (T >> 3) + 3;
  U32 high = T - 1, pos = 0, s;
  for (s = 0; s <= max_sv; s++)
    if (norm[s] == -1) cell[high--] = (U8)s;
  for (s = 0; s <= max_sv; s++) {
    int i;
    for (i = 0; i < norm[s]; i++) {
      cell[pos] = (U8)s;
      pos = (pos + step) & mask;
      while (pos > high) pos = (pos + step) & mask;
    }
  }
}

/* zstd FSE_buildCTable_wksp; call site src/fse_common.hpp:65-68.  A.4. */
void fq28o_build_ctable(uint16_t *state_table, int32_t *dfs, uint32_t *dnb,
                        const int16_t *norm, unsigned max_sv,
                        unsigned table_log) {
  const U32 T = 1u << table_log;
  U8 cell[1u << FQ28O_FSE_MAX_TABLELOG];
  U32 cumul[258];
  U32 u, s, total = 0;
  fq28o_spread(cell, norm, max_sv, table_log);
  cumul[0] = 0;
  for (s = 0; s <= max_sv; s++)
    cumul[s + 1] = cumul[s] + (norm[s] == -1 ? 1u : (U32)norm[s]);
  for (u = 0; u < T; u++) state_table[cumul[cell[u]]++] = (U16)(T + u);
  for (s = 0; s <= max_sv; s++) {
    switch (norm[s]) {
    case 0:
      dnb[s] = ((table_log + 1) << 16) - T;
      dfs[s] = 0;
      break;
    case -1:
    case 1:
      dnb[s] = (table_log << 16) - T;
      dfs[s] = (int32_t)(total - 1);
      total++;
      break;
    default: {
      const U32 max_bits_out = table_log - hb32((U32)norm[s] - 1);
      const U32 min_state_plus = (U32)norm[s] << max_bits_out;
      dnb[s] = (max_bits_out << 16) - min_state_plus;
      dfs[s] = (int32_t)(total - (U32)norm[s]);
      total += (U32)norm[s];
    }
    }
  }
}

/* zstd FSE_buildDTable_wksp; call site src/fse_common.hpp:121-124.  A.6. */
void fq28o_build_dtable(uint32_t *cells, const int16_t *norm, unsigned max_sv,
                        unsigned table_log) {
  const U32 T = 1u << table_log;
  U8 cell[1u << FQ28O_FSE_MAX_TABLELOG];
  U32 next[256];
  U32 u, s;
  fq28o_spread(cell, norm, max_sv, table_log);
  for (s = 0; s <= max_sv; s++) next[s] = (norm[s] == -1) ? 1u : (U32)norm[s];
  for (u = 0; u < T; u++) {
    const U32 sym = cell[u];
    const U32 x = next[sym]++;
    const U32 nb = table_log - hb32(x);
    const U32 ns = (x << nb) - T;
    cells[u] = (ns & 0xFFFFu) | (sym << 16) | (nb << 24);
  }
}

/* ===================================================================== */
/* bit streams (zstd lib/common/bitstream.h semantics, Appendix A.5/A.6)  */
/* ===================================================================== */

typedef struct {
  U64 acc;
  unsigned nbits;
  U8 *ptr, *start, *end;
  int overflow;
} bitw;

static void bw_init(bitw *w, U8 *dst, size_t cap) {
  w->acc = 0; w->nbits = 0; w->ptr = dst; w->start = dst; w->end = dst + cap;
  w->overflow = 0;
}
/* BIT_addBits + BIT_flushBitsFast: LSB-first concatenation; flush cadence
 * does not change the bytes. */
static inline void bw_add(bitw *w, U32 value, unsigned nb) {
  w->acc |= (U64)(value & ((1u << nb) - 1u)) << w->nbits;
  w->nbits += nb;
  if (w->nbits >= 32) {
    if (w->ptr + 4 <= w->end) {
      U32 lo = (U32)w->acc;
      memcpy(w->ptr, &lo, 4);
      w->ptr += 4;
    } else {
      w->overflow = 1;
    }
    w->acc >>= 32;
    w->nbits -= 32;
  }
}
/* BIT_closeCStream: end mark bit, size = ceil(bits/8), 0 when overflowed */
static size_t bw_close(bitw *w) {
  bw_add(w, 1, 1);
  while (w->nbits > 0) {
    if (w->ptr < w->end) *w->ptr++ = (U8)w->acc; else w->overflow = 1;
    w->acc >>= 8;
    w->nbits = w->nbits > 8 ? w->nbits - 8 : 0;
  }
  if (w->overflow) return 0;
  return (size_t)(w->ptr - w->start);
}

typedef struct {
  const U8 *src;
  U64 size;
  U64 pos; /* number of stream bits not yet consumed (below the end mark) */
  int bad;
} bitr;

/* BIT_initDStream: locate the end mark in the last byte */
static int br_init(bitr *r, const U8 *src, size_t size) {
  r->src = src; r->size = size; r->bad = 0; r->pos = 0;
  if (size == 0 || src[size - 1] == 0) { r->bad = 1; return -1; }
  r->pos = (U64)(size - 1) * 8 + hb32(src[size - 1]);
  return 0;
}
/* BIT_readBits: take the next nb bits going downward.  One unaligned 64-bit
 * load per call (what BIT_reloadDStream amounts to), byte gather near the
 * end of the buffer. */
static inline U32 br_read(bitr *r, unsigned nb) {
  U64 w = 0, byte;
  if (nb == 0) return 0;
  if (r->pos < nb) { r->bad = 1; r->pos = 0; return 0; }
  r->pos -= nb;
  byte = r->pos >> 3;
  if (byte + 8 <= r->size) {
    memcpy(&w, r->src + byte, 8);
  } else {
    unsigned i;
    for (i = 0; byte + i < r->size; i++) w |= (U64)r->src[byte + i] << (8 * i);
  }
  return (U32)((w >> (r->pos & 7)) & ((1u << nb) - 1u));
}

/* ===================================================================== */
/* codec tables (FSE_Encoder / FSE_Decoder ctors, src/fse_common.hpp:46-71, */
/* 107-127)                                                               */
/* ===================================================================== */

struct fq28o_codec {
  unsigned n_models, alphabet;
  U32 *logs;       /* [n_models] */
  size_t *toff;    /* [n_models] cell offset of each context's tables */
  U16 *state_tab;  /* CTable next-state cells, packed */
  int32_t *dfs;    /* [n_models*alphabet] */
  U32 *dnb;        /* [n_models*alphabet] */
  U32 *dcells;     /* DTable cells, packed */
};

static fq28o_codec *codec_create(const int16_t *norm, const U32 *logs,
                                 unsigned n_models, unsigned alphabet) {
  fq28o_codec *c = (fq28o_codec *)calloc(1, sizeof(*c));
  size_t total = 0;
  unsigned ctx;
  c->n_models = n_models; c->alphabet = alphabet;
  c->logs = (U32 *)malloc(sizeof(U32) * n_models);
  c->toff = (size_t *)malloc(sizeof(size_t) * n_models);
  for (ctx = 0; ctx < n_models; ctx++) {
    c->logs[ctx] = logs[ctx];
    c->toff[ctx] = total;
    total += (size_t)1 << logs[ctx];
  }
  c->state_tab = (U16 *)malloc(sizeof(U16) * total);
  c->dcells = (U32 *)malloc(sizeof(U32) * total);
  c->dfs = (int32_t *)malloc(sizeof(int32_t) * n_models * alphabet);
  c->dnb = (U32 *)malloc(sizeof(U32) * n_models * alphabet);
  for (ctx = 0; ctx < n_models; ctx++) {
    fq28o_build_ctable(c->state_tab + c->toff[ctx], c->dfs + (size_t)ctx * alphabet,
                       c->dnb + (size_t)ctx * alphabet, norm + (size_t)ctx * alphabet,
                       alphabet - 1, logs[ctx]);
    fq28o_build_dtable(c->dcells + c->toff[ctx], norm + (size_t)ctx * alphabet,
                       alphabet - 1, logs[ctx]);
  }
  return c;
}
fq28o_codec *fq28o_codec_seq(const fq28o_ft_seq *ft) {
  return codec_create(&ft->norm[0][0], ft->logs, FQ28O_SEQ_MODELS, FQ28O_SEQ_ALPHABET);
}
fq28o_codec *fq28o_codec_qual(const fq28o_ft_qual *ft) {
  return codec_create(&ft->norm[0][0], ft->logs, FQ28O_QUAL_MODELS, FQ28O_QUAL_ALPHABET);
}
void fq28o_codec_free(fq28o_codec *c) {
  if (!c) return;
  free(c->logs); free(c->toff); free(c->state_tab); free(c->dfs); free(c->dnb);
  free(c->dcells); free(c);
}

/* FSE_encodeSymbol (zstd fse.h), Appendix A.5 */
static inline void enc_symbol(const fq28o_codec *c, bitw *w, U32 *states,
                              unsigned ctx, unsigned sym) {
  const U32 v = states[ctx];
  const U32 nb = (v + c->dnb[(size_t)ctx * c->alphabet + sym]) >> 16;
  bw_add(w, v, nb);
  states[ctx] = c->state_tab[c->toff[ctx] +
                             (size_t)((int32_t)(v >> nb) +
                                      c->dfs[(size_t)ctx * c->alphabet + sym])];
}
/* FSE_Encoder::startChunk src/fse_common.hpp:77-83: FSE_initCState = 1<<log */
static void enc_start(const fq28o_codec *c, U32 *states) {
  unsigned ctx;
  for (ctx = 0; ctx < c->n_models; ctx++) states[ctx] = 1u << c->logs[ctx];
}
/* FSE_Encoder::endChunk src/fse_common.hpp:86-90: flush ctx 0..N-1, close */
static size_t enc_end(const fq28o_codec *c, bitw *w, const U32 *states) {
  unsigned ctx;
  for (ctx = 0; ctx < c->n_models; ctx++) bw_add(w, states[ctx], c->logs[ctx]);
  return bw_close(w);
}
/* FSE_Decoder::startChunk src/fse_common.hpp:130-139: states N-1..0 */
static void dec_start(const fq28o_codec *c, bitr *r, U32 *states) {
  unsigned i;
  for (i = c->n_models; i > 0; --i) states[i - 1] = br_read(r, c->logs[i - 1]);
}
/* FSE_decodeSymbol (zstd fse.h), Appendix A.6 */
static inline unsigned dec_symbol(const fq28o_codec *c, bitr *r, U32 *states,
                                  unsigned ctx) {
  const U32 e = c->dcells[c->toff[ctx] + states[ctx]];
  states[ctx] = (e & 0xFFFFu) + br_read(r, e >> 24);
  return (e >> 16) & 0xFFu;
}

size_t fq28o_bound_seq(size_t n) { /* src/workspace.h:21-29 */
  if (n < 1024) return (size_t)1024 * FQ28O_SEQ_MODELS;
  return n / 4 + 1024;
}
size_t fq28o_bound_qual(size_t n) { /* src/workspace.h:31-35 */
  size_t a = (size_t)1024 * FQ28O_QUAL_MODELS, b = n * 7 / 8 + 1024;
  return a > b ? a : b;
}

/* ===================================================================== */
/* parsing / chunking                                                    */
/* ===================================================================== */

/* src/fastq_io.cpp:67-125 */
size_t fq28o_parse_records(const char *data, size_t size, fq28o_rec *recs,
                           size_t cap, size_t *n_recs, int *err) {
  size_t processed = 0, n = 0;
  int ln = 0;
  fq28o_rec rec;
  const char *line_start = data;
  memset(&rec, 0, sizeof(rec));
  if (err) *err = 0;
  for (;;) {
    const char *line_end;
    size_t line_len;
    if (processed == size) {
      if (n_recs) *n_recs = n;
      return ln == 0 ? size : (size_t)rec.hdr_off;
    }
    line_end = (const char *)memchr(line_start, '\n', size - processed);
    if (!line_end) {
      if (n_recs) *n_recs = n;
      return ln == 0 ? (size_t)(line_start - data) : (size_t)rec.hdr_off;
    }
    line_len = (size_t)(line_end - line_start);
    if (line_len > 65535) { /* narrow_cast<readlen_t> throws, :95 */
      if (err) *err = FQ28O_ERR_LONG;
      if (n_recs) *n_recs = n;
      return (size_t)(ln == 0 ? (size_t)(line_start - data) : rec.hdr_off);
    }
    switch (ln) {
    case 0:
      rec.hdr_off = (U64)(line_start - data);
      rec.hdr_len = (U32)line_len;
      if (line_len == 0 || line_start[0] != '@') {
        if (err) *err = FQ28O_ERR_FORMAT;
        if (n_recs) *n_recs = n;
        return (size_t)rec.hdr_off;
      }
      break;
    case 1:
      rec.seq_off = (U64)(line_start - data);
      rec.len = (U32)line_len;
      break;
    case 2:
      if (line_len == 0 || line_start[0] != '+') {
        if (err) *err = FQ28O_ERR_FORMAT;
        if (n_recs) *n_recs = n;
        return (size_t)rec.hdr_off;
      }
      break;
    case 3:
      if (line_len != rec.len) {
        if (err) *err = FQ28O_ERR_FORMAT;
        if (n_recs) *n_recs = n;
        return (size_t)rec.hdr_off;
      }
      rec.qual_off = (U64)(line_start - data);
      if (recs) {
        if (n >= cap) {
          if (err) *err = FQ28O_ERR_CAP;
          if (n_recs) *n_recs = n;
          return (size_t)rec.hdr_off;
        }
        recs[n] = rec;
      }
      n++;
      ln = -1;
      break;
    }
    ++ln;
    line_start = line_end + 1;
    processed += line_len + 1;
  }
}

/* src/fastq_io.cpp:23-65 (boundary rule only).  The window of chunk k is
 * [s_k, min(s_k + R, size)); the chunk ends where the first incomplete record
 * of the window starts; reading stops after the window that reaches EOF
 * (a trailing partial record is dropped). */
long fq28o_split_chunks(const char *data, size_t size, size_t R, uint64_t *offs,
                        size_t cap) {
  size_t s = 0, bytes_left = size, partial = 0;
  long n = 0;
  if (cap < 1) return FQ28O_ERR_CAP;
  offs[0] = 0;
  while (bytes_left > 0) {
    size_t to_read = R - partial < bytes_left ? R - partial : bytes_left;
    size_t win = partial + to_read, n_recs = 0, used;
    int err = 0;
    used = fq28o_parse_records(data + s, win, NULL, 0, &n_recs, &err);
    if (err) return err;
    if (used == 0) return FQ28O_ERR_FORMAT; /* record longer than R: reference UB */
    if ((size_t)n + 1 >= cap) return FQ28O_ERR_CAP;
    partial = win - used;
    bytes_left -= to_read;
    s += used;
    offs[++n] = s;
  }
  return n;
}

/* ===================================================================== */
/* frequency tables                                                      */
/* ===================================================================== */

static inline int base2bits(char c) { /* src/fse_sequence.cpp:6-14 */
  switch (c) {
  case 'A': return 0;
  case 'C': return 1;
  case 'G': return 2;
  case 'T': return 3;
  default: return -1;
  }
}
/* src/fse_quality.h:40-44 */
static inline unsigned qual_ctx(unsigned q, unsigned q1, unsigned q2) {
  unsigned ctx = ((((q1 > q2 ? q1 : q2) << 6) + q) & 0xFFFu);
  ctx += (unsigned)(q1 == q2) << 12;
  return ctx;
}

/* src/fse_sequence.cpp:145-169 (without the fill(1) prior, see make_ft) */
int fq28o_hist_seq(const char *data, const fq28o_rec *recs, size_t n, uint32_t *counts) {
  size_t r;
  for (r = 0; r < n; r++) {
    unsigned ctx = FQ28O_SEQ_INITIAL_CTX;
    const char *p = data + recs[r].seq_off;
    U32 i;
    for (i = 0; i < recs[r].len; i++) {
      int sym;
      if (p[i] == 'N') continue; /* :157-158 ctx unchanged */
      sym = base2bits(p[i]);
      if (sym < 0) return FQ28O_ERR_ALPHABET;
      counts[ctx * 4 + (unsigned)sym]++;
      ctx = (ctx >> 2) + ((unsigned)sym << 6); /* addSymUpper, fse_sequence.h:22-24 */
    }
  }
  return 0;
}
/* src/fse_quality.cpp:69-97 */
int fq28o_hist_qual(const char *data, const fq28o_rec *recs, size_t n, uint32_t *counts) {
  size_t r;
  for (r = 0; r < n; r++) {
    unsigned ctx = qual_ctx(0, 0, 0), q1 = 0, q2 = 0;
    const unsigned char *p = (const unsigned char *)data + recs[r].qual_off;
    U32 i;
    for (i = 0; i < recs[r].len; i++) {
      unsigned q = (unsigned)p[i] - FQ28O_QUAL_OFFSET;
      if (q > 63) return FQ28O_ERR_ALPHABET;
      counts[ctx * 64 + q]++;
      ctx = qual_ctx(q, q1, q2);
      q2 = q1; q1 = q;
    }
  }
  return 0;
}

/* makeNormalizedFreqTable src/fse_common.hpp:179-200; the +1 prior is the
 * fill(1) at src/fse_sequence.cpp:149-150 / src/fse_quality.cpp:75-76 */
static void make_ft(const uint32_t *counts, unsigned n_models, unsigned alphabet,
                    int16_t *norm, uint32_t *logs, uint32_t *max_log) {
  unsigned ctx, s;
  unsigned tmp[256];
  *max_log = 0;
  memset(norm, 0, sizeof(int16_t) * (size_t)n_models * alphabet);
  for (ctx = 0; ctx < n_models; ctx++) {
    size_t total = 0;
    for (s = 0; s < alphabet; s++) {
      tmp[s] = counts[(size_t)ctx * alphabet + s] + 1u;
      total += tmp[s];
    }
    logs[ctx] = fq28o_optimal_table_log(0, total, alphabet - 1);
    fq28o_normalize_count(norm + (size_t)ctx * alphabet, logs[ctx], tmp, total,
                          alphabet - 1, 1);
    if (logs[ctx] > *max_log) *max_log = logs[ctx];
  }
}
void fq28o_make_ft_seq(const uint32_t *counts, fq28o_ft_seq *ft) {
  make_ft(counts, FQ28O_SEQ_MODELS, FQ28O_SEQ_ALPHABET, &ft->norm[0][0], ft->logs, &ft->max_log);
}
void fq28o_make_ft_qual(const uint32_t *counts, fq28o_ft_qual *ft) {
  make_ft(counts, FQ28O_QUAL_MODELS, FQ28O_QUAL_ALPHABET, &ft->norm[0][0], ft->logs, &ft->max_log);
}

/* ===================================================================== */
/* sequence / quality codecs                                             */
/* ===================================================================== */

/* replaceAndEncodeNs src/fse_sequence.cpp:35-51 + encodeRecord :53-112.
 * The context of base i is (b[i-1]<<6 | b[i-2]<<4 | b[i-3]<<2 | b[i-4]) with
 * the virtual prefix b[-1..-4] = T,C,C,T (0xD7); symbols are coded for
 * i = L-1 .. 0.  This equals the reference's two-loop formulation for every
 * L >= 1 (checked by tests/test_oracle_golden.py::test_seq_ctx_formulation). */
long fq28o_encode_seq(const fq28o_codec *c, char *data, const fq28o_rec *recs,
                      size_t n, uint8_t *dst, size_t cap, uint16_t *n_count,
                      uint16_t *n_pos, size_t n_pos_cap, size_t *n_npos) {
  U32 states[FQ28O_SEQ_MODELS];
  bitw w;
  size_t r, npos_n = 0;
  bw_init(&w, dst, cap);
  enc_start(c, states);
  for (r = 0; r < n; r++) {
    char *p = data + recs[r].seq_off;
    const U32 L = recs[r].len;
    U32 i;
    U16 cnt = 0, prev = 0;
    if (L < 3) return FQ28O_ERR_SHORT;
    for (i = 0; i < L; i++) {
      if (p[i] == 'N') {
        cnt++;
        if (npos_n >= n_pos_cap) return FQ28O_ERR_CAP;
        n_pos[npos_n++] = (U16)(i - prev);
        p[i] = 'A';
        prev = (U16)i;
      } else if (base2bits(p[i]) < 0) {
        return FQ28O_ERR_ALPHABET;
      }
    }
    n_count[r] = cnt;
    {
      /* ext[j] = 2-bit base of position j-4 over the virtual prefix T,C,C,T;
       * ctx(pos) = ext[pos+3]<<6 | ext[pos+2]<<4 | ext[pos+1]<<2 | ext[pos].
       * Rolling form: going from pos to pos-1 shifts the window down, exactly
       * what addBaseLower does at src/fse_sequence.cpp:84. */
      unsigned ctx;
      U32 pos = L - 1;
#define EXT(j) ((j) >= 4 ? (unsigned)base2bits(p[(j)-4]) : ((FQ28O_SEQ_INITIAL_CTX >> (2 * (j))) & 3u))
      ctx = (EXT(pos + 3) << 6) | (EXT(pos + 2) << 4) | (EXT(pos + 1) << 2) | EXT(pos);
      for (;;) {
        enc_symbol(c, &w, states, ctx, (unsigned)base2bits(p[pos]));
        if (pos == 0) break;
        --pos;
        ctx = ((ctx << 2) & 0xFFu) | EXT(pos);
      }
#undef EXT
    }
  }
  if (n_npos) *n_npos = npos_n;
  return (long)enc_end(c, &w, states);
}

/* src/fse_quality.cpp:5-53: ctx_i = calcContext(q[i-1], q[i-2], q[i-3]),
 * q[<0] = 0, coded for i = L-1 .. 0 (valid for L >= 3) */
long fq28o_encode_qual(const fq28o_codec *c, const char *data,
                       const fq28o_rec *recs, size_t n, uint8_t *dst, size_t cap) {
  U32 *states = (U32 *)malloc(sizeof(U32) * FQ28O_QUAL_MODELS);
  bitw w;
  size_t r;
  long ret;
  bw_init(&w, dst, cap);
  enc_start(c, states);
  for (r = 0; r < n; r++) {
    const unsigned char *p = (const unsigned char *)data + recs[r].qual_off;
    const U32 L = recs[r].len;
    U32 i;
    if (L < 3) { free(states); return FQ28O_ERR_SHORT; }
    for (i = 0; i < L; i++)
      if ((unsigned)p[i] - FQ28O_QUAL_OFFSET > 63) { free(states); return FQ28O_ERR_ALPHABET; }
    for (i = L; i > 0; --i) {
      const U32 pos = i - 1;
      const unsigned q = pos >= 1 ? p[pos - 1] - FQ28O_QUAL_OFFSET : 0;
      const unsigned q1 = pos >= 2 ? p[pos - 2] - FQ28O_QUAL_OFFSET : 0;
      const unsigned q2 = pos >= 3 ? p[pos - 3] - FQ28O_QUAL_OFFSET : 0;
      enc_symbol(c, &w, states, qual_ctx(q, q1, q2), p[pos] - FQ28O_QUAL_OFFSET);
    }
  }
  ret = (long)enc_end(c, &w, states);
  free(states);
  return ret;
}

/* src/fse_sequence.cpp:114-143, records n-1..0 (src/workspace.cpp:84-87) */
int fq28o_decode_seq(const fq28o_codec *c, const uint8_t *src, size_t size,
                     char *data, const fq28o_rec *recs, size_t n,
                     const uint16_t *n_count, const uint16_t *n_pos, size_t n_npos) {
  static const char acgt[4] = {'A', 'C', 'G', 'T'};
  U32 states[FQ28O_SEQ_MODELS];
  bitr br;
  size_t r, npos_idx = n_npos;
  if (br_init(&br, src, size)) return -1;
  dec_start(c, &br, states);
  for (r = n; r > 0; --r) {
    const fq28o_rec *rec = &recs[r - 1];
    char *p = data + rec->seq_off;
    unsigned ctx = FQ28O_SEQ_INITIAL_CTX;
    U32 i, cnt = n_count[r - 1], np = 0;
    for (i = 0; i < rec->len; i++) {
      const unsigned sym = dec_symbol(c, &br, states, ctx);
      p[i] = acgt[sym & 3];
      ctx = (ctx >> 2) + (sym << 6);
    }
    if (cnt > npos_idx) return -2;
    npos_idx -= cnt;
    for (i = 0; i < cnt; i++) {
      np = (np + n_pos[npos_idx + i]) & 0xFFFFu; /* readlen_t arithmetic */
      if (np >= rec->len) return -2;
      p[np] = 'N';
    }
  }
  if (br.bad || br.pos != 0) return -3; /* BIT_endOfDStream */
  return 0;
}

/* src/fse_quality.cpp:55-67 */
int fq28o_decode_qual(const fq28o_codec *c, const uint8_t *src, size_t size,
                      char *data, const fq28o_rec *recs, size_t n) {
  U32 *states = (U32 *)malloc(sizeof(U32) * FQ28O_QUAL_MODELS);
  bitr br;
  size_t r;
  int ret = 0;
  if (br_init(&br, src, size)) { free(states); return -1; }
  dec_start(c, &br, states);
  for (r = n; r > 0; --r) {
    const fq28o_rec *rec = &recs[r - 1];
    char *p = data + rec->qual_off;
    unsigned ctx = qual_ctx(0, 0, 0), q1 = 0, q2 = 0;
    U32 i;
    for (i = 0; i < rec->len; i++) {
      const unsigned q = dec_symbol(c, &br, states, ctx);
      p[i] = (char)(q + FQ28O_QUAL_OFFSET);
      ctx = qual_ctx(q, q1, q2);
      q2 = q1; q1 = q;
    }
  }
  if (br.bad || br.pos != 0) ret = -3;
  free(states);
  return ret;
}

/* src/workspace.cpp:62-80 */
size_t fq28o_layout_chunk(char *out, size_t cap, const char *headers,
                          const uint32_t *hdr_lens, const uint16_t *readlens,
                          size_t n, fq28o_rec *recs) {
  size_t i, pos = 0, hpos = 0;
  for (i = 0; i < n; i++) {
    const size_t need = (size_t)hdr_lens[i] + 2u * readlens[i] + 5;
    if (pos + need > cap) return 0;
    recs[i].hdr_off = pos; recs[i].hdr_len = hdr_lens[i];
    memcpy(out + pos, headers + hpos, hdr_lens[i]);
    pos += hdr_lens[i]; hpos += hdr_lens[i];
    out[pos++] = '\n';
    recs[i].seq_off = pos; recs[i].len = readlens[i];
    pos += readlens[i];
    out[pos++] = '\n'; out[pos++] = '+'; out[pos++] = '\n';
    recs[i].qual_off = pos;
    pos += readlens[i];
    out[pos++] = '\n';
  }
  return pos;
}

/* ===================================================================== */
/* multi-threaded CPU baseline (src/process.cpp:32-105 threading model)   */
/* ===================================================================== */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
  U8 *seq, *qual;
  size_t seq_len, qual_len;
  U16 *readlens, *n_count, *n_pos;
  size_t n_rec, n_npos;
  U32 *hdr_lens;
  char *headers;
  size_t hdr_bytes;
} chunk_out;

typedef struct {
  const char *fastq;
  const uint64_t *offs;
  long n_chunks;
  const fq28o_codec *cs, *cq;
  chunk_out *outs;
  long next;       /* next chunk to pull (under mtx), like readNextChunk */
  pthread_mutex_t mtx;
  int err;
  int decode;
  int mismatch;
} bench_ctx;

static void *bench_worker(void *arg) {
  bench_ctx *b = (bench_ctx *)arg;
  char *buf = NULL;
  size_t buf_cap = 0;
  fq28o_rec *recs = NULL;
  size_t recs_cap = 0;
  for (;;) {
    long k;
    size_t len, n_recs = 0, i, tot = 0;
    int err = 0;
    chunk_out *o;
    pthread_mutex_lock(&b->mtx);
    k = b->next++;
    pthread_mutex_unlock(&b->mtx);
    if (k >= b->n_chunks) break;
    len = (size_t)(b->offs[k + 1] - b->offs[k]);
    o = &b->outs[k];
    if (len > buf_cap) { free(buf); buf = (char *)malloc(len); buf_cap = len; }
    if (!b->decode) {
      /* the reader hands each worker a private copy of the chunk bytes
       * (ifstream::read into chunk.raw_data, src/fastq_io.cpp:47) */
      memcpy(buf, b->fastq + b->offs[k], len);
      fq28o_parse_records(buf, len, NULL, 0, &n_recs, &err);
      if (n_recs > recs_cap) { free(recs); recs = (fq28o_rec *)malloc(sizeof(*recs) * n_recs); recs_cap = n_recs; }
      fq28o_parse_records(buf, len, recs, recs_cap, &n_recs, &err);
      if (err) { b->err = err; break; }
      for (i = 0; i < n_recs; i++) tot += recs[i].len;
      o->n_rec = n_recs;
      o->readlens = (U16 *)malloc(2 * n_recs + 2);
      o->n_count = (U16 *)malloc(2 * n_recs + 2);
      o->n_pos = (U16 *)malloc(2 * tot + 2);
      o->hdr_lens = (U32 *)malloc(4 * n_recs + 4);
      for (i = 0; i < n_recs; i++) { o->readlens[i] = (U16)recs[i].len; o->hdr_lens[i] = recs[i].hdr_len; }
      o->seq = (U8 *)malloc(fq28o_bound_seq(tot));
      o->qual = (U8 *)malloc(fq28o_bound_qual(tot));
      {
        long s = fq28o_encode_seq(b->cs, buf, recs, n_recs, o->seq, fq28o_bound_seq(tot),
                                  o->n_count, o->n_pos, tot + 1, &o->n_npos);
        long q = fq28o_encode_qual(b->cq, buf, recs, n_recs, o->qual, fq28o_bound_qual(tot));
        if (s <= 0 || q <= 0) { b->err = s <= 0 ? (int)(s ? s : FQ28O_ERR_CAP) : (int)(q ? q : FQ28O_ERR_CAP); break; }
        o->seq_len = (size_t)s; o->qual_len = (size_t)q;
        /* shrink like vector::resize keeps capacity: no realloc in the timed path */
      }
      /* headers are out-of-path work (host tokeniser); keep the raw bytes for
       * the decode leg only */
      {
        size_t hb = 0;
        for (i = 0; i < n_recs; i++) hb += recs[i].hdr_len;
        o->headers = (char *)malloc(hb + 1);
        o->hdr_bytes = hb;
        hb = 0;
        for (i = 0; i < n_recs; i++) {
          memcpy(o->headers + hb, b->fastq + b->offs[k] + recs[i].hdr_off, recs[i].hdr_len);
          hb += recs[i].hdr_len;
        }
      }
    } else {
      size_t wrote;
      if (o->n_rec > recs_cap) { free(recs); recs = (fq28o_rec *)malloc(sizeof(*recs) * o->n_rec); recs_cap = o->n_rec; }
      wrote = fq28o_layout_chunk(buf, len, o->headers, o->hdr_lens, o->readlens, o->n_rec, recs);
      if (wrote != len) { b->err = FQ28O_ERR_FORMAT; break; }
      if (fq28o_decode_seq(b->cs, o->seq, o->seq_len, buf, recs, o->n_rec, o->n_count, o->n_pos, o->n_npos) ||
          fq28o_decode_qual(b->cq, o->qual, o->qual_len, buf, recs, o->n_rec)) {
        b->err = -9; break;
      }
      if (memcmp(buf, b->fastq + b->offs[k], len) != 0) b->mismatch = 1;
    }
  }
  free(buf); free(recs);
  return NULL;
}

static void run_workers(bench_ctx *b, int threads) {
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  int i;
  b->next = 0;
  for (i = 0; i < threads; i++) pthread_create(&th[i], NULL, bench_worker, b);
  for (i = 0; i < threads; i++) pthread_join(th[i], NULL);
  free(th);
}

uint64_t fq28o_fnv1a(const uint8_t *p, size_t n, uint64_t h) {
  size_t i;
  for (i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
  return h;
}

int fq28o_bench(const char *fastq, size_t size, size_t sample_bytes,
                size_t reading_size, int threads, int do_decompress,
                fq28o_bench_result *res) {
  bench_ctx b;
  fq28o_ft_seq *fts = (fq28o_ft_seq *)calloc(1, sizeof(*fts));
  fq28o_ft_qual *ftq = (fq28o_ft_qual *)calloc(1, sizeof(*ftq));
  uint32_t *cs = (uint32_t *)calloc(256 * 4, 4), *cq = (uint32_t *)calloc(8192 * 64, 4);
  uint64_t *offs;
  size_t offs_cap = size / (reading_size ? reading_size : 1) * 2 + 16;
  double t0;
  long k;
  memset(res, 0, sizeof(*res));
  memset(&b, 0, sizeof(b));
  if (threads < 1) threads = 1;

  /* analyzeDataset src/prepare.cpp:42-47: first chunk of a reader with R = S */
  t0 = now_s();
  {
    size_t win = sample_bytes < size ? sample_bytes : size, n_recs = 0;
    int err = 0;
    fq28o_rec *recs;
    fq28o_parse_records(fastq, win, NULL, 0, &n_recs, &err);
    recs = (fq28o_rec *)malloc(sizeof(*recs) * (n_recs + 1));
    fq28o_parse_records(fastq, win, recs, n_recs, &n_recs, &err);
    if (err || n_recs == 0) { res->err = err ? err : FQ28O_ERR_FORMAT; free(recs); goto done; }
    if ((err = fq28o_hist_seq(fastq, recs, n_recs, cs)) || (err = fq28o_hist_qual(fastq, recs, n_recs, cq))) {
      res->err = err; free(recs); goto done;
    }
    fq28o_make_ft_seq(cs, fts);
    fq28o_make_ft_qual(cq, ftq);
    free(recs);
  }
  b.cs = fq28o_codec_seq(fts);
  b.cq = fq28o_codec_qual(ftq);
  res->t_analyze_s = now_s() - t0;

  offs = (uint64_t *)malloc(sizeof(uint64_t) * offs_cap);
  k = fq28o_split_chunks(fastq, size, reading_size, offs, offs_cap);
  if (k <= 0) { res->err = k ? (int)k : FQ28O_ERR_FORMAT; free(offs); goto done; }
  b.fastq = fastq; b.offs = offs; b.n_chunks = k;
  b.outs = (chunk_out *)calloc((size_t)k, sizeof(chunk_out));
  pthread_mutex_init(&b.mtx, NULL);

  t0 = now_s();
  b.decode = 0;
  run_workers(&b, threads);
  res->t_compress_s = now_s() - t0;
  res->n_chunks = (uint64_t)k;
  res->fastq_bytes = offs[k];
  if (!b.err) {
    uint64_t h = 1469598103934665603ULL;
    long c;
    size_t i;
    for (c = 0; c < k; c++) {
      res->seq_bytes += b.outs[c].seq_len;
      res->qual_bytes += b.outs[c].qual_len;
      res->n_records += b.outs[c].n_rec;
      for (i = 0; i < b.outs[c].seq_len; i++) { h ^= b.outs[c].seq[i]; h *= 1099511628211ULL; }
      for (i = 0; i < b.outs[c].qual_len; i++) { h ^= b.outs[c].qual[i]; h *= 1099511628211ULL; }
    }
    res->checksum = h;
    if (do_decompress) {
      t0 = now_s();
      b.decode = 1;
      run_workers(&b, threads);
      res->t_decompress_s = now_s() - t0;
      res->roundtrip_ok = !b.mismatch && !b.err;
    }
  }
  res->err = b.err;
  for (k = 0; k < b.n_chunks; k++) {
    chunk_out *o = &b.outs[k];
    free(o->seq); free(o->qual); free(o->readlens); free(o->n_count); free(o->n_pos);
    free(o->hdr_lens); free(o->headers);
  }
  free(b.outs); free(offs);
  pthread_mutex_destroy(&b.mtx);
done:
  fq28o_codec_free((fq28o_codec *)b.cs);
  fq28o_codec_free((fq28o_codec *)b.cq);
  free(fts); free(ftq); free(cs); free(cq);
  return res->err;
}
