/*
 * tps_fastx.c -- streaming FASTQ / FASTA (.gz) reader that fills scan batches.
 *
 * Replaces, for the hot path, the reference's input side: check_file_type / unzip_file
 * (Topsicle/allsteps.py:36-50, 127-149, Bio.SeqIO.parse records) and the re-parse that
 * writes the `_trc_over_` subset (Topsicle/main.py:68-86).  The reference hands every read
 * to Python as a SeqRecord; here records are only *indexed* (title / sequence / quality
 * byte ranges inside the raw text) and their bases are gathered back to back into the
 * caller's (pinned) batch buffer, ready for tps_submit.  Titles and qualities are touched
 * again only for the few reads that pass the TRC cutoff.
 *
 * Record semantics follow Biopython as the reference sees it:
 *   id          = title.split(None, 1)[0]  (first whitespace-delimited token, "" if none)
 *   description = title line after '@' / '>' with trailing whitespace stripped
 *   FASTQ       = 4-line records; sequence / quality lines right-stripped; lengths must agree
 *   FASTA       = '>' title, sequence = lines right-stripped and joined, ' ' and '\r' removed
 * Unlike the reference (which logs a parse error and silently stops, allsteps.py:147-149)
 * malformed input is an error.
 *
 * Plain files are mmap-ed and indexed by several threads (a window is cut into segments,
 * every segment finds its first record start on its own; the pieces must chain exactly or
 * the window is re-indexed sequentially).  `.gz` files are inflated with zlib into chunk buffers
 * that the batch then owns: plain gzip by the calling thread, BGZF (bgzip) by all parser threads.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/topsicle_host.h"
#include "tps_pgz.h"

typedef struct rec_vec {
  tps_fastx_rec *v;
  size_t n, cap;
  uint64_t first, end; /* first record start / end of the last complete record (abs. in window) */
  int status;          /* 0 ok, <0 error */
  uint64_t err_at;
} rec_vec;

typedef struct tps_fastx {
  int format;
  int is_gz;
  int fd;
  gzFile gz;
  const uint8_t *map;
  uint64_t map_len, pos;
  /* gz streaming */
  uint8_t *carry;
  uint64_t carry_len;
  int gz_eof;
  /* plain gzip: parallel two-pass inflate over the mapped file (tps_pgz.c); zlib (gz) only with TPS_FX_NO_PGZ=1 */
  tps_pgz *pgz;
  /* BGZF (bgzip) input: independent <= 64 KiB deflate blocks, inflated by all parser threads */
  int is_bgzf;
  const uint8_t *zmap;
  uint64_t zlen, zpos;
  uint64_t window_bytes;
  uint64_t clip_bases; /* tps_fastx_set_clip: ends kept of a record longer than a whole batch (0 = such a record is an error) */
  int threads;
  int slow_only;      /* never use the one-pass FASTQ reader (tests / tuning) */
  uint64_t n_records; /* records delivered so far */
  char err[256];
} tps_fastx;

static __thread char g_open_err[256];

#include <time.h>
static double dbg_now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static int fx_fail(tps_fastx *fx, int code, const char *fmt, ...) {
  char *dst = fx ? fx->err : g_open_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 256, fmt, ap);
  va_end(ap);
  return code;
}

const char *tps_fastx_last_error(const tps_fastx *fx) { return fx ? fx->err : g_open_err; }

static inline int is_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

/* length of [p, p+n) without trailing whitespace */
static inline uint64_t rstrip_len(const uint8_t *p, uint64_t n) {
  while (n && is_ws(p[n - 1])) --n;
  return n;
}

static int vec_push(rec_vec *rv, const tps_fastx_rec *r) {
  if (rv->n == rv->cap) {
    size_t nc = rv->cap ? rv->cap * 2 : 1024;
    tps_fastx_rec *nv = (tps_fastx_rec *)realloc(rv->v, nc * sizeof(*nv));
    if (!nv) return -1;
    rv->v = nv;
    rv->cap = nc;
  }
  rv->v[rv->n++] = *r;
  return 0;
}

static void set_title(tps_fastx_rec *r, const uint8_t *w, uint64_t t0, uint64_t tl) {
  r->title_off = t0;
  r->title_len = (uint32_t)rstrip_len(w + t0, tl);
  uint32_t i = 0;
  while (i < r->title_len && is_ws(w[t0 + i])) ++i;
  uint32_t j = i;
  while (j < r->title_len && !is_ws(w[t0 + j])) ++j;
  r->id_off = i;
  r->id_len = j - i;
}

/* end of the line starting at `at`: index of '\n' or `end` when none (sets *has_nl) */
static inline uint64_t line_end(const uint8_t *w, uint64_t at, uint64_t end, int *has_nl) {
  const uint8_t *q = (const uint8_t *)memchr(w + at, '\n', end - at);
  *has_nl = q != NULL;
  return q ? (uint64_t)(q - w) : end;
}

/* ---- FASTQ: parse records starting in [from, seg_end); a record may run up to win_end.
 * `final` = the window ends at end of file (a last line without '\n' is complete). */
static void index_fastq(const uint8_t *w, uint64_t from, uint64_t seg_end, uint64_t win_end, int final,
                        rec_vec *rv) {
  uint64_t p = from;
  rv->first = from;
  rv->end = from;
  while (p < seg_end) {
    if (w[p] == '\n' || w[p] == '\r') { /* blank line between records */
      ++p;
      rv->end = p;
      continue;
    }
    if (w[p] != '@') {
      rv->status = TPS_FX_EFORMAT;
      rv->err_at = p;
      return;
    }
    int nl;
    uint64_t e1 = line_end(w, p, win_end, &nl);
    if (!nl) break; /* title incomplete */
    uint64_t s0 = e1 + 1;
    if (s0 >= win_end) break;
    uint64_t e2 = line_end(w, s0, win_end, &nl);
    if (!nl) break;
    uint64_t p0 = e2 + 1;
    if (p0 >= win_end) break;
    if (w[p0] != '+') { /* multi-line FASTQ is undefined in the reference's use; rejected */
      rv->status = TPS_FX_EFORMAT;
      rv->err_at = p0;
      return;
    }
    uint64_t e3 = line_end(w, p0, win_end, &nl);
    if (!nl) break;
    uint64_t q0 = e3 + 1;
    uint64_t e4;
    /* The quality line is as long as the sequence line: jump over it instead of scanning it (half of
     * a FASTQ file is quality text).  Accepted only if a newline sits exactly there and the next line
     * starts a record (or the window ends); anything else takes the scanning path, which validates. */
    const uint64_t j = q0 + (e2 - s0);
    if (j < win_end && w[j] == '\n' && (j + 1 == win_end || w[j + 1] == '@' || w[j + 1] == '\n')) {
      e4 = j;
      nl = 1;
    } else {
      e4 = q0 <= win_end ? line_end(w, q0, win_end, &nl) : win_end;
      if (!nl && !final) break; /* quality line may continue in the next window */
    }
    tps_fastx_rec r;
    memset(&r, 0, sizeof(r));
    set_title(&r, w, p + 1, e1 - (p + 1));
    r.seq_off = s0;
    r.seq_len = (uint32_t)rstrip_len(w + s0, e2 - s0);
    r.seq_raw_len = (uint32_t)(e2 - s0);
    r.qual_off = q0;
    if (rstrip_len(w + q0, e4 - q0) != r.seq_len) {
      rv->status = TPS_FX_EFORMAT;
      rv->err_at = q0;
      return;
    }
    if (vec_push(rv, &r)) {
      rv->status = TPS_FX_ENOMEM;
      return;
    }
    p = nl ? e4 + 1 : e4;
    rv->end = p;
  }
}

/* first FASTQ record start at or after `a` (a > window start): a line that begins with '@'
 * and whose second-next line begins with '+' (a quality line may begin with '@', but then
 * the second-next line is a sequence line, which never begins with '+'). */
static uint64_t find_fastq_start(const uint8_t *w, uint64_t a, uint64_t win_end) {
  int nl;
  uint64_t p = line_end(w, a - 1, win_end, &nl); /* a-1: `a` itself may be a line start */
  if (!nl) return win_end;
  ++p;
  while (p < win_end) {
    if (w[p] == '@') {
      uint64_t e1 = line_end(w, p, win_end, &nl);
      if (!nl) return win_end;
      uint64_t e2 = e1 + 1 < win_end ? line_end(w, e1 + 1, win_end, &nl) : win_end;
      if (e2 >= win_end || !nl) return win_end;
      if (e2 + 1 < win_end && w[e2 + 1] == '+') return p;
      if (e2 + 1 >= win_end) return win_end;
    }
    uint64_t e = line_end(w, p, win_end, &nl);
    if (!nl) return win_end;
    p = e + 1;
  }
  return win_end;
}

/* ---- FASTA */
static void index_fasta(const uint8_t *w, uint64_t from, uint64_t seg_end, uint64_t win_end, int final,
                        rec_vec *rv) {
  uint64_t p = from;
  rv->first = from;
  rv->end = from;
  while (p < seg_end) {
    int nl;
    if (w[p] != '>') { /* text before the first record / blank lines: skipped */
      uint64_t e = line_end(w, p, win_end, &nl);
      if (!nl && !final) return;
      p = nl ? e + 1 : e;
      rv->end = p;
      continue;
    }
    uint64_t e1 = line_end(w, p, win_end, &nl);
    if (!nl && !final) return;
    tps_fastx_rec r;
    memset(&r, 0, sizeof(r));
    set_title(&r, w, p + 1, e1 - (p + 1));
    uint64_t q = nl ? e1 + 1 : e1;
    r.seq_off = q;
    uint64_t bases = 0;
    uint32_t lines = 0, special = 0;
    int complete = 0;
    while (1) {
      if (q >= win_end) {
        complete = final;
        break;
      }
      if (w[q] == '>') {
        complete = 1;
        break;
      }
      uint64_t e = line_end(w, q, win_end, &nl);
      if (!nl && !final) break;
      uint64_t n = rstrip_len(w + q, e - q);
      uint64_t blanks = 0;
      for (uint64_t j = 0; j < n; ++j) blanks += (w[q + j] == ' ') | (w[q + j] == '\r');
      bases += n - blanks;
      special |= blanks != 0;
      if (n) ++lines;
      q = nl ? e + 1 : e;
    }
    if (!complete) return;
    if (bases > 0xFFFFFFFFull || q - r.seq_off > 0xFFFFFFFFull) {
      rv->status = TPS_FX_ECAPACITY;
      rv->err_at = p;
      return;
    }
    r.seq_len = (uint32_t)bases;
    r.seq_raw_len = (uint32_t)(q - r.seq_off);
    r.flags = (lines > 1 || special || r.seq_raw_len != r.seq_len + 1u) ? 1u : 0u;
    if (vec_push(rv, &r)) {
      rv->status = TPS_FX_ENOMEM;
      return;
    }
    p = q;
    rv->end = p;
  }
}

static uint64_t find_fasta_start(const uint8_t *w, uint64_t a, uint64_t win_end) {
  int nl;
  uint64_t p = line_end(w, a - 1, win_end, &nl);
  if (!nl) return win_end;
  ++p;
  while (p < win_end) {
    if (w[p] == '>') return p;
    uint64_t e = line_end(w, p, win_end, &nl);
    if (!nl) return win_end;
    p = e + 1;
  }
  return win_end;
}

/* Copy with non-temporal stores: the destination (a pinned batch buffer, GBs per batch) is only read
 * back by the DMA engine, so it should not displace the file text from the caches nor cost a
 * read-for-ownership of every line. */
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void copy_stream_avx2(uint8_t *dst, const uint8_t *src, size_t n) {
  size_t head = (32u - ((uintptr_t)dst & 31u)) & 31u;
  if (head > n) head = n;
  memcpy(dst, src, head);
  dst += head;
  src += head;
  n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
    __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64));
    __m256i d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a);
    _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c);
    _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  for (; i + 32 <= n; i += 32) _mm256_stream_si256((__m256i *)(dst + i), _mm256_loadu_si256((const __m256i *)(src + i)));
  memcpy(dst + i, src + i, n - i);
}
static int g_have_avx2 = -1;
static inline void copy_stream(uint8_t *dst, const uint8_t *src, size_t n) {
  if (g_have_avx2 < 0) g_have_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
  if (g_have_avx2 && n >= 256) copy_stream_avx2(dst, src, n);
  else memcpy(dst, src, n);
}
static inline void copy_fence(void) { _mm_sfence(); }
#else
static inline void copy_stream(uint8_t *dst, const uint8_t *src, size_t n) { memcpy(dst, src, n); }
static inline void copy_fence(void) {}
#endif

/* copy the bases of record r to dst (r->seq_len bytes) */
static void gather_seq(const uint8_t *w, const tps_fastx_rec *r, uint8_t *dst) {
  if (!(r->flags & 1u)) {
    copy_stream(dst, w + r->seq_off, r->seq_len);
    return;
  }
  uint64_t q = r->seq_off, end = r->seq_off + r->seq_raw_len, o = 0;
  while (q < end) {
    int nl;
    uint64_t e = line_end(w, q, end, &nl);
    uint64_t n = rstrip_len(w + q, e - q);
    for (uint64_t j = 0; j < n; ++j) {
      uint8_t c = w[q + j];
      if (c != ' ' && c != '\r') dst[o++] = c;
    }
    q = nl ? e + 1 : e;
  }
}

/* Index [0, win_end) of w with `threads` segments.  Returns the merged record list in *out
 * (caller frees out->v); out->end = bytes consumed. */
static int index_window(tps_fastx *fx, const uint8_t *w, uint64_t win_end, int final, rec_vec *out) {
  memset(out, 0, sizeof(*out));
  int T = fx->threads;
  if (T < 1) T = 1;
  if (win_end < (uint64_t)T * (1u << 20)) T = 1; /* small windows: not worth splitting */
  void (*index_fn)(const uint8_t *, uint64_t, uint64_t, uint64_t, int, rec_vec *) =
      fx->format == TPS_FX_FASTQ ? index_fastq : index_fasta;
  if (T > 1) {
    rec_vec *parts = (rec_vec *)calloc((size_t)T, sizeof(rec_vec));
    uint64_t *starts = (uint64_t *)calloc((size_t)T + 1, sizeof(uint64_t));
    if (!parts || !starts) {
      free(parts);
      free(starts);
      return fx_fail(fx, TPS_FX_ENOMEM, "out of memory");
    }
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int i = 0; i < T; ++i) {
      uint64_t a = win_end / (uint64_t)T * (uint64_t)i;
      starts[i] = i == 0 ? 0
                         : (fx->format == TPS_FX_FASTQ ? find_fastq_start(w, a, win_end)
                                                       : find_fasta_start(w, a, win_end));
    }
    starts[T] = win_end;
    for (int i = T - 1; i >= 0; --i) /* segments that found nothing inherit the next start */
      if (starts[i] > starts[i + 1]) starts[i] = starts[i + 1];
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int i = 0; i < T; ++i) {
      if (starts[i] < starts[i + 1]) index_fn(w, starts[i], starts[i + 1], win_end, final, &parts[i]);
      else parts[i].first = parts[i].end = starts[i];
    }
    /* the pieces must chain: piece i ends where piece i+1 begins (or piece i stopped early) */
    int ok = 1;
    size_t total = 0;
    int last = -1;
    for (int i = 0; i < T && ok; ++i) {
      if (parts[i].status) ok = 0;
      else if (starts[i] < starts[i + 1]) {
        if (last >= 0 && parts[last].end != parts[i].first) ok = 0;
        last = i;
        total += parts[i].n;
        if (parts[i].end < starts[i + 1]) { /* incomplete record: nothing after it counts */
          for (int j = i + 1; j < T; ++j) parts[j].n = 0;
          break;
        }
      }
    }
    if (ok) {
      out->v = (tps_fastx_rec *)malloc((total ? total : 1) * sizeof(tps_fastx_rec));
      if (!out->v) ok = 0;
    }
    if (ok) {
      out->end = 0;
      for (int i = 0; i < T; ++i) {
        if (starts[i] >= starts[i + 1]) continue;
        memcpy(out->v + out->n, parts[i].v, parts[i].n * sizeof(tps_fastx_rec));
        out->n += parts[i].n;
        if (parts[i].end > out->end) out->end = parts[i].end;
        if (parts[i].end < starts[i + 1]) break;
      }
      out->cap = out->n;
    }
    for (int i = 0; i < T; ++i) free(parts[i].v);
    free(parts);
    free(starts);
    if (ok) return TPS_FX_OK;
    free(out->v);
    memset(out, 0, sizeof(*out)); /* fall through to the sequential index (also reports errors) */
  }
  index_fn(w, 0, win_end, win_end, final, out);
  if (out->status == TPS_FX_EFORMAT && out->n > 0) {
    /* the records before the malformed one are delivered as they are (the reference keeps what it parsed
     * before a bad record, allsteps.py:137-149); the next window starts at the bad record and reports it */
    out->status = 0;
    return TPS_FX_OK;
  }
  if (out->status == TPS_FX_EFORMAT) {
    free(out->v);
    out->v = NULL;
    return fx_fail(fx, TPS_FX_EFORMAT,
                   "malformed %s record #%llu near byte %llu of the current window (4-line FASTQ / FASTA expected; "
                   "sequence and quality lengths must agree)",
                   fx->format == TPS_FX_FASTQ ? "FASTQ" : "FASTA",
                   (unsigned long long)(fx->n_records + out->n + 1), (unsigned long long)out->err_at);
  }
  if (out->status) {
    int st = out->status;
    free(out->v);
    out->v = NULL;
    return fx_fail(fx, st, st == TPS_FX_ENOMEM ? "out of memory" : "record longer than 4 Gbases");
  }
  return TPS_FX_OK;
}

/* ------------------------------------------------------------------------------ public API */
/* ------------------------------------------------------------------ gzip input
 * Plain gzip is one sequential deflate stream: zlib inflates it on the reader thread (gzread).  BGZF
 * (bgzip, the block-gzip of htslib) is a series of gzip members of <= 64 KiB whose compressed size is in
 * an extra header field, so the block boundaries are known without inflating and every parser thread
 * inflates its own blocks.  Python's gzip module -- what the reference opens `.gz` files with,
 * allsteps.py:141-143 -- reads both the same way. */
static int bgzf_header(const uint8_t *p, uint64_t avail, uint32_t *bsize, uint32_t *hdr) {
  if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
  const uint32_t xlen = (uint32_t)p[10] | ((uint32_t)p[11] << 8);
  if (12ull + xlen > avail) return 0;
  uint32_t at = 12, end = 12 + xlen;
  while (at + 4 <= end) {
    const uint32_t slen = (uint32_t)p[at + 2] | ((uint32_t)p[at + 3] << 8);
    if (p[at] == 'B' && p[at + 1] == 'C' && slen == 2 && at + 6 <= end) {
      *bsize = ((uint32_t)p[at + 4] | ((uint32_t)p[at + 5] << 8)) + 1u;
      *hdr = end;
      return *bsize >= end + 8u && *bsize <= avail && !(p[3] & ~4); /* no name / comment / hcrc fields */
    }
    at += 4 + slen;
  }
  return 0;
}

typedef struct bgzf_blk {
  uint64_t zoff, out;
  uint32_t zsize, hdr, isize;
} bgzf_blk;

/* Append inflated text to chunk[*win .. cap): as much as fits.  Sets fx->gz_eof at the end of the input.
 * Returns 0, or <0 after fx_fail. */
/* Window buffers of compressed input are hundreds of megabytes, written once by all parser threads and freed with
 * the batch: 2 MiB-aligned and advised as huge pages, so that first touch is one fault per 2 MiB instead of 512
 * (with sixteen threads faulting at once the address-space lock made a cold window cost as much as inflating it). */
/* Window buffers of compressed input are hundreds of megabytes, live for one batch and are handed back by
 * tps_fastx_release; a fresh one costs its page faults again (a 512 MiB window: ~0.1 s of a batch's 0.13 s at
 * 2 Gbases/s), so the last few are kept and reused (process-wide, by all open files). */
#define WPOOL_LIVE 64
#define WPOOL_KEEP 3
static struct {
  pthread_mutex_t mu;
  void *live[WPOOL_LIVE];
  uint64_t live_sz[WPOOL_LIVE];
  void *kept[WPOOL_KEEP];
  uint64_t kept_sz[WPOOL_KEEP];
} wpool = {PTHREAD_MUTEX_INITIALIZER, {0}, {0}, {0}, {0}};

static uint8_t *window_alloc(uint64_t bytes) {
  void *p = NULL;
  const uint64_t sz = (bytes + (2u << 20) - 1) & ~(uint64_t)((2u << 20) - 1);
  pthread_mutex_lock(&wpool.mu);
  for (int i = 0; i < WPOOL_KEEP && !p; ++i)
    if (wpool.kept[i] && wpool.kept_sz[i] >= sz && wpool.kept_sz[i] <= 2 * sz) {
      p = wpool.kept[i];
      wpool.kept[i] = NULL;
      for (int j = 0; j < WPOOL_LIVE; ++j)
        if (!wpool.live[j]) {
          wpool.live[j] = p;
          wpool.live_sz[j] = wpool.kept_sz[i];
          break;
        }
    }
  pthread_mutex_unlock(&wpool.mu);
  if (p) return (uint8_t *)p;
  if (posix_memalign(&p, 2u << 20, sz)) return NULL;
#ifdef MADV_HUGEPAGE
  madvise(p, sz, MADV_HUGEPAGE);
#endif
  pthread_mutex_lock(&wpool.mu);
  for (int j = 0; j < WPOOL_LIVE; ++j)
    if (!wpool.live[j]) {
      wpool.live[j] = p;
      wpool.live_sz[j] = sz;
      break;
    }
  pthread_mutex_unlock(&wpool.mu);
  return (uint8_t *)p;
}

/* Hands a window buffer back: kept for reuse if it is one of ours and there is room, else freed. */
static void window_free(void *p) {
  if (!p) return;
  uint64_t sz = 0;
  void *drop = p;
  pthread_mutex_lock(&wpool.mu);
  for (int j = 0; j < WPOOL_LIVE; ++j)
    if (wpool.live[j] == p) {
      sz = wpool.live_sz[j];
      wpool.live[j] = NULL;
      break;
    }
  if (sz && !getenv("TPS_FX_NO_POOL")) {
    int at = -1;
    for (int i = 0; i < WPOOL_KEEP; ++i)
      if (!wpool.kept[i]) at = i;
    if (at < 0) { /* full: the smallest one goes */
      at = 0;
      for (int i = 1; i < WPOOL_KEEP; ++i)
        if (wpool.kept_sz[i] < wpool.kept_sz[at]) at = i;
      if (wpool.kept_sz[at] < sz) {
        drop = wpool.kept[at];
        wpool.kept[at] = NULL;
      } else {
        at = -1;
      }
    }
    if (at >= 0) {
      wpool.kept[at] = p;
      wpool.kept_sz[at] = sz;
      if (drop == p) drop = NULL;
    }
  }
  pthread_mutex_unlock(&wpool.mu);
  free(drop);
}

static int gz_fill(tps_fastx *fx, uint8_t *chunk, uint64_t *win, uint64_t cap) {
  if (fx->pgz) { /* may stop short of cap: the inflater only runs stretches that keep every thread busy */
    if (*win < cap && !fx->gz_eof) {
      const int64_t n = tps_pgz_read(fx->pgz, chunk + *win, cap - *win);
      if (n < 0) return fx_fail(fx, TPS_FX_EIO, "gzip input: %s", tps_pgz_error(fx->pgz));
      *win += (uint64_t)n;
      if (tps_pgz_eof(fx->pgz)) fx->gz_eof = 1;
    }
    return TPS_FX_OK;
  }
  if (!fx->is_bgzf) {
    while (*win < cap && !fx->gz_eof) {
      uint64_t ask = cap - *win;
      if (ask > (1u << 30)) ask = 1u << 30;
      int n = gzread(fx->gz, chunk + *win, (unsigned)ask);
      if (n < 0) return fx_fail(fx, TPS_FX_EIO, "gzip read error");
      if (n == 0) fx->gz_eof = 1;
      *win += (uint64_t)n;
    }
    return TPS_FX_OK;
  }
  bgzf_blk *blk = NULL;
  size_t nb = 0, capb = 0;
  uint64_t out = *win, z = fx->zpos;
  while (z < fx->zlen) {
    uint32_t bsize = 0, hdr = 0;
    if (!bgzf_header(fx->zmap + z, fx->zlen - z, &bsize, &hdr)) {
      free(blk);
      return fx_fail(fx, TPS_FX_EIO, "corrupt BGZF block header at compressed offset %llu", (unsigned long long)z);
    }
    const uint8_t *t = fx->zmap + z + bsize - 4;
    const uint32_t isize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    if (out + isize > cap) break;
    if (nb == capb) {
      capb = capb ? capb * 2 : 4096;
      bgzf_blk *nv = (bgzf_blk *)realloc(blk, capb * sizeof(*nv));
      if (!nv) {
        free(blk);
        return fx_fail(fx, TPS_FX_ENOMEM, "out of memory");
      }
      blk = nv;
    }
    blk[nb].zoff = z; blk[nb].out = out; blk[nb].zsize = bsize; blk[nb].hdr = hdr; blk[nb].isize = isize;
    ++nb;
    out += isize;
    z += bsize;
  }
  int bad = 0;
  const int bgzf_zlib = getenv("TPS_FX_BGZF_ZLIB") != NULL && atoi(getenv("TPS_FX_BGZF_ZLIB")) != 0;
#pragma omp parallel for num_threads(fx->threads) schedule(dynamic, 8) if (nb > 16)
  for (int64_t i = 0; i < (int64_t)nb; ++i) {
    const bgzf_blk *b = &blk[i];
    if (b->isize == 0) continue; /* the empty end-of-file block */
    const uint8_t *t = fx->zmap + b->zoff + b->zsize - 8;
    const uint32_t want = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    /* the table-driven byte decoder of tps_pgz.c with the folded CRC-32 (about twice zlib's rate on FASTQ text);
     * zlib when it does not apply (TPS_FX_BGZF_ZLIB=1 forces zlib: A/B, tests) */
    uint32_t got = 0;
    int rc = bgzf_zlib ? -3 : tps_pgz_inflate_block(fx->zmap + b->zoff + b->hdr, b->zsize - b->hdr - 8u, chunk + b->out,
                                                    b->isize, &got);
    int ok = rc == 0 && got == want;
    if (rc == -3) {
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      ok = inflateInit2(&zs, -15) == Z_OK;
      if (ok) {
        zs.next_in = (Bytef *)(fx->zmap + b->zoff + b->hdr);
        zs.avail_in = b->zsize - b->hdr - 8u;
        zs.next_out = chunk + b->out;
        zs.avail_out = b->isize;
        ok = inflate(&zs, Z_FINISH) == Z_STREAM_END && zs.total_out == b->isize;
        inflateEnd(&zs);
      }
      if (ok) ok = (uint32_t)crc32(crc32(0L, Z_NULL, 0), chunk + b->out, b->isize) == want;
    }
    if (!ok) {
#pragma omp atomic write
      bad = 1;
    }
  }
  free(blk);
  if (bad) return fx_fail(fx, TPS_FX_EIO, "corrupt BGZF block (inflate / CRC) near compressed offset %llu",
                          (unsigned long long)fx->zpos);
  fx->zpos = z;
  *win = out;
  if (z >= fx->zlen) fx->gz_eof = 1;
  return TPS_FX_OK;
}

int tps_fastx_open(tps_fastx **out, const char *path, int threads) {
  if (!out || !path) return fx_fail(NULL, TPS_FX_EINVAL, "null argument");
  *out = NULL;
  tps_fastx *fx = (tps_fastx *)calloc(1, sizeof(*fx));
  if (!fx) return fx_fail(NULL, TPS_FX_ENOMEM, "out of memory");
  fx->fd = -1;
  fx->threads = threads > 0 ? threads : 1;
  fx->window_bytes = 4ull << 30;
  size_t pl = strlen(path);
  fx->is_gz = pl >= 3 && strcmp(path + pl - 3, ".gz") == 0; /* by suffix, allsteps.py:37,141 */
  uint8_t first = 0;
  if (fx->is_gz) {
    /* BGZF?  Map the compressed file and look at the first block header */
    int zfd = open(path, O_RDONLY);
    struct stat zst;
    if (zfd >= 0 && fstat(zfd, &zst) == 0 && zst.st_size >= 28) {
      void *zm = mmap(NULL, (size_t)zst.st_size, PROT_READ, MAP_PRIVATE, zfd, 0);
      uint32_t bs = 0, hd = 0;
      if (zm != MAP_FAILED && bgzf_header((const uint8_t *)zm, (uint64_t)zst.st_size, &bs, &hd)) {
        fx->is_bgzf = 1;
        fx->zmap = (const uint8_t *)zm;
        fx->zlen = (uint64_t)zst.st_size;
        fx->fd = zfd;
        zfd = -1;
      } else if (zm != MAP_FAILED && !getenv("TPS_FX_NO_PGZ") &&
                 (fx->pgz = tps_pgz_open((const uint8_t *)zm, (uint64_t)zst.st_size, fx->threads)) != NULL) {
        fx->zmap = (const uint8_t *)zm; /* plain gzip: inflated by all parser threads (tps_pgz.c) */
        fx->zlen = (uint64_t)zst.st_size;
        fx->fd = zfd;
        zfd = -1;
        madvise(zm, (size_t)zst.st_size, MADV_SEQUENTIAL);
      } else if (zm != MAP_FAILED) {
        munmap(zm, (size_t)zst.st_size);
      }
    }
    if (zfd >= 0) close(zfd);
  }
  if (fx->is_bgzf || fx->pgz) {
    fx->carry = (uint8_t *)malloc(1u << 17);
    uint64_t got = 0;
    if (!fx->carry || gz_fill(fx, fx->carry, &got, 1u << 17)) {
      int rc = fx->carry ? TPS_FX_EIO : TPS_FX_ENOMEM;
      snprintf(g_open_err, sizeof(g_open_err), "%s", fx->carry ? fx->err : "out of memory");
      tps_pgz_close(fx->pgz);
      munmap((void *)fx->zmap, fx->zlen);
      close(fx->fd);
      free(fx->carry);
      free(fx);
      return rc;
    }
    fx->carry_len = got;
    uint64_t i = 0; /* check_file_type: first line, stripped (allsteps.py:40) */
    while (i < fx->carry_len && is_ws(fx->carry[i]) && fx->carry[i] != '\n') ++i;
    first = i < fx->carry_len ? fx->carry[i] : 0;
  } else if (fx->is_gz) {
    fx->gz = gzopen(path, "rb");
    if (!fx->gz) {
      int rc = fx_fail(NULL, TPS_FX_EIO, "cannot open %s: %s", path, strerror(errno));
      free(fx);
      return rc;
    }
    gzbuffer(fx->gz, 1u << 20);
    fx->carry = (uint8_t *)malloc(1u << 16);
    int n = fx->carry ? gzread(fx->gz, fx->carry, 1u << 16) : -1;
    if (n < 0) {
      int rc = fx_fail(NULL, TPS_FX_EIO, "cannot read %s (not a gzip stream?)", path);
      gzclose(fx->gz);
      free(fx->carry);
      free(fx);
      return rc;
    }
    fx->carry_len = (uint64_t)n;
    if (n < (1 << 16)) fx->gz_eof = 1;
    uint64_t i = 0; /* check_file_type: first line, stripped (allsteps.py:40) */
    while (i < fx->carry_len && is_ws(fx->carry[i]) && fx->carry[i] != '\n') ++i;
    first = i < fx->carry_len ? fx->carry[i] : 0;
  } else {
    fx->fd = open(path, O_RDONLY);
    struct stat st;
    if (fx->fd < 0 || fstat(fx->fd, &st) != 0) {
      int rc = fx_fail(NULL, TPS_FX_EIO, "cannot open %s: %s", path, strerror(errno));
      if (fx->fd >= 0) close(fx->fd);
      free(fx);
      return rc;
    }
    fx->map_len = (uint64_t)st.st_size;
    if (fx->map_len) {
      void *m = mmap(NULL, fx->map_len, PROT_READ, MAP_PRIVATE, fx->fd, 0);
      if (m == MAP_FAILED) {
        int rc = fx_fail(NULL, TPS_FX_EIO, "cannot mmap %s: %s", path, strerror(errno));
        close(fx->fd);
        free(fx);
        return rc;
      }
      fx->map = (const uint8_t *)m;
      madvise(m, fx->map_len, MADV_SEQUENTIAL);
      uint64_t i = 0;
      while (i < fx->map_len && is_ws(fx->map[i]) && fx->map[i] != '\n') ++i;
      first = i < fx->map_len ? fx->map[i] : 0;
    }
  }
  if (first == '@') fx->format = TPS_FX_FASTQ;
  else if (first == '>') fx->format = TPS_FX_FASTA;
  else {
    int rc = fx_fail(NULL, TPS_FX_EFORMAT, "%s: format cannot be identified (first line starts with neither '@' nor '>')",
                     path);
    if (fx->gz) gzclose(fx->gz);
    tps_pgz_close(fx->pgz);
    if (fx->map) munmap((void *)fx->map, fx->map_len);
    if (fx->zmap) munmap((void *)fx->zmap, fx->zlen);
    if (fx->fd >= 0) close(fx->fd);
    free(fx->carry);
    free(fx);
    return rc;
  }
  *out = fx;
  return TPS_FX_OK;
}

/* Dropping the page-table entries of a multi-GB mapping is the expensive part of munmap and runs on
 * one thread under the exclusive mm lock; madvise(MADV_DONTNEED) does the same work under the shared
 * lock, so it can be spread over the parser threads first (the page cache itself is untouched). */
static void unmap_parallel(const uint8_t *map, uint64_t len, int threads) {
  const uint64_t chunk = 32ull << 20;
  const int64_t n = (int64_t)((len + chunk - 1) / chunk);
  if (threads > 1 && n > 1) {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
      uint64_t a = (uint64_t)i * chunk, b = a + chunk < len ? a + chunk : len;
      madvise((void *)(map + a), b - a, MADV_DONTNEED);
    }
  }
  munmap((void *)map, len);
}

void tps_fastx_close(tps_fastx *fx) {
  if (!fx) return;
  if (fx->gz) gzclose(fx->gz);
  tps_pgz_close(fx->pgz);
  if (fx->zmap) munmap((void *)fx->zmap, fx->zlen);
  if (fx->map) unmap_parallel(fx->map, fx->map_len, fx->threads);
  if (fx->fd >= 0) close(fx->fd);
  free(fx->carry);
  free(fx);
}

int tps_fastx_format(const tps_fastx *fx) { return fx ? fx->format : 0; }
void tps_fastx_inflate_stats(const tps_fastx *fx, tps_pgz_stats *out) {
  if (!out) return;
  memset(out, 0, sizeof(*out));
  if (fx && fx->pgz) tps_pgz_get_stats(fx->pgz, out);
}
void tps_fastx_set_window(tps_fastx *fx, uint64_t bytes) {
  if (fx && bytes >= 4096) fx->window_bytes = bytes;
}
void tps_fastx_set_clip(tps_fastx *fx, uint64_t clip_bases) {
  if (fx) fx->clip_bases = clip_bases;
}
void tps_fastx_release(void *owner) { window_free(owner); }
void tps_fastx_set_two_pass(tps_fastx *fx, int on) {
  if (fx) fx->slow_only = on;
}

/* Acquire the next window of raw text (a view of the mapping, or a freshly inflated chunk) and index its
 * records.  Returns 0 with *rv filled, 1 at end of file, <0 on error.  `want` = raw bytes to look at. */
static int acquire_indexed(tps_fastx *fx, uint64_t want, const uint8_t **w_out, uint8_t **chunk_out, uint64_t *win_out,
                           int *final_out, rec_vec *rv_out) {
  const uint8_t *w = NULL;
  uint8_t *chunk = NULL;
  uint64_t win = 0;
  int final = 0;
  rec_vec rv;
  memset(&rv, 0, sizeof(rv));
  for (;;) {
    if (!fx->is_gz) {
      if (fx->pos >= fx->map_len) return 1;
      w = fx->map + fx->pos;
      win = fx->map_len - fx->pos;
      if (win > want) win = want;
      final = fx->pos + win == fx->map_len;
    } else {
      if (fx->carry_len == 0 && fx->gz_eof) return 1;
      uint64_t cap = fx->carry_len > want ? fx->carry_len : want;
      if (!chunk) {
        chunk = window_alloc(cap + 1);
        if (!chunk) return fx_fail(fx, TPS_FX_ENOMEM, "out of memory for a %llu-byte chunk", (unsigned long long)cap);
        memcpy(chunk, fx->carry, fx->carry_len);
        win = fx->carry_len;
      } else {
        uint8_t *nc = window_alloc(cap + 1);
        if (!nc) {
          window_free(chunk);
          return fx_fail(fx, TPS_FX_ENOMEM, "out of memory for a %llu-byte chunk", (unsigned long long)cap);
        }
        memcpy(nc, chunk, win);
        window_free(chunk);
        chunk = nc;
      }
      {
        int frc = gz_fill(fx, chunk, &win, cap);
        if (frc) {
          window_free(chunk);
          return frc;
        }
      }
      w = chunk;
      final = fx->gz_eof;
    }
    double t_ix = dbg_now();
    int rc = index_window(fx, w, win, final, &rv);
    if (getenv("TPS_FX_DEBUG")) fprintf(stderr, "[fastx] index %.1f MB in %.4f s, %zu records\n", win / 1e6, dbg_now() - t_ix, rv.n);
    if (rc) {
      window_free(chunk);
      return rc;
    }
    if (rv.n > 0 || final) break;
    /* not one complete record in the window: widen it */
    free(rv.v);
    memset(&rv, 0, sizeof(rv));
    if (want >= (1ull << 40)) {
      window_free(chunk);
      return fx_fail(fx, TPS_FX_ECAPACITY, "record larger than 1 TiB");
    }
    want *= 2;
  }
  *w_out = w;
  *chunk_out = chunk;
  *win_out = win;
  *final_out = final;
  *rv_out = rv;
  return 0;
}

static void gather_ends(const uint8_t *w, const tps_fastx_rec *r, uint32_t end_len, uint8_t *dst);

/* Next batch: at most reads_cap records and bases_cap bases, in file order.
 *   bases_out[offsets_out[i] .. offsets_out[i+1]) = bases of record i; recs_out[i] indexes its text
 *   relative to *raw_base, which stays valid until tps_fastx_release(*raw_owner) (gz input) or
 *   tps_fastx_close (plain input, *raw_owner == NULL).
 * *n_reads == 0 means end of file. */
int tps_fastx_next(tps_fastx *fx, uint64_t bases_cap, uint32_t reads_cap, uint8_t *bases_out,
                   uint64_t *offsets_out, tps_fastx_rec *recs_out, uint32_t *n_reads,
                   const uint8_t **raw_base, void **raw_owner) {
  if (!fx || !bases_out || !offsets_out || !recs_out || !n_reads || !raw_base || !raw_owner)
    return fx_fail(fx, TPS_FX_EINVAL, "null argument");
  *n_reads = 0;
  *raw_base = NULL;
  *raw_owner = NULL;
  offsets_out[0] = 0;
  if (reads_cap == 0 || bases_cap == 0) return fx_fail(fx, TPS_FX_EINVAL, "zero capacity");
  /* raw bytes that can hold a full batch: FASTQ carries the quality line too */
  uint64_t want = (fx->format == TPS_FX_FASTQ ? 2 : 1) * bases_cap + bases_cap / 32 + (uint64_t)reads_cap * 128 + 4096;
  if (want > fx->window_bytes) want = fx->window_bytes;
  const uint8_t *w = NULL;
  uint8_t *chunk = NULL;
  uint64_t win = 0;
  int final = 0;
  rec_vec rv;
  {
    int rc = acquire_indexed(fx, want, &w, &chunk, &win, &final, &rv);
    if (rc == 1) return TPS_FX_OK; /* end of file */
    if (rc) return rc;
  }
  /* longest prefix within the caps */
  uint32_t n = 0;
  uint64_t nb = 0;
  while (n < rv.n && n < reads_cap && nb + rv.v[n].seq_len <= bases_cap) {
    nb += rv.v[n].seq_len;
    offsets_out[n + 1] = nb;
    ++n;
  }
  /* A record longer than a whole batch (a chromosome in a FASTA of contigs): the scan only looks at a bounded
   * stretch of either end (allsteps.py:176-177 the first / last 1000 bases, :266-268 at most maxlengthtelo bases
   * from the chosen end), so with tps_fastx_set_clip the record is delivered, alone, as its first and last
   * clip_bases bases back to back; recs_out[0] still describes the whole record. */
  int clipped = 0;
  if (n == 0 && rv.n > 0 && fx->clip_bases && 2 * fx->clip_bases <= bases_cap && 2 * fx->clip_bases <= 0xFFFFFFFFull &&
      rv.v[0].seq_len > 2 * fx->clip_bases) {
    clipped = 1;
    n = 1;
    nb = 2 * fx->clip_bases;
    offsets_out[1] = nb;
  }
  if (n == 0 && rv.n > 0) {
    uint32_t L = rv.v[0].seq_len;
    free(rv.v);
    window_free(chunk);
    return fx_fail(fx, TPS_FX_ECAPACITY,
                   "read #%llu has %u bases, more than the batch capacity of %llu%s",
                   (unsigned long long)(fx->n_records + 1), L, (unsigned long long)bases_cap,
                   fx->clip_bases ? " (which cannot even hold its two ends: raise the batch size, TOPSICLE_BATCH_BASES)" : "");
  }
  uint64_t consumed;
  if (n == rv.n) consumed = rv.end;
  else consumed = rv.v[n].title_off - 1; /* start of the first record left for the next call */
  if (n) memcpy(recs_out, rv.v, (size_t)n * sizeof(tps_fastx_rec));
  int T = fx->threads;
  double t_g = dbg_now();
  if (clipped) {
    gather_ends(w, &rv.v[0], (uint32_t)fx->clip_bases, bases_out);
  } else {
#pragma omp parallel num_threads(T) if (nb > (8u << 20))
    {
#pragma omp for schedule(dynamic, 16)
      for (int64_t i = 0; i < (int64_t)n; ++i) gather_seq(w, &rv.v[i], bases_out + offsets_out[i]);
      copy_fence(); /* non-temporal stores are visible before the batch is handed to the DMA engine */
    }
  }
  if (getenv("TPS_FX_DEBUG")) fprintf(stderr, "[fastx] gather %.1f MB in %.4f s\n", nb / 1e6, dbg_now() - t_g);
  free(rv.v);
  if (!fx->is_gz) {
    *raw_base = w;
    fx->pos += consumed;
    if (n == 0) fx->pos = fx->map_len; /* only blank / skipped text was left */
  } else {
    uint64_t left = win - consumed;
    if (n == 0) left = 0;
    uint8_t *nc = (uint8_t *)realloc(fx->carry, left ? left : 1);
    if (!nc) {
      window_free(chunk);
      return fx_fail(fx, TPS_FX_ENOMEM, "out of memory");
    }
    fx->carry = nc;
    memcpy(fx->carry, chunk + consumed, left);
    fx->carry_len = left;
    if (n) {
      *raw_base = chunk;
      *raw_owner = chunk;
    } else {
      window_free(chunk);
    }
  }
  fx->n_records += n;
  *n_reads = n;
  return TPS_FX_OK;
}

/* ------------------------------------------------------------------ ends batches (ends-first scanning)
 * Step 1 of the scan only reads the first and the last `end_len` bases of a read (allsteps.py:176-177), so
 * for inputs that are mostly not telomeric the interior of the reads need not be copied, uploaded or packed
 * at all: the index pass finds the record boundaries (one memchr over the sequence line, the quality line is
 * jumped over), then every record contributes head + tail (the whole read when L <= 2 * end_len), packed back
 * to back.  true_lens_out[i] = L.  The records' text stays addressable through recs_out / *raw_base, which is
 * where the caller takes the regions of the TRC-pass reads from (tps_submit_regions).
 * At most `raw_cap` bytes of file text, `reads_cap` records and `bases_cap` uploaded bases per batch. */
static void gather_ends(const uint8_t *w, const tps_fastx_rec *r, uint32_t end_len, uint8_t *dst) {
  const uint32_t L = r->seq_len;
  if (!(r->flags & 1u)) {
    if (L <= 2u * end_len) {
      memcpy(dst, w + r->seq_off, L);
    } else {
      memcpy(dst, w + r->seq_off, end_len);
      memcpy(dst + end_len, w + r->seq_off + (L - end_len), end_len);
    }
    return;
  }
  /* multi-line / blank-carrying sequence: filter it once, then take the ends */
  uint8_t *tmp = (uint8_t *)malloc(L ? L : 1);
  if (!tmp) { /* cannot happen for sane L; leave the slot as N so that nothing matches */
    memset(dst, 'N', L <= 2u * end_len ? L : 2u * end_len);
    return;
  }
  gather_seq(w, r, tmp);
  if (L <= 2u * end_len) {
    memcpy(dst, tmp, L);
  } else {
    memcpy(dst, tmp, end_len);
    memcpy(dst + end_len, tmp + (L - end_len), end_len);
  }
  free(tmp);
}

int tps_fastx_next_ends(tps_fastx *fx, uint64_t raw_cap, uint64_t bases_cap, uint32_t reads_cap, uint32_t end_len,
                        uint8_t *bases_out, uint64_t *starts_out, uint32_t *lens_out, uint32_t *true_lens_out,
                        tps_fastx_rec *recs_out, uint32_t *n_reads, uint64_t *span_out, uint64_t *true_bases_out,
                        const uint8_t **raw_base, void **raw_owner) {
  if (!fx || !bases_out || !starts_out || !lens_out || !true_lens_out || !recs_out || !n_reads || !span_out ||
      !true_bases_out || !raw_base || !raw_owner)
    return fx_fail(fx, TPS_FX_EINVAL, "null argument");
  *n_reads = 0;
  *span_out = 0;
  *true_bases_out = 0;
  *raw_base = NULL;
  *raw_owner = NULL;
  if (reads_cap == 0 || bases_cap == 0 || end_len == 0) return fx_fail(fx, TPS_FX_EINVAL, "zero capacity");
  uint64_t want = raw_cap < 4096 ? 4096 : raw_cap;
  if (want > fx->window_bytes) want = fx->window_bytes;
  const uint8_t *w = NULL;
  uint8_t *chunk = NULL;
  uint64_t win = 0;
  int final = 0;
  rec_vec rv;
  {
    int rc = acquire_indexed(fx, want, &w, &chunk, &win, &final, &rv);
    if (rc == 1) return TPS_FX_OK; /* end of file */
    if (rc) return rc;
  }
  (void)final;
  uint32_t n = 0;
  uint64_t nb = 0, tb = 0;
  while (n < rv.n && n < reads_cap) {
    const uint32_t L = rv.v[n].seq_len;
    const uint32_t m = L <= 2u * end_len ? L : 2u * end_len;
    if (nb + m > bases_cap) break;
    starts_out[n] = nb;
    lens_out[n] = m;
    true_lens_out[n] = L;
    nb += m;
    tb += L;
    ++n;
  }
  if (n == 0 && rv.n > 0) {
    free(rv.v);
    window_free(chunk);
    return fx_fail(fx, TPS_FX_ECAPACITY, "the batch capacity of %llu bases cannot hold the ends of one read",
                   (unsigned long long)bases_cap);
  }
  uint64_t consumed;
  if (n == rv.n) consumed = rv.end;
  else consumed = rv.v[n].title_off - 1;
  if (n) memcpy(recs_out, rv.v, (size_t)n * sizeof(tps_fastx_rec));
  int T = fx->threads;
#pragma omp parallel for num_threads(T) schedule(dynamic, 256) if (n > 4096)
  for (int64_t i = 0; i < (int64_t)n; ++i) gather_ends(w, &rv.v[i], end_len, bases_out + starts_out[i]);
  free(rv.v);
  if (!fx->is_gz) {
    *raw_base = w;
    fx->pos += consumed;
    if (n == 0) fx->pos = fx->map_len;
  } else {
    uint64_t left = win - consumed;
    if (n == 0) left = 0;
    uint8_t *nc = (uint8_t *)realloc(fx->carry, left ? left : 1);
    if (!nc) {
      window_free(chunk);
      return fx_fail(fx, TPS_FX_ENOMEM, "out of memory");
    }
    fx->carry = nc;
    memcpy(fx->carry, chunk + consumed, left);
    fx->carry_len = left;
    if (n) {
      *raw_base = chunk;
      *raw_owner = chunk;
    } else {
      window_free(chunk);
    }
  }
  fx->n_records += n;
  *n_reads = n;
  *span_out = nb;
  *true_bases_out = tb;
  return TPS_FX_OK;
}

/* ------------------------------------------------------------------ one-pass FASTQ -> span batch
 * In a 4-line FASTQ record the quality line is as long as the sequence line, so a record that starts
 * at byte x of the window owns at least 2*L + 6 bytes of text.  Placing its L bases at x/2 of the batch
 * buffer therefore never collides with the next record -- and needs no knowledge of the records before
 * it.  Every parser thread can then copy the sequence line WHILE it looks for its end (one pass over
 * the text, streaming stores), instead of indexing first and gathering after a global prefix sum.  The
 * batch is a "span batch": reads separated by small gaps (tps_submit_spans). */
#if defined(__x86_64__)
__attribute__((target("avx2"))) static int64_t copy_line_avx2(uint8_t *dst, const uint8_t *src, uint64_t n) {
  const __m256i nl = _mm256_set1_epi8('\n');
  uint64_t i = 0;
  if (n >= 32) {
    const __m256i v = _mm256_loadu_si256((const __m256i *)src);
    const uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, nl));
    if (m) {
      const uint32_t k = (uint32_t)__builtin_ctz(m);
      memcpy(dst, src, k);
      return (int64_t)k;
    }
    _mm256_storeu_si256((__m256i *)dst, v);
    i = 32u - ((uintptr_t)dst & 31u); /* 1..32: from here dst + i is 32-byte aligned */
  }
  for (; i + 32 <= n; i += 32) {
    const __m256i v = _mm256_loadu_si256((const __m256i *)(src + i));
    const uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, nl));
    if (m) {
      const uint32_t k = (uint32_t)__builtin_ctz(m);
      memcpy(dst + i, src + i, k);
      return (int64_t)(i + k);
    }
    if (((uintptr_t)(dst + i) & 31u) == 0) _mm256_stream_si256((__m256i *)(dst + i), v);
    else _mm256_storeu_si256((__m256i *)(dst + i), v);
  }
  const uint8_t *q = (const uint8_t *)memchr(src + i, '\n', n - i);
  if (q) {
    memcpy(dst + i, src + i, (size_t)(q - (src + i)));
    return (int64_t)(q - src);
  }
  memcpy(dst + i, src + i, n - i);
  return -1;
}
#endif

/* copy src[0..] to dst up to (not including) the first '\n' within n bytes; returns the line length or
 * -1 when there is no newline (n bytes were then copied) */
static int64_t copy_line(uint8_t *dst, const uint8_t *src, uint64_t n) {
#if defined(__x86_64__)
  if (g_have_avx2 < 0) g_have_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
  if (g_have_avx2) return copy_line_avx2(dst, src, n);
#endif
  const uint8_t *q = (const uint8_t *)memchr(src, '\n', n);
  memcpy(dst, src, q ? (size_t)(q - src) : n);
  return q ? (int64_t)(q - src) : -1;
}

#define TPS_FX_NEED_SLOW (-100) /* internal: this window must take the two-pass path */

/* records starting in [from, seg_end): parse + copy bases to dst[(record start) >> 1 ...] */
static void scan_fastq_fused(const uint8_t *w, uint64_t from, uint64_t seg_end, uint64_t win_end, int final,
                             uint8_t *dst, rec_vec *rv) {
  uint64_t p = from;
  rv->first = from;
  rv->end = from;
  while (p < seg_end) {
    if (w[p] == '\n' || w[p] == '\r') {
      ++p;
      rv->end = p;
      continue;
    }
    if (w[p] != '@') {
      rv->status = TPS_FX_EFORMAT;
      rv->err_at = p;
      return;
    }
    int nl;
    const uint64_t e1 = line_end(w, p, win_end, &nl);
    if (!nl) break;
    const uint64_t s0 = e1 + 1;
    if (s0 >= win_end) break;
    /* a complete record needs 2 * line bytes: a longer line cannot end inside this window, and
     * copying more than (win_end - p) / 2 bytes would leave the first win/2 bytes of the batch buffer */
    uint64_t lim = (win_end - p) / 2;
    if (lim > win_end - s0) lim = win_end - s0;
    const int64_t ll = copy_line(dst + (p >> 1), w + s0, lim);
    if (ll < 0) break; /* sequence line runs past the window */
    const uint64_t e2 = s0 + (uint64_t)ll;
    const uint64_t p0 = e2 + 1;
    if (p0 >= win_end) break;
    if (w[p0] != '+') {
      rv->status = TPS_FX_EFORMAT;
      rv->err_at = p0;
      return;
    }
    const uint64_t e3 = line_end(w, p0, win_end, &nl);
    if (!nl) break;
    const uint64_t q0 = e3 + 1;
    uint64_t e4;
    const uint64_t j = q0 + (e2 - s0);
    if (j < win_end && w[j] == '\n' && (j + 1 == win_end || w[j + 1] == '@' || w[j + 1] == '\n')) {
      e4 = j;
      nl = 1;
    } else {
      e4 = q0 <= win_end ? line_end(w, q0, win_end, &nl) : win_end;
      if (!nl && !final) break;
    }
    tps_fastx_rec r;
    memset(&r, 0, sizeof(r));
    set_title(&r, w, p + 1, e1 - (p + 1));
    r.seq_off = s0;
    r.seq_len = (uint32_t)rstrip_len(w + s0, e2 - s0);
    r.seq_raw_len = (uint32_t)(e2 - s0);
    r.qual_off = q0;
    if (rstrip_len(w + q0, e4 - q0) != r.seq_len) {
      rv->status = TPS_FX_EFORMAT;
      rv->err_at = q0;
      return;
    }
    /* the x/2 placement needs record bytes >= 2 * (copied bytes): only odd trailing blanks break it */
    if ((e2 - s0) - r.seq_len > 4 || (nl ? e4 + 1 : e4) - p < 2 * (e2 - s0) + 2) {
      rv->status = TPS_FX_NEED_SLOW;
      return;
    }
    if (vec_push(rv, &r)) {
      rv->status = TPS_FX_ENOMEM;
      return;
    }
    p = nl ? e4 + 1 : e4;
    rv->end = p;
  }
}

static int next_spans_contiguous(tps_fastx *fx, uint64_t span_cap, uint32_t reads_cap, uint8_t *bases_out,
                                 uint64_t *starts_out, uint32_t *lens_out, tps_fastx_rec *recs_out, uint32_t *n_reads,
                                 uint64_t *span_used, const uint8_t **raw_base, void **raw_owner);

/* Next batch as a span batch: read i = bases_out[starts_out[i] .. + lens_out[i]); *span_used bytes of
 * bases_out are meaningful.  starts_out must hold reads_cap + 1 entries.  Otherwise like tps_fastx_next. */
int tps_fastx_next_spans(tps_fastx *fx, uint64_t span_cap, uint32_t reads_cap, uint8_t *bases_out,
                         uint64_t *starts_out, uint32_t *lens_out, tps_fastx_rec *recs_out, uint32_t *n_reads,
                         uint64_t *span_used, const uint8_t **raw_base, void **raw_owner) {
  if (!fx || !bases_out || !starts_out || !lens_out || !recs_out || !n_reads || !span_used || !raw_base || !raw_owner)
    return fx_fail(fx, TPS_FX_EINVAL, "null argument");
  *n_reads = 0;
  *span_used = 0;
  *raw_base = NULL;
  *raw_owner = NULL;
  if (reads_cap == 0 || span_cap < 64) return fx_fail(fx, TPS_FX_EINVAL, "zero capacity");
  if (fx->format != TPS_FX_FASTQ || fx->slow_only)
    return next_spans_contiguous(fx, span_cap, reads_cap, bases_out, starts_out, lens_out, recs_out, n_reads, span_used,
                                 raw_base, raw_owner);
  uint64_t want = 2 * span_cap; /* a record at window byte x lands at x/2 */
  if (want > fx->window_bytes) want = fx->window_bytes;
  const uint8_t *w = NULL;
  uint8_t *chunk = NULL;
  uint64_t win = 0;
  int final = 0;
  if (!fx->is_gz) {
    if (fx->pos >= fx->map_len) return TPS_FX_OK;
    w = fx->map + fx->pos;
    win = fx->map_len - fx->pos;
    if (win > want) win = want;
    final = fx->pos + win == fx->map_len;
  } else {
    if (fx->carry_len == 0 && fx->gz_eof) return TPS_FX_OK;
    if (fx->carry_len >= want) /* a carried record as large as the batch: the two-pass path reports it */
      return next_spans_contiguous(fx, span_cap, reads_cap, bases_out, starts_out, lens_out, recs_out, n_reads,
                                   span_used, raw_base, raw_owner);
    chunk = window_alloc(want + 1);
    if (!chunk) return fx_fail(fx, TPS_FX_ENOMEM, "out of memory for a %llu-byte chunk", (unsigned long long)want);
    memcpy(chunk, fx->carry, fx->carry_len);
    win = fx->carry_len;
    {
      int frc = gz_fill(fx, chunk, &win, want);
      if (frc) {
        window_free(chunk);
        return frc;
      }
    }
    w = chunk;
    final = fx->gz_eof;
  }
  int T = fx->threads;
  if (T < 1) T = 1;
  if (win < (uint64_t)T * (1u << 20)) T = 1;
  rec_vec *parts = (rec_vec *)calloc((size_t)T, sizeof(rec_vec));
  uint64_t *sst = (uint64_t *)calloc((size_t)T + 1, sizeof(uint64_t));
  if (!parts || !sst) {
    free(parts);
    free(sst);
    window_free(chunk);
    return fx_fail(fx, TPS_FX_ENOMEM, "out of memory");
  }
  double t_ix = dbg_now();
#pragma omp parallel for num_threads(T) schedule(static, 1)
  for (int i = 0; i < T; ++i) sst[i] = i == 0 ? 0 : find_fastq_start(w, win / (uint64_t)T * (uint64_t)i, win);
  sst[T] = win;
  for (int i = T - 1; i >= 0; --i)
    if (sst[i] > sst[i + 1]) sst[i] = sst[i + 1];
#pragma omp parallel num_threads(T)
  {
#pragma omp for schedule(static, 1)
    for (int i = 0; i < T; ++i) {
      if (sst[i] < sst[i + 1]) scan_fastq_fused(w, sst[i], sst[i + 1], win, final, bases_out, &parts[i]);
      else parts[i].first = parts[i].end = sst[i];
    }
    copy_fence();
  }
  /* the pieces must chain exactly; anything odd (format error, pathological blanks, a segment start
   * guessed wrong) sends this window through the validating two-pass path */
  int ok = 1, last = -1;
  size_t total = 0;
  uint64_t end = 0;
  for (int i = 0; i < T && ok; ++i) {
    if (parts[i].status) ok = 0;
    else if (sst[i] < sst[i + 1]) {
      if (last >= 0 && parts[last].end != parts[i].first) ok = 0;
      last = i;
      total += parts[i].n;
      end = parts[i].end;
      if (parts[i].end < sst[i + 1]) { /* incomplete record: nothing after it counts */
        for (int j = i + 1; j < T; ++j) parts[j].n = 0;
        break;
      }
    }
  }
  if (getenv("TPS_FX_DEBUG"))
    fprintf(stderr, "[fastx] one-pass %.1f MB in %.4f s, %zu records, ok=%d\n", win / 1e6, dbg_now() - t_ix, total, ok);
  if (!ok || (total == 0 && !final)) {
    for (int i = 0; i < T; ++i) free(parts[i].v);
    free(parts);
    free(sst);
    if (chunk) { /* gz: the bytes already inflated become the carry of the two-pass path */
      free(fx->carry);
      fx->carry = chunk;
      fx->carry_len = win;
    }
    return next_spans_contiguous(fx, span_cap, reads_cap, bases_out, starts_out, lens_out, recs_out, n_reads, span_used,
                                 raw_base, raw_owner);
  }
  uint32_t n = 0;
  uint64_t consumed = end;
  for (int i = 0; i < T; ++i) {
    for (size_t k = 0; k < parts[i].n; ++k) {
      if (n == reads_cap) {
        consumed = parts[i].v[k].title_off - 1;
        goto capped;
      }
      const tps_fastx_rec *r = &parts[i].v[k];
      recs_out[n] = *r;
      starts_out[n] = (r->title_off - 1) >> 1;
      lens_out[n] = r->seq_len;
      ++n;
    }
  }
capped:
  for (int i = 0; i < T; ++i) free(parts[i].v);
  free(parts);
  free(sst);
  *span_used = n ? starts_out[n - 1] + lens_out[n - 1] : 0;
  if (!fx->is_gz) {
    *raw_base = w;
    fx->pos += consumed;
    if (n == 0) fx->pos = fx->map_len;
  } else {
    uint64_t left = n ? win - consumed : 0;
    uint8_t *nc = (uint8_t *)realloc(fx->carry, left ? left : 1);
    if (!nc) {
      window_free(chunk);
      return fx_fail(fx, TPS_FX_ENOMEM, "out of memory");
    }
    fx->carry = nc;
    memcpy(fx->carry, chunk + consumed, left);
    fx->carry_len = left;
    if (n) {
      *raw_base = chunk;
      *raw_owner = chunk;
    } else {
      window_free(chunk);
    }
  }
  fx->n_records += n;
  *n_reads = n;
  return TPS_FX_OK;
}

/* two-pass reader presented as a span batch (FASTA, and the fallback of the one-pass FASTQ reader) */
static int next_spans_contiguous(tps_fastx *fx, uint64_t span_cap, uint32_t reads_cap, uint8_t *bases_out,
                                 uint64_t *starts_out, uint32_t *lens_out, tps_fastx_rec *recs_out, uint32_t *n_reads,
                                 uint64_t *span_used, const uint8_t **raw_base, void **raw_owner) {
  int rc = tps_fastx_next(fx, span_cap, reads_cap, bases_out, starts_out, recs_out, n_reads, raw_base, raw_owner);
  if (rc) return rc;
  const uint32_t n = *n_reads;
  *span_used = n ? starts_out[n] : 0;
  for (uint32_t i = 0; i < n; ++i) lens_out[i] = (uint32_t)(starts_out[i + 1] - starts_out[i]);
  return TPS_FX_OK;
}

/* ---------------------------------------------------------------- rawcount CSV formatter
 * Text of `DataFrame(rows, columns=['tail','position','pattern','count']).to_csv()` as
 * rawCountPattern + main.py:150 produce it (allsteps.py:401-416, 464): header
 * ",tail,position,pattern,count", then one line "{row},{tail},{w*slide},{literal},{count}" per
 * window (major) and literal (minor), '\n' line ends.  counts = uint8[n_windows][n_patterns].
 * Returns the number of bytes written, or -(bytes needed) if cap is too small. */
static inline char *put_u64(char *p, uint64_t v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

int64_t tps_format_rawcount(const uint8_t *counts, uint32_t n_windows, uint32_t n_patterns, uint32_t slide,
                            const char *tail, const char *const *patterns, char *out, uint64_t cap) {
  static const char header[] = ",tail,position,pattern,count\n";
  size_t tl = strlen(tail), maxpat = 0;
  for (uint32_t p = 0; p < n_patterns; ++p) {
    size_t l = strlen(patterns[p]);
    if (l > maxpat) maxpat = l;
  }
  uint64_t need = sizeof(header) - 1 + (uint64_t)n_windows * n_patterns * (20 + tl + 10 + maxpat + 3 + 5);
  if (need > cap) return -(int64_t)need;
  char *p = out;
  memcpy(p, header, sizeof(header) - 1);
  p += sizeof(header) - 1;
  uint64_t row = 0;
  for (uint32_t w = 0; w < n_windows; ++w) {
    char pos[24];
    char *pe = put_u64(pos, (uint64_t)w * slide);
    size_t pl = (size_t)(pe - pos);
    for (uint32_t q = 0; q < n_patterns; ++q, ++row) {
      p = put_u64(p, row);
      *p++ = ',';
      memcpy(p, tail, tl);
      p += tl;
      *p++ = ',';
      memcpy(p, pos, pl);
      p += pl;
      *p++ = ',';
      size_t l = strlen(patterns[q]);
      memcpy(p, patterns[q], l);
      p += l;
      *p++ = ',';
      p = put_u64(p, counts[(uint64_t)w * n_patterns + q]);
      *p++ = '\n';
    }
  }
  return (int64_t)(p - out);
}

/* Indices of the records of a batch whose id equals `id` (`seq.id != read`, allsteps.py:258, 381).
 * Returns the number of matches (at most `cap` indices are stored). */
uint32_t tps_fastx_find_id(const uint8_t *raw_base, const tps_fastx_rec *recs, uint32_t n_reads, const char *id,
                           uint32_t id_len, uint32_t *out_idx, uint32_t cap) {
  uint32_t found = 0;
  for (uint32_t i = 0; i < n_reads; ++i) {
    if (recs[i].id_len != id_len) continue;
    if (memcmp(raw_base + recs[i].title_off + recs[i].id_off, id, id_len) != 0) continue;
    if (found < cap) out_idx[found] = i;
    ++found;
  }
  return found;
}

/* The ids of records idx[0..n) joined by '\n' into out (capacity cap bytes): one call instead of one string copy
 * per TRC-pass read.  Returns the bytes written, or -1 if cap is too small. */
int64_t tps_fastx_join_ids(const uint8_t *raw_base, const tps_fastx_rec *recs, const uint32_t *idx, uint32_t n,
                           uint8_t *out, uint64_t cap) {
  uint64_t at = 0;
  for (uint32_t j = 0; j < n; ++j) {
    const tps_fastx_rec *r = &recs[idx[j]];
    if (at + r->id_len + 1 > cap) return -1;
    memcpy(out + at, raw_base + r->title_off + r->id_off, r->id_len);
    at += r->id_len;
    out[at++] = '\n';
  }
  return (int64_t)at;
}

/* Region batch of the ends-first mode: for records idx[0..n) (TRC-pass reads of an ends batch) copy the first
 * (tails[j] == 0) or last (tails[j] == 1) min(L, maxlen) bases back to back into dst (capacity cap bytes), fill
 * starts / lens.  Returns how many records were placed (a prefix: the first one that does not fit stops it). */
uint32_t tps_fastx_gather_regions(const uint8_t *raw_base, const tps_fastx_rec *recs, const uint32_t *idx,
                                  const uint8_t *tails, uint32_t n, uint32_t maxlen, uint64_t cap, uint8_t *dst,
                                  uint64_t *starts, uint32_t *lens) {
  uint64_t at = 0;
  uint32_t placed = 0;
  for (; placed < n; ++placed) {
    const tps_fastx_rec *r = &recs[idx[placed]];
    const uint32_t L = r->seq_len, k = L < maxlen ? L : maxlen;
    if (at + k > cap) break;
    starts[placed] = at;
    lens[placed] = k;
    at += k;
  }
#pragma omp parallel for schedule(dynamic, 16) if (placed > 64)
  for (int64_t j = 0; j < (int64_t)placed; ++j) {
    const tps_fastx_rec *r = &recs[idx[j]];
    const uint32_t L = r->seq_len, k = lens[j];
    const uint32_t from = tails[j] ? L - k : 0u;
    if (!(r->flags & 1u)) {
      memcpy(dst + starts[j], raw_base + r->seq_off + from, k);
    } else { /* multi-line / blank-carrying sequence: filter it, then slice */
      uint8_t *tmp = (uint8_t *)malloc(L ? L : 1);
      if (tmp) {
        gather_seq(raw_base, r, tmp);
        memcpy(dst + starts[j], tmp + from, k);
        free(tmp);
      } else {
        memset(dst + starts[j], 'N', k);
      }
    }
  }
  return placed;
}

/* SeqIO.write text (main.py:86) of records idx[0..n), back to back in out: FASTQ `@title\nseq\n+\nqual\n`, FASTA
 * `>title\n` + the sequence wrapped at 60 columns.  ends[j] = end offset of record j's text.  Returns the bytes
 * written, or -1 if cap is too small.  One call (threads) for all TRC-pass reads of a batch. */
int64_t tps_fastx_records_text(const uint8_t *raw_base, const tps_fastx_rec *recs, const uint32_t *idx, uint32_t n,
                               int format, uint8_t *out, uint64_t cap, uint64_t *ends) {
  uint64_t at = 0;
  for (uint32_t j = 0; j < n; ++j) {
    const tps_fastx_rec *r = &recs[idx[j]];
    const uint64_t L = r->seq_len;
    at += format == TPS_FX_FASTQ ? 1 + (uint64_t)r->title_len + 1 + L + 3 + L + 1
                                 : 1 + (uint64_t)r->title_len + 1 + L + (L + 59) / 60;
    ends[j] = at;
  }
  if (at > cap) return -1;
#pragma omp parallel for schedule(dynamic, 8) if (n > 32)
  for (int64_t j = 0; j < (int64_t)n; ++j) {
    const tps_fastx_rec *r = &recs[idx[j]];
    const uint32_t L = r->seq_len;
    uint8_t *o = out + (j ? ends[j - 1] : 0);
    *o++ = format == TPS_FX_FASTQ ? '@' : '>';
    memcpy(o, raw_base + r->title_off, r->title_len);
    o += r->title_len;
    *o++ = '\n';
    if (format == TPS_FX_FASTQ) {
      gather_seq(raw_base, r, o);
      o += L;
      memcpy(o, "\n+\n", 3);
      o += 3;
      memcpy(o, raw_base + r->qual_off, L);
      o += L;
      *o++ = '\n';
    } else {
      uint8_t *tmp = (uint8_t *)malloc(L ? L : 1);
      if (!tmp) continue; /* cannot happen for sane L */
      gather_seq(raw_base, r, tmp);
      for (uint32_t a = 0; a < L; a += 60) {
        const uint32_t k = L - a < 60 ? L - a : 60;
        memcpy(o, tmp + a, k);
        o += k;
        *o++ = '\n';
      }
      free(tmp);
    }
  }
  return (int64_t)at;
}
