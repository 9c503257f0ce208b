/*
 * topsicle_host.h -- C-ABI of libtps_host.so, the host-side half of the drop-in: the FASTQ / FASTA reader
 * that feeds pinned batches to libtopsicle_b200.so (include/topsicle_b200.h), the text formatters of the
 * reference's output files, and the synth-v1 workload generator of bench.py.  No CUDA in here.
 *
 * The reference (jaeyoungchoilab/Topsicle) has no FFI; every entry point below names the reference code it
 * replaces (file:line in the reference tree).  Plain pointers and sizes only.  Functions that return int
 * return TPS_FX_OK or a negative TPS_FX_E* code and leave a message for tps_fastx_last_error; they never throw.
 * A reader handle is used by one thread at a time (its parsing is itself multi-threaded, OpenMP).
 */
#ifndef TOPSICLE_HOST_H
#define TOPSICLE_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* input formats (allsteps.py:36-50 check_file_type: first character '@' or '>') */
#define TPS_FX_FASTQ 1
#define TPS_FX_FASTA 2

#define TPS_FX_OK 0
#define TPS_FX_EIO (-1)       /* open / read / inflate failure */
#define TPS_FX_EFORMAT (-2)   /* neither FASTQ nor FASTA, or a malformed record (the records before it were delivered) */
#define TPS_FX_ENOMEM (-3)
#define TPS_FX_ECAPACITY (-4) /* one record does not fit the caller's batch buffers */
#define TPS_FX_EINVAL (-5)

/* One record of a batch: where its title, sequence and quality text lie in the batch's raw text
 * (`*raw_base` of the tps_fastx_next* call).  This is what Bio.SeqIO's record object carries for the
 * reference: `record.id` = title up to the first blank, `record.description` = the whole title
 * (main.py:83-86 re-writes both through SeqIO.write).  48 bytes. */
typedef struct tps_fastx_rec {
  uint64_t title_off;   /* first title byte (after '@' / '>') */
  uint64_t seq_off;     /* first byte of the first sequence line */
  uint64_t qual_off;    /* FASTQ: first byte of the quality line; FASTA: 0 */
  uint32_t title_len;   /* right-stripped */
  uint32_t id_off;      /* id = title[id_off : id_off + id_len] */
  uint32_t id_len;
  uint32_t seq_len;     /* bases after stripping */
  uint32_t seq_raw_len; /* raw bytes spanned by the sequence lines (FASTA: incl. newlines) */
  uint32_t flags;       /* bit 0: sequence needs the filtered copy (multi-line or inner blanks) */
} tps_fastx_rec;

typedef struct tps_fastx tps_fastx;

/* Open `path` (plain, gzip or BGZF by content; FASTQ or FASTA by its first character), `threads` parser threads.
 * Replaces: check_file_type + unzip_file's open (allsteps.py:36-50, 127-146). */
int tps_fastx_open(tps_fastx **out, const char *path, int threads);
void tps_fastx_close(tps_fastx *fx);
/* TPS_FX_FASTQ / TPS_FX_FASTA. */
int tps_fastx_format(const tps_fastx *fx);
/* Last error text of `fx` (or of a failed tps_fastx_open when fx == NULL; thread-local). */
const char *tps_fastx_last_error(const tps_fastx *fx);
/* Raw bytes examined per call (default 512 MiB) / force the validating two-pass reader (tests, tuning). */
void tps_fastx_set_window(tps_fastx *fx, uint64_t bytes);
/* A record with more bases than a whole batch can hold normally ends the file with TPS_FX_ECAPACITY.  With
 * clip_bases > 0, tps_fastx_next / tps_fastx_next_spans deliver such a record as a batch of its own that holds its
 * first and last clip_bases bases back to back (its tps_fastx_rec still describes the whole record).  The scan is
 * unchanged by that when clip_bases >= max(maxlengthtelo, 1000) and 2 * clip_bases > minSeqLength: it looks at
 * no other base (reference: Topsicle/allsteps.py:176-177, 266-268). */
void tps_fastx_set_clip(tps_fastx *fx, uint64_t clip_bases);
void tps_fastx_set_two_pass(tps_fastx *fx, int on);

/* Next batch of records in file order: at most reads_cap records and bases_cap bases, bases of record i at
 * bases_out[offsets_out[i] .. offsets_out[i+1]) (offsets_out holds reads_cap + 1 entries), recs_out[i] relative to
 * *raw_base, which stays valid until tps_fastx_release(*raw_owner) (compressed input) or tps_fastx_close (plain
 * input: a view of the file mapping, *raw_owner == NULL).  *n_reads == 0: end of file.
 * Replaces: `for record in SeqIO.parse(handle, fmt)` (allsteps.py:143-146, 174; main.py:83; allsteps.py:252). */
int tps_fastx_next(tps_fastx *fx, uint64_t bases_cap, uint32_t reads_cap, uint8_t *bases_out, uint64_t *offsets_out,
                   tps_fastx_rec *recs_out, uint32_t *n_reads, const uint8_t **raw_base, void **raw_owner);
/* The same as a span batch for tps_submit_spans: read i = bases_out[starts_out[i] .. + lens_out[i]), *span_used
 * bytes of bases_out meaningful.  4-line FASTQ is read in ONE pass (a record at window byte x is placed at x / 2). */
int tps_fastx_next_spans(tps_fastx *fx, uint64_t span_cap, uint32_t reads_cap, uint8_t *bases_out, uint64_t *starts_out,
                         uint32_t *lens_out, tps_fastx_rec *recs_out, uint32_t *n_reads, uint64_t *span_used,
                         const uint8_t **raw_base, void **raw_owner);
/* Ends batch for tps_submit_ends: of every record only the first and last end_len bases (the whole read when it
 * has at most 2 * end_len), true_lens_out[i] = its real length; at most raw_cap bytes of file text per call.
 * Step 1 reads nothing else of a read (allsteps.py:176-177). */
int tps_fastx_next_ends(tps_fastx *fx, uint64_t raw_cap, uint64_t bases_cap, uint32_t reads_cap, uint32_t end_len,
                        uint8_t *bases_out, uint64_t *starts_out, uint32_t *lens_out, uint32_t *true_lens_out,
                        tps_fastx_rec *recs_out, uint32_t *n_reads, uint64_t *span_out, uint64_t *true_bases_out,
                        const uint8_t **raw_base, void **raw_owner);
void tps_fastx_release(void *owner);

/* Indices (at most cap are stored) of the records whose id equals `id`; returns how many match.
 * Replaces: `if record.id == read` (allsteps.py:258, 381). */
uint32_t tps_fastx_find_id(const uint8_t *raw_base, const tps_fastx_rec *recs, uint32_t n_reads, const char *id,
                           uint32_t id_len, uint32_t *out_idx, uint32_t cap);
/* ids of records idx[0..n) joined by '\n' into out; bytes written or -1 (cap too small).  (record.id, main.py:85,138) */
int64_t tps_fastx_join_ids(const uint8_t *raw_base, const tps_fastx_rec *recs, const uint32_t *idx, uint32_t n,
                           uint8_t *out, uint64_t cap);
/* Region batch for tps_submit_regions: the first (tails[j] == 0) or last (1) min(L, maxlen) bases of records idx[j],
 * packed into dst (cap bytes) with starts / lens; returns how many records fit (a prefix).
 * These are the only bases steps 2/3 read (allsteps.py:263-271, 395-396). */
uint32_t tps_fastx_gather_regions(const uint8_t *raw_base, const tps_fastx_rec *recs, const uint32_t *idx,
                                  const uint8_t *tails, uint32_t n, uint32_t maxlen, uint64_t cap, uint8_t *dst,
                                  uint64_t *starts, uint32_t *lens);
/* Text of SeqIO.write(record, handle, "fastq" | "fasta") for records idx[0..n), back to back in out; ends[j] = end
 * of record j's text.  Bytes written or -1.  Replaces: the subset-file loop (main.py:82-86). */
int64_t tps_fastx_records_text(const uint8_t *raw_base, const tps_fastx_rec *recs, const uint32_t *idx, uint32_t n,
                               int format, uint8_t *out, uint64_t cap, uint64_t *ends);

/* Text of `rawCountPattern(...).to_csv()` (allsteps.py:401-416, 464; main.py:150) for counts[n_windows][n_patterns]
 * (uint8): header ",tail,position,pattern,count", one line per window (major) and literal (minor), '\n' line ends.
 * Bytes written, or -(bytes needed) if cap is too small. */
int64_t tps_format_rawcount(const uint8_t *counts, uint32_t n_windows, uint32_t n_patterns, uint32_t slide,
                            const char *tail, const char *const *patterns, char *out, uint64_t cap);

/* ---- parallel inflate of plain gzip (csrc/tps_pgz.c), what the reader uses for `.gz` input that is not BGZF.
 * Replaces: gzip.open(filepath, 'rt') in unzip_file (allsteps.py:142-146), one zlib stream on one core.
 * The compressed file is cut into one piece per thread; every thread finds the first deflate block that starts in
 * its piece, decodes from there with an unknown 32 KiB window into 16-bit symbols, the pieces must chain exactly,
 * then all are resolved to bytes in parallel; CRC-32 and length of every gzip member are verified. */
typedef struct tps_pgz tps_pgz;
typedef struct tps_pgz_stats {
  uint64_t stretches;           /* calls that inflated a stretch of the file */
  uint64_t segments;            /* pieces decoded (one per thread and stretch when every block-start search succeeds) */
  uint64_t chain_breaks;        /* pieces thrown away because the piece before did not land on their start */
  uint64_t members;             /* gzip members completed (CRC-32 and length verified) */
  uint64_t text_bytes;          /* inflated bytes */
  uint64_t parallel_text_bytes; /* ... of which decoded from a guessed block start with an unknown window */
} tps_pgz_stats;
/* zmap[0..zlen) = the whole compressed file (mapped).  NULL if it does not start with a gzip member header. */
tps_pgz *tps_pgz_open(const uint8_t *zmap, uint64_t zlen, int threads);
void tps_pgz_close(tps_pgz *g);
/* Up to cap bytes of inflated text into dst; 0 = end of file; -1 = corrupt input / out of memory (tps_pgz_error). */
int64_t tps_pgz_read(tps_pgz *g, uint8_t *dst, uint64_t cap);
int tps_pgz_eof(const tps_pgz *g);
const char *tps_pgz_error(const tps_pgz *g);
void tps_pgz_get_stats(const tps_pgz *g, tps_pgz_stats *out);
void tps_pgz_set_piece(tps_pgz *g, uint64_t bytes); /* compressed bytes per thread and stretch (tests, tuning) */
/* One complete raw deflate stream with an empty window that inflates to exactly `want` <= 65536 bytes (a BGZF
 * block): bytes to dst, their CRC-32 to *crc_out.  0 ok, -1 corrupt or of another length, -3 not applicable. */
int tps_pgz_inflate_block(const uint8_t *z, uint64_t zlen, uint8_t *dst, uint64_t want, uint32_t *crc_out);
/* The same counters of the reader's own inflater (all zero unless `fx` reads plain gzip). */
void tps_fastx_inflate_stats(const tps_fastx *fx, tps_pgz_stats *out);

/* ---- synth-v1 workload generator (no reference counterpart: the reference ships no benchmark) */
typedef struct tps_synth_cfg {
  uint64_t seed;
  uint32_t len_kind;      /* 0 fixed(len_a); 1 lognormal(mu=len_a, sigma=len_b) clipped [len_min,len_max];
                             2 len_a + Exp(mean len_b) capped len_max */
  double len_a, len_b;
  uint32_t len_min, len_max;
  double f_telo;          /* fraction of telomeric reads */
  uint32_t telo_min, telo_max;
  double sub_rate, ins_rate, del_rate;
  double n_rate;          /* per-base probability of 'N' */
  double near_frac;       /* fraction of near-threshold reads */
  double lower_frac;      /* fraction of reads with a 500-base lower-case stretch */
  uint32_t motif_len;
  char motif[32];
} tps_synth_cfg;

/* offsets_out[0] = 0, offsets_out[i+1] = offsets_out[i] + length of read first_read + i */
int tps_synth_lengths(const tps_synth_cfg *c, uint64_t first_read, uint32_t n_reads, uint64_t *offsets_out);
/* bases of reads [first_read, first_read + n_reads) at bases_out[offsets[i] ..); kinds_out (optional) = read class */
int tps_synth_fill(const tps_synth_cfg *c, uint64_t first_read, uint32_t n_reads, const uint64_t *offsets,
                   uint8_t *bases_out, uint8_t *kinds_out, int n_threads);
int tps_host_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* TOPSICLE_HOST_H */
