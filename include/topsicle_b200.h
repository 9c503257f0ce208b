/*
 * topsicle_b200.h -- C-ABI of the B200-native per-read telomere scan.
 *
 * The reference (jaeyoungchoilab/Topsicle) has no FFI: its boundary for this path is
 * the Python functions star-exported from Topsicle/allsteps.py and driven per file by
 * Topsicle/main.py:process_file.  Each entry point below names the reference
 * interface it replaces (file:line in the reference tree).  Plain pointers and
 * sizes only; no torch / C++ types.  All functions return 0 on success or a negative
 * TPS_E* code; they never throw.  One context per (device, host thread).
 *
 * Batch model: a batch is `n_reads` reads whose bases are concatenated back to back
 * in one byte buffer (ASCII, any case, any byte value; only ACGTacgt can match) with
 * `offsets[n_reads + 1]` giving each read's start (offsets[0] == 0, offsets[n] ==
 * total bases).  One call scans the whole batch on the GPU:
 *   K1 ingest/pack   -> replaces str.upper() + the regex engine's view of the bytes
 *                       (allsteps.py:176-177, 267-271)
 *   K2 TRC           -> allsteps.py:175-198   (patternTRC_count, per read)
 *   K3 windows       -> allsteps.py:207-225, 275-297, 395-416 (seq_cut_windows, counts)
 *   K4 change point  -> allsteps.py:304-315 + ruptures 1.1.9 Binseg(l2).predict(n_bkps=1)
 */
#ifndef TOPSICLE_B200_H
#define TOPSICLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPS_ABI_VERSION 1
#define TPS_MAX_PATTERNS 64     /* literals per scan (reference: 2 * unique k-mers, <= 2*len(pattern)) */
#define TPS_MAX_PATTERN_LEN 32  /* bases per literal */

/* error codes */
#define TPS_OK 0
#define TPS_EINVAL (-1)    /* bad argument / unsupported parameter combination */
#define TPS_ECUDA (-2)     /* CUDA runtime error (see tps_last_error) */
#define TPS_ENOMEM (-3)    /* allocation failed */
#define TPS_ECAPACITY (-4) /* batch larger than max_batch_* or rawcount buffer too small */
#define TPS_ESTATE (-5)    /* wait on unknown batch id / slot busy */
#define TPS_ENODEVICE (-6) /* no CUDA device: there is NO CPU fallback */

/* tps_row.status */
#define TPS_ST_FILTERED 0 /* L <= min_seq_length                       (allsteps.py:175) */
#define TPS_ST_BELOW 1    /* scanned, TRC <= cutoff                    (allsteps.py:194,197) */
#define TPS_ST_PASS 2     /* TRC > cutoff, change point found          (allsteps.py:330-331) */
#define TPS_ST_BADSEG 3   /* TRC > cutoff but fewer than 7 windows: the reference raises
                             ruptures.BadSegmentationParameters (n>=1) or IndexError (n==0) */

/* tps_params.flags */
#define TPS_FLAG_STEP1_ONLY 1u    /* patternTRC_count alone: no windows / change point (allsteps.py:152-204) */
#define TPS_FLAG_FORCE_FORWARD 2u /* bound_detect / rawCountPattern called with tail='forward': the caller,
                                     not step 1, picks the end (allsteps.py:294-297, 413-416) */
#define TPS_FLAG_FORCE_REVERSE 4u /* ... tail='reverse' */

/* tps_row.tail */
#define TPS_TAIL_FORWARD 0
#define TPS_TAIL_REVERSE 1

/* Scan parameters == the arguments main.py:process_file passes down
 * (main.py:57, 129-130, 147-148) plus capacities. */
typedef struct tps_params {
  uint32_t struct_size;     /* = sizeof(tps_params) */
  uint32_t n_patterns;      /* P; literals in reference order (patterns_to_search, allsteps.py:84-125) */
  uint8_t pattern_len[TPS_MAX_PATTERNS];
  char patterns[TPS_MAX_PATTERNS][TPS_MAX_PATTERN_LEN]; /* upper-case ACGT, not NUL-terminated */
  uint32_t min_seq_length;  /* keep reads with L >  min_seq_length          (allsteps.py:175) */
  uint32_t no_bp;           /* head/tail length, reference uses 1000        (main.py:57) */
  uint32_t count_threshold; /* pass iff best count >= this; host computes the smallest c with
                               c / (no_bp / len(pattern)) > cutoff in float64 (allsteps.py:178-198) */
  uint32_t window_size;     /* W; window text is W-1 bases                  (allsteps.py:219-224) */
  uint32_t slide;           /* s                                            (main.py:212-215) */
  uint32_t trimfirst;       /* t                                            (allsteps.py:267-271) */
  uint32_t maxlengthtelo;   /* M = min(maxlengthtelo, L)                    (allsteps.py:263-264) */
  uint32_t want_rawcount;   /* also return counts[w][p]                     (allsteps.py:401-416) */
  uint32_t n_slots;         /* batches in flight (1..4), async pipeline depth */
  uint32_t max_batch_reads; /* capacity per batch */
  uint32_t max_pass_reads;  /* capacity for TRC-pass reads per batch (0 = max_batch_reads); each costs
                               4 bytes x windows-per-read of device memory */
  uint32_t flags;           /* TPS_FLAG_* */
  uint64_t max_batch_bases; /* capacity per batch (bytes of sequence) */
  uint64_t rawcount_capacity; /* per batch, in count elements (uint8 each); 0 if !want_rawcount */
} tps_params;

/* One row per input read, in input order (the reference emits rows in file order,
 * main.py:125-138). 40 bytes. */
typedef struct tps_row {
  uint32_t length;       /* L */
  uint8_t status;        /* TPS_ST_* */
  uint8_t tail;          /* TPS_TAIL_*          (allsteps.py:193-198) */
  uint8_t best_pattern;  /* first-max literal index in the chosen end (allsteps.py:190-191) */
  uint8_t reserved0;
  uint16_t match_count;  /* count of that literal: TRC = match_count / (no_bp/len(pattern)) */
  uint16_t head_max;     /* max_p count in seq[:no_bp] */
  uint16_t tail_max;     /* max_p count in reversed seq[-no_bp:] */
  uint16_t reserved1;
  uint32_t n_windows;    /* nW of the chosen end (0 unless status >= PASS) */
  int32_t bkp;           /* change point index b*, -1 if none */
  int32_t telo_length;   /* trimfirst + slide * b*  (allsteps.py:312-315), -1 if none */
  uint32_t reserved2;
  uint64_t rawcount_offset; /* element offset of counts[0][0] in the rawcount buffer, ~0 if none */
} tps_row;

typedef struct tps_ctx tps_ctx;

/* Library identity. */
int tps_abi_version(void);
const char *tps_build_info(void);

/* Number of CUDA devices visible to the library (0 if none / no driver). */
int tps_device_count(void);

/* Create / destroy a scan context on CUDA device `device`.
 * Replaces: the per-call setup in patternTRC_count / bound_detect
 * (re.compile of every literal, allsteps.py:167-168, 240-241). */
int tps_create(tps_ctx **out, int device, const tps_params *params);
void tps_destroy(tps_ctx *ctx);
/* Last error text for `ctx` (or for a failed tps_create when ctx == NULL). */
const char *tps_last_error(const tps_ctx *ctx);

/* Pinned host staging memory for tps_submit (pageable memory also works, slower). */
void *tps_alloc_pinned(size_t bytes);
void tps_free_pinned(void *p);

/* Asynchronous scan of one batch held in HOST memory.  The caller keeps `bases` and
 * `offsets` alive and unmodified until tps_wait(batch_id) returns.
 * Replaces: process_file's step 1 + per-read step 2/3 loop (main.py:57, 125-150). */
int tps_submit(tps_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads,
               uint64_t batch_id);
/* Same for a batch whose reads are NOT packed back to back: read i is bases[starts[i] .. starts[i] +
 * lengths[i]); reads do not overlap, anything between them is ignored, `n_span` bytes of `bases` are
 * uploaded.  This is what a one-pass parser produces (it can place every read without first knowing the
 * lengths of all reads before it, see csrc/tps_fastx.c).  Same lifetime rules as tps_submit. */
int tps_submit_spans(tps_ctx *ctx, const uint8_t *bases, uint64_t n_span, const uint64_t *starts,
                     const uint32_t *lengths, uint32_t n_reads, uint64_t batch_id);
/* Ends-first scanning (optional, for inputs whose reads are mostly NOT telomeric): step 1 only looks at the
 * first and last `no_bp` bases of a read (allsteps.py:176-177) and steps 2/3 only at `[trimfirst,
 * min(L, maxlengthtelo))` of one end of the TRC-pass reads (allsteps.py:263-271), so the interior of long
 * reads never has to cross PCIe.
 *   tps_submit_ends: read i is uploaded as its head followed by its tail, `lengths[i]` = min(L, 2*no_bp)
 *   bytes (the whole read when L <= 2*no_bp), `true_lengths[i]` = L.  Runs K1 + K2 only; the rows carry
 *   the real length and status FILTERED / BELOW / PASS (n_windows, bkp, telo_length unset).
 *   tps_submit_regions: read i is the first (tails[i] = TPS_TAIL_FORWARD) or last (TPS_TAIL_REVERSE)
 *   min(L, maxlengthtelo) bases of a read that passed step 1 with that tail.  Runs K1..K4 with the tail
 *   forced per read and the length filter off; n_windows, bkp, telo_length, status PASS / BADSEG and the raw
 *   counts equal those of a whole-read scan; the step-1 fields of these rows (match_count, head_max,
 *   tail_max, best_pattern) describe the uploaded region only -- keep those of the ends batch.  Same span-batch layout and lifetime rules as tps_submit_spans. */
int tps_submit_ends(tps_ctx *ctx, const uint8_t *bases, uint64_t n_span, const uint64_t *starts,
                    const uint32_t *lengths, const uint32_t *true_lengths, uint32_t n_reads, uint64_t batch_id);
int tps_submit_regions(tps_ctx *ctx, const uint8_t *bases, uint64_t n_span, const uint64_t *starts,
                       const uint32_t *lengths, const uint8_t *tails, uint32_t n_reads, uint64_t batch_id);
/* Scan, under the parameters of `ctx`, the batch that `owner` has in flight as `batch_id`, without a
 * second upload or a second K1: ctx's K2..K4 are enqueued on the owner's stream and read the owner's
 * packed reads.  Same device; collect with tps_wait(ctx, batch_id, ...).  This is how several
 * telophrases are scanned from one pass over the input.
 * Replaces: the reference's outer loop over telo_phrases, which re-parses and re-scans the whole
 * dataset once per phrase (main.py:206-235). */
int tps_submit_shared(tps_ctx *ctx, tps_ctx *owner, uint64_t batch_id);
/* Block until batch `batch_id` is done; copies n_reads rows to rows_out.  If the context
 * was created with want_rawcount, copies *rawcount_elems count elements (uint8,
 * [n_windows][P] per PASS read at tps_row.rawcount_offset) to rawcounts_out, whose capacity
 * in elements is rawcount_cap (TPS_ECAPACITY if too small; call again with a larger one). */
int tps_wait(tps_ctx *ctx, uint64_t batch_id, tps_row *rows_out, uint32_t *n_pass_out,
             uint8_t *rawcounts_out, uint64_t rawcount_cap, uint64_t *rawcount_elems);

/* Block until batch `batch_id` is done and report its sizes WITHOUT releasing it, so that the
 * caller can size rawcounts_out for tps_wait: *n_pass_out = reads that passed the TRC cutoff,
 * *rawcount_elems = count elements the batch produced. */
int tps_batch_info(tps_ctx *ctx, uint64_t batch_id, uint32_t *n_pass_out, uint64_t *rawcount_elems);

/* Scan a batch already resident in DEVICE memory (d_bases must be 16-byte aligned, readable up to
 * n_bases rounded up to a multiple of 2048 bytes, and unchanged until the scan has finished: K1 reads it
 * through the bulk-copy engine, K2 / K3 re-read the ASCII bytes of groups that hold a non-ACGT byte);
 * rows are written to d_rows_out (device).
 * Enqueued on the context's stream 0; returns without synchronising. */
int tps_scan_device(tps_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets,
                    uint32_t n_reads, uint64_t n_bases, tps_row *d_rows_out);
/* Same, on batch slot `slot` (0 .. n_slots-1): every slot has its own stream and scratch buffers, so
 * scans of independent device-resident batches issued on different slots may overlap on the GPU. */
int tps_scan_device_slot(tps_ctx *ctx, uint32_t slot, const uint8_t *d_bases, const uint64_t *d_offsets,
                         uint32_t n_reads, uint64_t n_bases, tps_row *d_rows_out);
int tps_sync(tps_ctx *ctx);

/* CUDA-event timings (ms), recorded on the scan's own stream, of the tps_scan_device call
 * `back` calls ago (0 = most recent; a ring of TPS_TIMING_RING calls is kept):
 * ms[0] = K1 pack, ms[1] = K2 TRC, ms[2] = K3+K4 windows/change point, ms[3] = whole scan.
 * Blocks until that scan has finished. */
#define TPS_N_TIMINGS 4
#define TPS_TIMING_RING 256
int tps_get_timings(tps_ctx *ctx, uint32_t back, float ms[TPS_N_TIMINGS]);
/* Timeline of overlapping scans: ms[i] = time from the start of the timed scan `base_back` calls ago to
 * event i of the scan `back` calls ago (0 = K1 start, 1 = K1 end, 2 = K2 end, 3 = K4 end). */
int tps_get_timeline(tps_ctx *ctx, uint32_t back, uint32_t base_back, float ms[TPS_N_TIMINGS]);
/* Device time (ms) from event `from_event` of the tps_scan_device call `from_back` calls ago of context `from` to
 * event `to_event` of the call `to_back` calls ago of context `to` (same device; events as in tps_get_timeline):
 * the span of a run of overlapping scans issued through several contexts, measured on the device.  Blocks until both
 * scans have finished. */
int tps_elapsed_between(tps_ctx *from, uint32_t from_back, uint32_t from_event, tps_ctx *to, uint32_t to_back,
                        uint32_t to_event, float *ms);
/* Number of kernels this context has launched so far. */
uint64_t tps_kernel_launches(const tps_ctx *ctx);

/* Overview heat map (the `next` row f4 of SURVEY 8): the data half of
 * Topsicle/descriptive_plot.py:233-313 `patterns_vs_match_heatmap`, i.e. what `overview_plot.py
 * --recfindingpattern --rawcount` writes to heatmap_rawcount_{i}.csv.  For every read longer than
 * min_seq_length and every k-mer of `patterns` (n_patterns x k upper-case ACGT characters, the ORIGIN k-mers of
 * pattern_scramble_telo, no complements): the leftmost non-overlapping matches of `kmer(.{match_len - k})` in
 * `seq[skip:upto]` (strand 0) and in the complement of `reversed(seq)[skip:upto]` (strand 1).
 * sel_out[((read * 2 + strand) * n_patterns + p) * W + w], W = ceil((upto - skip) / 32): bit b of word w set
 * <=> a match starts at position 32 w + b of that slice.  Synchronous, stateless (own buffers, K1 + one
 * warp per (read, strand)); reads are back to back in `bases` as for tps_submit.  Errors: tps_last_error(NULL). */
int tps_follow_scan(int device, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads, const char *patterns,
                    uint32_t n_patterns, uint32_t k, uint32_t match_len, uint32_t min_seq_length, uint32_t skip,
                    uint32_t upto, uint32_t *sel_out, uint64_t sel_words);

/* Test hooks: copy internal device arrays of slot 0 to the host after a scan.
 * what: 0 = 2-bit code words (uint32 per 16 bases), 1 = invalid-group flag words
 * (uint32 per 512 bases), 2 = validity masks as K2/K3 see them (uint16 per 16 bases: 0xFFFF for an
 * unflagged group, else rebuilt from the group's ASCII bytes), 3 = pass list (uint32 read indices, unordered),
 * 4 = window sums (row i = pass-list entry i, row stride in elements from 5): uint32 c_w per window under the
 * plain window kernel; under the bit-parallel kernel, which keeps them in shared memory, uint16 sums of c_w over
 * the groups of five windows [5j, 5j+5) -- what its change point reads -- and only if the context was created with
 * TPS_K3_DEBUG_GS=1 in the environment (else TPS_ESTATE), 5 = uint32[4] {row stride of 4, 1 if the bit-parallel
 * window kernel is in use, pass-list capacity, window-start positions per tile}. */
int tps_debug_copy(tps_ctx *ctx, int what, void *dst, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* TOPSICLE_B200_H */
