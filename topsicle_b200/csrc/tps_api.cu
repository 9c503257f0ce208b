/*
 * tps_api.cu -- C-ABI (include/topsicle_b200.h) over the sm_100a kernels.
 *
 * One tps_ctx owns `n_slots` independent batch slots, each with its own CUDA stream,
 * device buffers and pinned result staging, so that H2D of batch i+1 overlaps the kernels
 * of batch i and the D2H of batch i-1 (async H2D / compute / D2H pipeline per device).
 * There is no CPU fallback: without a CUDA device tps_create fails with TPS_ENODEVICE.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>

#include "tps_kernels.cuh"

namespace {

thread_local char g_create_error[512] = "";

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t e_in = nullptr, e_k1 = nullptr, e_tail = nullptr; /* split mode: slot -> pack -> tail -> slot */
  cudaEvent_t done = nullptr;
  uint8_t *d_bases = nullptr;
  uint32_t *d_codes = nullptr;
  uint32_t *d_flags = nullptr;
  const uint8_t *last_bases = nullptr; /* ASCII batch of the last scan (slot- or caller-owned) */
  uint64_t *d_off = nullptr;
  uint32_t *d_len = nullptr;        /* read lengths of a span batch */
  bool has_lens = false;
  uint32_t *d_true_len = nullptr;   /* ends batch: real read lengths */
  uint8_t *d_tails = nullptr;       /* region batch: per-read forced tail */
  bool ends = false, regions = false; /* kind of the batch in flight */
  tps_row *d_rows = nullptr;
  uint32_t *d_pass = nullptr;
  uint32_t *d_counters = nullptr;
  uint8_t *d_raw = nullptr;
  uint32_t *d_cw = nullptr;         /* c_w rows of passing reads */
  TpsReadItem *d_items = nullptr;   /* bit-parallel K3: one record per TRC-pass read, listed by K2 */
  uint32_t *d_tile_done = nullptr;  /* bit-parallel K3, split mode: CTAs done per passing read */
  uint16_t *d_gs_debug = nullptr;   /* TPS_K3_DEBUG_GS=1: copy of every read's group sums */
  tps_row *h_rows = nullptr;        /* pinned */
  uint32_t *h_counters = nullptr;   /* pinned */
  uint64_t batch_id = 0;
  uint32_t n_reads = 0;
  bool busy = false;
};

}  // namespace

struct tps_ctx {
  int device = 0;
  tps_params p;
  TpsPatTable pt;
  int n_sms = 0;
  uint32_t k2_lin_words = 0, k2_nq_max = 0, k2_smem = 0;
  uint32_t k3_lin_words = 0, k3_tile_words = 0, k3_tile_bases = 0, k3_tiles_max = 0, k3_smem = 0, k3_grid = 0;
  uint32_t k4_grid = 0;
  /* bit-parallel K3 with the change point fused (tps_window_bp_kernel): the default whenever the literals are
   * distinct strings of one length K <= 8, W - K >= 32, W <= 2048 and no raw-count tables are wanted */
  bool k3_bitpar = false;
  void (*k3n_fn)(const TpsScanArgs, const TpsPatTable) = nullptr;
  uint32_t k3n_lin_words = 0, k3n_stride = 0, k3n_tile_bases = 0, k3n_tiles_max = 0, k3n_smem = 0, k3n_grid = 0, k3n_nz = 0;
  uint32_t k3n_gs_cap = 0, k3n_tile_windows = 0, k3n_no_groups = 0;
  bool k3n_debug_gs = false; /* TPS_K3_DEBUG_GS=1: the group sums of every read are also written out (tests) */
  uint32_t k3n_no_split = 0;
  uint32_t cw_stride = 0, max_pass = 0;
  int kt = 0; /* template K of the K2/K3 instantiation in use (0 = generic) */
  void (*k2_fn)(const TpsScanArgs, const TpsPatTable) = nullptr;
  void (*k2r_fn)(const TpsScanArgs, const TpsPatTable) = nullptr; /* register-staged K2 (no_bp <= 1000, P <= 32) */
  /* K2 with the literal set in the instruction stream (constant-bank masks, unrolled): complement-paired sets of
   * 5..8 literals of one length 3..8 without self-overlap -- every default CLI run */
  void (*k2c_fn)(const TpsScanArgs, const TpsPairMasks) = nullptr;
  TpsPairMasks k2c_masks;
  bool k2_reg = false;
  uint32_t k2r_smem = 0;
  void (*k3_fn)(const TpsScanArgs, const TpsPatTable) = nullptr;
  int k1_grid = 0, k1_unroll = 4;
  /* split mode (default): every K1 of the context runs on `pack_stream` (low priority), every K2..K4 on
   * `tail_stream` (high priority), so the tail kernels of batch i run under the K1 of batch i+1 -- the
   * HBM-bound pack kernel leaves issue slots and registers free, the issue-bound tail kernels use them --
   * and never more than one K1 is resident.  The slot streams keep the copies and the ordering. */
  bool split = false;
  cudaStream_t pack_stream = nullptr, tail_stream = nullptr;
  bool k1_tma = false;          /* K1 through the bulk-copy engine (tps_pack_tma_kernel) */
  uint32_t k1t_stages = 0, k1t_smem = 0, k1t_unroll = 4; /* stage = 4 * k1t_unroll KiB */
  int k1t_grid = 0;
  void (*k1t_fn)(const uint4 *, uint32_t *, uint32_t *, uint64_t, uint32_t) = nullptr;
  void (*k1_fn)(const uint4 *, uint32_t *, uint32_t *, uint64_t) = nullptr;
  uint64_t cap_tiles = 0;
  Slot slots[4];
  cudaEvent_t ev[TPS_TIMING_RING][4]; /* CUDA-event ring: one set of 4 events per timed scan */
  uint64_t scan_seq = 0;               /* number of timed scans enqueued so far */
  uint64_t launches = 0;
  char err[512] = "";
};

namespace {

/* The pack / tail stream pair is one per DEVICE, shared by every context of the process on that device
 * (reference-counted): the contexts of several pattern sets or telophrases then queue their K1s on one stream --
 * never two persistent HBM-bound pack kernels resident at once -- and their K2..K4 on the other. */
struct DevStreams {
  cudaStream_t pack = nullptr, tail = nullptr;
  int refs = 0;
};
std::mutex g_dev_mutex;
DevStreams g_dev_streams[64];

int acquire_dev_streams(int device, cudaStream_t *pack, cudaStream_t *tail) {
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  DevStreams &d = g_dev_streams[device & 63];
  if (d.refs == 0) {
    int pr_lo = 0, pr_hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi) != cudaSuccess) return -1;
    if (cudaStreamCreateWithPriority(&d.pack, cudaStreamNonBlocking, pr_lo) != cudaSuccess) return -1;
    if (cudaStreamCreateWithPriority(&d.tail, cudaStreamNonBlocking, pr_hi) != cudaSuccess) {
      cudaStreamDestroy(d.pack);
      d.pack = nullptr;
      return -1;
    }
  }
  ++d.refs;
  *pack = d.pack;
  *tail = d.tail;
  return 0;
}

void release_dev_streams(int device) {
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  DevStreams &d = g_dev_streams[device & 63];
  if (d.refs > 0 && --d.refs == 0) {
    cudaStreamSynchronize(d.pack);
    cudaStreamSynchronize(d.tail);
    cudaStreamDestroy(d.pack);
    cudaStreamDestroy(d.tail);
    d.pack = d.tail = nullptr;
  }
}

int fail(tps_ctx *ctx, int code, const char *fmt, ...) {
  char *dst = ctx ? ctx->err : g_create_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define TPS_CUDA(ctx, call)                                                                     \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? TPS_ENOMEM : TPS_ECUDA, "%s: %s", #call, \
                  cudaGetErrorString(e_));                                                      \
  } while (0)

uint32_t lin_words_for(uint64_t n_max) {
  uint64_t ng = (n_max + 30) / 16 + 1;
  return (uint32_t)((2 + ng + 1) / 2 + 2);
}

int build_pattern_table(const tps_params *p, TpsPatTable *pt) {
  memset(pt, 0, sizeof(*pt));
  if (p->n_patterns < 1 || p->n_patterns > TPS_MAX_PATTERNS)
    return fail(nullptr, TPS_EINVAL, "n_patterns must be in 1..%d", TPS_MAX_PATTERNS);
  pt->n = p->n_patterns;
  for (uint32_t i = 0; i < p->n_patterns; ++i) {
    uint32_t k = p->pattern_len[i];
    if (k < 1 || k > TPS_MAX_PATTERN_LEN)
      return fail(nullptr, TPS_EINVAL, "pattern %u: length %u not in 1..%d", i, k, TPS_MAX_PATTERN_LEN);
    pt->len[i] = (uint8_t)k;
    for (uint32_t j = 0; j < k; ++j) {
      uint32_t c = tps_ascii_code((uint8_t)p->patterns[i][j]);
      if (c == 0xFFu)
        return fail(nullptr, TPS_EINVAL, "pattern %u has a non-ACGT character (regex metacharacters and "
                                         "'|' patterns are undefined in the reference)", i);
      pt->lo[i] |= (c & 1u) << j;
      pt->hi[i] |= ((c >> 1) & 1u) << j;
    }
    /* proper border <=> two occurrences can overlap <=> greedy != count of all occurrences */
    bool bordered = false;
    for (uint32_t b = 1; b < k && !bordered; ++b) {
      bool eq = true;
      for (uint32_t j = 0; j < b && eq; ++j)
        eq = tps_ascii_code((uint8_t)p->patterns[i][j]) == tps_ascii_code((uint8_t)p->patterns[i][k - b + j]);
      bordered = eq;
    }
    pt->bordered[i] = bordered;
    if (bordered) {
      pt->brow[i] = (uint8_t)pt->n_bordered++;
      pt->bordered_mask |= 1ull << i;
    }
  }
  /* second half = base-wise complements of the first half (code ^ 2: same plane 0, inverted plane 1)? */
  pt->paired = p->n_patterns >= 2 && p->n_patterns % 2 == 0 && !getenv("TPS_K2_NO_PAIRS");
  const uint32_t half = p->n_patterns / 2;
  for (uint32_t i = 0; pt->paired && i < half; ++i) {
    const uint32_t k = pt->len[i], km = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    if (pt->len[i + half] != k || pt->lo[i + half] != pt->lo[i] || pt->hi[i + half] != ((~pt->hi[i]) & km)) pt->paired = 0;
  }
  return TPS_OK;
}

int log2_ceil(uint64_t v) {
  int b = 0;
  while (b < 64 && (1ull << b) < v) ++b;
  return b;
}

}  // namespace

extern "C" {

int tps_abi_version(void) { return TPS_ABI_VERSION; }

const char *tps_build_info(void) {
  return "topsicle_b200 sm_100a; kernels: tps_pack_tma_kernel, tps_pack_kernel, tps_trc_reg_kernel<K>, tps_trc_kernel<K>, tps_window_bp_kernel<K>, tps_window_kernel<K>, tps_changepoint_kernel; " __DATE__;
}

const char *tps_last_error(const tps_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

int tps_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void *tps_alloc_pinned(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void tps_free_pinned(void *p) {
  if (p) cudaFreeHost(p);
}

void tps_destroy(tps_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < 4; ++i) {
    Slot &s = ctx->slots[i];
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.d_bases); cudaFree(s.d_codes); cudaFree(s.d_flags);
    cudaFree(s.d_off); cudaFree(s.d_len); cudaFree(s.d_true_len); cudaFree(s.d_tails); cudaFree(s.d_rows); cudaFree(s.d_pass); cudaFree(s.d_counters);
    cudaFree(s.d_raw); cudaFree(s.d_cw); cudaFree(s.d_items); cudaFree(s.d_tile_done); cudaFree(s.d_gs_debug);
    if (s.h_rows) cudaFreeHost(s.h_rows);
    if (s.h_counters) cudaFreeHost(s.h_counters);
    if (s.done) cudaEventDestroy(s.done);
    if (s.e_in) cudaEventDestroy(s.e_in);
    if (s.e_k1) cudaEventDestroy(s.e_k1);
    if (s.e_tail) cudaEventDestroy(s.e_tail);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (ctx->pack_stream) { /* shared with the other contexts of this device: drained, then released */
    cudaStreamSynchronize(ctx->pack_stream);
    cudaStreamSynchronize(ctx->tail_stream);
    release_dev_streams(ctx->device);
    ctx->pack_stream = ctx->tail_stream = nullptr;
  }
  for (int r = 0; r < TPS_TIMING_RING; ++r)
    for (int i = 0; i < 4; ++i)
      if (ctx->ev[r][i]) cudaEventDestroy(ctx->ev[r][i]);
  delete ctx;
}

int tps_create(tps_ctx **out, int device, const tps_params *params) {
  if (!out || !params) return fail(nullptr, TPS_EINVAL, "null argument");
  *out = nullptr;
  if (params->struct_size != sizeof(tps_params))
    return fail(nullptr, TPS_EINVAL, "tps_params.struct_size %u != %zu (ABI mismatch)", params->struct_size,
                sizeof(tps_params));
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, TPS_ENODEVICE, "no CUDA device visible; topsicle_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(nullptr, TPS_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
  const tps_params &p = *params;
  if (p.window_size < 1 || p.window_size > 8192) return fail(nullptr, TPS_EINVAL, "window_size must be in 1..8192");
  if (p.slide < 1) return fail(nullptr, TPS_EINVAL, "slide must be >= 1");
  if (p.no_bp < 1 || p.no_bp > 32768) return fail(nullptr, TPS_EINVAL, "no_bp must be in 1..32768");
  if (p.n_slots < 1 || p.n_slots > 4) return fail(nullptr, TPS_EINVAL, "n_slots must be in 1..4");
  if (p.max_batch_reads < 1 || p.max_batch_bases < 1) return fail(nullptr, TPS_EINVAL, "batch capacities must be >= 1");
  if (p.flags & ~(TPS_FLAG_STEP1_ONLY | TPS_FLAG_FORCE_FORWARD | TPS_FLAG_FORCE_REVERSE))
    return fail(nullptr, TPS_EINVAL, "unknown bits in tps_params.flags");
  if ((p.flags & TPS_FLAG_FORCE_FORWARD) && (p.flags & TPS_FLAG_FORCE_REVERSE))
    return fail(nullptr, TPS_EINVAL, "FORCE_FORWARD and FORCE_REVERSE are exclusive");
  TpsPatTable pt;
  int rc = build_pattern_table(&p, &pt);
  if (rc) return rc;
  uint32_t kmin = 255;
  for (uint32_t i = 0; i < pt.n; ++i) kmin = pt.len[i] < kmin ? pt.len[i] : kmin;
  const uint32_t cnt_max = (p.window_size - 1) / kmin > 0 ? (p.window_size - 1) / kmin : 1;
  if (p.want_rawcount && cnt_max > 255)
    return fail(nullptr, TPS_EINVAL, "want_rawcount needs (window_size-1)/min(pattern_len) <= 255");
  /* windows of the longest possible region */
  const uint64_t reg_max = p.maxlengthtelo > p.trimfirst ? (uint64_t)p.maxlengthtelo - p.trimfirst : 0;
  const uint64_t nw_max = reg_max >= p.window_size ? (reg_max - p.window_size) / p.slide + 1 : 0;
  if (nw_max > 0xFFFFFFF0ull / 4) return fail(nullptr, TPS_EINVAL, "too many windows per read");
  /* exact change-point compare needs n^6 * cmax^2 < 2^128 (see tps_bitops.h) */
  if (6 * log2_ceil(nw_max ? nw_max : 1) + 2 * log2_ceil((uint64_t)cnt_max * pt.n) >= 128)
    return fail(nullptr, TPS_EINVAL, "maxlengthtelo/slide/windowSize combination exceeds the exact "
                                     "128-bit change-point range (%llu windows)", (unsigned long long)nw_max);

  tps_ctx *ctx = new (std::nothrow) tps_ctx();
  if (!ctx) return fail(nullptr, TPS_ENOMEM, "out of host memory");
  memset(ctx->ev, 0, sizeof(ctx->ev));
  ctx->device = device;
  ctx->p = p;
  ctx->pt = pt;
#define TPS_CC(call)                                                               \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      int c_ = fail(nullptr, e_ == cudaErrorMemoryAllocation ? TPS_ENOMEM : TPS_ECUDA, "%s: %s", #call, \
                    cudaGetErrorString(e_));                                       \
      tps_destroy(ctx);                                                            \
      return c_;                                                                   \
    }                                                                              \
  } while (0)
  TPS_CC(cudaSetDevice(device));
  cudaDeviceProp prop;
  TPS_CC(cudaGetDeviceProperties(&prop, device));
  ctx->n_sms = prop.multiProcessorCount;

  /* pick the K2/K3 instantiation: K = common literal length if all literals share one <= 8 */
  uint32_t kmax = 0;
  for (uint32_t i = 0; i < pt.n; ++i) kmax = pt.len[i] > kmax ? pt.len[i] : kmax;
  ctx->kt = (kmin == kmax && kmax <= 8) ? (int)kmax : 0;
  switch (ctx->kt) {
#define TPS_PICK(KK) case KK: ctx->k2_fn = tps_trc_kernel<KK>; ctx->k2r_fn = tps_trc_reg_kernel<KK>; ctx->k3_fn = tps_window_kernel<KK>; ctx->k3n_fn = tps_window_bp_kernel<KK>; break;
    TPS_PICK(1) TPS_PICK(2) TPS_PICK(3) TPS_PICK(4) TPS_PICK(5) TPS_PICK(6) TPS_PICK(7) TPS_PICK(8)
#undef TPS_PICK
    default: ctx->k2_fn = tps_trc_kernel<0>; ctx->k2r_fn = tps_trc_reg_kernel<0>; ctx->k3_fn = tps_window_kernel<0>; break;
  }
  const uint32_t pm_words = 2 * pt.n * (ctx->kt > 0 ? (uint32_t)ctx->kt : 1u);
  /* K2 geometry */
  ctx->k2_lin_words = lin_words_for(p.no_bp);
  ctx->k2_nq_max = (p.no_bp + 31) / 32;
  ctx->k2_smem = (pm_words + TPS_K2_WARPS * (3 * ctx->k2_lin_words + pt.n_bordered * ctx->k2_nq_max + TPS_MAX_PATTERNS)) * 4;
  ctx->k2_reg = p.no_bp <= 1000 && pt.n <= 32 && !getenv("TPS_K2_SMEM_PATH");
  ctx->k2r_smem = (pm_words + TPS_K2R_WARPS * pt.n_bordered * 32) * 4;
  {
    const char *e = getenv("TPS_K2_CONST"); /* 0 = always the table-driven K2 (A/B, tests) */
    const uint32_t U = pt.n / 2;
    if (ctx->k2_reg && pt.paired && pt.n_bordered == 0 && !(e && atoi(e) == 0)) {
#define TPS_PICK_U(KK, UU) case UU: ctx->k2c_fn = tps_trc_const_kernel<KK, UU>; break;
#define TPS_PICK_K(KK) case KK: switch (U) { TPS_PICK_U(KK, 5) TPS_PICK_U(KK, 6) TPS_PICK_U(KK, 7) TPS_PICK_U(KK, 8) default: break; } break;
      switch (ctx->kt) {
        TPS_PICK_K(3) TPS_PICK_K(4) TPS_PICK_K(5) TPS_PICK_K(6) TPS_PICK_K(7) TPS_PICK_K(8)
        default: break;
      }
#undef TPS_PICK_K
#undef TPS_PICK_U
    }
    if (ctx->k2c_fn) {
      memset(&ctx->k2c_masks, 0, sizeof(ctx->k2c_masks));
      for (uint32_t q = 0; q < U; ++q)
        for (int j = 0; j < ctx->kt; ++j) {
          ctx->k2c_masks.x[q][j] = 0u - ((pt.lo[q] >> j) & 1u);
          ctx->k2c_masks.y[q][j] = 0u - ((pt.hi[q] >> j) & 1u);
        }
    }
  }
  /* K3 geometry: a tile stages tile_bases + W positions; keep that <= 4096 (128 words) when W allows */
  const uint32_t w32 = (p.window_size + 31) / 32 * 32;
  ctx->k3_tile_bases = w32 + 1024 <= 4096 ? 4096 - w32 : 1024;
  const uint32_t tile_n = ctx->k3_tile_bases + p.window_size;
  ctx->k3_lin_words = lin_words_for(tile_n);
  ctx->k3_tile_words = (tile_n + 31) / 32 + 1;
  ctx->k3_tiles_max = (uint32_t)((reg_max + ctx->k3_tile_bases - 1) / ctx->k3_tile_bases);
  if (ctx->k3_tiles_max == 0) ctx->k3_tiles_max = 1;
  ctx->k3_smem = (pm_words + 3 * ctx->k3_lin_words + 3 * ctx->k3_tile_words + 3 + 2 * pt.n * ctx->k3_tile_words +
                  pt.n_bordered * ctx->k3_tile_words) * 4;
  {
    /* bit-parallel K3: needs distinct literals of one length (at most one literal matches at a position) */
    bool distinct = true;
    for (uint32_t i = 0; i < pt.n && distinct; ++i)
      for (uint32_t j = i + 1; j < pt.n && distinct; ++j)
        distinct = !(pt.len[i] == pt.len[j] && pt.lo[i] == pt.lo[j] && pt.hi[i] == pt.hi[j]);
    const char *e = getenv("TPS_K3_BITPAR"); /* 0 = tps_window_kernel + tps_changepoint_kernel (A/B, and the fallback) */
    ctx->k3_bitpar = ctx->kt > 0 && distinct && !p.want_rawcount && p.window_size >= (uint32_t)ctx->kt + 32u &&
                     p.window_size <= 2048u && 4ull * p.slide + p.window_size <= 32u * TPS_K3N_THREADS /* a tile holds a group */ &&
                     5ull * cnt_max * pt.n <= 65535u /* group sums as uint16 */ &&
                     nw_max / 5 * 2 <= 32768 /* ... of a whole read in shared memory */ && !(e && atoi(e) == 0);
    if (ctx->k3_bitpar) {
      const uint32_t D = p.window_size - (uint32_t)ctx->kt, NT = TPS_K3N_THREADS;
      /* a tile stages (windows - 1) * slide + W <= 32 NT positions; whole groups of five windows */
      ctx->k3n_tile_windows = (uint32_t)(((32ull * NT - p.window_size) / p.slide + 1) / 5 * 5);
      ctx->k3n_tile_bases = ctx->k3n_tile_windows * p.slide;
      ctx->k3n_lin_words = lin_words_for(32u * NT);
      ctx->k3n_stride = NT + (D >> 5) + 2u;
      ctx->k3n_tiles_max = (uint32_t)(((nw_max + 4) / 5 + ctx->k3n_tile_windows / 5 - 1) / (ctx->k3n_tile_windows / 5));
      if (ctx->k3n_tiles_max == 0) ctx->k3n_tiles_max = 1;
      ctx->k3n_nz = 1;
      while ((1u << ctx->k3n_nz) <= pt.n) ctx->k3n_nz++;
      ctx->k3n_gs_cap = (uint32_t)(((nw_max + 4) / 5 + 7) & ~7ull); /* group sums of the longest read, in shared memory */
      if (ctx->k3n_gs_cap == 0) ctx->k3n_gs_cap = 8;
      /* raw[2] | pm | lin | ori | alignment | Z | Zhi | UP | CP | SP | brows | gsum */
      ctx->k3n_smem = (2 * TPS_K3N_RAW_WORDS + pm_words + 3 * ctx->k3n_lin_words + 3 * (NT + 1) + 3 + 4 * (NT + 1) +
                       (ctx->k3n_nz > 4 ? 4 * (NT + 1) : 0) + 2 * (NT + 2) + (pt.n_bordered ? 2 * (NT + 2) : 0) +
                       2 * ((pt.n + 3u) & ~3u) * ctx->k3n_stride + pt.n_bordered * (NT + 1) + 4 + ctx->k3n_gs_cap / 2) * 4;
      const char *ng = getenv("TPS_K3_NO_GROUPS"); /* 1 = no five-window fast path (A/B) */
      ctx->k3n_no_groups = ng && atoi(ng) != 0;
      const char *dg = getenv("TPS_K3_DEBUG_GS");
      ctx->k3n_debug_gs = dg && atoi(dg) != 0;
      const char *sp = getenv("TPS_K3_SPLIT"); /* 0 = a read is always one CTA's work (A/B) */
      ctx->k3n_no_split = sp && atoi(sp) == 0;
    }
  }
  ctx->cw_stride = (uint32_t)(nw_max ? nw_max : 1);
  ctx->max_pass = p.max_pass_reads ? p.max_pass_reads : p.max_batch_reads;
  if (ctx->max_pass > p.max_batch_reads) ctx->max_pass = p.max_batch_reads;
  if ((uint64_t)ctx->max_pass * (ctx->k3_tiles_max > ctx->k3n_tiles_max ? ctx->k3_tiles_max : ctx->k3n_tiles_max) > 0xFFFFFFF0ull) {
    int c_ = fail(nullptr, TPS_EINVAL, "max_pass_reads * tiles per read overflows the work counter");
    tps_destroy(ctx);
    return c_;
  }
  const uint32_t smem_limit = (uint32_t)prop.sharedMemPerBlockOptin;
  if (ctx->k3_bitpar && ctx->k3n_smem > smem_limit) ctx->k3_bitpar = false; /* very many literals: the plain kernel */
  if (ctx->k2_smem > smem_limit || ctx->k3_smem > smem_limit) {
    int c_ = fail(nullptr, TPS_EINVAL, "parameters need %u / %u bytes of shared memory (limit %u)", ctx->k2_smem,
                  ctx->k3_smem, smem_limit);
    tps_destroy(ctx);
    return c_;
  }
  TPS_CC(cudaFuncSetAttribute(ctx->k2_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->k2_smem));
  TPS_CC(cudaFuncSetAttribute(ctx->k3_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->k3_smem));
  if (ctx->k2r_smem > 48 * 1024)
    TPS_CC(cudaFuncSetAttribute(ctx->k2r_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->k2r_smem));
  if (ctx->k3_bitpar) {
    int occn = 0;
    TPS_CC(cudaFuncSetAttribute(ctx->k3n_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->k3n_smem));
    TPS_CC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occn, ctx->k3n_fn, TPS_K3N_THREADS, ctx->k3n_smem));
    ctx->k3n_grid = (uint32_t)(ctx->n_sms * (occn > 0 ? occn : 1));
  }
  int occ1 = 0, occ3 = 0, occ4 = 0;
  {
    const char *e = getenv("TPS_K1_UNROLL"); /* tuning knob: 2, 4 (default) or 8 tiles in flight per warp */
    ctx->k1_unroll = e ? atoi(e) : 4;
    if (ctx->k1_unroll == 8) ctx->k1_fn = tps_pack_kernel<8>;
    else { ctx->k1_unroll = 4; ctx->k1_fn = tps_pack_kernel<4>; }
#ifdef TPS_TUNING
    if (const char *mb = getenv("TPS_K1_MINB")) { /* register-capped builds: more resident CTAs */
      ctx->k1_unroll = 4;
      switch (atoi(mb)) {
        case 5: ctx->k1_fn = tps_pack_kernel<4, 5>; break;
        case 6: ctx->k1_fn = tps_pack_kernel<4, 6>; break;
        case 8: ctx->k1_fn = tps_pack_kernel<4, 8>; break;
        default: break;
      }
    }
#endif
#ifdef TPS_TUNING
    const char *pv = getenv("TPS_K1_PROBE");
    if (pv && atoi(pv) == 1) { ctx->k1_unroll = 4; ctx->k1_fn = tps_pack_probe<4, 1>; }
    if (pv && atoi(pv) == 2) { ctx->k1_unroll = 4; ctx->k1_fn = tps_pack_probe<4, 2>; }
    if (pv && atoi(pv) == 3) { ctx->k1_unroll = 8; ctx->k1_fn = tps_pack_probe<8, 2>; }
    if (pv && atoi(pv) == 4) { ctx->k1_unroll = 4; ctx->k1_fn = tps_pack_probe<4, 4>; }
    if (pv && atoi(pv) == 5) { ctx->k1_unroll = 4; ctx->k1_fn = tps_pack_probe<4, 5>; }
#endif
  }
  TPS_CC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, ctx->k1_fn, TPS_K1_THREADS, 0));
  TPS_CC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, ctx->k3_fn, TPS_K3_THREADS, ctx->k3_smem));
  TPS_CC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, tps_changepoint_kernel, TPS_K4_THREADS, 0));
  ctx->k1_grid = ctx->n_sms * (occ1 > 0 ? occ1 : 1);
  if (const char *e = getenv("TPS_K1_CTAS_PER_SM")) { /* tuning knob: leave room for other streams' kernels */
    int v = atoi(e);
    if (v >= 1 && v <= occ1) ctx->k1_grid = ctx->n_sms * v;
  }
  {
    /* K1 variant: bulk-copy (TMA) staged, TPS_K1_STAGES x 16 KiB per CTA (default), or register-staged
     * (TPS_K1_TMA=0, kept for A/B measurements) */
    const char *e = getenv("TPS_K1_TMA");
    ctx->k1_tma = !(e && atoi(e) == 0);
    if (ctx->k1_tma) {
      const char *se = getenv("TPS_K1_STAGES");
      int stages = se ? atoi(se) : 2;
      if (stages < 2) stages = 2;
      if (stages > 12) stages = 12;
      ctx->k1t_stages = (uint32_t)stages;
      if (const char *ke = getenv("TPS_K1_STAGE_KB")) ctx->k1t_unroll = atoi(ke) <= 8 ? 2u : 4u;
      ctx->k1t_smem = ctx->k1t_stages * (TPS_K1T_STAGE_BYTES(ctx->k1t_unroll) + 16u);
      int want_per_sm = 3; /* 3 CTAs x 2 stages x 16 KiB = 96 KiB in flight per SM measured best (profiles/README.md) */
      if (const char *ce = getenv("TPS_K1_CTAS_PER_SM")) want_per_sm = atoi(ce);
      if (want_per_sm < 1) want_per_sm = 1;
      if (want_per_sm > 6) want_per_sm = 6;
      ctx->k1t_fn = ctx->k1t_unroll == 2 ? tps_pack_tma_kernel<2> : tps_pack_tma_kernel<4>;
      TPS_CC(cudaFuncSetAttribute(ctx->k1t_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->k1t_smem));
      int occt = 0;
      TPS_CC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occt, ctx->k1t_fn, TPS_K1T_THREADS, ctx->k1t_smem));
      if (occt < 1) occt = 1;
      const int per_sm = want_per_sm < occt ? want_per_sm : occt;
      ctx->k1t_grid = ctx->n_sms * per_sm;
    }
  }
  ctx->k3_grid = (uint32_t)(ctx->n_sms * (occ3 > 0 ? occ3 : 1));
  ctx->k4_grid = (uint32_t)(ctx->n_sms * (occ4 > 0 ? (occ4 > 4 ? 4 : occ4) : 1));

  {
    const char *e = getenv("TPS_SPLIT_STREAMS"); /* 0 = everything of a batch on its slot's stream */
    ctx->split = !(e && atoi(e) == 0);
    if (ctx->split && acquire_dev_streams(device, &ctx->pack_stream, &ctx->tail_stream) != 0) {
      int c_ = fail(nullptr, TPS_ECUDA, "cannot create the pack / tail streams: %s", cudaGetErrorString(cudaGetLastError()));
      tps_destroy(ctx);
      return c_;
    }
  }
  ctx->cap_tiles = (p.max_batch_bases + 511) / 512;
  const uint64_t cap_pad = ((p.max_batch_bases + 2047) / 2048) * 2048;
  for (uint32_t i = 0; i < p.n_slots; ++i) {
    Slot &s = ctx->slots[i];
    TPS_CC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    TPS_CC(cudaEventCreateWithFlags(&s.e_in, cudaEventDisableTiming));
    TPS_CC(cudaEventCreateWithFlags(&s.e_k1, cudaEventDisableTiming));
    TPS_CC(cudaEventCreateWithFlags(&s.e_tail, cudaEventDisableTiming));
    TPS_CC(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    TPS_CC(cudaMalloc(&s.d_bases, cap_pad));
    TPS_CC(cudaMemset(s.d_bases, 'N', cap_pad));
    /* + 64 bytes: the bit-parallel K3 prefetches code / flag words in aligned 16-byte chunks */
    TPS_CC(cudaMalloc(&s.d_codes, ctx->cap_tiles * 32 * sizeof(uint32_t) + 64));
    TPS_CC(cudaMalloc(&s.d_flags, ctx->cap_tiles * sizeof(uint32_t) + 64));
    TPS_CC(cudaMemset(s.d_flags, 0, ctx->cap_tiles * sizeof(uint32_t) + 64));
    TPS_CC(cudaMalloc(&s.d_off, ((uint64_t)p.max_batch_reads + 1) * sizeof(uint64_t)));
    TPS_CC(cudaMalloc(&s.d_len, (uint64_t)p.max_batch_reads * sizeof(uint32_t)));
    TPS_CC(cudaMalloc(&s.d_true_len, (uint64_t)p.max_batch_reads * sizeof(uint32_t)));
    TPS_CC(cudaMalloc(&s.d_tails, (uint64_t)p.max_batch_reads));
    TPS_CC(cudaMalloc(&s.d_rows, (uint64_t)p.max_batch_reads * sizeof(tps_row)));
    TPS_CC(cudaMalloc(&s.d_pass, (uint64_t)ctx->max_pass * sizeof(uint32_t)));
    TPS_CC(cudaMalloc(&s.d_counters, 8 * sizeof(uint32_t)));
    if (p.want_rawcount) TPS_CC(cudaMalloc(&s.d_raw, p.rawcount_capacity ? p.rawcount_capacity : 1));
    if (ctx->k3_bitpar) { /* c_w never leaves the SM: group sums in shared memory */
      TPS_CC(cudaMalloc(&s.d_items, (uint64_t)ctx->max_pass * sizeof(TpsReadItem)));
      TPS_CC(cudaMalloc(&s.d_tile_done, (uint64_t)ctx->max_pass * sizeof(uint32_t)));
      /* group-sum rows for the split mode (few passing reads); as uint16 a fifth of a tenth of the plain kernel's c_w */
      TPS_CC(cudaMalloc(&s.d_cw, (uint64_t)ctx->max_pass * ctx->k3n_gs_cap * sizeof(uint16_t)));
      if (ctx->k3n_debug_gs) TPS_CC(cudaMalloc(&s.d_gs_debug, (uint64_t)ctx->max_pass * ctx->k3n_gs_cap * sizeof(uint16_t)));
    } else {
      TPS_CC(cudaMalloc(&s.d_cw, (uint64_t)ctx->max_pass * ctx->cw_stride * sizeof(uint32_t)));
    }
    TPS_CC(cudaHostAlloc(&s.h_rows, (uint64_t)p.max_batch_reads * sizeof(tps_row), cudaHostAllocDefault));
    TPS_CC(cudaHostAlloc(&s.h_counters, 8 * sizeof(uint32_t), cudaHostAllocDefault));
  }
  for (int r = 0; r < TPS_TIMING_RING; ++r)
    for (int i = 0; i < 4; ++i) TPS_CC(cudaEventCreate(&ctx->ev[r][i]));
#undef TPS_CC
  *out = ctx;
  return TPS_OK;
}

}  // extern "C"

namespace {

/* What enqueue_scan works on.  d_bases / d_off may be slot- or caller-owned.  With `packed_by` the batch was
 * already packed (K1) into that slot's code / flag buffers by another context: only K2..K4 run, reading them,
 * on the streams of `stream_owner`. */
struct ScanJob {
  const uint8_t *d_bases = nullptr;
  const uint64_t *d_off = nullptr;
  uint32_t n_reads = 0;
  uint64_t n_bases = 0;
  tps_row *d_rows = nullptr;
  bool timed = false;                  /* record the CUDA-event ring (device-resident scans) */
  const Slot *packed_by = nullptr;
  const uint32_t *d_len = nullptr;     /* span batch: read lengths */
  tps_ctx *stream_owner = nullptr;
  const uint32_t *d_true_len = nullptr; /* ends batch: real read lengths (step 1 only) */
  const uint8_t *d_tails = nullptr;     /* region batch: per-read forced tail */
};

/* Enqueue K1..K4 for one batch; `st` is the slot's stream (copies, ordering). */
int enqueue_scan(tps_ctx *ctx, Slot &s, cudaStream_t st, const ScanJob &job) {
  const uint8_t *d_bases = job.d_bases;
  const uint64_t *d_off = job.d_off;
  const uint32_t n_reads = job.n_reads;
  const uint64_t n_bases = job.n_bases;
  tps_row *d_rows = job.d_rows;
  const bool timed = job.timed;
  const Slot *packed_by = job.packed_by;
  const uint32_t *d_len = job.d_len;
  tps_ctx *stream_owner = job.stream_owner;
  const uint32_t *d_true_len = job.d_true_len;
  const uint8_t *d_tails = job.d_tails;
  const tps_params &p = ctx->p;
  cudaEvent_t *ev = ctx->ev[ctx->scan_seq % TPS_TIMING_RING];
  /* split mode: `st` (the slot's stream) orders the batch against its copies; K1 goes to the context's
   * pack stream, K2..K4 to its tail stream.  A follower context (packed_by) runs on the owner's streams. */
  tps_ctx *so = stream_owner ? stream_owner : ctx;
  const bool split = so->split;
  cudaStream_t sp = split ? so->pack_stream : st; /* K1 */
  cudaStream_t sh = split ? so->tail_stream : st; /* K2..K4 */
  if (split) {
    TPS_CUDA(ctx, cudaEventRecord(s.e_in, st));
    TPS_CUDA(ctx, cudaStreamWaitEvent(packed_by ? sh : sp, s.e_in, 0));
  } else {
    TPS_CUDA(ctx, cudaMemsetAsync(s.d_counters, 0, 8 * sizeof(uint32_t), st));
  }
  if (timed) TPS_CUDA(ctx, cudaEventRecord(ev[0], sp));
  const uint64_t n_tiles = (n_bases + 511) / 512;
  if (n_tiles && !packed_by && ctx->k1_tma) {
    const uint64_t st_tiles = TPS_K1T_STAGE_TILES(ctx->k1t_unroll);
    const uint64_t want = (n_tiles + st_tiles - 1) / st_tiles;
    int grid = (int)(want < (uint64_t)ctx->k1t_grid ? want : (uint64_t)ctx->k1t_grid);
    ctx->k1t_fn<<<grid, TPS_K1T_THREADS, ctx->k1t_smem, sp>>>(
        reinterpret_cast<const uint4 *>(d_bases), s.d_codes, s.d_flags, n_tiles, ctx->k1t_stages);
    ctx->launches++;
  } else if (n_tiles && !packed_by) {
    const uint64_t per_cta = (uint64_t)(TPS_K1_THREADS / 32) * ctx->k1_unroll;
    uint64_t want = (n_tiles + per_cta - 1) / per_cta;
    int grid = (int)(want < (uint64_t)ctx->k1_grid ? want : (uint64_t)ctx->k1_grid);
    ctx->k1_fn<<<grid, TPS_K1_THREADS, 0, sp>>>(reinterpret_cast<const uint4 *>(d_bases), s.d_codes, s.d_flags,
                                                     n_tiles);
    ctx->launches++;
  }
  if (timed) TPS_CUDA(ctx, cudaEventRecord(ev[1], sp));
  if (split) {
    if (!packed_by) {
      TPS_CUDA(ctx, cudaEventRecord(s.e_k1, sp));
      TPS_CUDA(ctx, cudaStreamWaitEvent(sh, s.e_k1, 0));
    }
    TPS_CUDA(ctx, cudaMemsetAsync(s.d_counters, 0, 8 * sizeof(uint32_t), sh));
  }
  TpsScanArgs a;
  memset(&a, 0, sizeof(a));
  const Slot &src = packed_by ? *packed_by : s;
  a.pk.codes = src.d_codes;
  a.pk.flags = src.d_flags;
  a.pk.bases = d_bases;
  s.last_bases = d_bases;
  a.offsets = d_off;
  a.lens = d_len;
  a.true_lens = d_true_len; /* ends batch: step 1 only, the windows need bases that were not uploaded */
  a.force_tails = d_tails;
  a.n_reads = n_reads;
  a.rows = d_rows;
  a.pass_list = s.d_pass;
  a.counters = s.d_counters;
  a.min_seq_length = p.min_seq_length;
  a.no_bp = p.no_bp;
  a.count_threshold = p.count_threshold;
  a.W = p.window_size;
  a.slide = p.slide;
  a.trimfirst = p.trimfirst;
  a.maxlengthtelo = p.maxlengthtelo;
  a.want_rawcount = p.want_rawcount;
  a.flags = p.flags | (d_true_len ? TPS_FLAG_STEP1_ONLY : 0u);
  a.raw = s.d_raw;
  a.raw_capacity = p.want_rawcount ? p.rawcount_capacity : 0;
  a.cw = s.d_cw;
  a.cw_stride = ctx->cw_stride;
  a.max_pass = ctx->max_pass;
  a.tile_bases = ctx->k3_tile_bases;
  a.tiles_max = ctx->k3_tiles_max;
  a.nq_max = ctx->k2_nq_max;
  a.items = ctx->k3_bitpar ? s.d_items : nullptr;
  a.nz = ctx->k3n_nz;
  a.bp_tile_windows = ctx->k3n_tile_windows;
  a.gs_cap = ctx->k3n_gs_cap;
  a.no_groups = ctx->k3n_no_groups;
  a.gs_debug = ctx->k3_bitpar && ctx->k3n_debug_gs ? s.d_gs_debug : nullptr;
  a.gs_rows = ctx->k3_bitpar ? reinterpret_cast<uint16_t *>(s.d_cw) : nullptr;
  a.tile_done = s.d_tile_done;
  a.no_split = ctx->k3n_no_split;
  if (n_reads) {
    a.lin_words = ctx->k2_lin_words;
    if (ctx->k2c_fn)
      ctx->k2c_fn<<<(n_reads + TPS_K2R_WARPS - 1) / TPS_K2R_WARPS, TPS_K2R_WARPS * 32, 0, sh>>>(a, ctx->k2c_masks);
    else if (ctx->k2_reg)
      ctx->k2r_fn<<<(n_reads + TPS_K2R_WARPS - 1) / TPS_K2R_WARPS, TPS_K2R_WARPS * 32, ctx->k2r_smem, sh>>>(a, ctx->pt);
    else
      ctx->k2_fn<<<(n_reads + TPS_K2_WARPS - 1) / TPS_K2_WARPS, TPS_K2_WARPS * 32, ctx->k2_smem, sh>>>(a, ctx->pt);
    ctx->launches++;
  }
  if (timed) TPS_CUDA(ctx, cudaEventRecord(ev[2], sh));
  if (n_reads && !(a.flags & TPS_FLAG_STEP1_ONLY)) {
    if (ctx->k3_bitpar) { /* window counts and change points in one launch */
      a.lin_words = ctx->k3n_lin_words;
      a.tile_words = ctx->k3n_stride;
      ctx->k3n_fn<<<ctx->k3n_grid, TPS_K3N_THREADS, ctx->k3n_smem, sh>>>(a, ctx->pt);
      ctx->launches += 1;
    } else {
      a.lin_words = ctx->k3_lin_words;
      a.tile_words = ctx->k3_tile_words;
      ctx->k3_fn<<<ctx->k3_grid, TPS_K3_THREADS, ctx->k3_smem, sh>>>(a, ctx->pt);
      tps_changepoint_kernel<<<ctx->k4_grid, TPS_K4_THREADS, 0, sh>>>(a);
      ctx->launches += 2;
    }
  }
  if (timed) {
    TPS_CUDA(ctx, cudaEventRecord(ev[3], sh));
    ctx->scan_seq++;
  }
  if (split) { /* the slot's stream continues (D2H, next batch) only after the tail kernels */
    TPS_CUDA(ctx, cudaEventRecord(s.e_tail, sh));
    TPS_CUDA(ctx, cudaStreamWaitEvent(st, s.e_tail, 0));
  }
  TPS_CUDA(ctx, cudaGetLastError());
  return TPS_OK;
}

}  // namespace

extern "C" {

static int submit_common(tps_ctx *ctx, const uint8_t *bases, uint64_t n_bases, const uint64_t *starts,
                         const uint32_t *lengths, uint32_t n_reads, uint64_t batch_id,
                         const uint32_t *true_lens = nullptr, const uint8_t *tails = nullptr) {
  const tps_params &p = ctx->p;
  if (n_reads > p.max_batch_reads) return fail(ctx, TPS_ECAPACITY, "batch has %u reads, capacity %u", n_reads, p.max_batch_reads);
  if (n_bases > p.max_batch_bases) return fail(ctx, TPS_ECAPACITY, "batch has %llu bases, capacity %llu",
                                               (unsigned long long)n_bases, (unsigned long long)p.max_batch_bases);
  Slot *sl = nullptr;
  for (uint32_t i = 0; i < p.n_slots; ++i) {
    if (ctx->slots[i].busy && ctx->slots[i].batch_id == batch_id) return fail(ctx, TPS_ESTATE, "batch id already in flight");
    if (!ctx->slots[i].busy && !sl) sl = &ctx->slots[i];
  }
  if (!sl) return fail(ctx, TPS_ESTATE, "all %u slots busy: call tps_wait first", p.n_slots);
  TPS_CUDA(ctx, cudaSetDevice(ctx->device));
  Slot &s = *sl;
  cudaStream_t st = s.stream;
  /* span batches carry n starts + n lengths, back-to-back batches n+1 offsets */
  TPS_CUDA(ctx, cudaMemcpyAsync(s.d_off, starts, ((uint64_t)n_reads + (lengths ? 0 : 1)) * sizeof(uint64_t),
                                cudaMemcpyHostToDevice, st));
  if (lengths && n_reads)
    TPS_CUDA(ctx, cudaMemcpyAsync(s.d_len, lengths, (uint64_t)n_reads * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  if (true_lens && n_reads)
    TPS_CUDA(ctx, cudaMemcpyAsync(s.d_true_len, true_lens, (uint64_t)n_reads * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  if (tails && n_reads) TPS_CUDA(ctx, cudaMemcpyAsync(s.d_tails, tails, n_reads, cudaMemcpyHostToDevice, st));
  if (n_bases) TPS_CUDA(ctx, cudaMemcpyAsync(s.d_bases, bases, n_bases, cudaMemcpyHostToDevice, st));
  s.has_lens = lengths != nullptr;
  s.ends = true_lens != nullptr;
  s.regions = tails != nullptr;
  ScanJob job;
  job.d_bases = s.d_bases;
  job.d_off = s.d_off;
  job.n_reads = n_reads;
  job.n_bases = n_bases;
  job.d_rows = s.d_rows;
  job.d_len = lengths ? s.d_len : nullptr;
  job.d_true_len = true_lens ? s.d_true_len : nullptr;
  job.d_tails = tails ? s.d_tails : nullptr;
  int rc = enqueue_scan(ctx, s, st, job);
  if (rc) return rc;
  if (n_reads)
    TPS_CUDA(ctx, cudaMemcpyAsync(s.h_rows, s.d_rows, (uint64_t)n_reads * sizeof(tps_row), cudaMemcpyDeviceToHost, st));
  TPS_CUDA(ctx, cudaMemcpyAsync(s.h_counters, s.d_counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TPS_CUDA(ctx, cudaEventRecord(s.done, st));
  s.busy = true;
  s.batch_id = batch_id;
  s.n_reads = n_reads;
  return TPS_OK;
}

int tps_submit(tps_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads, uint64_t batch_id) {
  if (!ctx || !offsets || (!bases && n_reads)) return fail(ctx, TPS_EINVAL, "null argument");
  if (offsets[0] != 0) return fail(ctx, TPS_EINVAL, "offsets[0] must be 0");
  return submit_common(ctx, bases, offsets[n_reads], offsets, nullptr, n_reads, batch_id);
}

int tps_submit_spans(tps_ctx *ctx, const uint8_t *bases, uint64_t n_span, const uint64_t *starts,
                     const uint32_t *lengths, uint32_t n_reads, uint64_t batch_id) {
  if (!ctx || (n_reads && (!starts || !lengths)) || (!bases && n_span)) return fail(ctx, TPS_EINVAL, "null argument");
  if (n_reads && starts[n_reads - 1] + lengths[n_reads - 1] > n_span)
    return fail(ctx, TPS_EINVAL, "last read ends at %llu, beyond the %llu uploaded bytes",
                (unsigned long long)(starts[n_reads - 1] + lengths[n_reads - 1]), (unsigned long long)n_span);
  static const uint64_t zero = 0;
  return submit_common(ctx, bases, n_span, n_reads ? starts : &zero, n_reads ? lengths : nullptr, n_reads, batch_id);
}

int tps_submit_ends(tps_ctx *ctx, const uint8_t *bases, uint64_t n_span, const uint64_t *starts,
                    const uint32_t *lengths, const uint32_t *true_lengths, uint32_t n_reads, uint64_t batch_id) {
  if (!ctx || (n_reads && (!starts || !lengths || !true_lengths)) || (!bases && n_span))
    return fail(ctx, TPS_EINVAL, "null argument");
  if (n_reads && starts[n_reads - 1] + lengths[n_reads - 1] > n_span)
    return fail(ctx, TPS_EINVAL, "last read ends beyond the %llu uploaded bytes", (unsigned long long)n_span);
  static const uint64_t zero = 0;
  static const uint32_t zero32 = 0;
  return submit_common(ctx, bases, n_span, n_reads ? starts : &zero, n_reads ? lengths : nullptr, n_reads, batch_id,
                       n_reads ? true_lengths : &zero32, nullptr);
}

int tps_submit_regions(tps_ctx *ctx, const uint8_t *bases, uint64_t n_span, const uint64_t *starts,
                       const uint32_t *lengths, const uint8_t *tails, uint32_t n_reads, uint64_t batch_id) {
  if (!ctx || (n_reads && (!starts || !lengths || !tails)) || (!bases && n_span))
    return fail(ctx, TPS_EINVAL, "null argument");
  if (ctx->p.flags & TPS_FLAG_STEP1_ONLY) return fail(ctx, TPS_EINVAL, "region batches need a context that runs steps 2/3");
  if (n_reads && starts[n_reads - 1] + lengths[n_reads - 1] > n_span)
    return fail(ctx, TPS_EINVAL, "last read ends beyond the %llu uploaded bytes", (unsigned long long)n_span);
  for (uint32_t i = 0; i < n_reads; ++i)
    if (tails[i] > TPS_TAIL_REVERSE) return fail(ctx, TPS_EINVAL, "tails[%u] = %u is not a TPS_TAIL_* value", i, tails[i]);
  static const uint64_t zero = 0;
  static const uint8_t zero8 = 0;
  return submit_common(ctx, bases, n_span, n_reads ? starts : &zero, n_reads ? lengths : nullptr, n_reads, batch_id,
                       nullptr, n_reads ? tails : &zero8);
}

int tps_submit_shared(tps_ctx *ctx, tps_ctx *owner, uint64_t batch_id) {
  if (!ctx || !owner || ctx == owner) return fail(ctx, TPS_EINVAL, "tps_submit_shared needs two distinct contexts");
  if (ctx->device != owner->device) return fail(ctx, TPS_EINVAL, "contexts are on different devices");
  Slot *so = nullptr, *sl = nullptr;
  for (uint32_t i = 0; i < owner->p.n_slots; ++i)
    if (owner->slots[i].busy && owner->slots[i].batch_id == batch_id) so = &owner->slots[i];
  if (!so) return fail(ctx, TPS_ESTATE, "batch %llu is not in flight in the owner context", (unsigned long long)batch_id);
  const tps_params &p = ctx->p;
  if (so->n_reads > p.max_batch_reads)
    return fail(ctx, TPS_ECAPACITY, "batch has %u reads, capacity %u", so->n_reads, p.max_batch_reads);
  for (uint32_t i = 0; i < p.n_slots; ++i) {
    if (ctx->slots[i].busy && ctx->slots[i].batch_id == batch_id) return fail(ctx, TPS_ESTATE, "batch id already in flight");
    if (!ctx->slots[i].busy && !sl) sl = &ctx->slots[i];
  }
  if (!sl) return fail(ctx, TPS_ESTATE, "all %u slots busy: call tps_wait first", p.n_slots);
  TPS_CUDA(ctx, cudaSetDevice(ctx->device));
  Slot &s = *sl;
  cudaStream_t st = so->stream; /* owner's stream: ordered after its H2D + K1 and before its slot is reused */
  if (so->regions) return fail(ctx, TPS_EINVAL, "a region batch belongs to one context (its reads were chosen by that context's step 1)");
  s.ends = so->ends;
  s.regions = false;
  ScanJob job;
  job.d_bases = so->d_bases;
  job.d_off = so->d_off;
  job.n_reads = so->n_reads;
  job.d_rows = s.d_rows;
  job.packed_by = so;
  job.d_len = so->has_lens ? so->d_len : nullptr;
  job.stream_owner = owner;
  job.d_true_len = so->ends ? so->d_true_len : nullptr;
  int rc = enqueue_scan(ctx, s, st, job);
  if (rc) return rc;
  if (so->n_reads)
    TPS_CUDA(ctx, cudaMemcpyAsync(s.h_rows, s.d_rows, (uint64_t)so->n_reads * sizeof(tps_row), cudaMemcpyDeviceToHost, st));
  TPS_CUDA(ctx, cudaMemcpyAsync(s.h_counters, s.d_counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TPS_CUDA(ctx, cudaEventRecord(s.done, st));
  s.busy = true;
  s.batch_id = batch_id;
  s.n_reads = so->n_reads;
  return TPS_OK;
}

int tps_wait(tps_ctx *ctx, uint64_t batch_id, tps_row *rows_out, uint32_t *n_pass_out, uint8_t *rawcounts_out,
             uint64_t rawcount_cap, uint64_t *rawcount_elems) {
  if (!ctx) return TPS_EINVAL;
  Slot *sl = nullptr;
  for (uint32_t i = 0; i < ctx->p.n_slots; ++i)
    if (ctx->slots[i].busy && ctx->slots[i].batch_id == batch_id) sl = &ctx->slots[i];
  if (!sl) return fail(ctx, TPS_ESTATE, "batch %llu is not in flight", (unsigned long long)batch_id);
  Slot &s = *sl;
  TPS_CUDA(ctx, cudaSetDevice(ctx->device));
  TPS_CUDA(ctx, cudaEventSynchronize(s.done));
  if (rows_out && s.n_reads) memcpy(rows_out, s.h_rows, (uint64_t)s.n_reads * sizeof(tps_row));
  if (n_pass_out) *n_pass_out = s.h_counters[0];
  if (s.h_counters[4] & TPS_OVF_PASS) {
    s.busy = false;
    return fail(ctx, TPS_ECAPACITY, "%u reads passed the TRC cutoff but max_pass_reads is %u", s.h_counters[0],
                ctx->max_pass);
  }
  uint64_t elems = 0;
  memcpy(&elems, s.h_counters + 2, sizeof(uint64_t));
  if (rawcount_elems) *rawcount_elems = elems;
  if (ctx->p.want_rawcount) {
    if (s.h_counters[4] & TPS_OVF_RAWCOUNT) {
      s.busy = false;
      return fail(ctx, TPS_ECAPACITY, "rawcount_capacity %llu too small: batch needs %llu elements",
                  (unsigned long long)ctx->p.rawcount_capacity, (unsigned long long)elems);
    }
    if (rawcounts_out) {
      if (elems > rawcount_cap)
        return fail(ctx, TPS_ECAPACITY, "rawcounts_out holds %llu elements, batch has %llu (slot kept; call again)",
                    (unsigned long long)rawcount_cap, (unsigned long long)elems);
      if (elems) TPS_CUDA(ctx, cudaMemcpy(rawcounts_out, s.d_raw, elems, cudaMemcpyDeviceToHost));
    }
  }
  s.busy = false;
  return TPS_OK;
}

int tps_batch_info(tps_ctx *ctx, uint64_t batch_id, uint32_t *n_pass_out, uint64_t *rawcount_elems) {
  if (!ctx) return TPS_EINVAL;
  Slot *sl = nullptr;
  for (uint32_t i = 0; i < ctx->p.n_slots; ++i)
    if (ctx->slots[i].busy && ctx->slots[i].batch_id == batch_id) sl = &ctx->slots[i];
  if (!sl) return fail(ctx, TPS_ESTATE, "batch %llu is not in flight", (unsigned long long)batch_id);
  TPS_CUDA(ctx, cudaSetDevice(ctx->device));
  TPS_CUDA(ctx, cudaEventSynchronize(sl->done));
  if (n_pass_out) *n_pass_out = sl->h_counters[0];
  if (rawcount_elems) memcpy(rawcount_elems, sl->h_counters + 2, sizeof(uint64_t));
  return TPS_OK;
}

int tps_scan_device(tps_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t n_reads,
                    uint64_t n_bases, tps_row *d_rows_out) {
  return tps_scan_device_slot(ctx, 0, d_bases, d_offsets, n_reads, n_bases, d_rows_out);
}

int tps_scan_device_slot(tps_ctx *ctx, uint32_t slot, const uint8_t *d_bases, const uint64_t *d_offsets,
                         uint32_t n_reads, uint64_t n_bases, tps_row *d_rows_out) {
  if (!ctx || !d_offsets || !d_rows_out || (!d_bases && n_bases)) return fail(ctx, TPS_EINVAL, "null argument");
  if (slot >= ctx->p.n_slots) return fail(ctx, TPS_EINVAL, "slot %u out of range (context has %u)", slot, ctx->p.n_slots);
  if (n_reads > ctx->p.max_batch_reads || n_bases > ctx->p.max_batch_bases)
    return fail(ctx, TPS_ECAPACITY, "batch (%u reads, %llu bases) exceeds context capacity", n_reads,
                (unsigned long long)n_bases);
  if ((uintptr_t)d_bases & 15u) return fail(ctx, TPS_EINVAL, "d_bases must be 16-byte aligned");
  TPS_CUDA(ctx, cudaSetDevice(ctx->device));
  Slot &s = ctx->slots[slot];
  if (s.busy) return fail(ctx, TPS_ESTATE, "slot %u busy with a submitted batch", slot);
  ScanJob job;
  job.d_bases = d_bases;
  job.d_off = d_offsets;
  job.n_reads = n_reads;
  job.n_bases = n_bases;
  job.d_rows = d_rows_out;
  job.timed = true;
  return enqueue_scan(ctx, s, s.stream, job);
}

int tps_sync(tps_ctx *ctx) {
  if (!ctx) return TPS_EINVAL;
  TPS_CUDA(ctx, cudaSetDevice(ctx->device));
  for (uint32_t i = 0; i < ctx->p.n_slots; ++i) TPS_CUDA(ctx, cudaStreamSynchronize(ctx->slots[i].stream));
  return TPS_OK;
}

int tps_get_timings(tps_ctx *ctx, uint32_t back, float ms[TPS_N_TIMINGS]) {
  if (!ctx || !ms) return TPS_EINVAL;
  if (back >= TPS_TIMING_RING || back >= ctx->scan_seq)
    return fail(ctx, TPS_ESTATE, "timed scan %u steps back is not recorded (ring of %d)", back, TPS_TIMING_RING);
  cudaEvent_t *ev = ctx->ev[(ctx->scan_seq - 1 - back) % TPS_TIMING_RING];
  TPS_CUDA(ctx, cudaEventSynchronize(ev[3]));
  for (int i = 0; i < 3; ++i) TPS_CUDA(ctx, cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  TPS_CUDA(ctx, cudaEventElapsedTime(&ms[3], ev[0], ev[3]));
  return TPS_OK;
}

int tps_get_timeline(tps_ctx *ctx, uint32_t back, uint32_t base_back, float ms[TPS_N_TIMINGS]) {
  if (!ctx || !ms) return TPS_EINVAL;
  if (back >= TPS_TIMING_RING || back >= ctx->scan_seq || base_back >= TPS_TIMING_RING || base_back >= ctx->scan_seq)
    return fail(ctx, TPS_ESTATE, "timed scan %u / %u steps back is not recorded (ring of %d)", back, base_back,
                TPS_TIMING_RING);
  cudaEvent_t *ev = ctx->ev[(ctx->scan_seq - 1 - back) % TPS_TIMING_RING];
  cudaEvent_t *e0 = ctx->ev[(ctx->scan_seq - 1 - base_back) % TPS_TIMING_RING];
  TPS_CUDA(ctx, cudaEventSynchronize(ev[3]));
  TPS_CUDA(ctx, cudaEventSynchronize(e0[3]));
  for (int i = 0; i < 4; ++i) TPS_CUDA(ctx, cudaEventElapsedTime(&ms[i], e0[0], ev[i]));
  return TPS_OK;
}

int tps_elapsed_between(tps_ctx *from, uint32_t from_back, uint32_t from_event, tps_ctx *to, uint32_t to_back,
                        uint32_t to_event, float *ms) {
  if (!from || !to || !ms || from_event > 3 || to_event > 3) return TPS_EINVAL;
  if (from->device != to->device) return fail(to, TPS_EINVAL, "the two scans ran on different devices");
  if (from_back >= TPS_TIMING_RING || from_back >= from->scan_seq || to_back >= TPS_TIMING_RING || to_back >= to->scan_seq)
    return fail(to, TPS_ESTATE, "timed scan %u / %u steps back is not recorded (ring of %d)", from_back, to_back,
                TPS_TIMING_RING);
  cudaEvent_t *e0 = from->ev[(from->scan_seq - 1 - from_back) % TPS_TIMING_RING];
  cudaEvent_t *e1 = to->ev[(to->scan_seq - 1 - to_back) % TPS_TIMING_RING];
  TPS_CUDA(to, cudaSetDevice(to->device));
  TPS_CUDA(to, cudaEventSynchronize(e0[3]));
  TPS_CUDA(to, cudaEventSynchronize(e1[3]));
  TPS_CUDA(to, cudaEventElapsedTime(ms, e0[from_event], e1[to_event]));
  return TPS_OK;
}

int tps_follow_scan(int device, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads, const char *patterns,
                    uint32_t n_patterns, uint32_t k, uint32_t match_len, uint32_t min_seq_length, uint32_t skip,
                    uint32_t upto, uint32_t *sel_out, uint64_t sel_words) {
  if (!offsets || !patterns || !sel_out || (!bases && n_reads)) return fail(nullptr, TPS_EINVAL, "null argument");
  if (n_patterns < 1 || n_patterns > 16) return fail(nullptr, TPS_EINVAL, "the follower scan takes 1..16 k-mers");
  if (k < 1 || k > 8) return fail(nullptr, TPS_EINVAL, "the follower scan takes k-mers of 1..8 bases");
  if (match_len < k || match_len > 64) return fail(nullptr, TPS_EINVAL, "match_len must be in k..64");
  if (upto <= skip || upto - skip > 65536) return fail(nullptr, TPS_EINVAL, "need skip < upto <= skip + 65536");
  if (offsets[0] != 0) return fail(nullptr, TPS_EINVAL, "offsets[0] must be 0");
  const uint32_t wpr = (upto - skip + 31) / 32;
  if (sel_words < (uint64_t)n_reads * 2 * n_patterns * wpr)
    return fail(nullptr, TPS_ECAPACITY, "sel_out holds %llu words, %llu needed", (unsigned long long)sel_words,
                (unsigned long long)((uint64_t)n_reads * 2 * n_patterns * wpr));
  if (n_reads == 0) return TPS_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, TPS_ENODEVICE, "no CUDA device visible; topsicle_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(nullptr, TPS_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
  tps_params pp;
  memset(&pp, 0, sizeof(pp));
  pp.n_patterns = n_patterns;
  for (uint32_t i = 0; i < n_patterns; ++i) {
    pp.pattern_len[i] = (uint8_t)k;
    memcpy(pp.patterns[i], patterns + (size_t)i * k, k);
  }
  TpsPatTable pt;
  int rc = build_pattern_table(&pp, &pt);
  if (rc) return rc;
  const uint64_t n_bases = offsets[n_reads];
  const uint64_t n_tiles = (n_bases + 511) / 512;
  const uint64_t cap_pad = ((n_bases + 2047) / 2048) * 2048 + 2048;
  uint8_t *d_bases = nullptr;
  uint32_t *d_codes = nullptr, *d_flags = nullptr, *d_sel = nullptr;
  uint64_t *d_off = nullptr;
  const uint64_t sel_bytes = (uint64_t)n_reads * 2 * n_patterns * wpr * sizeof(uint32_t);
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&d_bases, cap_pad);
  if (e == cudaSuccess) e = cudaMemset(d_bases, 'N', cap_pad);
  if (e == cudaSuccess) e = cudaMalloc(&d_codes, (n_tiles + 1) * 32 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_flags, (n_tiles + 4) * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_off, ((uint64_t)n_reads + 1) * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_sel, sel_bytes);
  if (e == cudaSuccess) e = cudaMemset(d_sel, 0, sel_bytes); /* a batch of empty reads launches nothing */
  if (e == cudaSuccess) e = cudaMemcpy(d_bases, bases, n_bases, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_off, offsets, ((uint64_t)n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n_tiles) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) {
      const uint32_t stages = 2, smem1 = stages * (TPS_K1T_STAGE_BYTES(4) + 16u);
      const uint64_t want = (n_tiles + TPS_K1T_STAGE_TILES(4) - 1) / TPS_K1T_STAGE_TILES(4);
      const uint64_t cap = (uint64_t)prop.multiProcessorCount * 3;
      tps_pack_tma_kernel<4><<<(int)(want < cap ? want : cap), TPS_K1T_THREADS, smem1>>>(
          reinterpret_cast<const uint4 *>(d_bases), d_codes, d_flags, n_tiles, stages);
      TpsFollowArgs a;
      memset(&a, 0, sizeof(a));
      a.pk.codes = d_codes;
      a.pk.flags = d_flags;
      a.pk.bases = d_bases;
      a.offsets = d_off;
      a.n_reads = n_reads;
      a.min_seq_length = min_seq_length;
      a.skip = skip;
      a.upto = upto;
      a.match_len = match_len;
      a.lin_words = lin_words_for(upto - skip);
      a.words_per_row = wpr;
      a.sel = d_sel;
      const uint32_t smem5 = (4 * n_patterns * k + TPS_K5_WARPS * (3 * a.lin_words + n_patterns * wpr)) * 4;
      void (*fn)(const TpsFollowArgs, const TpsPatTable) = nullptr;
      switch (k) {
#define TPS_PICK5(KK) case KK: fn = tps_follow_kernel<KK>; break;
        TPS_PICK5(1) TPS_PICK5(2) TPS_PICK5(3) TPS_PICK5(4) TPS_PICK5(5) TPS_PICK5(6) TPS_PICK5(7) TPS_PICK5(8)
#undef TPS_PICK5
      }
      if (smem5 > (uint32_t)prop.sharedMemPerBlockOptin) {
        e = cudaErrorInvalidValue;
      } else {
        if (smem5 > 48 * 1024) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5);
        if (e == cudaSuccess) {
          fn<<<(2 * n_reads + TPS_K5_WARPS - 1) / TPS_K5_WARPS, TPS_K5_WARPS * 32, smem5>>>(a, pt);
          e = cudaGetLastError();
        }
      }
    }
  }
  if (e == cudaSuccess) e = cudaMemcpy(sel_out, d_sel, sel_bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_bases); cudaFree(d_codes); cudaFree(d_flags); cudaFree(d_off); cudaFree(d_sel);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(nullptr, e == cudaErrorMemoryAllocation ? TPS_ENOMEM : TPS_ECUDA, "tps_follow_scan: %s", cudaGetErrorString(e));
  }
  return TPS_OK;
}

uint64_t tps_kernel_launches(const tps_ctx *ctx) { return ctx ? ctx->launches : 0; }

int tps_debug_copy(tps_ctx *ctx, int what, void *dst, size_t bytes) {
  if (!ctx || !dst) return TPS_EINVAL;
  Slot &s = ctx->slots[0];
  TPS_CUDA(ctx, cudaSetDevice(ctx->device));
  TPS_CUDA(ctx, cudaStreamSynchronize(s.stream));
  const void *src = nullptr;
  size_t cap = 0;
  switch (what) {
    case 0: src = s.d_codes; cap = ctx->cap_tiles * 32 * sizeof(uint32_t); break;
    case 1: src = s.d_flags; cap = ctx->cap_tiles * sizeof(uint32_t); break;
    case 2: { /* validity masks as K2/K3 derive them (flag word + ASCII bytes of flagged groups) */
      const uint64_t ng = bytes / sizeof(uint16_t);
      if (!s.last_bases || ng > ctx->cap_tiles * 32) return fail(ctx, TPS_EINVAL, "no scanned batch / too many groups");
      uint16_t *tmp = nullptr;
      TPS_CUDA(ctx, cudaMalloc(&tmp, ng * sizeof(uint16_t) + 2));
      tps_debug_valid_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, s.stream>>>(s.d_flags, s.last_bases, tmp, ng);
      cudaError_t e = cudaMemcpyAsync(dst, tmp, ng * sizeof(uint16_t), cudaMemcpyDeviceToHost, s.stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
      cudaFree(tmp);
      if (e != cudaSuccess) return fail(ctx, TPS_ECUDA, "debug validity copy: %s", cudaGetErrorString(e));
      return TPS_OK;
    }
    case 3: src = s.d_pass; cap = (size_t)ctx->max_pass * sizeof(uint32_t); break;
    case 4:
      if (ctx->k3_bitpar && !ctx->k3n_debug_gs)
        return fail(ctx, TPS_ESTATE, "the bit-parallel window kernel keeps its window sums in shared memory (set TPS_K3_DEBUG_GS=1)");
      src = ctx->k3_bitpar ? (const void *)s.d_gs_debug : (const void *)s.d_cw;
      cap = ctx->k3_bitpar ? (size_t)ctx->max_pass * ctx->k3n_gs_cap * sizeof(uint16_t)
                           : (size_t)ctx->max_pass * ctx->cw_stride * sizeof(uint32_t);
      break;
    case 5: { /* geometry: c_w row stride (elements), which K3 is in use, pass capacity, tile size of that K3,
               * which K2 is in use (0 shared-memory staged, 1 register staged, 2 literal set in the instructions) */
      const uint32_t info[5] = {ctx->k3_bitpar ? ctx->k3n_gs_cap : ctx->cw_stride, ctx->k3_bitpar ? 1u : 0u, ctx->max_pass,
                                ctx->k3_bitpar ? ctx->k3n_tile_bases : ctx->k3_tile_bases,
                                ctx->k2c_fn ? 2u : (ctx->k2_reg ? 1u : 0u)};
      if (bytes > sizeof(info)) return fail(ctx, TPS_EINVAL, "debug info is %zu bytes", sizeof(info));
      memcpy(dst, info, bytes);
      return TPS_OK;
    }
    default: return fail(ctx, TPS_EINVAL, "unknown debug array %d", what);
  }
  if (bytes > cap) return fail(ctx, TPS_EINVAL, "debug copy of %zu bytes exceeds array size %zu", bytes, cap);
  TPS_CUDA(ctx, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return TPS_OK;
}

}  // extern "C"
