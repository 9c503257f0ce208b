/*
 * tps_kernels.cuh -- sm_100a kernels of the per-read telomere scan.
 *
 *   K1  tps_pack_tma_kernel     ASCII -> 2-bit codes + invalid-group flags (tps_pack_kernel: register-staged twin)
 *   K2  tps_trc_kernel<K>       per read: head/tail greedy counts, TRC decision, pass list
 *   K3  tps_window_kernel<K>    per (passing read, tile): window counts c_w (+ raw counts)
 *   K4  tps_changepoint_kernel  per passing read: exact single change point
 *
 * Reference semantics: /root/reference/Topsicle/allsteps.py:152-204 (step 1),
 * :207-225 + :227-338 (step 2), :359-464 (step 3), ruptures 1.1.9 Binseg/CostL2.
 * Nothing here is a dense contraction; the path is HBM-bound integer/bit work, so no
 * tensor cores: coalesced 128-bit loads (K1), shared-memory staging of read tiles with a
 * window halo (K3), warp ballot/redux reductions (K2), exact 128-bit rational argmax (K4).
 *
 * K2/K3 are templated on K = the common literal length (1..8) so the per-position match is a
 * fully unrolled chain of LOP3 over register-resident shifted bit planes; K = 0 is the
 * generic path (mixed lengths or literals longer than 8).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/topsicle_b200.h"
#include "tps_bitops.h"

#define TPS_FULL 0xFFFFFFFFu

struct TpsPatTable {
  uint32_t lo[TPS_MAX_PATTERNS];  /* bit j = code bit 0 of literal base j */
  uint32_t hi[TPS_MAX_PATTERNS];  /* bit j = code bit 1 */
  uint8_t len[TPS_MAX_PATTERNS];
  uint8_t bordered[TPS_MAX_PATTERNS]; /* literal overlaps itself -> greedy walk needed */
  uint8_t brow[TPS_MAX_PATTERNS];     /* row index among bordered literals */
  uint32_t n;
  uint32_t n_bordered;
  uint64_t bordered_mask; /* bit p = literal p overlaps itself */
  uint32_t paired;        /* n even and literal p + n/2 is the base-wise complement of literal p (what
                             patterns_to_search builds, allsteps.py:104-120): code ^ 2, i.e. the same plane-0
                             bits and inverted plane-1 bits, so the two share half of the match */
};

struct TpsPacked {
  const uint32_t *codes; /* 1 word / 16 bases */
  const uint32_t *flags; /* 1 bit / 16 bases */
  const uint8_t *bases;  /* the ASCII batch itself: exact validity of a flagged group is recomputed from it */
};

/* counters[]: [0] n_pass, [1] K3 work cursor, [2,3] rawcount cursor (u64), [4] overflow flags,
 * [5] K4 work cursor, [6] read items of the bit-parallel K3 */
#define TPS_OVF_RAWCOUNT 1u
#define TPS_OVF_PASS 2u

struct TpsScanArgs {
  TpsPacked pk;
  const uint64_t *offsets; /* read starts; without `lens`, offsets[r+1] ends read r (back-to-back batch) */
  const uint32_t *lens;    /* read lengths of a span batch (reads separated by gaps), or null */
  const uint32_t *true_lens;   /* ends batch: the read's real length L (the batch holds head + tail only), or null */
  const uint8_t *force_tails;  /* region batch: per-read TPS_TAIL_* chosen by an earlier step-1 scan, or null */
  uint32_t n_reads;
  tps_row *rows;
  uint32_t *pass_list;
  uint32_t *counters;
  uint32_t min_seq_length, no_bp, count_threshold;
  uint32_t W, slide, trimfirst, maxlengthtelo, want_rawcount, flags;
  uint8_t *raw;
  uint64_t raw_capacity;
  uint32_t *cw;        /* c_w of passing read i at cw + i * cw_stride */
  uint32_t cw_stride;
  uint32_t max_pass;   /* capacity of pass_list / cw */
  uint32_t lin_words;  /* words per linear plane buffer */
  uint32_t tile_words; /* K3: oriented words per tile incl. halo (+1) */
  uint32_t tile_bases; /* K3: window-start positions per tile */
  uint32_t tiles_max;  /* K3: tiles per read at the longest region */
  uint32_t nq_max;     /* K2: ceil(no_bp/32) */
  /* bit-parallel K3 (tps_window_bp_kernel): K2 appends one TpsReadItem per TRC-pass read with at least 7
   * windows to `items` (counters[6] = number of records) */
  struct TpsReadItem *items;
  uint32_t nz;              /* count planes, 2^nz > P */
  uint32_t bp_tile_windows; /* windows per tile at most (a multiple of 5) */
  uint32_t gs_cap;          /* uint16 group sums per read the kernel's shared memory holds */
  uint32_t no_groups;       /* 1 = every window on its own (A/B of the five-window fast path) */
  uint16_t *gs_debug;       /* test hook (TPS_K3_DEBUG_GS=1): group sums also go to gs_debug + slot * gs_cap, else null */
  /* split mode (few passing reads: fewer than half as many as resident CTAs): the tiles of a read are dealt over
   * 2 or 4 CTAs, whose group sums meet in gs_rows + slot * gs_cap; tile_done[slot] counts the CTAs that are done
   * (zeroed by K2 at the append) and the last one finds the change point */
  uint16_t *gs_rows;
  uint32_t *tile_done;
  uint32_t no_split;        /* 1 = never split (A/B) */
};

/* One work item of the bit-parallel K3 = one TRC-pass read: what a CTA needs to walk its tiles, worked out once
 * by K2 (which has the read's start, length and orientation at hand) so that a read starts with one 32-byte load. */
struct __align__(16) TpsReadItem {
  uint64_t g_edge; /* forward: batch index of region position 0; reverse: one past the batch index of region position 0 */
  uint32_t n_windows;
  uint32_t read;   /* row index */
  uint32_t rev;    /* 1 = reverse tail: region position j is batch index g_edge - 1 - j */
  uint32_t slot;   /* pass-list slot */
  uint32_t reserved[2];
};


/* ------------------------------------------------------------------------------------ K1 */
__device__ __forceinline__ uint4 tps_ldg_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

#define TPS_K1_THREADS 256

/* Validity of 16-base group g (bit i = base 16g + i is one of ACGTacgt).  K1 only flags groups that hold
 * another byte (N, IUPAC codes, anything); the few consumers that meet a flagged group -- K2 and K3 touch
 * the read ends and the regions of TRC-pass reads, ~0.5 % of the batch -- rebuild the exact mask from the
 * 16 ASCII bytes, which are still resident.  Keeping that out of K1 removes a fifth of its instructions. */
__device__ __forceinline__ uint32_t tps_group_valid(const uint32_t *__restrict__ flags, const uint8_t *__restrict__ bases,
                                                    uint64_t g) {
  const uint32_t fw = __ldg(flags + (g >> 5));
  if (!((fw >> (g & 31)) & 1u)) return 0xFFFFu;
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bases) + g);
  return tps_exact_mask16_simd(v.x, v.y, v.z, v.w);
}

#ifdef TPS_TUNING
/* tuning probes (not used by the product): V=1 codes only, V=2 load + trivial store */
template <int U, int V>
__global__ void __launch_bounds__(TPS_K1_THREADS)
tps_pack_probe(const uint4 *__restrict__ bases, uint32_t *__restrict__ codes, uint32_t *__restrict__ flags,
               uint64_t n_tiles) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = (uint64_t)blockIdx.x * (TPS_K1_THREADS / 32) + (threadIdx.x >> 5);
  const uint64_t stride = (uint64_t)gridDim.x * (TPS_K1_THREADS / 32) * U;
  const uint64_t n_main = n_tiles / U * U;
  for (uint64_t t0 = warp * U; t0 < n_main; t0 += stride) {
    uint4 v[U];
    const uint64_t g0 = t0 * 32 + lane;
    const uint4 *src = bases + g0;
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = tps_ldg_stream(src + u * 32);
    uint32_t *cp = codes + g0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (V == 1) {
        uint32_t w0 = v[u].x, w1 = v[u].y, w2 = v[u].z, w3 = v[u].w;
        cp[u * 32] = ((w0 >> 1) & 0x03030303u) | ((w1 << 1) & 0x0C0C0C0Cu) | ((w2 << 3) & 0x30303030u) |
                     ((w3 << 5) & 0xC0C0C0C0u);
      } else if (V == 4) { /* codes + flags, no exact-mask path */
        uint32_t bad;
        cp[u * 32] = tps_pack16(v[u].x, v[u].y, v[u].z, v[u].w, &bad);
        const uint32_t fl = __ballot_sync(TPS_FULL, bad != 0u);
        if (lane == 0) flags[t0 + u] = fl;
      } else if (V == 5) { /* codes + per-lane bad OR-ed into one store, no ballot */
        uint32_t bad;
        cp[u * 32] = tps_pack16(v[u].x, v[u].y, v[u].z, v[u].w, &bad);
        if (bad) flags[t0 + u] = 1;
      } else {
        cp[u * 32] = v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
      }
    }
  }
}
#endif

/* One warp converts 512 consecutive bases per tile: lane l loads bytes [16l,16l+16) as one
 * 128-bit load (512 B coalesced per warp instruction), emits one code word (128 B coalesced
 * store per warp), and the warp emits one flag word per tile by ballot (lane 0 stores the U
 * flag words of its U consecutive tiles as 128-bit vectors).  U tiles are loaded before any is
 * converted so each thread keeps U x 16 B in flight.  U must be a multiple of 4. */
template <int U, int MINB = 4>
__global__ void __launch_bounds__(TPS_K1_THREADS, MINB)
tps_pack_kernel(const uint4 *__restrict__ bases, uint32_t *__restrict__ codes,
                uint32_t *__restrict__ flags, uint64_t n_tiles) {
  static_assert(U % 4 == 0, "U must be a multiple of 4 (vector flag stores)");
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = (uint64_t)blockIdx.x * (TPS_K1_THREADS / 32) + (threadIdx.x >> 5);
  const uint64_t stride = (uint64_t)gridDim.x * (TPS_K1_THREADS / 32) * U;
  const uint64_t n_main = n_tiles / U * U;
  for (uint64_t t0 = warp * U; t0 < n_main; t0 += stride) {
    uint4 v[U];
    uint32_t fl[U];
    const uint64_t g0 = t0 * 32 + lane;
    const uint4 *src = bases + g0;
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = tps_ldg_stream(src + u * 32);
    uint32_t *cp = codes + g0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t bad;
      cp[u * 32] = tps_pack16(v[u].x, v[u].y, v[u].z, v[u].w, &bad);
      fl[u] = __ballot_sync(TPS_FULL, bad != 0u);
    }
    if (lane == 0) {
      uint4 *fp = reinterpret_cast<uint4 *>(flags + t0); /* t0 is a multiple of U: 16-byte aligned */
#pragma unroll
      for (int u = 0; u < U; u += 4) fp[u / 4] = make_uint4(fl[u], fl[u + 1], fl[u + 2], fl[u + 3]);
    }
  }
  const uint64_t tt = n_main + warp; /* at most U-1 trailing tiles, one warp each */
  if (tt < n_tiles) {
    const uint64_t g = tt * 32 + lane;
    const uint4 v = tps_ldg_stream(bases + g);
    uint32_t bad;
    codes[g] = tps_pack16(v.x, v.y, v.z, v.w, &bad);
    const uint32_t f = __ballot_sync(TPS_FULL, bad != 0u);
    if (lane == 0) flags[tt] = f;
  }
}

/* ------------------------------------------------------------- K1, bulk-copy (TMA) staged
 * Same conversion, but the ASCII bytes reach the SM through the bulk async-copy engine
 * (cp.async.bulk global -> shared, completion counted on an mbarrier) instead of per-lane
 * 128-bit loads: one producer lane keeps `n_stages` chunks of 16 KiB in flight per CTA without
 * holding a single register for them, eight consumer warps read their 512-byte tiles back with
 * conflict-free LDS.128 (lane l owns bytes [16l, 16l+16) of a tile, exactly as in the
 * register-staged kernel), hand the stage back through an `empty` mbarrier and convert.
 * Chunk c of the batch goes to CTA c mod gridDim.x, so at any time the whole grid streams one
 * contiguous span of the input. */
#define TPS_K1T_CWARPS 8
#define TPS_K1T_THREADS (TPS_K1T_CWARPS * 32 + 32)
/* U = tiles per consumer warp and stage: a stage holds 8 * U tiles = 4 * U KiB (U = 2 or 4) */
#define TPS_K1T_STAGE_TILES(U) (TPS_K1T_CWARPS * (U))
#define TPS_K1T_STAGE_BYTES(U) (TPS_K1T_STAGE_TILES(U) * 512)

__device__ __forceinline__ uint32_t tps_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tps_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tps_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tps_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tps_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TPS_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TPS_DONE_%=;\n"
      "bra TPS_WAIT_%=;\n"
      "TPS_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tps_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int U>
__global__ void __launch_bounds__(TPS_K1T_THREADS)
tps_pack_tma_kernel(const uint4 *__restrict__ bases, uint32_t *__restrict__ codes, uint32_t *__restrict__ flags,
                    uint64_t n_tiles, uint32_t n_stages) {
  extern __shared__ __align__(128) uint8_t k1t_smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(k1t_smem + (size_t)n_stages * TPS_K1T_STAGE_BYTES(U));
  const uint32_t full0 = tps_smem_addr(bars), empty0 = full0 + 8u * n_stages;
  const uint32_t data0 = tps_smem_addr(k1t_smem);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint64_t n_chunks = (n_tiles + TPS_K1T_STAGE_TILES(U) - 1) / TPS_K1T_STAGE_TILES(U);
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < n_stages; ++s) {
      tps_mbar_init(full0 + 8u * s, 1u);
      tps_mbar_init(empty0 + 8u * s, TPS_K1T_CWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t s = 0, ph = 0;
  if (warp == TPS_K1T_CWARPS) { /* producer */
    if (lane == 0) {
      for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        tps_mbar_wait(empty0 + 8u * s, ph ^ 1u); /* all eight consumer warps have read the stage */
        const uint64_t t0 = c * TPS_K1T_STAGE_TILES(U);
        const uint64_t left = n_tiles - t0;
        const uint32_t bytes = (uint32_t)(left < TPS_K1T_STAGE_TILES(U) ? left : TPS_K1T_STAGE_TILES(U)) * 512u;
        tps_mbar_expect_tx(full0 + 8u * s, bytes);
        tps_bulk_g2s(data0 + s * TPS_K1T_STAGE_BYTES(U), reinterpret_cast<const uint8_t *>(bases) + t0 * 512u, bytes,
                     full0 + 8u * s);
        if (++s == n_stages) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    tps_mbar_wait(full0 + 8u * s, ph);
    const uint4 *sp = reinterpret_cast<const uint4 *>(k1t_smem + (size_t)s * TPS_K1T_STAGE_BYTES(U)) +
                      warp * (U * 32) + lane;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = sp[u * 32];
    __syncwarp();
    if (lane == 0) tps_mbar_arrive(empty0 + 8u * s);
    if (++s == n_stages) { s = 0; ph ^= 1u; }
    const uint64_t t0 = c * TPS_K1T_STAGE_TILES(U) + warp * U;
    const uint64_t g0 = t0 * 32 + lane;
    if (t0 + U <= n_tiles) {
      uint32_t fl[U];
      uint32_t *cp = codes + g0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        uint32_t bad;
        cp[u * 32] = tps_pack16(v[u].x, v[u].y, v[u].z, v[u].w, &bad);
        fl[u] = __ballot_sync(TPS_FULL, bad != 0u);
      }
      if (lane == 0) { /* t0 is a multiple of U: the flag words of the warp's tiles go out as one vector */
        if constexpr (U == 4) *reinterpret_cast<uint4 *>(flags + t0) = make_uint4(fl[0], fl[1], fl[2], fl[3]);
        else *reinterpret_cast<uint2 *>(flags + t0) = make_uint2(fl[0], fl[1]);
      }
    } else { /* last chunk of the batch: tiles past n_tiles were not copied */
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t0 + u < n_tiles) {
          uint32_t bad;
          codes[g0 + u * 32] = tps_pack16(v[u].x, v[u].y, v[u].z, v[u].w, &bad);
          const uint32_t f = __ballot_sync(TPS_FULL, bad != 0u);
          if (lane == 0) flags[t0 + u] = f;
        }
      }
    }
  }
}

/* Test hook: the validity mask K2/K3 would see for every group of the batch. */
__global__ void tps_debug_valid_kernel(const uint32_t *__restrict__ flags, const uint8_t *__restrict__ bases,
                                       uint16_t *__restrict__ out, uint64_t n_groups) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) out[g] = (uint16_t)tps_group_valid(flags, bases, g);
}

/* ----------------------------------------------------------------- staging (K2 and K3) */
/* Linear plane buffers: three arrays of `lw` 32-bit words (plane0, plane1, valid).  Viewed
 * as uint16: entries 0,1 are a zero pad, entry 2+i holds 16-base group (g0>>4)+i in linear
 * base order.  Bit index of base g0+x is 32 + phase + x, phase = g0 & 15. */
__device__ __forceinline__ void tps_stage_linear(const TpsPacked &pk, uint64_t g0, uint32_t n,
                                                 uint32_t *lin, uint32_t lw, uint32_t tid,
                                                 uint32_t nthreads) {
  uint16_t *l0 = reinterpret_cast<uint16_t *>(lin);
  uint16_t *l1 = reinterpret_cast<uint16_t *>(lin + lw);
  uint16_t *lv = reinterpret_cast<uint16_t *>(lin + 2 * lw);
  const uint64_t gfirst = g0 >> 4;
  const uint32_t ng = n ? (uint32_t)(((g0 + n + 15) >> 4) - gfirst) : 0u;
  for (uint32_t i = tid; i < 2 * lw; i += nthreads) {
    if (i < 2 || i >= 2 + ng) {
      l0[i] = 0;
      l1[i] = 0;
      lv[i] = 0;
    } else {
      const uint64_t g = gfirst + (i - 2);
      uint32_t y = tps_linear_planes(__ldg(pk.codes + g));
      uint32_t v = tps_group_valid(pk.flags, pk.bases, g);
      l0[i] = (uint16_t)(y & 0xFFFFu);
      l1[i] = (uint16_t)(y >> 16);
      lv[i] = (uint16_t)v;
    }
  }
}

/* Oriented word q (positions 32q..32q+31 of the slice of n bases; reversed slice if rev). */
__device__ __forceinline__ void tps_oriented_word(const uint32_t *lin, uint32_t lw, uint32_t phase,
                                                  uint32_t n, bool rev, uint32_t q, uint32_t &p0,
                                                  uint32_t &p1, uint32_t &v) {
  if ((uint64_t)q * 32u >= n) {
    p0 = p1 = v = 0u;
    return;
  }
  const uint32_t rem = n - q * 32u;
  const uint32_t m = rem >= 32u ? TPS_FULL : ((1u << rem) - 1u);
  const uint32_t *l0 = lin, *l1 = lin + lw, *lv = lin + 2 * lw;
  if (!rev) {
    const uint32_t idx = 1u + q;
    p0 = __funnelshift_r(l0[idx], l0[idx + 1], phase);
    p1 = __funnelshift_r(l1[idx], l1[idx + 1], phase);
    v = __funnelshift_r(lv[idx], lv[idx + 1], phase) & m;
  } else {
    const uint32_t o = phase + n - q * 32u; /* = 32 + phase + (n - 32q - 32) */
    const uint32_t idx = o >> 5, sh = o & 31u;
    p0 = __brev(__funnelshift_r(l0[idx], l0[idx + 1], sh));
    p1 = __brev(__funnelshift_r(l1[idx], l1[idx + 1], sh));
    v = __brev(__funnelshift_r(lv[idx], lv[idx + 1], sh)) & m;
  }
}

/* ------------------------------------------------------------------ per-position matching */
/* Shared-memory literal masks: pm[p * KS + j] = (x, y) with x = all-ones iff code bit 0 of
 * base j of literal p is 1, y likewise for code bit 1.  KS = max(K, 1). */
__device__ __forceinline__ void tps_build_pattern_masks(uint2 *pm, const TpsPatTable &pt, uint32_t ks,
                                                        uint32_t tid, uint32_t nthreads) {
  for (uint32_t i = tid; i < pt.n * ks; i += nthreads) {
    const uint32_t p = i / ks, j = i - p * ks;
    pm[i] = make_uint2(0u - ((pt.lo[p] >> j) & 1u), 0u - ((pt.hi[p] >> j) & 1u));
  }
}

/* Planes of 32 consecutive positions shifted by 0..K-1, valid = all K bases valid. */
template <int K>
struct TpsWin {
  uint32_t X[K > 0 ? K : 1], Y[K > 0 ? K : 1], V;
  uint32_t a0, b0, a1, b1, av, bv;
};

template <int K>
__device__ __forceinline__ void tps_win_init(TpsWin<K> &w, uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1,
                                             uint32_t av, uint32_t bv) {
  if constexpr (K > 0) {
    w.V = av;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      w.X[j] = __funnelshift_r(a0, b0, j);
      w.Y[j] = __funnelshift_r(a1, b1, j);
      if (j) w.V &= __funnelshift_r(av, bv, j);
    }
  } else {
    w.a0 = a0; w.b0 = b0; w.a1 = a1; w.b1 = b1; w.av = av; w.bv = bv;
  }
}

/* Match word of literal p: bit i = literal occurs at position 32q+i (all bases valid). */
template <int K>
__device__ __forceinline__ uint32_t tps_win_match(const TpsWin<K> &w, const uint2 *pm, const TpsPatTable &pt,
                                                  uint32_t p) {
  if constexpr (K > 0) {
    uint32_t M = w.V;
    const uint2 *c = pm + p * K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const uint2 m = c[j];
      M &= ~(w.X[j] ^ m.x);
      M &= ~(w.Y[j] ^ m.y);
    }
    return M;
  } else {
    uint32_t M = TPS_FULL;
    const uint32_t lo = pt.lo[p], hi = pt.hi[p], k = pt.len[p];
    for (uint32_t j = 0; j < k; ++j) {
      const uint32_t cx = 0u - ((lo >> j) & 1u), cy = 0u - ((hi >> j) & 1u);
      M &= ~(__funnelshift_r(w.a0, w.b0, j) ^ cx) & ~(__funnelshift_r(w.a1, w.b1, j) ^ cy) &
           __funnelshift_r(w.av, w.bv, j);
    }
    return M;
  }
}

/* Match words of literal p and of its complement p + n/2 (TpsPatTable.paired): the plane-0 comparison is
 * shared, the plane-1 comparison of the complement is the inverted one -- 14 LOP3 for the pair instead of 18. */
template <int K>
__device__ __forceinline__ void tps_win_match_pair(const TpsWin<K> &w, const uint2 *pm, uint32_t p, uint32_t &M,
                                                   uint32_t &Mc) {
  static_assert(K > 0, "paired matching needs a common literal length");
  const uint2 *c = pm + p * K;
  uint32_t tx = 0u, ty = 0u, tyc = 0u;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const uint2 m = c[j];
    tx |= w.X[j] ^ m.x;
    ty |= w.Y[j] ^ m.y;
    tyc |= ~(w.Y[j] ^ m.y);
  }
  M = w.V & ~tx & ~ty;
  Mc = w.V & ~tx & ~tyc;
}

/* ---- step-1 row logic shared by the K2 kernels ------------------------------------------------------- */
__device__ __forceinline__ void tps_row_init(tps_row &row, uint32_t true_len) {
  row.length = true_len;
  row.status = TPS_ST_FILTERED;
  row.tail = 0; row.best_pattern = 0; row.reserved0 = 0;
  row.match_count = 0; row.head_max = 0; row.tail_max = 0; row.reserved1 = 0;
  row.n_windows = 0; row.bkp = -1; row.telo_length = -1; row.reserved2 = 0;
  row.rawcount_offset = ~0ull;
}

/* The 40-byte row as five 64-bit stores (rows are 8-byte aligned). */
__device__ __forceinline__ void tps_store_row(tps_row *dst, const tps_row &row) {
  uint2 *d = reinterpret_cast<uint2 *>(dst);
  d[0] = make_uint2(row.length, (uint32_t)row.status | ((uint32_t)row.tail << 8) | ((uint32_t)row.best_pattern << 16) |
                                    ((uint32_t)row.reserved0 << 24));
  d[1] = make_uint2((uint32_t)row.match_count | ((uint32_t)row.head_max << 16),
                    (uint32_t)row.tail_max | ((uint32_t)row.reserved1 << 16));
  d[2] = make_uint2(row.n_windows, (uint32_t)row.bkp);
  d[3] = make_uint2((uint32_t)row.telo_length, row.reserved2);
  d[4] = make_uint2((uint32_t)row.rawcount_offset, (uint32_t)(row.rawcount_offset >> 32));
}

/* Tail decision, cutoff test and (lane 0) the append to the pass list, from the first-max counts of the two
 * ends: ms / ps of seq[:no_bp], me / pe of the reversed seq[-no_bp:] (allsteps.py:190-198). */
__device__ __forceinline__ void tps_trc_decide(const TpsScanArgs &a, uint32_t n_patterns, uint32_t r, uint64_t off,
                                               uint32_t L, uint32_t Lt, uint32_t lane, uint32_t ms, uint32_t ps,
                                               uint32_t me, uint32_t pe, tps_row &row) {
  bool fwd = ms > me; /* tie -> reverse, allsteps.py:193-198 */
  if (a.flags & TPS_FLAG_FORCE_FORWARD) fwd = true; /* caller-chosen tail, allsteps.py:294-297 */
  if (a.flags & TPS_FLAG_FORCE_REVERSE) fwd = false;
  if (a.force_tails) fwd = a.force_tails[r] == TPS_TAIL_FORWARD;
  const uint32_t cnt = fwd ? ms : me;
  row.tail = fwd ? TPS_TAIL_FORWARD : TPS_TAIL_REVERSE;
  row.best_pattern = (uint8_t)(fwd ? ps : pe);
  row.match_count = (uint16_t)cnt;
  row.head_max = (uint16_t)ms;
  row.tail_max = (uint16_t)me;
  /* a region batch holds reads that passed step 1 on their whole ends; its own count may see fewer bases */
  row.status = (cnt >= a.count_threshold || a.force_tails) ? TPS_ST_PASS : TPS_ST_BELOW;
  if (row.status == TPS_ST_PASS && lane == 0 && !(a.flags & TPS_FLAG_STEP1_ONLY)) {
    const uint32_t M = Lt < a.maxlengthtelo ? Lt : a.maxlengthtelo;
    const uint32_t nreg = M > a.trimfirst ? M - a.trimfirst : 0u;
    const uint32_t nW = nreg >= a.W ? (nreg - a.W) / a.slide + 1u : 0u;
    row.n_windows = nW;
    if (nW < 7u) row.status = TPS_ST_BADSEG; /* ruptures' sanity_check: jump 5, min_size 2, one breakpoint need n >= 7 */
    const uint32_t slot = atomicAdd(a.counters + 0, 1u);
    if (slot < a.max_pass) {
      a.pass_list[slot] = r;
      if (a.items && nW >= 7u) { /* one work item per read for the bit-parallel K3 */
        TpsReadItem it;
        it.g_edge = fwd ? off + a.trimfirst : off + L - a.trimfirst;
        it.n_windows = nW;
        it.read = r;
        it.rev = fwd ? 0u : 1u;
        it.slot = slot;
        it.reserved[0] = it.reserved[1] = 0u;
        a.items[atomicAdd(a.counters + 6, 1u)] = it;
        a.tile_done[slot] = 0u;
      }
      if (a.want_rawcount && nW) {
        const unsigned long long elems = (unsigned long long)nW * n_patterns;
        const unsigned long long at = atomicAdd(reinterpret_cast<unsigned long long *>(a.counters + 2), elems);
        if (at + elems <= a.raw_capacity) row.rawcount_offset = at;
        else atomicOr(a.counters + 4, TPS_OVF_RAWCOUNT);
      }
    } else {
      atomicOr(a.counters + 4, TPS_OVF_PASS);
    }
  }
}

/* ------------------------------------------------------------------------------------ K2 */
#define TPS_K2_WARPS 4

/* Greedy counts of every literal over one oriented slice (head or reversed tail), one warp.
 * Returns (max count, first-max literal index) -- allsteps.py:181-191. */
template <int K>
__device__ __forceinline__ void tps_trc_end(const TpsScanArgs &a, const TpsPatTable &pt, const uint2 *pm,
                                            uint64_t g0, uint32_t n, bool rev, uint32_t *lin, uint32_t *mrows,
                                            uint32_t *cnts, uint32_t lane, uint32_t &best, uint32_t &bestp) {
  const uint32_t lw = a.lin_words;
  const uint32_t phase = (uint32_t)(g0 & 15u);
  const uint32_t nq = (n + 31u) >> 5;
  tps_stage_linear(a.pk, g0, n, lin, lw, lane, 32u);
  for (uint32_t p = lane; p < pt.n; p += 32u) cnts[p] = 0u;
  __syncwarp();
  for (uint32_t qc = 0; qc < nq; qc += 32u) {
    const uint32_t q = qc + lane;
    uint32_t a0, a1, av, b0, b1, bv;
    tps_oriented_word(lin, lw, phase, n, rev, q, a0, a1, av);
    tps_oriented_word(lin, lw, phase, n, rev, q + 1u, b0, b1, bv);
    TpsWin<K> win;
    tps_win_init<K>(win, a0, b0, a1, b1, av, bv);
    for (uint32_t p = 0; p < pt.n; ++p) {
      const uint32_t M = tps_win_match<K>(win, pm, pt, p);
      if (pt.bordered[p]) { /* warp-uniform */
        if (q < nq) mrows[pt.brow[p] * a.nq_max + q] = M;
      } else {
        const uint32_t c = __reduce_add_sync(TPS_FULL, tps_popc32(M));
        if (lane == 0) cnts[p] += c;
      }
    }
  }
  __syncwarp();
  /* self-overlapping literals: each lane walks one literal's match row */
  for (uint32_t p = lane; p < pt.n; p += 32u) {
    if (pt.bordered[p]) {
      const uint32_t k = pt.len[p];
      cnts[p] = (n >= k) ? tps_greedy_count(mrows + pt.brow[p] * a.nq_max, 0, (int32_t)(n - k), k) : 0u;
    }
  }
  __syncwarp();
  best = 0u;
  bestp = 0u;
  for (uint32_t p = 0; p < pt.n; ++p) {
    const uint32_t c = cnts[p];
    if (c > best) { /* strict: first maximum wins */
      best = c;
      bestp = p;
    }
  }
  __syncwarp();
}

/* dynamic shared memory (words): pm[2 * P * max(K,1)] | per warp: lin[3*lin_words] |
 * mrows[n_bordered * nq_max] | cnts[TPS_MAX_PATTERNS] */
template <int K>
__global__ void __launch_bounds__(TPS_K2_WARPS * 32)
tps_trc_kernel(const TpsScanArgs a, const TpsPatTable pt) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr uint32_t KS = K > 0 ? K : 1;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint2 *pm = reinterpret_cast<uint2 *>(smem);
  tps_build_pattern_masks(pm, pt, KS, threadIdx.x, TPS_K2_WARPS * 32);
  __syncthreads();
  const uint32_t per_warp = 3u * a.lin_words + pt.n_bordered * a.nq_max + TPS_MAX_PATTERNS;
  uint32_t *lin = smem + 2u * pt.n * KS + warp * per_warp;
  uint32_t *mrows = lin + 3u * a.lin_words;
  uint32_t *cnts = mrows + pt.n_bordered * a.nq_max;
  const uint32_t r = blockIdx.x * TPS_K2_WARPS + warp;
  if (r >= a.n_reads) return;
  const uint64_t off = a.offsets[r];
  const uint32_t L = a.lens ? a.lens[r] : (uint32_t)(a.offsets[r + 1] - off);
  /* ends batch: the slice geometry comes from the uploaded head + tail (L bases), the length filter, the row
   * and the window count from the read's real length */
  const uint32_t Lt = a.true_lens ? a.true_lens[r] : L;
  tps_row row;
  tps_row_init(row, Lt);
  if (Lt > a.min_seq_length || a.force_tails) { /* strict, allsteps.py:175; region batches passed it already */
    const uint32_t n = L < a.no_bp ? L : a.no_bp;
    uint32_t ms, ps, me, pe;
    tps_trc_end<K>(a, pt, pm, off, n, false, lin, mrows, cnts, lane, ms, ps);         /* seq[:no_bp] */
    tps_trc_end<K>(a, pt, pm, off + L - n, n, true, lin, mrows, cnts, lane, me, pe);  /* seq[-no_bp:][::-1] */
    tps_trc_decide(a, pt.n, r, off, L, Lt, lane, ms, ps, me, pe, row);
  }
  if (lane == 0) a.rows[r] = row;
}

/* ---------------------------------------------------------------------- K2, register-staged
 * Fast path of step 1 for no_bp <= 1000 and <= 32 literals (every CLI run: no_bp is 1000,
 * main.py:57).  The slice of one read end (<= 1000 bases + <= 15 bases of group misalignment
 * <= 1024) fits one 32-bit word of each bit plane per lane, so staging needs no shared memory:
 * lane l converts groups 2l, 2l+1 to linear planes in registers, the oriented words come from
 * warp shuffles, every literal costs one LOP3 chain + POPC + REDUX, the first maximum is one
 * REDUX.MAX over (count << 8 | 255 - index).  Only self-overlapping literals still write their
 * match row to shared memory for the greedy walk. */
template <int K>
__device__ __forceinline__ void tps_trc_stage_reg(const TpsScanArgs &a, uint64_t g0, uint32_t n, bool rev, uint32_t lane,
                                                  TpsWin<K> &win) {
  const uint32_t phase = (uint32_t)(g0 & 15u);
  const uint64_t gfirst = g0 >> 4;
  const uint32_t ng = n ? (uint32_t)(((g0 + n + 15) >> 4) - gfirst) : 0u; /* <= 64 */
  /* linear planes of bases [16*gfirst + 32*lane, +32): bit x of the slice is linear bit phase + x */
  uint32_t L0 = 0u, L1 = 0u, LV = 0u;
#pragma unroll
  for (uint32_t h = 0; h < 2u; ++h) {
    const uint32_t gi = 2u * lane + h;
    const bool in = gi < ng;
    const uint64_t g = gfirst + gi;
    uint32_t y = 0u, v = in ? 0xFFFFu : 0u;
    bool flagged = false;
    if (in) {
      y = tps_linear_planes(__ldg(a.pk.codes + g));
      flagged = (__ldg(a.pk.flags + (g >> 5)) >> (g & 31)) & 1u;
    }
    /* exact validity of the (rare) flagged groups, one group at a time by the whole warp: 16 lanes test one
     * ASCII byte each and a ballot is the mask -- a dozen instructions instead of the 65 of the per-lane
     * bit-sliced formula that every lane would execute for the sake of one */
    for (uint32_t fb = __ballot_sync(TPS_FULL, flagged); fb; fb &= fb - 1u) {
      const uint32_t src = (uint32_t)__ffs((int)fb) - 1u;
      const uint64_t gg = gfirst + 2u * src + h;
      const uint32_t c = lane < 16u ? (uint32_t)__ldg(a.pk.bases + 16u * gg + lane) : 0x41u;
      const uint32_t m = __ballot_sync(TPS_FULL, tps_byte_is_acgt(c) != 0) & 0xFFFFu;
      if (lane == src) v = m;
    }
    L0 |= (y & 0xFFFFu) << (16u * h);
    L1 |= (y >> 16) << (16u * h);
    LV |= v << (16u * h);
  }
  /* oriented word `lane`: positions 32*lane .. 32*lane+31 of the (reversed) slice */
  uint32_t a0, a1, av;
  if (!rev) {
    uint32_t n0 = __shfl_down_sync(TPS_FULL, L0, 1), n1 = __shfl_down_sync(TPS_FULL, L1, 1),
             nv = __shfl_down_sync(TPS_FULL, LV, 1);
    if (lane == 31u) n0 = n1 = nv = 0u;
    a0 = __funnelshift_r(L0, n0, phase);
    a1 = __funnelshift_r(L1, n1, phase);
    av = __funnelshift_r(LV, nv, phase);
  } else {
    /* reversed position j is slice base n-1-j: word `lane` is the bit-reversal of the 32 linear bits
     * starting at o = phase + n - 32*lane - 32 (o < 0: the part below bit 0 is masked out below) */
    const int32_t o = (int32_t)(phase + n) - 32 * (int32_t)lane - 32;
    const int32_t idx = o >> 5; /* floor; >= -1 for every word that holds slice bases */
    const uint32_t sh = (uint32_t)o & 31u;
    const int lo_src = idx & 31, hi_src = (idx + 1) & 31;
    uint32_t l0 = __shfl_sync(TPS_FULL, L0, lo_src), l1 = __shfl_sync(TPS_FULL, L1, lo_src),
             lv = __shfl_sync(TPS_FULL, LV, lo_src);
    uint32_t h0 = __shfl_sync(TPS_FULL, L0, hi_src), h1 = __shfl_sync(TPS_FULL, L1, hi_src),
             hv = __shfl_sync(TPS_FULL, LV, hi_src);
    if (idx < 0) l0 = l1 = lv = 0u;
    if (idx + 1 < 0 || idx + 1 > 31) h0 = h1 = hv = 0u;
    a0 = __brev(__funnelshift_r(l0, h0, sh));
    a1 = __brev(__funnelshift_r(l1, h1, sh));
    av = __brev(__funnelshift_r(lv, hv, sh));
  }
  {
    const uint32_t done = 32u * lane;
    const uint32_t rem = n > done ? n - done : 0u;
    av &= rem >= 32u ? TPS_FULL : ((1u << rem) - 1u);
  }
  uint32_t b0 = __shfl_down_sync(TPS_FULL, a0, 1), b1 = __shfl_down_sync(TPS_FULL, a1, 1),
           bv = __shfl_down_sync(TPS_FULL, av, 1);
  if (lane == 31u) b0 = b1 = bv = 0u;
  tps_win_init<K>(win, a0, b0, a1, b1, av, bv);
}

template <int K>
__device__ __forceinline__ void tps_trc_end_reg(const TpsScanArgs &a, const TpsPatTable &pt, const uint2 *pm,
                                                uint64_t g0, uint32_t n, bool rev, uint32_t *mrows,
                                                uint32_t lane, uint32_t &best, uint32_t &bestp) {
  TpsWin<K> win;
  tps_trc_stage_reg<K>(a, g0, n, rev, lane, win);
  const uint32_t nq = (n + 31u) >> 5;
  uint32_t mine = 0u; /* count of literal `lane` */
  bool done = false;
  if constexpr (K > 0) {
    if (pt.paired) { /* literal p and its complement p + U together: shared plane-0 compare, one REDUX for both */
      const uint32_t U = pt.n >> 1;
#pragma unroll 2
      for (uint32_t p = 0; p < U; ++p) {
        uint32_t M, Mc;
        tps_win_match_pair<K>(win, pm, p, M, Mc);
        const uint32_t c = __reduce_add_sync(TPS_FULL, tps_popc32(M) | (tps_popc32(Mc) << 16));
        if (lane == p) mine = c & 0xFFFFu;
        if (lane == p + U) mine = c >> 16;
      }
      done = true;
    }
  }
  if (!done) {
#pragma unroll 4
    for (uint32_t p = 0; p < pt.n; ++p) { /* branch-free: occurrences of every literal */
      const uint32_t M = tps_win_match<K>(win, pm, pt, p);
      const uint32_t c = __reduce_add_sync(TPS_FULL, tps_popc32(M));
      if (lane == p) mine = c;
    }
  }
  if (pt.n_bordered) { /* self-overlapping literals: occurrences != greedy count, redo those exactly */
    for (uint64_t bm = pt.bordered_mask; bm; bm &= bm - 1) {
      const uint32_t p = (uint32_t)__ffsll((long long)bm) - 1u;
      const uint32_t M = tps_win_match<K>(win, pm, pt, p);
      if (lane < nq) mrows[pt.brow[p] * 32u + lane] = M;
    }
    __syncwarp();
    if (lane < pt.n && pt.bordered[lane]) {
      const uint32_t k = pt.len[lane];
      mine = (n >= k) ? tps_greedy_count(mrows + pt.brow[lane] * 32u, 0, (int32_t)(n - k), k) : 0u;
    }
    __syncwarp();
  }
  /* first maximum in literal order (allsteps.py:190-191): max over (count, -index) */
  const uint32_t key = lane < pt.n ? ((mine << 8) | (255u - lane)) : 0u;
  const uint32_t kmax = __reduce_max_sync(TPS_FULL, key);
  best = kmax >> 8;
  bestp = best ? 255u - (kmax & 255u) : 0u;
}

#ifndef TPS_K2R_WARPS
#define TPS_K2R_WARPS 8
#endif

/* dynamic shared memory (words): pm[2 * P * max(K,1)] | per warp: mrows[n_bordered * 32] */
template <int K>
__global__ void __launch_bounds__(TPS_K2R_WARPS * 32)
tps_trc_reg_kernel(const TpsScanArgs a, const TpsPatTable pt) {
  extern __shared__ __align__(16) uint32_t smem[];
  constexpr uint32_t KS = K > 0 ? K : 1;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint2 *pm = reinterpret_cast<uint2 *>(smem);
  tps_build_pattern_masks(pm, pt, KS, threadIdx.x, TPS_K2R_WARPS * 32);
  __syncthreads();
  uint32_t *mrows = smem + 2u * pt.n * KS + warp * (pt.n_bordered * 32u);
  const uint32_t r = blockIdx.x * TPS_K2R_WARPS + warp;
  if (r >= a.n_reads) return;
  const uint64_t off = a.offsets[r];
  const uint32_t L = a.lens ? a.lens[r] : (uint32_t)(a.offsets[r + 1] - off);
  /* ends batch: the slice geometry comes from the uploaded head + tail (L bases), the length filter, the row
   * and the window count from the read's real length */
  const uint32_t Lt = a.true_lens ? a.true_lens[r] : L;
  tps_row row;
  tps_row_init(row, Lt);
  if (Lt > a.min_seq_length || a.force_tails) { /* strict, allsteps.py:175; region batches passed it already */
    const uint32_t n = L < a.no_bp ? L : a.no_bp;
    uint32_t ms, ps, me, pe;
    tps_trc_end_reg<K>(a, pt, pm, off, n, false, mrows, lane, ms, ps);         /* seq[:no_bp] */
    tps_trc_end_reg<K>(a, pt, pm, off + L - n, n, true, mrows, lane, me, pe);  /* seq[-no_bp:][::-1] */
    tps_trc_decide(a, pt.n, r, off, L, Lt, lane, ms, ps, me, pe, row);
  }
  if (lane == 0) tps_store_row(a.rows + r, row);
}

/* K2 for the literal sets every CLI run builds (patterns_to_search: the U distinct k-mers of pattern + pattern and
 * their base-wise complements, allsteps.py:104-120) when none overlaps itself: U and K are template parameters, the
 * pair loop is unrolled and the 0 / ~0 plane masks of the literals are kernel parameters, i.e. constant-bank operands
 * of the LOP3s -- no mask table in shared memory, no block barrier, no loop or select instructions.  The counts of
 * a pair come out of one REDUX as a warp-uniform value, so the first maximum in literal order (allsteps.py:190-191)
 * is a running max over (count << 8 | 255 - index). */
#define TPS_K2C_MAXU 8
struct TpsPairMasks {
  uint32_t x[TPS_K2C_MAXU][8], y[TPS_K2C_MAXU][8]; /* [pair p][base j]: all-ones if plane 0 / 1 of literal p's base j is 1 */
};

template <int K, int U>
__device__ __forceinline__ void tps_trc_end_const(const TpsScanArgs &a, const TpsPairMasks &pm, uint64_t g0, uint32_t n,
                                                  bool rev, uint32_t lane, uint32_t &best, uint32_t &bestp) {
  TpsWin<K> win;
  tps_trc_stage_reg<K>(a, g0, n, rev, lane, win);
  /* 16-bit keys count << 4 | 15 - index (count <= 1000 / 3, index < 16), the pair's two side by side: the sum over
   * the lanes is linear in the counts, the tie-breakers are added behind it, the running maximum is one packed
   * 16x2 max per pair */
  uint32_t key2 = 0u;
#pragma unroll
  for (int p = 0; p < U; ++p) {
    uint32_t tx = 0u, ty = 0u, tyc = 0u;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      tx |= win.X[j] ^ pm.x[p][j];
      ty |= win.Y[j] ^ pm.y[p][j];
      tyc |= ~(win.Y[j] ^ pm.y[p][j]);
    }
    const uint32_t M = win.V & ~tx & ~ty, Mc = win.V & ~tx & ~tyc;
    const uint32_t c = __reduce_add_sync(TPS_FULL, tps_popc32(M) * 16u + tps_popc32(Mc) * (16u << 16));
    key2 = __vmaxu2(key2, c + ((15u - (uint32_t)p) | ((15u - (uint32_t)(p + U)) << 16)));
  }
  const uint32_t klo = key2 & 0xFFFFu, khi = key2 >> 16;
  const uint32_t key = klo > khi ? klo : khi;
  best = key >> 4;
  bestp = best ? 15u - (key & 15u) : 0u;
}

template <int K, int U>
__global__ void __launch_bounds__(TPS_K2R_WARPS * 32)
tps_trc_const_kernel(const TpsScanArgs a, const TpsPairMasks pm) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t r = blockIdx.x * TPS_K2R_WARPS + warp;
  if (r >= a.n_reads) return;
  const uint64_t off = a.offsets[r];
  const uint32_t L = a.lens ? a.lens[r] : (uint32_t)(a.offsets[r + 1] - off);
  const uint32_t Lt = a.true_lens ? a.true_lens[r] : L; /* ends batch: see tps_trc_reg_kernel */
  tps_row row;
  tps_row_init(row, Lt);
  if (Lt > a.min_seq_length || a.force_tails) {
    const uint32_t n = L < a.no_bp ? L : a.no_bp;
    uint32_t ms, ps, me, pe;
    tps_trc_end_const<K, U>(a, pm, off, n, false, lane, ms, ps);
    tps_trc_end_const<K, U>(a, pm, off + L - n, n, true, lane, me, pe);
    tps_trc_decide(a, 2u * U, r, off, L, Lt, lane, ms, ps, me, pe, row);
  }
  if (lane == 0) tps_store_row(a.rows + r, row);
}

/* ------------------------------------------------------------------------------------ K3 */
#ifndef TPS_K3_THREADS
#define TPS_K3_THREADS 128
#endif
#define TPS_K3_PSPLIT 2 /* literals of one 32-position word are split over this many threads */

/* One work item = one tile of one passing read: window starts in [tb0, tb0 + tile_bases) of
 * the oriented region z = oriented[trimfirst : min(L, maxlengthtelo)].
 *
 * Per tile: (1) stage the tile + W-base halo as oriented bit planes, (2) one match word per
 * (literal, 32 positions), stored next to (3) the running popcount of the literal's row up to
 * that word, so that (4) a window's count of a border-free literal is two 64-bit shared-memory
 * loads, two masked POPCs and a subtraction:
 *     cnt = pre[qe] + popc(row[qe] & below(be)) - pre[q0] - popc(row[q0] & below(b0))
 * with [ls, to] the window's start positions and e = to + 1.  Self-overlapping literals keep a
 * plain copy of their row for the greedy walk (exactly re.finditer's non-overlapping count).
 *
 * dynamic shared memory (words): pm[2*P*max(K,1)] | lin[3*lin_words] | ori[3*tile_words] | pad to 16 B |
 * rp[tile_words][P] (uint2 {row word, running popcount}) | brows[n_bordered*tile_words] */
template <int K>
__global__ void __launch_bounds__(TPS_K3_THREADS)
tps_window_kernel(const TpsScanArgs a, const TpsPatTable pt) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t s_item;
  constexpr uint32_t KS = K > 0 ? K : 1;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t lw = a.lin_words, tw = a.tile_words;
  uint2 *pm = reinterpret_cast<uint2 *>(smem);
  uint32_t *lin = smem + 2u * pt.n * KS;
  uint32_t *ori = lin + 3u * lw;
  uint2 *rp = reinterpret_cast<uint2 *>(ori + 3u * tw + ((4u - ((3u * lw + 3u * tw) & 3u)) & 3u)); /* 16-byte aligned */
  uint32_t *brows = reinterpret_cast<uint32_t *>(rp + (size_t)pt.n * tw);
  tps_build_pattern_masks(pm, pt, KS, tid, TPS_K3_THREADS);
  uint32_t n_pass = a.counters[0];
  if (n_pass > a.max_pass) n_pass = a.max_pass;
  const uint32_t n_items = n_pass * a.tiles_max;
  const uint32_t W = a.W, s = a.slide, t = a.trimfirst;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(a.counters + 1, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint32_t pi = item / a.tiles_max;
    const uint32_t tb0 = (item - pi * a.tiles_max) * a.tile_bases;
    const uint32_t r = a.pass_list[pi];
    const uint64_t off = a.offsets[r];
    const uint32_t L = a.lens ? a.lens[r] : (uint32_t)(a.offsets[r + 1] - off);
    const uint32_t M = L < a.maxlengthtelo ? L : a.maxlengthtelo; /* allsteps.py:263-264 */
    const uint32_t nreg = M > t ? M - t : 0u;                      /* |z|, allsteps.py:267-271 */
    const uint32_t nW = nreg >= W ? (nreg - W) / s + 1u : 0u;      /* allsteps.py:219 */
    if (tb0 >= nreg || nW == 0u) continue;
    /* windows whose start lies in [tb0, tb0 + tile_bases) */
    const uint32_t wlo = (tb0 + s - 1u) / s;
    uint32_t whi = (tb0 + a.tile_bases + s - 1u) / s;
    if (whi > nW) whi = nW;
    if (wlo >= whi) continue;
    const bool rev = a.rows[r].tail == TPS_TAIL_REVERSE;
    const uint64_t raw_off = a.rows[r].rawcount_offset;
    uint32_t *cw = a.cw + (size_t)pi * a.cw_stride;
    /* oriented positions [tb0, tb0 + tn) are staged; a window reaches W-2 past its start */
    uint32_t tn = a.tile_bases + W;
    if (tn > nreg - tb0) tn = nreg - tb0;
    /* forward: read bases [off+t+tb0, +tn); reverse: region position j is read index
     * L-1-(t+j), so the slice is read bases [off+L-t-tb0-tn, off+L-t-tb0) reversed */
    const uint64_t g0 = rev ? (off + L - t - tb0 - tn) : (off + t + tb0);
    const uint32_t phase = (uint32_t)(g0 & 15u);
    tps_stage_linear(a.pk, g0, tn, lin, lw, tid, TPS_K3_THREADS);
    __syncthreads();
    const uint32_t nq = (tn + 31u) >> 5;
    for (uint32_t q = tid; q < tw; q += TPS_K3_THREADS) {
      uint32_t p0, p1, v;
      tps_oriented_word(lin, lw, phase, tn, rev, q, p0, p1, v);
      ori[q] = p0;
      ori[tw + q] = p1;
      ori[2u * tw + q] = v;
    }
    __syncthreads();
    /* (2) match words; words past the staged slice are zero.  rp is word-major: rp[q * P + p] */
    const uint32_t P = pt.n;
    for (uint32_t i = tid; i < tw * TPS_K3_PSPLIT; i += TPS_K3_THREADS) {
      const uint32_t h = i / tw, q = i - h * tw;
      if (q < nq) {
        TpsWin<K> win;
        tps_win_init<K>(win, ori[q], ori[q + 1], ori[tw + q], ori[tw + q + 1], ori[2u * tw + q],
                        ori[2u * tw + q + 1]);
        for (uint32_t p = h; p < P; p += TPS_K3_PSPLIT) rp[q * P + p].x = tps_win_match<K>(win, pm, pt, p);
      } else {
        for (uint32_t p = h; p < P; p += TPS_K3_PSPLIT) rp[q * P + p].x = 0u;
      }
    }
    __syncthreads();
    /* plain copies of the rows of self-overlapping literals for the greedy walk */
    for (uint64_t bm = pt.bordered_mask; bm; bm &= bm - 1) {
      const uint32_t p = (uint32_t)__ffsll((long long)bm) - 1u;
      for (uint32_t q = tid; q < tw; q += TPS_K3_THREADS) brows[pt.brow[p] * tw + q] = rp[q * P + p].x;
    }
    /* (3) exclusive running popcount per literal row: one warp per literal, each lane sums a run of
     * consecutive words, one warp scan joins the runs */
    {
      const uint32_t per = (tw + 31u) / 32u; /* words per lane */
      for (uint32_t p = warp; p < P; p += TPS_K3_THREADS / 32) {
        const uint32_t qa = lane * per;
        uint32_t run = 0u;
        for (uint32_t j = 0; j < per; ++j) {
          const uint32_t q = qa + j;
          if (q < tw) run += tps_popc32(rp[q * P + p].x);
        }
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t up = __shfl_up_sync(TPS_FULL, inc, o);
          if ((int)lane >= o) inc += up;
        }
        uint32_t acc = inc - run;
        for (uint32_t j = 0; j < per; ++j) {
          const uint32_t q = qa + j;
          if (q < tw) {
            uint2 *e = rp + q * P + p;
            const uint32_t x = e->x;
            e->y = acc;
            acc += tps_popc32(x);
          }
        }
      }
    }
    __syncthreads();
    /* (4) windows */
    const bool want_raw = raw_off != ~0ull;
    for (uint32_t w = wlo + tid; w < whi; w += TPS_K3_THREADS) {
      const uint32_t ls = w * s - tb0;
      uint32_t c = 0;
      uint8_t *rawp = want_raw ? a.raw + raw_off + (uint64_t)w * P : nullptr;
      if constexpr (K > 0) {
        if (W - 1u >= (uint32_t)K) {
          const uint32_t e = ls + (W - (uint32_t)K); /* one past the last start position */
          const uint2 *r0 = rp + (ls >> 5) * P, *re = rp + (e >> 5) * P;
          const uint32_t m0 = (1u << (ls & 31u)) - 1u, me = (1u << (e & 31u)) - 1u;
          /* occurrences of every literal (branch-free); P even: two literals per 128-bit load */
          if ((P & 3u) == 0u) {
            const uint4 *v0 = reinterpret_cast<const uint4 *>(r0), *ve = reinterpret_cast<const uint4 *>(re);
            for (uint32_t p = 0; p < P; p += 4u) {
              const uint4 a0 = v0[p >> 1], ae = ve[p >> 1], b0 = v0[(p >> 1) + 1u], be = ve[(p >> 1) + 1u];
              uint32_t c0 = ae.y + tps_popc32(ae.x & me) - a0.y - tps_popc32(a0.x & m0);
              uint32_t c1 = ae.w + tps_popc32(ae.z & me) - a0.w - tps_popc32(a0.z & m0);
              uint32_t c2 = be.y + tps_popc32(be.x & me) - b0.y - tps_popc32(b0.x & m0);
              uint32_t c3 = be.w + tps_popc32(be.z & me) - b0.w - tps_popc32(b0.z & m0);
              c0 = c0 ? c0 : 1u; c1 = c1 ? c1 : 1u; c2 = c2 ? c2 : 1u; c3 = c3 ? c3 : 1u; /* `... or 1` */
              c += (c0 + c1) + (c2 + c3);
              if (want_raw) *reinterpret_cast<uint32_t *>(rawp + p) = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
            }
          } else {
            for (uint32_t p = 0; p < P; ++p) {
              const uint2 x0 = r0[p], xe = re[p];
              uint32_t cnt = xe.y + tps_popc32(xe.x & me) - x0.y - tps_popc32(x0.x & m0);
              cnt = cnt ? cnt : 1u;
              c += cnt;
              if (want_raw) rawp[p] = (uint8_t)cnt;
            }
          }
          /* self-overlapping literals: replace the occurrence count by the greedy one */
          for (uint64_t bm = pt.bordered_mask; bm; bm &= bm - 1) {
            const uint32_t p = (uint32_t)__ffsll((long long)bm) - 1u;
            const uint2 x0 = r0[p], xe = re[p];
            uint32_t occ = xe.y + tps_popc32(xe.x & me) - x0.y - tps_popc32(x0.x & m0);
            uint32_t g = tps_greedy_count(brows + pt.brow[p] * tw, (int32_t)ls, (int32_t)e - 1, (uint32_t)K);
            occ = occ ? occ : 1u;
            g = g ? g : 1u;
            c = c - occ + g;
            if (want_raw) rawp[p] = (uint8_t)g;
          }
        } else {
          c = P; /* window text shorter than the literals: every count is floored to 1 */
          if (want_raw)
            for (uint32_t p = 0; p < P; ++p) rawp[p] = 1u;
        }
      } else {
        for (uint32_t p = 0; p < P; ++p) {
          const uint32_t k = pt.len[p];
          uint32_t cnt = 0;
          if (W - 1u >= k) {
            const uint32_t e = ls + (W - k);
            if (pt.bordered[p]) {
              cnt = tps_greedy_count(brows + pt.brow[p] * tw, (int32_t)ls, (int32_t)e - 1, k);
            } else {
              const uint2 x0 = rp[(ls >> 5) * P + p], xe = rp[(e >> 5) * P + p];
              cnt = xe.y + tps_popc32(xe.x & ((1u << (e & 31u)) - 1u)) - x0.y -
                    tps_popc32(x0.x & ((1u << (ls & 31u)) - 1u));
            }
          }
          cnt = cnt ? cnt : 1u;
          c += cnt;
          if (want_raw) rawp[p] = (uint8_t)cnt;
        }
      }
      cw[w] = c;
    }
  }
}

/* ------------------------------------------------------------------------------------ K4 */
#define TPS_K4_THREADS 128

/* Single change point of c_w[0..n) by one CTA of TPS_K4_THREADS threads, the exact form of ruptures
 * Binseg(model="l2", jump=5, min_size=2).predict(n_bkps=1):
 * argmax over b in {5,10,...}, 2 <= b <= n-2 of (n*S_b - b*T)^2 / (b*(n-b)), ties -> larger b.
 * c_w comes through a reader whose elements are single windows (G = 1) or the sums of the groups of five
 * windows [5j, 5j+5) (G = 5: the candidates are the multiples of 5, so group sums are all that is needed).
 * Every thread owns one run of consecutive groups, so every candidate's S_b is a running sum inside one thread:
 *   pass 1  sums the run; one block-wide exclusive scan turns the run sums into S at the run starts and T;
 *   pass 2  walks the run's candidates with a float32 SCREEN of the gain (relative error < 1e-6) and keeps the
 *           thread's largest; a block-wide max gives gtop;
 *   pass 3  only threads whose screen value reaches gtop * (1 - 1e-4) walk their run again and put every candidate
 *           above that bar through the exact comparator (float64 cross-multiplication, 128-bit integers for
 *           ties and near-ties); their winners go to a short shared list that thread 0 reduces.
 * Any candidate that can be the exact argmax passes the screen, so the result is the exact one; the screen only
 * spares the other ~99 % of the candidates the 64-bit arithmetic.  A constant signal (every gain 0) sends every
 * candidate to pass 3, which is the old cost.  c_w comes through a reader (global rows are read with
 * ld.global.cg: in the fused kernel other CTAs wrote them).  Result valid in thread 0. */
struct TpsCpShared {
  uint64_t wsum[TPS_K4_THREADS / 32];
  float wmax[TPS_K4_THREADS / 32];
  uint32_t n_cand;
  tps_cand cand[TPS_K4_THREADS];
};

/* c_w readers: uint32 in global memory (tps_window_kernel's rows), uint16 in global memory (rows of the
 * bit-parallel kernel that do not fit its scratch) or uint16 staged in shared memory */
struct TpsCwGlobal32 {
  const uint32_t *p;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return __ldcg(p + i); }
};
struct TpsCwGlobal16 {
  const uint16_t *p;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return __ldcg(p + i); }
};
struct TpsCwShared16 {
  const uint16_t *p;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return p[i]; }
};
/* sum of the reader elements that make up one group of five windows: five elements of one window each
 * (G = 1) or one element that already is the group sum (G = 5) */
template <int G, class RD>
__device__ __forceinline__ uint32_t tps_cw5(const RD &cw, uint32_t i) {
  if constexpr (G == 5) return cw(i);
  else return cw(i) + cw(i + 1u) + cw(i + 2u) + cw(i + 3u) + cw(i + 4u);
}

__device__ __forceinline__ float tps_gain_screen(uint32_t n, uint64_t S, uint64_t T, uint32_t b) {
  const float d = (float)((int64_t)((uint64_t)n * S) - (int64_t)((uint64_t)b * T));
  return d * d * __frcp_rn((float)b * (float)(n - b));
}

template <int G, class RD>
__device__ __forceinline__ int32_t tps_changepoint_block(const RD cw, uint32_t n, TpsCpShared &sh, uint32_t tid) {
  static_assert(G == 1 || G == 5, "reader elements are windows or groups of five windows");
  constexpr uint32_t STEP = 5 / G; /* elements per candidate */
  constexpr uint32_t NWARP = TPS_K4_THREADS / 32;
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  const uint32_t n_el = (n + G - 1u) / G; /* reader elements */
  const uint32_t chunk = STEP * ((n + 5u * TPS_K4_THREADS - 1u) / (5u * TPS_K4_THREADS)); /* elements per thread */
  const uint32_t w0 = tid * chunk < n_el ? tid * chunk : n_el;
  const uint32_t w1 = n_el - w0 < chunk ? n_el : w0 + chunk;
  uint64_t sum = 0;
  {
    uint32_t w = w0;
    for (; w + STEP <= w1; w += STEP) sum += tps_cw5<G>(cw, w);
    for (; w < w1; ++w) sum += cw(w);
  }
  uint64_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t up = __shfl_up_sync(TPS_FULL, inc, o);
    if ((int)lane >= o) inc += up;
  }
  if (lane == 31u) sh.wsum[warp] = inc;
  if (tid == 0) sh.n_cand = 0u;
  __syncthreads();
  uint64_t S0 = inc - sum, T = 0; /* S0 = sum_{w < w0} c_w */
#pragma unroll
  for (uint32_t i = 0; i < NWARP; ++i) {
    if (i < warp) S0 += sh.wsum[i];
    T += sh.wsum[i];
  }
  /* pass 2: float32 screen */
  float gmax = -1.0f;
  {
    uint64_t S = S0;
    for (uint32_t e = w0; e < w1; e += STEP) {
      const uint32_t b = e * G;
      if (b >= 2u && n - b >= 2u) gmax = fmaxf(gmax, tps_gain_screen(n, S, T, b));
      if (e + STEP <= w1) S += tps_cw5<G>(cw, e);
    }
  }
  float gtop = gmax;
#pragma unroll
  for (int o = 16; o; o >>= 1) gtop = fmaxf(gtop, __shfl_xor_sync(TPS_FULL, gtop, o));
  if (lane == 0) sh.wmax[warp] = gtop;
  __syncthreads();
#pragma unroll
  for (uint32_t i = 0; i < NWARP; ++i) gtop = fmaxf(gtop, sh.wmax[i]);
  const float bar = gtop * (1.0f - 1e-4f); /* gtop >= 0 whenever a candidate exists */
  /* pass 3: exact comparison of the candidates at the top */
  if (gmax >= bar && gmax >= 0.0f) {
    tps_cand best;
    best.b = -1; best.d = 0; best.den = 1; best.num_f = 0.0; best.den_f = 1.0;
    uint64_t S = S0;
    for (uint32_t e = w0; e < w1; e += STEP) {
      const uint32_t b = e * G;
      if (b >= 2u && n - b >= 2u && tps_gain_screen(n, S, T, b) >= bar) {
        const tps_cand c = tps_make_cand(n, S, T, b);
        if (tps_cand_better(&best, &c)) best = c;
      }
      if (e + STEP <= w1) S += tps_cw5<G>(cw, e);
    }
    if (best.b >= 0) sh.cand[atomicAdd(&sh.n_cand, 1u)] = best;
  }
  __syncthreads();
  int32_t best_b = -1;
  if (tid == 0) {
    const uint32_t nc = sh.n_cand;
    if (nc) {
      tps_cand best = sh.cand[0];
      for (uint32_t i = 1; i < nc; ++i) {
        const tps_cand oth = sh.cand[i];
        if (tps_cand_better(&best, &oth)) best = oth;
      }
      best_b = best.b;
    }
  }
  return best_b;
}

/* Thread 0 stores the change point of read r (x[bkp] = trimfirst + slide * b, allsteps.py:304,312), or marks the
 * read TPS_ST_BADSEG: ruptures' sanity_check needs n >= 7 for jump 5, min_size 2, one breakpoint. */
__device__ __forceinline__ void tps_store_changepoint(const TpsScanArgs &a, uint32_t r, int32_t best_b) {
  tps_row *row = a.rows + r;
  if (best_b >= 0) {
    row->bkp = best_b;
    row->telo_length = (int32_t)(a.trimfirst + a.slide * (uint32_t)best_b);
  } else {
    row->status = TPS_ST_BADSEG;
  }
}

/* Stand-alone K4 (behind tps_window_kernel): persistent CTAs, one passing read at a time. */
__global__ void __launch_bounds__(TPS_K4_THREADS)
tps_changepoint_kernel(const TpsScanArgs a) {
  __shared__ TpsCpShared s_cp;
  __shared__ uint32_t s_pi;
  const uint32_t tid = threadIdx.x;
  uint32_t n_pass = a.counters[0];
  if (n_pass > a.max_pass) n_pass = a.max_pass;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_pi = atomicAdd(a.counters + 5, 1u);
    __syncthreads();
    const uint32_t pi = s_pi;
    if (pi >= n_pass) break;
    const uint32_t r = a.pass_list[pi];
    const uint32_t nW = a.rows[r].n_windows;
    int32_t best_b = -1;
    if (nW >= 7u) best_b = tps_changepoint_block<1>(TpsCwGlobal32{a.cw + (size_t)pi * a.cw_stride}, nW, s_cp, tid);
    if (tid == 0) tps_store_changepoint(a, r, best_b);
  }
}

/* ------------------------------------------------------------------- K3 + K4, bit-parallel (default)
 * tps_window_bp_kernel<K>: the window counts of step 2 without a per-(window, literal) loop, and the change
 * point of a read by the CTA that finishes its last tile.
 *
 *   c_w = sum_p max(cnt_p(w), 1) = occ_U(w) + #{p : cnt_p(w) == 0}         (allsteps.py:279-291, `... or 1`)
 *
 * The literals of one scan are distinct strings of one length K, so at most one of them matches at a position:
 * the sum of the occurrence counts is the number of set bits of the UNION row U = OR_p m_p inside the window's
 * D = W - K start positions, one prefix-popcount difference per window instead of P.  Whether literal p occurs in
 * a window at all is bit j of the dilated row R_p[j] = OR_{d<D} m_p[j+d]; with S = "bits at or below the highest
 * set bit" and Pf = "bits above the lowest set bit" of a match word (one FLO / one negate each),
 * x32[q] = S[q] | Pf[q+1] is the 32-dilation, and R[q] = x32[q] | .. | x32[q+a-1] | funnel(x32[q+a-1], x32[q+a], r)
 * for D = 32a + r.  The P rows R_p are added bit-sliced (carry-save adders, 2 LOP3 each) into `nz` count planes,
 * 32 positions per instruction; a window reads its count of present literals off the planes at its start bit.
 * Self-overlapping literals (greedy != occurrences): CF marks starts that have another start of the same literal
 * less than K ahead; a window without such a start inside is exact as it stands, the few others redo those
 * literals with the greedy walk.
 *
 * Work item = one passing read, listed by K2: a CTA walks the read's tiles one after the other (a tile is
 * <= 32 * TPS_K3N_THREADS oriented positions from the start of its first window to the end of its last, one
 * 32-position word per thread; tiles are balanced and hold whole groups of five windows), keeps the sums of c_w
 * over the groups of five windows [5j, 5j+5) -- all the change point reads: its candidates are the multiples of
 * 5 -- in shared memory and finishes with tps_changepoint_block: no second launch, no trip through global memory
 * between the window counts and the change point, and the change points overlap the window counting of the
 * CTA's neighbours.  The code words of the next tile and the record of the next read are prefetched (cp.async).
 * With few passing reads (at most twice as many as CTAs) a work unit is every second or fourth tile of a read: the
 * group sums of the parts meet in a global uint16 row per read (gs_rows), tile_done counts the finished parts, and
 * the CTA that completes a read fetches the row and finds the change point.
 *
 * dynamic shared memory (words): raw[2][TPS_K3N_RAW_WORDS] | pm[2PK] | lin[3*lin_words] | ori[3*(NT+1)] | pad to
 * 16 B | Z[NT+1] uint4 | Zhi[NT+1] uint4 (nz > 4) | UP[NT+2] uint2 | CP[NT+2] uint2 (bordered) | SP[P4][sp_stride]
 * uint2 {S, Pf}, P4 = P rounded up to a multiple of 4 | brows[n_bordered][NT+1] | gsum[gs_cap] uint16 */
#define TPS_K3N_THREADS 128

#define TPS_CSA(h, l, x, y, z)                       \
  do {                                               \
    const uint32_t x_ = (x), y_ = (y), z_ = (z);     \
    (h) = (x_ & y_) | (x_ & z_) | (y_ & z_);         \
    (l) = x_ ^ y_ ^ z_;                              \
  } while (0)

__device__ __forceinline__ uint32_t tps_block_excl_scan(uint32_t v, uint32_t *wt, uint32_t tid) {
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(TPS_FULL, inc, o);
    if ((int)lane >= o) inc += up;
  }
  if (lane == 31u) wt[warp] = inc;
  __syncthreads();
  uint32_t before = inc - v;
#pragma unroll
  for (uint32_t i = 0; i < TPS_K3N_THREADS / 32; ++i)
    if (i < warp) before += wt[i];
  return before;
}

/* R[q] of one literal from its {S, Pf} slots q, q+1, ...: MODE 1: D = 32 + dr (dr > 0), MODE 2: D = 32 da,
 * MODE 3: D = 96, MODE 0: any D >= 32 */
template <int MODE>
__device__ __forceinline__ uint32_t tps_dilate(const uint2 *__restrict__ sp, uint32_t da, uint32_t dr) {
  if constexpr (MODE == 1) {
    const uint2 s0 = sp[0], s1 = sp[1], s2 = sp[2];
    const uint32_t x0 = s0.x | s1.y, x1 = s1.x | s2.y;
    return x0 | __funnelshift_r(x0, x1, dr);
  } else if constexpr (MODE == 3) {
    const uint2 s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3];
    return (s0.x | s1.x | s1.y) | (s2.x | s2.y | s3.y);
  } else if constexpr (MODE == 2) {
    uint2 v = sp[0];
    uint32_t acc = v.x;
    for (uint32_t j = 1; j < da; ++j) {
      v = sp[j];
      acc |= v.x | v.y;
    }
    return acc | sp[da].y;
  } else {
    uint2 s0 = sp[0], s1 = sp[1];
    uint32_t xi = s0.x | s1.y, acc = xi;
    for (uint32_t j = 1; j < da; ++j) {
      s0 = s1;
      s1 = sp[j + 1u];
      xi = s0.x | s1.y;
      acc |= xi;
    }
    if (dr) {
      s0 = s1;
      s1 = sp[da + 1u];
      acc |= __funnelshift_r(xi, s0.x | s1.y, dr);
    }
    return acc;
  }
}

/* count planes of word q: the P presence rows added bit-sliced, four at a time */
template <int MODE>
__device__ __forceinline__ void tps_presence_planes(const uint2 *__restrict__ sp, uint32_t P4, uint32_t stride,
                                                    uint32_t da, uint32_t dr, uint32_t nz, uint4 *zlo, uint4 *zhi) {
  uint32_t c1 = 0u, c2 = 0u, c4 = 0u, c8 = 0u, c16 = 0u, c32 = 0u, c64 = 0u;
  for (uint32_t pb = 0; pb < P4; pb += 4u) {
    const uint32_t r0 = tps_dilate<MODE>(sp, da, dr);
    const uint32_t r1 = tps_dilate<MODE>(sp + stride, da, dr);
    const uint32_t r2 = tps_dilate<MODE>(sp + 2u * stride, da, dr);
    const uint32_t r3 = tps_dilate<MODE>(sp + 3u * stride, da, dr);
    sp += 4u * stride;
    uint32_t ta, tb, fa;
    TPS_CSA(ta, c1, c1, r0, r1);
    TPS_CSA(tb, c1, c1, r2, r3);
    TPS_CSA(fa, c2, c2, ta, tb);
    uint32_t cy = c4 & fa; /* ripple the fours */
    c4 ^= fa;
    if (nz > 4u) {
      uint32_t t = c8 & cy;
      c8 ^= cy;
      cy = c16 & t; c16 ^= t;
      t = c32 & cy; c32 ^= cy;
      c64 ^= t;
    } else {
      c8 ^= cy;
    }
  }
  *zlo = make_uint4(c1, c2, c4, c8);
  if (nz > 4u) *zhi = make_uint4(c16, c32, c64, 0u);
}

__device__ __forceinline__ void tps_cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tps_smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void tps_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

/* raw code / flag words of one tile, prefetched into shared memory one tile ahead */
#define TPS_K3N_RAW_CODE_WORDS 272 /* 258 groups of a 4096-position slice at any phase + alignment to 16 bytes */
#define TPS_K3N_RAW_FLAG_WORDS 16
#define TPS_K3N_RAW_WORDS (TPS_K3N_RAW_CODE_WORDS + TPS_K3N_RAW_FLAG_WORDS)

/* issue the asynchronous copies of the code and flag words that cover bases [g0, g0 + n): code words from the
 * 16-byte boundary at or below group g0 / 16, flag words from the 16-byte boundary at or below its flag word */
__device__ __forceinline__ void tps_prefetch_codes(const TpsPacked &pk, uint64_t g0, uint32_t n, uint32_t *raw,
                                                   uint32_t tid) {
  const uint64_t gfirst = g0 >> 4;
  const uint32_t ng = (uint32_t)(((g0 + n + 15u) >> 4) - gfirst); /* groups */
  const uint64_t w0 = gfirst & ~3ull;
  const uint32_t nchunk = (ng + (uint32_t)(gfirst & 3ull) + 3u) >> 2;
  if (tid < nchunk) tps_cp_async16(raw + 4u * tid, pk.codes + w0 + 4u * tid);
  const uint64_t f0 = (gfirst >> 5) & ~3ull;
  const uint32_t nf = ((ng + (uint32_t)(gfirst & 127ull) + 31u) >> 5) + 3u >> 2;
  if (tid >= 96u && tid - 96u < nf) tps_cp_async16(raw + TPS_K3N_RAW_CODE_WORDS + 4u * (tid - 96u), pk.flags + f0 + 4u * (tid - 96u));
}

/* tps_stage_linear from the prefetched words: entry e of the uint16 views holds group gfirst + e - 2 */
__device__ __forceinline__ void tps_stage_entry_raw(const uint32_t *raw, const uint8_t *__restrict__ bases, uint64_t gfirst,
                                                    uint32_t ng, uint32_t e, uint16_t *l0, uint16_t *l1, uint16_t *lv) {
  uint32_t y = 0u, v = 0u;
  const uint32_t gi = e - 2u;
  if (gi < ng) { /* e < 2 wraps around: pad */
    y = tps_linear_planes(raw[(uint32_t)(gfirst & 3ull) + gi]);
    v = 0xFFFFu;
    const uint32_t fb = (uint32_t)(gfirst & 127ull) + gi; /* flag bit relative to the first prefetched flag word */
    if ((raw[TPS_K3N_RAW_CODE_WORDS + (fb >> 5)] >> (fb & 31u)) & 1u) { /* rare: N, IUPAC */
      const uint4 b = __ldg(reinterpret_cast<const uint4 *>(bases) + (gfirst + gi));
      v = tps_exact_mask16_simd(b.x, b.y, b.z, b.w);
    }
  }
  l0[e] = (uint16_t)(y & 0xFFFFu);
  l1[e] = (uint16_t)(y >> 16);
  lv[e] = (uint16_t)v;
}

/* one window of the bit-parallel kernel on its own: c_w from the union row, the count planes and, where a
 * self-overlapping literal has two starts less than K apart inside the window, the greedy walk */
template <int K>
__device__ __forceinline__ uint32_t tps_bp_window(uint32_t ls, uint32_t D, uint32_t P, uint32_t nz, uint32_t nb,
                                                  const TpsPatTable &pt, const uint2 *UP, const uint4 *Z, const uint4 *Zhi,
                                                  const uint2 *CP, const uint32_t *brows, uint32_t brow_stride) {
  const uint32_t e = ls + D; /* start positions [ls, e) */
  const uint32_t q0 = ls >> 5, qe = e >> 5, b0 = ls & 31u;
  const uint32_t m0 = (1u << b0) - 1u, me = (1u << (e & 31u)) - 1u;
  const uint2 u0 = UP[q0], ue = UP[qe];
  uint32_t c = ue.y + tps_popc32(ue.x & me) - u0.y - tps_popc32(u0.x & m0); /* occurrences of all literals */
  const uint4 z = Z[q0];
  uint32_t present = ((z.x >> b0) & 1u) | (((z.y >> b0) & 1u) << 1) | (((z.z >> b0) & 1u) << 2) | (((z.w >> b0) & 1u) << 3);
  if (nz > 4u) {
    const uint4 zh = Zhi[q0];
    present |= (((zh.x >> b0) & 1u) << 4) | (((zh.y >> b0) & 1u) << 5) | (((zh.z >> b0) & 1u) << 6);
  }
  c += P - present; /* absent literals count 1 each */
  if (nb) {
    const uint2 f0 = CP[q0], fe = CP[qe];
    if (fe.y + tps_popc32(fe.x & me) - f0.y - tps_popc32(f0.x & m0)) {
      for (uint64_t bm = pt.bordered_mask; bm; bm &= bm - 1) {
        const uint32_t p = (uint32_t)__ffsll((long long)bm) - 1u;
        const uint32_t *row = brows + pt.brow[p] * brow_stride;
        uint32_t occ = tps_range_popcount(row, (int32_t)ls, (int32_t)e - 1);
        uint32_t g = tps_greedy_count(row, (int32_t)ls, (int32_t)e - 1, (uint32_t)K);
        occ = occ ? occ : 1u;
        g = g ? g : 1u;
        c = c - occ + g;
      }
    }
  }
  return c;
}

template <int K>
#ifdef TPS_K3N_MINB
__global__ void __launch_bounds__(TPS_K3N_THREADS, TPS_K3N_MINB)
#else
__global__ void __launch_bounds__(TPS_K3N_THREADS)
#endif
tps_window_bp_kernel(const TpsScanArgs a, const TpsPatTable pt) {
  static_assert(K > 0, "the bit-parallel window kernel needs a common literal length");
  constexpr uint32_t NT = TPS_K3N_THREADS;
  extern __shared__ __align__(16) uint32_t smem[];
  /* software pipeline: while read i is processed its CTA holds the record of read i+1 (cp.async, issued at the
   * start of read i) and the index of read i+2 (atomic, issued at the start of read i, stored at its end); every
   * tile issues the copy of the next tile's code words -- the next tile of the read or tile 0 of read i+1 -- and
   * collects it at its end.  No step of the loop waits for global memory. */
  __shared__ TpsReadItem s_item[2];
  __shared__ uint32_t s_idx[3];
  __shared__ uint32_t s_last;
  __shared__ uint32_t s_wt[2][NT / 32];
  __shared__ TpsCpShared s_cp;
  const uint32_t tid = threadIdx.x, q = threadIdx.x;
  const uint32_t P = pt.n, P4 = (P + 3u) & ~3u, lw = a.lin_words, stride = a.tile_words, nb = pt.n_bordered;
  const uint32_t W = a.W, s = a.slide;
  const uint32_t D = W - (uint32_t)K, da = D >> 5, dr = D & 31u, nz = a.nz;
  const uint32_t mode = (da == 1u && dr) ? 1u : (dr ? 0u : (da == 3u ? 3u : 2u));
  /* five-window fast path: the five starts of a group lie within 32 positions of the first */
  const bool grouped = 4u * s < 32u && !a.no_groups;
  const uint32_t gm = 1u | (1u << s) | (1u << (2u * s)) | (1u << (3u * s)) | (1u << (4u * s)); /* the five starts */
  const uint32_t lm1 = (1u << s) - 1u, lm2 = (1u << (2u * s)) - 1u, lm3 = (1u << (3u * s)) - 1u, lm4 = (1u << (4u * s)) - 1u;
  uint32_t *raw = smem; /* two buffers of TPS_K3N_RAW_WORDS */
  uint2 *pm = reinterpret_cast<uint2 *>(smem + 2u * TPS_K3N_RAW_WORDS);
  uint32_t *lin = smem + 2u * TPS_K3N_RAW_WORDS + 2u * P * K;
  uint32_t *ori = lin + 3u * lw;
  uint32_t *al = ori + 3u * (NT + 1u);
  al += (4u - ((uint32_t)(al - smem) & 3u)) & 3u;
  uint4 *Z = reinterpret_cast<uint4 *>(al);            /* NT + 1 entries each: the fast path reads word q0 + 1 */
  uint4 *Zhi = Z + (NT + 1u);
  uint2 *UP = reinterpret_cast<uint2 *>(Zhi + (nz > 4u ? NT + 1u : 0u));
  uint2 *CP = UP + (NT + 2u);
  uint2 *SP = CP + (nb ? NT + 2u : 0u);
  uint32_t *brows = reinterpret_cast<uint32_t *>(SP + (size_t)P4 * stride);
  /* group sums of the read in hand (16-byte aligned: the split mode fills them with cp.async) */
  uint16_t *gsum = reinterpret_cast<uint16_t *>((reinterpret_cast<uintptr_t>(brows + nb * (NT + 1u)) + 15u) & ~(uintptr_t)15u);
  tps_build_pattern_masks(pm, pt, K, tid, NT);
  const uint32_t n_items = a.counters[6];
  /* words that stay zero for the whole kernel: the pad word behind the oriented planes and the per-word tables,
   * the slots of SP behind the tile and of the rows that pad P to a multiple of 4, the pad word of the plain rows */
  if (tid < 3u) ori[tid * (NT + 1u) + NT] = 0u;
  if (tid == 3u) Z[NT] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 4u && nz > 4u) Zhi[NT] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 5u) UP[NT] = make_uint2(0u, 0u);
  if (tid == 6u && nb) CP[NT] = make_uint2(0u, 0u);
  for (uint32_t i = tid; i < P4 * stride; i += NT) SP[i] = make_uint2(0u, 0u);
  for (uint32_t i = tid; i < nb * (NT + 1u); i += NT) brows[i] = 0u;

  /* tile t of a read with n_windows windows: windows [t * per, min(n_windows, (t + 1) * per)), balanced, whole
   * groups of five; staged positions from the start of its first window to the end of its last */
  const uint32_t gmax = a.bp_tile_windows / 5u;
  auto tiles_of = [&](uint32_t nW, uint32_t &per) {
    const uint32_t ng = (nW + 4u) / 5u, nt = (ng + gmax - 1u) / gmax;
    per = (ng + nt - 1u) / nt * 5u;
    return nt;
  };
  auto prefetch_tile = [&](const TpsReadItem &it, uint32_t per, uint32_t t, uint32_t *buf) {
    const uint32_t wlo = t * per, whi = wlo + per < it.n_windows ? wlo + per : it.n_windows;
    const uint32_t tb0 = wlo * s, tn = (whi - wlo - 1u) * s + W;
    const uint64_t g0 = it.rev ? it.g_edge - tb0 - tn : it.g_edge + tb0;
    tps_prefetch_codes(a.pk, g0, tn, buf, tid);
  };

  /* A work unit is a read, or -- when the batch holds at most twice (once) as many passing reads as the grid has
   * CTAs, so that whole reads would leave CTAs idle or the last ones alone at the end -- every second (fourth) tile
   * of a read: unit u = (read u / parts, tiles t = u % parts, + parts, ...). */
  const uint32_t parts = a.no_split ? 1u : (n_items <= gridDim.x ? 4u : (n_items <= 2u * gridDim.x ? 2u : 1u));
  const uint32_t n_units = n_items * parts;
  bool have = false; /* the code words of this CTA's next tile are in raw[buf] */
  if (tid == 0) {
    const uint32_t base = atomicAdd(a.counters + 1, 2u);
    s_idx[0] = base; s_idx[1] = base + 1u;
    if (base < n_units) s_item[0] = a.items[base / parts];
  }
  __syncthreads();
  if (s_idx[0] < n_units) {
    uint32_t per0;
    const uint32_t nt0 = tiles_of(s_item[0].n_windows, per0), part0 = s_idx[0] % parts;
    if (part0 < nt0) {
      prefetch_tile(s_item[0], per0, part0, raw);
      have = true;
    }
  }
  tps_cp_async_wait_all();
  __syncthreads();
  uint32_t buf = 0u; /* raw buffer that holds the tile in hand */

  for (uint32_t i = 0;; ++i) {
    const uint32_t unit = s_idx[i % 3u];
    if (unit >= n_units) break;
    const TpsReadItem it = s_item[i & 1u];
    const uint32_t part = unit % parts;
    const uint32_t idx_next = s_idx[(i + 1u) % 3u];
    uint32_t idx_next2 = 0u;
    if (tid == 0) { /* record of unit i+1, index of unit i+2 */
      if (idx_next < n_units) {
        const TpsReadItem *src = a.items + idx_next / parts;
        tps_cp_async16(&s_item[(i + 1u) & 1u], src);
        tps_cp_async16(reinterpret_cast<uint8_t *>(&s_item[(i + 1u) & 1u]) + 16, reinterpret_cast<const uint8_t *>(src) + 16);
      }
      idx_next2 = atomicAdd(a.counters + 1, 1u);
    }
    const uint32_t nW = it.n_windows;
    uint32_t per;
    const uint32_t ntiles = tiles_of(nW, per);
    const bool rev = it.rev != 0u;
    if (part < ntiles && !have) { /* the unit before had nothing to prefetch for (an empty unit): fetch now */
      prefetch_tile(it, per, part, raw + buf * TPS_K3N_RAW_WORDS);
      tps_cp_async_wait_all();
      __syncthreads();
    }
    have = false;
    uint16_t *gdst = parts == 1u ? gsum : a.gs_rows + (size_t)it.slot * a.gs_cap;

    for (uint32_t t = part; t < ntiles; t += parts, buf ^= 1u) {
      const uint32_t wlo = t * per, whi = wlo + per < nW ? wlo + per : nW;
      const uint32_t tb0 = wlo * s, tn = (whi - wlo - 1u) * s + W;
      const uint64_t g0 = rev ? it.g_edge - tb0 - tn : it.g_edge + tb0;
      const uint32_t *rawk = raw + buf * TPS_K3N_RAW_WORDS;
      /* code words of the next tile: of this unit, or the first tile of the next unit (its record was asked for at
       * the start of this unit and has arrived by the end of any tile; a unit's first tile waits for it here) */
      if (t + parts < ntiles) {
        prefetch_tile(it, per, t + parts, raw + (buf ^ 1u) * TPS_K3N_RAW_WORDS);
      } else if (idx_next < n_units) {
        if (t == part) {
          tps_cp_async_wait_all();
          __syncthreads();
        }
        const TpsReadItem &nx = s_item[(i + 1u) & 1u];
        uint32_t pern;
        const uint32_t ntn = tiles_of(nx.n_windows, pern), partn = idx_next % parts;
        if (partn < ntn) {
          prefetch_tile(nx, pern, partn, raw + (buf ^ 1u) * TPS_K3N_RAW_WORDS);
          have = true;
        }
      }
      const uint32_t phase = (uint32_t)(g0 & 15u);
      {
        uint16_t *l0 = reinterpret_cast<uint16_t *>(lin), *l1 = reinterpret_cast<uint16_t *>(lin + lw),
                 *lv = reinterpret_cast<uint16_t *>(lin + 2 * lw);
        const uint64_t gfirst = g0 >> 4;
        const uint32_t ng = (uint32_t)(((g0 + tn + 15u) >> 4) - gfirst);
        tps_stage_entry_raw(rawk, a.pk.bases, gfirst, ng, tid, l0, l1, lv);
        tps_stage_entry_raw(rawk, a.pk.bases, gfirst, ng, tid + NT, l0, l1, lv);
        if (tid < 2u * lw - 2u * NT) tps_stage_entry_raw(rawk, a.pk.bases, gfirst, ng, tid + 2u * NT, l0, l1, lv);
      }
      __syncthreads();
      {
        uint32_t p0, p1, v;
        tps_oriented_word(lin, lw, phase, tn, rev, q, p0, p1, v);
        ori[q] = p0;
        ori[(NT + 1u) + q] = p1;
        ori[2u * (NT + 1u) + q] = v;
      }
      __syncthreads();
      /* (1) match words of word q: union, {S, Pf} halves of the 32-dilation, plain rows of bordered literals */
      uint32_t U = 0u;
      {
        TpsWin<K> win;
        tps_win_init<K>(win, ori[q], ori[q + 1u], ori[(NT + 1u) + q], ori[(NT + 1u) + q + 1u], ori[2u * (NT + 1u) + q],
                        ori[2u * (NT + 1u) + q + 1u]);
        auto halves = [](uint32_t m) { /* bits at or below the highest match | bits above the lowest match */
          return make_uint2(__funnelshift_rc(TPS_FULL, 0u, (uint32_t)__clz((int)m)), (m | (0u - m)) << 1);
        };
        uint2 *sp = SP + q;
        if (pt.paired) {
          const uint32_t H = P >> 1;
          uint2 *spc = sp + (size_t)H * stride;
          for (uint32_t p = 0; p < H; ++p, sp += stride, spc += stride) {
            uint32_t m, mc;
            tps_win_match_pair<K>(win, pm, p, m, mc);
            U |= m | mc;
            *sp = halves(m);
            *spc = halves(mc);
          }
        } else {
          for (uint32_t p = 0; p < P; ++p, sp += stride) {
            const uint32_t m = tps_win_match<K>(win, pm, pt, p);
            U |= m;
            *sp = halves(m);
          }
        }
        for (uint64_t bm = pt.bordered_mask; bm; bm &= bm - 1) { /* uniform; none for most pattern sets */
          const uint32_t p = (uint32_t)__ffsll((long long)bm) - 1u;
          brows[pt.brow[p] * (NT + 1u) + q] = tps_win_match<K>(win, pm, pt, p);
        }
      }
      {
        const uint32_t before = tps_block_excl_scan(tps_popc32(U), s_wt[0], tid); /* barrier inside */
        UP[q] = make_uint2(U, before);
      }
      /* (2) presence rows R_p of word q, added bit-sliced into the count planes */
      switch (mode) {
        case 1: tps_presence_planes<1>(SP + q, P4, stride, da, dr, nz, Z + q, Zhi + q); break;
        case 2: tps_presence_planes<2>(SP + q, P4, stride, da, dr, nz, Z + q, Zhi + q); break;
        case 3: tps_presence_planes<3>(SP + q, P4, stride, da, dr, nz, Z + q, Zhi + q); break;
        default: tps_presence_planes<0>(SP + q, P4, stride, da, dr, nz, Z + q, Zhi + q); break;
      }
      if (nb) { /* starts of a self-overlapping literal with another start of it less than K ahead */
        uint32_t CF = 0u;
        for (uint32_t bi = 0; bi < nb; ++bi) {
          const uint32_t m0 = brows[bi * (NT + 1u) + q], m1 = brows[bi * (NT + 1u) + q + 1u];
#pragma unroll
          for (int d = 1; d < K; ++d) CF |= m0 & __funnelshift_r(m0, m1, d);
        }
        const uint32_t before = tps_block_excl_scan(tps_popc32(CF), s_wt[1], tid);
        CP[q] = make_uint2(CF, before);
      }
      __syncthreads();
      /* (3) windows, one group of five per thread: sum_i c_(w+i) = sum_i [A(e_i) - A(ls_i)] + 5 P - sum_i present(ls_i)
       * with A = prefix popcount of U.  The five starts ls + i s sit in one 32-bit view of the rows taken at
       * ls (a funnel shift over words q0, q0+1), so A(ls_i) - A(ls_0) is a popcount under a constant mask, the
       * same at the ends, and the five `present` bits of a count plane are one popcount under the mask gm */
      const uint32_t n_groups = (whi - wlo + 4u) / 5u;
      uint16_t *gs = gdst + wlo / 5u;
      for (uint32_t gi = tid; gi < n_groups; gi += NT) {
        const uint32_t ls = 5u * gi * s; /* tile origin = start of window wlo */
        const uint32_t w0 = wlo + 5u * gi;
        uint32_t sum;
        bool fast = grouped && w0 + 5u <= whi;
        if (fast && nb) { /* no two close starts of a self-overlapping literal anywhere in the five windows */
          const uint32_t ee = ls + 4u * s + D;
          const uint2 f0 = CP[ls >> 5], fe = CP[ee >> 5];
          fast = fe.y + tps_popc32(fe.x & ((1u << (ee & 31u)) - 1u)) == f0.y + tps_popc32(f0.x & ((1u << (ls & 31u)) - 1u));
        }
        if (fast) {
          const uint32_t q0 = ls >> 5, b0 = ls & 31u, e = ls + D, qe = e >> 5, be = e & 31u;
          const uint2 u0 = UP[q0], ue = UP[qe];
          const uint32_t ua = __funnelshift_r(u0.x, UP[q0 + 1u].x, b0), va = __funnelshift_r(ue.x, UP[qe + 1u].x, be);
          const uint32_t a0 = u0.y + tps_popc32(u0.x & ((1u << b0) - 1u)), ae = ue.y + tps_popc32(ue.x & ((1u << be) - 1u));
          sum = 5u * (ae - a0 + P) + (tps_popc32(va & lm1) + tps_popc32(va & lm2) + tps_popc32(va & lm3) + tps_popc32(va & lm4)) -
                (tps_popc32(ua & lm1) + tps_popc32(ua & lm2) + tps_popc32(ua & lm3) + tps_popc32(ua & lm4));
          const uint4 z0 = Z[q0], z1 = Z[q0 + 1u];
          uint32_t present = tps_popc32(__funnelshift_r(z0.x, z1.x, b0) & gm) + 2u * tps_popc32(__funnelshift_r(z0.y, z1.y, b0) & gm) +
                             4u * tps_popc32(__funnelshift_r(z0.z, z1.z, b0) & gm) + 8u * tps_popc32(__funnelshift_r(z0.w, z1.w, b0) & gm);
          if (nz > 4u) {
            const uint4 h0 = Zhi[q0], h1 = Zhi[q0 + 1u];
            present += 16u * tps_popc32(__funnelshift_r(h0.x, h1.x, b0) & gm) + 32u * tps_popc32(__funnelshift_r(h0.y, h1.y, b0) & gm) +
                       64u * tps_popc32(__funnelshift_r(h0.z, h1.z, b0) & gm);
          }
          sum -= present;
        } else {
          sum = 0u;
          for (uint32_t j = 0; j < 5u && w0 + j < whi; ++j)
            sum += tps_bp_window<K>(ls + j * s, D, P, nz, nb, pt, UP, Z, Zhi, CP, brows, NT + 1u);
        }
        gs[gi] = (uint16_t)sum;
      }
      tps_cp_async_wait_all();
      __syncthreads();
    }
    bool mine = true; /* this CTA finds the read's change point */
    if (parts > 1u) { /* the CTA that completes the read collects the group sums of all its parts */
      __syncthreads();
      if (tid == 0) {
        __threadfence(); /* cumulative: the barrier ordered the CTA's stores before it */
        s_last = atomicAdd(a.tile_done + it.slot, 1u) == parts - 1u;
      }
      __syncthreads();
      mine = s_last != 0u;
      if (mine) {
        __threadfence();
        for (uint32_t j = tid; j < ((nW + 4u) / 5u + 7u) >> 3; j += NT) tps_cp_async16(gsum + 8u * j, gdst + 8u * j);
        tps_cp_async_wait_all();
        __syncthreads();
      }
    }
    if (mine) {
      if (a.gs_debug) /* test hook */
        for (uint32_t j = tid; j < (nW + 4u) / 5u; j += NT) a.gs_debug[(size_t)it.slot * a.gs_cap + j] = gsum[j];
      /* the read's change point, straight from its group sums in shared memory */
      const int32_t best_b = tps_changepoint_block<5>(TpsCwShared16{gsum}, nW, s_cp, tid);
      if (tid == 0) tps_store_changepoint(a, it.read, best_b);
    }
    if (tid == 0) s_idx[(i + 2u) % 3u] = idx_next2;
    tps_cp_async_wait_all(); /* the record of the next unit, if no tile of this unit waited for it */
    __syncthreads();
  }
}

/* ------------------------------------------------------------------------------------ K5
 * Overview heat map (descriptive_plot.py:233-313 `patterns_vs_match_heatmap`): for every read longer than
 * min_seq_length and every ORIGIN k-mer (no complements), the leftmost non-overlapping matches of
 * `pattern(.{finding})` -- the k-mer followed by `finding` more characters of any kind, match length
 * match_len = k + finding = len(telopattern) -- in `seq[skip:upto]` (strand 0) and in the complement of
 * `reversed(seq)[skip:upto]` (strand 1: the reversed slice is matched against the complemented k-mers).
 * One warp per (read, strand); output = one bit per selected match start, sel[read][strand][pattern][word].
 * The host turns the bits into the reference's (Pattern, Match, read id) rows.
 *
 * dynamic shared memory (words): pm[2 * U * K] | pmc[2 * U * K] | per warp: lin[3 * lin_words] | rows[U * words_per_row] */
#define TPS_K5_WARPS 4

struct TpsFollowArgs {
  TpsPacked pk;
  const uint64_t *offsets; /* n_reads + 1 */
  uint32_t n_reads;
  uint32_t min_seq_length, skip, upto, match_len;
  uint32_t lin_words, words_per_row;
  uint32_t *sel; /* [n_reads][2][U][words_per_row] */
};

template <int K>
__global__ void __launch_bounds__(TPS_K5_WARPS * 32)
tps_follow_kernel(const TpsFollowArgs a, const TpsPatTable pt) {
  static_assert(K > 0, "the follower scan needs a common k-mer length");
  extern __shared__ __align__(16) uint32_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t U = pt.n, wpr = a.words_per_row, lw = a.lin_words;
  uint2 *pm = reinterpret_cast<uint2 *>(smem);
  uint2 *pmc = pm + U * K;
  tps_build_pattern_masks(pm, pt, K, threadIdx.x, TPS_K5_WARPS * 32);
  for (uint32_t i = threadIdx.x; i < U * K; i += TPS_K5_WARPS * 32) { /* complement: code ^ 2 = plane 1 inverted */
    const uint32_t p = i / K, j = i - p * K;
    pmc[i] = make_uint2(0u - ((pt.lo[p] >> j) & 1u), ~(0u - ((pt.hi[p] >> j) & 1u)));
  }
  __syncthreads();
  uint32_t *lin = smem + 4u * U * K + warp * (3u * lw + U * wpr);
  uint32_t *rows = lin + 3u * lw;
  const uint32_t idx = blockIdx.x * TPS_K5_WARPS + warp;
  if (idx >= 2u * a.n_reads) return;
  const uint32_t r = idx >> 1, strand = idx & 1u;
  const uint64_t off = a.offsets[r];
  const uint32_t L = (uint32_t)(a.offsets[r + 1] - off);
  uint32_t *out = a.sel + ((size_t)idx * U) * wpr;
  const uint32_t hi = L < a.upto ? L : a.upto;
  const uint32_t n = (L > a.min_seq_length && hi > a.skip) ? hi - a.skip : 0u; /* len(seq[skip:upto]) */
  if (n < a.match_len) { /* filtered read, or a slice too short for a single match */
    for (uint32_t i = lane; i < U * wpr; i += 32u) out[i] = 0u;
    return;
  }
  /* strand 0: read[skip, hi); strand 1: reversed read -> the slice read[L - hi, L - skip) reversed */
  const bool rev = strand != 0u;
  const uint64_t g0 = rev ? off + (L - hi) : off + a.skip;
  const uint32_t phase = (uint32_t)(g0 & 15u);
  tps_stage_linear(a.pk, g0, n, lin, lw, lane, 32u);
  for (uint32_t i = lane; i < U * wpr; i += 32u) rows[i] = 0u;
  __syncwarp();
  const uint32_t nq = (n + 31u) >> 5;
  const uint32_t last = n - a.match_len; /* last admissible match start */
  for (uint32_t qc = 0; qc < nq; qc += 32u) {
    const uint32_t q = qc + lane;
    uint32_t a0, a1, av, b0, b1, bv;
    tps_oriented_word(lin, lw, phase, n, rev, q, a0, a1, av);
    tps_oriented_word(lin, lw, phase, n, rev, q + 1u, b0, b1, bv);
    TpsWin<K> win;
    tps_win_init<K>(win, a0, b0, a1, b1, av, bv);
    if (q < nq && q < wpr) {
      /* starts beyond `last` have no room for the `finding` characters behind the k-mer */
      const uint32_t lim = last >= 32u * q ? last - 32u * q : 0u;
      const uint32_t room = last < 32u * q ? 0u : (lim >= 31u ? TPS_FULL : ((2u << lim) - 1u));
      for (uint32_t p = 0; p < U; ++p) rows[p * wpr + q] = tps_win_match<K>(win, rev ? pmc : pm, pt, p) & room;
    }
  }
  __syncwarp();
  if (lane < U) { /* leftmost non-overlapping selection, in place: a match consumes match_len positions */
    uint32_t *row = rows + lane * wpr;
    uint32_t pos = 0u;
    for (uint32_t wi = 0; wi < nq && wi < wpr; ++wi) {
      uint32_t w = row[wi], keep = 0u;
      const uint32_t base = 32u * wi;
      if (pos > base) w &= pos - base >= 32u ? 0u : (TPS_FULL << (pos - base));
      while (w) {
        const uint32_t b = (uint32_t)__ffs((int)w) - 1u;
        keep |= 1u << b;
        pos = base + b + a.match_len;
        w &= pos - base >= 32u ? 0u : (TPS_FULL << (pos - base));
      }
      row[wi] = keep;
    }
  }
  __syncwarp();
  for (uint32_t i = lane; i < U * wpr; i += 32u) out[i] = rows[i];
}
