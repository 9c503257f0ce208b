/*
 * tps_pgz.c -- parallel inflate of plain gzip input (what the reference opens with gzip.open,
 * allsteps.py:142-146: every README example and the demo input are `.fastq.gz`).
 *
 * A gzip member is ONE deflate stream: block k can only be decoded after block k-1 (its Huffman tables sit in
 * its own header, but its back-references reach 32 KiB into the text before it), which is why zlib inflates a
 * `.fastq.gz` on one core at ~0.2 GB/s of text.  The way around it (the idea of pugz, Kerbiriou & Chikhi 2019),
 * written from scratch here:
 *
 *   1. cut the next stretch of the compressed file into one piece per thread;
 *   2. every thread but the first finds the first deflate block that starts in its piece: it tries each bit
 *      offset, keeps an offset whose dynamic-Huffman header is a complete prefix code (zlib's own validity
 *      rules), whose block decodes to text bytes only, and which is followed by another valid block header;
 *   3. every thread decodes from its block start to the next thread's block start.  It does not know the 32 KiB
 *      of text before its start, so it decodes into 16-bit SYMBOLS: 0..255 = a literal byte, 256 + i = "byte i of
 *      the unknown window"; copies of symbols are symbols.  Thread k must land EXACTLY on the block start thread
 *      k+1 found (else everything after k is thrown away and redone from where k stopped).  Once the last
 *      32 KiB a piece has produced hold no symbol of the unknown window any more (in FASTQ text: a few hundred
 *      kilobytes into a piece of eight megabytes) the piece goes over to plain BYTE output at the next block;
 *      the first piece of a stretch knows its window and decodes to bytes from the start;
 *   4. the last 32 KiB of every piece are resolved in order (a cheap sequential chain), then the symbolic fronts
 *      of all pieces are translated to bytes and their byte parts copied, in parallel, straight into the caller's
 *      buffer, each thread taking the CRC-32 of its bytes on the way (folded with PCLMULQDQ); the CRCs are
 *      combined and compared with the gzip trailer at the end of the member.
 *
 * The block decoders (tps_pgz_decode.inc, compiled for symbols and for bytes) look up to three literals, or a
 * length code with its base and extra-bit count, in one 12-bit table and distances in a direct 10-bit table.
 *
 * Anything unusual (a stream that is not text, stored / fixed blocks, several members, a wrong guess) costs
 * speed, never correctness: the first piece of every stretch starts at a known position with a known window, and
 * the chain check plus the member's CRC-32 / ISIZE are the same guarantees zlib gives.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h> /* crc32, crc32_combine */
#ifdef _OPENMP
#include <omp.h>
#endif

#include "tps_pgz.h"

#include <stdio.h>
#include <sys/mman.h>
#include <time.h>
static double pgz_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- CRC-32 (gzip polynomial) by carry-less multiplication: folds 64 bytes per step (Gopal et al., "Fast CRC
 * computation for generic polynomials using PCLMULQDQ", Intel 2009) instead of zlib's table walk, which at about
 * 2 GB/s per thread was a third of the time of the pass that turns symbols into bytes.  Checked against zlib on
 * first use; zlib's crc32 is used when the CPU lacks the instruction or the check fails. */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define PGZ_HAVE_CLMUL 1
__attribute__((target("pclmul,sse4.1"))) static uint32_t crc32_clmul_raw(const uint8_t *buf, size_t len, uint32_t crc) {
  /* len >= 64 and a multiple of 16; crc = the register (bit-inverted CRC) on both sides */
  static const uint64_t __attribute__((aligned(16))) k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
  static const uint64_t __attribute__((aligned(16))) k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
  static const uint64_t __attribute__((aligned(16))) k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
  static const uint64_t __attribute__((aligned(16))) poly[2] = {0x01db710641ull, 0x01f7011641ull};
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = _mm_load_si128((const __m128i *)k1k2);
  buf += 64;
  len -= 64;
  while (len >= 64) { /* four independent 128-bit lanes, each folded over 512 bits */
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
    x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
    x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
    y6 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
    y8 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
    x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
    x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64;
    len -= 64;
  }
  x0 = _mm_load_si128((const __m128i *)k3k4); /* the four lanes into one */
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) {
    x2 = _mm_loadu_si128((const __m128i *)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16;
    len -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10); /* 128 -> 64 bits */
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_loadl_epi64((const __m128i *)k5k0);
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_load_si128((const __m128i *)poly); /* Barrett reduction to 32 bits */
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

static int pgz_clmul_ok = -1; /* -1 not probed yet */

static uint32_t pgz_crc32_zlib(uint32_t c, const uint8_t *p, uint64_t n) {
  while (n) {
    const uint64_t step = n > (1u << 30) ? (1u << 30) : n;
    c = (uint32_t)crc32(c, p, (uInt)step);
    p += step;
    n -= step;
  }
  return c;
}

/* crc32(c, p, n) of zlib, faster */
static uint32_t pgz_crc32(uint32_t c, const uint8_t *p, uint64_t n) {
#ifdef PGZ_HAVE_CLMUL
  if (pgz_clmul_ok < 0) { /* probe: CPU support and agreement with zlib on lengths around the folding steps */
    const char *env = getenv("TPS_PGZ_CLMUL"); /* 0 = zlib's crc32 (A/B, tests) */
    int ok = !(env && atoi(env) == 0) && __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (ok) {
      uint8_t t[64 * 5 + 48];
      uint32_t x = 12345u;
      for (size_t i = 0; i < sizeof(t); ++i) {
        x = x * 1664525u + 1013904223u;
        t[i] = (uint8_t)(x >> 24);
      }
      for (size_t len = 64; len <= sizeof(t) && ok; len += 16)
        ok = ~crc32_clmul_raw(t, len, ~0x89abcdefu) == (uint32_t)crc32(0x89abcdefu, t, (uInt)len);
    }
    pgz_clmul_ok = ok;
  }
  if (pgz_clmul_ok && n >= 64) {
    const uint64_t body = n & ~(uint64_t)15;
    c = ~crc32_clmul_raw(p, body, ~c);
    p += body;
    n -= body;
  }
#endif
  return pgz_crc32_zlib(c, p, n);
}

#define PGZ_WSIZE 32768u
#define LIT_TB 11 /* primary table bits, literal / length code */
#define DST_TB 8  /* primary table bits, distance code */
#define MAX_SUB 5120 /* > 286 long codes x 16-entry subtables */

typedef struct huff {
  uint32_t tab[(1u << LIT_TB) + MAX_SUB]; /* entry: sym << 16 | kind << 8 | len; kind 1 = subtable (sym = offset, len = its bits) */
  uint32_t tb;
} huff;

typedef struct bitrd {
  const uint8_t *base, *p, *end;
  uint64_t buf;
  uint32_t cnt;
} bitrd;

static inline void br_init(bitrd *b, const uint8_t *base, uint64_t len, uint64_t bitpos) {
  b->base = base;
  b->end = base + len;
  b->p = base + (bitpos >> 3);
  b->buf = 0;
  b->cnt = 0;
  if (b->p < b->end) {
    b->buf = (uint64_t)*b->p++ >> (bitpos & 7);
    b->cnt = 8 - (uint32_t)(bitpos & 7);
  }
}
static inline void br_refill(bitrd *b) {
  if (b->p + 8 <= b->end) { /* libdeflate-style branch-light refill */
    uint64_t w;
    memcpy(&w, b->p, 8);
    b->buf |= w << b->cnt;
    b->p += (63 - b->cnt) >> 3;
    b->cnt |= 56;
  } else {
    while (b->cnt <= 56 && b->p < b->end) {
      b->buf |= (uint64_t)*b->p++ << b->cnt;
      b->cnt += 8;
    }
  }
}
static inline uint64_t br_pos(const bitrd *b) { return (uint64_t)(b->p - b->base) * 8 - b->cnt; }
static inline uint32_t br_bits(bitrd *b, uint32_t n) { /* n <= 16, after a refill */
  const uint32_t v = (uint32_t)(b->buf & ((1u << n) - 1u));
  b->buf >>= n;
  b->cnt -= n;
  return v;
}

static inline uint32_t rev_bits(uint32_t c, uint32_t n) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < n; ++i) r |= ((c >> i) & 1u) << (n - 1 - i);
  return r;
}

/* Canonical Huffman decoding table from code lengths (RFC 1951 3.2.2) under zlib's validity rules
 * (inftrees.c): an over-subscribed set is invalid; an incomplete set is invalid unless it is a single code of
 * length 1 and `allow_single`.  Returns 0 if valid. */
static int huff_build(huff *h, const uint8_t *lens, uint32_t n, uint32_t tb, int allow_single) {
  uint32_t count[16] = {0}, next[16], maxlen = 0, ncodes = 0;
  for (uint32_t i = 0; i < n; ++i) {
    count[lens[i]]++;
    if (lens[i]) {
      ++ncodes;
      if (lens[i] > maxlen) maxlen = lens[i];
    }
  }
  h->tb = tb;
  const uint32_t psize = 1u << tb;
  if (ncodes == 0) { /* no codes at all: legal for the distance code of a block without matches */
    if (!allow_single) return -1;
    for (uint32_t i = 0; i < psize; ++i) h->tab[i] = 0; /* len 0 = invalid code when used */
    return 0;
  }
  int32_t left = 1;
  for (uint32_t l = 1; l <= 15; ++l) {
    left <<= 1;
    left -= (int32_t)count[l];
    if (left < 0) return -1;
  }
  if (left > 0 && !(allow_single && maxlen == 1)) return -1;
  uint32_t code = 0;
  count[0] = 0;
  for (uint32_t l = 1; l <= 15; ++l) {
    code = (code + count[l - 1]) << 1;
    next[l] = code;
  }
  for (uint32_t i = 0; i < psize; ++i) h->tab[i] = 0;
  /* subtable sizes: longest code behind every primary index */
  uint8_t sub[1u << LIT_TB];
  int have_long = 0;
  if (maxlen > tb) {
    memset(sub, 0, psize);
    uint32_t nx[16];
    memcpy(nx, next, sizeof(nx));
    for (uint32_t s = 0; s < n; ++s) {
      const uint32_t l = lens[s];
      if (!l) continue;
      const uint32_t r = rev_bits(nx[l]++, l);
      if (l > tb) {
        have_long = 1;
        const uint32_t pi = r & (psize - 1);
        if (l - tb > sub[pi]) sub[pi] = (uint8_t)(l - tb);
      }
    }
  }
  uint32_t suboff = psize;
  if (have_long) {
    for (uint32_t pi = 0; pi < psize; ++pi) {
      if (!sub[pi]) continue;
      if (suboff + (1u << sub[pi]) > psize + MAX_SUB) return -1;
      h->tab[pi] = (suboff << 16) | (1u << 8) | sub[pi];
      for (uint32_t k = 0; k < (1u << sub[pi]); ++k) h->tab[suboff + k] = 0;
      suboff += 1u << sub[pi];
    }
  }
  for (uint32_t s = 0; s < n; ++s) {
    const uint32_t l = lens[s];
    if (!l) continue;
    const uint32_t r = rev_bits(next[l]++, l);
    if (l <= tb) {
      for (uint32_t k = r; k < psize; k += 1u << l) h->tab[k] = (s << 16) | l;
    } else {
      const uint32_t pe = h->tab[r & (psize - 1)];
      const uint32_t off = pe >> 16, sb = pe & 255u;
      for (uint32_t k = r >> tb; k < (1u << sb); k += 1u << (l - tb)) h->tab[off + k] = (s << 16) | (l - tb);
    }
  }
  return 0;
}

/* one symbol; returns -1 on an invalid code.  The reader was refilled (>= 32 bits unless at the end).
 * tb is a compile-time constant at every call site (the function is inlined). */
static inline __attribute__((always_inline)) int32_t huff_sym_tb(const huff *h, bitrd *b, const uint32_t tb) {
  uint32_t e = h->tab[b->buf & ((1u << tb) - 1u)];
  if (__builtin_expect((e >> 8) & 1u, 0)) {
    b->buf >>= tb;
    b->cnt -= tb;
    e = h->tab[(e >> 16) + (uint32_t)(b->buf & ((1u << (e & 255u)) - 1u))];
  }
  const uint32_t l = e & 255u;
  if (__builtin_expect(l == 0 || l > b->cnt, 0)) return -1;
  b->buf >>= l;
  b->cnt -= l;
  return (int32_t)(e >> 16);
}
static inline int32_t huff_sym(const huff *h, bitrd *b) { return huff_sym_tb(h, b, h->tb); }

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

#define FAST_TB 12  /* index bits of the literal-run table */
#define DFAST_TB 10 /* index bits of the direct distance table */
#define PGZ_F_LEN 0x40u
#define PGZ_F_EOB 0x80u

typedef struct blockhdr {
  int final, type;
  huff lit, dst;
  uint32_t stored_len;
  /* Literal runs (build_fast): the next FAST_TB bits -> up to three literals whose codes fit them whole.
   * entry: bits 0-3 code bits consumed | bits 4-5 literals (0 = not a literal run: take the general path) |
   * bits 8-15, 16-23, 24-31 the literals.  FASTQ text is mostly literals -- bases at 2-3 bits, quality characters at
   * 5-6 -- and a one-symbol-per-lookup decoder is bound by the lookup's latency, so two or three per lookup nearly
   * double or triple its rate.
   * A length code that fits is there as well: bit 6 set | bits 0-3 code bits | bits 8-10 extra bits | bits 16-24
   * base length; the end of block: bit 7 set | bits 0-3 code bits. */
  uint32_t fast[1u << FAST_TB];
  /* distance codes, direct: bits 0-3 code bits (0 = code longer than DFAST_TB, unused or invalid: general path) |
   * bits 8-11 extra bits | bits 16-30 base distance */
  uint32_t dfast[1u << DFAST_TB];
} blockhdr;

/* The literal-run table of a block with Huffman codes; needs h->lit. */
static void build_fast(blockhdr *h) {
  const uint32_t *tab = h->lit.tab;
  const uint32_t pmask = (1u << LIT_TB) - 1u;
  for (uint32_t i = 0; i < (1u << FAST_TB); ++i) {
    uint32_t bits = 0, n = 0, pay = 0, j = i, room = FAST_TB;
    while (n < 3 && room) {
      const uint32_t e = tab[j & pmask];
      const uint32_t l = e & 255u;
      /* the code must be fully inside the bits that are known; long codes, invalid codes, lengths and the end of
       * block go the general way */
      if (((e >> 8) & 1u) || l == 0 || l > room || (e >> 16) >= 256u) break;
      pay |= (e >> 16) << (8u * n);
      ++n;
      bits += l;
      room -= l;
      j >>= l;
    }
    if (n) {
      h->fast[i] = bits | (n << 4) | (pay << 8);
      continue;
    }
    const uint32_t e = tab[i & pmask], l = e & 255u, sym = e >> 16;
    uint32_t f = 0u;
    if (!((e >> 8) & 1u) && l != 0 && l <= FAST_TB && sym >= 256u) {
      if (sym == 256u) f = PGZ_F_EOB | l;
      else if (sym - 257u < 29u) f = PGZ_F_LEN | l | ((uint32_t)LEN_EXTRA[sym - 257u] << 8) | ((uint32_t)LEN_BASE[sym - 257u] << 16);
    }
    h->fast[i] = f;
  }
  const uint32_t *dt = h->dst.tab;
  for (uint32_t i = 0; i < (1u << DFAST_TB); ++i) {
    const uint32_t e = dt[i & ((1u << DST_TB) - 1u)];
    uint32_t l = e & 255u, sym = e >> 16, ok = 1;
    if ((e >> 8) & 1u) { /* subtable: the code is DST_TB + its own bits long */
      const uint32_t sb = l;
      if (DST_TB + sb <= DFAST_TB) {
        const uint32_t e2 = dt[sym + ((i >> DST_TB) & ((1u << sb) - 1u))];
        l = (e2 & 255u) ? DST_TB + (e2 & 255u) : 0u;
        sym = e2 >> 16;
      } else {
        /* the subtable index is only partly inside the DFAST_TB bits: the entry is right if the code it selects
         * is short enough to lie inside them */
        const uint32_t known = DFAST_TB - DST_TB;
        const uint32_t e2 = dt[sym + ((i >> DST_TB) & ((1u << known) - 1u))];
        if ((e2 & 255u) && (e2 & 255u) <= known) {
          l = DST_TB + (e2 & 255u);
          sym = e2 >> 16;
        } else {
          ok = 0;
        }
      }
    }
    h->dfast[i] = (ok && l != 0 && l <= DFAST_TB && sym < 30u)
                      ? (l | ((uint32_t)DST_EXTRA[sym] << 8) | ((uint32_t)DST_BASE[sym] << 16)) : 0u;
  }
}

/* Parse one block header at the reader's position.  0 ok, -1 invalid / out of data. */
static int read_block_header(bitrd *b, blockhdr *h) {
  br_refill(b);
  if (b->cnt < 3) return -1;
  h->final = (int)br_bits(b, 1);
  h->type = (int)br_bits(b, 2);
  if (h->type == 3) return -1;
  if (h->type == 0) {
    const uint32_t drop = b->cnt & 7u; /* to the byte boundary */
    b->buf >>= drop;
    b->cnt -= drop;
    br_refill(b);
    if (b->cnt < 32) return -1;
    const uint32_t len = br_bits(b, 16), nlen = br_bits(b, 16);
    if ((len ^ nlen) != 0xFFFFu) return -1;
    h->stored_len = len;
    return 0;
  }
  uint8_t lens[320];
  if (h->type == 1) {
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    if (huff_build(&h->lit, lens, 288, LIT_TB, 0)) return -1;
    for (int i = 0; i < 32; ++i) lens[i] = 5; /* 30 and 31 complete the code; using them is an error (RFC 1951 3.2.6) */
    return huff_build(&h->dst, lens, 32, DST_TB, 1);
  }
  if (b->cnt < 14) return -1;
  const uint32_t hlit = br_bits(b, 5) + 257, hdist = br_bits(b, 5) + 1, hclen = br_bits(b, 4) + 4;
  if (hlit > 286 || hdist > 30) return -1;
  uint8_t cl[19] = {0};
  br_refill(b);
  for (uint32_t i = 0; i < hclen; ++i) {
    if (b->cnt < 3) {
      br_refill(b);
      if (b->cnt < 3) return -1;
    }
    cl[CL_ORDER[i]] = (uint8_t)br_bits(b, 3);
  }
  huff clh;
  if (huff_build(&clh, cl, 19, 7, 0)) return -1;
  uint32_t i = 0;
  while (i < hlit + hdist) {
    br_refill(b);
    const int32_t s = huff_sym(&clh, b);
    if (s < 0) return -1;
    if (s < 16) {
      lens[i++] = (uint8_t)s;
    } else {
      uint32_t rep, val = 0;
      if (b->cnt < 7) return -1;
      if (s == 16) {
        if (i == 0) return -1;
        val = lens[i - 1];
        rep = 3 + br_bits(b, 2);
      } else if (s == 17) {
        rep = 3 + br_bits(b, 3);
      } else {
        rep = 11 + br_bits(b, 7);
      }
      if (i + rep > hlit + hdist) return -1;
      while (rep--) lens[i++] = (uint8_t)val;
    }
  }
  if (lens[256] == 0) return -1; /* no end-of-block code */
  if (huff_build(&h->lit, lens, hlit, LIT_TB, 0)) return -1;
  return huff_build(&h->dst, lens + hlit, hdist, DST_TB, 1);
}

typedef struct obuf16 { /* symbols */
  uint16_t *out;
  uint64_t n, cap;
  int fixed; /* the buffer is not ours to grow */
} obuf16;
typedef struct obuf8 { /* bytes */
  uint8_t *out;
  uint64_t n, cap;
  int fixed;
} obuf8;

/* One piece of a stretch.  Its text is sym.n symbols followed by byt.n bytes: a piece that starts with an unknown
 * window decodes to symbols until its last 32 KiB hold none that refers to the unknown window (in FASTQ text that
 * takes a few hundred kilobytes of a piece's eight megabytes), and to bytes from the next block on -- half the
 * store traffic, and nothing left to translate afterwards; the first piece of a stretch knows its window and
 * decodes to bytes from the start. */
typedef struct seg {
  uint64_t start_bit; /* a block starts here */
  uint64_t stop_bit;  /* decode whole blocks until the position is >= this */
  uint64_t end_bit;   /* where the decode stopped (a block boundary) */
  obuf16 sym;
  obuf8 byt;
  uint8_t *bwin;      /* the 32 KiB before the byte phase (owned unless it is the stretch's known window) */
  int bwin_owned;
  uint64_t scanned;   /* symbols already searched for window references */
  uint64_t unres_end; /* 1 + index of the last symbol that refers to the unknown window */
  int status;    /* 0 ok, -1 corrupt, -2 out of memory */
  int saw_final; /* stopped behind the member's last block */
  int symbolic;  /* started with an unknown window */
  uint32_t crc;
} seg;

/* Output buffers are megabytes per thread and written once per stretch: 2 MiB-aligned and advised as huge pages, so
 * that first touch costs one fault per 2 MiB instead of 512 (with eight threads faulting at once the kernel's
 * address-space lock made the first stretches several times slower than the rest). */
static void *big_alloc(uint64_t nbytes) {
  const uint64_t bytes = (nbytes + (2u << 20) - 1) & ~(uint64_t)((2u << 20) - 1);
  void *p = NULL;
  if (posix_memalign(&p, 2u << 20, bytes)) return NULL;
#ifdef MADV_HUGEPAGE
  madvise(p, bytes, MADV_HUGEPAGE);
#endif
  return p;
}
static uint16_t *sym_alloc(uint64_t n) { return (uint16_t *)big_alloc(n * sizeof(uint16_t)); }

static int grow_sym(obuf16 *s, uint64_t need) {
  if (s->n + need <= s->cap) return 0;
  if (s->fixed) return -1;
  uint64_t nc = s->cap ? s->cap * 2 : (1u << 22);
  while (nc < s->n + need) nc *= 2;
  uint16_t *nv = sym_alloc(nc);
  if (!nv) return -1;
  if (s->n) memcpy(nv, s->out, s->n * sizeof(uint16_t));
  free(s->out);
  s->out = nv;
  s->cap = nc;
  return 0;
}
static int grow_byte(obuf8 *s, uint64_t need) {
  if (s->n + need <= s->cap) return 0;
  if (s->fixed) return -1;
  uint64_t nc = s->cap ? s->cap * 2 : (1u << 22);
  while (nc < s->n + need) nc *= 2;
  uint8_t *nv = (uint8_t *)big_alloc(nc);
  if (!nv) return -1;
  if (s->n) memcpy(nv, s->out, s->n);
  free(s->out);
  s->out = nv;
  s->cap = nc;
  return 0;
}

static inline int is_text(uint32_t c) { return c == 10 || c == 13 || c == 9 || (c >= 32 && c < 127); }

#define PGZ_T uint16_t
#define PGZ_OBUF obuf16
#define PGZ_NAME(x) x##_sym
#define PGZ_BYTES 0
#define PGZ_CHUNK 8u
#include "tps_pgz_decode.inc"
#undef PGZ_T
#undef PGZ_OBUF
#undef PGZ_NAME
#undef PGZ_BYTES
#undef PGZ_CHUNK
#define PGZ_T uint8_t
#define PGZ_OBUF obuf8
#define PGZ_NAME(x) x##_byte
#define PGZ_BYTES 1
#define PGZ_CHUNK 16u
#include "tps_pgz_decode.inc"
#undef PGZ_T
#undef PGZ_OBUF
#undef PGZ_NAME
#undef PGZ_BYTES
#undef PGZ_CHUNK

static int pgz_fast_on = 1; /* TPS_PGZ_FAST=0 (read by tps_pgz_open): one symbol per lookup everywhere (A/B, tests) */

static int pgz_bytes_on = 1; /* TPS_PGZ_BYTES=0 (read by tps_pgz_open): pieces stay symbolic to their end (A/B, tests) */

static int decode_block_text(bitrd *b, const blockhdr *h, obuf16 *s, const uint16_t *win, uint64_t max_out) {
  return decode_block_impl_sym(b, h, s, win, 1, max_out);
}

/* Decode whole blocks from s->start_bit until the position reaches s->stop_bit or the member ends.
 * win = the 32768 symbols before a symbolic piece; a piece with a known window comes with s->bwin set. */
static void decode_segment(const uint8_t *z, uint64_t zlen, seg *s, const uint16_t *win) {
  bitrd b;
  br_init(&b, z, zlen, s->start_bit);
  blockhdr *h = (blockhdr *)malloc(sizeof(blockhdr));
  if (!h) {
    s->status = -2;
    return;
  }
  s->status = 0;
  int bytes = s->bwin != NULL;
  for (;;) {
    if (read_block_header(&b, h)) {
      s->status = -1;
      break;
    }
    int rc;
    const int fast = h->type != 0 && pgz_fast_on;
    if (fast) build_fast(h);
    if (bytes) rc = fast ? decode_block_fast_byte(&b, h, &s->byt, s->bwin) : decode_block_impl_byte(&b, h, &s->byt, s->bwin, 0, 0);
    else rc = fast ? decode_block_fast_sym(&b, h, &s->sym, win) : decode_block_impl_sym(&b, h, &s->sym, win, 0, 0);
    if (rc) {
      s->status = rc;
      break;
    }
    s->end_bit = br_pos(&b);
    if (h->final) {
      s->saw_final = 1;
      break;
    }
    if (s->end_bit >= s->stop_bit) break;
    if (!bytes && pgz_bytes_on) { /* do the last 32 KiB still refer to the unknown window? */
      const uint16_t *o = s->sym.out;
      uint64_t k = s->scanned, last = s->unres_end;
      for (; k < s->sym.n; ++k)
        if (o[k] >= 256u) last = k + 1;
      s->scanned = k;
      s->unres_end = last;
      if (s->sym.n >= PGZ_WSIZE && s->sym.n - last >= PGZ_WSIZE) {
        s->bwin = (uint8_t *)malloc(PGZ_WSIZE);
        if (!s->bwin) {
          s->status = -2;
          break;
        }
        s->bwin_owned = 1;
        for (uint32_t j = 0; j < PGZ_WSIZE; ++j) s->bwin[j] = (uint8_t)o[s->sym.n - PGZ_WSIZE + j];
        bytes = 1;
      }
    }
  }
  free(h);
}

/* One complete raw deflate stream that starts with an empty window and inflates to exactly `want` <= 65536 bytes (a
 * BGZF block) -> dst, with the CRC-32 of the bytes.  Uses the byte decoders above; thread-safe (per-thread tables and
 * scratch).  0 ok, -1 corrupt / another length, -3 not applicable (want too large, out of memory). */
int tps_pgz_inflate_block(const uint8_t *z, uint64_t zlen, uint8_t *dst, uint64_t want, uint32_t *crc_out) {
  static __thread blockhdr *h = NULL;
  static __thread uint8_t *scratch = NULL;
  static const uint8_t no_window[PGZ_WSIZE] = {0};
  if (!z || !dst || !crc_out || want > 65536u) return -3;
  if (!h) h = (blockhdr *)malloc(sizeof(blockhdr));
  if (!scratch) scratch = (uint8_t *)malloc(65536u + 2048u);
  if (!h || !scratch) return -3;
  obuf8 o;
  o.out = scratch;
  o.n = 0;
  o.cap = 65536u + 2048u;
  o.fixed = 1;
  bitrd b;
  br_init(&b, z, zlen, 0);
  for (;;) {
    if (read_block_header(&b, h)) return -1;
    int rc;
    if (h->type != 0 && pgz_fast_on) {
      build_fast(h);
      rc = decode_block_fast_byte(&b, h, &o, no_window);
    } else {
      rc = decode_block_impl_byte(&b, h, &o, no_window, 0, 0);
    }
    if (rc || o.n > want) return -1;
    if (h->final) break;
  }
  if (o.n != want) return -1;
  memcpy(dst, scratch, want);
  *crc_out = pgz_crc32((uint32_t)crc32(0L, Z_NULL, 0), dst, want);
  return 0;
}

/* First bit position >= from (and < limit) where a dynamic-Huffman block starts: valid header, the block decodes
 * to text, and another valid block header follows.  Returns ~0 if none. */
static uint64_t find_block_start(const uint8_t *z, uint64_t zlen, uint64_t from, uint64_t limit, const uint16_t *symwin) {
  blockhdr *h = (blockhdr *)malloc(sizeof(blockhdr)), *h2 = (blockhdr *)malloc(sizeof(blockhdr));
  obuf16 t;
  memset(&t, 0, sizeof(t));
  uint64_t found = ~0ull;
  if (h && h2) {
    for (uint64_t pos = from; pos < limit; ++pos) {
      /* cheap pre-filter on the first 17 header bits: BFINAL 0, BTYPE 2, HLIT <= 29, HDIST <= 29 */
      const uint64_t byte = pos >> 3;
      if (byte + 4 > zlen) break;
      uint32_t w = (uint32_t)z[byte] | ((uint32_t)z[byte + 1] << 8) | ((uint32_t)z[byte + 2] << 16) | ((uint32_t)z[byte + 3] << 24);
      w >>= pos & 7;
      if ((w & 7u) != 4u) continue;          /* BFINAL = 0, BTYPE = 10b */
      if (((w >> 3) & 31u) > 29u) continue;  /* HLIT */
      if (((w >> 8) & 31u) > 29u) continue;  /* HDIST */
      /* the HCLEN + 4 three-bit lengths of the code-length code must form a complete prefix code (zlib accepts
       * nothing else): a Kraft sum over 57 bits read in one go turns away nearly every position that got here */
      {
        const uint64_t at = pos + 17;
        if ((at >> 3) + 8 <= zlen) {
          uint64_t v;
          memcpy(&v, z + (at >> 3), 8);
          v >>= at & 7;
          const uint32_t hclen = ((w >> 13) & 15u) + 4u;
          uint32_t kraft = 0;
          for (uint32_t i = 0; i < hclen; ++i) {
            const uint32_t l = (uint32_t)(v >> (3u * i)) & 7u;
            kraft += l ? 128u >> l : 0u;
          }
          if (kraft != 128u) continue;
        }
      }
      bitrd b;
      br_init(&b, z, zlen, pos);
      if (read_block_header(&b, h)) continue;
      t.n = 0;
      if (decode_block_text(&b, h, &t, symwin, 1u << 22)) continue;
      if (t.n < 1024) continue; /* a real block of a FASTQ stream holds tens of kilobytes */
      if (read_block_header(&b, h2)) continue;
      found = pos;
      break;
    }
  }
  free(t.out);
  free(h);
  free(h2);
  return found;
}

struct tps_pgz {
  const uint8_t *z;
  uint64_t zlen;
  int threads;
  uint64_t pos_bit;          /* next block of the current member */
  int in_member;
  int eof;
  uint8_t window[PGZ_WSIZE]; /* last text of the current member (shorter at its start: wlen) */
  uint32_t wlen;
  uint32_t crc;
  uint64_t isize;
  uint16_t *bufs[256];       /* symbol buffers of the pieces, kept from stretch to stretch */
  uint64_t bufcap[256];
  uint8_t *bbufs[256];       /* byte buffers of the pieces, likewise */
  uint64_t bbufcap[256];
  uint8_t *q;                /* text produced but not yet handed out */
  uint64_t q_len, q_off;
  double ratio;              /* text bytes per compressed byte so far */
  uint64_t piece;            /* compressed bytes per thread and stretch */
  tps_pgz_stats st;
  char err[160];
};

static int pgz_fail(tps_pgz *g, const char *msg) {
  strncpy(g->err, msg, sizeof(g->err) - 1);
  return -1;
}

/* gzip member header at byte offset `at` (RFC 1952); returns the offset of the deflate data or 0. */
static uint64_t gzip_header(const uint8_t *z, uint64_t zlen, uint64_t at) {
  if (at + 18 > zlen || z[at] != 0x1f || z[at + 1] != 0x8b || z[at + 2] != 8) return 0;
  const uint8_t flg = z[at + 3];
  if (flg & 0xE0) return 0;
  uint64_t p = at + 10;
  if (flg & 4) {
    if (p + 2 > zlen) return 0;
    p += 2 + ((uint64_t)z[p] | ((uint64_t)z[p + 1] << 8));
  }
  if (flg & 8) {
    while (p < zlen && z[p]) ++p;
    ++p;
  }
  if (flg & 16) {
    while (p < zlen && z[p]) ++p;
    ++p;
  }
  if (flg & 2) p += 2;
  return p < zlen ? p : 0;
}

tps_pgz *tps_pgz_open(const uint8_t *zmap, uint64_t zlen, int threads) {
  const uint64_t d = gzip_header(zmap, zlen, 0);
  if (!d) return NULL;
  tps_pgz *g = (tps_pgz *)calloc(1, sizeof(*g));
  if (!g) return NULL;
  g->z = zmap;
  g->zlen = zlen;
  g->threads = threads > 0 ? (threads > 256 ? 256 : threads) : 1;
  g->pos_bit = d * 8;
  g->in_member = 1;
  g->crc = (uint32_t)crc32(0L, Z_NULL, 0);
  pgz_clmul_ok = -1; /* probed here, on one thread, under this open's environment */
  (void)pgz_crc32(0u, zmap, 0);
  {
    const char *f = getenv("TPS_PGZ_FAST");
    pgz_fast_on = !(f && atoi(f) == 0);
    f = getenv("TPS_PGZ_BYTES");
    pgz_bytes_on = !(f && atoi(f) == 0);
  }
  if (getenv("TPS_PGZ_DEBUG"))
    fprintf(stderr, "[pgz] open: %d threads, literal-run tables %s, crc32 by %s\n", g->threads, pgz_fast_on ? "on" : "off",
            pgz_clmul_ok ? "pclmulqdq" : "zlib");
  g->ratio = 4.5;
  g->piece = 4u << 20;
  const char *e = getenv("TPS_PGZ_PIECE"); /* compressed bytes per thread and stretch (tests, tuning) */
  if (e && atoll(e) >= (1 << 16)) g->piece = (uint64_t)atoll(e);
  return g;
}

void tps_pgz_close(tps_pgz *g) {
  if (!g) return;
  for (int i = 0; i < 256; ++i) free(g->bufs[i]);
  for (int i = 0; i < 256; ++i) free(g->bbufs[i]);
  free(g->q);
  free(g);
}

const char *tps_pgz_error(const tps_pgz *g) { return g ? g->err : "out of memory"; }
void tps_pgz_get_stats(const tps_pgz *g, tps_pgz_stats *out) {
  if (g && out) *out = g->st;
}
void tps_pgz_set_piece(tps_pgz *g, uint64_t bytes) {
  if (g && bytes >= (1u << 16)) g->piece = bytes;
}

/* The member ended at pos_bit: check its trailer, move to the next member or to the end of the file. */
static int end_member(tps_pgz *g) {
  uint64_t p = (g->pos_bit + 7) >> 3;
  if (p + 8 > g->zlen) return pgz_fail(g, "gzip trailer missing (truncated file)");
  const uint32_t crc = (uint32_t)g->z[p] | ((uint32_t)g->z[p + 1] << 8) | ((uint32_t)g->z[p + 2] << 16) | ((uint32_t)g->z[p + 3] << 24);
  const uint32_t isz = (uint32_t)g->z[p + 4] | ((uint32_t)g->z[p + 5] << 8) | ((uint32_t)g->z[p + 6] << 16) | ((uint32_t)g->z[p + 7] << 24);
  if (crc != g->crc || isz != (uint32_t)g->isize) return pgz_fail(g, "gzip CRC-32 / length check failed (corrupt file)");
  p += 8;
  g->st.members++;
  while (p < g->zlen && g->z[p] == 0) ++p; /* zero padding between / behind members */
  if (p >= g->zlen) {
    g->eof = 1;
    g->in_member = 0;
    return 0;
  }
  const uint64_t d = gzip_header(g->z, g->zlen, p);
  if (!d) return pgz_fail(g, "data behind the gzip member is not another member");
  g->pos_bit = d * 8;
  g->wlen = 0;
  g->crc = (uint32_t)crc32(0L, Z_NULL, 0);
  g->isize = 0;
  return 0;
}

/* Inflate the next stretch of the member; the text goes to dst (at most cap bytes) or, if the stretch turns out
 * larger, to the internal queue.  Returns the bytes written to dst, -1 on error. */
static int64_t next_stretch(tps_pgz *g, uint8_t *dst, uint64_t cap) {
  const double t_in = pgz_now();
  int T = g->threads;
  /* size the stretch so that its text most likely fits the caller's buffer */
  uint64_t piece = g->piece;
  const uint64_t want_comp = (uint64_t)((double)cap / g->ratio * 0.85);
  if ((uint64_t)T * piece > want_comp) {
    piece = want_comp / (uint64_t)T;
    if (piece < (1u << 18)) {
      piece = 1u << 18;
      T = (int)(want_comp / piece);
      if (T < 1) T = 1;
    }
  }
  const uint64_t start_byte = g->pos_bit >> 3;
  if (start_byte + (uint64_t)T * piece > g->zlen) {
    const uint64_t left = g->zlen - start_byte;
    if (left < (uint64_t)T * (1u << 18)) T = (int)(left >> 18) > 0 ? (int)(left >> 18) : 1;
    piece = left / (uint64_t)T + 1;
  }
  seg *sg = (seg *)calloc((size_t)T, sizeof(seg));
  uint16_t *symwin = (uint16_t *)malloc(PGZ_WSIZE * sizeof(uint16_t));
  if (!sg || !symwin) {
    free(sg); free(symwin);
    return pgz_fail(g, "out of memory");
  }
  for (uint32_t i = 0; i < PGZ_WSIZE; ++i) symwin[i] = (uint16_t)(256u + i);
  const int dbg = getenv("TPS_PGZ_DEBUG") != NULL;
  const double t_a = pgz_now();
  /* 2. block starts */
  uint64_t *starts = (uint64_t *)malloc(((size_t)T + 1) * sizeof(uint64_t));
  starts[0] = g->pos_bit;
#pragma omp parallel for num_threads(T) schedule(static, 1)
  for (int i = 1; i < T; ++i) {
    const uint64_t from = (start_byte + (uint64_t)i * piece) * 8, lim = (start_byte + (uint64_t)(i + 1) * piece) * 8;
    starts[i] = find_block_start(g->z, g->zlen, from, lim < g->zlen * 8 ? lim : g->zlen * 8, symwin);
  }
  int ns = 1; /* pieces that found a start, in order */
  for (int i = 1; i < T; ++i)
    if (starts[i] != ~0ull && starts[i] > starts[ns - 1]) starts[ns++] = starts[i];
  const uint64_t stretch_end = (start_byte + (uint64_t)T * piece) * 8;
  uint8_t *bwin0 = (uint8_t *)calloc(1, PGZ_WSIZE); /* the known window of the first piece, zero-padded in front */
  if (!bwin0) {
    free(sg); free(symwin); free(starts);
    return pgz_fail(g, "out of memory");
  }
  memcpy(bwin0 + (PGZ_WSIZE - g->wlen), g->window, g->wlen);
  for (int i = 0; i < ns; ++i) {
    sg[i].sym.out = g->bufs[i];
    sg[i].sym.cap = g->bufcap[i];
    sg[i].byt.out = g->bbufs[i];
    sg[i].byt.cap = g->bbufcap[i];
    g->bufs[i] = NULL;
    g->bbufs[i] = NULL;
    const uint64_t expect = (uint64_t)((double)piece * g->ratio * 1.5) + (1u << 20); /* no doubling on the way */
    const uint64_t want = (expect + expect / 2 + (1u << 23) - 1) & ~(uint64_t)((1u << 23) - 1);
    /* with headroom: the running ratio moves a little from stretch to stretch.  A piece that may go over to bytes
     * touches only the front of its symbol buffer (untouched pages cost nothing) */
    if (i > 0 && sg[i].sym.cap < expect) {
      free(sg[i].sym.out);
      sg[i].sym.out = sym_alloc(want);
      sg[i].sym.cap = sg[i].sym.out ? want : 0;
    }
    if ((i == 0 || pgz_bytes_on) && sg[i].byt.cap < expect) {
      free(sg[i].byt.out);
      sg[i].byt.out = (uint8_t *)big_alloc(want);
      sg[i].byt.cap = sg[i].byt.out ? want : 0;
    }
    sg[i].start_bit = starts[i];
    sg[i].stop_bit = i + 1 < ns ? starts[i + 1] : stretch_end;
    sg[i].symbolic = i > 0;
    if (i == 0) sg[i].bwin = bwin0; /* bytes from the start */
  }
  const double t_b = pgz_now();
  /* 3. decode */
#pragma omp parallel for num_threads(ns) schedule(static, 1)
  for (int i = 0; i < ns; ++i) decode_segment(g->z, g->zlen, &sg[i], symwin);
  const double t_c = pgz_now();
  /* the chain: segment i must stop exactly where segment i+1 started */
  int good = 0;
  for (int i = 0; i < ns; ++i) {
    if (sg[i].status) break;
    good = i + 1;
    if (sg[i].saw_final) break;
    if (i + 1 < ns && sg[i].end_bit != sg[i + 1].start_bit) {
      g->st.chain_breaks++;
      break;
    }
  }
  int64_t ret = -1;
  if (good == 0) {
    pgz_fail(g, sg[0].status == -2 ? "out of memory" : "corrupt deflate stream");
  } else {
    /* 4. windows in order, then everything to bytes in parallel */
    uint8_t *wins = (uint8_t *)malloc((size_t)good * PGZ_WSIZE);
    uint32_t *wlens = (uint32_t *)malloc((size_t)good * sizeof(uint32_t));
    uint64_t total = 0;
    for (int i = 0; i < good; ++i) total += sg[i].sym.n + sg[i].byt.n;
    uint8_t *target = dst;
    int to_queue = 0;
    if (total > cap) { /* the estimate was too small: keep the text, hand it out piecewise */
      free(g->q);
      g->q = (uint8_t *)malloc(total ? total : 1);
      g->q_len = total;
      g->q_off = 0;
      target = g->q;
      to_queue = 1;
    }
    if (!wins || !wlens || !target) {
      pgz_fail(g, "out of memory");
    } else {
      /* window before segment i = last 32 KiB of the text up to its start */
      memcpy(wins, g->window, g->wlen);
      wlens[0] = g->wlen;
      for (int i = 1; i < good; ++i) {
        const seg *p = &sg[i - 1];
        const uint8_t *pw = wins + (size_t)(i - 1) * PGZ_WSIZE;
        uint8_t *w = wins + (size_t)i * PGZ_WSIZE;
        const uint32_t pwl = wlens[i - 1];
        const uint64_t plen = p->sym.n + p->byt.n;
        const uint64_t take = plen < PGZ_WSIZE ? plen : PGZ_WSIZE;
        const uint32_t keep = take < PGZ_WSIZE ? (uint32_t)((PGZ_WSIZE - take) < pwl ? (PGZ_WSIZE - take) : pwl) : 0;
        memcpy(w, pw + (pwl - keep), keep);
        /* the last `take` elements of the piece: the end of its symbols, then its bytes */
        const uint64_t from_bytes = p->byt.n < take ? p->byt.n : take, from_syms = take - from_bytes;
        for (uint64_t k = 0; k < from_syms; ++k) {
          const uint16_t v = p->sym.out[p->sym.n - from_syms + k];
          /* symbol 256 + j = byte j of the 32 KiB window, whose last pwl bytes are known (a valid stream never
           * reaches further back than the member's start) */
          w[keep + k] = v < 256 ? (uint8_t)v : (v - 256u >= PGZ_WSIZE - pwl ? pw[v - 256u - (PGZ_WSIZE - pwl)] : 0);
        }
        memcpy(w + keep + from_syms, p->byt.out + (p->byt.n - from_bytes), from_bytes);
        wlens[i] = keep + (uint32_t)take;
      }
      uint64_t *offs = (uint64_t *)malloc(((size_t)good + 1) * sizeof(uint64_t));
      offs[0] = 0;
      for (int i = 0; i < good; ++i) offs[i + 1] = offs[i] + sg[i].sym.n + sg[i].byt.n;
      int bad_ref = 0;
#pragma omp parallel for num_threads(good) schedule(static, 1)
      for (int i = 0; i < good; ++i) {
        const seg *s = &sg[i];
        uint8_t *o = target + offs[i];
        const uint8_t *w = wins + (size_t)i * PGZ_WSIZE;
        const uint32_t wl = wlens[i], miss = PGZ_WSIZE - wl;
        const uint16_t *so = s->sym.out;
        const uint64_t sn = s->sym.n;
        uint32_t c = (uint32_t)crc32(0L, Z_NULL, 0);
        /* a chunk of symbols to bytes, then its CRC while the bytes are still in the cache */
        for (uint64_t k0 = 0; k0 < sn; k0 += 1u << 15) {
          const uint64_t k1 = k0 + (1u << 15) < sn ? k0 + (1u << 15) : sn;
          uint64_t k = k0;
          while (k < k1) {
            /* a little into a piece nearly everything is resolved: whole runs of 64 narrow at once */
            if (k + 64 <= k1) {
              uint16_t any = 0;
              for (int j = 0; j < 64; ++j) any |= so[k + j];
              if (any < 256) {
                for (int j = 0; j < 64; ++j) o[k + j] = (uint8_t)so[k + j];
                k += 64;
                continue;
              }
            }
            const uint64_t ke = k + 64 <= k1 ? k + 64 : k1;
            for (; k < ke; ++k) {
              const uint16_t v = so[k];
              if (v < 256) o[k] = (uint8_t)v;
              else if (s->symbolic && v - 256u >= miss) o[k] = w[v - 256u - miss];
              else {
                o[k] = 0;
                bad_ref = 1; /* a reference before the start of the member: corrupt */
              }
            }
          }
          c = pgz_crc32(c, o + k0, k1 - k0);
        }
        /* the piece's bytes: copied and summed chunk by chunk */
        for (uint64_t k0 = 0; k0 < s->byt.n; k0 += 1u << 16) {
          const uint64_t len = s->byt.n - k0 < (1u << 16) ? s->byt.n - k0 : (1u << 16);
          memcpy(o + sn + k0, s->byt.out + k0, len);
          c = pgz_crc32(c, o + sn + k0, len);
        }
        sg[i].crc = c;
      }
      if (dbg)
        fprintf(stderr, "[pgz] stretch: %d pieces of %.2f MB, %d starts, %d good; setup %.4f s, search %.4f s, decode %.4f s, resolve+crc %.4f s, %.1f MB text\n",
                T, piece / 1e6, ns, good, t_a - t_in, t_b - t_a, t_c - t_b, pgz_now() - t_c, total / 1e6);
      if (bad_ref) {
        pgz_fail(g, "corrupt deflate stream (reference before the start of the member)");
      } else {
        for (int i = 0; i < good; ++i)
          g->crc = (uint32_t)crc32_combine(g->crc, sg[i].crc, (z_off_t)(sg[i].sym.n + sg[i].byt.n));
        g->isize += total;
        /* new window */
        const seg *l = &sg[good - 1];
        if (total >= PGZ_WSIZE) {
          memcpy(g->window, target + total - PGZ_WSIZE, PGZ_WSIZE);
          g->wlen = PGZ_WSIZE;
        } else {
          const uint32_t keep = (uint32_t)((PGZ_WSIZE - total) < g->wlen ? (PGZ_WSIZE - total) : g->wlen);
          memmove(g->window, g->window + (g->wlen - keep), keep);
          memcpy(g->window + keep, target, total);
          g->wlen = keep + (uint32_t)total;
        }
        g->pos_bit = l->end_bit;
        g->st.stretches++;
        g->st.segments += (uint64_t)good;
        g->st.text_bytes += total;
        if (good > 1) g->st.parallel_text_bytes += total - sg[0].sym.n - sg[0].byt.n;
        const uint64_t comp = (l->end_bit >> 3) - start_byte;
        if (comp > (1u << 16)) g->ratio = 0.5 * g->ratio + 0.5 * ((double)total / (double)comp);
        if (g->ratio < 1.0) g->ratio = 1.0;
        ret = 0;
        if (l->saw_final && end_member(g)) ret = -1;
        if (ret == 0) ret = to_queue ? 0 : (int64_t)total;
      }
      free(offs);
    }
    free(wins);
    free(wlens);
  }
  for (int i = 0; i < ns; ++i) { /* keep the buffers: fresh ones cost their page faults again */
    g->bufs[i] = sg[i].sym.out;
    g->bufcap[i] = sg[i].sym.cap;
    g->bbufs[i] = sg[i].byt.out;
    g->bbufcap[i] = sg[i].byt.cap;
    if (sg[i].bwin_owned) free(sg[i].bwin);
  }
  free(sg);
  free(symwin);
  free(bwin0);
  free(starts);
  return ret;
}

int64_t tps_pgz_read(tps_pgz *g, uint8_t *dst, uint64_t cap) {
  if (!g || !dst) return -1;
  uint64_t got = 0;
  while (got < cap) {
    if (g->q_off < g->q_len) {
      uint64_t n = g->q_len - g->q_off;
      if (n > cap - got) n = cap - got;
      memcpy(dst + got, g->q + g->q_off, n);
      g->q_off += n;
      got += n;
      if (g->q_off == g->q_len) {
        free(g->q);
        g->q = NULL;
        g->q_len = g->q_off = 0;
      }
      continue;
    }
    if (g->eof) break;
    /* a stretch pays when every thread gets a full piece: with less room than that left in the caller's buffer,
     * hand back what there is (the caller comes back with a fresh buffer) */
    if (got && (double)(cap - got) < 0.6 * (double)g->threads * (double)g->piece * g->ratio) break;
    const int64_t n = next_stretch(g, dst + got, cap - got);
    if (n < 0) return -1;
    got += (uint64_t)n;
  }
  return (int64_t)got;
}

int tps_pgz_eof(const tps_pgz *g) { return g && g->eof && g->q_off >= g->q_len; }
