/*
 * tps_bitops.h -- bit-level building blocks shared by the CUDA kernels (device) and
 * the host unit tests (tests/csrc).  Pure integer code, no CUDA intrinsics.
 *
 * Packed read format produced by K1 (tps_pack_kernel), per group of 16 consecutive
 * bases (= 4 little-endian 32-bit words w0..w3 of ASCII):
 *
 *   code word (uint32): base g = 4*i + m (word i, byte m) has its 2-bit code at bits
 *       8*m + 2*i (+1).  code = (ascii >> 1) & 3  ->  A=0 C=1 T=2 G=3, case-insensitive
 *       (this is what `.upper()` + literal matching needs: allsteps.py:176-177,267-271).
 *       complement(code) = code ^ 2.
 *   flag bit: 1 iff any of the 16 bytes is not in {A,C,G,T,a,c,g,t}; such bytes can never
 *       be part of a match (the reference's regex literals contain only ACGT).
 *   exact mask (uint16, only written for flagged groups): bit g = 1 iff base g is valid.
 */
#ifndef TPS_BITOPS_H
#define TPS_BITOPS_H

#include <stdint.h>

#if defined(__CUDACC__)
#define TPS_HD __host__ __device__ __forceinline__
#else
#define TPS_HD static inline
#endif

/* ---- K1: 16 ASCII bytes -> code word, and "group has an invalid byte" flag ---------- */
/* Field layout: base g = 4*i + m (word i, byte m) owns bits 8m+2i, 8m+2i+1. */
TPS_HD uint32_t tps_pack16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t *bad_out) {
  /* bits 1,2 of every byte -> 2-bit fields at 8m+2i */
  uint32_t u = ((w0 >> 1) & 0x03030303u) | ((w1 << 1) & 0x0C0C0C0Cu) | ((w2 << 3) & 0x30303030u) |
               ((w3 << 5) & 0xC0C0C0C0u);
  /* A byte is one of ACGTacgt <=> bit7=0, bit6=1, bit3=0, bit4 == z, bit0 == !z with
   * z = bit2 & ~bit1 (1 only for T/t); bit5 is the case bit and is ignored.  The relational
   * part is evaluated per word at bit position 4 using left shifts only (they run on the FMA
   * pipe as IMAD.SHL, keeping the ALU pipe for the LOP3s). */
#define TPS_REL4(w) ((((w) ^ (((w) << 2) & ~((w) << 3))) | ~((((w) << 4)) ^ (((w) << 2) & ~((w) << 3)))))
  uint32_t inv = (TPS_REL4(w0) | TPS_REL4(w1)) | (TPS_REL4(w2) | TPS_REL4(w3));
#undef TPS_REL4
  /* bits 7,6,3 must read 0,1,0 in every byte: OR of (w ^ 0x40) over the four words */
  uint32_t hi = ((w0 ^ 0x40404040u) | (w1 ^ 0x40404040u)) | ((w2 ^ 0x40404040u) | (w3 ^ 0x40404040u));
  *bad_out = (inv & 0x10101010u) | (hi & 0xC8C8C8C8u);
  return u;
}

/* even-bit field word (bit 8m+2i for base 4i+m) -> 16 bits in linear base order */
TPS_HD uint32_t tps_fields_to_linear16(uint32_t x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0F0F0F0Fu;
  x = (x | (x >> 4)) & 0x00FF00FFu;
  x = (x | (x >> 8)) & 0x0000FFFFu;
  uint32_t t = (x ^ (x >> 3)) & 0x0A0Au;
  x ^= t ^ (t << 3);
  t = (x ^ (x >> 6)) & 0x00CCu;
  x ^= t ^ (t << 6);
  return x;
}

/* exact validity bits (bit g = base g is ACGTacgt) of a flagged group, branch-free */
TPS_HD uint32_t tps_exact_mask16_simd(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  /* every tested ASCII bit gathered to the even field positions 8m+2i */
  uint32_t u1 = ((w0 >> 1) & 0x01010101u) | ((w1 << 1) & 0x04040404u) | ((w2 << 3) & 0x10101010u) |
                ((w3 << 5) & 0x40404040u);
  uint32_t u2 = ((w0 >> 2) & 0x01010101u) | (w1 & 0x04040404u) | ((w2 << 2) & 0x10101010u) |
                ((w3 << 4) & 0x40404040u);
  uint32_t u0 = (w0 & 0x01010101u) | ((w1 << 2) & 0x04040404u) | ((w2 << 4) & 0x10101010u) |
                ((w3 << 6) & 0x40404040u);
  uint32_t u4 = ((w0 >> 4) & 0x01010101u) | ((w1 >> 2) & 0x04040404u) | (w2 & 0x10101010u) |
                ((w3 << 2) & 0x40404040u);
  uint32_t u7 = ((w0 >> 7) & 0x01010101u) | ((w1 >> 5) & 0x04040404u) | ((w2 >> 3) & 0x10101010u) |
                ((w3 >> 1) & 0x40404040u);
  uint32_t u6 = ((w0 >> 6) & 0x01010101u) | ((w1 >> 4) & 0x04040404u) | ((w2 >> 2) & 0x10101010u) |
                (w3 & 0x40404040u);
  uint32_t u3 = ((w0 >> 3) & 0x01010101u) | ((w1 >> 1) & 0x04040404u) | ((w2 << 1) & 0x10101010u) |
                ((w3 << 3) & 0x40404040u);
  uint32_t z = u2 & ~u1;
  uint32_t inv = (u4 ^ z) | ~(u0 ^ z) | u7 | u3 | ~u6;
  return (~tps_fields_to_linear16(inv)) & 0xFFFFu;
}

TPS_HD int tps_byte_is_acgt(uint32_t c) {
  c &= 0xDFu;
  return c == 0x41u || c == 0x43u || c == 0x47u || c == 0x54u;
}

/* exact validity bits of a 16-base group (slow path, only for flagged groups) */
TPS_HD uint32_t tps_exact_mask16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  uint32_t w[4] = {w0, w1, w2, w3};
  uint32_t m = 0;
  for (int g = 0; g < 16; ++g) {
    uint32_t c = (w[g >> 2] >> (8 * (g & 3))) & 0xFFu;
    m |= (uint32_t)tps_byte_is_acgt(c) << g;
  }
  return m;
}

/* 2-bit code of base g (0..15) inside a code word */
TPS_HD uint32_t tps_code_at(uint32_t u, uint32_t g) { return (u >> (8 * (g & 3) + 2 * (g >> 2))) & 3u; }

/* code word -> linear bit planes: low 16 bits = plane0 (code bit 0) of bases 0..15 in order,
 * high 16 bits = plane1 (code bit 1).  Bit 8m + 2i + pl of the code word (base 4i + m, plane pl) goes to bit
 * 16 pl + 4i + m: a permutation of the five INDEX bits [m1 m0 i1 i0 pl] -> [pl i1 i0 m1 m0], done as four
 * exchanges of two index bits (delta swaps: positions 4<->0, 3<->2, 2<->1, 1<->0), 16 instructions instead of the
 * 28 of compressing each plane and transposing it. */
TPS_HD uint32_t tps_linear_planes(uint32_t y) {
  uint32_t t = (y ^ (y >> 15)) & 0x0000AAAAu;
  y ^= t ^ (t << 15);
  t = (y ^ (y >> 4)) & 0x00F000F0u;
  y ^= t ^ (t << 4);
  t = (y ^ (y >> 2)) & 0x0C0C0C0Cu;
  y ^= t ^ (t << 2);
  t = (y ^ (y >> 1)) & 0x22222222u;
  y ^= t ^ (t << 1);
  return y;
}

/* ---- pattern properties ------------------------------------------------------------- */
/* ASCII base -> code, or 0xFF if not ACGT (upper or lower case) */
TPS_HD uint32_t tps_ascii_code(uint32_t c) { return tps_byte_is_acgt(c) ? ((c >> 1) & 3u) : 0xFFu; }

/* ---- greedy (leftmost, non-overlapping) count over a match bitmask -------------------
 * mask bit j = literal of length k matches at position j.  Counts matches with start in
 * [from, to] (inclusive), restarting the greedy scan at `from` -- exactly
 * len(list(re.finditer(literal, text))) for text = s[from : to + k]
 * (allsteps.py:182-183, 281, 288). */
TPS_HD uint32_t tps_greedy_count(const uint32_t *mask, int32_t from, int32_t to, uint32_t k) {
  uint32_t cnt = 0;
  int32_t pos = from;
  while (pos <= to) {
    int32_t wi = pos >> 5;
    uint32_t w = mask[wi] & (0xFFFFFFFFu << (pos & 31));
    const int32_t wlast = to >> 5;
    while (w == 0u) {
      if (++wi > wlast) return cnt;
      w = mask[wi];
    }
    /* index of lowest set bit */
    uint32_t low = w & (0u - w);
    int32_t b = 0;
#if defined(__CUDA_ARCH__)
    b = __ffs((int)w) - 1;
    (void)low;
#else
    while (!((low >> b) & 1u)) ++b;
#endif
    int32_t at = (wi << 5) + b;
    if (at > to) return cnt;
    ++cnt;
    pos = at + (int32_t)k;
  }
  return cnt;
}

/* popcount of mask bits with index in [from, to] (inclusive); for border-free literals this
 * equals the greedy count */
TPS_HD uint32_t tps_popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(v);
#else
  return (uint32_t)__builtin_popcount(v);
#endif
}

TPS_HD uint32_t tps_range_popcount(const uint32_t *mask, int32_t from, int32_t to) {
  if (to < from) return 0;
  int32_t w0 = from >> 5, w1 = to >> 5;
  uint32_t first = 0xFFFFFFFFu << (from & 31);
  uint32_t last = 0xFFFFFFFFu >> (31 - (to & 31));
  if (w0 == w1) return tps_popc32(mask[w0] & first & last);
  uint32_t c = tps_popc32(mask[w0] & first) + tps_popc32(mask[w1] & last);
  for (int32_t w = w0 + 1; w < w1; ++w) c += tps_popc32(mask[w]);
  return c;
}

/* ---- change point: exact comparison of two candidate gains --------------------------
 * gain(b) is proportional to  d^2/den  with d = n*S_b - b*T, den = b*(n-b)
 * (equivalent to ruptures' C(0,n) - C(0,b) - C(b,n) with CostL2; the common factor 1/(n*P^2)
 * cancels).  tps_cand_better returns 1 if candidate B is better than A under ruptures'
 * `max((gain, bkp))` rule: larger gain, ties -> larger b.
 * The decision is exact: a float64 cross-multiplication settles every pair whose gains differ by
 * more than 1e-9 relative (the float64 evaluation is good to ~1e-15), anything closer -- ties and
 * near-ties -- is compared as 128-bit integers (d^2 * den fits: checked at tps_create). */
typedef struct tps_cand {
  int64_t d;
  uint64_t den;
  double num_f; /* (double)d * (double)d */
  double den_f;
  int32_t b;
} tps_cand;

TPS_HD int tps_cand_better(const tps_cand *A, const tps_cand *B) {
  if (A->b < 0) return B->b >= 0;
  if (B->b < 0) return 0;
  const double l = B->num_f * A->den_f, r = A->num_f * B->den_f;
  const double diff = l - r, big = l > r ? l : r;
  if (diff > 1e-9 * big) return 1;
  if (-diff > 1e-9 * big) return 0;
  const uint64_t da = (uint64_t)(A->d < 0 ? -A->d : A->d), db = (uint64_t)(B->d < 0 ? -B->d : B->d);
  const unsigned __int128 le = (unsigned __int128)db * db * (unsigned __int128)A->den;
  const unsigned __int128 re = (unsigned __int128)da * da * (unsigned __int128)B->den;
  if (le != re) return le > re;
  return B->b > A->b;
}

TPS_HD tps_cand tps_make_cand(uint32_t n, uint64_t S_b, uint64_t T, uint32_t b) {
  tps_cand c;
  c.d = (int64_t)((uint64_t)n * S_b) - (int64_t)((uint64_t)b * T);
  c.den = (uint64_t)b * (uint64_t)(n - b);
  const double df = (double)c.d;
  c.num_f = df * df;
  c.den_f = (double)c.den;
  c.b = (int32_t)b;
  return c;
}

#endif /* TPS_BITOPS_H */
