// Host build of topsicle_b200/csrc/tps_bitops.h for CPU unit tests (tests/test_bitops_host.py).
// Test infrastructure: compiled with g++ into tests/csrc/_build/libbitops_host.so.
#include "../../topsicle_b200/csrc/tps_bitops.h"

extern "C" {
uint32_t t_pack16(const uint8_t *b, uint32_t *bad) {
  uint32_t w[4];
  for (int i = 0; i < 4; ++i)
    w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  return tps_pack16(w[0], w[1], w[2], w[3], bad);
}
uint32_t t_exact_mask16_simd(const uint8_t *b) {
  uint32_t w[4];
  for (int i = 0; i < 4; ++i)
    w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  return tps_exact_mask16_simd(w[0], w[1], w[2], w[3]);
}
uint32_t t_exact_mask16(const uint8_t *b) {
  uint32_t w[4];
  for (int i = 0; i < 4; ++i)
    w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  return tps_exact_mask16(w[0], w[1], w[2], w[3]);
}
uint32_t t_code_at(uint32_t u, uint32_t g) { return tps_code_at(u, g); }
uint32_t t_linear_planes(uint32_t u) { return tps_linear_planes(u); }
uint32_t t_ascii_code(uint32_t c) { return tps_ascii_code(c); }
uint32_t t_greedy_count(const uint32_t *m, int32_t from, int32_t to, uint32_t k) { return tps_greedy_count(m, from, to, k); }
uint32_t t_range_popcount(const uint32_t *m, int32_t from, int32_t to) { return tps_range_popcount(m, from, to); }
// returns argmax b over candidates (same loop as K4, sequential), -1 if none
int32_t t_change_point(const uint32_t *cw, uint32_t n) {
  if (n < 7) return -1;
  uint64_t T = 0, S = 0;
  for (uint32_t i = 0; i < n; ++i) T += cw[i];
  tps_cand best; best.b = -1; best.d = 0; best.den = 1; best.num_f = 0.0; best.den_f = 1.0;
  uint32_t pos = 0;
  for (uint32_t b = 0; b < n; b += 5) {
    while (pos < b) S += cw[pos++];
    if (b >= 2 && n - b >= 2) {
      tps_cand c = tps_make_cand(n, S, T, b);
      if (tps_cand_better(&best, &c)) best = c;
    }
  }
  return best.b;
}
}
