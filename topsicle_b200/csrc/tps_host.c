/*
 * tps_host.c -- host-side C helpers (no CUDA): the `synth-v1` workload generator used by
 * bench.py / tests, built into topsicle_b200/libtps_host.so.
 *
 * synth-v1: every read has its own xoshiro256** stream seeded from (seed, global read index),
 * so any shard of a configuration can be generated independently (per GPU rank, per batch)
 * and always yields the same reads.
 *   - background bases iid uniform over ACGT;
 *   - a fraction f_telo of reads is telomeric: half "forward" (read begins with the motif
 *     repeated, random phase), half "reverse" (read ends with the reverse-complement motif
 *     repeated); telomere length ~ U[telo_min, telo_max] capped at L - 1000;
 *   - 1 % of reads are near-threshold: U[0.3,0.9] * (1000/len(motif)) motif copies scattered
 *     over the first (or last) 1000 bases;
 *   - sequencing errors on every template base: substitution / insertion / deletion rates;
 *   - 0.05 % of bases become 'N'; 1 % of reads carry a 500-base lower-case stretch.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/topsicle_host.h"

typedef struct { uint64_t s[4]; } rng_t;

static inline uint64_t splitmix64(uint64_t *x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t *r) {
  uint64_t *s = r->s;
  const uint64_t result = rotl(s[1] * 5, 7) * 9;
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl(s[3], 45);
  return result;
}
static inline void rng_seed(rng_t *r, uint64_t seed, uint64_t stream, uint64_t purpose) {
  uint64_t x = seed ^ (stream * 0xD1B54A32D192ED03ull) ^ (purpose * 0x8CB92BA72F3D8DD7ull);
  for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&x);
}
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

static uint32_t synth_length(const tps_synth_cfg *c, uint64_t read) {
  rng_t r;
  rng_seed(&r, c->seed, read, 1);
  double L;
  if (c->len_kind == 0) {
    L = c->len_a;
  } else if (c->len_kind == 1) {
    double u1 = rng_unit(&r), u2 = rng_unit(&r);
    if (u1 < 1e-300) u1 = 1e-300;
    double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    L = exp(c->len_a + c->len_b * z);
  } else {
    double u = rng_unit(&r);
    if (u < 1e-300) u = 1e-300;
    L = c->len_a - c->len_b * log(u);
  }
  if (L < (double)c->len_min) L = (double)c->len_min;
  if (L > (double)c->len_max) L = (double)c->len_max;
  return (uint32_t)L;
}

/* offsets_out[0] = 0, offsets_out[i+1] = offsets_out[i] + L(first_read + i) */
int tps_synth_lengths(const tps_synth_cfg *c, uint64_t first_read, uint32_t n_reads, uint64_t *offsets_out) {
  if (!c || !offsets_out) return -1;
  offsets_out[0] = 0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n_reads; ++i) offsets_out[i + 1] = synth_length(c, first_read + (uint64_t)i);
  for (uint32_t i = 0; i < n_reads; ++i) offsets_out[i + 1] += offsets_out[i];
  return 0;
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};

static inline char comp(char b) {
  switch (b) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return b;
  }
}

/* kind: 0 background, 1 forward telomeric, 2 reverse telomeric, 3 near-threshold fwd, 4 near-threshold rev */
static void synth_read(const tps_synth_cfg *c, uint64_t read, uint8_t *out, uint32_t L, uint8_t *kind_out) {
  rng_t r;
  rng_seed(&r, c->seed, read, 2);
  const uint32_t m = c->motif_len;
  char rcm[32];
  for (uint32_t i = 0; i < m; ++i) rcm[i] = comp(c->motif[m - 1 - i]);
  double u = rng_unit(&r);
  int kind = 0;
  if (u < c->f_telo) kind = (rng_next(&r) & 1) ? 2 : 1;
  else if (u < c->f_telo + c->near_frac) kind = (rng_next(&r) & 1) ? 4 : 3;
  uint32_t T = 0, phase = (uint32_t)(rng_next(&r) % (m ? m : 1));
  if (kind == 1 || kind == 2) {
    T = c->telo_min + (uint32_t)(rng_unit(&r) * (double)(c->telo_max - c->telo_min + 1));
    uint32_t cap = L > 1000 ? L - 1000 : 0;
    if (T > cap) T = cap;
  }
  const uint32_t t_sub = (uint32_t)(c->sub_rate * 4294967296.0);
  const uint32_t t_ins = t_sub + (uint32_t)(c->ins_rate * 4294967296.0);
  const uint32_t t_del = t_ins + (uint32_t)(c->del_rate * 4294967296.0);
  const uint32_t t_n = (uint32_t)(c->n_rate * 4294967296.0);
  uint32_t ti = 0, o = 0;
  while (o < L) {
    uint64_t x = rng_next(&r);
    uint32_t ev = (uint32_t)x;
    uint32_t rb = (uint32_t)(x >> 32) & 3u;
    if (ev >= t_ins && ev < t_del) { /* deletion: skip a template base */
      ++ti;
      continue;
    }
    char b;
    if (ev >= t_sub && ev < t_ins) { /* insertion: random base, template not advanced */
      b = BASES[rb];
    } else {
      if (kind == 1 && ti < T) b = c->motif[(phase + ti) % m];
      else if (kind == 2 && ti + T >= L) b = rcm[(phase + ti) % m];
      else b = BASES[(uint32_t)(x >> 40) & 3u];
      if (ev < t_sub) { /* substitution: a different base */
        char nb = BASES[rb];
        if (nb == b) nb = BASES[(rb + 1u) & 3u];
        b = nb;
      }
      ++ti;
    }
    if (t_n && (uint32_t)((x * 0x9E3779B97F4A7C15ull) >> 32) < t_n) b = 'N';
    out[o++] = (uint8_t)b;
  }
  if (kind >= 3 && L >= m) { /* scatter motif copies into the terminal 1000 bases */
    uint32_t span = L < 1000 ? L : 1000;
    double frac = 0.3 + 0.6 * rng_unit(&r);
    uint32_t copies = (uint32_t)(frac * (1000.0 / (double)m));
    for (uint32_t k = 0; k < copies && span >= m; ++k) {
      uint32_t at = (uint32_t)(rng_next(&r) % (span - m + 1));
      for (uint32_t j = 0; j < m; ++j) {
        if (kind == 3) out[at + j] = (uint8_t)c->motif[j];
        else out[L - span + at + j] = (uint8_t)rcm[j];
      }
    }
  }
  if (rng_unit(&r) < c->lower_frac && L > 500) {
    uint32_t at = (uint32_t)(rng_next(&r) % (L - 500));
    for (uint32_t j = 0; j < 500; ++j) out[at + j] |= 0x20;
  }
  if (kind_out) *kind_out = (uint8_t)kind;
}

/* Fill bases_out[offsets[i] .. offsets[i+1]) for reads first_read .. first_read+n_reads-1.
 * kinds_out (optional, n_reads bytes) receives the read class. */
int tps_synth_fill(const tps_synth_cfg *c, uint64_t first_read, uint32_t n_reads, const uint64_t *offsets,
                   uint8_t *bases_out, uint8_t *kinds_out, int n_threads) {
  if (!c || !offsets || !bases_out || c->motif_len < 1 || c->motif_len > 32) return -1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < (int64_t)n_reads; ++i) {
    uint32_t L = (uint32_t)(offsets[i + 1] - offsets[i]);
    synth_read(c, first_read + (uint64_t)i, bases_out + offsets[i], L, kinds_out ? kinds_out + i : NULL);
  }
  return 0;
}

int tps_host_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
