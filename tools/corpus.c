/*
 * corpus.c -- seeded synthetic corpus for tests and bench.
 * Bench/test tooling only: not part of the product library, not part of the
 * oracle.  Byte-reproducible from (seed, unit index, unit length) so every
 * rank / the CPU arm / the GPU arm see identical inputs.
 *
 * Classes (per unit, chosen by splitmix64(seed*0x100000001B3 + index) % 100
 * when klass < 0):  <50 text, <75 records, <90 random, else runs.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static inline uint64_t sm64_next(uint64_t *s)
{
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t sm64_hash(uint64_t x) { return sm64_next(&x); }

enum { VOCAB = 2048, MAXW = 10 };
static uint8_t vocab[VOCAB][MAXW];
static uint8_t vocab_len[VOCAB];
__attribute__((constructor)) static void vocab_init(void)
{
  uint64_t s = 0xC0FFEE;
  for (int w = 0; w < VOCAB; w++) {
    int len = 2 + (int)(sm64_next(&s) % 9);
    vocab_len[w] = (uint8_t)len;
    for (int k = 0; k < len; k++) vocab[w][k] = (uint8_t)('a' + sm64_next(&s) % 26);
  }
}

static void gen_text(uint8_t *dst, size_t n, uint64_t *s)
{
  size_t o = 0;
  unsigned word = 0;
  while (o < n) {
    uint64_t x = sm64_next(s) & 0xFFFFF;
    unsigned idx = (unsigned)((x * x * x) >> 49); /* floor(2048 * (x/2^20)^3) */
    const uint8_t *w = vocab[idx];
    int len = vocab_len[idx];
    for (int k = 0; k < len && o < n; k++) dst[o++] = w[k];
    word++;
    if (o < n) dst[o++] = (word % 12 == 0) ? '\n' : ' ';
  }
}

static void gen_records(uint8_t *dst, size_t n, uint64_t *s)
{
  uint32_t base = (uint32_t)sm64_next(s);
  uint32_t counter = (uint32_t)(sm64_next(s) & 0xFFFF);
  size_t o = 0;
  uint32_t i = 0;
  while (o < n) {
    uint64_t u = sm64_next(s);
    uint8_t rec[16];
    uint32_t a = counter + i;
    uint32_t b = base + i * (uint32_t)(u % 4);
    uint8_t c = (uint8_t)((u >> 8) % 4);
    uint32_t d = (uint32_t)((u >> 16) % 1000);
    memcpy(rec, &a, 4);
    memcpy(rec + 4, &b, 4);
    rec[8] = rec[9] = rec[10] = rec[11] = c;
    memcpy(rec + 12, &d, 4);
    size_t k = n - o < 16 ? n - o : 16;
    memcpy(dst + o, rec, k);
    o += k;
    i++;
  }
}

static void gen_random(uint8_t *dst, size_t n, uint64_t *s)
{
  size_t o = 0;
  while (o < n) {
    uint64_t u = sm64_next(s);
    size_t k = n - o < 8 ? n - o : 8;
    memcpy(dst + o, &u, k);
    o += k;
  }
}

static void gen_runs(uint8_t *dst, size_t n, uint64_t *s)
{
  size_t o = 0;
  while (o < n) {
    uint64_t u = sm64_next(s);
    uint8_t v = (uint8_t)(u % 256);
    size_t run = 1 + (size_t)((u >> 8) % 4096);
    if (run > n - o) run = n - o;
    memset(dst + o, v, run);
    o += run;
  }
}

/* period-7 pattern (config 5b variant) */
static void gen_period7(uint8_t *dst, size_t n, uint64_t *s)
{
  uint8_t pat[7];
  uint64_t u = sm64_next(s);
  for (int k = 0; k < 7; k++) pat[k] = (uint8_t)(u >> (8 * k));
  for (size_t o = 0; o < n; o++) dst[o] = pat[o % 7];
}

/* single repeated byte (config 5b) */
static void gen_const(uint8_t *dst, size_t n, uint64_t *s) { memset(dst, (int)(sm64_next(s) & 0xff), n); }

int fb_corpus_class(uint64_t seed, uint64_t index)
{
  unsigned r = (unsigned)(sm64_hash(seed * 0x100000001B3ull + index) % 100);
  return r < 50 ? 0 : r < 75 ? 1 : r < 90 ? 2 : 3;
}

/* klass: -1 mixed, 0 text, 1 records, 2 random, 3 runs, 4 const byte, 5 period-7 */
void fb_corpus_unit(uint8_t *dst, size_t n, uint64_t seed, uint64_t index, int klass)
{
  uint64_t s = sm64_hash(seed ^ (index * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull));
  if (klass < 0) klass = fb_corpus_class(seed, index);
  switch (klass) {
  case 0: gen_text(dst, n, &s); break;
  case 1: gen_records(dst, n, &s); break;
  case 2: gen_random(dst, n, &s); break;
  case 3: gen_runs(dst, n, &s); break;
  case 4: gen_const(dst, n, &s); break;
  default: gen_period7(dst, n, &s); break;
  }
}

/* nunit fixed-size units [first, first+nunit) laid out back to back */
void fb_corpus_fill(uint8_t *dst, uint64_t first, uint64_t nunit, uint32_t unit_size, uint64_t seed, int klass)
{
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < (int64_t)nunit; i++)
    fb_corpus_unit(dst + (size_t)i * unit_size, unit_size, seed, first + (uint64_t)i, klass);
}

/* config 3: unit i has length 1024 + splitmix64(seed', first+i) % (15*1024+1) */
uint32_t fb_corpus_var_len(uint64_t seed, uint64_t index)
{
  return 1024u + (uint32_t)(sm64_hash(seed * 0x9E3779B97F4A7C15ull + index * 2 + 1) % (15u * 1024u + 1u));
}
/* off has nunit+1 entries (filled here); returns total bytes; dst may be NULL to size only */
uint64_t fb_corpus_fill_var(uint8_t *dst, uint64_t *off, uint64_t first, uint64_t nunit, uint64_t seed, int klass)
{
  uint64_t o = 0;
  for (uint64_t i = 0; i < nunit; i++) {
    off[i] = o;
    o += fb_corpus_var_len(seed, first + i);
  }
  off[nunit] = o;
  if (dst) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)nunit; i++)
      fb_corpus_unit(dst + off[i], (size_t)(off[i + 1] - off[i]), seed, first + (uint64_t)i, klass);
  }
  return o;
}
