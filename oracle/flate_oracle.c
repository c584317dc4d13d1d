/*
 * flate_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of gmlewis/moonbit-flate's deflate-fast encoder and
 * inflate decoder.  Every function names the reference file:line it follows
 * (paths relative to /root/reference).  See flate_oracle.h for who may use it
 * and for how it is pinned ("parity by transcription + reference KATs + zlib").
 *
 * Reference quirks that change bytes/behaviour and are reproduced on purpose:
 *  D1  DeflateFast.prev is never populated (deflate-fast.mbt:114,:138,:157,:349)
 *      -> match_len returns 0 for candidates in the previous block.
 *  D2  "store instead" test is ssize < (size+size)>>4 (huffman-bit-writer.mbt:527,:780).
 *  D3  new_dict compresses the dictionary into the output (deflate.mbt:108-151).
 *  D5  more_bits returns the reader's error unwrapped (inflate.mbt:789-799), so
 *      input exhausted at a block header / extra-bits refill surfaces as plain
 *      ioeof, while huff_sym / data_block wrap it into unexpected-EOF
 *      (inflate.mbt:824,:719,:753).
 */
#include "flate_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ================================================================== */
/* constants (deflate-fast.mbt:12-55,:89-92; huffman-bit-writer.mbt:11-44;
 * inflate.mbt:22-34,:69-78)                                           */
enum {
  TABLE_BITS = 14,
  TABLE_SIZE = 1 << TABLE_BITS,
  TABLE_MASK = TABLE_SIZE - 1,
  TABLE_SHIFT = 32 - TABLE_BITS,
  BASE_MATCH_LENGTH = 3,
  MAX_MATCH_LENGTH = 258,
  BASE_MATCH_OFFSET = 1,
  MAX_MATCH_OFFSET = 1 << 15,
  MAX_STORE_BLOCK_SIZE = 65535,
  INPUT_MARGIN = 16 - 1,
  MIN_NON_LITERAL_BLOCK_SIZE = 1 + 1 + INPUT_MARGIN,
  OFFSET_CODE_COUNT = 30,
  END_BLOCK_MARKER = 256,
  LENGTH_CODES_START = 257,
  CODEGEN_CODE_COUNT = 19,
  BAD_CODE = 255,
  MAX_NUM_LIT = 286,
  MAX_NUM_DIST = 30,
  NUM_CODES = 19,
  MAX_CODE_LEN = 16,
  MAX_BITS_LIMIT = 16,
  HUFFMAN_CHUNK_BITS = 9,
  HUFFMAN_NUM_CHUNKS = 1 << HUFFMAN_CHUNK_BITS,
  HUFFMAN_COUNT_MASK = 15,
  HUFFMAN_VALUE_SHIFT = 4
};
#define INT32_MAXV 2147483647
#define BUFFER_RESET (INT32_MAXV - MAX_STORE_BLOCK_SIZE * 2)
#define MATCH_TYPE (1u << 30)
#define LENGTH_SHIFT 22
#define OFFSET_MASK ((1u << LENGTH_SHIFT) - 1)

/* 32-bit wrapping add, as MoonBit Int arithmetic */
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

/* ================================================================== */
/* bits.mbt:11-21 (table generated instead of listed)                  */
static uint32_t rev8(uint32_t x)
{
  uint32_t r = 0;
  for (int i = 0; i < 8; i++)
    if (x & (1u << i)) r |= 0x80u >> i;
  return r;
}
uint32_t orc_reverse16(uint32_t x) { return rev8((x & 0xff00) >> 8) | (rev8(x & 0xff) << 8); }
/* huffman-code.mbt:283-286 */
uint32_t orc_reverse_bits(uint32_t number, uint32_t bit_length)
{
  return orc_reverse16(number << (16 - bit_length));
}

/* ================================================================== */
/* LUTs.  huffman-bit-writer.mbt:49-85 gives the base/extra tables; the two
 * 256-entry code LUTs of token.mbt:30-61 are "largest code whose base <= x",
 * which is how they are regenerated here (checked in tests against spot
 * values of the reference listing).                                    */
static const int length_extra_bits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                          2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint32_t length_base[29] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  10,
                                         12, 14, 16, 20, 24, 28, 32, 40, 48, 56,
                                         64, 80, 96, 112, 128, 160, 192, 224, 255};
static const int offset_extra_bits[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                          6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint32_t offset_base[30] = {
    0x0000, 0x0001, 0x0002, 0x0003, 0x0004, 0x0006, 0x0008, 0x000c, 0x0010, 0x0018,
    0x0020, 0x0030, 0x0040, 0x0060, 0x0080, 0x00c0, 0x0100, 0x0180, 0x0200, 0x0300,
    0x0400, 0x0600, 0x0800, 0x0c00, 0x1000, 0x1800, 0x2000, 0x3000, 0x4000, 0x6000};
/* huffman-bit-writer.mbt:83-85 == inflate.mbt:424-426 */
static const int codegen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

static uint8_t length_codes[256];
static uint8_t offset_codes[256];

__attribute__((constructor)) static void orc_init_luts(void)
{
  for (int x = 0; x < 256; x++) {
    int c = 0;
    while (c + 1 < 29 && length_base[c + 1] <= (uint32_t)x) c++;
    length_codes[x] = (uint8_t)c;
    c = 0;
    while (c + 1 < 30 && offset_base[c + 1] <= (uint32_t)x) c++;
    offset_codes[x] = (uint8_t)c;
  }
}

/* token.mbt:69-123 */
static inline uint32_t literal_token(uint32_t lit) { return lit; }
static inline uint32_t match_token(uint32_t xlength, uint32_t xoffset)
{
  return MATCH_TYPE + (xlength << LENGTH_SHIFT) + xoffset;
}
static inline uint32_t tok_offset(uint32_t t) { return t & OFFSET_MASK; }
static inline uint32_t tok_length(uint32_t t) { return (t - MATCH_TYPE) >> LENGTH_SHIFT; }
static inline int length_code(uint32_t len) { return length_codes[len]; }
static inline int offset_code(uint32_t off)
{
  if (off < 256) return offset_codes[off];
  if ((off >> 7) < 256) return offset_codes[off >> 7] + 14;
  return offset_codes[off >> 14] + 28;
}
uint32_t orc_token_offset(uint32_t tok) { return tok_offset(tok); }
int orc_length_code(uint32_t xlen) { return length_code(xlen); }
int orc_offset_code(uint32_t xoff) { return offset_code(xoff); }

/* ================================================================== */
/* growable vectors                                                    */
typedef struct {
  uint8_t *p;
  size_t len, cap;
} bytevec;
static void bv_push(bytevec *v, const uint8_t *src, size_t n)
{
  if (v->len + n > v->cap) {
    size_t nc = v->cap ? v->cap * 2 : 4096;
    while (nc < v->len + n) nc *= 2;
    v->p = (uint8_t *)realloc(v->p, nc);
    v->cap = nc;
  }
  memcpy(v->p + v->len, src, n);
  v->len += n;
}
typedef struct {
  uint32_t *p;
  size_t len, cap;
} tokvec;
static inline void tv_push(tokvec *v, uint32_t t)
{
  if (v->len == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 65536 + 16;
    v->p = (uint32_t *)realloc(v->p, v->cap * sizeof(uint32_t));
  }
  v->p[v->len++] = t;
}

/* ================================================================== */
/* deflate-fast.mbt                                                    */
static inline uint32_t load32(const uint8_t *b, int i) /* :58-63 */
{
  return (uint32_t)b[i] | ((uint32_t)b[i + 1] << 8) | ((uint32_t)b[i + 2] << 16) |
         ((uint32_t)b[i + 3] << 24);
}
static inline uint64_t load64(const uint8_t *b, int i) /* :66-75 */
{
  return (uint64_t)load32(b, i) | ((uint64_t)load32(b, i + 4) << 32);
}
static inline int hash4(uint32_t u) { return (int)((u * 0x1e35a7bdu) >> TABLE_SHIFT); } /* :78-81 */

typedef struct { /* :95-98 */
  uint32_t val;
  int32_t offset;
} table_entry;

typedef struct { /* :104-108 */
  table_entry table[TABLE_SIZE];
  const uint8_t *prev; /* previous block; the reference never populates it (D1) */
  int prev_len;
  int32_t cur;
} deflate_fast;

static void df_new(deflate_fast *e) /* :111-117 */
{
  memset(e->table, 0, sizeof e->table);
  e->prev = NULL;
  e->prev_len = 0;
  e->cur = MAX_STORE_BLOCK_SIZE;
}

static void df_shift_offsets(deflate_fast *e) /* :366-389 */
{
  if (e->prev_len == 0) {
    memset(e->table, 0, sizeof e->table);
    e->cur = MAX_MATCH_OFFSET + 1;
    return;
  }
  for (int i = 0; i < TABLE_SIZE; i++) {
    int32_t v = wadd(wadd(e->table[i].offset, -e->cur), MAX_MATCH_OFFSET + 1);
    if (v < 0) v = 0;
    e->table[i].offset = v;
  }
  e->cur = MAX_MATCH_OFFSET + 1;
}

static void df_reset(deflate_fast *e) /* :348-358 */
{
  e->prev_len = 0;
  e->cur = wadd(e->cur, MAX_MATCH_OFFSET);
  if (e->cur >= BUFFER_RESET) df_shift_offsets(e);
}

static void emit_literal(tokvec *dst, const uint8_t *lit, int n) /* :273-279 */
{
  for (int i = 0; i < n; i++) tv_push(dst, literal_token(lit[i]));
}

static int df_match_len(const deflate_fast *e, int s, int t, const uint8_t *src, int n) /* :286-342 */
{
  int s1 = s + MAX_MATCH_LENGTH - 4;
  if (s1 > n) s1 = n;
  if (t >= 0) { /* inside the current block */
    int a_length = s1 - s;
    for (int i = 0; i < a_length; i++)
      if (src[s + i] != src[t + i]) return i;
    return a_length;
  }
  /* match in the previous block: prev is always empty, so tp = 0 + t < 0 (D1) */
  int tp = e->prev_len + t;
  if (tp < 0) return 0;
  int a_length = s1 - s;
  int b_length = e->prev_len - tp;
  if (b_length > a_length) b_length = a_length;
  for (int i = 0; i < b_length; i++)
    if (src[s + i] != e->prev[tp + i]) return i;
  int nn = b_length;
  if (s + nn == s1) return nn;
  a_length = s1 - (s + nn);
  for (int i = 0; i < a_length; i++)
    if (src[s + nn + i] != src[i]) return i + nn;
  return a_length + nn;
}

static void df_encode(deflate_fast *e, tokvec *dst, const uint8_t *src, int n) /* :123-270 */
{
  if (e->cur >= BUFFER_RESET) df_shift_offsets(e);
  if (n < MIN_NON_LITERAL_BLOCK_SIZE) {
    e->cur = wadd(e->cur, MAX_STORE_BLOCK_SIZE);
    e->prev_len = 0;
    emit_literal(dst, src, n);
    return;
  }
  const int s_limit = n - INPUT_MARGIN;
  int next_emit = 0;
  int s = 0;
  uint32_t cv = load32(src, s);
  int next_hash = hash4(cv);

  for (;;) {
    int skip = 32;
    int next_s = s;
    table_entry candidate = {0, 0};
    for (;;) { /* :183-202 */
      s = next_s;
      int bytes_between_hash_lookups = skip >> 5;
      next_s = s + bytes_between_hash_lookups;
      skip += bytes_between_hash_lookups;
      if (next_s > s_limit) goto emit_remainder;
      candidate = e->table[next_hash & TABLE_MASK];
      uint32_t now = load32(src, next_s);
      e->table[next_hash & TABLE_MASK].offset = wadd(s, e->cur);
      e->table[next_hash & TABLE_MASK].val = cv;
      next_hash = hash4(now);
      int32_t offset = wadd(s, -wadd(candidate.offset, -e->cur));
      if (offset > MAX_MATCH_OFFSET || cv != candidate.val) {
        cv = now;
        continue;
      }
      break;
    }
    emit_literal(dst, src + next_emit, s - next_emit); /* :207 */
    for (;;) {                                         /* :217-266 */
      s += 4;
      int t = wadd(wadd(candidate.offset, -e->cur), 4);
      int l = df_match_len(e, s, t, src, n);
      tv_push(dst, match_token((uint32_t)(l + 4 - BASE_MATCH_LENGTH),
                               (uint32_t)(s - t - BASE_MATCH_OFFSET)));
      s += l;
      next_emit = s;
      if (s >= s_limit) goto emit_remainder;
      uint64_t x = load64(src, s - 1);
      int prev_hash = hash4((uint32_t)x);
      e->table[prev_hash & TABLE_MASK].offset = wadd(e->cur, s - 1);
      e->table[prev_hash & TABLE_MASK].val = (uint32_t)x;
      x >>= 8;
      int curr_hash = hash4((uint32_t)x);
      candidate = e->table[curr_hash & TABLE_MASK];
      e->table[curr_hash & TABLE_MASK].offset = wadd(e->cur, s);
      e->table[curr_hash & TABLE_MASK].val = (uint32_t)x;
      int32_t offset = wadd(s, -wadd(candidate.offset, -e->cur));
      if (offset > MAX_MATCH_OFFSET || (uint32_t)x != candidate.val) {
        cv = (uint32_t)(x >> 8);
        next_hash = hash4(cv);
        s += 1;
        break;
      }
    }
  }
emit_remainder: /* :152-159 */
  if (next_emit < n) emit_literal(dst, src + next_emit, n - next_emit);
  e->cur = wadd(e->cur, n);
  /* slice_copy(self.prev, src) copies min(prev.length(), n) = 0 bytes (D1) */
}

/* ================================================================== */
/* huffman-code.mbt                                                    */
typedef struct { /* :37-40 */
  uint32_t code;
  uint32_t len;
} hcode;
typedef struct { /* :28-31 */
  uint32_t literal;
  int32_t freq;
} literal_node;
typedef struct { /* :45-62 */
  int level;
  int32_t last_freq, next_char_freq, next_pair_freq, needed;
} level_info;
typedef struct { /* :9-13 */
  hcode codes[MAX_NUM_LIT];
  literal_node freqcache[MAX_NUM_LIT + 1];
  int32_t bit_count[17];
  int ncodes;
} huffman_encoder;

static void he_new(huffman_encoder *h, int size) /* :16-26 */
{
  memset(h, 0, sizeof *h);
  h->ncodes = size;
}

static int he_bit_length(const huffman_encoder *h, const int32_t *freq, int nfreq) /* :83-91 */
{
  int32_t total = 0;
  for (int i = 0; i < nfreq; i++)
    if (freq[i] != 0) total = wadd(total, (int32_t)((uint32_t)freq[i] * h->codes[i].len));
  return total;
}

/* simple-quicksort.mbt:45 -- both comparators are strict total orders over
 * distinct literals, so any correct sort yields the same permutation. */
static int cmp_by_frequency(const void *pa, const void *pb) /* huffman-code.mbt:346-351 */
{
  const literal_node *a = (const literal_node *)pa, *b = (const literal_node *)pb;
  if (a->freq == b->freq) return (int32_t)a->literal < (int32_t)b->literal ? -1 : 1;
  return a->freq < b->freq ? -1 : 1;
}
static int cmp_by_literal(const void *pa, const void *pb) /* :354-356 */
{
  const literal_node *a = (const literal_node *)pa, *b = (const literal_node *)pb;
  return (int32_t)a->literal < (int32_t)b->literal ? -1 : ((int32_t)a->literal > (int32_t)b->literal);
}

/* :112-244.  list has room for list[n]; returns max_bits actually used and
 * fills h->bit_count[1..max_bits]. */
static int he_bit_counts(huffman_encoder *h, literal_node *list, int n, int max_bits)
{
  if (max_bits >= MAX_BITS_LIMIT) abort();
  list[n].literal = 0xffffffffu; /* max_node(), :77-79 */
  list[n].freq = INT32_MAXV;
  if (max_bits > n - 1) max_bits = n - 1;

  level_info levels[MAX_BITS_LIMIT];
  int32_t leaf_counts[MAX_BITS_LIMIT][MAX_BITS_LIMIT];
  memset(levels, 0, sizeof levels);
  memset(leaf_counts, 0, sizeof leaf_counts);

  for (int level = 1; level <= max_bits; level++) {
    levels[level].level = level;
    levels[level].last_freq = list[1].freq;
    levels[level].next_char_freq = list[2].freq;
    levels[level].next_pair_freq = wadd(list[0].freq, list[1].freq);
    levels[level].needed = 0;
    leaf_counts[level][level] = 2;
    if (level == 1) levels[level].next_pair_freq = INT32_MAXV;
  }
  levels[max_bits].needed = 2 * n - 4;

  int level = max_bits;
  for (;;) {
    level_info *l = &levels[level];
    if (l->next_pair_freq == INT32_MAXV && l->next_char_freq == INT32_MAXV) {
      l->needed = 0;
      levels[level + 1].next_pair_freq = INT32_MAXV;
      level++;
      continue;
    }
    int32_t prev_freq = l->last_freq;
    if (l->next_char_freq < l->next_pair_freq) {
      int nn = leaf_counts[level][level] + 1;
      l->last_freq = l->next_char_freq;
      leaf_counts[level][level] = nn;
      l->next_char_freq = list[nn].freq;
    } else {
      l->last_freq = l->next_pair_freq;
      for (int i = 0; i < level; i++) leaf_counts[level][i] = leaf_counts[level - 1][i];
      levels[l->level - 1].needed = 2;
    }
    l->needed--;
    if (l->needed == 0) {
      if (l->level == max_bits) break;
      levels[l->level + 1].next_pair_freq = wadd(prev_freq, l->last_freq);
      level++;
    } else {
      while (levels[level - 1].needed > 0) level--;
    }
  }
  if (leaf_counts[max_bits][max_bits] != n) abort();
  int bits = 1;
  const int32_t *counts = leaf_counts[max_bits];
  for (int lv = max_bits; lv > 0; lv--) {
    h->bit_count[bits] = counts[lv] - counts[lv - 1];
    bits++;
  }
  return max_bits;
}

/* :250-280 */
static void he_assign_encoding_and_size(huffman_encoder *h, int nbit_count, literal_node *list, int nlist)
{
  uint32_t code = 0;
  for (int n = 0; n < nbit_count; n++) {
    int bits = h->bit_count[n];
    code <<= 1;
    if (n == 0 || bits == 0) continue;
    literal_node *chunk = list + (nlist - bits);
    qsort(chunk, (size_t)bits, sizeof *chunk, cmp_by_literal);
    for (int i = 0; i < bits; i++) {
      uint32_t key = chunk[i].literal;
      h->codes[key].code = orc_reverse_bits(code & 0xffff, (uint32_t)n);
      h->codes[key].len = (uint32_t)n;
      code++;
    }
    nlist -= bits;
  }
}

static void he_generate(huffman_encoder *h, const int32_t *freq, int nfreq, int max_bits) /* :295-343 */
{
  literal_node *list = h->freqcache;
  int count = 0;
  for (int i = 0; i < nfreq; i++) {
    if (freq[i] != 0) {
      list[count].literal = (uint32_t)(i & 0xffff);
      list[count].freq = freq[i];
      count++;
    } else {
      h->codes[i].len = 0;
    }
  }
  if (count <= 2) {
    for (int i = 0; i < count; i++) {
      uint32_t key = list[i].literal & 0xffff;
      h->codes[key].code = (uint32_t)(i & 0xffff);
      h->codes[key].len = 1;
    }
    return;
  }
  qsort(list, (size_t)count, sizeof *list, cmp_by_frequency);
  int mb = he_bit_counts(h, list, count, max_bits);
  he_assign_encoding_and_size(h, mb + 1, list, count);
}

void orc_huff_generate(const int32_t *freq, int nfreq, int max_bits, uint8_t *len_out, uint16_t *code_out)
{
  huffman_encoder *h = (huffman_encoder *)malloc(sizeof *h);
  he_new(h, nfreq);
  he_generate(h, freq, nfreq, max_bits);
  for (int i = 0; i < nfreq; i++) {
    len_out[i] = (uint8_t)h->codes[i].len;
    if (code_out) code_out[i] = (uint16_t)h->codes[i].code;
  }
  free(h);
}

/* static huff_offset (huffman-code.mbt:691-726): symbol 0 has length 1 */
static huffman_encoder g_huff_offset;
__attribute__((constructor)) static void orc_init_huff_offset(void)
{
  he_new(&g_huff_offset, OFFSET_CODE_COUNT);
  g_huff_offset.codes[0].code = 0;
  g_huff_offset.codes[0].len = 1;
}

/* ================================================================== */
/* huffman-bit-writer.mbt                                              */
typedef struct { /* :88-109 */
  bytevec out;   /* stands for the underlying &@io.Writer */
  uint64_t bits;
  uint32_t nbits;
  int32_t codegen_freq[CODEGEN_CODE_COUNT];
  int32_t literal_freq[MAX_NUM_LIT];
  int32_t offset_freq[OFFSET_CODE_COUNT];
  uint8_t codegen[MAX_NUM_LIT + OFFSET_CODE_COUNT + 1];
  huffman_encoder literal_encoding, offset_encoding, codegen_encoding;
  int err_internal;
} bit_writer;

static void bw_new(bit_writer *w) /* :112-136 */
{
  memset(w, 0, sizeof *w);
  he_new(&w->literal_encoding, MAX_NUM_LIT);
  he_new(&w->codegen_encoding, CODEGEN_CODE_COUNT);
  he_new(&w->offset_encoding, OFFSET_CODE_COUNT);
}

/* total bits written so far (bytes handed to the sink + pending bits).  The
 * reference's 240-byte staging (:37-44,:193-196) does not affect bytes, so the
 * restatement hands each 6-byte spill straight to the sink. */
static inline uint64_t bw_bitpos(const bit_writer *w) { return (uint64_t)w->out.len * 8 + w->nbits; }

static void bw_flush(bit_writer *w) /* :139-158 */
{
  while (w->nbits != 0) {
    uint8_t b = (uint8_t)w->bits;
    bv_push(&w->out, &b, 1);
    w->bits >>= 8;
    if (w->nbits > 8) w->nbits -= 8;
    else w->nbits = 0;
  }
  w->bits = 0;
}

static inline void bw_spill48(bit_writer *w) /* :181-198, :394-411 */
{
  if (w->nbits >= 48) {
    uint8_t b[6];
    uint64_t bits = w->bits;
    for (int i = 0; i < 6; i++) b[i] = (uint8_t)(bits >> (8 * i));
    w->bits >>= 48;
    w->nbits -= 48;
    bv_push(&w->out, b, 6);
  }
}
static inline void bw_write_bits(bit_writer *w, int32_t b, uint32_t nb) /* :170-199 */
{
  w->bits |= (uint64_t)(uint32_t)b << w->nbits;
  w->nbits += nb;
  bw_spill48(w);
}
static inline void bw_write_code(bit_writer *w, hcode c) /* :387-412 */
{
  w->bits |= (uint64_t)c.code << w->nbits;
  w->nbits += c.len;
  bw_spill48(w);
}
static void bw_write_bytes(bit_writer *w, const uint8_t *p, size_t n) /* :202-225 */
{
  if ((w->nbits & 7) != 0) {
    w->err_internal = 1; /* "write_bytes with unfinished bits" */
    return;
  }
  while (w->nbits != 0) {
    uint8_t b = (uint8_t)w->bits;
    bv_push(&w->out, &b, 1);
    w->bits >>= 8;
    w->nbits -= 8;
  }
  bv_push(&w->out, p, n);
}

static void bw_generate_codegen(bit_writer *w, int num_literals, int num_offsets,
                                const huffman_encoder *lit_enc, const huffman_encoder *off_enc) /* :241-330 */
{
  for (int i = 0; i < CODEGEN_CODE_COUNT; i++) w->codegen_freq[i] = 0;
  uint8_t *codegen = w->codegen;
  for (int i = 0; i < num_literals; i++) codegen[i] = (uint8_t)lit_enc->codes[i].len;
  for (int i = 0; i < num_offsets; i++) codegen[num_literals + i] = (uint8_t)off_enc->codes[i].len;
  codegen[num_literals + num_offsets] = BAD_CODE;

  uint8_t size = codegen[0];
  int count = 1;
  int out_index = 0;
  for (int in_index = 1; size != BAD_CODE; in_index++) {
    uint8_t next_size = codegen[in_index];
    if (next_size == size) {
      count++;
      continue;
    }
    if (size != 0) {
      codegen[out_index++] = size;
      w->codegen_freq[size]++;
      count--;
      while (count >= 3) {
        int n = 6;
        if (n > count) n = count;
        codegen[out_index++] = 16;
        codegen[out_index++] = (uint8_t)(n - 3);
        w->codegen_freq[16]++;
        count -= n;
      }
    } else {
      while (count >= 11) {
        int n = 138;
        if (n > count) n = count;
        codegen[out_index++] = 18;
        codegen[out_index++] = (uint8_t)(n - 11);
        w->codegen_freq[18]++;
        count -= n;
      }
      if (count >= 3) {
        codegen[out_index++] = 17;
        codegen[out_index++] = (uint8_t)(count - 3);
        w->codegen_freq[17]++;
        count = 0;
      }
    }
    count--;
    for (; count >= 0; count--) {
      codegen[out_index++] = size;
      w->codegen_freq[size]++;
    }
    size = next_size;
    count = 1;
  }
  codegen[out_index] = BAD_CODE;
}

static int bw_dynamic_size(bit_writer *w, const huffman_encoder *lit_enc, const huffman_encoder *off_enc,
                           int extra_bits, int *num_codegens_out) /* :335-360 */
{
  int num_codegens = CODEGEN_CODE_COUNT;
  while (num_codegens > 4 && w->codegen_freq[codegen_order[num_codegens - 1]] == 0) num_codegens--;
  int32_t header = 3 + 5 + 5 + 4 + 3 * num_codegens +
                   he_bit_length(&w->codegen_encoding, w->codegen_freq, CODEGEN_CODE_COUNT) +
                   w->codegen_freq[16] * 2 + w->codegen_freq[17] * 3 + w->codegen_freq[18] * 7;
  int32_t size = header + he_bit_length(lit_enc, w->literal_freq, MAX_NUM_LIT) +
                 he_bit_length(off_enc, w->offset_freq, OFFSET_CODE_COUNT) + extra_bits;
  *num_codegens_out = num_codegens;
  return size;
}

static int stored_size(int inp_length, int *storable) /* :375-384 */
{
  if (inp_length == 0) {
    *storable = 0;
    return 0;
  }
  if (inp_length <= MAX_STORE_BLOCK_SIZE) {
    *storable = 1;
    return (inp_length + 5) * 8;
  }
  *storable = 0;
  return 0;
}

static void bw_write_dynamic_header(bit_writer *w, int num_literals, int num_offsets, int num_codegens,
                                    int is_eof) /* :421-471 */
{
  int first_bits = 4;
  if (is_eof) first_bits = 5;
  bw_write_bits(w, first_bits, 3);
  bw_write_bits(w, num_literals - 257, 5);
  bw_write_bits(w, num_offsets - 1, 5);
  bw_write_bits(w, num_codegens - 4, 4);
  for (int i = 0; i < num_codegens; i++)
    bw_write_bits(w, (int32_t)w->codegen_encoding.codes[codegen_order[i]].len, 3);
  int i = 0;
  for (;;) {
    int code_word = w->codegen[i++];
    if (code_word == BAD_CODE) break;
    bw_write_code(w, w->codegen_encoding.codes[code_word]);
    switch (code_word) {
    case 16: bw_write_bits(w, w->codegen[i++], 2); break;
    case 17: bw_write_bits(w, w->codegen[i++], 3); break;
    case 18: bw_write_bits(w, w->codegen[i++], 7); break;
    default: break;
    }
  }
}

static void bw_write_stored_header(bit_writer *w, int length, int is_eof) /* :474-487 */
{
  bw_write_bits(w, is_eof ? 1 : 0, 3);
  bw_flush(w);
  bw_write_bits(w, length, 16);
  bw_write_bits(w, (~length) & 0xffff, 16);
}

static void bw_index_tokens(bit_writer *w, const uint32_t *tokens, size_t ntok, int *num_literals,
                            int *num_offsets) /* :550-593 */
{
  for (int i = 0; i < MAX_NUM_LIT; i++) w->literal_freq[i] = 0;
  for (int i = 0; i < OFFSET_CODE_COUNT; i++) w->offset_freq[i] = 0;
  for (size_t k = 0; k < ntok; k++) {
    uint32_t t = tokens[k];
    if (t < MATCH_TYPE) {
      w->literal_freq[t]++;
      continue;
    }
    uint32_t length = tok_length(t);
    uint32_t offset = tok_offset(t);
    w->literal_freq[LENGTH_CODES_START + length_code(length)]++;
    w->offset_freq[offset_code(offset)]++;
  }
  int nl = MAX_NUM_LIT;
  while (w->literal_freq[nl - 1] == 0) nl--;
  int no = OFFSET_CODE_COUNT;
  while (no > 0 && w->offset_freq[no - 1] == 0) no--;
  if (no == 0) {
    w->offset_freq[0] = 1;
    no = 1;
  }
  he_generate(&w->literal_encoding, w->literal_freq, MAX_NUM_LIT, 15);
  he_generate(&w->offset_encoding, w->offset_freq, OFFSET_CODE_COUNT, 15);
  *num_literals = nl;
  *num_offsets = no;
}

static void bw_write_tokens(bit_writer *w, const uint32_t *tokens, size_t ntok, const hcode *le_codes,
                            const hcode *oe_codes) /* :596-731 */
{
  for (size_t k = 0; k < ntok; k++) {
    uint32_t t = tokens[k];
    if (t < MATCH_TYPE) {
      bw_write_code(w, le_codes[t]);
      continue;
    }
    uint32_t length = tok_length(t);
    int lc = length_code(length);
    bw_write_code(w, le_codes[lc + LENGTH_CODES_START]);
    uint32_t extra_length_bits = (uint32_t)length_extra_bits[lc];
    if (extra_length_bits > 0) bw_write_bits(w, (int32_t)(length - length_base[lc]), extra_length_bits);
    uint32_t offset = tok_offset(t);
    int oc = offset_code(offset);
    bw_write_code(w, oe_codes[oc]);
    uint32_t extra_offset_bits = (uint32_t)offset_extra_bits[oc];
    if (extra_offset_bits > 0) bw_write_bits(w, (int32_t)(offset - offset_base[oc]), extra_offset_bits);
  }
}

/* :496-542.  tokens has room for one more element (the EOB push of :507). */
static void bw_write_block_dynamic(bit_writer *w, tokvec *tokens, int eof, const uint8_t *input, int input_len)
{
  tv_push(tokens, END_BLOCK_MARKER);
  int num_literals, num_offsets, num_codegens;
  bw_index_tokens(w, tokens->p, tokens->len, &num_literals, &num_offsets);
  bw_generate_codegen(w, num_literals, num_offsets, &w->literal_encoding, &w->offset_encoding);
  he_generate(&w->codegen_encoding, w->codegen_freq, CODEGEN_CODE_COUNT, 7);
  int size = bw_dynamic_size(w, &w->literal_encoding, &w->offset_encoding, 0, &num_codegens);
  int storable;
  int ssize = stored_size(input_len, &storable);
  if (storable && ssize < ((size + size) >> 4)) { /* D2 */
    bw_write_stored_header(w, input_len, eof);
    bw_write_bytes(w, input, (size_t)input_len);
    return;
  }
  bw_write_dynamic_header(w, num_literals, num_offsets, num_codegens, eof);
  bw_write_tokens(w, tokens->p, tokens->len, w->literal_encoding.codes, w->offset_encoding.codes);
}

/* :738-824, histogram :831-836 */
static int bw_write_block_huff(bit_writer *w, int eof, const uint8_t *input, int input_len)
{
  for (int i = 0; i < MAX_NUM_LIT; i++) w->literal_freq[i] = 0;
  for (int i = 0; i < input_len; i++) w->literal_freq[input[i]]++;
  w->literal_freq[END_BLOCK_MARKER] = 1;
  int num_literals = END_BLOCK_MARKER + 1;
  w->offset_freq[0] = 1; /* the other 29 entries keep whatever the last block left */
  int num_offsets = 1;
  he_generate(&w->literal_encoding, w->literal_freq, MAX_NUM_LIT, 15);
  bw_generate_codegen(w, num_literals, num_offsets, &w->literal_encoding, &g_huff_offset);
  he_generate(&w->codegen_encoding, w->codegen_freq, CODEGEN_CODE_COUNT, 7);
  int num_codegens;
  int size = bw_dynamic_size(w, &w->literal_encoding, &g_huff_offset, 0, &num_codegens);
  int storable;
  int ssize = stored_size(input_len, &storable);
  if (storable && ssize < ((size + size) >> 4)) { /* D2 */
    bw_write_stored_header(w, input_len, eof);
    bw_write_bytes(w, input, (size_t)input_len);
    return ORC_BLK_STORED;
  }
  bw_write_dynamic_header(w, num_literals, num_offsets, num_codegens, eof);
  const hcode *encoding = w->literal_encoding.codes;
  for (int i = 0; i < input_len; i++) bw_write_code(w, encoding[input[i]]);
  bw_write_code(w, encoding[END_BLOCK_MARKER]);
  return ORC_BLK_HUFF;
}

/* ================================================================== */
/* deflate.mbt: Compressor                                             */
typedef struct {
  uint32_t *tokens;
  size_t tok_len, tok_cap;
  uint32_t *blk_ntok;
  uint8_t *blk_kind;
  uint64_t *blk_bits;
  size_t nblk, blk_cap;
  int overflow;
} recorder;

struct orc_writer { /* deflate.mbt:46-78 (vestigial hash-chain fields omitted) */
  bit_writer w;
  deflate_fast best_speed;
  uint8_t window[MAX_STORE_BLOCK_SIZE];
  int window_end;
  int sync;
  tokvec tokens;
  int closed; /* err == writer_closed_error */
  recorder *rec;
};

static void rec_block(orc_writer *c, int kind, const uint32_t *tok, size_t ntok, uint64_t bits)
{
  recorder *r = c->rec;
  if (!r) return;
  if (r->nblk >= r->blk_cap) {
    r->overflow = 1;
    return;
  }
  if (r->blk_ntok) r->blk_ntok[r->nblk] = (uint32_t)ntok;
  if (r->blk_kind) r->blk_kind[r->nblk] = (uint8_t)kind;
  if (r->blk_bits) r->blk_bits[r->nblk] = bits;
  r->nblk++;
  if (r->tokens) {
    if (r->tok_len + ntok > r->tok_cap) {
      r->overflow = 1;
      return;
    }
    memcpy(r->tokens + r->tok_len, tok, ntok * sizeof(uint32_t));
    r->tok_len += ntok;
  }
}

orc_writer *orc_writer_new(void) /* writer.mbt:10-15, deflate.mbt:81-100 */
{
  orc_writer *c = (orc_writer *)calloc(1, sizeof *c);
  bw_new(&c->w);
  df_new(&c->best_speed);
  return c;
}

static int comp_fill_store(orc_writer *c, const uint8_t *b, size_t n) /* deflate.mbt:222-229 */
{
  size_t room = (size_t)(MAX_STORE_BLOCK_SIZE - c->window_end);
  size_t k = room < n ? room : n;
  memcpy(c->window + c->window_end, b, k);
  c->window_end += (int)k;
  return (int)k;
}

static void comp_enc_speed(orc_writer *c) /* deflate.mbt:236-277 */
{
  uint64_t bit0 = bw_bitpos(&c->w);
  if (c->window_end < MAX_STORE_BLOCK_SIZE) {
    if (!c->sync) return;
    if (c->window_end < 128) {
      if (c->window_end == 0) return;
      if (c->window_end <= 16) { /* write_stored_block, deflate.mbt:186-196 */
        bw_write_stored_header(&c->w, c->window_end, 0);
        bw_write_bytes(&c->w, c->window, (size_t)c->window_end);
        rec_block(c, ORC_BLK_STORED, NULL, 0, bw_bitpos(&c->w) - bit0);
      } else {
        int k = bw_write_block_huff(&c->w, 0, c->window, c->window_end);
        rec_block(c, k, NULL, 0, bw_bitpos(&c->w) - bit0);
      }
      c->window_end = 0;
      df_reset(&c->best_speed);
      return;
    }
  }
  c->tokens.len = 0;
  df_encode(&c->best_speed, &c->tokens, c->window, c->window_end);
  size_t ntok = c->tokens.len;
  if ((int)ntok > c->window_end - (c->window_end >> 4)) {
    int k = bw_write_block_huff(&c->w, 0, c->window, c->window_end);
    rec_block(c, k, c->tokens.p, ntok, bw_bitpos(&c->w) - bit0);
  } else {
    bw_write_block_dynamic(&c->w, &c->tokens, 0, c->window, c->window_end);
    rec_block(c, ORC_BLK_DYNAMIC, c->tokens.p, ntok, bw_bitpos(&c->w) - bit0);
  }
  c->window_end = 0;
}

int64_t orc_writer_write(orc_writer *c, const uint8_t *b, size_t n) /* deflate.mbt:280-294 */
{
  if (c->closed) return -1;
  size_t total = n;
  while (n > 0) {
    comp_enc_speed(c);
    int k = comp_fill_store(c, b, n);
    b += k;
    n -= (size_t)k;
  }
  return (int64_t)total;
}

orc_writer *orc_writer_new_dict(const uint8_t *dict, size_t n) /* writer.mbt:25-31, deflate.mbt:108-151 */
{
  orc_writer *c = orc_writer_new();
  if (n > (size_t)MAX_MATCH_OFFSET) { /* window_size = 1<<15, deflate.mbt:116-118 */
    dict += n - MAX_MATCH_OFFSET;
    n = MAX_MATCH_OFFSET;
  }
  memcpy(c->window, dict, n); /* slice_copy(self.window, b); hash chains are vestigial */
  c->window_end = (int)n;
  return c;
}

int orc_writer_close(orc_writer *c) /* deflate.mbt:157-183 */
{
  if (c->closed) return 0;
  c->sync = 1;
  comp_enc_speed(c);
  bw_write_stored_header(&c->w, 0, 1);
  bw_flush(&c->w);
  c->closed = 1;
  return 0;
}

const uint8_t *orc_writer_data(const orc_writer *c, size_t *len)
{
  *len = c->w.out.len;
  return c->w.out.p;
}

void orc_writer_free(orc_writer *c)
{
  if (!c) return;
  free(c->w.out.p);
  free(c->tokens.p);
  free(c);
}

size_t orc_deflate_bound(size_t n)
{
  /* every code <= 15 bits, a match token (>= 4 bytes) <= 48 bits, block header < 400 B */
  size_t nblk = n / MAX_STORE_BLOCK_SIZE + 1;
  return 2 * n + 1024 * nblk + 16;
}

int64_t orc_deflate_ex(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len,
                       uint32_t *tokens, size_t tok_cap, uint32_t *blk_ntok, uint8_t *blk_kind,
                       uint64_t *blk_bits, size_t blk_cap)
{
  recorder r;
  memset(&r, 0, sizeof r);
  r.tokens = tokens;
  r.tok_cap = tok_cap;
  r.blk_ntok = blk_ntok;
  r.blk_kind = blk_kind;
  r.blk_bits = blk_bits;
  r.blk_cap = blk_cap;
  orc_writer *c = orc_writer_new();
  c->rec = &r;
  orc_writer_write(c, src, n);
  orc_writer_close(c);
  int64_t ret = (int64_t)r.nblk;
  size_t len = c->w.out.len;
  if (out_len) *out_len = len;
  if (dst) {
    if (len > cap) ret = -1;
    else memcpy(dst, c->w.out.p, len);
  }
  if (r.overflow) ret = -1;
  orc_writer_free(c);
  return ret;
}

int64_t orc_deflate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
  orc_writer *c = orc_writer_new();
  orc_writer_write(c, src, n);
  orc_writer_close(c);
  int64_t len = (int64_t)c->w.out.len;
  if ((size_t)len > cap) len = -1;
  else memcpy(dst, c->w.out.p, (size_t)len);
  orc_writer_free(c);
  return len;
}

/* ================================================================== */
/* dict-decoder.mbt                                                    */
struct orc_dict { /* :29-36 */
  uint8_t *hist;
  int size;
  int wr_pos, rd_pos;
  int full;
};

static int slice_copy(uint8_t *dst, int dlen, const uint8_t *src, int slen) /* :188-194 */
{
  int n = dlen < slen ? dlen : slen;
  for (int i = 0; i < n; i++) dst[i] = src[i];
  return n;
}

static void dd_init(orc_dict *dd, int size, const uint8_t *dict, size_t n) /* :42-60 */
{
  dd->hist = (uint8_t *)calloc((size_t)size, 1);
  dd->size = size;
  dd->wr_pos = dd->rd_pos = 0;
  dd->full = 0;
  if (n > (size_t)size) {
    dict += n - (size_t)size;
    n = (size_t)size;
  }
  dd->wr_pos = slice_copy(dd->hist, size, dict, (int)n);
  if (dd->wr_pos == size) {
    dd->wr_pos = 0;
    dd->full = 1;
  }
  dd->rd_pos = dd->wr_pos;
}
static inline int dd_hist_size(const orc_dict *dd) { return dd->full ? dd->size : dd->wr_pos; } /* :64-69 */
static inline int dd_avail_read(const orc_dict *dd) { return dd->wr_pos - dd->rd_pos; }        /* :73-75 */
static inline int dd_avail_write(const orc_dict *dd) { return dd->size - dd->wr_pos; }         /* :79-81 */
static inline void dd_write_byte(orc_dict *dd, uint8_t c) { dd->hist[dd->wr_pos++] = c; }      /* :103-106 */

static int dd_write_copy(orc_dict *dd, int dist, int length) /* :114-154 */
{
  int dst_base = dd->wr_pos;
  int dst_pos = dst_base;
  int src_pos = dst_pos - dist;
  int end_pos = dst_pos + length;
  if (end_pos > dd->size) end_pos = dd->size;
  if (src_pos < 0) {
    src_pos += dd->size;
    dst_pos += slice_copy(dd->hist + dst_pos, end_pos - dst_pos, dd->hist + src_pos, dd->size - src_pos);
    src_pos = 0;
  }
  while (dst_pos < end_pos)
    dst_pos += slice_copy(dd->hist + dst_pos, end_pos - dst_pos, dd->hist + src_pos, dst_pos - src_pos);
  dd->wr_pos = dst_pos;
  return dst_pos - dst_base;
}

static int dd_try_write_copy(orc_dict *dd, int dist, int length) /* :163-185 */
{
  int dst_pos = dd->wr_pos;
  int end_pos = dst_pos + length;
  if (dst_pos < dist || end_pos > dd->size) return 0;
  int dst_base = dst_pos;
  int src_pos = dst_pos - dist;
  while (dst_pos < end_pos)
    dst_pos += slice_copy(dd->hist + dst_pos, end_pos - dst_pos, dd->hist + src_pos, dst_pos - src_pos);
  dd->wr_pos = dst_pos;
  return dst_pos - dst_base;
}

/* :200-209; returns a view (pointer + length) */
static const uint8_t *dd_read_flush(orc_dict *dd, int *len)
{
  const uint8_t *p = dd->hist + dd->rd_pos;
  *len = dd->wr_pos - dd->rd_pos;
  dd->rd_pos = dd->wr_pos;
  if (dd->wr_pos == dd->size) {
    dd->wr_pos = 0;
    dd->rd_pos = 0;
    dd->full = 1;
  }
  return p;
}

orc_dict *orc_dict_new(int size, const uint8_t *dict, size_t n)
{
  orc_dict *d = (orc_dict *)calloc(1, sizeof *d);
  dd_init(d, size, dict, n);
  return d;
}
void orc_dict_free(orc_dict *d)
{
  if (!d) return;
  free(d->hist);
  free(d);
}
int orc_dict_hist_size(const orc_dict *d) { return dd_hist_size(d); }
int orc_dict_avail_read(const orc_dict *d) { return dd_avail_read(d); }
int orc_dict_avail_write(const orc_dict *d) { return dd_avail_write(d); }
int orc_dict_write(orc_dict *d, const uint8_t *p, int n) /* write_slice :87-89 + write_mark :96-98 */
{
  int cnt = slice_copy(d->hist + d->wr_pos, d->size - d->wr_pos, p, n);
  d->wr_pos += cnt;
  return cnt;
}
void orc_dict_write_byte(orc_dict *d, uint8_t c) { dd_write_byte(d, c); }
int orc_dict_write_copy(orc_dict *d, int dist, int length) { return dd_write_copy(d, dist, length); }
int orc_dict_try_write_copy(orc_dict *d, int dist, int length) { return dd_try_write_copy(d, dist, length); }
int orc_dict_read_flush(orc_dict *d, uint8_t *out)
{
  int len;
  const uint8_t *p = dd_read_flush(d, &len);
  memcpy(out, p, (size_t)len);
  return len;
}

/* ================================================================== */
/* inflate.mbt                                                         */
#define MAX_LINKS (1 << (15 - HUFFMAN_CHUNK_BITS))
typedef struct { /* :81-86 */
  int min;
  uint32_t chunks[HUFFMAN_NUM_CHUNKS];
  uint32_t links[HUFFMAN_NUM_CHUNKS][MAX_LINKS];
  int nlinks;     /* links.length() */
  int link_width; /* length of every links[i] */
  uint32_t link_mask;
} huffman_decoder;

static void hd_clear(huffman_decoder *h) /* :89-92 */
{
  h->min = 0;
  memset(h->chunks, 0, sizeof h->chunks);
  h->nlinks = 0;
  h->link_width = 0;
  h->link_mask = 0;
}

static int hd_initialize(huffman_decoder *h, const int *lengths, int nlen) /* :100-223 */
{
  if (h->min != 0) hd_clear(h);
  int count[MAX_CODE_LEN];
  memset(count, 0, sizeof count);
  int min = 0, max = 0;
  for (int i = 0; i < nlen; i++) {
    int n = lengths[i];
    if (n == 0) continue;
    if (min == 0 || n < min) min = n;
    if (n > max) max = n;
    count[n]++;
  }
  if (max == 0) return 1;
  int code = 0;
  int nextcode[MAX_CODE_LEN];
  memset(nextcode, 0, sizeof nextcode);
  for (int i = min; i <= max; i++) {
    code <<= 1;
    nextcode[i] = code;
    code += count[i];
  }
  if (code != (1 << max) && !(code == 1 && max == 1)) return 0;

  h->min = min;
  if (max > HUFFMAN_CHUNK_BITS) {
    uint32_t num_links = 1u << (max - HUFFMAN_CHUNK_BITS);
    h->link_mask = num_links - 1;
    int link = nextcode[HUFFMAN_CHUNK_BITS + 1] >> 1;
    h->nlinks = HUFFMAN_NUM_CHUNKS - link;
    h->link_width = (int)num_links;
    for (uint32_t j = (uint32_t)link; j < HUFFMAN_NUM_CHUNKS; j++) {
      int reverse = (int)orc_reverse16(j & 0xffff);
      reverse >>= (16 - HUFFMAN_CHUNK_BITS);
      uint32_t off = j - (uint32_t)link;
      h->chunks[reverse] = (off << HUFFMAN_VALUE_SHIFT) | (uint32_t)(HUFFMAN_CHUNK_BITS + 1);
      memset(h->links[off], 0, num_links * sizeof(uint32_t));
    }
  }
  for (int i = 0; i < nlen; i++) {
    int n = lengths[i];
    if (n == 0) continue;
    int c = nextcode[n];
    nextcode[n]++;
    uint32_t chunk = ((uint32_t)i << HUFFMAN_VALUE_SHIFT) | (uint32_t)n;
    int reverse = (int)orc_reverse16((uint32_t)c & 0xffff);
    reverse >>= (16 - n);
    if (n <= HUFFMAN_CHUNK_BITS) {
      for (int off = reverse; off < HUFFMAN_NUM_CHUNKS; off += 1 << n) h->chunks[off] = chunk;
    } else {
      int j = reverse & (HUFFMAN_NUM_CHUNKS - 1);
      uint32_t value = h->chunks[j] >> HUFFMAN_VALUE_SHIFT;
      uint32_t *linktab = h->links[value];
      reverse >>= HUFFMAN_CHUNK_BITS;
      for (int off = reverse; off < h->link_width; off += 1 << (n - HUFFMAN_CHUNK_BITS)) linktab[off] = chunk;
    }
  }
  return 1;
}

/* fixed_huffman_decoder (:886-939).  The reference lists the 512 chunks; they
 * are what initialize() yields for the RFC 1951 3.2.6 lengths with min = 7
 * (tests compare spot values with the reference listing). */
static huffman_decoder g_fixed;
__attribute__((constructor)) static void orc_init_fixed(void)
{
  int bits[288];
  for (int i = 0; i < 144; i++) bits[i] = 8;
  for (int i = 144; i < 256; i++) bits[i] = 9;
  for (int i = 256; i < 280; i++) bits[i] = 7;
  for (int i = 280; i < 288; i++) bits[i] = 8;
  hd_clear(&g_fixed);
  hd_initialize(&g_fixed, bits, 288);
  g_fixed.min = 7;
}
uint32_t orc_fixed_chunk(int i) { return g_fixed.chunks[i & (HUFFMAN_NUM_CHUNKS - 1)]; }

enum { ERR_NONE = -1 };
enum { STEP_NEXT_BLOCK, STEP_HUFFMAN_BLOCK, STEP_COPY_DATA };
enum { STATE_INIT, STATE_DICT };

struct orc_reader { /* :257-291 */
  const uint8_t *in; /* stands for r : &Reader */
  size_t in_len, in_pos;
  int64_t roffset;
  uint32_t b, nb;
  huffman_decoder h1, h2;
  int bits[MAX_NUM_LIT + MAX_NUM_DIST];
  int codebits[NUM_CODES];
  orc_dict dict;
  uint8_t buf[4];
  int step, step_state;
  int final_flag;
  int err; /* ERR_NONE or ORC_* (ORC_OK == ioeof); 5 == ioeof from an exhausted more_bits (D5) */
  int64_t err_off;
  const uint8_t *to_read;
  int to_read_len;
  const huffman_decoder *hl, *hd;
  int copy_len, copy_dist;
};

static orc_reader *rd_alloc(const uint8_t *comp, size_t n, const uint8_t *dict, size_t dn) /* :320-342 */
{
  orc_reader *f = (orc_reader *)calloc(1, sizeof *f);
  f->in = comp;
  f->in_len = n;
  hd_clear(&f->h1);
  hd_clear(&f->h2);
  dd_init(&f->dict, MAX_MATCH_OFFSET, dict, dn);
  f->step = STEP_NEXT_BLOCK;
  f->step_state = STATE_INIT;
  f->err = ERR_NONE;
  return f;
}
orc_reader *orc_reader_new(const uint8_t *comp, size_t n) { return rd_alloc(comp, n, NULL, 0); }
orc_reader *orc_reader_new_dict(const uint8_t *comp, size_t n, const uint8_t *dict, size_t dn)
{
  return rd_alloc(comp, n, dict, dn);
}
void orc_reader_free(orc_reader *f)
{
  if (!f) return;
  free(f->dict.hist);
  free(f);
}
int64_t orc_reader_roffset(const orc_reader *f) { return f->roffset; }

static void set_corrupt(orc_reader *f) /* corrupt_input_error(self.roffset), :38-40 */
{
  f->err = ORC_CORRUPT;
  f->err_off = f->roffset;
}

/* :789-799.  Returns 0, or 1 when the reader is exhausted: the reference
 * hands back the reader's own eof, NOT no_eof(...) (D5). */
static int more_bits(orc_reader *f)
{
  if (f->in_pos >= f->in_len) return 1;
  uint8_t c = f->in[f->in_pos++];
  f->roffset++;
  f->b |= (uint32_t)c << (f->nb & 31);
  f->nb += 8;
  return 0;
}
#define NEED_BITS(f, k, onfail)             \
  while ((f)->nb < (uint32_t)(k)) {         \
    if (more_bits(f)) {                     \
      (f)->err = ORC_EOF_AT_REFILL;         \
      onfail;                               \
    }                                       \
  }

/* :803-854.  Returns the symbol, or -1 with f->err set. */
static int huff_sym(orc_reader *f, const huffman_decoder *h)
{
  uint32_t n = (uint32_t)h->min;
  uint32_t nb = f->nb, b = f->b;
  for (;;) {
    while (nb < n) {
      if (f->in_pos >= f->in_len) {
        f->b = b;
        f->nb = nb;
        f->err = ORC_UNEXPECTED_EOF; /* no_eof(e) */
        return -1;
      }
      uint8_t c = f->in[f->in_pos++];
      f->roffset++;
      b |= (uint32_t)c << (nb & 31);
      nb += 8;
    }
    uint32_t chunk = h->chunks[b & (HUFFMAN_NUM_CHUNKS - 1)];
    n = chunk & HUFFMAN_COUNT_MASK;
    if (n > HUFFMAN_CHUNK_BITS) {
      chunk = h->links[chunk >> HUFFMAN_VALUE_SHIFT][(b >> HUFFMAN_CHUNK_BITS) & h->link_mask];
      n = chunk & HUFFMAN_COUNT_MASK;
    }
    if (n <= nb) {
      if (n == 0) {
        f->b = b;
        f->nb = nb;
        set_corrupt(f);
        return -1;
      }
      f->b = b >> (n & 31);
      f->nb = nb - n;
      return (int)(chunk >> HUFFMAN_VALUE_SHIFT);
    }
  }
}

static void finish_block(orc_reader *f) /* :769-777 */
{
  if (f->final_flag) {
    if (dd_avail_read(&f->dict) > 0) f->to_read = dd_read_flush(&f->dict, &f->to_read_len);
    f->err = ORC_OK; /* ioeof */
  }
  f->step = STEP_NEXT_BLOCK;
}

/* @io.read_full over the in-memory reader: (cnt, err) with err = eof when
 * nothing could be read and unexpected-eof on a short read; both become
 * unexpected-eof through no_eof (:781-786). */
static int read_full(orc_reader *f, uint8_t *dst, int want, int *got)
{
  size_t avail = f->in_len - f->in_pos;
  int k = (size_t)want < avail ? want : (int)avail;
  memcpy(dst, f->in + f->in_pos, (size_t)k);
  f->in_pos += (size_t)k;
  *got = k;
  return k < want;
}

static void copy_data(orc_reader *f) /* :742-766 */
{
  int room = dd_avail_write(&f->dict);
  if (room > f->copy_len) room = f->copy_len;
  int cnt;
  int short_read = read_full(f, f->dict.hist + f->dict.wr_pos, room, &cnt);
  f->roffset += cnt;
  f->copy_len -= cnt;
  f->dict.wr_pos += cnt; /* write_mark */
  if (short_read) {
    f->err = ORC_UNEXPECTED_EOF;
    return;
  }
  if (dd_avail_write(&f->dict) == 0 || f->copy_len > 0) {
    f->to_read = dd_read_flush(&f->dict, &f->to_read_len);
    f->step = STEP_COPY_DATA;
    return;
  }
  finish_block(f);
}

static void data_block(orc_reader *f) /* :708-737 */
{
  f->nb = 0;
  f->b = 0;
  int nr;
  int short_read = read_full(f, f->buf, 4, &nr);
  f->roffset += nr;
  if (short_read) {
    f->err = ORC_UNEXPECTED_EOF;
    return;
  }
  int n = f->buf[0] | (f->buf[1] << 8);
  int nn = f->buf[2] | (f->buf[3] << 8);
  if ((nn & 0xffff) != ((~n) & 0xffff)) {
    set_corrupt(f);
    return;
  }
  if (n == 0) {
    f->to_read = dd_read_flush(&f->dict, &f->to_read_len);
    finish_block(f);
    return;
  }
  f->copy_len = n;
  copy_data(f);
}

static int read_huffman(orc_reader *f) /* :429-548; returns 0 ok, 1 with f->err set */
{
  NEED_BITS(f, 5 + 5 + 4, return 1);
  int nlit = (int)(f->b & 0x1F) + 257;
  if (nlit > MAX_NUM_LIT) {
    set_corrupt(f);
    return 1;
  }
  f->b >>= 5;
  int ndist = (int)(f->b & 0x1F) + 1;
  if (ndist > MAX_NUM_DIST) {
    set_corrupt(f);
    return 1;
  }
  f->b >>= 5;
  int nclen = (int)(f->b & 0xF) + 4;
  f->b >>= 4;
  f->nb -= 5 + 5 + 4;

  for (int i = 0; i < nclen; i++) {
    NEED_BITS(f, 3, return 1);
    f->codebits[codegen_order[i]] = (int)(f->b & 0x7);
    f->b >>= 3;
    f->nb -= 3;
  }
  for (int i = nclen; i < NUM_CODES; i++) f->codebits[codegen_order[i]] = 0;
  if (!hd_initialize(&f->h1, f->codebits, NUM_CODES)) {
    set_corrupt(f);
    return 1;
  }

  int i = 0;
  int n = nlit + ndist;
  while (i < n) {
    int x = huff_sym(f, &f->h1);
    if (x < 0) return 1;
    if (x < 16) {
      f->bits[i++] = x;
      continue;
    }
    int rep;
    uint32_t nb;
    int b;
    switch (x) {
    case 16:
      rep = 3;
      nb = 2;
      if (i == 0) {
        set_corrupt(f);
        return 1;
      }
      b = f->bits[i - 1];
      break;
    case 17:
      rep = 3;
      nb = 3;
      b = 0;
      break;
    case 18:
      rep = 11;
      nb = 7;
      b = 0;
      break;
    default:
      f->err = ORC_INTERNAL; /* "unexpected length code" */
      return 1;
    }
    NEED_BITS(f, nb, return 1);
    rep += (int)(f->b & ((1u << nb) - 1));
    f->b >>= nb;
    f->nb -= nb;
    if (i + rep > n) {
      set_corrupt(f);
      return 1;
    }
    for (int j = 0; j < rep; j++) f->bits[i++] = b;
  }
  if (!hd_initialize(&f->h1, f->bits, nlit) || !hd_initialize(&f->h2, f->bits + nlit, ndist)) {
    set_corrupt(f);
    return 1;
  }
  if (f->h1.min < f->bits[END_BLOCK_MARKER]) f->h1.min = f->bits[END_BLOCK_MARKER];
  return 0;
}

static void huffman_block(orc_reader *f);

static void copy_history_then_literals(orc_reader *f, int start_in_copy);

static void read_literal(orc_reader *f) { copy_history_then_literals(f, 0); }  /* :565-684 */
static void copy_history(orc_reader *f) { copy_history_then_literals(f, 1); } /* :689-704 */

/* read_literal and copy_history call each other in tail position in the
 * reference (:584, :683, :703); restated as one loop. */
static void copy_history_then_literals(orc_reader *f, int start_in_copy)
{
  int in_copy = start_in_copy;
  for (;;) {
    if (in_copy) { /* copy_history :689-704 */
      int cnt = dd_try_write_copy(&f->dict, f->copy_dist, f->copy_len);
      if (cnt == 0) cnt = dd_write_copy(&f->dict, f->copy_dist, f->copy_len);
      f->copy_len -= cnt;
      if (dd_avail_write(&f->dict) == 0 || f->copy_len > 0) {
        f->to_read = dd_read_flush(&f->dict, &f->to_read_len);
        f->step = STEP_HUFFMAN_BLOCK;
        f->step_state = STATE_DICT;
        return;
      }
      in_copy = 0;
    }
    /* read_literal :565-684 */
    int v = huff_sym(f, f->hl);
    if (v < 0) return;
    uint32_t n;
    int length;
    if (v < 256) {
      dd_write_byte(&f->dict, (uint8_t)v);
      if (dd_avail_write(&f->dict) == 0) {
        f->to_read = dd_read_flush(&f->dict, &f->to_read_len);
        f->step = STEP_HUFFMAN_BLOCK;
        f->step_state = STATE_INIT;
        return;
      }
      continue;
    }
    if (v == 256) {
      finish_block(f);
      return;
    }
    if (v < 265) {
      length = v - (257 - 3);
      n = 0;
    } else if (v < 269) {
      length = v * 2 - (265 * 2 - 11);
      n = 1;
    } else if (v < 273) {
      length = v * 4 - (269 * 4 - 19);
      n = 2;
    } else if (v < 277) {
      length = v * 8 - (273 * 8 - 35);
      n = 3;
    } else if (v < 281) {
      length = v * 16 - (277 * 16 - 67);
      n = 4;
    } else if (v < 285) {
      length = v * 32 - (281 * 32 - 131);
      n = 5;
    } else if (v < MAX_NUM_LIT) {
      length = 258;
      n = 0;
    } else {
      set_corrupt(f);
      return;
    }
    if (n > 0) {
      NEED_BITS(f, n, return);
      length += (int)(f->b & ((1u << n) - 1));
      f->b >>= n;
      f->nb -= n;
    }
    int dist;
    if (f->hd == NULL) {
      NEED_BITS(f, 5, return);
      dist = (int)rev8((f->b & 0x1F) << 3);
      f->b >>= 5;
      f->nb -= 5;
    } else {
      dist = huff_sym(f, f->hd);
      if (dist < 0) return;
    }
    if (dist < 4) {
      dist++;
    } else if (dist < MAX_NUM_DIST) {
      uint32_t nb = (uint32_t)(dist - 2) >> 1;
      int extra = (dist & 1) << nb;
      NEED_BITS(f, nb, return);
      extra |= (int)(f->b & ((1u << nb) - 1));
      f->b >>= nb;
      f->nb -= nb;
      dist = (1 << (nb + 1)) + 1 + extra;
    } else {
      set_corrupt(f);
      return;
    }
    if (dist > dd_hist_size(&f->dict)) {
      set_corrupt(f);
      return;
    }
    f->copy_len = length;
    f->copy_dist = dist;
    in_copy = 1;
  }
}

static void huffman_block(orc_reader *f) /* :555-560 */
{
  if (f->step_state == STATE_INIT) read_literal(f);
  else copy_history(f);
}

static void next_block(orc_reader *f) /* :345-379 */
{
  NEED_BITS(f, 1 + 2, return);
  f->final_flag = (f->b & 1) == 1;
  f->b >>= 1;
  uint32_t typ = f->b & 3;
  f->b >>= 2;
  f->nb -= 1 + 2;
  switch (typ) {
  case 0: data_block(f); break;
  case 1:
    f->hl = &g_fixed;
    f->hd = NULL;
    huffman_block(f);
    break;
  case 2:
    if (read_huffman(f) == 0) {
      f->hl = &f->h1;
      f->hd = &f->h2;
      huffman_block(f);
    }
    break;
  default: set_corrupt(f); break;
  }
}

size_t orc_reader_read(orc_reader *f, uint8_t *buf, size_t n, int *status, int64_t *err_off) /* :382-407 */
{
  for (;;) {
    if (f->to_read_len > 0) {
      size_t k = (size_t)f->to_read_len;
      if (n < k) k = n;
      for (size_t i = 0; i < k; i++) buf[i] = f->to_read[i];
      f->to_read += k;
      f->to_read_len -= (int)k;
      if (f->to_read_len == 0) {
        *status = f->err;
        if (err_off) *err_off = f->err_off;
        return k;
      }
      *status = ERR_NONE;
      return k;
    }
    if (f->err != ERR_NONE) {
      *status = f->err;
      if (err_off) *err_off = f->err_off;
      return 0;
    }
    switch (f->step) {
    case STEP_NEXT_BLOCK: next_block(f); break;
    case STEP_HUFFMAN_BLOCK: huffman_block(f); break;
    default: copy_data(f); break;
    }
    if (f->err != ERR_NONE && f->to_read_len == 0) f->to_read = dd_read_flush(&f->dict, &f->to_read_len);
  }
}

int orc_inflate(const uint8_t *comp, size_t n, uint8_t *out, size_t cap, size_t *out_len, int64_t *err_off,
                int64_t *consumed)
{
  orc_reader *f = orc_reader_new(comp, n);
  size_t total = 0;
  int status = ERR_NONE;
  int64_t eo = 0;
  int overflow = 0;
  while (status == ERR_NONE) {
    size_t room = cap - total;
    size_t k;
    if (room == 0) {
      /* keep draining to learn the true status, but remember the overflow */
      uint8_t tmp[4096];
      k = orc_reader_read(f, tmp, sizeof tmp, &status, &eo);
      if (k > 0) overflow = 1;
      continue;
    }
    k = orc_reader_read(f, out + total, room, &status, &eo);
    total += k;
  }
  if (out_len) *out_len = total;
  if (err_off) *err_off = eo;
  if (consumed) *consumed = f->roffset;
  orc_reader_free(f);
  if (overflow) return ORC_DST_TOO_SMALL;
  return status;
}

/* Timing helper of bench.py's CPU arm (no reference counterpart): deflate + inflate of streams first, first + stride,
 * ... of src[off[i] .. off[i+1]) on the calling thread; 1 if every stream round-trips. */
int orc_codec_pass(const uint8_t *src, const uint64_t *off, uint64_t ns, uint64_t first, uint64_t stride)
{
  size_t maxlen = 0;
  for (uint64_t i = first; i < ns; i += stride)
    if (off[i + 1] - off[i] > maxlen) maxlen = (size_t)(off[i + 1] - off[i]);
  const size_t bound = orc_deflate_bound(maxlen);
  uint8_t *dst = (uint8_t *)malloc(bound ? bound : 1), *out = (uint8_t *)malloc(maxlen + 1);
  int ok = dst && out;
  for (uint64_t i = first; ok && i < ns; i += stride) {
    const size_t n = (size_t)(off[i + 1] - off[i]);
    const int64_t c = orc_deflate(src + off[i], n, dst, bound);
    size_t ol = 0;
    int64_t eo = 0, cons = 0;
    if (c < 0 || orc_inflate(dst, (size_t)c, out, n, &ol, &eo, &cons) != ORC_OK || ol != n || memcmp(out, src + off[i], n) != 0) ok = 0;
  }
  free(dst);
  free(out);
  return ok;
}
