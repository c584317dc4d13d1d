/* blockpar_model.c -- CPU model of the block-parallel parse of multi-block streams (BlockParJob in
 * moonbit_flate_b200/csrc/kernels.h), built on the oracle's own encode (test infrastructure only).
 *
 * The kernel's rule, restated: round 1 parses every block from an empty table; every block keeps its
 * normalised end table (per bucket the distance of the last position from the block end, 0 beyond 32768);
 * round r > 1 parses again exactly the blocks whose predecessor's end table changed in round r-1, seeded
 * from that table.  When no end table changes any more, every block's tokens must equal those of the
 * sequential parse (DeflateFast::encode called block after block on one table, deflate.mbt:236-277). */
#include "../../oracle/flate_oracle.c"

static void snap_end_table(const deflate_fast *e, uint16_t *out)
{
  for (int i = 0; i < TABLE_SIZE; i++) {
    const int64_t d = (int64_t)e->cur - e->table[i].offset; /* cur = block end once encode has returned */
    out[i] = (d >= 1 && d <= MAX_MATCH_OFFSET) ? (uint16_t)d : 0; /* 32768 fits in 16 bits */
  }
}

/* src: nblk full blocks of 65535 bytes.  Returns 1 when the fixpoint equals the sequential parse, else 0
 * (-1: did not converge within nblk + 2 rounds).  rounds / parses: work the fixpoint needed. */
int fbm_blockpar_check(const uint8_t *src, int nblk, int *rounds, long *parses)
{
  const int B = MAX_STORE_BLOCK_SIZE;
  deflate_fast *e = (deflate_fast *)malloc(sizeof *e);
  tokvec tv = {0};
  /* sequential reference */
  uint32_t **seq = (uint32_t **)calloc(nblk, sizeof *seq), **par = (uint32_t **)calloc(nblk, sizeof *par);
  size_t *seq_n = (size_t *)calloc(nblk, sizeof *seq_n), *par_n = (size_t *)calloc(nblk, sizeof *par_n);
  df_new(e);
  for (int b = 0; b < nblk; b++) {
    tv.len = 0;
    df_encode(e, &tv, src + (size_t)b * B, B);
    seq[b] = (uint32_t *)malloc(tv.len * 4 + 4);
    memcpy(seq[b], tv.p, tv.len * 4);
    seq_n[b] = tv.len;
  }
  /* fixpoint */
  uint16_t *tabs = (uint16_t *)malloc((size_t)2 * nblk * TABLE_SIZE * 2);
  uint8_t *lat = (uint8_t *)calloc(2 * (size_t)nblk, 1), *chg = (uint8_t *)calloc(2 * (size_t)nblk, 1);
  uint16_t *fresh = (uint16_t *)malloc(TABLE_SIZE * 2);
  int round = 1, ok = -1;
  long np = 0;
  for (;; round++) {
    uint8_t *lat_prev = lat + (size_t)((round - 1) & 1) * nblk, *lat_next = lat + (size_t)(round & 1) * nblk;
    uint8_t *chg_prev = chg + (size_t)((round - 1) & 1) * nblk, *chg_next = chg + (size_t)(round & 1) * nblk;
    int any = 0;
    for (int b = 0; b < nblk; b++) { lat_next[b] = round == 1 ? 0 : lat_prev[b]; chg_next[b] = 0; }
    for (int b = 0; b < nblk; b++) {
      if (!(round == 1 || (b > 0 && chg_prev[b - 1]))) continue;
      any = 1;
      np++;
      df_new(e);
      if (round > 1 && b > 0) { /* seed: what block b-1 left within reach of this block's start */
        const uint16_t *t = tabs + ((size_t)lat_prev[b - 1] * nblk + (b - 1)) * TABLE_SIZE;
        for (int i = 0; i < TABLE_SIZE; i++)
          if (t[i]) {
            e->table[i].offset = e->cur - (int32_t)t[i];
            e->table[i].val = load32(src + (size_t)b * B - t[i], 0);
          }
      }
      tv.len = 0;
      df_encode(e, &tv, src + (size_t)b * B, B);
      free(par[b]);
      par[b] = (uint32_t *)malloc(tv.len * 4 + 4);
      memcpy(par[b], tv.p, tv.len * 4);
      par_n[b] = tv.len;
      if (b + 1 < nblk) {
        const int ob = round == 1 ? 0 : lat_prev[b], nb = ob ^ 1;
        snap_end_table(e, fresh);
        const uint16_t *oldt = tabs + ((size_t)ob * nblk + b) * TABLE_SIZE;
        const int diff = round == 1 || memcmp(oldt, fresh, TABLE_SIZE * 2) != 0;
        memcpy(tabs + ((size_t)nb * nblk + b) * TABLE_SIZE, fresh, TABLE_SIZE * 2);
        lat_next[b] = (uint8_t)nb;
        chg_next[b] = (uint8_t)diff;
      }
    }
    if (!any) { round--; break; }
    if (round > nblk + 2) goto done;
  }
  ok = 1;
  for (int b = 0; b < nblk; b++)
    if (par_n[b] != seq_n[b] || memcmp(par[b], seq[b], seq_n[b] * 4) != 0) { ok = 0; break; }
done:
  if (rounds) *rounds = round;
  if (parses) *parses = np;
  for (int b = 0; b < nblk; b++) { free(seq[b]); free(par[b]); }
  free(seq); free(par); free(seq_n); free(par_n); free(tabs); free(lat); free(chg); free(fresh); free(tv.p); free(e);
  return ok;
}
