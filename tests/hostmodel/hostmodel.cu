// hostmodel.cu -- CPU-side checks of GPU kernel logic (test infrastructure).
//
// 1. Exposes the host compilation of the K3 building blocks
//    (moonbit_flate_b200/csrc/huff_build.cuh: the very code k_build_codes runs
//    per thread) so the CPU suite can diff it against the oracle.
// 2. A lane-by-lane emulation of K1's 32-wide batched probe (parse.cu): same
//    batch schedule, same intra-batch bucket resolution, same commit rule, run
//    sequentially over 32 "lanes", to check the batching argument against the
//    oracle's sequential parse without a GPU.
#include "../../moonbit_flate_b200/csrc/common.cuh"
#include "../../moonbit_flate_b200/csrc/huff_build.cuh"

#include <cstring>
#include <vector>

using namespace fb;

extern "C" void fbm_generate(const uint32_t *freq, int nsym, int max_bits, uint8_t *len, uint16_t *code)
{
  static HuffScratch S;
  warp_generate(freq, nsym, max_bits, len, code, S);
}

extern "C" int fbm_build_block(uint32_t *freq, int kind, uint32_t n, uint32_t *codeout, uint32_t *hdr_words,
                               uint32_t *hdr_nbits, uint32_t *blk_bits)
{
  static HuffScratch S;
  BlockBuild r = build_block_warp(freq, kind, n, codeout, hdr_words, S);
  *hdr_nbits = r.hdr_nbits;
  *blk_bits = r.blk_bits;
  return r.kind;
}

extern "C" void fbm_codes(uint32_t xlen, uint32_t xoff, int *out)
{
  int c, nb;
  uint32_t ex;
  length_code_of(xlen, c, nb, ex);
  out[0] = c; out[1] = nb; out[2] = (int)ex;
  offset_code_of(xoff, c, nb, ex);
  out[3] = c; out[4] = nb; out[5] = (int)ex;
}

static inline uint32_t ld32(const uint8_t *p)
{
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}

// Emulation of k_parse<MULTI> for one stream of L bytes.  tokens_out receives
// the concatenated tokens of all parsed blocks, blk_ntok[b] the per-block counts.
extern "C" int64_t fbm_parse_stream(const uint8_t *src, uint64_t L, uint32_t *tokens_out, uint64_t tok_cap,
                                    uint32_t *blk_ntok, uint64_t blk_cap)
{
  const bool MULTI = L >= (uint64_t)kBlockSize + 128;
  std::vector<uint32_t> table(kTableSize, MULTI ? 0u : 0xffffu);
  std::vector<uint32_t> sched(512);
  {
    uint32_t d = 0;
    for (int k = 0; k < 512; k++) {
      sched[k] = d;
      d = d + 1 + (d >> 5);
      if (d > (1u << 20)) d = 1u << 20;
    }
  }
  const uint32_t nblk = (uint32_t)((L + kBlockSize - 1) / kBlockSize);
  uint64_t total = 0;
  for (uint32_t b = 0; b < nblk; b++) {
    if (b >= blk_cap) return -1;
    blk_ntok[b] = 0;
    const uint64_t boff = (uint64_t)b * kBlockSize;
    const int n = (int)((L - boff) < (uint64_t)kBlockSize ? (L - boff) : (uint64_t)kBlockSize);
    if (n < 128 || L < 128) continue;
    const uint8_t *srcb = src + boff;
    uint32_t *tok = tokens_out + total;
    const uint32_t S0 = (uint32_t)boff;
    const int s_limit = n - kInputMargin;
    int s = 0, next_emit = 0;
    uint32_t ntok = 0;
    bool modeM = false;
    int loop_p0 = 0, k0 = 0;
    auto emit_lits = [&](int a, int e) {
      for (int i = a; i < e; i++) tok[ntok++] = srcb[i];
    };
    if (total + (uint64_t)n > tok_cap) return -1;
    for (;;) {
      int pos[32], cand[32];
      bool probe[32], fail[32], active[32], hit[32];
      uint32_t cv[32], h[32], old[32];
      for (int lane = 0; lane < 32; lane++) {
        int step;
        bool loopprobe;
        if (modeM) {
          if (lane == 0) { pos[lane] = s - 1; step = 0; probe[lane] = false; loopprobe = false; }
          else if (lane == 1) { pos[lane] = s; step = 0; probe[lane] = true; loopprobe = false; }
          else { pos[lane] = s + 1 + (lane - 2); step = 1; probe[lane] = true; loopprobe = true; }
        } else {
          const int k = k0 + lane;
          const uint32_t d = k < 32 ? (uint32_t)k : (k < 512 ? sched[k] : (1u << 20));
          pos[lane] = loop_p0 + (int)d;
          step = 1 + (int)(d >> 5);
          probe[lane] = true; loopprobe = true;
        }
        fail[lane] = loopprobe && (pos[lane] + step > s_limit);
        active[lane] = !fail[lane];
        cv[lane] = 0; h[lane] = 0x10000u | (uint32_t)lane; old[lane] = 0;
        if (active[lane]) {
          cv[lane] = ld32(srcb + pos[lane]);
          h[lane] = hash4(cv[lane]);
          old[lane] = table[h[lane]];
        }
      }
      unsigned hitm = 0, failm = 0;
      for (int lane = 0; lane < 32; lane++) {
        int lower = -1;
        for (int q = lane - 1; q >= 0; q--)
          if (h[q] == h[lane]) { lower = q; break; }
        bool ok;
        if (lower >= 0) {
          cand[lane] = pos[lower];
          ok = (pos[lane] - cand[lane]) <= kMaxMatchOffset;
        } else if (MULTI) {
          const uint32_t D = (S0 + (uint32_t)pos[lane] + 1u) - old[lane];
          ok = (old[lane] != 0) && (D <= (uint32_t)kMaxMatchOffset);
          cand[lane] = pos[lane] - (int)D;
        } else {
          const int D = pos[lane] - (int)old[lane];
          ok = (D >= 1) && (D <= kMaxMatchOffset);
          cand[lane] = (int)old[lane];
        }
        hit[lane] = active[lane] && probe[lane] && ok && ld32(srcb + cand[lane]) == cv[lane];
        if (hit[lane]) hitm |= 1u << lane;
        if (fail[lane]) failm |= 1u << lane;
      }
      const unsigned evt = hitm | failm;
      const int m = evt ? __builtin_ffs((int)evt) - 1 : 32;
      const bool mhit = evt && ((hitm >> m) & 1u);
      for (int lane = 0; lane < 32; lane++) { // commits in lane order == highest committed lane wins
        const bool in = (m == 32) || (mhit ? lane <= m : lane < m);
        if (active[lane] && in) table[h[lane]] = MULTI ? (S0 + (uint32_t)pos[lane] + 1u) : (uint32_t)(uint16_t)pos[lane];
      }
      if (m == 32) {
        if (modeM) { modeM = false; loop_p0 = s + 1; k0 = 30; }
        else k0 += 32;
        continue;
      }
      if (!mhit) break;
      const int s_hit = pos[m], c = cand[m];
      emit_lits(next_emit, s_hit);
      const int s2 = s_hit + 4, t = c + 4;
      int ext = 0;
      if (t >= 0) {
        int s1 = s2 + kMaxMatchLength - 4;
        if (s1 > n) s1 = n;
        const int a = s1 - s2;
        while (ext < a && srcb[s2 + ext] == srcb[t + ext]) ext++;
      }
      tok[ntok++] = kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1);
      s = s2 + ext;
      next_emit = s;
      if (s >= s_limit) break;
      modeM = true;
    }
    emit_lits(next_emit, n);
    blk_ntok[b] = ntok;
    total += ntok;
  }
  return (int64_t)total;
}

// ---------------------------------------------------------------------------
// v2: emulation of the multi-match batch of parse.cu.  A post-match batch of 32
// consecutive positions is evaluated once (bytes, hash, old table entry,
// 4-byte verify, up to 8 bytes of speculative extension per lane); if no two
// lanes share a bucket, every lane's old-entry candidate is what sequential
// execution would read no matter which lanes end up inserted, so the batch is
// consumed match after match from registers.  Otherwise (and outside post-match
// batches) the single-event batch of v1 runs.
extern "C" int64_t fbm_parse_stream_v2(const uint8_t *src, uint64_t L, uint32_t *tokens_out, uint64_t tok_cap,
                                       uint32_t *blk_ntok, uint64_t blk_cap, uint64_t *stats /*[4]*/)
{
  const bool MULTI = L >= (uint64_t)kBlockSize + 128;
  std::vector<uint32_t> table(kTableSize, MULTI ? 0u : 0xffffu);
  std::vector<uint32_t> sched(512);
  {
    uint32_t d = 0;
    for (int k = 0; k < 512; k++) {
      sched[k] = d;
      d = d + 1 + (d >> 5);
      if (d > (1u << 20)) d = 1u << 20;
    }
  }
  const uint32_t nblk = (uint32_t)((L + kBlockSize - 1) / kBlockSize);
  uint64_t total = 0;
  for (uint32_t b = 0; b < nblk; b++) {
    if (b >= blk_cap) return -1;
    blk_ntok[b] = 0;
    const uint64_t boff = (uint64_t)b * kBlockSize;
    const int n = (int)((L - boff) < (uint64_t)kBlockSize ? (L - boff) : (uint64_t)kBlockSize);
    if (n < 128 || L < 128) continue;
    const uint8_t *srcb = src + boff;
    uint32_t *tok = tokens_out + total;
    const uint32_t S0 = (uint32_t)boff;
    const int s_limit = n - kInputMargin;
    int s = 0, next_emit = 0;
    uint32_t ntok = 0;
    bool modeM = false;
    int loop_p0 = 0, k0 = 0;
    auto emit_lits = [&](int a, int e) {
      for (int i = a; i < e; i++) tok[ntok++] = srcb[i];
    };
    auto enc = [&](int pos) -> uint32_t { return MULTI ? (S0 + (uint32_t)pos + 1u) : (uint32_t)(uint16_t)pos; };
    if (total + (uint64_t)n > tok_cap) return -1;
    bool done = false;
    while (!done) {
      // ------------------------------------------------ multi-match fast batch
      if (modeM && s + 31 <= s_limit) {
        int pos[32], cand[32], extl[32], avail[32];
        uint32_t cv[32], h[32], old[32];
        bool hit[32];
        const int base = s - 1;
        for (int l = 0; l < 32; l++) {
          pos[l] = base + l;
          cv[l] = ld32(srcb + pos[l]);
          h[l] = hash4(cv[l]);
          old[l] = table[h[l]];
        }
        // width W of the conflict-free lane prefix.  The kernel finds it from a speculative insert +
        // read-back: in every group of lanes sharing a bucket exactly one (arbitrary) lane wins, and
        // W = lowest losing lane.  The model tries both extremes of that arbitrary choice (lowest or
        // highest lane of a group wins), selected by bit 0 of stats[3] when stats is given.
        int W = 32;
        {
          const bool lowest_wins = stats ? (stats[3] & 1) != 0 : true;
          for (int l = 0; l < 32; l++) {
            bool loser = false;
            for (int q = 0; q < 32; q++)
              if (q != l && h[q] == h[l]) {
                if (lowest_wins ? q < l : q > l) loser = true;
              }
            if (loser) { W = l; break; }
          }
        }
        if (W >= 2) {
          if (stats) stats[0]++;
          for (int l = 0; l < 32; l++) {
            bool ok;
            if (MULTI) {
              const uint32_t D = enc(pos[l]) - old[l];
              ok = old[l] != 0 && D <= (uint32_t)kMaxMatchOffset;
              cand[l] = pos[l] - (int)D;
            } else {
              cand[l] = (int)old[l];
              ok = (uint32_t)(pos[l] - cand[l] - 1) < (uint32_t)kMaxMatchOffset;
            }
            hit[l] = l != 0 && l < W && ok && ld32(srcb + cand[l]) == cv[l];
            avail[l] = l <= 23 ? 8 : (l <= 27 ? 4 : 0);
            extl[l] = 0;
            if (hit[l]) {
              const int t = cand[l] + 4;
              if (t < 0) { avail[l] = 99; extl[l] = 0; } // previous block: match_len == 0 (D1); final
              else {
                while (extl[l] < avail[l] && srcb[pos[l] + 4 + extl[l]] == srcb[t + extl[l]]) extl[l]++;
              }
            }
          }
          uint32_t keep = 0;
          int cur = 1;
          for (;;) {
            int m = -1;
            for (int l = cur; l < W; l++)
              if (hit[l]) { m = l; break; }
            if (m < 0) { // no more hits below W: lanes cur-1..W-1 inserted, lanes cur..W-1 are literals
              for (int l = cur - 1; l < W; l++) keep |= 1u << l;
              for (int l = cur; l < W; l++) tok[ntok++] = cv[l] & 0xff;
              next_emit = base + W;
              modeM = false;
              loop_p0 = base + cur + 1;
              k0 = W - 1 - cur;
              break;
            }
            if (stats) stats[1]++;
            for (int l = cur - 1; l <= m; l++) keep |= 1u << l;
            for (int l = cur; l < m; l++) tok[ntok++] = cv[l] & 0xff;
            const int s2 = pos[m] + 4, t = cand[m] + 4;
            int ext = extl[m];
            if (t >= 0 && ext == avail[m]) { // continue the comparison cooperatively
              int s1 = s2 + kMaxMatchLength - 4;
              if (s1 > n) s1 = n;
              const int a = s1 - s2;
              while (ext < a && srcb[s2 + ext] == srcb[t + ext]) ext++;
            }
            tok[ntok++] = kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1);
            s = s2 + ext;
            next_emit = s;
            if (s >= s_limit) { done = true; break; }
            const int ncur = s - base; // lane of the new probe(s)
            if (ncur >= W) break;      // next batch starts at s (still mode M)
            cur = ncur;
          }
          for (int l = 0; l < 32; l++)
            if ((keep >> l) & 1u) table[h[l]] = enc(pos[l]); // only kept lanes insert (they share no bucket)
          continue;
        }
        if (stats) stats[2]++;
      }
      // ------------------------------------------------ generic single-event batch (v1)
      int pos[32], cand[32];
      bool probe[32], fail[32], active[32], hit[32];
      uint32_t cv[32], h[32], old[32];
      for (int lane = 0; lane < 32; lane++) {
        int step;
        bool loopprobe;
        if (modeM) {
          if (lane == 0) { pos[lane] = s - 1; step = 0; probe[lane] = false; loopprobe = false; }
          else if (lane == 1) { pos[lane] = s; step = 0; probe[lane] = true; loopprobe = false; }
          else { pos[lane] = s + 1 + (lane - 2); step = 1; probe[lane] = true; loopprobe = true; }
        } else {
          const int k = k0 + lane;
          const uint32_t d = k < 32 ? (uint32_t)k : (k < 512 ? sched[k] : (1u << 20));
          pos[lane] = loop_p0 + (int)d;
          step = 1 + (int)(d >> 5);
          probe[lane] = true; loopprobe = true;
        }
        fail[lane] = loopprobe && (pos[lane] + step > s_limit);
        active[lane] = !fail[lane];
        cv[lane] = 0; h[lane] = 0x10000u | (uint32_t)lane; old[lane] = 0;
        if (active[lane]) {
          cv[lane] = ld32(srcb + pos[lane]);
          h[lane] = hash4(cv[lane]);
          old[lane] = table[h[lane]];
        }
      }
      unsigned hitm = 0, failm = 0;
      for (int lane = 0; lane < 32; lane++) {
        int lower = -1;
        for (int q = lane - 1; q >= 0; q--)
          if (h[q] == h[lane]) { lower = q; break; }
        bool ok;
        if (lower >= 0) {
          cand[lane] = pos[lower];
          ok = (pos[lane] - cand[lane]) <= kMaxMatchOffset;
        } else if (MULTI) {
          const uint32_t D = (S0 + (uint32_t)pos[lane] + 1u) - old[lane];
          ok = (old[lane] != 0) && (D <= (uint32_t)kMaxMatchOffset);
          cand[lane] = pos[lane] - (int)D;
        } else {
          const int D = pos[lane] - (int)old[lane];
          ok = (D >= 1) && (D <= kMaxMatchOffset);
          cand[lane] = (int)old[lane];
        }
        hit[lane] = active[lane] && probe[lane] && ok && ld32(srcb + cand[lane]) == cv[lane];
        if (hit[lane]) hitm |= 1u << lane;
        if (fail[lane]) failm |= 1u << lane;
      }
      const unsigned evt = hitm | failm;
      const int m = evt ? __builtin_ffs((int)evt) - 1 : 32;
      const bool mhit = evt && ((hitm >> m) & 1u);
      for (int lane = 0; lane < 32; lane++) {
        const bool in = (m == 32) || (mhit ? lane <= m : lane < m);
        if (active[lane] && in) table[h[lane]] = enc(pos[lane]);
      }
      if (m == 32) {
        if (modeM) { modeM = false; loop_p0 = s + 1; k0 = 30; }
        else k0 += 32;
        continue;
      }
      if (!mhit) break;
      const int s_hit = pos[m], c = cand[m];
      emit_lits(next_emit, s_hit);
      const int s2 = s_hit + 4, t = c + 4;
      int ext = 0;
      if (t >= 0) {
        int s1 = s2 + kMaxMatchLength - 4;
        if (s1 > n) s1 = n;
        const int a = s1 - s2;
        while (ext < a && srcb[s2 + ext] == srcb[t + ext]) ext++;
      }
      tok[ntok++] = kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1);
      s = s2 + ext;
      next_emit = s;
      if (s >= s_limit) break;
      modeM = true;
    }
    emit_lits(next_emit, n);
    blk_ntok[b] = ntok;
    total += ntok;
  }
  return (int64_t)total;
}

// ---------------------------------------------------------------------------
// v3: emulation of the unified fast batch (an experiment of round 2, measured and not adopted: experiments/README.md;
// kept because the model pins the argument the variant rests on).  Besides post-match batches (lane 0 =
// insert(s-1), lane 1 = probe(s), lanes 2.. = a fresh probe loop), a probe loop that is still in its
// consecutive regime (probe index k <= 32: positions loop_p0 + k, deflate-fast.mbt:178-187) continues as a fast
// batch as well: lane l = probe k0 + l at base + l, valid while k0 + l <= 32; past that the reference probes
// every second position, so in the first segment of such a batch hits are only looked for below
// nc = 33 - k0.  After the first match of a batch a fresh loop starts and the limit is gone (33 positions reach
// beyond lane 31).  Everything else (conflict-free prefix W, walk, keep / emit masks) is v2's.
// stats: [0] fast batches, [1] matches taken in fast batches, [2] fast attempts with W < 2, [3] bit 0 = lowest lane
// of a bucket group wins the speculative insert, [4] generic batches, [5] bytes advanced by fast batches,
// [6] bytes advanced by generic batches, [7] fast batches entered from a continuing loop
extern "C" int64_t fbm_parse_stream_v3(const uint8_t *src, uint64_t L, uint32_t *tokens_out, uint64_t tok_cap,
                                       uint32_t *blk_ntok, uint64_t blk_cap, uint64_t *stats /*[8]*/)
{
  const bool MULTI = L >= (uint64_t)kBlockSize + 128;
  std::vector<uint32_t> table(kTableSize, MULTI ? 0u : 0xffffu);
  std::vector<uint32_t> sched(512);
  {
    uint32_t d = 0;
    for (int k = 0; k < 512; k++) {
      sched[k] = d;
      d = d + 1 + (d >> 5);
      if (d > (1u << 20)) d = 1u << 20;
    }
  }
  const uint32_t nblk = (uint32_t)((L + kBlockSize - 1) / kBlockSize);
  uint64_t total = 0;
  for (uint32_t b = 0; b < nblk; b++) {
    if (b >= blk_cap) return -1;
    blk_ntok[b] = 0;
    const uint64_t boff = (uint64_t)b * kBlockSize;
    const int n = (int)((L - boff) < (uint64_t)kBlockSize ? (L - boff) : (uint64_t)kBlockSize);
    if (n < 128 || L < 128) continue;
    const uint8_t *srcb = src + boff;
    uint32_t *tok = tokens_out + total;
    const uint32_t S0 = (uint32_t)boff;
    const int s_limit = n - kInputMargin;
    int s = 0, next_emit = 0;
    uint32_t ntok = 0;
    bool modeM = false;
    int loop_p0 = 0, k0 = 0;
    auto emit_lits = [&](int a, int e) {
      for (int i = a; i < e; i++) tok[ntok++] = srcb[i];
    };
    auto enc = [&](int pos) -> uint32_t { return MULTI ? (S0 + (uint32_t)pos + 1u) : (uint32_t)(uint16_t)pos; };
    if (total + (uint64_t)n > tok_cap) return -1;
    bool done = false;
    while (!done) {
      // ------------------------------------------------ fast batch: 32 consecutive positions base .. base + 31
      const bool contin = !modeM && k0 <= 31;
      const int base = modeM ? s - 1 : loop_p0 + k0;
      if ((modeM || contin) && base + 33 <= s_limit) {
        const int f0 = modeM ? 1 : 0;            // first probing lane (lane 0 of a post-match batch only inserts)
        const int nc = modeM ? 64 : 33 - k0;     // lanes of the first segment that are probes of the running loop
        int pos[32], cand[32], extl[32], avail[32];
        uint32_t cv[32], h[32], old[32];
        bool hit[32];
        for (int l = 0; l < 32; l++) {
          pos[l] = base + l;
          cv[l] = ld32(srcb + pos[l]);
          h[l] = hash4(cv[l]);
          old[l] = table[h[l]];
        }
        int W = 32;
        {
          const bool lowest_wins = stats ? (stats[3] & 1) != 0 : true;
          for (int l = 0; l < 32; l++) {
            bool loser = false;
            for (int q = 0; q < 32; q++)
              if (q != l && h[q] == h[l]) {
                if (lowest_wins ? q < l : q > l) loser = true;
              }
            if (loser) { W = l; break; }
          }
        }
        if (W >= 2) {
          if (stats) { stats[0]++; if (contin) stats[7]++; }
          const int start_pos = modeM ? s : base;
          // a generic batch without an event advances the loop but emits nothing (:198-199): catch up
          if (contin && next_emit < base) { emit_lits(next_emit, base); next_emit = base; }
          for (int l = 0; l < 32; l++) {
            bool ok;
            if (MULTI) {
              const uint32_t D = enc(pos[l]) - old[l];
              ok = old[l] != 0 && D <= (uint32_t)kMaxMatchOffset;
              cand[l] = pos[l] - (int)D;
            } else {
              cand[l] = (int)old[l];
              ok = (uint32_t)(pos[l] - cand[l] - 1) < (uint32_t)kMaxMatchOffset;
            }
            hit[l] = l >= f0 && l < W && ok && ld32(srcb + cand[l]) == cv[l];
            avail[l] = l <= 23 ? 8 : (l <= 27 ? 4 : 0);
            extl[l] = 0;
            if (hit[l]) {
              const int t = cand[l] + 4;
              if (t < 0) { avail[l] = 99; extl[l] = 0; }
              else {
                while (extl[l] < avail[l] && srcb[pos[l] + 4 + extl[l]] == srcb[t + extl[l]]) extl[l]++;
              }
            }
          }
          uint32_t keep = 0;
          int cur = f0;
          bool first = true;
          for (;;) {
            const int lim = (first && nc < W) ? nc : W;
            int m = -1;
            for (int l = cur; l < lim; l++)
              if (hit[l]) { m = l; break; }
            if (m < 0) { // no hit: lanes cur .. lim-1 are inserted literals, the loop goes on at lane lim
              for (int l = (cur > 0 ? cur - 1 : 0); l < lim; l++) keep |= 1u << l;
              for (int l = cur; l < lim; l++) tok[ntok++] = cv[l] & 0xff;
              next_emit = base + lim;
              if (first && !modeM) { k0 += lim; }             // same loop, lim more probes done
              else { loop_p0 = base + cur + 1; k0 = lim - 1 - cur; }
              modeM = false;
              break;
            }
            if (stats) stats[1]++;
            for (int l = (cur > 0 ? cur - 1 : 0); l <= m; l++) keep |= 1u << l;
            for (int l = cur; l < m; l++) tok[ntok++] = cv[l] & 0xff;
            const int s2 = pos[m] + 4, t = cand[m] + 4;
            int ext = extl[m];
            if (t >= 0 && ext == avail[m]) {
              int s1 = s2 + kMaxMatchLength - 4;
              if (s1 > n) s1 = n;
              const int a = s1 - s2;
              while (ext < a && srcb[s2 + ext] == srcb[t + ext]) ext++;
            }
            tok[ntok++] = kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1);
            s = s2 + ext;
            next_emit = s;
            modeM = true;
            first = false;
            if (s >= s_limit) { done = true; break; }
            const int ncur = s - base;
            if (ncur >= W) break;
            cur = ncur;
          }
          for (int l = 0; l < 32; l++)
            if ((keep >> l) & 1u) table[h[l]] = enc(pos[l]);
          if (stats) stats[5] += (uint64_t)((modeM ? s : loop_p0 + (k0 <= 32 ? k0 : 0)) - start_pos);
          continue;
        }
        if (stats) stats[2]++;
      }
      // ------------------------------------------------ generic single-event batch (v1)
      if (stats) stats[4]++;
      const int gstart = modeM ? s : loop_p0;
      int pos[32], cand[32];
      bool probe[32], fail[32], active[32], hit[32];
      uint32_t cv[32], h[32], old[32];
      for (int lane = 0; lane < 32; lane++) {
        int step;
        bool loopprobe;
        if (modeM) {
          if (lane == 0) { pos[lane] = s - 1; step = 0; probe[lane] = false; loopprobe = false; }
          else if (lane == 1) { pos[lane] = s; step = 0; probe[lane] = true; loopprobe = false; }
          else { pos[lane] = s + 1 + (lane - 2); step = 1; probe[lane] = true; loopprobe = true; }
        } else {
          const int k = k0 + lane;
          const uint32_t d = k < 32 ? (uint32_t)k : (k < 512 ? sched[k] : (1u << 20));
          pos[lane] = loop_p0 + (int)d;
          step = 1 + (int)(d >> 5);
          probe[lane] = true; loopprobe = true;
        }
        fail[lane] = loopprobe && (pos[lane] + step > s_limit);
        active[lane] = !fail[lane];
        cv[lane] = 0; h[lane] = 0x10000u | (uint32_t)lane; old[lane] = 0;
        if (active[lane]) {
          cv[lane] = ld32(srcb + pos[lane]);
          h[lane] = hash4(cv[lane]);
          old[lane] = table[h[lane]];
        }
      }
      unsigned hitm = 0, failm = 0;
      for (int lane = 0; lane < 32; lane++) {
        int lower = -1;
        for (int q = lane - 1; q >= 0; q--)
          if (h[q] == h[lane]) { lower = q; break; }
        bool ok;
        if (lower >= 0) {
          cand[lane] = pos[lower];
          ok = (pos[lane] - cand[lane]) <= kMaxMatchOffset;
        } else if (MULTI) {
          const uint32_t D = (S0 + (uint32_t)pos[lane] + 1u) - old[lane];
          ok = (old[lane] != 0) && (D <= (uint32_t)kMaxMatchOffset);
          cand[lane] = pos[lane] - (int)D;
        } else {
          const int D = pos[lane] - (int)old[lane];
          ok = (D >= 1) && (D <= kMaxMatchOffset);
          cand[lane] = (int)old[lane];
        }
        hit[lane] = active[lane] && probe[lane] && ok && ld32(srcb + cand[lane]) == cv[lane];
        if (hit[lane]) hitm |= 1u << lane;
        if (fail[lane]) failm |= 1u << lane;
      }
      const unsigned evt = hitm | failm;
      const int m = evt ? __builtin_ffs((int)evt) - 1 : 32;
      const bool mhit = evt && ((hitm >> m) & 1u);
      for (int lane = 0; lane < 32; lane++) {
        const bool in = (m == 32) || (mhit ? lane <= m : lane < m);
        if (active[lane] && in) table[h[lane]] = enc(pos[lane]);
      }
      if (m == 32) {
        if (modeM) { modeM = false; loop_p0 = s + 1; k0 = 30; }
        else k0 += 32;
        continue;
      }
      if (!mhit) break;
      const int s_hit = pos[m], c = cand[m];
      emit_lits(next_emit, s_hit);
      const int s2 = s_hit + 4, t = c + 4;
      int ext = 0;
      if (t >= 0) {
        int s1 = s2 + kMaxMatchLength - 4;
        if (s1 > n) s1 = n;
        const int a = s1 - s2;
        while (ext < a && srcb[s2 + ext] == srcb[t + ext]) ext++;
      }
      tok[ntok++] = kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1);
      s = s2 + ext;
      next_emit = s;
      if (stats) stats[6] += (uint64_t)(s - gstart);
      if (s >= s_limit) break;
      modeM = true;
    }
    emit_lits(next_emit, n);
    blk_ntok[b] = ntok;
    total += ntok;
  }
  return (int64_t)total;
}

// ---------------------------------------------------------------------------
// v4: emulation of the fixed-window parse (experiments/parse_windows.cu, round 2: measured and not adopted, see
// experiments/README.md; kept because the model pins the window walk's exactness).  The block is cut into windows of 32
// consecutive positions [32 j, 32 j + 32); lane l of window j always holds position 32 j + l, so the loads of the
// next windows can be issued before the current one has been walked (the kernel pipelines them; that changes
// when values arrive, not which values are used -- stale table entries are detected by a re-read and replaced --
// so the model reads everything in place).
//
// State between windows: `pend` = the insert of position s - 1 after a match is still to be done (:246-251),
// and the running probe loop (loop_p0, k0): probe k sits at loop_p0 + d_k, d_k = k for k <= 32 (:178-187); the
// probe of s itself right after a match (:252-260) is probe -1 of the loop that starts at s + 1.  A window is
// entered at the lane of the next thing to do (cur0); lanes below it lie inside an earlier match: they neither
// probe nor insert, and they are ignored when the conflict-free lane prefix W is determined.
extern "C" int64_t fbm_parse_stream_v4(const uint8_t *src, uint64_t L, uint32_t *tokens_out, uint64_t tok_cap,
                                       uint32_t *blk_ntok, uint64_t blk_cap, uint64_t *stats /*[8]*/)
{
  const bool MULTI = L >= (uint64_t)kBlockSize + 128;
  std::vector<uint32_t> table(kTableSize, MULTI ? 0u : 0xffffu);
  std::vector<uint32_t> sched(512);
  {
    uint32_t d = 0;
    for (int k = 0; k < 512; k++) {
      sched[k] = d;
      d = d + 1 + (d >> 5);
      if (d > (1u << 20)) d = 1u << 20;
    }
  }
  const uint32_t nblk = (uint32_t)((L + kBlockSize - 1) / kBlockSize);
  uint64_t total = 0;
  for (uint32_t b = 0; b < nblk; b++) {
    if (b >= blk_cap) return -1;
    blk_ntok[b] = 0;
    const uint64_t boff = (uint64_t)b * kBlockSize;
    const int n = (int)((L - boff) < (uint64_t)kBlockSize ? (L - boff) : (uint64_t)kBlockSize);
    if (n < 128 || L < 128) continue;
    const uint8_t *srcb = src + boff;
    uint32_t *tok = tokens_out + total;
    const uint32_t S0 = (uint32_t)boff;
    const int s_limit = n - kInputMargin;
    int next_emit = 0;
    uint32_t ntok = 0;
    bool pend = false;
    int loop_p0 = 0, k0 = 0;
    auto emit_lits = [&](int a, int e) {
      for (int i = a; i < e; i++) tok[ntok++] = srcb[i];
    };
    auto enc = [&](int pos) -> uint32_t { return MULTI ? (S0 + (uint32_t)pos + 1u) : (uint32_t)(uint16_t)pos; };
    auto match_len = [&](int s2, int t) -> int { // :286-342; t < 0 -> 0 (D1)
      if (t < 0) return 0;
      int s1 = s2 + kMaxMatchLength - 4;
      if (s1 > n) s1 = n;
      int ext = 0;
      while (s2 + ext < s1 && srcb[s2 + ext] == srcb[t + ext]) ext++;
      return ext;
    };
    if (total + (uint64_t)n > tok_cap) return -1;
    bool done = false;
    while (!done) {
      // ------------------------------------------------ fast window
      if (k0 <= 31) {
        const int P = pend ? loop_p0 - 2 : loop_p0 + k0;
        const int j = P >> 5, base = j << 5, cur0 = P & 31;
        if (base + 33 <= s_limit) {
          int pos[32], cand[32];
          uint32_t cv[32], h[32], old[32];
          bool hit[32];
          for (int l = 0; l < 32; l++) {
            pos[l] = base + l;
            cv[l] = ld32(srcb + pos[l]);
            h[l] = hash4(cv[l]);
            old[l] = table[h[l]];
          }
          int W = 32;
          {
            const bool lowest_wins = stats ? (stats[3] & 1) != 0 : true;
            for (int l = cur0; l < 32; l++) {
              bool loser = false;
              for (int q = cur0; q < 32; q++)
                if (q != l && h[q] == h[l]) {
                  if (lowest_wins ? q < l : q > l) loser = true;
                }
              if (loser) { W = l; break; }
            }
          }
          if (W >= (cur0 + 2 < 32 ? cur0 + 2 : 32)) {
            if (stats) stats[0]++;
            for (int l = 0; l < 32; l++) {
              bool ok;
              if (MULTI) {
                const uint32_t D = enc(pos[l]) - old[l];
                ok = old[l] != 0 && D <= (uint32_t)kMaxMatchOffset;
                cand[l] = pos[l] - (int)D;
              } else {
                cand[l] = (int)old[l];
                ok = (uint32_t)(pos[l] - cand[l] - 1) < (uint32_t)kMaxMatchOffset;
              }
              hit[l] = l < W && ok && ld32(srcb + cand[l]) == cv[l];
            }
            if (!pend && next_emit < P) { emit_lits(next_emit, P); next_emit = P; }
            uint32_t keep = 0;
            int cur = cur0;
            if (pend) { keep |= 1u << cur0; cur = cur0 + 1; pend = false; } // insert(s - 1); probe -1 follows
            int lim = cur + 33 - k0 < W ? cur + 33 - k0 : W;               // probes of the running loop only
            for (;;) {
              int m = -1;
              for (int l = cur; l < lim; l++)
                if (hit[l]) { m = l; break; }
              if (m < 0) {
                for (int l = cur; l < lim; l++) { keep |= 1u << l; tok[ntok++] = cv[l] & 0xff; }
                if (lim > cur) next_emit = base + lim;
                k0 += lim - cur;
                break;
              }
              if (stats) stats[1]++;
              for (int l = cur; l <= m; l++) keep |= 1u << l;
              for (int l = cur; l < m; l++) tok[ntok++] = cv[l] & 0xff;
              const int s2 = pos[m] + 4, t = cand[m] + 4;
              const int ext = match_len(s2, t);
              tok[ntok++] = kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1);
              const int s = s2 + ext;
              next_emit = s;
              loop_p0 = s + 1;
              k0 = -1;
              if (s >= s_limit) { done = true; break; }
              const int ncur = s - base;
              if (ncur - 1 < W) keep |= 1u << (ncur - 1); // insert(s - 1) lies in this window's prefix
              else pend = true;
              if (ncur >= W) break;
              cur = ncur;
              lim = W;
            }
            for (int l = 0; l < 32; l++)
              if ((keep >> l) & 1u) table[h[l]] = enc(pos[l]);
            continue;
          }
          if (stats) stats[2]++;
        }
      }
      // ------------------------------------------------ generic single-event batch
      if (stats) stats[4]++;
      int pos[32], cand[32];
      bool probe[32], fail[32], active[32], hit[32];
      uint32_t cv[32], h[32], old[32];
      for (int lane = 0; lane < 32; lane++) {
        int step;
        bool loopprobe;
        if (pend && lane == 0) { pos[lane] = loop_p0 - 2; step = 0; probe[lane] = false; loopprobe = false; }
        else {
          const int k = pend ? lane - 2 : k0 + lane;
          const int d = k < 32 ? k : (k < 512 ? (int)sched[k] : (1 << 20));
          pos[lane] = loop_p0 + d;
          step = k < 0 ? 0 : 1 + (d >> 5);
          probe[lane] = true; loopprobe = k >= 0;
        }
        fail[lane] = loopprobe && (pos[lane] + step > s_limit);
        active[lane] = !fail[lane];
        cv[lane] = 0; h[lane] = 0x10000u | (uint32_t)lane; old[lane] = 0;
        if (active[lane]) {
          cv[lane] = ld32(srcb + pos[lane]);
          h[lane] = hash4(cv[lane]);
          old[lane] = table[h[lane]];
        }
      }
      unsigned hitm = 0, failm = 0;
      for (int lane = 0; lane < 32; lane++) {
        int lower = -1;
        for (int q = lane - 1; q >= 0; q--)
          if (h[q] == h[lane]) { lower = q; break; }
        bool ok;
        if (lower >= 0) {
          cand[lane] = pos[lower];
          ok = (pos[lane] - cand[lane]) <= kMaxMatchOffset;
        } else if (MULTI) {
          const uint32_t D = (S0 + (uint32_t)pos[lane] + 1u) - old[lane];
          ok = (old[lane] != 0) && (D <= (uint32_t)kMaxMatchOffset);
          cand[lane] = pos[lane] - (int)D;
        } else {
          const int D = pos[lane] - (int)old[lane];
          ok = (D >= 1) && (D <= kMaxMatchOffset);
          cand[lane] = (int)old[lane];
        }
        hit[lane] = active[lane] && probe[lane] && ok && ld32(srcb + cand[lane]) == cv[lane];
        if (hit[lane]) hitm |= 1u << lane;
        if (fail[lane]) failm |= 1u << lane;
      }
      const unsigned evt = hitm | failm;
      const int m = evt ? __builtin_ffs((int)evt) - 1 : 32;
      const bool mhit = evt && ((hitm >> m) & 1u);
      for (int lane = 0; lane < 32; lane++) {
        const bool in = (m == 32) || (mhit ? lane <= m : lane < m);
        if (active[lane] && in) table[h[lane]] = enc(pos[lane]);
      }
      if (m == 32) {
        if (pend) { pend = false; k0 = 30; } // lanes 1..31 were probes -1..29
        else k0 += 32;
        continue;
      }
      if (!mhit) break;
      const int s_hit = pos[m], c = cand[m];
      emit_lits(next_emit, s_hit);
      const int s2 = s_hit + 4, t = c + 4;
      const int ext = match_len(s2, t);
      tok[ntok++] = kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1);
      const int s = s2 + ext;
      next_emit = s;
      if (s >= s_limit) break;
      pend = true; loop_p0 = s + 1; k0 = -1;
    }
    emit_lits(next_emit, n);
    blk_ntok[b] = ntok;
    total += ntok;
  }
  return (int64_t)total;
}
