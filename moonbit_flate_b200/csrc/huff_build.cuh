// huff_build.cuh -- K3 building blocks: length-limited canonical code
// construction, codegen RLE, dynamic-block header.  One WARP per block.
//
//   HuffmanEncoder::generate / bit_counts / assign_encoding_and_size
//       huffman-code.mbt:112-343
//   HuffmanBitWriter::generate_codegen / dynamic_size / write_dynamic_header
//       huffman-bit-writer.mbt:241-360, :421-471
//
// The reference's bit_counts is Go's lazy "boundary package-merge": a serial walk
// of ~2n*L steps.  Here the same package-merge lists are built eagerly, level by
// level, so that a warp can work on them in parallel:
//     list_1 = the leaves (ascending (freq, literal)),
//     list_k = merge(leaves, pairs(list_{k-1})), ties: PAIR FIRST
// (the reference takes a leaf only if next_char < next_pair, :186-190), each
// truncated to its first 2n-2 items (the most any level is ever asked for; the
// truncation loses no pair because a list never has more than 2n-1 items).
// Walking down from the top level with m_top = 2n-2, m_{k-1} = 2 * (pairs among
// the first m_k items of list_k) gives the number of leaves each level uses,
// which is exactly the reference's leaf_counts[max_bits][k]; symbol i (ascending
// order) gets length #{k : i < leaves_k} (:236-242, :260-278).
//
// The code is written once for device and host: FB_PFOR distributes a loop over
// the 32 lanes on the device and is a plain loop on the host, where the CPU
// test-suite diffs every output against the oracle (tests/test_hostmodel.py).
#pragma once
#include "common.cuh"

#if defined(__CUDA_ARCH__)
#define FB_LANE ((int)(threadIdx.x & 31))
#define FB_NLANES 32
#define FB_WSYNC() __syncwarp()
#else
#define FB_LANE 0
#define FB_NLANES 1
#define FB_WSYNC() ((void)0)
#endif
#define FB_PFOR(i, n) for (int i = FB_LANE; i < (int)(n); i += FB_NLANES)

#if defined(__CUDACC__)
#define FB_HD __host__ __device__
#else
#define FB_HD
#endif

namespace fb {

FB_HD inline unsigned brev32(unsigned v)
{
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  return (v >> 16) | (v << 16);
#endif
}

FB_HD inline int wsum(int v)
{
#if defined(__CUDA_ARCH__)
  return __reduce_add_sync(0xffffffffu, v);
#else
  return v;
#endif
}
FB_HD inline int wmin(int v)
{
#if defined(__CUDA_ARCH__)
  return __reduce_min_sync(0xffffffffu, v);
#else
  return v;
#endif
}
FB_HD inline int wmax(int v)
{
#if defined(__CUDA_ARCH__)
  return __reduce_max_sync(0xffffffffu, v);
#else
  return v;
#endif
}

constexpr int kMaxSyms = 288;

// Per-warp working set (shared memory on the device), for alphabets with at most NS symbols IN USE.  The warp's
// scratch is what caps the kernel's occupancy: blocks whose literal/length alphabet has <= 128 used symbols
// (small blocks, text) are built with the 6.9 KB variant, 32 warps per SM instead of 18.
template <int NS>
struct HuffScratchT {
  static constexpr int kSyms = NS;
  static constexpr int kPad = NS <= 32 ? 32 : NS <= 64 ? 64 : NS <= 128 ? 128 : NS <= 256 ? 256 : 512; // bitonic sort width
  static constexpr int kMaskWords = 2 * NS / 32;                                                      // one bit per list position
  uint32_t keys[kPad];              // (freq << 9) | literal, ascending == by_frequency (huffman-code.mbt:346-351)
  uint32_t lists[4 * NS];           // package-merge lists (weights), double buffered; later: run table, header staging
  uint32_t pairs[NS];
  uint32_t leafmask[16][kMaskWords]; // bit p set: item p of list_k is a leaf
  uint32_t run[16];                 // canonical code counters
  // block assembly
  uint32_t freq[320];               // lit/len (286) ++ offset (30) histogram
  uint32_t cgfreq[32];
  uint8_t len[320];
  uint16_t code[320];
  uint8_t cglen[32];
  uint16_t cgcode[32];
  uint8_t cg[kNumLit + kNumDist + 4];
  uint16_t items[kNumLit + kNumDist + 4]; // code-length symbols of the header: symbol | extra << 8
  int misc[8];
};
using HuffScratch = HuffScratchT<kMaxSyms>;
static_assert(4 * 128 * 4 >= 2 * 320 * 2 && 4 * 128 * 4 >= kHdrWords * 4, "lists[] also hosts the run table and the header staging");

#if defined(__CUDA_ARCH__)
// One package-merge level on the device: each lane owns up to Q leaves and Q pairs; their (branch-free)
// binary searches advance in lockstep so the shared-memory loads are independent.
template <int Q, int TOP, class SC>
__device__ __forceinline__ void pm_merge_level(SC &S, int k, int n, int np, int cap, uint32_t *cur)
{
  int lo[Q];
  uint32_t w[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = FB_LANE + 32 * q;
    w[q] = i < n ? (S.keys[i] >> 9) : 0u;
    lo[q] = 0;
  }
#pragma unroll
  for (int step = TOP; step > 0; step >>= 1) { // lo = #pairs with weight <= w
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int c = lo[q] + step;
      if (c <= np && S.pairs[c - 1] <= w[q]) lo[q] = c;
    }
  }
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = FB_LANE + 32 * q;
    if (i < n) {
      const int pos = i + lo[q];
      atomicOr(&S.leafmask[k][pos >> 5], 1u << (pos & 31));
      if (pos < cap) cur[pos] = w[q];
    }
  }
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int jx = FB_LANE + 32 * q;
    w[q] = jx < np ? S.pairs[jx] : 0u;
    lo[q] = 0;
  }
#pragma unroll
  for (int step = TOP; step > 0; step >>= 1) { // lo = #leaves with weight < w
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const int c = lo[q] + step;
      if (c <= n && (S.keys[c - 1] >> 9) < w[q]) lo[q] = c;
    }
  }
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int jx = FB_LANE + 32 * q;
    if (jx < np) {
      const int pos = jx + lo[q];
      if (pos < cap) cur[pos] = w[q];
    }
  }
}
#endif

// HuffmanEncoder::generate (:295-343): freq[0..nsym) -> len[] (0 for unused symbols), code[] bit-reversed.
template <class SC>
FB_HD inline void warp_generate(const uint32_t *freq, int nsym, int max_bits, uint8_t *len, uint16_t *code, SC &S)
{
  constexpr int kMaskWords = SC::kMaskWords;
  // gather the used symbols (any order: the sort below orders them), pad to a power of two
  int first = nsym, n = 0;
#if defined(__CUDA_ARCH__)
  {
    const int lane = FB_LANE;
    const unsigned ltm = (1u << lane) - 1u;
    for (int base = 0; base < nsym; base += 32) {
      const int i = base + lane;
      const uint32_t f = i < nsym ? freq[i] : 0u;
      if (i < nsym) { len[i] = 0; code[i] = 0; }
      const unsigned m = __ballot_sync(0xffffffffu, f != 0);
      if (f) {
        S.keys[n + __popc(m & ltm)] = (f << 9) | (uint32_t)i;
        if (i < first) first = i;
      }
      n += __popc(m);
    }
    first = wmin(first);
  }
#else
  for (int i = 0; i < nsym; i++) {
    len[i] = 0; code[i] = 0;
    if (freq[i]) { if (i < first) first = i; S.keys[n++] = (freq[i] << 9) | (uint32_t)i; }
  }
#endif
  int P = 32;
  while (P < n) P <<= 1;
  FB_PFOR(i, P - n) S.keys[n + i] = 0xffffffffu;
  FB_WSYNC();
  if (n <= 2) { // :326-336: codes 0, 1 in literal order, length 1
    FB_PFOR(i, nsym) if (freq[i]) { len[i] = 1; code[i] = (uint16_t)(i != first); }
    FB_WSYNC();
    return;
  }
  // sort ascending (any correct sort: keys are distinct) -- bitonic network over P >= n entries
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#if defined(__CUDA_ARCH__)
      if (SC::kPad >= 256 && P >= 256) { // several independent compare-exchanges per lane and stage
        uint32_t a[8], b[8];
        int ia[8];
        const int nq = P >> 6;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const int t = FB_LANE + 32 * q;
          ia[q] = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          if (q < nq) { a[q] = S.keys[ia[q]]; b[q] = S.keys[ia[q] | j]; }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const bool up = (ia[q] & k) == 0;
          if (q < nq && (a[q] > b[q]) == up) { S.keys[ia[q]] = b[q]; S.keys[ia[q] | j] = a[q]; }
        }
      } else
#endif
      FB_PFOR(t, P / 2) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const uint32_t a = S.keys[i], b = S.keys[p];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { S.keys[i] = b; S.keys[p] = a; }
      }
      FB_WSYNC();
    }
  }
  // package-merge lists (bit_counts :112-244)
  const int mb = max_bits < n - 1 ? max_bits : n - 1; // :126-129
  const int cap = 2 * n - 2;
  uint32_t *prev = S.lists, *cur = S.lists + 2 * SC::kSyms;
  FB_PFOR(i, 16 * kMaskWords) (&S.leafmask[0][0])[i] = 0;
  FB_PFOR(i, n) prev[i] = S.keys[i] >> 9;
  int plen = n;
  FB_WSYNC();
  for (int k = 2; k <= mb; k++) {
    const int np = plen >> 1;
    FB_PFOR(j, np) S.pairs[j] = prev[2 * j] + prev[2 * j + 1];
    FB_WSYNC();
#if defined(__CUDA_ARCH__)
    if (n <= 32) pm_merge_level<1, 32>(S, k, n, np, cap, cur);
    else if (n <= 64) pm_merge_level<2, 64>(S, k, n, np, cap, cur);
    else if (SC::kSyms <= 128 || n <= 128) pm_merge_level<4, 128>(S, k, n, np, cap, cur);
    else pm_merge_level<9, 256>(S, k, n, np, cap, cur);
#else
    FB_PFOR(i, n) { // leaf i sits after every pair of weight <= its own
      const uint32_t w = S.keys[i] >> 9;
      int lo = 0, hi = np;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (S.pairs[mid] <= w) lo = mid + 1; else hi = mid;
      }
      const int pos = i + lo;
      S.leafmask[k][pos >> 5] |= 1u << (pos & 31);
      if (pos < cap) cur[pos] = w;
    }
    FB_PFOR(j, np) { // pair j sits after every leaf of weight < its own
      const uint32_t w = S.pairs[j];
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((S.keys[mid] >> 9) < w) lo = mid + 1; else hi = mid;
      }
      const int pos = j + lo;
      if (pos < cap) cur[pos] = w;
    }
#endif
    plen = n + np < cap ? n + np : cap;
    FB_WSYNC();
    uint32_t *t = prev; prev = cur; cur = t;
  }
  // leaves used per level, top down (every lane computes the same values)
  int nl[17];
#pragma unroll
  for (int k = 0; k < 17; k++) nl[k] = 0;
  {
    int m = cap;
    for (int k = mb; k >= 1; k--) {
      int lv;
      if (k == 1) lv = m < n ? m : n;
      else { // leaves among the first m items of list_k
#if defined(__CUDA_ARCH__)
        const int wi = FB_LANE;
        int c = 0;
        if (wi < kMaskWords) {
          const int lo_bit = 32 * wi;
          const unsigned sel = m >= lo_bit + 32 ? 0xffffffffu : (m > lo_bit ? ((1u << (m - lo_bit)) - 1u) : 0u);
          c = __popc(S.leafmask[k][wi] & sel);
        }
        lv = __reduce_add_sync(0xffffffffu, c);
#else
        lv = 0;
        for (int wi = 0; wi < kMaskWords; wi++) {
          const int lo_bit = 32 * wi;
          const unsigned sel = m >= lo_bit + 32 ? 0xffffffffu : (m > lo_bit ? ((1u << (m - lo_bit)) - 1u) : 0u);
          lv += __builtin_popcount(S.leafmask[k][wi] & sel);
        }
#endif
      }
      nl[k] = lv;
      m = 2 * (m - lv);
    }
  }
  // code length of sorted position i = #{k : i < leaves_k}; counts per length = bit_count[] (:236-242)
  FB_PFOR(i, n) {
    int l = 0;
    for (int k = 1; k <= mb; k++) l += (i < nl[k]);
    len[S.keys[i] & 511u] = (uint8_t)l;
  }
  // canonical codes: per length in literal order, consecutive, stored bit-reversed (:250-280)
  unsigned next_code[17];
  {
    unsigned c = 0;
    next_code[0] = 0;
    for (int b = 1; b <= 16; b++) {
      c <<= 1;
      next_code[b] = c;
      // bit_count[b] = leaves_{mb-b+1} - leaves_{mb-b}
      const int hiK = mb - b + 1, loK = mb - b;
      if (b <= mb) c += (unsigned)(nl[hiK] - (loK >= 1 ? nl[loK] : 0));
    }
  }
  FB_PFOR(i, 16) S.run[i] = 0;
  FB_WSYNC();
#if defined(__CUDA_ARCH__)
  {
    const int lane = FB_LANE;
    const unsigned ltm = (1u << lane) - 1u;
    for (int base = 0; base < nsym; base += 32) {
      const int i = base + lane;
      const int l = i < nsym ? (int)len[i] : 0;
      const unsigned peers = __match_any_sync(0xffffffffu, l);
      if (l) {
        const unsigned v = next_code[l] + S.run[l] + (unsigned)__popc(peers & ltm);
        code[i] = (uint16_t)(__brev(v) >> (32 - l));
      }
      __syncwarp();
      if (l && (peers >> lane) <= 1u) S.run[l] += (unsigned)__popc(peers); // highest lane of each group
      __syncwarp();
    }
  }
#else
  for (int i = 0; i < nsym; i++) {
    const int l = len[i];
    if (l) {
      const unsigned v = next_code[l] + S.run[l]++;
      code[i] = (uint16_t)(brev32(v) >> (32 - l));
    }
  }
#endif
  FB_WSYNC();
}

struct BitAcc {
  uint32_t *words;
  uint64_t acc;
  int nacc;
  int nwords;
  FB_HD void put(uint32_t v, int nb)
  {
    acc |= (uint64_t)v << nacc;
    nacc += nb;
    if (nacc >= 32) {
      words[nwords++] = (uint32_t)acc;
      acc >>= 32;
      nacc -= 32;
    }
  }
  FB_HD int finish()
  {
    const int total = nwords * 32 + nacc;
    if (nacc) words[nwords++] = (uint32_t)acc;
    return total;
  }
};

// symbols a run of `count` equal code lengths `size` emits (generate_codegen, hbw:241-330)
FB_HD inline int codegen_run_items(int size, int count)
{
  if (size != 0) {
    const int c = count - 1;
    return 1 + c / 6 + (c % 6 >= 3 ? 1 : c % 6);
  }
  const int r = count % 138;
  return count / 138 + (r >= 3 ? 1 : r);
}

// bits of one codegen item (symbol | extra << 8): the symbol's code, then 2 / 3 / 7 extra bits behind 16 / 17 / 18
FB_HD inline void codegen_item_bits(uint16_t item, const uint16_t *cgcode, const uint8_t *cglen, uint32_t &v, int &nb)
{
  const int sym = item & 0xff, extra = item >> 8;
  v = cgcode[sym];
  nb = cglen[sym];
  const int xb = sym == 16 ? 2 : (sym == 17 ? 3 : (sym == 18 ? 7 : 0));
  v |= (uint32_t)extra << nb;
  nb += xb;
}

FB_HD inline void fb_add(uint32_t *p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}

// Everything write_block_dynamic / write_block_huff decide before the first
// payload bit (hbw:496-534, :738-787): codes, codegen, "store instead" test,
// header bit string, exact block size.  gfreq = lit/len[286] ++ offset[30]
// histogram of the block (EOB already counted); kind = kKindDynamic / kKindHuff.
struct BlockBuild {
  int kind;           // may turn into kKindStored (quirk D2 test)
  uint32_t hdr_nbits; // bits in hdr_words
  uint32_t blk_bits;  // header + payload + EOB
};

template <class SC>
FB_HD inline BlockBuild build_block_warp(const uint32_t *gfreq, int kind, uint32_t n, uint32_t *codeout,
                                         uint32_t *hdr_words, SC &S)
{
  BlockBuild res;
  res.kind = kind;
  res.hdr_nbits = 0;
  res.blk_bits = 0;
  int last_lit = 0, last_off = -1;
  FB_PFOR(i, 320) {
    const uint32_t f = i < kNumLit + kNumDist ? (gfreq ? gfreq[i] : S.freq[i]) : 0u;
    S.freq[i] = f;
    if (f) {
      if (i < kNumLit) { if (i > last_lit) last_lit = i; }
      else if (i - kNumLit > last_off) last_off = i - kNumLit;
    }
  }
  last_lit = wmax(last_lit);
  last_off = wmax(last_off);
  FB_WSYNC();
  int num_literals, num_offsets;
  if (kind == kKindDynamic) { // index_tokens tail (hbw:574-592)
    num_literals = last_lit + 1; // >= 257: EOB is counted
    num_offsets = last_off + 1;
    if (num_offsets == 0) {
      if (FB_LANE == 0) S.freq[kNumLit] = 1;
      num_offsets = 1;
      FB_WSYNC();
    }
    warp_generate(S.freq, kNumLit, 15, S.len, S.code, S);
    warp_generate(S.freq + kNumLit, kNumDist, 15, S.len + kNumLit, S.code + kNumLit, S);
  } else { // write_block_huff (hbw:747-758): literal-only, static huff_offset (huffman-code.mbt:691)
    num_literals = kEob + 1;
    num_offsets = 1;
    warp_generate(S.freq, kNumLit, 15, S.len, S.code, S);
    FB_PFOR(i, kNumDist) { S.len[kNumLit + i] = (uint8_t)(i == 0); S.code[kNumLit + i] = 0; }
    FB_WSYNC();
  }

  // generate_codegen (hbw:241-330).  The reference walks the <= 316 code lengths once and emits, run by run,
  // the code-length symbols 0..15 / 16 (repeat previous 3..6 times) / 17 (3..10 zeros) / 18 (11..138 zeros) with
  // their extra values.  What a run of `count` equal lengths `size` emits depends on (size, count) alone:
  //   size != 0: size, then c = count - 1 copies as  c / 6 x (16, extra 3), one (16, extra c % 6 - 3) if c % 6 >= 3,
  //              and the remaining c % 6 < 3 copies written out;
  //   size == 0: count / 138 x (18, extra 127), one (18, extra r - 11) if r = count % 138 >= 11, else one
  //              (17, extra r - 3) if r >= 3, else r zeros written out.
  // So the runs are found in parallel, every run computes how many symbols it emits, a prefix sum places them,
  // and every run writes its own.  The symbols are kept as items (symbol | extra << 8) in S.items; the header
  // emission below places their bits with a second prefix sum.
  FB_PFOR(i, 32) S.cgfreq[i] = 0;
  FB_PFOR(i, num_literals) S.cg[i] = S.len[i];
  FB_PFOR(i, num_offsets) S.cg[num_literals + i] = S.len[kNumLit + i];
  FB_WSYNC();
  const int ncl = num_literals + num_offsets;
  uint16_t *rstart = reinterpret_cast<uint16_t *>(S.lists);       // [<= 317] first position of every run (+ end)
  uint16_t *roff = reinterpret_cast<uint16_t *>(S.lists) + 320;   // [<= 317] first item of every run
  int nruns = 0;
#if defined(__CUDA_ARCH__)
  {
    const int lane = FB_LANE;
    const unsigned ltm = (1u << lane) - 1u;
    for (int base = 0; base < ncl; base += 32) {
      const int i = base + lane;
      const bool st = i < ncl && (i == 0 || S.cg[i] != S.cg[i - 1]);
      const unsigned m = __ballot_sync(0xffffffffu, st);
      if (st) rstart[nruns + __popc(m & ltm)] = (uint16_t)i;
      nruns += __popc(m);
    }
  }
#else
  for (int i = 0; i < ncl; i++)
    if (i == 0 || S.cg[i] != S.cg[i - 1]) rstart[nruns++] = (uint16_t)i;
#endif
  if (FB_LANE == 0) rstart[nruns] = (uint16_t)ncl;
  FB_WSYNC();
  // symbols every run emits -> exclusive prefix sum
  int nitems = 0;
#if defined(__CUDA_ARCH__)
  for (int base = 0; base < nruns; base += 32) {
    const int r = base + FB_LANE;
    int e = 0;
    if (r < nruns) e = codegen_run_items(S.cg[rstart[r]], (int)rstart[r + 1] - (int)rstart[r]);
    int x = e;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (FB_LANE >= o) x += y;
    }
    if (r < nruns) roff[r] = (uint16_t)(nitems + x - e);
    nitems += __shfl_sync(0xffffffffu, x, 31);
  }
#else
  for (int r = 0; r < nruns; r++) {
    roff[r] = (uint16_t)nitems;
    nitems += codegen_run_items(S.cg[rstart[r]], (int)rstart[r + 1] - (int)rstart[r]);
  }
#endif
  FB_WSYNC();
  FB_PFOR(r, nruns) {
    const int size = S.cg[rstart[r]];
    int count = (int)rstart[r + 1] - (int)rstart[r];
    uint16_t *it = S.items + roff[r];
    int o = 0;
    if (size != 0) {
      it[o++] = (uint16_t)size;
      int c = count - 1;
      const int full = c / 6, r6 = c % 6;
      for (int q = 0; q < full; q++) it[o++] = (uint16_t)(16 | (3 << 8));
      int rem = r6;
      if (r6 >= 3) { it[o++] = (uint16_t)(16 | ((r6 - 3) << 8)); rem = 0; }
      for (int q = 0; q < rem; q++) it[o++] = (uint16_t)size;
      fb_add(&S.cgfreq[size], (uint32_t)(1 + rem));
      if (full + (r6 >= 3)) fb_add(&S.cgfreq[16], (uint32_t)(full + (r6 >= 3)));
    } else {
      const int full = count / 138, r138 = count % 138;
      for (int q = 0; q < full; q++) it[o++] = (uint16_t)(18 | (127 << 8));
      int rem = r138;
      if (r138 >= 11) { it[o++] = (uint16_t)(18 | ((r138 - 11) << 8)); rem = 0; }
      else if (r138 >= 3) { it[o++] = (uint16_t)(17 | ((r138 - 3) << 8)); rem = 0; }
      for (int q = 0; q < rem; q++) it[o++] = 0;
      if (full + (r138 >= 11)) fb_add(&S.cgfreq[18], (uint32_t)(full + (r138 >= 11)));
      if (r138 >= 3 && r138 < 11) fb_add(&S.cgfreq[17], 1u);
      if (rem) fb_add(&S.cgfreq[0], (uint32_t)rem);
    }
  }
  FB_WSYNC();
  warp_generate(S.cgfreq, kNumCodegen, 7, S.cglen, S.cgcode, S);

  // dynamic_size (hbw:335-360) + payload bits
  const int order[kNumCodegen] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int num_codegens = kNumCodegen;
  while (num_codegens > 4 && S.cgfreq[order[num_codegens - 1]] == 0) num_codegens--;
  int size = 0;
  int extra = 0; // extra bits, not part of `size` (callers pass 0)
  FB_PFOR(i, kNumLit + kNumDist) {
    if (i < kNumLit) {
      size += (int)(S.freq[i] * S.len[i]);
      if (i >= kLenCodesStart + 8 && i < kLenCodesStart + 28) extra += (int)(S.freq[i] * (uint32_t)((i - kLenCodesStart - 4) >> 2));
    } else if (kind == kKindDynamic) {
      const int d = i - kNumLit;
      size += (int)(S.freq[i] * S.len[i]);
      if (d >= 4) extra += (int)(S.freq[i] * (uint32_t)((d - 2) >> 1));
    }
  }
  FB_PFOR(i, kNumCodegen) size += (int)(S.cgfreq[i] * S.cglen[i]);
  size = wsum(size);
  extra = wsum(extra);
  const int data_bits = size - 0; // lit + offset + codegen code bits so far
  int cg_bits = 0;
  FB_PFOR(i, kNumCodegen) cg_bits += (int)(S.cgfreq[i] * S.cglen[i]);
  cg_bits = wsum(cg_bits);
  if (kind != kKindDynamic) size += 1; // offset_freq[0] (forced to 1, hbw:756) * huff_offset.codes[0].len
  size += 3 + 5 + 5 + 4 + 3 * num_codegens + (int)(S.cgfreq[16] * 2 + S.cgfreq[17] * 3 + S.cgfreq[18] * 7);
  // "store instead" test (hbw:526-531, :779-784; quirk D2: (size+size)>>4)
  const int ssize = ((int)n + 5) * 8;
  if (ssize < ((size + size) >> 4)) {
    res.kind = kKindStored;
    return res;
  }

  // write_dynamic_header (hbw:421-471): 3 + 5 + 5 + 4 bits, the code-length code's lengths in codegen order, then
  // the items (code, then 2 / 3 / 7 extra bits behind symbols 16 / 17 / 18).  Bit positions by prefix sum, bits
  // ORed into a staging copy of the header words.
  uint32_t *stage = S.lists; // [kHdrWords] (the package-merge lists are free by now)
  FB_PFOR(i, kHdrWords) stage[i] = 0;
  FB_WSYNC();
  const int fixed_bits = 17 + 3 * num_codegens;
  if (FB_LANE == 0) {
    BitAcc ba;
    ba.words = stage;
    ba.acc = 0; ba.nacc = 0; ba.nwords = 0;
    ba.put(4, 3); // BFINAL = 0, BTYPE = 10
    ba.put((uint32_t)(num_literals - 257), 5);
    ba.put((uint32_t)(num_offsets - 1), 5);
    ba.put((uint32_t)(num_codegens - 4), 4);
    for (int i = 0; i < num_codegens; i++) ba.put(S.cglen[order[i]], 3);
    ba.finish();
  }
  FB_WSYNC();
  int bitpos = fixed_bits;
#if defined(__CUDA_ARCH__)
  for (int base = 0; base < nitems; base += 32) {
    const int k = base + FB_LANE;
    uint32_t v = 0;
    int nb = 0;
    if (k < nitems) codegen_item_bits(S.items[k], S.cgcode, S.cglen, v, nb);
    int x = nb;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (FB_LANE >= o) x += y;
    }
    if (nb) {
      const int at = bitpos + x - nb;
      const uint64_t sh = (uint64_t)v << (at & 31);
      atomicOr(&stage[at >> 5], (uint32_t)sh);
      if ((uint32_t)(sh >> 32)) atomicOr(&stage[(at >> 5) + 1], (uint32_t)(sh >> 32));
    }
    bitpos += __shfl_sync(0xffffffffu, x, 31);
  }
#else
  for (int k = 0; k < nitems; k++) {
    uint32_t v = 0;
    int nb = 0;
    codegen_item_bits(S.items[k], S.cgcode, S.cglen, v, nb);
    const uint64_t sh = (uint64_t)v << (bitpos & 31);
    stage[bitpos >> 5] |= (uint32_t)sh;
    stage[(bitpos >> 5) + 1] |= (uint32_t)(sh >> 32);
    bitpos += nb;
  }
#endif
  FB_WSYNC();
  FB_PFOR(i, (bitpos + 31) >> 5) hdr_words[i] = stage[i];
  if (FB_LANE == 0) S.misc[0] = bitpos;
  FB_WSYNC();
  const int hdr_bits = S.misc[0];
  res.hdr_nbits = (uint32_t)hdr_bits;
  // total bits of the block = header + sum(freq*len) + extra bits
  res.blk_bits = (uint32_t)(hdr_bits + (data_bits - cg_bits) + extra);
  FB_PFOR(i, kNumLit + kNumDist) codeout[i] = (uint32_t)S.code[i] | ((uint32_t)S.len[i] << 16);
  FB_WSYNC();
  return res;
}

} // namespace fb
