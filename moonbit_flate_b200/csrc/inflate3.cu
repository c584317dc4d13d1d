// inflate3.cu -- K6 fast path: one warp per stream, the 32 lanes decode ONE block in parallel.
//
// Huffman decoding is a serial chain (the position of a symbol is known only once the previous one has been
// decoded), ~100 dependent cycles per symbol on a GPU thread; a 64 KiB segment of 25..65 k symbols costs
// milliseconds per stream however many streams run side by side.  The chain is broken speculatively:
//
//   * the bits of a block are cut into 32 equal ranges; lane i starts decoding at the first bit of range i
//     -- almost certainly not a symbol boundary -- and runs to the end of its range.  Huffman streams
//     self-synchronise: after a few (wrong) symbols the lane falls onto true boundaries, so the position
//     p_i where it leaves its range is usually the true one;
//   * lane i then takes p_{i-1} as its start and decodes its range again, counting output bytes and
//     back-references; this is repeated until no start moves (lane 0 starts at the true first bit, so
//     lane k is final after round k at the latest: usually two rounds, always <= 33);
//   * the chain is valid up to the first lane that met the end-of-block code; an exclusive scan of the
//     per-lane counts gives every lane its output position, and a last pass writes the literals to
//     their final place and records the back-references {dst, len, dist};
//   * the recorded copies are replayed in order by the warp (DictDecoder::write_copy semantics,
//     dict-decoder.mbt:114-185): records whose source lies before the first unresolved destination are
//     independent and are copied one per lane, long ones by the whole warp.
//
// Block headers, stored blocks and the tables (inflate.mbt:345-548) are handled warp-uniformly as in the
// exact kernel's fast sibling.  Anything unusual -- a header the reference rejects, an invalid symbol on
// the true path, a distance beyond the output, a slot that is too small, truncated input -- puts the
// stream on the fallback list: k_inflate (inflate.cu) then re-decodes it with the reference's exact error
// behaviour.  A stream that completes here consumed only real bits and every block ended in its EOB, so
// the reference decodes it to the same bytes with status EOF.
#include "common.cuh"
#include "kernels.h"
#include "../../include/flate_b200.h"

#include <cstdio>
#include <cstdlib>

namespace fb {
namespace par {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kLB = 10;
constexpr int kDB = 8;
constexpr int kWarps = 4;
constexpr uint32_t kLitLim = 256u << 4; // table entry = (symbol << 4) | code length
constexpr uint32_t P_OK = 0, P_EOB = 1, P_BAD = 2;
constexpr uint32_t kMinRange = 256;     // bits per lane at least
#ifndef FB_INF_LANE_COPY
#define FB_INF_LANE_COPY 16 // measured per GiB: 8 -> 15.3 ms, 12 -> 10.31, 16 -> 10.33, 24 -> 10.83
#endif
constexpr int kLaneCopyMax = FB_INF_LANE_COPY; // longest back-reference one lane copies by itself
#ifndef FB_INF_BIG_STEP
#define FB_INF_BIG_STEP 3 // measured per GiB, runs / mixed: old loop 4.48 / 9.59 ms, 3 -> 3.10 / 9.79, 9 -> 2.83 / 11.06
#endif
constexpr int kBigStep = FB_INF_BIG_STEP; // loads in flight per lane while the warp copies a long back-reference
#ifndef FB_INF_PF
#define FB_INF_PF 1 // lane read-ahead: 0 none, 1 into L1, 2 into L2 only
#endif
#ifndef FB_INF_PF_WORDS
#define FB_INF_PF_WORDS 24
#endif
constexpr uint32_t kPrefetchWords = FB_INF_PF_WORDS;  // lane read-ahead: three 32-byte sectors
#ifndef FB_INF_SYNC_BITS
#define FB_INF_SYNC_BITS 1024 // measured per GiB: 256 11.0 ms, 512 10.46, 768 10.27, 1024 10.32, 1536 10.43, 2560 10.85
#endif
constexpr uint32_t kSyncBits = FB_INF_SYNC_BITS;   // round-0 run-in of a lane (bits)

struct Tab {
  uint16_t first[16], count[16], offs[16];
};

struct Smem { // one warp
  uint16_t lit[1 << kLB];
  uint16_t dist[1 << kDB];
  uint16_t cl[128];
  uint16_t lsorted[288];
  uint16_t dsorted[32];
  Tab tl, td;
  uint8_t lens[320];
  uint8_t cl_lens[32];
};

__constant__ uint32_t c3_len_tab[32] = {
    3,           4,           5,           6,           7,           8,           9,           10,
    11 | 1 << 16, 13 | 1 << 16, 15 | 1 << 16, 17 | 1 << 16, 19 | 2 << 16, 23 | 2 << 16, 27 | 2 << 16, 31 | 2 << 16,
    35 | 3 << 16, 43 | 3 << 16, 51 | 3 << 16, 59 | 3 << 16, 67 | 4 << 16, 83 | 4 << 16, 99 | 4 << 16, 115 | 4 << 16,
    131 | 5 << 16, 163 | 5 << 16, 195 | 5 << 16, 227 | 5 << 16, 258, 0, 0, 0};
__constant__ uint32_t c3_dist_tab[32] = {
    1,            2,            3,             4,             5 | 1 << 16,    7 | 1 << 16,    9 | 2 << 16,     13 | 2 << 16,
    17 | 3 << 16, 25 | 3 << 16, 33 | 4 << 16,  49 | 4 << 16,  65 | 5 << 16,   97 | 5 << 16,   129 | 6 << 16,   193 | 6 << 16,
    257 | 7 << 16, 385 | 7 << 16, 513 | 8 << 16, 769 | 8 << 16, 1025 | 9 << 16, 1537 | 9 << 16, 2049 | 10 << 16, 3073 | 10 << 16,
    4097 | 11 << 16, 6145 | 11 << 16, 8193 | 12 << 16, 12289 | 12 << 16, 16385 | 13 << 16, 24577 | 13 << 16, 0, 0};
__constant__ uint8_t c3_code_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// HuffmanDecoder::initialize (inflate.mbt:100-223), warp-cooperative: canonical description only.
// false = the reference rejects the code (or it is empty): exact path.
__device__ bool warp_canon(const uint8_t *lens, int nsym, uint16_t *sorted, Tab *tab, int *mn_out, int *mx_out)
{
  const int lane = threadIdx.x & 31;
  int c = 0;
  if (lane >= 1 && lane <= 15)
    for (int i = 0; i < nsym; i++) c += (lens[i] == lane);
  const unsigned nz = __ballot_sync(kFull, c != 0);
  if (nz == 0) return false;
  const int mn = __ffs(nz) - 1, mx = 31 - __clz(nz);
  int code = 0, off = 0, code_at_max = 0, my_first = 0, my_off = 0;
  for (int L = 1; L <= 15; L++) {
    const int cL = __shfl_sync(kFull, c, L);
    code <<= 1;
    if (lane == L) { my_first = code; my_off = off; }
    code += cL;
    off += cL;
    if (L == mx) code_at_max = code;
  }
  if (code_at_max != (1 << mx) && !(code_at_max == 1 && mx == 1)) return false; // :161
  if (lane < 16) {
    if (lane == 0) my_first = mx; // first[0] is not a code length's entry: it keeps the longest length in use
    tab->first[lane] = (uint16_t)my_first;
    tab->count[lane] = (uint16_t)c;
    tab->offs[lane] = (uint16_t)my_off;
  }
  if (c) {
    int k = my_off;
    for (int i = 0; i < nsym; i++)
      if (lens[i] == lane) sorted[k++] = (uint16_t)i;
  }
  *mn_out = mn;
  *mx_out = mx;
  __syncwarp();
  return true;
}

__device__ void warp_fill_lut(uint16_t *lut, int lut_bits, const uint16_t *sorted, const Tab *tab, int mn, int mx)
{
  const int lane = threadIdx.x & 31;
  for (int idx = lane; idx < (1 << lut_bits); idx += 32) {
    const unsigned r = __brev((unsigned)idx);
    uint32_t e = 0;
    for (int L = mn; L <= lut_bits && L <= mx; L++) {
      const unsigned d = (r >> (32 - L)) - tab->first[L];
      if (d < tab->count[L]) {
        e = (uint32_t)((sorted[tab->offs[L] + d] << 4) | L);
        break;
      }
    }
    lut[idx] = (uint16_t)e;
  }
  __syncwarp();
}

// code longer than the direct table: (sym << 4) | len, 0 if none matches
__device__ __forceinline__ uint32_t canon_long(uint32_t bits, int from, const Tab *tab, const uint16_t *sorted)
{
  const unsigned r = __brev(bits);
  for (int L = from; L <= 15; L++) {
    const unsigned d = (r >> (32 - L)) - tab->first[L];
    if (d < tab->count[L]) return ((uint32_t)sorted[tab->offs[L] + d] << 4) | (uint32_t)L;
  }
  return 0;
}

// Warp-uniform bit reader for the headers: every lane holds the same state and issues the same loads.
struct UBits {
  const uint32_t *w;
  const uint8_t *end;
  uint64_t bb;
  int nb;
  __device__ __forceinline__ void init(const uint8_t *in, uint64_t len)
  {
    end = in + len;
    bb = 0; nb = 0;
    const uint8_t *p = in;
    while (reinterpret_cast<uintptr_t>(p) & 3) {
      if (p < end) bb |= (uint64_t)__ldg(p) << nb;
      nb += 8;
      p++;
    }
    w = reinterpret_cast<const uint32_t *>(p);
  }
  __device__ __forceinline__ void refill()
  {
    if (nb < 32) {
      uint32_t v = 0;
      if (reinterpret_cast<const uint8_t *>(w) < end) v = __ldg(w);
      bb |= (uint64_t)v << nb;
      nb += 32;
      w++;
    }
  }
  __device__ __forceinline__ uint32_t peek() const { return (uint32_t)bb; }
  __device__ __forceinline__ void drop(int n) { bb >>= n; nb -= n; }
  __device__ __forceinline__ uint32_t take(int n)
  {
    const uint32_t v = (uint32_t)bb & ((1u << n) - 1u);
    drop(n);
    return v;
  }
  __device__ __forceinline__ int64_t consumed_bits(const uint8_t *in) const
  {
    return (int64_t)(reinterpret_cast<const uint8_t *>(w) - in) * 8 - nb;
  }
};

// Per-lane bit reader: a 96-bit window (lo, hi, nx) with bit offset bo.  A decode step peeks at bo < 32 (32 valid
// bits), may peek once more further in (peek_at: offsets < 64, still 32 valid bits) and moves bo with skip();
// norm() at the end of the step brings bo back below 32 -- one refill per step for the whole warp instead of one
// per consumed code, which matters because with 32 lanes in different places the refill branch is taken by some
// lane at nearly every step.  Positions are 32-bit word indices from the aligned word at or below the first
// byte (the base pointer is warp-uniform).
struct LBits {
  const uint32_t *base;
  uint32_t nwords; // words that may be read
  uint32_t wi;     // index of `lo`
  uint32_t lo, hi, nx;
  int bo;
  __device__ __forceinline__ uint32_t ld(uint32_t i) const { return i < nwords ? __ldg(base + i) : 0u; }
  // abit = bit position counted from `base`
  __device__ __forceinline__ void init(const uint32_t *b, uint32_t nw, uint32_t abit)
  {
    base = b;
    nwords = nw;
    wi = abit >> 5;
    bo = (int)(abit & 31u);
    lo = ld(wi); hi = ld(wi + 1); nx = ld(wi + 2);
  }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, bo); } // bo < 32
  __device__ __forceinline__ uint32_t peek_at(int b) const // b < 64 (the funnel shift takes b modulo 32)
  {
    const bool up = b >= 32;
    return __funnelshift_r(up ? hi : lo, up ? nx : hi, b);
  }
  __device__ __forceinline__ void skip(int n) { bo += n; }
  __device__ __forceinline__ void norm()
  {
    while (bo >= 32) {
      lo = hi; hi = nx;
      wi++;
      nx = ld(wi + 2);
      bo -= 32;
      // every lane streams through its own range: without this each 32-byte sector is a demand miss, and
      // with 32 lanes in different places a warp would be waiting on one of them at nearly every step
#if FB_INF_PF == 2
      if ((wi & 7u) == 0u && wi + kPrefetchWords < nwords) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + wi + kPrefetchWords));
#elif FB_INF_PF == 1
      if ((wi & 7u) == 0u && wi + kPrefetchWords < nwords) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + wi + kPrefetchWords));
#endif
    }
  }
  __device__ __forceinline__ void norm1() // after at most 32 skipped bits (bo < 64)
  {
    if (bo >= 32) {
      lo = hi; hi = nx;
      wi++;
      nx = ld(wi + 2);
      bo -= 32;
#if FB_INF_PF == 2
      if ((wi & 7u) == 0u && wi + kPrefetchWords < nwords) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + wi + kPrefetchWords));
#elif FB_INF_PF == 1
      if ((wi & 7u) == 0u && wi + kPrefetchWords < nwords) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + wi + kPrefetchWords));
#endif
    }
  }
  __device__ __forceinline__ uint32_t abit() const { return (wi << 5) + (uint32_t)bo; }
};

#ifdef FB_INFLATE_STEPSTAT
__device__ unsigned long long g_stepstat[2];
#endif

// shared-memory loads by 32-bit shared address (the tables are reached through a reference, which would
// otherwise cost a generic-to-shared conversion per access)
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr)
{
  unsigned short v;
  asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

// Every lane with `run` decodes from bit `start` until it reaches bit `e` (or EOB, or something invalid);
// bit positions are relative to `in`.  WRITE: literals go to out[obase + ...], back-references to rec[...]
// with absolute destinations.  One symbol per step and lane; the step is straight-line code, and its two
// optional parts (a code longer than the direct table, a length/distance pair) are entered by the whole
// warp when any lane needs them.
// RING (the CTA-per-stream kernel): `out` is a ring of kRing bytes in shared memory that holds the output by
// absolute position modulo kRing; the copies are replayed there before the block leaves for global memory.
constexpr uint32_t kRing = 128u * 1024u;
constexpr uint32_t kRingMask = kRing - 1u;

template <bool WRITE, bool RING = false>
__device__ __forceinline__ void decode_ranges(const Smem &sm, const uint32_t *s_len_tab, const uint32_t *s_dist_tab,
                                              const uint8_t *in, int64_t cur_len, uint32_t bend, bool run,
                                              uint32_t start, uint32_t e, uint32_t &p_out, uint32_t &flag_out,
                                              uint32_t &n_out, uint32_t &n_rec, uint8_t *out, uint32_t obase,
                                              uint2 *rec, uint32_t h0 = 0)
{
  const uint32_t lit_sa = (uint32_t)__cvta_generic_to_shared(sm.lit);
  const uint32_t dist_sa = (uint32_t)__cvta_generic_to_shared(sm.dist);
  const uint32_t ltab_sa = (uint32_t)__cvta_generic_to_shared(s_len_tab);
  const uint32_t dtab_sa = (uint32_t)__cvta_generic_to_shared(s_dist_tab);
  const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3);
  const uint32_t *wbase = reinterpret_cast<const uint32_t *>(in - lead);
  const uint32_t lead_bits = lead * 8u;
  LBits lb;
  lb.init(wbase, (uint32_t)((cur_len + lead + 3) >> 2), (run ? start : 0u) + lead_bits);
  const uint32_t e_abs = e + lead_bits, bend_abs = bend + lead_bits;
  const uint32_t lim_abs = e_abs < bend_abs + 1u ? e_abs : bend_abs + 1u; // a further symbol may start below this
  const bool long_lit = sm.tl.first[0] > (uint32_t)kLB, long_dist = sm.td.first[0] > (uint32_t)kDB; // codes beyond the direct tables?
  uint32_t flag = P_OK, cnt_out = 0, cnt_rec = 0;
  uint8_t *op = out + obase;
  bool act = run && start < e;
#ifdef FB_INFLATE_STEPSTAT
  uint32_t my_steps = 0, warp_steps = 0;
#endif
  while (__any_sync(kFull, act)) {
#ifdef FB_INFLATE_STEPSTAT
    my_steps += act; warp_steps++;
#endif
    const uint32_t bits = lb.peek();
    uint32_t e0 = lds_u16(lit_sa + ((bits & ((1u << kLB) - 1u)) << 1));
    if (long_lit && __any_sync(kFull, act && e0 == 0)) {
      if (act && e0 == 0) e0 = canon_long(bits, kLB + 1, &sm.tl, sm.lsorted);
    }
    const uint32_t cl = e0 & 15u, sym = e0 >> 4;
    const bool is_len = act && (sym - 257u) < 29u;
    uint32_t adv = cl, length = 0;
    if (__any_sync(kFull, is_len)) {
      if (is_len) {
        const uint32_t lt = lds_u32(ltab_sa + ((sym - 257u) << 2));
        const uint32_t xb = lt >> 16;
        length = (lt & 0xffffu) + ((bits >> cl) & ((1u << xb) - 1u));
        adv = cl + xb;
      }
    }
    if (act) {
      if (e0 == 0 || sym >= (uint32_t)kNumLit) { flag = P_BAD; act = false; }
      else if (sym < 256u) {
        // a second literal from the same peek, if the first leaves the lane inside its range and the next
        // code is a literal of the direct table (anything else waits for the next step)
        const uint32_t e1 = lds_u16(lit_sa + (((bits >> cl) & ((1u << kLB) - 1u)) << 1));
        const uint32_t c1 = e1 & 15u, s1 = e1 >> 4;
        const uint32_t ab0 = lb.abit() + cl;
        const bool two = e1 != 0 && s1 < 256u && ab0 < lim_abs;
        const uint32_t e2 = lds_u16(lit_sa + (((bits >> (cl + c1)) & ((1u << kLB) - 1u)) << 1));
        const uint32_t c2 = e2 & 15u, s2 = e2 >> 4;
        const uint32_t ab1 = ab0 + c1;
        // (the third look-up needs its 10 index bits inside the 32-bit peek: cl + c1 <= 22)
        const bool three = two && cl + c1 + (uint32_t)kLB <= 32u && e2 != 0 && s2 < 256u && ab1 < lim_abs;
        if (WRITE && !RING) {
          op[cnt_out] = (uint8_t)sym;
          if (two) op[cnt_out + 1] = (uint8_t)s1;
          if (three) op[cnt_out + 2] = (uint8_t)s2;
        }
        if (WRITE && RING) {
          const uint32_t at = obase + cnt_out;
          out[at & kRingMask] = (uint8_t)sym;
          if (two) out[(at + 1u) & kRingMask] = (uint8_t)s1;
          if (three) out[(at + 2u) & kRingMask] = (uint8_t)s2;
        }
        cnt_out += 1u + (two ? 1u : 0u) + (three ? 1u : 0u);
        lb.skip((int)(cl + (two ? c1 : 0u) + (three ? c2 : 0u)));
      } else {
        lb.skip((int)adv); // <= 20 bits: the distance below is still inside the window
        if (sym == 256u) { flag = P_EOB; act = false; }
      }
    }
    if (__any_sync(kFull, is_len && act)) {
      if (is_len && act) {
        const uint32_t dbits = lb.peek_at(lb.bo);
        uint32_t d = lds_u16(dist_sa + ((dbits & ((1u << kDB) - 1u)) << 1));
        if (long_dist && d == 0) d = canon_long(dbits, kDB + 1, &sm.td, sm.dsorted);
        if (d == 0 || (d >> 4) >= (uint32_t)kNumDist) { flag = P_BAD; act = false; }
        else {
          const uint32_t dl = d & 15u;
          const uint32_t dt = lds_u32(dtab_sa + ((d >> 4) << 2));
          const uint32_t dxb = dt >> 16;
          lb.skip((int)(dl + dxb));
          if (WRITE) {
            const uint32_t dist = (dt & 0xffffu) + ((dbits >> dl) & ((1u << dxb) - 1u));
            const uint32_t at = obase + cnt_out;
            if (dist > at + h0) { flag = P_BAD; act = false; } // dist > hist_size (inflate.mbt:677), h0 = dictionary
            else rec[cnt_rec] = make_uint2(at, length | ((dist - 1u) << 16));
          }
          cnt_out += length;
          cnt_rec++;
        }
      }
    }
    lb.norm();
    if (act) {
      const uint32_t ab = lb.abit();
      if (ab > bend_abs) { flag = P_BAD; act = false; } // ran past the end of the input
      else if (ab >= e_abs) act = false;
    }
  }
#ifdef FB_INFLATE_STEPSTAT
  if (WRITE) {
    const uint32_t tot = __reduce_add_sync(kFull, my_steps);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&g_stepstat[0], (unsigned long long)warp_steps * 32ull); atomicAdd(&g_stepstat[1], (unsigned long long)tot); }
  }
#endif
  p_out = (run && start < e) ? lb.abit() - lead_bits : start;
  flag_out = flag;
  n_out = cnt_out;
  n_rec = cnt_rec;
}

// The same for a block whose code has no length symbols (HLIT = 257: the literal-only blocks the encoder
// emits for incompressible data, which resynchronise poorly and therefore run many rounds): nothing but
// literals and EOB, up to three symbols per step.
// PHASES (counting only): the start is a guess that may be off by less than `nphase` bits (nphase = the longest
// code).  A code whose symbols nearly all have one length never falls back onto the true symbol boundaries from a
// wrong start -- but it runs into the end-of-block symbol (the rarest one, so about once per 2^length symbols)
// long before the range is over, which the true sequence cannot do in front of the block's last range.  So a lane
// that meets EOB inside its range starts over one bit further on; *start_used is where the try that got through
// (or the last one) began.
template <bool WRITE, bool RING = false, bool PHASES = false>
__device__ __forceinline__ void decode_ranges_lit(const Smem &sm, const uint8_t *in, int64_t cur_len, uint32_t bend,
                                                  bool run, uint32_t start, uint32_t e, uint32_t &p_out,
                                                  uint32_t &flag_out, uint32_t &n_out, uint8_t *out, uint32_t obase,
                                                  uint32_t nphase = 1, uint32_t *start_used = nullptr)
{
  uint32_t phase = 0;
  const uint32_t lit_sa = (uint32_t)__cvta_generic_to_shared(sm.lit);
  const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3);
  const uint32_t *wbase = reinterpret_cast<const uint32_t *>(in - lead);
  const uint32_t lead_bits = lead * 8u;
  LBits lb;
  lb.init(wbase, (uint32_t)((cur_len + lead + 3) >> 2), (run ? start : 0u) + lead_bits);
  const uint32_t e_abs = e + lead_bits, bend_abs = bend + lead_bits;
  const uint32_t lim_abs = e_abs < bend_abs + 1u ? e_abs : bend_abs + 1u; // a further symbol may start below this
  const bool long_lit = sm.tl.first[0] > (uint32_t)kLB; // codes beyond the direct table?
  uint32_t flag = P_OK, cnt_out = 0;
  uint8_t *op = out + obase;
  bool act = run && start < e;
  while (__any_sync(kFull, act)) {
    const uint32_t bits = lb.peek();
    uint32_t e0 = lds_u16(lit_sa + ((bits & ((1u << kLB) - 1u)) << 1));
    if (long_lit && __any_sync(kFull, act && e0 == 0)) {
      if (act && e0 == 0) e0 = canon_long(bits, kLB + 1, &sm.tl, sm.lsorted);
    }
    if (act) {
      const uint32_t c0 = e0 & 15u, s0 = e0 >> 4;
      if (e0 == 0 || s0 > 256u) { flag = P_BAD; act = false; }
      else if (s0 == 256u) {
        lb.skip((int)c0);
        if (PHASES && phase + 1u < nphase && lb.abit() < e_abs) { // a wrong start: the next one
          phase++;
          lb.init(wbase, (uint32_t)((cur_len + lead + 3) >> 2), start + phase + lead_bits);
          cnt_out = 0;
        } else { flag = P_EOB; act = false; }
      } else {
        // second symbol from the same 32-bit peek, if the first one leaves the lane inside its range and the
        // second is a literal of the direct table (anything else waits for the next step)
        const uint32_t e1 = lds_u16(lit_sa + (((bits >> c0) & ((1u << kLB) - 1u)) << 1));
        const uint32_t c1 = e1 & 15u, s1 = e1 >> 4;
        const uint32_t ab0 = lb.abit() + c0;
        const bool two = e1 != 0 && s1 < 256u && ab0 < lim_abs;
        const uint32_t e2 = lds_u16(lit_sa + (((bits >> (c0 + c1)) & ((1u << kLB) - 1u)) << 1));
        const uint32_t c2 = e2 & 15u, s2 = e2 >> 4;
        const uint32_t ab1 = ab0 + c1;
        const bool three = two && c0 + c1 + (uint32_t)kLB <= 32u && e2 != 0 && s2 < 256u && ab1 < lim_abs;
        if (WRITE && !RING) {
          op[cnt_out] = (uint8_t)s0;
          if (two) op[cnt_out + 1] = (uint8_t)s1;
          if (three) op[cnt_out + 2] = (uint8_t)s2;
        }
        if (WRITE && RING) {
          const uint32_t at = obase + cnt_out;
          out[at & kRingMask] = (uint8_t)s0;
          if (two) out[(at + 1u) & kRingMask] = (uint8_t)s1;
          if (three) out[(at + 2u) & kRingMask] = (uint8_t)s2;
        }
        cnt_out += 1u + (two ? 1u : 0u) + (three ? 1u : 0u);
        lb.skip((int)(c0 + (two ? c1 : 0u) + (three ? c2 : 0u)));
      }
    }
    lb.norm1();
    if (act) {
      const uint32_t ab = lb.abit();
      if (ab > bend_abs) { flag = P_BAD; act = false; } // ran past the end of the input
      else if (ab >= e_abs) act = false;
    }
  }
  p_out = (run && start < e) ? lb.abit() - lead_bits : start;
  flag_out = flag;
  n_out = cnt_out;
  if (PHASES) *start_used = start + phase;
}

// Replays records [0, nrec) in order, one warp.  RING: the output lives in a shared-memory ring (position modulo
// kRing, the CTA-per-stream kernel); otherwise `buf` is the stream's output slot in global memory.
template <bool RING>
__device__ void replay_records(uint8_t *buf, const uint2 *rec, uint32_t nrec, int lane)
{
  // positions are relative to the stream's output slot; a source may lie in front of it (preset dictionary): signed
  auto ld = [&](int64_t pos) -> uint8_t { return RING ? buf[(uint32_t)pos & kRingMask] : __ldcg(buf + pos); };
  auto st = [&](int64_t pos, uint8_t v) {
    if (RING) buf[(uint32_t)pos & kRingMask] = v;
    else buf[pos] = v;
  };
  uint2 nxt = make_uint2(0u, 0u); // the next group's records are requested while this group is being copied
  if ((uint32_t)lane < nrec) nxt = rec[lane];
  for (uint32_t g = 0; g < nrec; g += 32) {
    const uint32_t r = g + (uint32_t)lane;
    uint32_t dst = 0, len = 0, dist = 1;
    if (r < nrec) {
      const uint2 v = nxt;
      dst = v.x; len = v.y & 0xffffu; dist = (v.y >> 16) + 1u;
    }
    if (r + 32u < nrec) nxt = rec[r + 32u];
    const bool big = len > (uint32_t)kLaneCopyMax;
    // source interval: [dst - dist, dst - dist + min(len, dist))
    const uint32_t src = dst - dist;
    const uint32_t src_end = src + (len < dist ? len : dist);
    unsigned pending = __ballot_sync(kFull, r < nrec);
    while (pending) {
      const int p = __ffs(pending) - 1;
      const uint32_t dstp = __shfl_sync(kFull, dst, p);
      const bool bigp = __shfl_sync(kFull, (int)big, p) != 0;
      if (bigp) { // the whole warp copies record p
        const uint32_t lenp = __shfl_sync(kFull, len, p);
        const uint32_t dd = __shfl_sync(kFull, dist, p);
        const int64_t sp = (int64_t)dstp - (int64_t)dd;
        // Byte i of the copy is source byte i mod dd (the reference copies forward byte by byte, so a distance
        // shorter than the length repeats the pattern: dict-decoder.mbt:136-149), and all of those lie in front of
        // the destination: no dependencies inside the record.  kBigStep * 32 bytes per step with the loads first:
        // a 258-byte record costs three trips to memory instead of nine.
        if (dd >= lenp) { // (the usual case: source and destination do not overlap)
          if (lenp <= 32u) {
            if ((uint32_t)lane < lenp) st(dstp + lane, ld(sp + lane));
          } else
          for (uint32_t b0 = lane; b0 < lenp; b0 += 32u * kBigStep) {
            uint8_t v[kBigStep];
#pragma unroll
            for (int k = 0; k < kBigStep; k++)
              if (b0 + 32u * k < lenp) v[k] = ld(sp + b0 + 32u * k);
#pragma unroll
            for (int k = 0; k < kBigStep; k++)
              if (b0 + 32u * k < lenp) st(dstp + b0 + 32u * k, v[k]);
          }
        } else {
          uint32_t idx = dd > 31u ? (uint32_t)lane : (uint32_t)lane % dd;
          const uint32_t r32 = dd > 32u ? 32u : 32u % dd;
          for (uint32_t b0 = lane; b0 < lenp; b0 += 32u * kBigStep) {
            uint8_t v[kBigStep];
#pragma unroll
            for (int k = 0; k < kBigStep; k++) {
              if (b0 + 32u * k < lenp) v[k] = ld(sp + idx);
              idx += r32;
              if (idx >= dd) idx -= dd;
            }
#pragma unroll
            for (int k = 0; k < kBigStep; k++)
              if (b0 + 32u * k < lenp) st(dstp + b0 + 32u * k, v[k]);
          }
        }
        pending &= ~(1u << p);
        __syncwarp();
        continue;
      }
      // parallel round: every pending small record whose source lies before the first unresolved destination (it
      // reads nothing a pending record -- small or big -- still has to write)
      const bool mine = ((pending >> lane) & 1u) && !big && (lane == p || src_end <= dstp);
      const unsigned ready = __ballot_sync(kFull, mine);
      if ((ready >> lane) & 1u) {
        if (RING) {
          uint8_t v[kLaneCopyMax];
          if (dist >= len) {
#pragma unroll
            for (int k = 0; k < kLaneCopyMax; k++)
              if ((uint32_t)k < len) v[k] = buf[(src + (uint32_t)k) & kRingMask];
          } else {
#pragma unroll
            for (int k = 0; k < kLaneCopyMax; k++)
              if ((uint32_t)k < len) v[k] = buf[(src + ((uint32_t)k % dist)) & kRingMask];
          }
#pragma unroll
          for (int k = 0; k < kLaneCopyMax; k++)
            if ((uint32_t)k < len) buf[(dst + (uint32_t)k) & kRingMask] = v[k];
        } else {
          uint8_t *dp = buf + dst;
          const uint8_t *sp = dp - dist;
          uint8_t v[kLaneCopyMax];
          if (dist >= len) {
#pragma unroll
            for (int k = 0; k < kLaneCopyMax; k++)
              if ((uint32_t)k < len) v[k] = __ldcg(sp + k);
          } else {
#pragma unroll
            for (int k = 0; k < kLaneCopyMax; k++)
              if ((uint32_t)k < len) v[k] = __ldcg(sp + ((uint32_t)k % dist));
          }
#pragma unroll
          for (int k = 0; k < kLaneCopyMax; k++)
            if ((uint32_t)k < len) dp[k] = v[k];
        }
      }
      pending &= ~ready;
      __syncwarp();
    }
  }
}

template <int MINB>
__global__ void __launch_bounds__(kWarps * 32, MINB) k_inflate_par(InflateJob j)
{
  __shared__ Smem smem_all[kWarps];
  __shared__ uint32_t s_len_tab[32], s_dist_tab[32];
  Smem &sm = smem_all[threadIdx.x >> 5];
  if (threadIdx.x < 32) {
    s_len_tab[threadIdx.x] = c3_len_tab[threadIdx.x];
    s_dist_tab[threadIdx.x] = c3_dist_tab[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;

  for (;;) {
    uint32_t st32 = 0;
    if (lane == 0) st32 = atomicAdd(&j.counters[0], 1u);
    st32 = __shfl_sync(kFull, st32, 0);
    if (st32 >= j.nstreams) break;
    if (j.order) st32 = j.order[st32]; // device-resident calls: longest streams first
    if (j.avail) { // host-buffer call: wait until the H2D stream has delivered this stream's bytes
      if (lane == 0)
        while (*(volatile const uint32_t *)j.avail <= st32) __nanosleep(500);
      __syncwarp();
    }
    const uint8_t *in0 = j.comp + j.comp_off[st32];
    const uint8_t *in = in0;
    const uint64_t in_len = j.comp_off[st32 + 1] - j.comp_off[st32];
    uint8_t *out = j.out + j.out_off[st32];
    const uint64_t cap64 = j.out_off[st32 + 1] - j.out_off[st32];
    const uint32_t cap = cap64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)cap64;
    uint2 *rec = j.records + j.rec_off[st32];
    const uint32_t rec_cap = (uint32_t)(j.rec_off[st32 + 1] - j.rec_off[st32]);
    uint32_t nrec = 0, opos = 0;
    const uint32_t h0 = j.hist0 ? j.hist0[st32] : 0u; // preset dictionary in front of the slot
    int64_t cur_len = (int64_t)in_len; // bytes from `in` (re-based after every block) to the end
    UBits ub;
    ub.init(in, in_len);
    bool bail = in_len > 0x0fffffffull; // keep bit positions comfortably inside 32 bits
    bool done = false;
    uint32_t win_bits = j.window_bits ? j.window_bits : 96u * 1024u * 8u; // first block: a 65535-byte block fits

    while (!bail && !done) {
      // ---- block header (all lanes, uniform): next_block (inflate.mbt:345-379) ----
      ub.refill();
      const int final_flag = (int)ub.take(1);
      const int typ = (int)ub.take(2);
#ifdef FB_INFLATE_DEBUG
      if (lane == 0 && opos > 102800000u) printf("hdr: opos=%u final=%d typ=%d cur_len=%lld in_off=%lld nrec=%u win=%u\n", opos, final_flag, typ, (long long)cur_len, (long long)(in - in0), nrec, win_bits);
#endif
      if (typ == 3) { bail = true; break; }
      if (typ == 0) { // stored block (data_block / copy_data, :708-766)
        const int64_t p = (ub.consumed_bits(in) + 7) >> 3;
        if (p + 4 > cur_len) { bail = true; break; }
        const uint32_t sn = (uint32_t)__ldg(in + p) | ((uint32_t)__ldg(in + p + 1) << 8);
        const uint32_t nn = (uint32_t)__ldg(in + p + 2) | ((uint32_t)__ldg(in + p + 3) << 8);
        if (nn != ((~sn) & 0xffffu) || p + 4 + sn > cur_len || (uint64_t)opos + sn > cap) { bail = true; break; }
        const uint32_t sp = (uint32_t)(p + 4);
        for (uint32_t i = lane; i < sn; i += 32) out[opos + i] = __ldg(in + sp + i);
        opos += sn;
        cur_len -= (int64_t)sp + sn;
        in += sp + sn;
        ub.init(in, (uint64_t)cur_len);
        __syncwarp();
        if (final_flag) done = true;
        continue;
      }
      int mn1 = 0, mx1 = 0, mn2 = 0, mx2 = 0;
      bool lit_only = false;
      if (typ == 1) { // fixed_huffman_decoder (:886-939); distances are 5-bit codes
        for (int i = lane; i < 288; i += 32) sm.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        sm.lens[288 + lane] = 5;
        __syncwarp();
        warp_canon(sm.lens, 288, sm.lsorted, &sm.tl, &mn1, &mx1);
        warp_canon(sm.lens + 288, 32, sm.dsorted, &sm.td, &mn2, &mx2);
      } else { // read_huffman (:429-548)
        ub.refill();
        const int nlit = (int)ub.take(5) + 257;
        const int ndist = (int)ub.take(5) + 1;
        const int nclen = (int)ub.take(4) + 4;
        if (nlit > kNumLit || ndist > kNumDist) { bail = true; break; }
        lit_only = nlit == 257; // no length symbols in the code
        if (lane < 19) sm.cl_lens[lane] = 0;
        __syncwarp();
        for (int i = 0; i < nclen; i++) {
          ub.refill();
          const uint32_t v = ub.take(3);
          if (lane == 0) sm.cl_lens[c3_code_order[i]] = (uint8_t)v;
        }
        __syncwarp();
        int mnc = 0, mxc = 0;
        if (!warp_canon(sm.cl_lens, 19, sm.dsorted, &sm.td, &mnc, &mxc)) { bail = true; break; }
        warp_fill_lut(sm.cl, 7, sm.dsorted, &sm.td, mnc, mxc);
        bool herr = false;
        const int n = nlit + ndist;
        int i = 0, prev = 0;
        while (i < n) {
          ub.refill();
          const uint32_t e = sm.cl[ub.peek() & 127];
          const int len = (int)(e & 15), x = (int)(e >> 4);
          if (len == 0) { herr = true; break; }
          ub.drop(len);
          if (x < 16) {
            if (lane == 0) sm.lens[i] = (uint8_t)x;
            prev = x;
            i++;
            continue;
          }
          int rep, b = 0;
          if (x == 16) {
            if (i == 0) { herr = true; break; }
            b = prev;
            rep = 3 + (int)ub.take(2);
          } else if (x == 17) rep = 3 + (int)ub.take(3);
          else rep = 11 + (int)ub.take(7);
          if (i + rep > n) { herr = true; break; }
          for (int k = lane; k < rep; k += 32) sm.lens[i + k] = (uint8_t)b;
          i += rep;
          prev = b;
        }
        __syncwarp();
        if (herr) { bail = true; break; }
        uint8_t dl = 0;
        if (lane < ndist) dl = sm.lens[nlit + lane];
        __syncwarp();
        sm.lens[288 + lane] = (lane < ndist) ? dl : 0;
        __syncwarp();
        if (sm.lens[kEob] == 0) { bail = true; break; }
        if (!warp_canon(sm.lens, nlit, sm.lsorted, &sm.tl, &mn1, &mx1)) { bail = true; break; }
        if (!warp_canon(sm.lens + 288, ndist, sm.dsorted, &sm.td, &mn2, &mx2)) { bail = true; break; }
      }
      warp_fill_lut(sm.lit, kLB, sm.lsorted, &sm.tl, mn1, mx1);
      warp_fill_lut(sm.dist, kDB, sm.dsorted, &sm.td, mn2, mx2);
      // a code whose symbols nearly all share one length (incompressible data) barely self-synchronises
      bool flat_code;
      {
        const uint32_t c = lane < 16 ? sm.tl.count[lane] : 0u;
        const uint32_t mxc = __reduce_max_sync(kFull, c), tot = __reduce_add_sync(kFull, c);
        flat_code = mxc * 10u >= tot * 8u;
      }

      // ---- block body: 32 ranges, speculative starts, fixpoint over the hand-over positions ----
      // The ranges are cut from a WINDOW of the input behind the header, not from all that is left of the stream:
      // where the block ends is not known, and in a stream of many blocks everything behind its EOB would be
      // decoded for nothing (a 64 MiB stream spent 15 ms per block that way).  The window is the whole rest for a
      // stream that is about one block long, otherwise 1.25 x the size of the previous block; a block that is
      // longer than its window simply goes on with another window behind it (same tables).
      const int64_t b0s = ub.consumed_bits(in);
      if (b0s >= cur_len * 8) { bail = true; break; }
      uint32_t b0 = (uint32_t)b0s;
      const uint32_t bend = (uint32_t)(cur_len * 8), blk_first_bit = b0;
      uint32_t eob = 0;
      for (bool have_eob = false; !have_eob;) {
        const uint32_t wend = (uint64_t)b0 + win_bits < bend ? b0 + win_bits : bend;
        uint32_t R = (wend - b0 + 31u) / 32u;
        if (R < kMinRange) R = kMinRange;
        const uint64_t s64 = (uint64_t)b0 + (uint64_t)lane * R;
        const uint32_t s_nom = s64 < wend ? (uint32_t)s64 : wend;
        const uint32_t e_i = (s64 + R < wend) ? (uint32_t)(s64 + R) : wend;
        // Round 0 only has to find where each lane leaves its range, and a wrong start falls onto true symbol
        // boundaries within a few dozen symbols: every lane decodes just the last kSyncBits of their range (the
        // whole range when most codes have the same length, which synchronises poorly).  A lane that did not
        // synchronise is caught by the fixpoint below like any other wrong hand-over.
        uint32_t start = s_nom, p = 0, flag = P_OK, n_out = 0, n_rec = 0;
        if (!flat_code && e_i - s_nom > kSyncBits) start = e_i - kSyncBits; // lane 0 too: its full pass is round 1
        bool need = true;
        for (int round = 0; round < 34; round++) {
          uint32_t tp, tf, to, tr = 0;
          if (lit_only && flat_code) {
            // (round 0 starts every lane but the first at a guess: try the other phases too)
            uint32_t su = start;
            decode_ranges_lit<false, false, true>(sm, in, cur_len, bend, need, start, e_i, tp, tf, to, nullptr, 0u,
                                                  (round == 0 && lane > 0) ? (uint32_t)mx1 : 1u, &su);
            if (need) start = su;
          } else if (lit_only) decode_ranges_lit<false>(sm, in, cur_len, bend, need, start, e_i, tp, tf, to, nullptr, 0u);
          else decode_ranges<false>(sm, s_len_tab, s_dist_tab, in, cur_len, bend, need, start, e_i, tp, tf, to, tr,
                                    nullptr, 0u, nullptr);
          if (need) { p = tp; flag = tf; n_out = to; n_rec = tr; }
          const uint32_t pp = __shfl_up_sync(kFull, p, 1);
          const uint32_t pf = __shfl_up_sync(kFull, flag, 1);
          const uint32_t ns = lane == 0 ? b0 : (pf == P_OK ? pp : start);
          need = ns != start;
          start = ns;
          if (lane == 0) atomicAdd(&j.counters[5], 1u); // instrumentation: decode rounds
          if (!__any_sync(kFull, need)) break;
        }
        // the chain is valid up to the first lane that does not hand over: it must have met EOB -- or every lane
        // hands over and the block goes on behind the window
        const unsigned notok = __ballot_sync(kFull, flag != P_OK);
        int f = 31;
        if (notok == 0) {
          if (wend >= bend) { bail = true; break; } // no EOB before the end of the input
        } else {
          f = __ffs(notok) - 1;
          if (__shfl_sync(kFull, flag, f) != P_EOB) { bail = true; break; }
        }
        const bool mine = lane <= f;
        uint32_t xo = mine ? n_out : 0u, xr = mine ? n_rec : 0u;
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t yo = __shfl_up_sync(kFull, xo, o), yr = __shfl_up_sync(kFull, xr, o);
          if (lane >= o) { xo += yo; xr += yr; }
        }
        const uint32_t tot_o = __shfl_sync(kFull, xo, 31), tot_r = __shfl_sync(kFull, xr, 31);
        if (tot_o > cap - opos || tot_r > rec_cap - nrec) { bail = true; break; }
        {
          uint32_t tp, tf, to, tr = 0;
          if (lit_only)
            decode_ranges_lit<true>(sm, in, cur_len, bend, mine, start, e_i, tp, tf, to, out,
                                    opos + xo - (mine ? n_out : 0u));
          else
            decode_ranges<true>(sm, s_len_tab, s_dist_tab, in, cur_len, bend, mine, start, e_i, tp, tf, to, tr, out,
                                opos + xo - (mine ? n_out : 0u), rec + nrec + xr - (mine ? n_rec : 0u), h0);
          const bool bad = mine && (tf == P_BAD || tp != p || to != n_out || tr != n_rec);
          if (__any_sync(kFull, bad)) { bail = true; break; }
        }
        opos += tot_o;
        nrec += tot_r;
        const uint32_t pend = __shfl_sync(kFull, p, f);
#ifdef FB_INFLATE_DEBUG
        if (lane == 0 && opos > 102800000u) printf("  win: b0=%u wend=%u bend=%u f=%d notok=%08x tot_o=%u pend=%u\n", b0, wend, bend, f, notok, tot_o, pend);
#endif
        if (notok == 0) {
          if (pend <= b0 || pend >= bend) { bail = true; break; } // (no progress cannot happen: ranges are >= 256 bits)
          b0 = pend; // the block goes on behind the window, which doubles: the guess was too small
          if (win_bits < 0x40000000u) win_bits *= 2u;
        } else {
          eob = pend;
          have_eob = true;
        }
      }
      if (bail) break;
      if (lane == 0) atomicAdd(&j.counters[6], 1u); // instrumentation: blocks
      {
        const uint64_t w = ((uint64_t)(eob - blk_first_bit) * 5u) >> 2;
        const uint32_t wmin = 24u * 1024u * 8u; // a tiny block (a run of zeros) says little about the next one
        win_bits = w < wmin ? wmin : (w > 0x7fffffffull ? 0x7fffffffu : (uint32_t)w);
        if (j.window_bits == 0xffffffffu) win_bits = 0xffffffffu; // experiment switch: no windows
      }
      // continue after the EOB: re-base the uniform reader there
      if ((int64_t)eob > cur_len * 8) { bail = true; break; }
      in += eob >> 3;
      cur_len -= (int64_t)(eob >> 3);
      ub.init(in, (uint64_t)cur_len);
      ub.refill();
      ub.drop((int)(eob & 7u));
      __syncwarp();
      if (final_flag) done = true;
    }

    if (!bail && ub.consumed_bits(in) > cur_len * 8) bail = true;
    if (!bail) {
      __syncwarp();
      replay_records<false>(out, rec, nrec, lane);
    }
    if (lane == 0) {
      if (bail) {
        const uint32_t k = atomicAdd(&j.counters[2], 1u);
        j.fallback[k] = st32;
      } else {
        const int64_t cb = ub.consumed_bits(in);
        j.out_len[st32] = opos;
        j.status[st32] = FB200_ST_EOF;
        j.err_off[st32] = 0;
        if (j.consumed) j.consumed[st32] = (uint64_t)((int64_t)(in - in0) + ((cb + 7) >> 3));
      }
    }
    __syncwarp();
    if (j.group_done) { // host-buffer call: publish finished output groups so their D2H copy can start
      __threadfence();
      if (lane == 0) {
        const uint32_t g = st32 / j.group_streams;
        const uint32_t first = g * j.group_streams;
        const uint32_t cnt = (uint32_t)(j.nstreams - first < j.group_streams ? j.nstreams - first : j.group_streams);
        if (atomicAdd(&j.group_done[g], 1u) + 1u == cnt) {
          __threadfence_system();
          j.group_flag[g] = 1u;
        }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One CTA per stream: the same speculative decode with 32 * CW ranges per block, for calls with few (long) streams
// -- the Reader of one big stream, a handful of files -- where a warp per stream leaves the GPU idle and the
// stream crawls (17 blocks of a 1 MiB stream one after the other took 25 ms).  Warp 0 reads the block header and
// builds the tables; all warps decode ranges; hand-over positions, counts and the end of the chain travel
// through shared memory; the output of a window is assembled in a shared-memory ring -- literals written there,
// the recorded copies replayed there by warp 0 (rounds of shared-memory latency, no barrier; a CTA-wide replay with
// two barriers per round measured 4 x slower: the copies are a dependency chain of ~460 rounds per 64 KiB whatever
// the number of lanes) -- and leaves for global memory in one coalesced sweep.
#ifndef FB_CTA_WARPS
#define FB_CTA_WARPS 16 // (<= 32: one lane per warp in the replay's reductions)
#endif
constexpr int CW = FB_CTA_WARPS; // warps per CTA
constexpr int CL = CW * 32;  // ranges per block

struct CtaShared {
  Smem sm;
  uint32_t len_tab[32], dist_tab[32];
  uint32_t p[CL], flag[CL];
  uint32_t wo[CW], wr[CW];
  // block header, written by warp 0
  int bail, final_flag, typ, lit_only, flat_code;
  int64_t b0s;
  uint32_t sn, sp;
  // reductions
  uint32_t first_bad;
};

#ifdef FB_CTA_PROF
__device__ unsigned long long g_cta_prof[8]; // cycles of thread 0: header, rounds, write pass, replay, copy-out, other
#define CTA_T(k) do { if (tid == 0) { const long long t__ = clock64(); prof[k] += t__ - tprev; tprev = t__; } } while (0)
#else
#define CTA_T(k) do { } while (0)
#endif

__global__ void __launch_bounds__(CL, 1) k_inflate_cta(InflateJob j)
{
  extern __shared__ __align__(16) uint8_t cta_smem[];
  CtaShared &S = *reinterpret_cast<CtaShared *>(cta_smem);
  uint8_t *ring = cta_smem + ((sizeof(CtaShared) + 15) & ~(size_t)15); // [kRing]
  Smem &sm = S.sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef FB_CTA_PROF
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
  if (tid < 32) {
    S.len_tab[tid] = c3_len_tab[tid];
    S.dist_tab[tid] = c3_dist_tab[tid];
  }
  __syncthreads();

  for (uint32_t st32 = blockIdx.x; st32 < j.nstreams; st32 += gridDim.x) {
    if (j.avail) { // host-buffer call: wait until the H2D stream has delivered this stream's bytes
      if (tid == 0)
        while (*(volatile const uint32_t *)j.avail <= st32) __nanosleep(500);
      __syncthreads();
    }
    const uint8_t *in0 = j.comp + j.comp_off[st32];
    const uint8_t *in = in0;
    const uint64_t in_len = j.comp_off[st32 + 1] - j.comp_off[st32];
    uint8_t *out = j.out + j.out_off[st32];
    const uint64_t cap64 = j.out_off[st32 + 1] - j.out_off[st32];
    const uint32_t cap = cap64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)cap64;
    uint2 *rec = j.records + j.rec_off[st32];
    const uint32_t rec_cap = (uint32_t)(j.rec_off[st32 + 1] - j.rec_off[st32]);
    uint32_t opos = 0;
    const uint32_t h0 = j.hist0 ? j.hist0[st32] : 0u;
    int64_t cur_len = (int64_t)in_len;
    bool bail = in_len > 0x0fffffffull;
    bool done = false;
    uint32_t win_bits = j.window_bits ? j.window_bits : 96u * 1024u * 8u;
    int64_t hdr_bit = 0; // bit position (relative to `in`) of the next block header
    bool ring_ok = false; // the ring holds the 32768 bytes in front of opos (or all there is)

    while (!bail && !done) {
      // ---- block header: warp 0 (the code of the warp kernel), results through shared memory ----
      if (warp == 0) {
        UBits ub;
        ub.init(in, (uint64_t)cur_len);
        ub.refill();
        ub.drop((int)(hdr_bit & 7)); // `in` was advanced to the byte the header starts in
        bool hb = false;
        ub.refill();
        const int final_flag = (int)ub.take(1);
        const int typ = (int)ub.take(2);
        int lit_only = 0, flat = 0;
        uint32_t sn = 0, sp = 0;
        int mn1 = 0, mx1 = 0, mn2 = 0, mx2 = 0;
        if (typ == 3) hb = true;
        else if (typ == 0) {
          const int64_t p = (ub.consumed_bits(in) + 7) >> 3;
          if (p + 4 > cur_len) hb = true;
          else {
            sn = (uint32_t)__ldg(in + p) | ((uint32_t)__ldg(in + p + 1) << 8);
            const uint32_t nn = (uint32_t)__ldg(in + p + 2) | ((uint32_t)__ldg(in + p + 3) << 8);
            if (nn != ((~sn) & 0xffffu) || p + 4 + sn > cur_len || (uint64_t)opos + sn > cap) hb = true;
            sp = (uint32_t)(p + 4);
          }
        } else {
          if (typ == 1) {
            for (int i = lane; i < 288; i += 32) sm.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
            sm.lens[288 + lane] = 5;
            __syncwarp();
            warp_canon(sm.lens, 288, sm.lsorted, &sm.tl, &mn1, &mx1);
            warp_canon(sm.lens + 288, 32, sm.dsorted, &sm.td, &mn2, &mx2);
          } else {
            ub.refill();
            const int nlit = (int)ub.take(5) + 257;
            const int ndist = (int)ub.take(5) + 1;
            const int nclen = (int)ub.take(4) + 4;
            if (nlit > kNumLit || ndist > kNumDist) hb = true;
            if (!hb) {
              lit_only = nlit == 257;
              if (lane < 19) sm.cl_lens[lane] = 0;
              __syncwarp();
              for (int i = 0; i < nclen; i++) {
                ub.refill();
                const uint32_t v = ub.take(3);
                if (lane == 0) sm.cl_lens[c3_code_order[i]] = (uint8_t)v;
              }
              __syncwarp();
              int mnc = 0, mxc = 0;
              if (!warp_canon(sm.cl_lens, 19, sm.dsorted, &sm.td, &mnc, &mxc)) hb = true;
              if (!hb) {
                warp_fill_lut(sm.cl, 7, sm.dsorted, &sm.td, mnc, mxc);
                const int n = nlit + ndist;
                int i = 0, prev = 0;
                while (i < n) {
                  ub.refill();
                  const uint32_t e = sm.cl[ub.peek() & 127];
                  const int len = (int)(e & 15), x = (int)(e >> 4);
                  if (len == 0) { hb = true; break; }
                  ub.drop(len);
                  if (x < 16) {
                    if (lane == 0) sm.lens[i] = (uint8_t)x;
                    prev = x;
                    i++;
                    continue;
                  }
                  int rep, b = 0;
                  if (x == 16) {
                    if (i == 0) { hb = true; break; }
                    b = prev;
                    rep = 3 + (int)ub.take(2);
                  } else if (x == 17) rep = 3 + (int)ub.take(3);
                  else rep = 11 + (int)ub.take(7);
                  if (i + rep > n) { hb = true; break; }
                  for (int k = lane; k < rep; k += 32) sm.lens[i + k] = (uint8_t)b;
                  i += rep;
                  prev = b;
                }
                __syncwarp();
              }
              if (!hb) {
                uint8_t dl = 0;
                if (lane < ndist) dl = sm.lens[nlit + lane];
                __syncwarp();
                sm.lens[288 + lane] = (lane < ndist) ? dl : 0;
                __syncwarp();
                if (sm.lens[kEob] == 0) hb = true;
                if (!hb && !warp_canon(sm.lens, nlit, sm.lsorted, &sm.tl, &mn1, &mx1)) hb = true;
                if (!hb && !warp_canon(sm.lens + 288, ndist, sm.dsorted, &sm.td, &mn2, &mx2)) hb = true;
              }
            }
          }
          if (!hb) {
            warp_fill_lut(sm.lit, kLB, sm.lsorted, &sm.tl, mn1, mx1);
            warp_fill_lut(sm.dist, kDB, sm.dsorted, &sm.td, mn2, mx2);
            const uint32_t c = lane < 16 ? sm.tl.count[lane] : 0u;
            const uint32_t mxc = __reduce_max_sync(kFull, c), tot = __reduce_add_sync(kFull, c);
            flat = mxc * 10u >= tot * 8u;
          }
        }
        if (lane == 0) {
          S.bail = hb ? 1 : 0;
          S.final_flag = final_flag;
          S.typ = typ;
          S.lit_only = lit_only;
          S.flat_code = flat;
          S.b0s = ub.consumed_bits(in);
          S.sn = sn;
          S.sp = sp;
        }
      }
      __syncthreads();
      CTA_T(0);
      if (S.bail) { bail = true; break; }
      const int final_flag = S.final_flag, typ = S.typ;
      const bool lit_only = S.lit_only != 0, flat_code = S.flat_code != 0;
      if (typ == 0) { // stored block (data_block / copy_data, inflate.mbt:708-766)
        const uint32_t sn = S.sn, sp = S.sp;
        for (uint32_t i = tid; i < sn; i += CL) out[opos + i] = __ldg(in + sp + i);
        opos += sn;
        ring_ok = false;
        cur_len -= (int64_t)sp + sn;
        in += sp + sn;
        hdr_bit = 0;
        __syncthreads();
        if (final_flag) done = true;
        continue;
      }
      // ---- block body: CL ranges, speculative starts, fixpoint over the hand-over positions ----
      const int64_t b0s = S.b0s;
      if (b0s >= cur_len * 8) { bail = true; break; }
      uint32_t b0 = (uint32_t)b0s;
      const uint32_t bend = (uint32_t)(cur_len * 8), blk_first_bit = b0;
      uint32_t eob = 0;
      for (bool have_eob = false; !have_eob && !bail;) {
        const uint32_t wend = (uint64_t)b0 + win_bits < bend ? b0 + win_bits : bend;
        uint32_t R = (wend - b0 + (uint32_t)CL - 1u) / (uint32_t)CL;
        if (R < kMinRange) R = kMinRange;
        const uint64_t s64 = (uint64_t)b0 + (uint64_t)tid * R;
        const uint32_t s_nom = s64 < wend ? (uint32_t)s64 : wend;
        const uint32_t e_i = (s64 + R < wend) ? (uint32_t)(s64 + R) : wend;
        uint32_t start = s_nom, p = 0, flag = P_OK, n_out = 0, n_rec = 0;
        if (!flat_code && e_i - s_nom > kSyncBits) start = e_i - kSyncBits;
        bool need = true;
        for (int round = 0; round < CL + 2; round++) {
          uint32_t tp, tf, to, tr = 0;
          if (lit_only) decode_ranges_lit<false>(sm, in, cur_len, bend, need, start, e_i, tp, tf, to, nullptr, 0u);
          else decode_ranges<false>(sm, S.len_tab, S.dist_tab, in, cur_len, bend, need, start, e_i, tp, tf, to, tr, nullptr, 0u, nullptr);
          if (need) { p = tp; flag = tf; n_out = to; n_rec = tr; }
          S.p[tid] = p;
          S.flag[tid] = flag;
          __syncthreads();
          const uint32_t pp = tid ? S.p[tid - 1] : 0u, pf = tid ? S.flag[tid - 1] : P_OK;
          const uint32_t ns = tid == 0 ? b0 : (pf == P_OK ? pp : start);
          need = ns != start;
          start = ns;
          if (!__syncthreads_or(need ? 1 : 0)) break;
        }
        CTA_T(1);
        // the chain is valid up to the first range that does not hand over: it must have met EOB
        if (tid == 0) S.first_bad = 0xffffffffu;
        __syncthreads();
        if (flag != P_OK) atomicMin(&S.first_bad, (uint32_t)tid);
        __syncthreads();
        const uint32_t fb = S.first_bad;
        uint32_t f = CL - 1;
        if (fb == 0xffffffffu) {
          if (wend >= bend) { bail = true; break; } // no EOB before the end of the input
        } else {
          f = fb;
          if (S.flag[f] != P_EOB) { bail = true; break; }
        }
        const bool mine = (uint32_t)tid <= f;
        // CTA-wide exclusive scan of the output bytes / records of the ranges
        uint32_t xo = mine ? n_out : 0u, xr = mine ? n_rec : 0u;
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t yo = __shfl_up_sync(kFull, xo, o), yr = __shfl_up_sync(kFull, xr, o);
          if (lane >= o) { xo += yo; xr += yr; }
        }
        if (lane == 31) { S.wo[warp] = xo; S.wr[warp] = xr; }
        __syncthreads();
        uint32_t base_o = 0, base_r = 0, tot_o = 0, tot_r = 0;
#pragma unroll
        for (int w = 0; w < CW; w++) {
          if (w < warp) { base_o += S.wo[w]; base_r += S.wr[w]; }
          tot_o += S.wo[w];
          tot_r += S.wr[w];
        }
        if (tot_o > cap - opos || tot_r > rec_cap) { bail = true; break; }
        // The output of the window is assembled in the shared-memory ring if it leaves room for the 32768 bytes of
        // history (every window of a 65535-byte block does): literals are written there, the copies are replayed
        // there, and the finished bytes leave for global memory in one coalesced sweep.  Otherwise in place.
        const bool use_ring = tot_o <= kRing - (uint32_t)kMaxMatchOffset;
        if (use_ring && !ring_ok) { // bring the history in (first block, or the previous one went another way)
          const int64_t avail = (int64_t)opos + (int64_t)h0;
          const uint32_t hist = avail < (int64_t)kMaxMatchOffset ? (uint32_t)avail : (uint32_t)kMaxMatchOffset;
          for (uint32_t i = tid; i < hist; i += CL) ring[(opos - hist + i) & kRingMask] = out[(int64_t)opos - (int64_t)hist + (int64_t)i];
          __syncthreads();
        }
        {
          uint32_t tp, tf, to, tr = 0;
          const uint32_t my_o = opos + base_o + xo - (mine ? n_out : 0u);
          uint2 *my_rec = rec + base_r + xr - (mine ? n_rec : 0u);
          if (use_ring) {
            if (lit_only) decode_ranges_lit<true, true>(sm, in, cur_len, bend, mine, start, e_i, tp, tf, to, ring, my_o);
            else decode_ranges<true, true>(sm, S.len_tab, S.dist_tab, in, cur_len, bend, mine, start, e_i, tp, tf, to, tr, ring, my_o, my_rec, h0);
          } else {
            if (lit_only) decode_ranges_lit<true>(sm, in, cur_len, bend, mine, start, e_i, tp, tf, to, out, my_o);
            else decode_ranges<true>(sm, S.len_tab, S.dist_tab, in, cur_len, bend, mine, start, e_i, tp, tf, to, tr, out, my_o, my_rec, h0);
          }
          const bool bad = mine && (tf == P_BAD || tp != p || to != n_out || tr != n_rec);
          if (!use_ring) __threadfence_block();
          if (__syncthreads_or(bad ? 1 : 0)) { bail = true; break; }
        }
        CTA_T(2);
        // the copies of this window, in order (everything in front of it is final)
        if (use_ring) {
          if (warp == 0) replay_records<true>(ring, rec, tot_r, lane);
          __syncthreads();
          CTA_T(3);
          const bool al = (reinterpret_cast<uintptr_t>(out + opos) & 3u) == 0u && (opos & 3u) == 0u;
          if (al) { // words where the two sides are aligned alike (the ring is indexed by position)
            const uint32_t nw = tot_o >> 2;
            const uint32_t *rw = reinterpret_cast<const uint32_t *>(ring);
            uint32_t *ow = reinterpret_cast<uint32_t *>(out + opos);
            for (uint32_t i = tid; i < nw; i += CL) ow[i] = rw[((opos >> 2) + i) & (kRingMask >> 2)];
            for (uint32_t i = (nw << 2) + tid; i < tot_o; i += CL) out[opos + i] = ring[(opos + i) & kRingMask];
          } else {
            for (uint32_t i = tid; i < tot_o; i += CL) out[opos + i] = ring[(opos + i) & kRingMask];
          }
          ring_ok = true;
        } else {
          if (warp == 0) replay_records<false>(out, rec, tot_r, lane);
          __threadfence_block();
          ring_ok = false;
        }
        __syncthreads();
        CTA_T(4);
        opos += tot_o;
        const uint32_t pend = S.p[f];
        if (fb == 0xffffffffu) {
          if (pend <= b0 || pend >= bend) { bail = true; break; }
          b0 = pend; // the block goes on behind the window, which doubles: the guess was too small
          if (win_bits < 0x40000000u) win_bits *= 2u;
        } else {
          eob = pend;
          have_eob = true;
        }
        __syncthreads(); // S.p / S.flag are rewritten by the next window
      }
      if (bail) break;
      {
        const uint64_t w = ((uint64_t)(eob - blk_first_bit) * 5u) >> 2;
        const uint32_t wmin = 24u * 1024u * 8u;
        win_bits = w < wmin ? wmin : (w > 0x7fffffffull ? 0x7fffffffu : (uint32_t)w);
        if (j.window_bits == 0xffffffffu) win_bits = 0xffffffffu;
      }
      if ((int64_t)eob > cur_len * 8) { bail = true; break; }
      in += eob >> 3;
      cur_len -= (int64_t)(eob >> 3);
      hdr_bit = (int64_t)(eob & 7u);
      if (final_flag) done = true;
      __syncthreads(); // the tables are rebuilt by warp 0
    }

    // bits consumed: the stream ends hdr_bit bits into `in` (0 after a stored block: byte aligned)
    if (!bail && ((int64_t)(in - in0) * 8 + hdr_bit > (int64_t)in_len * 8)) bail = true;
    if (tid == 0) {
      if (bail) {
        const uint32_t k = atomicAdd(&j.counters[2], 1u);
        j.fallback[k] = st32;
      } else {
        j.out_len[st32] = opos;
        j.status[st32] = FB200_ST_EOF;
        j.err_off[st32] = 0;
        if (j.consumed) j.consumed[st32] = (uint64_t)((int64_t)(in - in0) + ((hdr_bit + 7) >> 3));
      }
    }
    __syncthreads();
    CTA_T(5);
#ifdef FB_CTA_PROF
    if (tid == 0)
      for (int k = 0; k < 8; k++) { atomicAdd(&g_cta_prof[k], (unsigned long long)prof[k]); prof[k] = 0; }
#endif
    if (j.group_done) { // host-buffer call: publish finished output groups so their D2H copy can start
      __threadfence();
      if (tid == 0) {
        const uint32_t g = st32 / j.group_streams;
        const uint32_t first = g * j.group_streams;
        const uint32_t cnt = (uint32_t)(j.nstreams - first < j.group_streams ? j.nstreams - first : j.group_streams);
        if (atomicAdd(&j.group_done[g], 1u) + 1u == cnt) {
          __threadfence_system();
          j.group_flag[g] = 1u;
        }
      }
      __syncthreads();
    }
  }
}

constexpr size_t kCtaSmemBytes = ((sizeof(CtaShared) + 15) & ~(size_t)15) + kRing;

} // namespace par

// record area of stream i: out capacity / 3 (a match yields >= 3 bytes) + 4
__global__ void k_rec_off(const uint64_t *out_off, uint64_t *rec_off, uint64_t ns)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= ns) rec_off[i] = (out_off[i] - out_off[0]) / 3 + 4 * i;
}

void launch_rec_off(const uint64_t *out_off, uint64_t *rec_off, uint64_t ns, cudaStream_t st)
{
  k_rec_off<<<(unsigned)((ns + 1 + 255) / 256), 256, 0, st>>>(out_off, rec_off, ns);
}

// Stream order for the decode kernel: a counting sort on the compressed size (256-byte classes), largest
// first.  A warp waits for its slowest stream and runs the union of its lanes' paths, so streams of similar
// size (similar symbol counts, similar literal / match mix) belong in the same warp, and the long ones
// should start first.
constexpr int kOrderBuckets = 1024;

__device__ __forceinline__ int order_bucket(uint64_t clen)
{
  const uint64_t b = clen >> 8;
  return (int)(b < (uint64_t)kOrderBuckets - 1 ? b : (uint64_t)kOrderBuckets - 1);
}

__global__ void k_order_count(InflateJob j, uint32_t *hist)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < j.nstreams) atomicAdd(&hist[order_bucket(j.comp_off[i + 1] - j.comp_off[i])], 1u);
}

__global__ void __launch_bounds__(kOrderBuckets) k_order_scan(uint32_t *hist)
{
  // start of bucket b = number of streams in larger buckets (descending order)
  __shared__ uint32_t sh[kOrderBuckets];
  const int t = threadIdx.x;
  sh[t] = hist[kOrderBuckets - 1 - t];
  __syncthreads();
  for (int o = 1; o < kOrderBuckets; o <<= 1) {
    const uint32_t v = t >= o ? sh[t - o] : 0u;
    __syncthreads();
    sh[t] += v;
    __syncthreads();
  }
  hist[kOrderBuckets - 1 - t] = sh[t] - hist[kOrderBuckets - 1 - t]; // exclusive
}

__global__ void k_order_scatter(InflateJob j, uint32_t *hist, uint32_t *order)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < j.nstreams) order[atomicAdd(&hist[order_bucket(j.comp_off[i + 1] - j.comp_off[i])], 1u)] = (uint32_t)i;
}

// decode order of device-resident calls
void launch_stream_order(const InflateJob &j, uint32_t *order_hist, cudaStream_t st)
{
  const unsigned gs = (unsigned)((j.nstreams + 255) / 256);
  cudaMemsetAsync(order_hist, 0, kOrderBuckets * sizeof(uint32_t), st);
  k_order_count<<<gs, 256, 0, st>>>(j, order_hist);
  k_order_scan<<<1, kOrderBuckets, 0, st>>>(order_hist);
  k_order_scatter<<<gs, 256, 0, st>>>(j, order_hist, const_cast<uint32_t *>(j.order));
}

// process-wide launch configuration, read once (preload_inflate3_kernels: when the first context is created)
static long g_win_kb = -1;
static int g_minb = 6; // 24 warps per SM at 80 registers: more occupancy only buys spills

static void inflate3_config()
{
  static bool done = false;
  if (done) return;
  if (const char *e = getenv("FB200_INFLATE_WINDOW_KB")) g_win_kb = atol(e); // first-block window; 0 = no windows at all
  if (const char *e = getenv("FB200_INFLATE_CTAS")) {
    const int m = atoi(e);
    if (m == 4 || m == 5 || m == 6 || m == 8 || m == 12) g_minb = m;
  }
  done = true;
}

void launch_inflate3(const InflateJob &j_in, int num_sms, cudaStream_t st)
{
  if (j_in.nstreams == 0) return;
  InflateJob j = j_in;
  inflate3_config();
  if (g_win_kb == 0) j.window_bits = 0xffffffffu;
  else if (g_win_kb > 0) j.window_bits = (uint32_t)(g_win_kb * 8192);
  // few streams: a CTA per stream (32 * 8 ranges per block) instead of a warp per stream
  if (j.nstreams <= (uint64_t)j.cta_streams) {
    const unsigned g = (unsigned)j.nstreams;
    par::k_inflate_cta<<<g, par::CL, par::kCtaSmemBytes, st>>>(j);
    return;
  }
  const int minb = g_minb;
  const uint64_t want = (j.nstreams + par::kWarps - 1) / par::kWarps;
  const uint64_t maxg = (uint64_t)num_sms * minb;
  const unsigned g = (unsigned)(want < maxg ? want : maxg);
  if (minb == 4) par::k_inflate_par<4><<<g, par::kWarps * 32, 0, st>>>(j);
  else if (minb == 5) par::k_inflate_par<5><<<g, par::kWarps * 32, 0, st>>>(j);
  else if (minb == 6) par::k_inflate_par<6><<<g, par::kWarps * 32, 0, st>>>(j);
  else if (minb == 8) par::k_inflate_par<8><<<g, par::kWarps * 32, 0, st>>>(j);
  else par::k_inflate_par<12><<<g, par::kWarps * 32, 0, st>>>(j);
}

#ifdef FB_CTA_PROF
extern "C" void fb200_debug_cta_prof(unsigned long long *out, int reset)
{
  cudaMemcpyFromSymbol(out, par::g_cta_prof, 64);
  if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(par::g_cta_prof, z, 64); }
}
#endif

#ifdef FB_INFLATE_STEPSTAT
extern "C" void fb200_debug_stepstat(unsigned long long *out)
{
  cudaMemcpyFromSymbol(out, par::g_stepstat, 16);
}
#endif

void preload_inflate3_kernels()
{
  inflate3_config();
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, par::k_inflate_par<4>);
  cudaFuncGetAttributes(&a, par::k_inflate_par<5>);
  cudaFuncGetAttributes(&a, par::k_inflate_par<6>);
  cudaFuncGetAttributes(&a, par::k_inflate_par<8>);
  cudaFuncGetAttributes(&a, par::k_inflate_par<12>);
  cudaFuncSetAttribute(par::k_inflate_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)par::kCtaSmemBytes);
  cudaFuncGetAttributes(&a, par::k_inflate_cta);
  cudaFuncGetAttributes(&a, k_rec_off);
  cudaFuncGetAttributes(&a, k_order_count);
  cudaFuncGetAttributes(&a, k_order_scan);
  cudaFuncGetAttributes(&a, k_order_scatter);
}

} // namespace fb
