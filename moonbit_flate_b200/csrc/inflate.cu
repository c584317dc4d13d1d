// inflate.cu -- K6: batched inflate, one warp per stream.
//
// Follows the Decompressor state machine of the reference
// (inflate.mbt:345-854) and DictDecoder's copy semantics
// (dict-decoder.mbt:114-185), restructured for a warp:
//   * lane 0 owns the LSB-first bit reader and decodes symbols; decode tables
//     live in shared memory (per warp): a direct lookup table for short codes
//     plus canonical first-code / count arrays for the long ones.  The symbol
//     a bit pattern decodes to is independent of the table layout, so results
//     match the reference's 9-bit chunks + links tables (inflate.mbt:100-223).
//   * the output goes straight to the final buffer (no 32 KiB window, no
//     suspension); back-references are copied by all 32 lanes, with
//     out[i] = out[i - dist] forward-copy semantics for overlapping copies.
//   * Error behaviour is the reference's: the same conditions map to corrupt /
//     unexpected EOF / plain eof, and the "corrupt input before offset N"
//     offset is reproduced by modelling its lazy byte-at-a-time refill:
//     roffset = ceil(max over requests of (bit position + bits required) / 8)
//     (more_bits inflate.mbt:789-799, huff_sym :803-854 with h.min, the
//     h1.min = max(min, len(EOB)) tweak :542-544).
#include "common.cuh"
#include "kernels.h"
#include "../../include/flate_b200.h"

#include <cstdlib>

namespace fb {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kLitLutBits = 10;
constexpr int kDistLutBits = 8;
constexpr int kInflateWarps = 4;

struct HuffTab {          // canonical description of one code
  uint16_t first[16];     // first code (MSB-first value) of each length
  uint16_t count[16];     // codes per length
  uint16_t offs[16];      // index of the first symbol of each length in sorted[]
};

struct WarpSmem {
  uint16_t lit_lut[1 << kLitLutBits];   // (sym << 4) | len, 0 = long code or invalid
  uint16_t dist_lut[1 << kDistLutBits]; // also hosts the code-length code while a header is read
  uint16_t lit_sorted[288];
  uint16_t dist_sorted[32];
  HuffTab lit, dist;
  uint8_t lens[288 + 32];
  uint8_t cl_lens[32];
};

// Build decode tables for lens[0..nsym) (HuffmanDecoder::initialize,
// inflate.mbt:100-223).  Returns false when the reference rejects the code
// (neither complete nor the single 1-bit code, :161).  *minlen = h.min.
__device__ bool warp_build(const uint8_t *lens, int nsym, int lut_bits, uint16_t *lut, uint16_t *sorted,
                           HuffTab *tab, int *minlen)
{
  const int lane = lane_id();
  // lane L counts the symbols of length L (L = 1..15)
  int c = 0;
  if (lane >= 1 && lane <= 15)
    for (int i = 0; i < nsym; i++) c += (lens[i] == lane);
  const unsigned nz = __ballot_sync(kFull, c != 0);
  const int lut_n = 1 << lut_bits;
  if (nz == 0) { // empty tree (:143-145): every lookup fails later
    for (int i = lane; i < lut_n; i += 32) lut[i] = 0;
    if (lane < 16) { tab->first[lane] = 0; tab->count[lane] = 0; tab->offs[lane] = 0; }
    *minlen = 0;
    __syncwarp();
    return true;
  }
  const int mn = __ffs(nz) - 1, mx = 31 - __clz(nz);
  int code = 0, off = 0, code_at_max = 0, my_first = 0, my_off = 0;
  for (int L = 1; L <= 15; L++) { // :148-154
    const int cL = __shfl_sync(kFull, c, L);
    code <<= 1;
    if (lane == L) { my_first = code; my_off = off; }
    code += cL;
    off += cL;
    if (L == mx) code_at_max = code;
  }
  if (code_at_max != (1 << mx) && !(code_at_max == 1 && mx == 1)) return false; // :161
  if (lane < 16) {
    tab->first[lane] = (uint16_t)my_first;
    tab->count[lane] = (uint16_t)c;
    tab->offs[lane] = (uint16_t)my_off;
  }
  if (c) { // symbols of my length in increasing order
    int k = my_off;
    for (int i = 0; i < nsym; i++)
      if (lens[i] == lane) sorted[k++] = (uint16_t)i;
  }
  __syncwarp();
  for (int idx = lane; idx < lut_n; idx += 32) {
    const unsigned r = __brev((unsigned)idx);
    uint16_t e = 0;
    for (int L = mn; L <= lut_bits && L <= mx; L++) {
      const unsigned d = (r >> (32 - L)) - tab->first[L];
      if (d < tab->count[L]) {
        e = (uint16_t)((sorted[tab->offs[L] + d] << 4) | L);
        break;
      }
    }
    lut[idx] = e;
  }
  *minlen = mn;
  __syncwarp();
  return true;
}

// lane-0 bit reader over in[0..in_len)
struct BitReader {
  const uint8_t *in;
  uint64_t in_len;
  uint64_t ipos;    // next byte to load into bb
  uint64_t bb;
  int nbb;
  uint64_t req_max; // furthest bit any request of the reference's lazy reader has needed

  __device__ __forceinline__ uint64_t bitpos() const { return ipos * 8 - (uint64_t)nbb; }
  __device__ __forceinline__ uint64_t remaining() const { return (in_len - ipos) * 8 + (uint64_t)nbb; }
  __device__ __forceinline__ void refill()
  {
    while (nbb <= 56 && ipos < in_len) {
      bb |= (uint64_t)__ldg(in + ipos) << nbb;
      ipos++;
      nbb += 8;
    }
  }
  __device__ __forceinline__ void consume(int n)
  {
    bb >>= n;
    nbb -= n;
  }
  __device__ __forceinline__ void request(int n)
  {
    const uint64_t e = bitpos() + (uint64_t)n;
    if (e > req_max) req_max = e;
  }
  // bytes the reference's reader has consumed so far (Decompressor.roffset)
  __device__ __forceinline__ uint64_t roffset() const { return (req_max + 7) >> 3; }
};

// huff_sym (inflate.mbt:803-854) against one table.  Returns the symbol, or
// -1 with *st set (FB200_ST_UNEXPECTED_EOF / FB200_ST_CORRUPT).
__device__ __forceinline__ int decode_sym(BitReader &br, const uint16_t *lut, int lut_bits, const HuffTab *tab,
                                          const uint16_t *sorted, int hmin, int *st)
{
  br.refill();
  const uint64_t R = br.remaining();
  if (R < (uint64_t)hmin) { // cannot even gather h.min bits: no_eof(eof) (:818-826)
    br.req_max = br.in_len * 8;
    *st = FB200_ST_UNEXPECTED_EOF;
    return -1;
  }
  const uint32_t bits = (uint32_t)br.bb; // bits past the end of input read as 0, as in the reference
  const uint32_t e = lut[bits & ((1u << lut_bits) - 1)];
  int n = (int)(e & 15), sym = (int)(e >> 4);
  if (n == 0) {
    const unsigned r = __brev(bits);
    for (int L = lut_bits + 1; L <= 15; L++) {
      const unsigned d = (r >> (32 - L)) - tab->first[L];
      if (d < tab->count[L]) {
        sym = sorted[tab->offs[L] + d];
        n = L;
        break;
      }
    }
  }
  if (n == 0) { // no code matches (empty / degenerate tree): corrupt (:842-847)
    br.request(hmin);
    *st = FB200_ST_CORRUPT;
    return -1;
  }
  if ((uint64_t)n > R) { // the code needs bits the input does not have
    br.req_max = br.in_len * 8;
    *st = FB200_ST_UNEXPECTED_EOF;
    return -1;
  }
  br.request(n > hmin ? n : hmin);
  br.consume(n);
  return sym;
}

// more_bits loop (`while self.nb < k { more_bits }`): returns false when the
// input is exhausted -- the reference then reports plain eof (quirk D5).
__device__ __forceinline__ bool need_bits(BitReader &br, int k, int *st)
{
  br.refill();
  if (br.remaining() < (uint64_t)k) {
    br.req_max = br.in_len * 8;
    *st = FB200_ST_EOF_AT_REFILL;
    return false;
  }
  br.request(k);
  return true;
}


// ------------------------------------------------------------------
// Fast path.  Valid streams (everything a deflate encoder produced) are decoded
// here with a lean lane-0 loop: 32-bit table entries that already carry the
// literal byte / length base / distance base and extra-bit counts, aligned
// word refills, no per-symbol end-of-input bookkeeping.  Anything unusual --
// a header the reference rejects, an invalid symbol, a distance beyond the
// output, an output slot that is too small, bits consumed past the end of the
// input -- makes the warp put the stream on the fallback list; k_inflate then
// re-decodes it from scratch with the reference's exact error behaviour.
// A stream that completes here consumed only real bits and every block ended
// in its EOB, so the reference decodes it to the same bytes with status EOF
// and roffset = ceil(bits consumed / 8) (argument in DESIGN.md).

constexpr int kFastLitBits = 10;
constexpr int kFastDistBits = 8;
constexpr int kFastWarps = 4;

// Table entries are (symbol << 4) | code length, 0 = not in the table (long code
// or invalid).  A lit/len entry is a literal iff 0 < e < (256 << 4).
constexpr uint32_t kLitLimit = 256u << 4;

struct FastSmem {
  uint16_t lit_lut[1 << kFastLitBits];
  uint16_t dist_lut[1 << kFastDistBits];
  uint16_t cl_lut[128];
  uint16_t lit_sorted[288];
  uint16_t dist_sorted[32];
  HuffTab lit, dist;
  uint8_t lens[288 + 32];
  uint8_t cl_lens[32];
};

// ceil(65536 / d): (i * r) >> 16 == i / d for i <= 258, d in [1, 31]
__constant__ uint32_t c_recip[32] = {0,    65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554,
                                     5958, 5462,  5042,  4682,  4370,  4096,  3856,  3641, 3450, 3277, 3121,
                                     2979, 2850,  2731,  2622,  2521,  2428,  2341,  2260, 2185, 2115};
// length symbol 257+c -> base | extra bits << 16 (inflate.mbt:591-615); distance symbol -> same (:656-670)
__constant__ uint32_t c_len_tab[32] = {
    3,           4,           5,           6,           7,           8,           9,           10,
    11 | 1 << 16, 13 | 1 << 16, 15 | 1 << 16, 17 | 1 << 16, 19 | 2 << 16, 23 | 2 << 16, 27 | 2 << 16, 31 | 2 << 16,
    35 | 3 << 16, 43 | 3 << 16, 51 | 3 << 16, 59 | 3 << 16, 67 | 4 << 16, 83 | 4 << 16, 99 | 4 << 16, 115 | 4 << 16,
    131 | 5 << 16, 163 | 5 << 16, 195 | 5 << 16, 227 | 5 << 16, 258, 0, 0, 0};
__constant__ uint32_t c_dist_tab[32] = {
    1,            2,            3,             4,             5 | 1 << 16,    7 | 1 << 16,    9 | 2 << 16,     13 | 2 << 16,
    17 | 3 << 16, 25 | 3 << 16, 33 | 4 << 16,  49 | 4 << 16,  65 | 5 << 16,   97 | 5 << 16,   129 | 6 << 16,   193 | 6 << 16,
    257 | 7 << 16, 385 | 7 << 16, 513 | 8 << 16, 769 | 8 << 16, 1025 | 9 << 16, 1537 | 9 << 16, 2049 | 10 << 16, 3073 | 10 << 16,
    4097 | 11 << 16, 6145 | 11 << 16, 8193 | 12 << 16, 12289 | 12 << 16, 16385 | 13 << 16, 24577 | 13 << 16, 0, 0};
__constant__ uint8_t c_code_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// canonical description only (counts, first codes, sorted symbols); false = reference rejects the code
// (or it is an empty tree, which the exact kernel handles)
__device__ bool warp_canon(const uint8_t *lens, int nsym, uint16_t *sorted, HuffTab *tab, int *mn_out, int *mx_out)
{
  const int lane = lane_id();
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
  // complete, or the single 1-bit code the reference also accepts (inflate.mbt:161); anything else: exact path
  if (code_at_max != (1 << mx) && !(code_at_max == 1 && mx == 1)) return false;
  if (lane < 16) {
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

__device__ void warp_fill_lut(uint16_t *lut, int lut_bits, const uint16_t *sorted, const HuffTab *tab, int mn, int mx)
{
  const int lane = lane_id();
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

// Warp-uniform bit reader: every lane holds the same state and issues the same
// (broadcast) loads, so the decode loop runs without divergence.  Aligned 32-bit
// refills; bits past the end of the input read as zero.
struct FastBits {
  const uint32_t *w;   // next aligned word to load
  const uint8_t *end;  // one past the last input byte
  uint64_t bb;
  int nb;
  __device__ __forceinline__ void init(const uint8_t *in, uint64_t len)
  {
    end = in + len;
    bb = 0; nb = 0;
    const uint8_t *p = in;
    while (reinterpret_cast<uintptr_t>(p) & 3) { // bytes up to the first aligned word (phantom zeros past the end)
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
  // bits consumed so far, relative to `in` (phantom words included)
  __device__ __forceinline__ int64_t consumed_bits(const uint8_t *in) const
  {
    return (int64_t)(reinterpret_cast<const uint8_t *>(w) - in) * 8 - nb;
  }
};

// symbol for a code longer than the table width: (sym << 4) | len, 0 if none matches
__device__ __forceinline__ uint32_t canon_long(uint32_t bits, int from, const HuffTab *tab, const uint16_t *sorted)
{
  const unsigned r = __brev(bits);
  for (int L = from; L <= 15; L++) {
    const unsigned d = (r >> (32 - L)) - tab->first[L];
    if (d < tab->count[L]) return ((uint32_t)sorted[tab->offs[L] + d] << 4) | (uint32_t)L;
  }
  return 0;
}

__global__ void __launch_bounds__(kFastWarps * 32, 10) k_inflate_fast(InflateJob j)
{
  __shared__ FastSmem smem_all[kFastWarps];
  FastSmem &sm = smem_all[threadIdx.x >> 5];
  const int lane = lane_id();

  for (;;) {
    uint32_t st32 = 0;
    if (lane == 0) st32 = atomicAdd(&j.counters[0], 1u);
    st32 = __shfl_sync(kFull, st32, 0);
    if (st32 >= j.nstreams) break;
    if (j.avail) { // host-buffer call: wait until the H2D stream has delivered this stream's bytes
      if (lane == 0)
        while (*(volatile const uint32_t *)j.avail <= st32) __nanosleep(500);
      __syncwarp();
    }

    const uint8_t *in = j.comp + j.comp_off[st32];
    const uint64_t in_len = j.comp_off[st32 + 1] - j.comp_off[st32];
    uint8_t *out = j.out + j.out_off[st32];
    const uint64_t cap64 = j.out_off[st32 + 1] - j.out_off[st32];
    const uint32_t cap = cap64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)cap64;
    int64_t cur_len = (int64_t)in_len; // bytes from `in` (re-based after stored blocks) to the end
    FastBits fb;
    fb.init(in, in_len);
    uint32_t opos = 0;
    const uint32_t h0 = j.hist0 ? j.hist0[st32] : 0u; // preset dictionary in front of the slot
    bool bail = in_len > 0x0fffffffull; // keep bit counts comfortably inside 32/64-bit ranges
    bool done = false;
    // deferred back-reference copy: byte loaded at one match, stored at the next (hides the L2 round trip)
    uint8_t *pend_ptr = nullptr;
    uint32_t pend_val = 0;

    while (!bail && !done) {
      // ---- block header (all lanes, uniform) ----
      fb.refill();
      const int final_flag = (int)fb.take(1);
      const int typ = (int)fb.take(2);
      if (typ == 3) { bail = true; break; }

      if (typ == 0) { // stored block
        const int64_t p = (fb.consumed_bits(in) + 7) >> 3;
        if (p + 4 > cur_len) { bail = true; break; }
        const uint32_t sn = (uint32_t)__ldg(in + p) | ((uint32_t)__ldg(in + p + 1) << 8);
        const uint32_t nn = (uint32_t)__ldg(in + p + 2) | ((uint32_t)__ldg(in + p + 3) << 8);
        if (nn != ((~sn) & 0xffffu) || p + 4 + sn > cur_len || (uint64_t)opos + sn > cap) { bail = true; break; }
        const uint32_t sp = (uint32_t)(p + 4);
        for (uint32_t i = lane; i < sn; i += 32) out[opos + i] = __ldg(in + sp + i);
        opos += sn;
        cur_len -= (int64_t)sp + sn; // re-base the bit reader at the byte after the payload
        in += sp + sn;
        fb.init(in, (uint64_t)cur_len);
        __syncwarp();
        if (final_flag) done = true;
        continue;
      }

      int mn1 = 0, mx1 = 0, mn2 = 0, mx2 = 0;
      if (typ == 1) {
        for (int i = lane; i < 288; i += 32) sm.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        sm.lens[288 + lane] = 5;
        __syncwarp();
        warp_canon(sm.lens, 288, sm.lit_sorted, &sm.lit, &mn1, &mx1);
        warp_canon(sm.lens + 288, 32, sm.dist_sorted, &sm.dist, &mn2, &mx2);
      } else {
        fb.refill();
        const int nlit = (int)fb.take(5) + 257;
        const int ndist = (int)fb.take(5) + 1;
        const int nclen = (int)fb.take(4) + 4;
        if (nlit > kNumLit || ndist > kNumDist) { bail = true; break; }
        if (lane < 19) sm.cl_lens[lane] = 0;
        __syncwarp();
        for (int i = 0; i < nclen; i++) {
          fb.refill();
          const uint32_t v = fb.take(3);
          if (lane == 0) sm.cl_lens[c_code_order[i]] = (uint8_t)v;
        }
        __syncwarp();
        int mnc = 0, mxc = 0;
        if (!warp_canon(sm.cl_lens, 19, sm.dist_sorted, &sm.dist, &mnc, &mxc)) { bail = true; break; }
        warp_fill_lut(sm.cl_lut, 7, sm.dist_sorted, &sm.dist, mnc, mxc);
        // code lengths: uniform decode, lane 0 writes
        bool herr = false;
        const int n = nlit + ndist;
        int i = 0, prev = 0;
        while (i < n) {
          fb.refill();
          const uint32_t e = sm.cl_lut[fb.peek() & 127];
          const int len = (int)(e & 15), x = (int)(e >> 4);
          if (len == 0) { herr = true; break; }
          fb.drop(len);
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
            rep = 3 + (int)fb.take(2);
          } else if (x == 17) rep = 3 + (int)fb.take(3);
          else rep = 11 + (int)fb.take(7);
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
        if (!warp_canon(sm.lens, nlit, sm.lit_sorted, &sm.lit, &mn1, &mx1)) { bail = true; break; }
        if (!warp_canon(sm.lens + 288, ndist, sm.dist_sorted, &sm.dist, &mn2, &mx2)) { bail = true; break; }
      }
      warp_fill_lut(sm.lit_lut, kFastLitBits, sm.lit_sorted, &sm.lit, mn1, mx1);
      warp_fill_lut(sm.dist_lut, kFastDistBits, sm.dist_sorted, &sm.dist, mn2, mx2);

      // ---- symbols: uniform decode; lane 0 stores literals, all lanes copy matches ----
      for (;;) {
        fb.refill(); // >= 32 bits: room for two lit/len codes
        uint32_t e = sm.lit_lut[fb.peek() & ((1u << kFastLitBits) - 1u)];
        if (e - 1u < kLitLimit - 1u) { // literal
          if (opos >= cap) { bail = true; break; } // slot full: the exact kernel reports it
          if (lane == 0) out[opos] = (uint8_t)(e >> 4);
          opos++;
          fb.drop((int)(e & 15u));
          e = sm.lit_lut[fb.peek() & ((1u << kFastLitBits) - 1u)];
          if (e - 1u < kLitLimit - 1u) { // second literal on the same refill
            if (opos >= cap) { bail = true; break; }
            if (lane == 0) out[opos] = (uint8_t)(e >> 4);
            opos++;
            fb.drop((int)(e & 15u));
            continue;
          }
        }
        if (e == 0) { // long code (or no code at all)
          e = canon_long(fb.peek(), kFastLitBits + 1, &sm.lit, sm.lit_sorted);
          if (e == 0) { bail = true; break; }
          if (e < kLitLimit) {
            if (opos >= cap) { bail = true; break; }
            if (lane == 0) out[opos] = (uint8_t)(e >> 4);
            opos++;
            fb.drop((int)(e & 15u));
            continue;
          }
        }
        fb.drop((int)(e & 15u));
        const uint32_t sym = e >> 4;
        if (sym == (uint32_t)kEob) break;
        if (sym >= (uint32_t)kNumLit) { bail = true; break; }
        // length base + extra bits (inflate.mbt:591-627), uniform index -> constant-cache broadcast
        const uint32_t lt = c_len_tab[sym - 257u];
        const uint32_t length = (lt & 0xffffu) + fb.take((int)(lt >> 16));
        fb.refill();
        uint32_t d = sm.dist_lut[fb.peek() & ((1u << kFastDistBits) - 1u)];
        if (d == 0) {
          d = canon_long(fb.peek(), kFastDistBits + 1, &sm.dist, sm.dist_sorted);
          if (d == 0) { bail = true; break; }
        }
        fb.drop((int)(d & 15u));
        if ((d >> 4) >= (uint32_t)kNumDist) { bail = true; break; }
        const uint32_t dt = c_dist_tab[d >> 4]; // distance base + extra bits (inflate.mbt:656-674)
        const uint32_t dist = (dt & 0xffffu) + fb.take((int)(dt >> 16));
        if (dist > opos + h0 || length > cap - opos) { bail = true; break; }
        // retire the previous deferred copy, then make lane-0 literal stores visible to the copy loads
        if (pend_ptr) { *pend_ptr = (uint8_t)pend_val; pend_ptr = nullptr; }
        __syncwarp();
        uint8_t *dp = out + opos;
        const uint8_t *sp8 = dp - dist;
        if (length <= 32u) { // one step: load now, store at the next match / block end
          if ((uint32_t)lane < length) {
            uint32_t i = (uint32_t)lane;
            if (dist < 32u) i -= ((i * c_recip[dist]) >> 16) * dist; // i % dist: the pattern repeats
            pend_val = sp8[i];
            pend_ptr = dp + lane;
          }
        } else if (dist >= 32u) {
          for (uint32_t base = 0; base < length; base += 32) {
            const uint32_t i = base + lane;
            if (i < length) dp[i] = sp8[i];
            __syncwarp();
          }
        } else {
          const uint32_t r = c_recip[dist];
          for (uint32_t i = lane; i < length; i += 32) dp[i] = sp8[i - ((i * r) >> 16) * dist];
          __syncwarp();
        }
        opos += length;
      }
      if (pend_ptr) { *pend_ptr = (uint8_t)pend_val; pend_ptr = nullptr; }
      __syncwarp();
      if (bail) break;
      if (final_flag) done = true;
      // bits consumed beyond the real input mean the stream is truncated: exact path
      if (fb.consumed_bits(in) > cur_len * 8) { bail = true; break; }
    }

    if (!bail && fb.consumed_bits(in) > cur_len * 8) bail = true;
    if (lane == 0) {
      if (bail) {
        const uint32_t k = atomicAdd(&j.counters[2], 1u);
        j.fallback[k] = st32;
      } else {
        const int64_t cb = fb.consumed_bits(in);
        j.out_len[st32] = opos;
        j.status[st32] = FB200_ST_EOF;
        j.err_off[st32] = 0;
        if (j.consumed) j.consumed[st32] = (uint64_t)((int64_t)(in - (j.comp + j.comp_off[st32])) + ((cb + 7) >> 3));
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

enum { EV_NONE = 0, EV_MATCH, EV_EOB, EV_ERR, EV_STORED, EV_TABLES, EV_FIXED };

__global__ void __launch_bounds__(kInflateWarps * 32) k_inflate(InflateJob j)
{
  __shared__ WarpSmem smem_all[kInflateWarps];
  WarpSmem &sm = smem_all[threadIdx.x >> 5];
  const int lane = lane_id();
  const uint8_t code_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

  const uint32_t nfallback = j.counters[2];
  for (;;) {
    uint32_t st32 = 0;
    if (lane == 0) st32 = atomicAdd(&j.counters[1], 1u);
    st32 = __shfl_sync(kFull, st32, 0);
    if (st32 >= nfallback) break;
    st32 = j.fallback[st32];
    const uint64_t h0 = j.hist0 ? j.hist0[st32] : 0u; // preset dictionary in front of the slot

    BitReader br;
    br.in = j.comp + j.comp_off[st32];
    br.in_len = j.comp_off[st32 + 1] - j.comp_off[st32];
    br.ipos = 0; br.bb = 0; br.nbb = 0; br.req_max = 0;
    uint8_t *out = j.out + j.out_off[st32];
    const uint64_t cap = j.out_off[st32 + 1] - j.out_off[st32];
    uint64_t opos = 0;
    int status = -1;
    int64_t err_off = 0;

    while (status < 0) {
      // ---------------- next_block (:345-379) ----------------
      int ev = EV_NONE, final_flag = 0;
      int nlit = 0, ndist = 0;
      uint64_t sp = 0; uint32_t sn = 0, savail = 0; // stored block: payload byte position, LEN, bytes present
      int hmin_cl = 0;
      if (lane == 0) {
        int est = -1;
        if (!need_bits(br, 3, &est)) { status = est; ev = EV_ERR; }
        else {
          final_flag = (int)(br.bb & 1);
          const int typ = (int)((br.bb >> 1) & 3);
          br.consume(3);
          if (typ == 0) { // data_block (:708-737)
            const uint64_t p = br.roffset(); // bits of the current byte are discarded
            if (p + 4 > br.in_len) {
              br.req_max = br.in_len * 8;
              status = FB200_ST_UNEXPECTED_EOF; ev = EV_ERR;
            } else {
              const uint32_t n = (uint32_t)__ldg(br.in + p) | ((uint32_t)__ldg(br.in + p + 1) << 8);
              const uint32_t nn = (uint32_t)__ldg(br.in + p + 2) | ((uint32_t)__ldg(br.in + p + 3) << 8);
              br.req_max = (p + 4) * 8;
              if (nn != ((~n) & 0xffffu)) {
                status = FB200_ST_CORRUPT; err_off = (int64_t)(p + 4); ev = EV_ERR;
              } else {
                sp = p + 4; sn = n;
                const uint64_t left = br.in_len - sp;
                savail = (uint32_t)(left < n ? left : n);
                ev = EV_STORED;
              }
            }
          } else if (typ == 1) {
            ev = EV_FIXED;
          } else if (typ == 2) { // read_huffman (:429-466): counts + code-length code lengths
            if (!need_bits(br, 14, &est)) { status = est; ev = EV_ERR; }
            else {
              nlit = (int)(br.bb & 0x1f) + 257;
              ndist = (int)((br.bb >> 5) & 0x1f) + 1;
              const int nclen = (int)((br.bb >> 10) & 0xf) + 4;
              if (nlit > kNumLit || ndist > kNumDist) {
                // the reference tests nlit before consuming, ndist after 5 bits; roffset is the same
                status = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); ev = EV_ERR;
              } else {
                br.consume(14);
                ev = EV_TABLES;
                for (int i = 0; i < 19; i++) sm.cl_lens[i] = 0;
                for (int i = 0; i < nclen; i++) {
                  if (!need_bits(br, 3, &est)) { status = est; ev = EV_ERR; break; }
                  sm.cl_lens[code_order[i]] = (uint8_t)(br.bb & 7);
                  br.consume(3);
                }
              }
            }
          } else {
            status = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); ev = EV_ERR;
          }
        }
      }
      ev = __shfl_sync(kFull, ev, 0);
      final_flag = __shfl_sync(kFull, final_flag, 0);
      __syncwarp();

      if (ev == EV_ERR) { status = __shfl_sync(kFull, status, 0); break; }

      if (ev == EV_STORED) { // copy_data (:742-766)
        sp = __shfl_sync(kFull, sp, 0);
        sn = __shfl_sync(kFull, sn, 0);
        savail = __shfl_sync(kFull, savail, 0);
        uint32_t ncopy = savail;
        int st_after = -1;
        if (opos + ncopy > cap) { ncopy = (uint32_t)(cap - opos); st_after = FB200_ST_DST_TOO_SMALL; }
        else if (savail < sn) st_after = FB200_ST_UNEXPECTED_EOF;
        for (uint32_t i = lane; i < ncopy; i += 32) out[opos + i] = __ldg(br.in + sp + i);
        opos += ncopy;
        if (lane == 0) {
          br.ipos = sp + savail; br.bb = 0; br.nbb = 0; br.req_max = br.ipos * 8;
        }
        __syncwarp();
        if (st_after >= 0) { status = st_after; break; }
        if (final_flag) { status = FB200_ST_EOF; break; } // finish_block (:769-777)
        continue;
      }

      int hmin_lit = 0, hmin_dist = 0;
      bool fixed = false;
      if (ev == EV_FIXED) { // fixed_huffman_decoder (:886-939), min = 7; distances are 5 reversed bits (:633-641)
        for (int i = lane; i < 288; i += 32) sm.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        for (int i = lane; i < 32; i += 32) sm.lens[288 + i] = 5;
        __syncwarp();
        warp_build(sm.lens, 288, kLitLutBits, sm.lit_lut, sm.lit_sorted, &sm.lit, &hmin_lit);
        warp_build(sm.lens + 288, 32, kDistLutBits, sm.dist_lut, sm.dist_sorted, &sm.dist, &hmin_dist);
        hmin_lit = 7;
        fixed = true;
      } else { // EV_TABLES: rest of read_huffman (:467-547)
        nlit = __shfl_sync(kFull, nlit, 0);
        ndist = __shfl_sync(kFull, ndist, 0);
        const bool okcl = warp_build(sm.cl_lens, 19, 7, sm.dist_lut, sm.dist_sorted, &sm.dist, &hmin_cl);
        int herr = -1;
        if (!okcl) {
          if (lane == 0) { herr = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); }
        } else if (lane == 0) {
          const int n = nlit + ndist;
          int i = 0;
          while (i < n) {
            int est = -1;
            const int x = decode_sym(br, sm.dist_lut, 7, &sm.dist, sm.dist_sorted, hmin_cl, &est);
            if (x < 0) { herr = est; err_off = (int64_t)br.roffset(); break; }
            if (x < 16) { sm.lens[i++] = (uint8_t)x; continue; }
            int rep, nb, b;
            if (x == 16) {
              rep = 3; nb = 2;
              if (i == 0) { herr = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); break; }
              b = sm.lens[i - 1];
            } else if (x == 17) { rep = 3; nb = 3; b = 0; }
            else { rep = 11; nb = 7; b = 0; }
            if (!need_bits(br, nb, &est)) { herr = est; break; }
            rep += (int)(br.bb & ((1u << nb) - 1));
            br.consume(nb);
            if (i + rep > n) { herr = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); break; }
            for (int k = 0; k < rep; k++) sm.lens[i++] = (uint8_t)b;
          }
        }
        herr = __shfl_sync(kFull, herr, 0);
        __syncwarp();
        if (herr >= 0) { status = herr; break; }
        // distance lengths sit right after the literal lengths; move them to their slot
        uint8_t dl = 0;
        if (lane < ndist) dl = sm.lens[nlit + lane];
        __syncwarp();
        if (lane < 32) sm.lens[288 + lane] = (lane < ndist) ? dl : 0;
        __syncwarp();
        const int eob_len = sm.lens[kEob];
        const bool ok1 = warp_build(sm.lens, nlit, kLitLutBits, sm.lit_lut, sm.lit_sorted, &sm.lit, &hmin_lit);
        const bool ok2 = ok1 && warp_build(sm.lens + 288, ndist, kDistLutBits, sm.dist_lut, sm.dist_sorted, &sm.dist, &hmin_dist);
        if (!ok1 || !ok2) { // :533-536
          if (lane == 0) err_off = (int64_t)br.roffset();
          status = FB200_ST_CORRUPT;
          break;
        }
        if (hmin_lit < eob_len) hmin_lit = eob_len; // :542-544
      }

      // ---------------- huffman_block: read_literal / copy_history (:565-704) ----------------
      for (;;) {
        int bev = EV_NONE, length = 0, dist = 0;
        if (lane == 0) {
          for (;;) {
            int est = -1;
            const int v = decode_sym(br, sm.lit_lut, kLitLutBits, &sm.lit, sm.lit_sorted, hmin_lit, &est);
            if (v < 0) { status = est; err_off = (int64_t)br.roffset(); bev = EV_ERR; break; }
            if (v < 256) {
              if (opos >= cap) { status = FB200_ST_DST_TOO_SMALL; bev = EV_ERR; break; }
              out[opos++] = (uint8_t)v;
              continue;
            }
            if (v == 256) { bev = EV_EOB; break; }
            int n;
            if (v < 265) { length = v - (257 - 3); n = 0; }
            else if (v < 269) { length = v * 2 - (265 * 2 - 11); n = 1; }
            else if (v < 273) { length = v * 4 - (269 * 4 - 19); n = 2; }
            else if (v < 277) { length = v * 8 - (273 * 8 - 35); n = 3; }
            else if (v < 281) { length = v * 16 - (277 * 16 - 67); n = 4; }
            else if (v < 285) { length = v * 32 - (281 * 32 - 131); n = 5; }
            else if (v < kNumLit) { length = 258; n = 0; }
            else { status = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); bev = EV_ERR; break; }
            if (n > 0) {
              if (!need_bits(br, n, &est)) { status = est; bev = EV_ERR; break; }
              length += (int)(br.bb & ((1u << n) - 1));
              br.consume(n);
            }
            if (fixed) { // 5 bits, bit-reversed (:633-641): a more_bits loop, so exhaustion is plain eof
              if (!need_bits(br, 5, &est)) { status = est; bev = EV_ERR; break; }
              dist = (int)(__brev((unsigned)(br.bb & 0x1f)) >> 27);
              br.consume(5);
            } else {
              dist = decode_sym(br, sm.dist_lut, kDistLutBits, &sm.dist, sm.dist_sorted, hmin_dist, &est);
              if (dist < 0) { status = est; err_off = (int64_t)br.roffset(); bev = EV_ERR; break; }
            }
            if (dist < 4) dist++;
            else if (dist < kNumDist) {
              const int nb = (dist - 2) >> 1;
              int extra = (dist & 1) << nb;
              if (!need_bits(br, nb, &est)) { status = est; bev = EV_ERR; break; }
              extra |= (int)(br.bb & ((1u << nb) - 1));
              br.consume(nb);
              dist = (1 << (nb + 1)) + 1 + extra;
            } else { status = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); bev = EV_ERR; break; }
            // dist > hist_size (:677): hist_size = min(dictionary + bytes produced, 32768), dist <= 32768
            if ((uint64_t)dist > opos + h0) { status = FB200_ST_CORRUPT; err_off = (int64_t)br.roffset(); bev = EV_ERR; break; }
            bev = EV_MATCH;
            break;
          }
        }
        bev = __shfl_sync(kFull, bev, 0);
        opos = __shfl_sync(kFull, opos, 0);
        if (bev == EV_MATCH) {
          length = __shfl_sync(kFull, length, 0);
          dist = __shfl_sync(kFull, dist, 0);
          int ncopy = length;
          bool too_small = false;
          if (opos + (uint64_t)ncopy > cap) { ncopy = (int)(cap - opos); too_small = true; }
          __syncwarp();
          uint8_t *dp = out + opos;
          const uint8_t *sp8 = dp - dist;
          if (dist >= 32) {
            for (int base = 0; base < ncopy; base += 32) {
              const int i = base + lane;
              if (i < ncopy) dp[i] = sp8[i];
              __syncwarp();
            }
          } else { // overlapping: the pattern of `dist` bytes repeats (dict-decoder.mbt:136-149)
            for (int i = lane; i < ncopy; i += 32) dp[i] = sp8[i % dist];
            __syncwarp();
          }
          opos += (uint64_t)ncopy;
          if (too_small) { status = FB200_ST_DST_TOO_SMALL; break; }
          continue;
        }
        if (bev == EV_ERR) status = __shfl_sync(kFull, status, 0);
        break; // EOB or error
      }
      if (status >= 0) break;
      if (final_flag) { status = FB200_ST_EOF; break; }
    }

    if (lane == 0) {
      j.out_len[st32] = opos;
      j.status[st32] = status;
      j.err_off[st32] = (status == FB200_ST_CORRUPT) ? err_off : 0;
      if (j.consumed) j.consumed[st32] = br.roffset();
    }
    __syncwarp();
  }
}

void launch_inflate(const InflateJob &j, int num_sms, bool fast_v1, cudaStream_t st)
{
  if (j.nstreams == 0) return;
  if (!fast_v1) {
    k_inflate<<<(unsigned)num_sms * 2, kInflateWarps * 32, 0, st>>>(j);
    return;
  }
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    const char *e = getenv("FB200_INFLATE_CTAS");
    ctas_per_sm = e ? atoi(e) : 12;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  uint64_t want = (j.nstreams + kFastWarps - 1) / kFastWarps;
  uint64_t maxg = (uint64_t)num_sms * ctas_per_sm;
  unsigned g = (unsigned)(want < maxg ? want : maxg);
  k_inflate_fast<<<g, kFastWarps * 32, 0, st>>>(j);
  // exact re-decode of whatever the fast path put on the fallback list (usually nothing)
  k_inflate<<<(unsigned)num_sms * 2, kInflateWarps * 32, 0, st>>>(j);
}

// CUDA loads kernels lazily, and loading one while another kernel spins on a host-fed watermark can
// deadlock: every kernel of this file is loaded when the context is created.
void preload_inflate_kernels()
{
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_inflate_fast);
  cudaFuncGetAttributes(&a, k_inflate);
}

} // namespace fb
