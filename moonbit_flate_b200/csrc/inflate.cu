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
// The exact decoder.  Valid streams are decoded by the fast kernel (inflate3.cu); whatever it puts on the
// fallback list -- a header the reference rejects, an invalid symbol, a distance beyond the output, a slot that
// is too small, truncated input -- is re-decoded here from scratch with the reference's exact error behaviour.

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

void launch_inflate_exact(const InflateJob &j, int num_sms, cudaStream_t st)
{
  if (j.nstreams == 0) return;
  k_inflate<<<(unsigned)num_sms * 2, kInflateWarps * 32, 0, st>>>(j);
}

void preload_inflate_kernels()
{
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_inflate);
}

} // namespace fb
