// inflate2.cu -- K6 fast path, two kernels (valid streams; anything unusual is
// handed to the exact kernel k_inflate in inflate.cu, which reproduces the
// reference's error behaviour).
//
//   k_inflate_decode  ONE THREAD PER STREAM walks the bit stream
//                     (next_block / read_huffman / huff_sym / read_literal,
//                     inflate.mbt:345-684).  The Huffman chain is inherently
//                     serial per stream, so a warp decodes 32 streams at once
//                     instead of one stream 32 times redundantly.  Literals go
//                     straight to their final position; a back-reference is only
//                     *recorded* (dst, len, dist), because its source bytes may
//                     not exist yet.  Decode tables live in shared memory,
//                     laid out [entry][lane] so that a lane stays in its own bank.
//   k_inflate_copy    one warp per stream replays the recorded copies in order
//                     (copy_history / DictDecoder::write_copy, inflate.mbt:689-704,
//                     dict-decoder.mbt:114-185).  Records whose source lies wholly
//                     before the first unresolved destination are independent and
//                     are copied in parallel, one per lane; long copies and stored
//                     blocks are copied by the whole warp.
#include "common.cuh"
#include "kernels.h"
#include "../../include/flate_b200.h"

#include <cstdlib>

namespace fb {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kLB = 9;  // lit/len direct table bits
constexpr int kDB = 5;  // distance direct table bits (also hosts the code-length code while a header is read)
constexpr uint32_t kLitLim = 256u << 4; // table entry = (symbol << 4) | code length; literal iff 0 < e < kLitLim
constexpr int kDecStreamsPerCta = 32;   // a CTA always owns the tables of 32 streams: 32 / S warps
constexpr uint32_t kStoredMark = 0xffffu;
constexpr int kLaneCopyMax = 24;        // longest copy a single lane performs in a parallel round

// S = streams (active lanes) per warp.  The tables cost 1760 bytes per stream whatever S is, so an SM holds
// 128 streams; a smaller S spreads them over more warps (S = 8: 16 warps per SM), which hides the
// dependent-issue latency of the per-stream chain and shrinks the union of divergent paths a warp runs.
template <int S>
struct DecSmem { // one warp
  uint16_t lit[1 << kLB][S];   // direct table: (symbol << 4) | length for codes of <= kLB bits
  uint16_t llong[288][S];      // symbols of the longer lit/len codes, canonical order
  uint16_t dist[1 << kDB][S];
  uint16_t dsorted[32][S];     // symbols of the longer distance (or code-length) codes, canonical order
  int16_t lbase[15 - kLB][S];  // per length L > kLB: index of its first symbol in llong - first code of L
  int16_t dbase[15 - kDB][S];
};

// Left-justified (15-bit) end of the code range of every length, two per register: a code longer than
// the direct table is located by comparing the next 15 stream bits against them -- registers only.
struct Lims {
  uint32_t p[8];
};

__constant__ uint32_t c2_len_tab[32] = {
    3,           4,           5,           6,           7,           8,           9,           10,
    11 | 1 << 16, 13 | 1 << 16, 15 | 1 << 16, 17 | 1 << 16, 19 | 2 << 16, 23 | 2 << 16, 27 | 2 << 16, 31 | 2 << 16,
    35 | 3 << 16, 43 | 3 << 16, 51 | 3 << 16, 59 | 3 << 16, 67 | 4 << 16, 83 | 4 << 16, 99 | 4 << 16, 115 | 4 << 16,
    131 | 5 << 16, 163 | 5 << 16, 195 | 5 << 16, 227 | 5 << 16, 258, 0, 0, 0};
__constant__ uint32_t c2_dist_tab[32] = {
    1,            2,            3,             4,             5 | 1 << 16,    7 | 1 << 16,    9 | 2 << 16,     13 | 2 << 16,
    17 | 3 << 16, 25 | 3 << 16, 33 | 4 << 16,  49 | 4 << 16,  65 | 5 << 16,   97 | 5 << 16,   129 | 6 << 16,   193 | 6 << 16,
    257 | 7 << 16, 385 | 7 << 16, 513 | 8 << 16, 769 | 8 << 16, 1025 | 9 << 16, 1537 | 9 << 16, 2049 | 10 << 16, 3073 | 10 << 16,
    4097 | 11 << 16, 6145 | 11 << 16, 8193 | 12 << 16, 12289 | 12 << 16, 16385 | 13 << 16, 24577 | 13 << 16, 0, 0};
__constant__ uint8_t c2_code_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// per-thread LSB-first bit reader.  A 64-bit window (hi:lo) with a bit offset bo < 32 always exposes
// >= 33 valid bits, enough for any code together with its extra bits (15 + 13), so a code and its extra
// bits cost one peek() (a funnel shift) and one drop().  Input arrives in aligned 16-byte blocks: `cur`
// is consumed word by word while `nxt` (requested when `cur` was first touched, from a line that an L2
// prefetch asked for 512 bytes earlier) is still in flight, so taking a word never waits for memory and
// the loop-carried chain of a stream is table lookups and ALU only.
struct TBits {
  uint32_t lo, hi;
  int bo;
  const uint8_t *abase; // 16-byte aligned address at or below the first byte
  uint32_t pos;         // byte offset of `cur` from abase
  uint32_t lim;         // bytes from abase to the end of the input
  uint32_t lead;        // bytes between abase and the first byte
  int k;                // next word of `cur`
  uint4 cur, nxt;
  __device__ __forceinline__ uint4 fetch(uint32_t off) const
  {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (off < lim) v = __ldg(reinterpret_cast<const uint4 *>(abase + off));
    return v;
  }
  __device__ __forceinline__ uint32_t next_word()
  {
    const uint32_t a = (k & 1) ? cur.y : cur.x, b = (k & 1) ? cur.w : cur.z;
    const uint32_t v = (k & 2) ? b : a;
    k++;
    if (k == 4) {
      cur = nxt;
      k = 0;
      pos += 16;
      nxt = fetch(pos + 16);
      if (pos + 512 < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(abase + pos + 512));
    }
    return v;
  }
  __device__ __forceinline__ void init(const uint8_t *in, uint64_t len)
  {
    const uintptr_t a = reinterpret_cast<uintptr_t>(in);
    lead = (uint32_t)(a & 15);
    abase = in - lead;
    lim = len ? (uint32_t)len + lead : 0u;
    pos = 0;
    cur = fetch(0);
    nxt = fetch(16);
    k = (int)(lead >> 2);
    lo = next_word();
    hi = next_word();
    bo = (int)(a & 3) * 8; // bytes of the first word that precede the stream
  }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, bo); }
  __device__ __forceinline__ void drop(int n)
  {
    bo += n;
    if (bo >= 32) {
      lo = hi;
      hi = next_word();
      bo -= 32;
    }
  }
  __device__ __forceinline__ uint32_t take(int n)
  {
    const uint32_t v = peek() & ((1u << n) - 1u);
    drop(n);
    return v;
  }
  __device__ __forceinline__ void refill() {}
  // bits consumed so far, relative to the first byte
  __device__ __forceinline__ int64_t consumed_bits() const
  {
    return ((int64_t)pos - (int64_t)lead) * 8 + 32 * k - 64 + bo;
  }
};

// HuffmanDecoder::initialize (inflate.mbt:100-223) for one thread: canonical description + direct table.
// false = the reference rejects the code, or it is empty: the exact kernel deals with it.
// HuffmanDecoder::initialize (inflate.mbt:100-223) for one thread: direct table for codes of <= LB bits,
// canonical description (limits in registers, bases + symbols in shared memory) for the longer ones.
// false = the reference rejects the code, or it is empty: the exact kernel deals with it.
template <int LB, int S>
__device__ __noinline__ bool build_table(const uint8_t *lens, int nsym, uint16_t (*lut)[S], int16_t (*base)[S],
                                         uint16_t (*sorted)[S], Lims &lims, int lane)
{
  uint16_t cnt[16], nxt[16], first[16], rank0[16];
#pragma unroll
  for (int l = 0; l < 16; l++) cnt[l] = 0;
  for (int i = 0; i < nsym; i++) cnt[lens[i]]++;
  uint32_t code = 0, off = 0, nshort = 0;
  int mx = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) lims.p[k] = 0;
#pragma unroll
  for (int L = 1; L <= 15; L++) {
    const uint32_t c = cnt[L];
    code <<= 1;
    first[L] = (uint16_t)code;
    nxt[L] = (uint16_t)off;
    rank0[L] = (uint16_t)off;
    if (L <= LB) nshort += c;
    else base[L - LB - 1][lane] = (int16_t)((int)off - (int)nshort - (int)code);
    code += c;
    off += c;
    if (c) mx = L;
    lims.p[L >> 1] |= (code << (15 - L)) << (16 * (L & 1)); // <= 1 << 15: fits 16 bits
  }
  if (off == 0) return false;
  if (code != (1u << 15) && !(mx == 1 && off == 1)) return false; // complete, or the single 1-bit code (:161)
  constexpr int nlut = 1 << LB;
  for (int k = 0; k < nlut; k++) lut[k][lane] = 0;
  for (int i = 0; i < nsym; i++) {
    const int l = lens[i];
    if (!l) continue;
    const uint32_t pos = nxt[l]++;
    if (l > LB) sorted[pos - nshort][lane] = (uint16_t)i;
    else {
      const uint32_t c = first[l] + (pos - rank0[l]);
      const uint32_t r = __brev(c) >> (32 - l);
      const uint16_t e = (uint16_t)((i << 4) | l);
      for (uint32_t k = r; k < (uint32_t)nlut; k += 1u << l) lut[k][lane] = e;
    }
  }
  return true;
}

// code longer than the direct table: (sym << 4) | len, 0 if no code matches.  The limits are nondecreasing
// in L, so the length is LB + 1 + the number of limits the next 15 stream bits have reached.
template <int LB, int S>
__device__ __forceinline__ uint32_t decode_long(uint32_t bits, const Lims &lims, int16_t (*base)[S],
                                                uint16_t (*sorted)[S], int lane)
{
  const uint32_t v15 = __brev(bits) >> 17;
  int L = LB + 1;
#pragma unroll
  for (int q = LB + 1; q <= 15; q++) L += (v15 >= ((lims.p[q >> 1] >> (16 * (q & 1))) & 0xffffu));
  if (L > 15) return 0;
  const int idx = (int)(v15 >> (15 - L)) + (int)base[L - LB - 1][lane];
  return ((uint32_t)sorted[idx][lane] << 4) | (uint32_t)L;
}

// Lanes decode different streams, so every long-running loop below is written warp-synchronously
// (`while (__any_sync(...)) { if (mine) one step }`): a lane that takes a different branch re-joins the
// others at the next iteration instead of running the rest of the loop on its own.
enum { DS_HDR = 0, DS_SYMS = 1, DS_DONE = 2 };

template <int S>
__global__ void __launch_bounds__(kDecStreamsPerCta / S * 32) k_inflate_decode(InflateJob j)
{
  extern __shared__ __align__(16) uint8_t dec_smem[];
  DecSmem<S> &sm = reinterpret_cast<DecSmem<S> *>(dec_smem)[threadIdx.x >> 5];
  __shared__ uint32_t s_len_tab[32], s_dist_tab[32];
  if (threadIdx.x < 32) {
    s_len_tab[threadIdx.x] = c2_len_tab[threadIdx.x];
    s_dist_tab[threadIdx.x] = c2_dist_tab[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint8_t lens[320];
  Lims llims, dlims;

  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&j.counters[0], (uint32_t)S);
    base = __shfl_sync(kFull, base, 0);
    if (base >= j.nstreams) break;
    // streams are taken in the order k_order_* produced: largest compressed size first, similar sizes together
    const bool have = lane < S && base + (uint32_t)lane < j.nstreams;
    const uint32_t st = have ? j.order[base + (uint32_t)lane] : 0u;
    const uint32_t sti = st;
    const uint8_t *in0 = j.comp + j.comp_off[sti];
    const uint8_t *in = in0;
    const uint64_t in_len = j.comp_off[sti + 1] - j.comp_off[sti];
    uint8_t *out = j.out + j.out_off[sti];
    const uint64_t cap64 = j.out_off[sti + 1] - j.out_off[sti];
    const uint32_t cap = cap64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)cap64;
    uint2 *rec = j.records + j.rec_off[sti];
    const uint32_t rec_cap = (uint32_t)(j.rec_off[sti + 1] - j.rec_off[sti]);
    uint32_t nrec = 0;
    int64_t cur_len = (int64_t)in_len;
    TBits tb;
    tb.init(in, have ? in_len : 0);
    uint32_t opos = 0;
    bool bail = in_len > 0x0fffffffull;
    int final_flag = 0;
    int state = (have && !bail) ? DS_HDR : DS_DONE;

    while (__any_sync(kFull, state != DS_DONE)) {
      // ---------------- block headers (next_block, inflate.mbt:345-379) ----------------
      int nlit = 0, ndist = 0, n = 0;
      bool dyn = false, fixed = false;
      if (state == DS_HDR) {
        tb.refill();
        final_flag = (int)tb.take(1);
        const int typ = (int)tb.take(2);
        if (typ == 3) { bail = true; state = DS_DONE; }
        else if (typ == 0) { // stored block (data_block / copy_data, :708-766)
          const int64_t p = (tb.consumed_bits() + 7) >> 3;
          bool ok = p + 4 <= cur_len;
          uint32_t sn = 0;
          if (ok) {
            sn = (uint32_t)__ldg(in + p) | ((uint32_t)__ldg(in + p + 1) << 8);
            const uint32_t nn = (uint32_t)__ldg(in + p + 2) | ((uint32_t)__ldg(in + p + 3) << 8);
            ok = nn == ((~sn) & 0xffffu) && p + 4 + sn <= cur_len && (uint64_t)opos + sn <= cap;
          }
          const uint32_t sp = (uint32_t)(p + 4);
          if (ok && sn >= 16) { // the warp copies it later; the source offset is parked in the destination bytes
            if (nrec >= rec_cap) ok = false;
            else {
              const uint64_t so = (uint64_t)(in - in0) + sp;
              for (int k = 0; k < 8; k++) out[opos + k] = (uint8_t)(so >> (8 * k));
              rec[nrec++] = make_uint2(opos, sn | (kStoredMark << 16));
            }
          } else if (ok) {
            for (uint32_t i = 0; i < sn; i++) out[opos + i] = __ldg(in + sp + i);
          }
          if (!ok) { bail = true; state = DS_DONE; }
          else {
            opos += sn;
            cur_len -= (int64_t)sp + sn; // re-base the bit reader at the byte after the payload
            in += sp + sn;
            tb.init(in, (uint64_t)cur_len);
            if (final_flag) state = DS_DONE;
          }
        } else if (typ == 1) {
          fixed = true;
        } else { // read_huffman (:429-466)
          tb.refill();
          nlit = (int)tb.take(5) + 257;
          ndist = (int)tb.take(5) + 1;
          const int nclen = (int)tb.take(4) + 4;
          if (nlit > kNumLit || ndist > kNumDist) { bail = true; state = DS_DONE; }
          else {
            uint8_t cl[19];
            for (int i = 0; i < 19; i++) cl[i] = 0;
            for (int i = 0; i < 19; i++) {
              if (i < nclen) {
                tb.refill();
                cl[c2_code_order[i]] = (uint8_t)tb.take(3);
              }
            }
            // the code-length code borrows the distance tables
            if (!build_table<kDB, S>(cl, 19, sm.dist, sm.dbase, sm.dsorted, dlims, lane)) { bail = true; state = DS_DONE; }
            else { dyn = true; n = nlit + ndist; }
          }
        }
      }
      __syncwarp();
      // code lengths of the dynamic blocks (:473-530), one code per step and lane
      {
        int i = 0;
        bool herr = false;
        while (__any_sync(kFull, dyn && !herr && i < n)) {
          if (dyn && !herr && i < n) {
            tb.refill();
            uint32_t e = sm.dist[tb.peek() & ((1u << kDB) - 1u)][lane];
            if (e == 0) e = decode_long<kDB, S>(tb.peek(), dlims, sm.dbase, sm.dsorted, lane);
            if (e == 0) herr = true;
            else {
              tb.drop((int)(e & 15u));
              const int x = (int)(e >> 4);
              if (x < 16) lens[i++] = (uint8_t)x;
              else {
                int rep, b = 0;
                if (x == 16) {
                  b = i ? lens[i - 1] : 0;
                  if (i == 0) herr = true;
                  rep = 3 + (int)tb.take(2);
                } else if (x == 17) rep = 3 + (int)tb.take(3);
                else rep = 11 + (int)tb.take(7);
                if (i + rep > n) herr = true;
                else for (int k = 0; k < rep; k++) lens[i++] = (uint8_t)b;
              }
            }
          }
        }
        if (dyn && (herr || lens[kEob] == 0)) { bail = true; state = DS_DONE; dyn = false; }
      }
      if (fixed) { // fixed_huffman_decoder (:886-939); distances are 5-bit codes
        for (int i = 0; i < 288; i++) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        for (int i = 0; i < 32; i++) lens[288 + i] = 5;
        nlit = 288; ndist = 32;
      }
      if (dyn || fixed) {
        if (!build_table<kLB, S>(lens, nlit, sm.lit, sm.lbase, sm.llong, llims, lane) ||
            !build_table<kDB, S>(lens + nlit, ndist, sm.dist, sm.dbase, sm.dsorted, dlims, lane)) {
          bail = true; state = DS_DONE;
        } else state = DS_SYMS;
      }
      __syncwarp();

      // ---------------- symbols (read_literal, :565-684): one symbol per step and lane.  The step is
      // straight-line code for every lane; the two optional parts (a code longer than the direct table,
      // a length/distance pair) are entered by the whole warp when any lane needs them ----------------
      for (int it = 0; it < 2048 && __any_sync(kFull, state == DS_SYMS); it++) {
        const bool act = state == DS_SYMS;
        const uint32_t bits = tb.peek();
        uint32_t e = act ? sm.lit[bits & ((1u << kLB) - 1u)][lane] : 1u;
        if (__any_sync(kFull, e == 0)) {
          if (e == 0) e = decode_long<kLB, S>(bits, llims, sm.lbase, sm.llong, lane);
        }
        const uint32_t sym = e >> 4;
        const bool is_lit = act && e != 0 && sym < 256u;
        const bool is_len = act && sym > 256u && sym < (uint32_t)kNumLit;
        uint32_t length = 0;
        if (act) {
          // a length code (<= 15 bits) and its extra bits (<= 5) come out of the same peek
          const uint32_t cl = e & 15u;
          const uint32_t lt = is_len ? s_len_tab[sym - 257u] : 0u;
          length = (lt & 0xffffu) + ((bits >> cl) & ((1u << (lt >> 16)) - 1u));
          tb.drop((int)(cl + (lt >> 16)));
          if (e == 0 || sym >= (uint32_t)kNumLit || (is_lit && opos >= cap)) { bail = true; state = DS_DONE; }
          else if (is_lit) out[opos++] = (uint8_t)sym;
          else if (sym == (uint32_t)kEob) {
            // bits consumed beyond the real input mean the stream is truncated: exact path
            if (tb.consumed_bits() > cur_len * 8) { bail = true; state = DS_DONE; }
            else state = final_flag ? DS_DONE : DS_HDR;
          }
        }
        if (__any_sync(kFull, is_len)) {
          if (is_len) {
            // distance code (<= 15 bits) and its extra bits (<= 13) come out of one 32-bit peek
            const uint32_t dbits = tb.peek();
            uint32_t d = sm.dist[dbits & ((1u << kDB) - 1u)][lane];
            if (d == 0) d = decode_long<kDB, S>(dbits, dlims, sm.dbase, sm.dsorted, lane);
            if (d == 0 || (d >> 4) >= (uint32_t)kNumDist) { bail = true; state = DS_DONE; }
            else {
              const uint32_t dl = d & 15u;
              const uint32_t dt = s_dist_tab[d >> 4];
              const uint32_t dist = (dt & 0xffffu) + ((dbits >> dl) & ((1u << (dt >> 16)) - 1u));
              tb.drop((int)(dl + (dt >> 16)));
              if (dist > opos || length > cap - opos || nrec >= rec_cap) { bail = true; state = DS_DONE; }
              else {
                rec[nrec++] = make_uint2(opos, length | ((dist - 1u) << 16));
                opos += length;
              }
            }
          }
        }
      }
      __syncwarp();
    }

    if (have) {
      if (!bail && tb.consumed_bits() > cur_len * 8) bail = true;
      if (bail) {
        const uint32_t k = atomicAdd(&j.counters[2], 1u);
        j.fallback[k] = st;
        j.nrec[st] = 0;
      } else {
        const int64_t cb = tb.consumed_bits();
        j.nrec[st] = nrec;
        j.out_len[st] = opos;
        j.status[st] = FB200_ST_EOF;
        j.err_off[st] = 0;
        if (j.consumed) j.consumed[st] = (uint64_t)((int64_t)(in - in0) + ((cb + 7) >> 3));
      }
    }
    __syncwarp();
  }
}

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

// ------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_inflate_copy(InflateJob j)
{
  const int lane = threadIdx.x & 31;
  const unsigned ltm = (1u << lane) - 1u;
  for (;;) {
    uint32_t st = 0;
    if (lane == 0) st = atomicAdd(&j.counters[3], 1u);
    st = __shfl_sync(kFull, st, 0);
    if (st >= j.nstreams) break;
    const uint32_t nrec = j.nrec[st];
    if (nrec == 0) continue;
    uint8_t *out = j.out + j.out_off[st];
    const uint8_t *in0 = j.comp + j.comp_off[st];
    const uint2 *rec = j.records + j.rec_off[st];
    for (uint32_t g = 0; g < nrec; g += 32) {
      const uint32_t r = g + (uint32_t)lane;
      uint32_t dst = 0, len = 0, dm1 = 0;
      if (r < nrec) {
        const uint2 v = rec[r];
        dst = v.x; len = v.y & 0xffffu; dm1 = v.y >> 16;
      }
      const bool big = (len > (uint32_t)kLaneCopyMax) || (dm1 == kStoredMark);
      // source interval of a lane-sized copy: [dst - dist, dst - dist + min(len, dist))
      const uint32_t dist = dm1 + 1u;
      const uint32_t src_end = dst - dist + (len < dist ? len : dist);
      unsigned pending = __ballot_sync(kFull, r < nrec);
      while (pending) {
        const int p = __ffs(pending) - 1;
        const uint32_t dstp = __shfl_sync(kFull, dst, p);
        const bool bigp = __shfl_sync(kFull, (int)big, p) != 0;
        if (bigp) { // the whole warp copies record p
          const uint32_t lenp = __shfl_sync(kFull, len, p);
          const uint32_t dm1p = __shfl_sync(kFull, dm1, p);
          uint8_t *dp = out + dstp;
          if (dm1p == kStoredMark) {
            uint64_t so = 0;
            for (int k = 0; k < 8; k++) so |= (uint64_t)__ldcg(dp + k) << (8 * k);
            __syncwarp();
            const uint8_t *sp = in0 + so;
            for (uint32_t i = lane; i < lenp; i += 32) dp[i] = __ldg(sp + i);
          } else {
            const uint32_t dd = dm1p + 1u;
            const uint8_t *sp = dp - dd;
            if (dd >= 32u) {
              for (uint32_t b = 0; b < lenp; b += 32) {
                const uint32_t i = b + lane;
                uint8_t v = 0;
                if (i < lenp) v = __ldcg(sp + i);
                if (i < lenp) dp[i] = v;
                if (dd < lenp) __syncwarp(); // later chunks may read what this one wrote
              }
            } else { // overlapping: the pattern of dd bytes repeats (dict-decoder.mbt:136-149)
              for (uint32_t i = lane; i < lenp; i += 32) dp[i] = __ldcg(sp + (i % dd));
            }
          }
          pending &= ~(1u << p);
          __syncwarp();
          continue;
        }
        // parallel round: every pending small record whose source lies before the first unresolved destination
        const bool mine = ((pending >> lane) & 1u) && !big && (lane == p || src_end <= dstp);
        // stop at the first pending big record: records after it may depend on it
        const unsigned bigm = __ballot_sync(kFull, big) & pending;
        const unsigned ready = __ballot_sync(kFull, mine) & (bigm ? ((1u << (__ffs(bigm) - 1)) - 1u) : kFull);
        if ((ready >> lane) & 1u) {
          uint8_t *dp = out + dst;
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
        pending &= ~ready;
        __syncwarp();
      }
      (void)ltm;
    }
  }
}

// ------------------------------------------------------------------
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

// decode order only (used by the other fast paths as well)
void launch_stream_order(const InflateJob &j, uint32_t *order_hist, cudaStream_t st)
{
  const unsigned gs = (unsigned)((j.nstreams + 255) / 256);
  cudaMemsetAsync(order_hist, 0, kOrderBuckets * sizeof(uint32_t), st);
  k_order_count<<<gs, 256, 0, st>>>(j, order_hist);
  k_order_scan<<<1, kOrderBuckets, 0, st>>>(order_hist);
  k_order_scatter<<<gs, 256, 0, st>>>(j, order_hist, const_cast<uint32_t *>(j.order));
}

template <int S>
static void launch_decode(const InflateJob &j, int num_sms, cudaStream_t st)
{
  static bool inited[64] = {}; // function attributes are per device
  const int smem = (kDecStreamsPerCta / S) * (int)sizeof(DecSmem<S>);
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev >= 0 && dev < 64 ? dev : 0;
  if (!inited[dev]) {
    cudaFuncSetAttribute(k_inflate_decode<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    inited[dev] = true;
  }
  const uint64_t want = (j.nstreams + kDecStreamsPerCta - 1) / kDecStreamsPerCta;
  const uint64_t maxg = (uint64_t)num_sms * 4; // 4 CTAs (128 streams) per SM by shared memory
  k_inflate_decode<S><<<(unsigned)(want < maxg ? want : maxg), kDecStreamsPerCta / S * 32, smem, st>>>(j);
}

void launch_inflate2(const InflateJob &j, int num_sms, uint32_t *order_hist, cudaStream_t st)
{
  if (j.nstreams == 0) return;
  static int spw = 0;
  if (!spw) {
    const char *e = getenv("FB200_DECODE_SPW");
    spw = e ? atoi(e) : 8;
    if (spw != 8 && spw != 16 && spw != 32) spw = 8;
  }
  launch_stream_order(j, order_hist, st);
  if (spw == 8) launch_decode<8>(j, num_sms, st);
  else if (spw == 16) launch_decode<16>(j, num_sms, st);
  else launch_decode<32>(j, num_sms, st);
  const uint64_t wantc = (j.nstreams + 3) / 4;
  const uint64_t maxc = (uint64_t)num_sms * 16;
  k_inflate_copy<<<(unsigned)(wantc < maxc ? wantc : maxc), 128, 0, st>>>(j);
}

// CUDA loads kernels lazily, and loading one while another kernel spins on a host-fed watermark can
// deadlock: every kernel of this file is loaded when the context is created.
void preload_inflate2_kernels()
{
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_inflate_decode<8>);
  cudaFuncGetAttributes(&a, k_inflate_decode<16>);
  cudaFuncGetAttributes(&a, k_inflate_decode<32>);
  cudaFuncGetAttributes(&a, k_inflate_copy);
  cudaFuncGetAttributes(&a, k_rec_off);
  cudaFuncGetAttributes(&a, k_order_count);
  cudaFuncGetAttributes(&a, k_order_scan);
  cudaFuncGetAttributes(&a, k_order_scatter);
}

} // namespace fb
