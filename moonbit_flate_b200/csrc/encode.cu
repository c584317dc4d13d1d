// encode.cu -- K2 histogram, K3 code construction + header, layout, K4 bit pack.
//
//   K2  k_histogram   HuffmanBitWriter::index_tokens counting loop / histogram
//                     (huffman-bit-writer.mbt:550-572, :831-836) and the block
//                     policy of Compressor::enc_speed (deflate.mbt:236-277)
//   K3  k_build_codes HuffmanEncoder::generate / bit_counts /
//                     assign_encoding_and_size (huffman-code.mbt:112-343),
//                     generate_codegen, dynamic_size, write_dynamic_header
//                     (huffman-bit-writer.mbt:241-360, :421-471)
//       k_layout      where each block / stream lands in the output
//   K4  k_pack        write_tokens, write_block_huff body, write_stored_header
//                     + write_bytes (huffman-bit-writer.mbt:596-824, :474-487,
//                     :202-225); k_trailer = Compressor::close's final empty
//                     stored block (deflate.mbt:171-176)
#include "common.cuh"
#include "huff_build.cuh"
#include "kernels.h"

namespace fb {

constexpr unsigned kFull = 0xffffffffu;

struct BlockRef {
  uint32_t st;
  uint32_t n;       // bytes in this block
  uint64_t src_off; // absolute source offset of the block
};

__device__ __forceinline__ BlockRef block_ref(const DeflateJob &j, uint64_t blk)
{
  BlockRef r;
  r.st = j.blk_stream[blk];
  const uint64_t o0 = j.stream_off[r.st];
  const uint64_t L = j.stream_off[r.st + 1] - o0;
  const uint64_t boff = (blk - j.stream_blk0[r.st]) * (uint64_t)kBlockSize;
  const uint64_t rem = L - boff;
  r.n = (uint32_t)(rem < (uint64_t)kBlockSize ? rem : (uint64_t)kBlockSize);
  r.src_off = o0 + boff;
  return r;
}

// ------------------------------------------------------------------
// K2: one CTA per block
__global__ void __launch_bounds__(256) k_histogram(DeflateJob j)
{
  __shared__ uint32_t hist[kFreqStride];
  const uint64_t blk = j.blk_begin + blockIdx.x;
  const BlockRef r = block_ref(j, blk);
  const uint32_t n = r.n;
  const uint32_t ntok = j.blk_ntok[blk];
  int kind;
  if (j.cont_prev && blk == 0) kind = kKindSkip;   // stand-in for the part of the stream earlier calls compressed
  else if (n <= 16) kind = kKindStored;            // deflate.mbt:248-249
  else if (n < 128) kind = kKindHuff;              // :250-252
  else kind = (ntok > n - (n >> 4)) ? kKindHuff : kKindDynamic; // :266
  for (int i = threadIdx.x; i < kFreqStride; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  if (kind == kKindDynamic) {
    const uint32_t *tok = j.tokens + r.src_off;
    for (uint32_t i = threadIdx.x; i < ntok; i += blockDim.x) {
      const uint32_t t = __ldcs(tok + i); // streaming: read once here, once more by K4
      if (t < kMatchType) {
        atomicAdd(&hist[t], 1u);
      } else {
        int lc, nb, oc;
        uint32_t ex;
        length_code_of((t - kMatchType) >> kLengthShift, lc, nb, ex);
        offset_code_of(t & kOffsetMask, oc, nb, ex);
        atomicAdd(&hist[kLenCodesStart + lc], 1u);
        atomicAdd(&hist[kNumLit + oc], 1u);
      }
    }
  } else if (kind == kKindHuff) {
    const uint8_t *src = j.src + r.src_off;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[__ldg(src + i)], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    hist[kEob] = 1; // EOB: pushed token 256 (hbw:507) / literal_freq[256] = 1 (hbw:754)
    j.blk_kind[blk] = (uint8_t)kind;
  }
  __syncthreads();
  uint32_t *out = j.blk_freq + blk * kFreqStride;
  for (int i = threadIdx.x; i < kFreqStride; i += blockDim.x) out[i] = hist[i];
}

void launch_histogram(const DeflateJob &j, cudaStream_t st)
{
  if (j.blk_end <= j.blk_begin) return;
  k_histogram<<<(unsigned)(j.blk_end - j.blk_begin), 256, 0, st>>>(j);
}

// ------------------------------------------------------------------
// K3: code construction.  One warp per block, working set in shared memory (huff_build.cuh).  Two passes: blocks
// whose literal/length alphabet has at most 128 symbols in use (small blocks, text) are built with the small
// scratch at 32 warps per SM; the pass marks the others (blk_hdr_nbits = kNeedsBigScratch) for the second pass
// with the full-size scratch at 18 warps per SM.
constexpr int kBuildWarpsBig = 18;   // 12.5 KB of scratch per warp
constexpr int kBuildWarpsSmall = 32; // 6.9 KB
constexpr int kSmallSyms = 128;
constexpr uint32_t kNeedsBigScratch = 0xffffffffu;

template <int NS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_build_codes(DeflateJob j, uint32_t *work_counter)
{
  extern __shared__ __align__(16) uint8_t build_smem[];
  using SC = HuffScratchT<NS>;
  SC &S = reinterpret_cast<SC *>(build_smem)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  for (;;) { // blocks differ a lot in cost (alphabet size): hand them out one at a time
    uint32_t b32 = 0;
    if (lane == 0) b32 = atomicAdd(work_counter, 1u);
    b32 = __shfl_sync(kFull, b32, 0);
    if (j.blk_begin + b32 >= j.blk_end) break;
    const uint64_t blk = j.blk_begin + b32;
    const int kind = j.blk_kind[blk];
    if (kind == kKindStored || kind == kKindSkip) continue;
    const uint32_t *gfreq = j.blk_freq + blk * kFreqStride;
    if (NS < kMaxSyms) { // small pass: count the literal/length symbols in use
      int used = 0;
      for (int i = lane; i < kNumLit; i += 32) used += gfreq[i] != 0;
      used = __reduce_add_sync(kFull, used);
      if (used > NS) {
        if (lane == 0) j.blk_hdr_nbits[blk] = kNeedsBigScratch;
        continue;
      }
    } else if (j.blk_hdr_nbits[blk] != kNeedsBigScratch) {
      continue; // built by the small pass
    }
    const BlockRef r = block_ref(j, blk);
    const BlockBuild res = build_block_warp(gfreq, kind, r.n, j.blk_code + blk * kFreqStride, j.blk_hdr + blk * kHdrWords, S);
    if (lane == 0) {
      j.blk_kind[blk] = (uint8_t)res.kind;
      j.blk_hdr_nbits[blk] = res.hdr_nbits;
      j.blk_bits[blk] = res.blk_bits;
    }
    __syncwarp();
  }
}

void launch_build_codes(const DeflateJob &j, int num_sms, cudaStream_t st)
{
  if (j.blk_end <= j.blk_begin) return;
  // enough CTAs to fill the GPU when the kernel has it to itself (the blocks are handed out dynamically)
  const uint64_t nb = j.blk_end - j.blk_begin;
  {
    const uint64_t want = (nb + kBuildWarpsSmall - 1) / kBuildWarpsSmall;
    const unsigned g = (unsigned)(want < (uint64_t)num_sms ? want : (uint64_t)num_sms);
    k_build_codes<kSmallSyms, kBuildWarpsSmall>
        <<<g, kBuildWarpsSmall * 32, kBuildWarpsSmall * (int)sizeof(HuffScratchT<kSmallSyms>), st>>>(j, j.work_counter);
  }
  {
    const uint64_t want = (nb + kBuildWarpsBig - 1) / kBuildWarpsBig;
    const unsigned g = (unsigned)(want < (uint64_t)num_sms ? want : (uint64_t)num_sms);
    k_build_codes<kMaxSyms, kBuildWarpsBig>
        <<<g, kBuildWarpsBig * 32, kBuildWarpsBig * (int)sizeof(HuffScratch), st>>>(j, j.work_counter + kBuildCounterStride);
  }
}

// ------------------------------------------------------------------
// layout: one thread per stream walks its blocks (bit-contiguous except for
// stored blocks, which pad to a byte after their 3 header bits, hbw:483-486).
__global__ void k_layout(DeflateJob j)
{
  const uint64_t st = j.st_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (st >= j.st_end) return;
  const uint64_t b0 = j.stream_blk0[st], b1 = j.stream_blk0[st + 1];
  const uint64_t L = j.stream_off[st + 1] - j.stream_off[st];
  uint64_t bit = j.cont_start_bit; // (0 unless the stream continues one of an earlier call: single-stream calls)
  for (uint64_t b = b0; b < b1; b++) {
    j.blk_bit_start[b] = bit;
    if (j.blk_kind[b] == kKindSkip) continue;
    if (j.blk_kind[b] == kKindStored) {
      const uint64_t boff = (b - b0) * (uint64_t)kBlockSize;
      const uint64_t n = (L - boff) < (uint64_t)kBlockSize ? (L - boff) : (uint64_t)kBlockSize;
      bit = ((bit + 3 + 7) & ~7ull) + 32 + 8 * n;
    } else {
      bit += j.blk_bits[b];
    }
  }
  j.stream_trailer_bit[st] = bit;
  if (j.cont_open) { // more data follows: no final block; the last byte may be partial
    j.stream_bytes[st] = (bit + 7) >> 3;
    return;
  }
  bit = ((bit + 3 + 7) & ~7ull) + 32; // final empty stored block (deflate.mbt:171)
  j.stream_bytes[st] = bit >> 3;
}

void launch_layout(const DeflateJob &j, cudaStream_t st)
{
  if (j.st_end <= j.st_begin) return;
  unsigned g = (unsigned)((j.st_end - j.st_begin + 127) / 128);
  k_layout<<<g, 128, 0, st>>>(j);
}

// ------------------------------------------------------------------
// K4: bit packing, one CTA per block.
constexpr int kPackThreads = 256;
#ifndef FB_PACK_ITEMS
#define FB_PACK_ITEMS 8 // measured: 4 -> 2.40 ms per GiB, 6 -> 2.24, 8 -> 2.12, 12 -> 2.71, 16 -> 2.66
#endif
constexpr int kItemsPerThread = FB_PACK_ITEMS;
constexpr int kChunkItems = kPackThreads * kItemsPerThread;
constexpr int kStageWords = kChunkItems * 48 / 32 + 4;

__device__ __forceinline__ void or_byte(uint32_t *dst32, uint64_t byte_pos, uint32_t v)
{
  if (v) atomicOr(&dst32[byte_pos >> 2], v << ((byte_pos & 3) * 8));
}

__device__ __forceinline__ void stage_put(uint32_t *stage, uint32_t bitoff, uint64_t val, int nb)
{
  if (nb == 0) return;
  // the (at most 48) bits land in up to three words: 32-bit funnel shifts instead of 64-bit ones
  const uint32_t w = bitoff >> 5, sh = bitoff & 31;
  const uint32_t vlo = (uint32_t)val, vhi = (uint32_t)(val >> 32);
  const uint32_t a = vlo << sh;
  const uint32_t b = __funnelshift_l(vlo, vhi, sh);
  const uint32_t c = __funnelshift_l(vhi, 0u, sh);
  if (a) atomicOr(&stage[w], a);
  if (b) atomicOr(&stage[w + 1], b);
  if (c) atomicOr(&stage[w + 2], c);
}

__global__ void __launch_bounds__(kPackThreads) k_pack(DeflateJob j)
{
  __shared__ uint32_t codes[kFreqStride];
  __shared__ uint32_t stage[kStageWords];
  __shared__ uint32_t wsum[kPackThreads / 32];
  __shared__ uint32_t carry_word;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t blk = j.blk_begin + blockIdx.x;
  if (j.dst_off[j.st_end] > j.dst_cap) return; // the range does not fit: nothing is written, the host reports the need
  const BlockRef r = block_ref(j, blk);
  const int kind = j.blk_kind[blk];
  if (kind == kKindSkip) return;
  uint32_t *dst32 = reinterpret_cast<uint32_t *>(j.dst);
  const uint64_t B0 = j.dst_off[r.st] * 8 + j.blk_bit_start[blk];
  const uint8_t *src = j.src + r.src_off;

  if (kind == kKindStored) {
    // write_stored_header(n, false) + write_bytes (hbw:474-487, :202-225): the 3
    // header bits are zero; LEN / ~LEN / payload start at the next byte boundary.
    const uint64_t p = (B0 + 3 + 7) >> 3;
    const uint32_t n = r.n;
    if (tid == 0) {
      or_byte(dst32, p + 0, n & 0xff);
      or_byte(dst32, p + 1, (n >> 8) & 0xff);
      or_byte(dst32, p + 2, (~n) & 0xff);
      or_byte(dst32, p + 3, ((~n) >> 8) & 0xff);
    }
    for (uint32_t i = tid; i < n; i += kPackThreads) or_byte(dst32, p + 4 + i, __ldg(src + i));
    return;
  }

  const uint64_t B1 = B0 + j.blk_bits[blk];
  for (int i = tid; i < kFreqStride; i += kPackThreads) codes[i] = j.blk_code[blk * kFreqStride + i];
  for (int i = tid; i < kStageWords; i += kPackThreads) stage[i] = 0;
  __syncthreads();

  uint64_t cur = B0;                // absolute bit cursor
  uint64_t win = B0 & ~31ull;       // absolute bit of stage[0] bit 0
  const uint32_t hdr_nbits = j.blk_hdr_nbits[blk];
  const uint32_t hdr_words = (hdr_nbits + 31) >> 5;
  const uint32_t *hdr = j.blk_hdr + blk * kHdrWords;
  const uint32_t *tok = j.tokens + r.src_off;
  const uint32_t nitems = (kind == kKindDynamic ? j.blk_ntok[blk] : r.n) + 1; // + EOB
  // chunk -1 is the header; chunks 0.. are tokens / literal bytes
  const int nchunks = (int)((nitems + kChunkItems - 1) / kChunkItems);
  for (int ch = -1; ch < nchunks; ch++) {
    uint64_t val[kItemsPerThread];
    int nb[kItemsPerThread];
    int tb = 0;
#pragma unroll
    for (int k = 0; k < kItemsPerThread; k++) { val[k] = 0; nb[k] = 0; }
    if (ch < 0) {
      // header words: one per thread (hdr_words <= kHdrWords <= kPackThreads)
      if ((uint32_t)tid < hdr_words) {
        val[0] = hdr[tid];
        nb[0] = (int)min(32u, hdr_nbits - 32u * (uint32_t)tid);
        tb = nb[0];
      }
    } else {
      const uint32_t base = (uint32_t)ch * kChunkItems + (uint32_t)tid * kItemsPerThread;
#pragma unroll
      for (int k = 0; k < kItemsPerThread; k++) {
        const uint32_t i = base + k;
        if (i < nitems) {
          uint32_t t;
          if (i == nitems - 1) t = kEob;
          else t = (kind == kKindDynamic) ? __ldcs(tok + i) : (uint32_t)__ldg(src + i);
          if (t < kMatchType) { // literal / EOB (hbw:609-612)
            const uint32_t e = codes[t];
            val[k] = e & 0xffff;
            nb[k] = (int)(e >> 16);
          } else { // length code, extra, offset code, extra (hbw:614-705)
            int lc, lnb, oc, onb;
            uint32_t lex, oex;
            length_code_of((t - kMatchType) >> kLengthShift, lc, lnb, lex);
            offset_code_of(t & kOffsetMask, oc, onb, oex);
            const uint32_t le = codes[kLenCodesStart + lc];
            const uint32_t oe = codes[kNumLit + oc];
            uint64_t v = le & 0xffff;
            int b = (int)(le >> 16);
            v |= (uint64_t)lex << b; b += lnb;
            v |= (uint64_t)(oe & 0xffff) << b; b += (int)(oe >> 16);
            v |= (uint64_t)oex << b; b += onb;
            val[k] = v;
            nb[k] = b;
          }
          tb += nb[k];
        }
      }
    }
    // exclusive scan of tb over the CTA
    uint32_t x = (uint32_t)tb;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kPackThreads / 32; w++) {
      const uint32_t s = wsum[w];
      if (w < warp) woff += s;
      total += s;
    }
    uint32_t off = (uint32_t)(cur - win) + woff + x - (uint32_t)tb;
#pragma unroll
    for (int k = 0; k < kItemsPerThread; k++) {
      stage_put(stage, off, val[k], nb[k]);
      off += (uint32_t)nb[k];
    }
    __syncthreads();
    cur += total;
    // flush complete words
    const uint32_t nfull = (uint32_t)((cur >> 5) - (win >> 5));
    const uint64_t gw0 = win >> 5;
    for (uint32_t i = tid; i < nfull; i += kPackThreads) {
      const uint64_t gw = gw0 + i;
      const uint32_t v = stage[i];
      if (gw * 32 >= B0 && (gw + 1) * 32 <= B1) __stcs(&dst32[gw], v);
      else if (v) atomicOr(&dst32[gw], v);
    }
    if (tid == 0) carry_word = stage[nfull];
    __syncthreads();
    for (uint32_t i = tid; i <= nfull + 2 && i < (uint32_t)kStageWords; i += kPackThreads) stage[i] = 0;
    __syncthreads();
    if (tid == 0) stage[0] = carry_word;
    win += (uint64_t)nfull * 32;
    __syncthreads();
  }
  if (tid == 0) {
    const uint32_t v = stage[0];
    if ((cur & 31) && v) atomicOr(&dst32[win >> 5], v);
    if (cur != B1) atomicAdd(&j.counters[4], 1u); // size mismatch: bug trap checked by the host
  }
}

// Compressor::close tail: write_stored_header(0, true) + flush (deflate.mbt:171-176):
// bits 1,0,0 then pad to a byte, then 00 00 FF FF.
__global__ void k_trailer(DeflateJob j)
{
  const uint64_t st = j.st_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (st >= j.st_end || j.dst_off[j.st_end] > j.dst_cap || j.cont_open) return;
  uint32_t *dst32 = reinterpret_cast<uint32_t *>(j.dst);
  const uint64_t bit = j.dst_off[st] * 8 + j.stream_trailer_bit[st];
  atomicOr(&dst32[bit >> 5], 1u << (bit & 31));
  const uint64_t p = (bit + 3 + 7) >> 3;
  or_byte(dst32, p + 2, 0xff);
  or_byte(dst32, p + 3, 0xff);
}

void launch_pack(const DeflateJob &j, cudaStream_t st)
{
  if (j.blk_end > j.blk_begin) k_pack<<<(unsigned)(j.blk_end - j.blk_begin), kPackThreads, 0, st>>>(j);
  if (j.st_end > j.st_begin) {
    unsigned g = (unsigned)((j.st_end - j.st_begin + 127) / 128);
    k_trailer<<<g, 128, 0, st>>>(j);
  }
}

// K4 ORs its bits into the output: the words of the range are cleared first.  Ranges are packed in order and a
// 32-bit word may be shared by the last stream of one range and the first of the next; it belongs to the earlier
// range (a range clears from the word boundary at or above its start to the word boundary at or above its end).
__global__ void __launch_bounds__(256) k_zero_range(DeflateJob j)
{
  if (j.dst_off[j.st_end] > j.dst_cap) return;
  const uint64_t a = (j.dst_off[j.st_begin] + 3) >> 2, b = (j.dst_off[j.st_end] + 3) >> 2; // words
  uint32_t *dst32 = reinterpret_cast<uint32_t *>(j.dst);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b - a < 64) {
    for (uint64_t k = a + t; k < b; k += stride) dst32[k] = 0;
    return;
  }
  // 16-byte stores between (aligned on the address, dst itself is only 4-byte aligned), single words at the ends
  const uint64_t bw = (reinterpret_cast<uintptr_t>(j.dst) >> 2) & 3;
  const uint64_t a4 = ((a + bw + 3) & ~3ull) - bw, b4 = ((b + bw) & ~3ull) - bw;
  for (uint64_t k = a + t; k < a4; k += stride) dst32[k] = 0;
  uint4 *d4 = reinterpret_cast<uint4 *>(dst32 + a4);
  const uint64_t n4 = (b4 - a4) >> 2;
  for (uint64_t k = t; k < n4; k += stride) __stcs(&d4[k], make_uint4(0, 0, 0, 0));
  for (uint64_t k = b4 + t; k < b; k += stride) dst32[k] = 0;
}

void launch_zero_range(const DeflateJob &j, cudaStream_t st)
{
  if (j.st_end <= j.st_begin) return;
  k_zero_range<<<296, 256, 0, st>>>(j);
}

// CUDA loads kernels lazily, and loading one while another kernel spins on a host-fed watermark can
// deadlock: every kernel of this file is loaded when the context is created.
void preload_encode_kernels()
{
  // function attributes are per device: set when the context is created
  cudaFuncSetAttribute(k_build_codes<kSmallSyms, kBuildWarpsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       kBuildWarpsSmall * (int)sizeof(HuffScratchT<kSmallSyms>));
  cudaFuncSetAttribute(k_build_codes<kMaxSyms, kBuildWarpsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       kBuildWarpsBig * (int)sizeof(HuffScratch));
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_histogram);
  cudaFuncGetAttributes(&a, k_build_codes<kSmallSyms, kBuildWarpsSmall>);
  cudaFuncGetAttributes(&a, k_build_codes<kMaxSyms, kBuildWarpsBig>);
  cudaFuncGetAttributes(&a, k_zero_range);
  cudaFuncGetAttributes(&a, k_layout);
  cudaFuncGetAttributes(&a, k_pack);
  cudaFuncGetAttributes(&a, k_trailer);
}

} // namespace fb
