// common.cuh -- constants and small device helpers shared by the sm_100a kernels.
//
// Constants are the parity contract of the reference:
//   deflate-fast.mbt:12-55,:89-92  table/hash/match limits
//   token.mbt:8-24                 token packing
//   huffman-bit-writer.mbt:11-85   alphabets, extra-bit tables, codegen order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

constexpr int kTableBits = 14;                  // deflate-fast.mbt:12
constexpr int kTableSize = 1 << kTableBits;     // :15
constexpr int kTableShift = 32 - kTableBits;    // :21
constexpr uint32_t kHashMul = 0x1e35a7bdu;      // :79
constexpr int kMaxMatchOffset = 1 << 15;        // :40
constexpr int kMaxMatchLength = 258;            // :34
constexpr int kBlockSize = 65535;               // max_store_block_size :46
constexpr int kInputMargin = 15;                // :89
constexpr int kBufferReset = 2147483647 - 2 * kBlockSize; // buffer_reset :55
constexpr uint32_t kMatchType = 1u << 30;       // token.mbt:24
constexpr int kLengthShift = 22;                // token.mbt:13
constexpr uint32_t kOffsetMask = (1u << kLengthShift) - 1;
constexpr int kNumLit = 286;                    // inflate.mbt:28
constexpr int kNumDist = 30;                    // inflate.mbt:31
constexpr int kNumCodegen = 19;                 // huffman-bit-writer.mbt:26
constexpr int kEob = 256;                       // :16
constexpr int kLenCodesStart = 257;             // :21

// DeflateFast.cur is 65535 before block 0 and grows by the block length; encode clears the table when it has
// reached buffer_reset (:129-132): at the start of block 32766 (65535 * 32767 >= buffer_reset), and then, cur
// restarting at 32769, every 32766 blocks again (checked against the running sum by tests/test_abi.py through fb200_debug_block_resets).
__host__ __device__ inline bool block_resets_table(uint64_t b)
{
  return b != 0 && b % 32766 == 0;
}

// block kinds (deflate.mbt:236-277)
constexpr int kKindStored = 0;   // n <= 16
constexpr int kKindHuff = 1;     // 17 <= n <= 127, or tokens > n - n/16
constexpr int kKindDynamic = 2;
constexpr int kKindSkip = 3;     // stand-in block of a continued stream (DeflateJob::cont_prev): no bits

// per-block header scratch: 14 + 19*3 + 316 * (7 + 7) bits worst case < 4608 bits
constexpr int kHdrWords = 160;

__host__ __device__ __forceinline__ uint32_t hash4(uint32_t u) { return (u * kHashMul) >> kTableShift; }

// token.mbt:107 length_code(xlen) with xlen = length - 3 in [0,255] and the
// extra-bit count / value of huffman-bit-writer.mbt:49-62, computed instead of
// looked up.  code < 8: no extra bits.  Otherwise nb = floor(log2 xlen) - 2,
// code = 4*nb + 4 + ((xlen >> nb) & 3), extra = xlen & ((1<<nb)-1); xlen 255 is
// the special code 28 (length 258) with no extra bits.
__host__ __device__ __forceinline__ void length_code_of(uint32_t xlen, int &code, int &nb, uint32_t &extra)
{
  if (xlen < 8) {
    code = (int)xlen; nb = 0; extra = 0;
  } else if (xlen == 255) {
    code = 28; nb = 0; extra = 0;
  } else {
#ifdef __CUDA_ARCH__
    int lg = 31 - __clz(xlen);
#else
    int lg = 31 - __builtin_clz(xlen);
#endif
    nb = lg - 2;
    code = 4 * nb + 4 + (int)((xlen >> nb) & 3);
    extra = xlen & ((1u << nb) - 1);
  }
}

// token.mbt:112-123 offset_code(xoff) with xoff = distance - 1 in [0,32767] and
// huffman-bit-writer.mbt:67-78.  xoff < 4: code = xoff.  Otherwise
// nb = floor(log2 xoff) - 1, code = 2*nb + 2 + ((xoff >> nb) & 1).
__host__ __device__ __forceinline__ void offset_code_of(uint32_t xoff, int &code, int &nb, uint32_t &extra)
{
  if (xoff < 4) {
    code = (int)xoff; nb = 0; extra = 0;
  } else {
#ifdef __CUDA_ARCH__
    int lg = 31 - __clz(xoff);
#else
    int lg = 31 - __builtin_clz(xoff);
#endif
    nb = lg - 1;
    code = 2 * nb + 2 + (int)((xoff >> nb) & 1);
    extra = xoff & ((1u << nb) - 1);
  }
}

#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Unaligned little-endian 32-bit load from global memory (deflate-fast.mbt:58-63)
// as two aligned word loads + funnel shift.  Reads up to byte (p|3)+4.
__device__ __forceinline__ uint32_t ld32u(const uint8_t *p)
{
  uintptr_t a = (uintptr_t)p;
  const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t lo = __ldg(q);
  uint32_t hi = sh ? __ldg(q + 1) : 0u;
  return __funnelshift_r(lo, hi, sh);
}
#endif

} // namespace fb
