// huff_build.cuh -- K3 building blocks: length-limited canonical code
// construction, codegen RLE, dynamic-block header.  Pure integer code that
// compiles for the device (one thread per block in k_build_codes) and for the
// host (the CPU test-suite checks it against the oracle without a GPU).
//
//   HuffmanEncoder::bit_counts / assign_encoding_and_size / generate
//       huffman-code.mbt:112-343
//   HuffmanBitWriter::generate_codegen / dynamic_size / write_dynamic_header
//       huffman-bit-writer.mbt:241-360, :421-471
#pragma once
#include "common.cuh"

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


constexpr int kMaxSyms = 288;
constexpr int kIntMax = 2147483647;

// HuffmanEncoder::bit_counts (huffman-code.mbt:112-244).  fr[0..n) are the
// non-zero frequencies in (freq, literal) order, fr[n] = MAX.  Fills
// bit_count[1..max_bits] and returns max_bits = min(limit, n-1).
FB_HD inline int bit_counts_dev(const int *fr, int n, int max_bits, int *bit_count)
{
  if (max_bits > n - 1) max_bits = n - 1;
  int last_freq[16], next_char[16], next_pair[16], needed[16];
  unsigned short leaf[16][16];
  for (int l = 0; l < 16; l++) {
    last_freq[l] = next_char[l] = next_pair[l] = needed[l] = 0;
    for (int k = 0; k < 16; k++) leaf[l][k] = 0;
  }
  for (int level = 1; level <= max_bits; level++) { // :149-163
    last_freq[level] = fr[1];
    next_char[level] = fr[2];
    next_pair[level] = fr[0] + fr[1];
    leaf[level][level] = 2;
    if (level == 1) next_pair[level] = kIntMax;
  }
  needed[max_bits] = 2 * n - 4; // :166
  int level = max_bits;
  for (;;) { // :170-224
    if (next_pair[level] == kIntMax && next_char[level] == kIntMax) {
      needed[level] = 0;
      next_pair[level + 1] = kIntMax;
      level++;
      continue;
    }
    const int prev_freq = last_freq[level];
    if (next_char[level] < next_pair[level]) { // leaf
      const int nn = leaf[level][level] + 1;
      last_freq[level] = next_char[level];
      leaf[level][level] = (unsigned short)nn;
      next_char[level] = fr[nn];
    } else { // pair from the level below
      last_freq[level] = next_pair[level];
      for (int i = 0; i < level; i++) leaf[level][i] = leaf[level - 1][i];
      needed[level - 1] = 2;
    }
    needed[level]--;
    if (needed[level] == 0) {
      if (level == max_bits) break;
      next_pair[level + 1] = (int)((unsigned)prev_freq + (unsigned)last_freq[level]);
      level++;
    } else {
      while (needed[level - 1] > 0) level--;
    }
  }
  int bits = 1;
  for (int lv = max_bits; lv > 0; lv--) { // :236-242
    bit_count[bits] = (int)leaf[max_bits][lv] - (int)leaf[max_bits][lv - 1];
    bits++;
  }
  return max_bits;
}

// HuffmanEncoder::generate (:295-343) + assign_encoding_and_size (:250-280).
// freq[0..nsym) -> len[0..nsym) (0 for unused symbols), code[] bit-reversed.
FB_HD inline void generate_dev(const uint32_t *freq, int nsym, int max_bits, unsigned char *len, unsigned short *code)
{
  unsigned int keys[kMaxSyms + 1]; // (freq << 9) | literal: ascending == by_frequency (:346-351)
  int fr[kMaxSyms + 1];
  int count = 0;
  for (int i = 0; i < nsym; i++) {
    len[i] = 0;
    code[i] = 0;
    if (freq[i] != 0) keys[count++] = (freq[i] << 9) | (unsigned)i;
  }
  if (count <= 2) { // :326-336: codes 0,1 in literal order, length 1
    for (int i = 0; i < count; i++) {
      const int sym = keys[i] & 511;
      len[sym] = 1;
      code[sym] = (unsigned short)i;
    }
    return;
  }
  // sort ascending (any correct sort: keys are distinct) -- shell sort
  for (int gap = count >> 1; gap > 0; gap = (gap == 2) ? 1 : (int)(gap * 5 / 11)) {
    for (int i = gap; i < count; i++) {
      const unsigned int v = keys[i];
      int k = i;
      while (k >= gap && keys[k - gap] > v) {
        keys[k] = keys[k - gap];
        k -= gap;
      }
      keys[k] = v;
    }
  }
  for (int i = 0; i < count; i++) fr[i] = (int)(keys[i] >> 9);
  fr[count] = kIntMax; // max_node (:77-79)
  int bit_count[17];
  const int mb = bit_counts_dev(fr, count, max_bits, bit_count);
  // lengths: the `bits` most frequent remaining symbols get length n (:260-278)
  int end = count;
  for (int nb = 1; nb <= mb; nb++) {
    const int bits = bit_count[nb];
    for (int i = end - bits; i < end; i++) len[keys[i] & 511] = (unsigned char)nb;
    end -= bits;
  }
  // canonical codes in literal order per length, stored bit-reversed (:268-277)
  unsigned next_code[17];
  unsigned c = 0;
  next_code[0] = 0;
  for (int nb = 1; nb <= mb; nb++) {
    c <<= 1;
    next_code[nb] = c;
    c += (unsigned)bit_count[nb];
  }
  for (int i = 0; i < nsym; i++) {
    const int l = len[i];
    if (l) {
      const unsigned v = next_code[l]++;
      code[i] = (unsigned short)(brev32(v) >> (32 - l));
    }
  }
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


// Everything write_block_dynamic / write_block_huff decide before the first
// payload bit (hbw:496-534, :738-787): codes, codegen, "store instead" test,
// header bit string, exact block size.  freq = lit/len[286] ++ offset[30]
// histogram of the block (EOB already counted); kind = kKindDynamic / kKindHuff.
struct BlockBuild {
  int kind;           // may turn into kKindStored (quirk D2 test)
  uint32_t hdr_nbits; // bits in hdr_words
  uint32_t blk_bits;  // header + payload + EOB
};

FB_HD inline BlockBuild build_block_dev(uint32_t *freq, int kind, uint32_t n, uint32_t *codeout, uint32_t *hdr_words)
{
  BlockBuild res;
  res.kind = kind;
  res.hdr_nbits = 0;
  res.blk_bits = 0;
  unsigned char len[kNumLit + kNumDist];
  unsigned short code[kNumLit + kNumDist];
  int num_literals, num_offsets;
  if (kind == kKindDynamic) { // index_tokens tail (hbw:574-592)
    num_literals = kNumLit;
    while (freq[num_literals - 1] == 0) num_literals--;
    num_offsets = kNumDist;
    while (num_offsets > 0 && freq[kNumLit + num_offsets - 1] == 0) num_offsets--;
    if (num_offsets == 0) {
      freq[kNumLit] = 1;
      num_offsets = 1;
    }
    generate_dev(freq, kNumLit, 15, len, code);
    generate_dev(freq + kNumLit, kNumDist, 15, len + kNumLit, code + kNumLit);
  } else { // write_block_huff (hbw:747-758): literal-only, static huff_offset (huffman-code.mbt:691)
    num_literals = kEob + 1;
    num_offsets = 1;
    generate_dev(freq, kNumLit, 15, len, code);
    for (int i = 0; i < kNumDist; i++) { len[kNumLit + i] = 0; code[kNumLit + i] = 0; }
    len[kNumLit] = 1;
  }

  // generate_codegen (hbw:241-330)
  unsigned char cg[kNumLit + kNumDist + 2];
  uint32_t cgfreq[kNumCodegen];
  for (int i = 0; i < kNumCodegen; i++) cgfreq[i] = 0;
  for (int i = 0; i < num_literals; i++) cg[i] = len[i];
  for (int i = 0; i < num_offsets; i++) cg[num_literals + i] = len[kNumLit + i];
  cg[num_literals + num_offsets] = 255;
  {
    unsigned char size = cg[0];
    int count = 1, out = 0;
    for (int in = 1; size != 255; in++) {
      const unsigned char next = cg[in];
      if (next == size) { count++; continue; }
      if (size != 0) {
        cg[out++] = size; cgfreq[size]++; count--;
        while (count >= 3) {
          const int nn = count < 6 ? count : 6;
          cg[out++] = 16; cg[out++] = (unsigned char)(nn - 3); cgfreq[16]++; count -= nn;
        }
      } else {
        while (count >= 11) {
          const int nn = count < 138 ? count : 138;
          cg[out++] = 18; cg[out++] = (unsigned char)(nn - 11); cgfreq[18]++; count -= nn;
        }
        if (count >= 3) {
          cg[out++] = 17; cg[out++] = (unsigned char)(count - 3); cgfreq[17]++; count = 0;
        }
      }
      count--;
      for (; count >= 0; count--) { cg[out++] = size; cgfreq[size]++; }
      size = next;
      count = 1;
    }
    cg[out] = 255;
  }
  unsigned char cglen[kNumCodegen];
  unsigned short cgcode[kNumCodegen];
  generate_dev(cgfreq, kNumCodegen, 7, cglen, cgcode);

  // dynamic_size (hbw:335-360)
  const int order[kNumCodegen] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int num_codegens = kNumCodegen;
  while (num_codegens > 4 && cgfreq[order[num_codegens - 1]] == 0) num_codegens--;
  int size = 3 + 5 + 5 + 4 + 3 * num_codegens + (int)(cgfreq[16] * 2 + cgfreq[17] * 3 + cgfreq[18] * 7);
  for (int i = 0; i < kNumCodegen; i++) size += (int)cgfreq[i] * cglen[i];
  uint32_t extra = 0; // extra bits, not part of `size` (callers pass 0)
  for (int i = 0; i < kNumLit; i++) {
    size += (int)freq[i] * len[i];
    if (i >= kLenCodesStart + 8 && i < kLenCodesStart + 28) extra += freq[i] * (uint32_t)((i - kLenCodesStart - 4) >> 2);
  }
  if (kind == kKindDynamic) {
    for (int i = 0; i < kNumDist; i++) {
      size += (int)freq[kNumLit + i] * len[kNumLit + i];
      if (i >= 4) extra += freq[kNumLit + i] * (uint32_t)((i - 2) >> 1);
    }
  } else {
    size += 1; // offset_freq[0] (forced to 1, hbw:756) * huff_offset.codes[0].len
  }
  // "store instead" test (hbw:526-531, :779-784; quirk D2: (size+size)>>4)
  const int ssize = ((int)n + 5) * 8;
  if (ssize < ((size + size) >> 4)) {
    res.kind = kKindStored;
    return res;
  }

  // write_dynamic_header (hbw:421-471)
  BitAcc ba;
  ba.words = hdr_words;
  ba.acc = 0; ba.nacc = 0; ba.nwords = 0;
  ba.put(4, 3); // BFINAL = 0, BTYPE = 10
  ba.put((uint32_t)(num_literals - 257), 5);
  ba.put((uint32_t)(num_offsets - 1), 5);
  ba.put((uint32_t)(num_codegens - 4), 4);
  for (int i = 0; i < num_codegens; i++) ba.put(cglen[order[i]], 3);
  for (int i = 0;;) {
    const int cw = cg[i++];
    if (cw == 255) break;
    ba.put(cgcode[cw], cglen[cw]);
    if (cw == 16) ba.put(cg[i++], 2);
    else if (cw == 17) ba.put(cg[i++], 3);
    else if (cw == 18) ba.put(cg[i++], 7);
  }
  const int hdr_bits = ba.finish();
  res.hdr_nbits = (uint32_t)hdr_bits;
  // total bits of the block = header + sum(freq*len) + extra bits
  uint32_t data_bits = extra;
  for (int i = 0; i < kNumLit; i++) data_bits += freq[i] * len[i];
  if (kind == kKindDynamic)
    for (int i = 0; i < kNumDist; i++) data_bits += freq[kNumLit + i] * len[kNumLit + i];
  res.blk_bits = (uint32_t)hdr_bits + data_bits;
  for (int i = 0; i < kNumLit + kNumDist; i++) codeout[i] = (uint32_t)code[i] | ((uint32_t)len[i] << 16);
  return res;
}

} // namespace fb
