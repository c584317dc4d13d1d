// kernels.h -- host-side launch interface of the CUDA kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

// Everything the deflate kernels need; all pointers are device pointers.
// Streams are the independent units: stream i is
// src[stream_off[i] .. stream_off[i+1]) and is cut into blocks of <= 65535
// bytes exactly as Compressor::write / enc_speed does (deflate.mbt:222-294).
struct DeflateJob {
  const uint8_t *src;
  const uint64_t *stream_off; // [nstreams+1]
  uint64_t nstreams;
  uint64_t nblocks;           // total over streams (host copy of stream_blk0[nstreams])
  // per stream
  uint64_t *stream_blk0;      // [nstreams+1] first block index (exclusive scan)
  uint64_t *stream_bytes;     // [nstreams]   compressed size
  uint64_t *stream_trailer_bit; // [nstreams] stream-relative bit offset of the final stored header
  uint64_t *dst_off;          // [nstreams+1] output byte offsets (exclusive scan of stream_bytes)
  // per block
  uint32_t *blk_stream;       // [nblocks] owning stream
  uint32_t *blk_ntok;         // [nblocks] tokens emitted by the parse (0 if not parsed)
  uint8_t *blk_kind;          // [nblocks] kKind*
  uint32_t *blk_bits;         // [nblocks] bits of header+data+EOB (huff/dynamic kinds)
  uint64_t *blk_bit_start;    // [nblocks] stream-relative start bit
  uint32_t *blk_hdr_nbits;    // [nblocks]
  uint32_t *blk_hdr;          // [nblocks][kHdrWords] header bit string, LSB first
  uint32_t *blk_freq;         // [nblocks][320] lit/len (286) + offset (30) histograms (+pad)
  uint32_t *blk_code;         // [nblocks][320] code | len<<16 for lit/len (286) + offset (30)
  // tokens of the block starting at source byte o live at tokens[o .. o+ntok)
  uint32_t *tokens;           // [n_total]
  uint32_t *counters;         // [8] work-stealing counters, zeroed per call
  // host-buffer calls: number of streams whose bytes have arrived (advanced by the H2D stream after
  // every chunk); the parse waits on it, so the copy overlaps the kernel.  Null: everything is resident.
  const uint32_t *avail;
  // range the post-parse kernels work on in this launch (whole batch: 0..nblocks / 0..nstreams)
  uint64_t blk_begin, blk_end, st_begin, st_end;
  uint32_t *work_counter;        // K3 work distribution of this launch (zeroed by the host)
  uint64_t dst_cap;              // K4 writes nothing for a range that would end beyond it (host reports the need)
  // output
  uint8_t *dst;
  // Continuation of ONE stream across calls (the streaming Writer: every call compresses the full 65535-byte
  // windows that have arrived, the hash table and the bit position carry over).  The call then holds a single
  // stream.  cont_prev: its first block is a stand-in for what came before -- only its last 32768 bytes are real
  // (the history a 4-byte candidate check may read, D1) -- it is never parsed and emits nothing, and the block
  // behind it is seeded with the end table the previous call left.  cont_open: more data follows, so the final
  // empty stored block is not written and the end table of the last block is kept.
  uint32_t cont_prev, cont_open;
  uint32_t cont_start_bit;  // bits of the first output byte that the previous call has used already (0..7)
  uint64_t cont_block_base; // cont_prev: block b >= 1 of this call is block cont_block_base + b - 1 of the whole stream (table resets)
};
constexpr int kFreqStride = 320;

// Block-parallel parse of multi-block streams.  The blocks of one stream are chained only through the hash
// table (what the previous block left within 32768 bytes of its end), so all blocks are parsed at once from a
// guessed table (empty in round 1) and a block is parsed again whenever the end table of its predecessor has
// changed; block 0 is exact after round 1, and the fixpoint -- no end table changed -- is the sequential result.
// Text needs about 5 rounds whatever the number of blocks.  All pointers are device pointers.
struct BlockParJob {
  const uint32_t *list;    // blocks to parse in this round (global block indices)
  uint32_t nlist;
  const uint64_t *mb_idx;  // [nblocks + 1] index of a block among the blocks of multi-block streams
  uint16_t *tabs;          // [2][nmb][1 << 14] normalised end tables (distance from the block end, 0 = none in reach)
  uint64_t nmb;
  const uint8_t *lat_prev; // [nmb] which of the two copies is a block's latest end table (before this round)
  uint8_t *lat_next;       // [nmb]
  uint8_t *chg_next;       // [nmb] the block's end table changed in this round
  int round;
  // DeflateJob::cont_prev / cont_open / cont_block_base of the call (see there)
  uint32_t cont_prev, cont_open;
  uint64_t cont_block_base;
};

// one-time device tables (probe schedule); call once per context
void launch_init_tables(cudaStream_t st);
// setup: per-stream block counts -> stream_blk0 (scan) ; then per-block stream ids
void launch_count_blocks(const DeflateJob &j, cudaStream_t st);
void launch_count_multi(const DeflateJob &j, cudaStream_t st);
void launch_fill_blocks(const DeflateJob &j, cudaStream_t st);
// K1: greedy LZ77 parse, one warp per stream (deflate-fast.mbt:123-342)
// gtables: parse_gtables_bytes(num_sms) bytes of device scratch on the launching GPU (tables of the warps that have
// no shared-memory table)
size_t parse_gtables_bytes(int num_sms);
void launch_parse(const DeflateJob &j, int num_sms, void *gtables, cudaStream_t st);
// the two halves of launch_parse: streams with one parsed block / streams with several (sequential per stream)
void launch_parse_single(const DeflateJob &j, int num_sms, void *gtables, cudaStream_t st);
void launch_parse_multi(const DeflateJob &j, int num_sms, void *gtables, cudaStream_t st);
// block-parallel rounds for the multi-block streams (see BlockParJob)
void launch_bp_flags(const DeflateJob &j, uint64_t *flags, cudaStream_t st);
void launch_bp_round(const DeflateJob &j, const BlockParJob &bp, const uint8_t *chg_prev, uint32_t *list, uint32_t *nlist,
                     cudaStream_t st);
void launch_parse_blocks(const DeflateJob &j, const BlockParJob &bp, uint32_t *counter, int num_sms, void *gtables,
                         cudaStream_t st);
// continuation: the latest end table of multi-block-stream block m -> out (1 << 14 uint16), after the last round
void launch_bp_save_table(const BlockParJob &bp, uint64_t m, const uint8_t *lat, uint16_t *out, cudaStream_t st);
// K2: block kind + histograms (huffman-bit-writer.mbt:550-593, :831)
void launch_histogram(const DeflateJob &j, cudaStream_t st);
// K3: code construction + codegen + header + sizes (huffman-code.mbt:112-343,
//     huffman-bit-writer.mbt:241-471)
// j.work_counter: two zeroed uint32 per launch, kBuildCounterStride apart (small-scratch pass, full-size pass)
constexpr int kBuildCounterStride = 256;
void launch_build_codes(const DeflateJob &j, int num_sms, cudaStream_t st);
// layout: per-stream bit offsets, stream sizes, output offsets
void launch_layout(const DeflateJob &j, cudaStream_t st);
// K4: bit packing (huffman-bit-writer.mbt:596-824, :474-487) + stream trailers
void launch_pack(const DeflateJob &j, cudaStream_t st);
// clears the output words of streams [st_begin, st_end) (K4 ORs into them); a word shared with the previous
// range belongs to that range's clear
void launch_zero_range(const DeflateJob &j, cudaStream_t st);
// out[i] = in[min(i * stride, n)], i in [0, cnt): group boundaries of an offset array
void launch_gather_u64(uint64_t *out, const uint64_t *in, uint64_t stride, uint64_t n, uint64_t cnt, cudaStream_t st);

// exclusive scan of n uint64 values (out may alias in); out has n+1 entries.  carry_from: start value read
// from device memory (may be out itself: the total the previous range's scan left there), or null for 0
// host_total: pinned host memory that receives the total as well (written by the kernel itself), or null
void launch_scan_u64(const uint64_t *in, uint64_t *out, uint64_t n, cudaStream_t st, const uint64_t *carry_from = nullptr,
                     uint64_t *host_total = nullptr);
// fixed-size segment offsets: off[i] = min(i*seg, n), i in [0, nseg]
void launch_fill_seg_off(uint64_t *off, uint64_t nseg, uint64_t seg, uint64_t n, cudaStream_t st);

// out[i] = in[i] + delta (mod 2^64), i in [0, cnt): re-bases offset arrays for chunked host calls
void launch_affine_u64(uint64_t *out, const uint64_t *in, uint64_t cnt, uint64_t delta, cudaStream_t st);

// frame reader: sizes = the u32 size array of a frame header (device or peer memory).  comp_off[k] = sum of
// sizes[first .. first + k), k in [0, count]; res[0] = sum of sizes[0 .. first), res[1] = comp_off[count]
void launch_frame_range(const uint32_t *sizes, uint64_t first, uint64_t count, uint64_t *comp_off, uint64_t *res,
                        cudaStream_t st);

void preload_parse_kernels();
void preload_encode_kernels();
void preload_inflate_kernels();

struct InflateJob {
  const uint8_t *comp;
  const uint64_t *comp_off; // [nstreams+1]
  uint64_t nstreams;
  uint8_t *out;
  const uint64_t *out_off;  // [nstreams+1] capacity slots
  uint64_t *out_len;        // [nstreams]
  int32_t *status;          // [nstreams]
  int64_t *err_off;         // [nstreams]
  uint64_t *consumed;       // [nstreams] or null
  uint32_t *fallback;       // [nstreams] streams the fast path hands to the exact kernel
  uint32_t *counters;       // [0] fast work counter, [1] exact work counter, [2] fallback count, [3] copy work counter
  // host-buffer calls (all null / 0 otherwise): streams whose input has arrived; per-group completion
  // counters and host-visible flags so that finished output groups can be copied back while the kernel runs
  const uint32_t *avail;
  uint32_t *group_done;     // [ngroups]
  volatile uint32_t *group_flag; // [ngroups] mapped pinned host memory
  uint32_t group_streams;   // streams per group
  // recorded back-references of the fast path
  uint2 *records;           // {dst, len | (dist-1) << 16}
  const uint64_t *rec_off;  // [nstreams+1] record area of each stream
  uint32_t *nrec;           // [nstreams] records written (0: nothing to copy, or stream handed to the exact kernel)
  const uint32_t *order;    // [nstreams] decode order (largest compressed size first)
  uint32_t window_bits;     // inflate3: speculation window of a stream's first block (0: the library's default)
  uint32_t cta_streams;     // calls with at most this many streams decode with one CTA per stream (256 ranges per block)
  // preset dictionary (&Reader::new_dict, inflate.mbt:310-317; DictDecoder::new, dict-decoder.mbt:42-60): hist0[i]
  // bytes of history (<= 32768) lie directly in front of stream i's output slot; back-references may reach into
  // them (dist > hist_size is the reference's only limit, inflate.mbt:677).  Null: no dictionaries.
  const uint32_t *hist0;
};
// K6: batched inflate.  Fast kernel: one warp per stream, the lanes decode one block in parallel (inflate3.cu);
// the exact kernel (inflate.cu) then re-decodes the streams on the fallback list with the reference's error behaviour.
void launch_inflate3(const InflateJob &j, int num_sms, cudaStream_t st);
void launch_inflate_exact(const InflateJob &j, int num_sms, cudaStream_t st);
void preload_inflate3_kernels();
// decode order for device-resident calls (largest compressed size first); order_hist: device scratch of 1024 uint32
void launch_stream_order(const InflateJob &j, uint32_t *order_hist, cudaStream_t st);
void launch_rec_off(const uint64_t *out_off, uint64_t *rec_off, uint64_t ns, cudaStream_t st);

} // namespace fb
