/*
 * flate_b200.h -- C ABI of libflate_b200.so, the B200-native (sm_100a CUDA)
 * drop-in for the deflate-fast encoder / inflate decoder hot path of
 * gmlewis/moonbit-flate.
 *
 * The reference has no FFI: its public surface is the MoonBit root package
 * (pkg.generated.mbti:9-45).  A MoonBit native-backend `extern "c"` shim
 * (moonbit_flate_b200/moonbit/) binds exactly the entry points below and
 * re-exposes Writer::new/write/close and Reader::new/read/close on top of them;
 * INTEGRATION.md shows the binding.  Plain pointers and sizes only; no
 * exceptions, no C++ or torch types cross this boundary.
 *
 * What each entry point replaces in the reference (paths in /root/reference):
 *   fb200_deflate_*      Writer::new + write + close, i.e. Compressor::write /
 *                        enc_speed / close (writer.mbt:10-58, deflate.mbt:157-294),
 *                        DeflateFast::encode (deflate-fast.mbt:123-342),
 *                        HuffmanBitWriter::write_block_dynamic / write_block_huff /
 *                        write_stored_header (huffman-bit-writer.mbt:474-824),
 *                        HuffmanEncoder::generate (huffman-code.mbt:295-343)
 *   fb200_inflate_*      Reader::new + Decompressor.read until ioeof
 *                        (inflate.mbt:305-407), next_block / read_huffman /
 *                        read_literal / copy_history / data_block
 *                        (inflate.mbt:345-777), DictDecoder (dict-decoder.mbt)
 *   fb200_writer_*       the streaming Writer object (writer.mbt:10-58)
 *   fb200_reader_*       the streaming Decompressor object (inflate.mbt:257-418)
 *
 * Every stream produced by fb200_deflate_* is byte-identical to what the
 * reference's Writer emits for the same bytes; every stream is decoded
 * bit-exactly as the reference's Decompressor would, including its error
 * classification and "corrupt input before offset N" offsets.
 *
 * There is no CPU fallback: every compute entry point fails with
 * FB200_ERR_CUDA when no sm_100 device / driver is usable.
 */
#ifndef FLATE_B200_H
#define FLATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_VERSION 1

/* ---- call-level return codes ---- */
#define FB200_OK 0
#define FB200_ERR_ARG (-1)           /* bad argument (NULL, offsets not monotone, ...) */
#define FB200_ERR_DST_TOO_SMALL (-2) /* output capacity insufficient; *out_len = need  */
#define FB200_ERR_CUDA (-3)          /* CUDA runtime / launch failure, or no device    */
#define FB200_ERR_CLOSED (-4)        /* write after close: "writer closed" (deflate.mbt:154) */
#define FB200_ERR_NOMEM (-5)

/* ---- per-stream inflate status (status[] arrays) ---- */
#define FB200_ST_EOF 0            /* final block reached: reference err == ioeof          */
#define FB200_ST_CORRUPT 1        /* "flate: corrupt input before offset N", N = err_off  */
#define FB200_ST_UNEXPECTED_EOF 2 /* @io.err_unexpected_eof (inflate.mbt:781-786)         */
#define FB200_ST_DST_TOO_SMALL 3  /* output slot too small (not a reference condition)    */
#define FB200_ST_INTERNAL 4       /* "flate: internal error: ..."                         */
#define FB200_ST_EOF_AT_REFILL 5  /* input ended inside more_bits (inflate.mbt:789-799):
                                     the reference reports plain ioeof here as well        */

typedef struct fb200_ctx fb200_ctx;

/* One context per host thread and GPU; not thread-safe.  device < 0 selects the
 * current CUDA device.  Owns a stream and grow-only device scratch. */
int fb200_create(fb200_ctx **out, int device);
void fb200_destroy(fb200_ctx *ctx);
/* Human-readable text of the last failure on this context ("" if none). */
const char *fb200_last_error(const fb200_ctx *ctx);
int fb200_version(void);

/* ------------------------------------------------------------------ */
/* Sizes.                                                              */
/* Upper bound on the compressed size of ONE stream of n bytes. */
uint64_t fb200_deflate_stream_bound(uint64_t n);
/* Upper bound for n bytes cut into seg_size-byte independent streams. */
uint64_t fb200_deflate_bound(uint64_t n, uint64_t seg_size);

/* ------------------------------------------------------------------ */
/* Batch deflate, HOST buffers (copies are part of the call).          */
/* src[0..n) is cut into ceil(n/seg_size) segments; each is compressed as an
 * independent, complete reference stream (ends with the 5-byte final stored
 * block, deflate.mbt:171).  Streams are laid out back to back in dst;
 * seg_off[i] .. seg_off[i+1] delimit stream i (seg_off has nseg+1 entries). */
int fb200_deflate_segments(fb200_ctx *ctx, const uint8_t *src, uint64_t n, uint64_t seg_size,
                           uint8_t *dst, uint64_t dst_cap, uint64_t *seg_off, uint64_t *out_len);
/* Same with explicit stream boundaries: stream i = src[src_off[i] .. src_off[i+1]). */
int fb200_deflate_streams(fb200_ctx *ctx, const uint8_t *src, const uint64_t *src_off, uint64_t nstreams,
                          uint8_t *dst, uint64_t dst_cap, uint64_t *dst_off, uint64_t *out_len);

/* Batch deflate, DEVICE buffers (no host<->device payload copies; the call
 * enqueues on the context stream and synchronises before returning).
 * d_src_off / d_dst_off are device arrays of nstreams+1 uint64.  *out_len
 * (host) receives the total compressed size. */
int fb200_deflate_streams_dev(fb200_ctx *ctx, const uint8_t *d_src, const uint64_t *d_src_off,
                              uint64_t nstreams, uint64_t n_total, uint8_t *d_dst, uint64_t dst_cap,
                              uint64_t *d_dst_off, uint64_t *out_len);
/* Fixed-size segments on device; d_seg_off: device array of nseg+1 uint64. */
int fb200_deflate_segments_dev(fb200_ctx *ctx, const uint8_t *d_src, uint64_t n, uint64_t seg_size,
                               uint8_t *d_dst, uint64_t dst_cap, uint64_t *d_seg_off, uint64_t *out_len);

/* ------------------------------------------------------------------ */
/* Batch inflate.                                                      */
/* Stream i = comp[comp_off[i] .. comp_off[i+1]); its output goes to
 * out[out_off[i] .. out_off[i+1]) (a capacity slot).  out_len[i] = bytes
 * produced (partial output before an error is kept, inflate.mbt:403-405),
 * status[i] one of FB200_ST_*, err_off[i] = the reference's roffset for
 * FB200_ST_CORRUPT, consumed[i] (may be NULL) = input bytes consumed. */
int fb200_inflate_batch(fb200_ctx *ctx, const uint8_t *comp, const uint64_t *comp_off, uint64_t nstreams,
                        uint8_t *out, const uint64_t *out_off, uint64_t *out_len, int32_t *status,
                        int64_t *err_off, uint64_t *consumed);
/* One stream decoded with a preset dictionary (&Reader::new_dict, inflate.mbt:310-317; DictDecoder::new,
 * dict-decoder.mbt:42-60): as if the uncompressed data started with dict, which has already been read.  Host
 * buffers; consumed may be NULL. */
int fb200_inflate_dict(fb200_ctx *ctx, const uint8_t *comp, uint64_t n, const uint8_t *dict, uint64_t dict_len,
                       uint8_t *out, uint64_t cap, uint64_t *out_len, int32_t *status, int64_t *err_off,
                       uint64_t *consumed);
/* Device-buffer variant: every pointer is a device pointer (consumed may be NULL). */
int fb200_inflate_batch_dev(fb200_ctx *ctx, const uint8_t *d_comp, const uint64_t *d_comp_off,
                            uint64_t nstreams, uint8_t *d_out, const uint64_t *d_out_off, uint64_t *d_out_len,
                            int32_t *d_status, int64_t *d_err_off, uint64_t *d_consumed);

/* ------------------------------------------------------------------ */
/* Asynchronous forms of the host-buffer calls: they return at once, the work runs on a helper thread owned by
 * the context, and fb200_wait blocks until it has finished and returns its result code.  One call in flight per
 * context (a second one returns FB200_ERR_ARG); every buffer and result pointer must stay valid until fb200_wait
 * returns.  Two contexts on one GPU overlap a deflate call with an inflate call: the H2D copy of one travels
 * beside the D2H copy of the other (PCIe is full duplex) and the kernels of both share the SMs. */
int fb200_deflate_segments_async(fb200_ctx *ctx, const uint8_t *src, uint64_t n, uint64_t seg_size,
                                 uint8_t *dst, uint64_t dst_cap, uint64_t *seg_off, uint64_t *out_len);
int fb200_deflate_streams_async(fb200_ctx *ctx, const uint8_t *src, const uint64_t *src_off, uint64_t nstreams,
                                uint8_t *dst, uint64_t dst_cap, uint64_t *dst_off, uint64_t *out_len);
int fb200_inflate_batch_async(fb200_ctx *ctx, const uint8_t *comp, const uint64_t *comp_off, uint64_t nstreams,
                              uint8_t *out, const uint64_t *out_off, uint64_t *out_len, int32_t *status,
                              int64_t *err_off, uint64_t *consumed);
int fb200_wait(fb200_ctx *ctx);

/* ------------------------------------------------------------------ */
/* Multi-GPU framing helpers (one process per GPU).  The
 * reference defines no container; this frame is an addition:
 *   magic "FB2\0" u32 | seg_size u32 | nseg u64 | comp_size u32[nseg] | streams */
#define FB200_FRAME_MAGIC 0x00324246u
uint64_t fb200_frame_header_bytes(uint64_t nseg);
/* Frame assembly over NVLink without a staging copy and without SM time: the
 * assembling rank allocates the frame buffer on its GPU and exports it with
 * CUDA IPC; every other rank of the box maps it and, once the all-gathered
 * segment sizes have given it its payload offset, puts its compacted streams
 * straight into the frame with an asynchronous peer copy (copy engines), which
 * runs beside whatever the context's compute stream does next.  handle = the
 * 64 opaque bytes of a cudaIpcMemHandle_t, to be sent to the peers by the
 * caller (torch.distributed / MPI / a pipe). */
#define FB200_IPC_HANDLE_BYTES 64
int fb200_mg_frame_alloc(fb200_ctx *ctx, uint64_t bytes, void **d_frame, uint8_t *handle);
int fb200_mg_frame_open(fb200_ctx *ctx, const uint8_t *handle, void **d_frame);
/* owner != 0: the pointer came from fb200_mg_frame_alloc (freed); else from _open (unmapped). */
int fb200_mg_frame_close(fb200_ctx *ctx, void *d_frame, int owner);
/* Asynchronous copy of n bytes from d_payload (this context's GPU) to d_frame + offset
 * (local or IPC-mapped), ordered after everything queued so far on the context stream. */
int fb200_mg_put(fb200_ctx *ctx, void *d_frame, uint64_t offset, const void *d_payload, uint64_t n);
/* Blocks until every fb200_mg_put / fb200_mg_get_async of this context has landed. */
int fb200_mg_wait(fb200_ctx *ctx);
/* Frame reader (decompress side, SURVEY.md 8e: "scatter compressed ranges, decode locally"): fetches the
 * compressed streams of segments [first, first + count) from a frame of frame_bytes bytes at d_frame -- on this
 * GPU, or on the assembling GPU and mapped with fb200_mg_frame_open (then the payload crosses NVLink as one peer
 * copy) -- into d_comp (device, capacity comp_cap) and writes their offsets, count + 1 values starting at 0, to
 * the device array d_comp_off: ready for fb200_inflate_batch_dev.  *seg_size / *nseg_total (may be NULL) receive
 * the header fields, *out_bytes the size of the fetched range.  Returns when the range has arrived. */
int fb200_mg_get(fb200_ctx *ctx, const void *d_frame, uint64_t frame_bytes, uint64_t first, uint64_t count,
                 uint8_t *d_comp, uint64_t comp_cap, uint64_t *d_comp_off, uint32_t *seg_size, uint64_t *nseg_total,
                 uint64_t *out_bytes);
/* The same, but returns once the offsets are in d_comp_off and the payload copy is queued (the header fields and
 * *out_bytes are final on return); fb200_mg_wait blocks until the payload has arrived.  Lets the caller run other
 * work of the same context (the deflate of the next batch) beside the transfer. */
int fb200_mg_get_async(fb200_ctx *ctx, const void *d_frame, uint64_t frame_bytes, uint64_t first, uint64_t count,
                       uint8_t *d_comp, uint64_t comp_cap, uint64_t *d_comp_off, uint32_t *seg_size,
                       uint64_t *nseg_total, uint64_t *out_bytes);

/* ------------------------------------------------------------------ */
/* Streaming objects mirroring the reference API (host buffers).       */
typedef struct fb200_writer fb200_writer;
/* sink is called with each run of compressed bytes (the &@io.Writer passed to
 * Writer::new, writer.mbt:10); return 0 on success, non-zero = sticky I/O error.
 * Like Compressor::write (deflate.mbt:280-294), a write compresses every 65535-byte
 * window it has completed and hands the bytes to the sink before it returns: at most
 * one window stays buffered, whatever the length of the stream.  sink == NULL selects
 * the pull form for hosts that cannot pass a callback: the bytes queue up inside the
 * object and are collected with fb200_writer_take after each write / close. */
typedef int (*fb200_sink_fn)(void *user, const uint8_t *data, uint64_t n);
fb200_writer *fb200_writer_new(fb200_ctx *ctx, fb200_sink_fn sink, void *user);
/* pull form: compressed bytes waiting / moves up to cap of them to dst, returns the count */
uint64_t fb200_writer_pending(const fb200_writer *w);
uint64_t fb200_writer_take(fb200_writer *w, uint8_t *dst, uint64_t cap);
/* Writer::new_dict (writer.mbt:25-31): the dictionary is compressed into the
 * output as if it had been written first (reference quirk, deflate_test.mbt:12-35). */
fb200_writer *fb200_writer_new_dict(fb200_ctx *ctx, fb200_sink_fn sink, void *user, const uint8_t *dict,
                                    uint64_t n);
/* impl @io.Writer for Writer (writer.mbt:45): returns bytes accepted (== n) or a
 * negative FB200_ERR_* (FB200_ERR_CLOSED after close). */
int64_t fb200_writer_write(fb200_writer *w, const uint8_t *data, uint64_t n);
/* impl @io.Closer for Writer (writer.mbt:53); a second close returns FB200_OK. */
int fb200_writer_close(fb200_writer *w);
void fb200_writer_free(fb200_writer *w);

typedef struct fb200_reader fb200_reader;
/* &Reader::new (inflate.mbt:305): comp must stay valid until the first read. */
fb200_reader *fb200_reader_new(fb200_ctx *ctx, const uint8_t *comp, uint64_t n);
/* &Reader::new_dict (inflate.mbt:310-317): the stream is decoded as if the uncompressed data started with
 * dict, which has already been read (only its last 32768 bytes matter, dict-decoder.mbt:49-52). */
fb200_reader *fb200_reader_new_dict(fb200_ctx *ctx, const uint8_t *comp, uint64_t n, const uint8_t *dict,
                                    uint64_t dict_len);
/* Decompressor::reset(r, dict) / make_reader (inflate.mbt:857-883): a new input (and dictionary, may be
 * NULL / 0) for an existing object; everything buffered is dropped. */
int fb200_reader_reset(fb200_reader *r, const uint8_t *comp, uint64_t n, const uint8_t *dict, uint64_t dict_len);
/* impl @io.Reader for Decompressor (inflate.mbt:382-407).  Returns the byte
 * count; *status = -1 while the reference returns (n, None), otherwise the
 * FB200_ST_* code delivered together with the last bytes. */
uint64_t fb200_reader_read(fb200_reader *r, uint8_t *buf, uint64_t n, int32_t *status, int64_t *err_off);
/* Input bytes the decoder consumed (the reference pulls its input byte by byte and stops right behind the final
 * block, inflate.mbt:789-799): whatever follows comp[consumed] -- a gzip / zlib trailer, the next member --
 * belongs to the caller.  Decodes the stream if no read has done so yet. */
uint64_t fb200_reader_consumed(fb200_reader *r);
/* impl @io.Closer for Decompressor (inflate.mbt:410-415): FB200_ST_EOF* -> 0. */
int fb200_reader_close(fb200_reader *r);
void fb200_reader_free(fb200_reader *r);

/* ------------------------------------------------------------------ */
/* Introspection used by the parity tests and the bench (device-side
 * intermediates of the last fb200_deflate_* call on this context).    */
typedef struct {
  uint64_t nblocks;      /* blocks over all streams (final empty stored block excluded) */
  uint64_t ntokens;      /* tokens of all parsed blocks                                 */
  uint64_t kernel_launches; /* kernels launched by the last call                        */
  uint64_t inflate_fallbacks; /* streams of the last inflate call re-decoded by the exact kernel */
} fb200_stats;
int fb200_last_stats(const fb200_ctx *ctx, fb200_stats *out);
/* The context's cudaStream_t (as void*): every kernel of this context is
 * launched on it, so callers can bracket calls with their own CUDA events. */
void *fb200_cuda_stream(const fb200_ctx *ctx);
/* Device time of each pipeline stage of the last deflate / inflate call,
 * measured with CUDA events on the context stream (milliseconds; slots below;
 * ms must have room for FB200_NUM_STAGES floats). */
#define FB200_STAGE_SETUP 0     /* block table: count, scan, fill           */
#define FB200_STAGE_PARSE 1     /* K1 lz77 parse (both instantiations)      */
#define FB200_STAGE_HISTOGRAM 2 /* K2                                       */
#define FB200_STAGE_BUILD 3     /* K3 code construction + headers           */
#define FB200_STAGE_LAYOUT 4    /* layout + output offset scan              */
#define FB200_STAGE_PACK 5      /* output clear + K4 bit pack + trailers    */
#define FB200_STAGE_INFLATE 6   /* K6                                       */
#define FB200_NUM_STAGES 7
int fb200_last_stage_ms(const fb200_ctx *ctx, float *ms);
/* Copy per-block results of the last deflate call to host arrays (each may be
 * NULL): token counts, kinds (0 stored, 1 huff-only, 2 dynamic), bit sizes;
 * tokens[] receives the concatenated token arrays (tok_cap entries max). */
int fb200_last_blocks(const fb200_ctx *ctx, uint32_t *blk_ntok, uint8_t *blk_kind, uint32_t *blk_bits,
                      uint64_t blk_cap, uint32_t *tokens, uint64_t tok_cap);

/* Test hook: 1 if block b of one stream starts with a cleared hash table (DeflateFast::encode calls
 * shift_offsets once cur has reached buffer_reset, deflate-fast.mbt:129-132); the closed form the
 * block-parallel parse uses, checked against the running sum by tests/test_abi.py. */
int fb200_debug_block_resets(uint64_t b);

/* ------------------------------------------------------------------ */
/* Environment switches read at fb200_create (everything else is compiled in):
 *   FB200_HOST_OVERLAP=0|1    host-buffer calls: overlap the H2D copy with the consuming kernel through a device
 *                             watermark (default 1; default 0 when CUDA_LAUNCH_BLOCKING is set or a profiler /
 *                             sanitizer / debugger injection variable is present: copy first, then launch)
 *   FB200_CHUNK_MB=n          size of the H2D chunks / D2H output groups of the host-buffer calls (32)
 *   FB200_GROUP_MB=n          input per K2..K4 group of the host-buffer deflate (64)
 *   FB200_DEFLATE_PIPELINE=0  host-buffer deflate: K2..K4 in one piece instead of group by group
 *   FB200_PARSE_BLOCKPAR=0|1|2  block-parallel parse of multi-block streams: never / when few streams / always
 *   FB200_PARSE_WARPS=s, FB200_PARSE_GWARPS=g  parse warps per SM with shared-memory / global-memory tables
 *   FB200_PARSE_CARVEOUT=p    shared-memory carve-out hint (percent) of the parse kernels; default: the smallest
 *                             one that holds the tables
 *   FB200_INFLATE_CTAS=c      inflate CTAs (4 warps each) per SM
 *   FB200_INFLATE_WINDOW_KB=k speculation window of a stream's first block (96; 0: the whole rest of the stream)
 *   FB200_INFLATE_CTA_STREAMS=n  calls with at most n streams inflate with one CTA per stream (296; 0: never)
 *   FB200_TRACE=1|2           inflate round / block counters (1), timeline of the host-buffer calls (2) on stderr */

#ifdef __cplusplus
}
#endif
#endif
