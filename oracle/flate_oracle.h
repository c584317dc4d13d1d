/*
 * flate_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the deflate-fast encoder and the inflate decoder of
 * gmlewis/moonbit-flate (reference files cited per function in flate_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library, and only as the checker / CPU baseline.
 * The product (libflate_b200.so) never links or calls it.
 *
 * Parity pinning: the MoonBit toolchain is absent from the build image, so the
 * reference itself cannot be run.  The restatement is pinned against every
 * known-answer the reference's own tests hold for this path (38-byte KAT of
 * deflate_test.mbt:12-35, token.mbt:95-99, bits.mbt:24-27,
 * huffman-code.mbt:289-292, the DictDecoder scenario of
 * dict-decoder_wbtest.mbt:9-291, the TestBestSpeed round-trip matrix of
 * deflate-fast_test.mbt:14-100) plus stdlib zlib as an independent inflater.
 * Compressed-byte goldens do not exist in the reference; byte parity beyond the
 * 38-byte length KAT is by source transcription.
 */
#ifndef FLATE_ORACLE_H
#define FLATE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes shared by the inflate entry points ---- */
#define ORC_OK 0              /* reached the final block: reference err == ioeof      */
#define ORC_CORRUPT 1         /* "flate: corrupt input before offset N" (N = err_off) */
#define ORC_UNEXPECTED_EOF 2  /* @io.err_unexpected_eof                               */
#define ORC_DST_TOO_SMALL 3   /* oracle-side only: caller's buffer too small          */
#define ORC_INTERNAL 4        /* "flate: internal error: ..."                         */
#define ORC_EOF_AT_REFILL 5   /* reference returns plain ioeof as well: the input ran
                                 out inside more_bits (inflate.mbt:789-799), which
                                 does not wrap the reader's eof (quirk D5)            */

/* ---- block kinds reported by the introspection API ---- */
#define ORC_BLK_STORED 0
#define ORC_BLK_HUFF 1    /* write_block_huff: literal-only dynamic block */
#define ORC_BLK_DYNAMIC 2 /* write_block_dynamic                          */

/* ------------------------------------------------------------------ */
/* Streaming Writer (writer.mbt:10-58, deflate.mbt:81-294).            */
typedef struct orc_writer orc_writer;
orc_writer *orc_writer_new(void);
/* Writer::new_dict quirk D3 (writer.mbt:25-31, deflate.mbt:108-151): the
 * dictionary is copied into the input window and compressed into the output. */
orc_writer *orc_writer_new_dict(const uint8_t *dict, size_t n);
/* returns bytes accepted, or -1 with the writer in the "writer closed" state */
int64_t orc_writer_write(orc_writer *w, const uint8_t *p, size_t n);
/* 0 = None (also for a second close, deflate.mbt:158-160) */
int orc_writer_close(orc_writer *w);
/* compressed bytes produced so far (valid until the next write/close/free) */
const uint8_t *orc_writer_data(const orc_writer *w, size_t *len);
void orc_writer_free(orc_writer *w);

/* One-shot: Writer::new + write(src) + close.  Returns the compressed size, or
 * -1 if cap is too small. */
int64_t orc_deflate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap);
/* Upper bound on orc_deflate output for n input bytes. */
size_t orc_deflate_bound(size_t n);
/* bench.py timing helper: deflate + inflate of every stride-th stream from `first`; 1 = all round-trip */
int orc_codec_pass(const uint8_t *src, const uint64_t *off, uint64_t ns, uint64_t first, uint64_t stride);

/* One-shot with per-block introspection.  tokens (may be NULL) receives the
 * concatenated token arrays of every parsed block (deflate-fast.mbt:123-270),
 * EOB not included; blk_ntok[b] is the token count of block b (0 for blocks
 * that were not parsed), blk_kind[b] one of ORC_BLK_*, blk_bits[b] the
 * number of bits the block occupies in the output (stored blocks: including
 * their byte-alignment padding).  Returns the number of blocks (the final
 * empty stored block written by close is not counted), or -1 on overflow. */
int64_t orc_deflate_ex(const uint8_t *src, size_t n, uint8_t *dst, size_t cap,
                       size_t *out_len, uint32_t *tokens, size_t tok_cap,
                       uint32_t *blk_ntok, uint8_t *blk_kind, uint64_t *blk_bits,
                       size_t blk_cap);

/* HuffmanEncoder::generate (huffman-code.mbt:295-343): code lengths and
 * bit-reversed codes for nfreq symbols. */
void orc_huff_generate(const int32_t *freq, int nfreq, int max_bits,
                       uint8_t *len_out, uint16_t *code_out);

/* ------------------------------------------------------------------ */
/* Streaming Decompressor (inflate.mbt:257-883).                       */
typedef struct orc_reader orc_reader;
orc_reader *orc_reader_new(const uint8_t *comp, size_t n);
orc_reader *orc_reader_new_dict(const uint8_t *comp, size_t n,
                                const uint8_t *dict, size_t dn);
/* impl @io.Reader for Decompressor (inflate.mbt:382-407).  Returns the byte
 * count; *status = -1 while the reference would return (n, None), else the
 * ORC_* code of the error returned alongside (ORC_OK == ioeof). */
size_t orc_reader_read(orc_reader *r, uint8_t *buf, size_t n, int *status,
                       int64_t *err_off);
/* bytes of input consumed so far (Decompressor.roffset) */
int64_t orc_reader_roffset(const orc_reader *r);
void orc_reader_free(orc_reader *r);

/* One-shot: read until an error/ioeof.  *out_len = bytes produced (partial
 * output before an error is delivered, inflate.mbt:403-405). */
int orc_inflate(const uint8_t *comp, size_t n, uint8_t *out, size_t cap,
                size_t *out_len, int64_t *err_off, int64_t *consumed);

/* ------------------------------------------------------------------ */
/* Small pieces exported for the reference's unit KATs.                */
uint32_t orc_token_offset(uint32_t tok);               /* token.mbt:90-92   */
uint32_t orc_reverse16(uint32_t x);                    /* bits.mbt:18-21    */
uint32_t orc_reverse_bits(uint32_t v, uint32_t nbits); /* huffman-code.mbt:283 */
int orc_length_code(uint32_t xlen);                    /* token.mbt:107     */
int orc_offset_code(uint32_t xoff);                    /* token.mbt:112     */
uint32_t orc_fixed_chunk(int i);                       /* inflate.mbt:886   */

/* DictDecoder (dict-decoder.mbt:29-209) for the white-box scenario. */
typedef struct orc_dict orc_dict;
orc_dict *orc_dict_new(int size, const uint8_t *dict, size_t n);
void orc_dict_free(orc_dict *d);
int orc_dict_hist_size(const orc_dict *d);
int orc_dict_avail_read(const orc_dict *d);
int orc_dict_avail_write(const orc_dict *d);
/* write_slice + slice_copy + write_mark fused: copies min(avail, n), returns it */
int orc_dict_write(orc_dict *d, const uint8_t *p, int n);
void orc_dict_write_byte(orc_dict *d, uint8_t c);
int orc_dict_write_copy(orc_dict *d, int dist, int length);
int orc_dict_try_write_copy(orc_dict *d, int dist, int length);
/* read_flush: returns count, copies the flushed bytes to out (cap >= size) */
int orc_dict_read_flush(orc_dict *d, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
