// api.cu -- C ABI of libflate_b200.so (include/flate_b200.h): context, device
// scratch, kernel sequencing, host<->device staging, and the streaming
// Writer / Decompressor objects that mirror the reference API
// (writer.mbt:10-58, inflate.mbt:257-418).
//
// No CPU fallback exists: every compute entry point runs the CUDA kernels or
// fails with FB200_ERR_CUDA.
#include "../../include/flate_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace fb;

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

} // namespace

struct fb200_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // deflate scratch
  DevBuf stream_blk0, stream_bytes, stream_trailer, dst_off_own, blk_stream, blk_ntok, blk_kind, blk_bits,
      blk_bit_start, blk_hdr_nbits, blk_hdr, blk_freq, blk_code, tokens, counters;
  // staging for the host-buffer entry points: two slots, so that the H2D copy of chunk c+1 and the
  // D2H copy of chunk c-1 overlap the kernels of chunk c (copy engines run beside the SMs)
  DevBuf h_src_off; // fixed-size segment offsets of the *_segments_dev entry point
  DevBuf p_in[2], p_out[2], p_off_in[2], p_off_out[2], p_off_abs[2], p_len[2], p_status[2], p_eoff[2], p_cons[2];
  DevBuf all_off, all_off2;
  DevBuf i_fallback, i_rec_off, i_nrec, i_records, i_order, i_order_hist;
  bool inflate_v1 = true; // FB200_INFLATE_V1=0 selects the experimental thread-per-stream path (inflate2.cu)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t e_in[2] = {}, e_comp[2] = {}, e_out[2] = {};
  uint64_t chunk_bytes = 128ull << 20;
  uint64_t *pinned = nullptr; // small pinned read-back area
  // last deflate job (for introspection)
  DeflateJob last{};
  uint64_t last_n_total = 0;
  fb200_stats stats{};
  // stage timing: event pair per stage, on the context stream
  cudaEvent_t ev0[FB200_NUM_STAGES] = {}, ev1[FB200_NUM_STAGES] = {};
  bool ev_used[FB200_NUM_STAGES] = {};
  void stage_begin(int s) { cudaEventRecord(ev0[s], stream); }
  void stage_end(int s) { cudaEventRecord(ev1[s], stream); ev_used[s] = true; }
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      char buf__[256];                                                                        \
      snprintf(buf__, sizeof buf__, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      ctx->err = buf__;                                                                       \
      cudaGetLastError();                                                                     \
      return FB200_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

extern "C" int fb200_version(void) { return FB200_VERSION; }

extern "C" int fb200_create(fb200_ctx **out, int device)
{
  if (!out) return FB200_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return FB200_ERR_CUDA; // no CPU fallback
  }
  fb200_ctx *ctx = new (std::nothrow) fb200_ctx();
  if (!ctx) return FB200_ERR_NOMEM;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) { delete ctx; return FB200_ERR_CUDA; }
  }
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    cudaGetLastError();
    return FB200_ERR_CUDA;
  }
  if (prop.major < 10) { // built for sm_100a only
    delete ctx;
    return FB200_ERR_CUDA;
  }
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost((void **)&ctx->pinned, 64 * sizeof(uint64_t)) != cudaSuccess) {
    delete ctx;
    cudaGetLastError();
    return FB200_ERR_CUDA;
  }
  for (int i = 0; i < FB200_NUM_STAGES; i++) {
    cudaEventCreate(&ctx->ev0[i]);
    cudaEventCreate(&ctx->ev1[i]);
  }
  cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
  for (int i = 0; i < 2; i++) {
    cudaEventCreateWithFlags(&ctx->e_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->e_comp[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->e_out[i], cudaEventDisableTiming);
  }
  if (const char *e = getenv("FB200_INFLATE_V1")) ctx->inflate_v1 = atoi(e) != 0;
  if (const char *e = getenv("FB200_CHUNK_MB")) {
    const long mb = atol(e);
    if (mb > 0) ctx->chunk_bytes = (uint64_t)mb << 20;
  }
  launch_init_tables(ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    delete ctx;
    cudaGetLastError();
    return FB200_ERR_CUDA;
  }
  *out = ctx;
  return FB200_OK;
}

extern "C" void fb200_destroy(fb200_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  DevBuf *all[] = {&ctx->stream_blk0, &ctx->stream_bytes, &ctx->stream_trailer, &ctx->dst_off_own, &ctx->blk_stream,
                   &ctx->blk_ntok, &ctx->blk_kind, &ctx->blk_bits, &ctx->blk_bit_start, &ctx->blk_hdr_nbits,
                   &ctx->blk_hdr, &ctx->blk_freq, &ctx->blk_code, &ctx->tokens, &ctx->counters, &ctx->h_src_off,
                   &ctx->all_off, &ctx->all_off2, &ctx->i_fallback, &ctx->i_rec_off, &ctx->i_nrec, &ctx->i_records, &ctx->i_order, &ctx->i_order_hist};
  for (int i = 0; i < 2; i++) {
    DevBuf *slot[] = {&ctx->p_in[i], &ctx->p_out[i], &ctx->p_off_in[i], &ctx->p_off_out[i], &ctx->p_off_abs[i],
                      &ctx->p_len[i], &ctx->p_status[i], &ctx->p_eoff[i], &ctx->p_cons[i]};
    for (DevBuf *b : slot) b->release();
    if (ctx->e_in[i]) cudaEventDestroy(ctx->e_in[i]);
    if (ctx->e_comp[i]) cudaEventDestroy(ctx->e_comp[i]);
    if (ctx->e_out[i]) cudaEventDestroy(ctx->e_out[i]);
  }
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  for (DevBuf *b : all) b->release();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int i = 0; i < FB200_NUM_STAGES; i++) {
    if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]);
    if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]);
  }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *fb200_last_error(const fb200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" uint64_t fb200_deflate_stream_bound(uint64_t n)
{
  // every code <= 15 bits per literal byte, a match token (>= 4 bytes) <= 48 bits,
  // block header < 640 B, one block per 65535 bytes, 5-byte final block
  const uint64_t nblk = n / kBlockSize + 1;
  return 2 * n + 640 * nblk + 16;
}

extern "C" uint64_t fb200_deflate_bound(uint64_t n, uint64_t seg_size)
{
  if (seg_size == 0) return 0;
  const uint64_t nseg = (n + seg_size - 1) / seg_size;
  return nseg * fb200_deflate_stream_bound(seg_size < n ? seg_size : n) + 16;
}

extern "C" uint64_t fb200_frame_header_bytes(uint64_t nseg) { return 16 + 4 * nseg; }

// ------------------------------------------------------------------
// deflate core: phase A = everything up to the output layout (returns the
// total compressed size), phase B = bit packing into d_dst.

static int deflate_phase_a(fb200_ctx *ctx, const uint8_t *d_src, const uint64_t *d_src_off, uint64_t ns,
                           uint64_t n_total, uint64_t *d_dst_off, uint64_t *total_out)
{
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ctx->stats = fb200_stats{};
  if (ns > 0xfffffff0ull) { ctx->err = "too many streams"; return FB200_ERR_ARG; }
  DeflateJob &j = ctx->last;
  j = DeflateJob{};
  j.src = d_src;
  j.stream_off = d_src_off;
  j.nstreams = ns;
  CK(ctx->stream_blk0.ensure((ns + 1) * 8));
  CK(ctx->stream_bytes.ensure((ns + 1) * 8));
  CK(ctx->stream_trailer.ensure((ns + 1) * 8));
  CK(ctx->counters.ensure(64));
  j.stream_blk0 = ctx->stream_blk0.as<uint64_t>();
  j.stream_bytes = ctx->stream_bytes.as<uint64_t>();
  j.stream_trailer_bit = ctx->stream_trailer.as<uint64_t>();
  j.dst_off = d_dst_off;
  j.counters = ctx->counters.as<uint32_t>();
  CK(cudaMemsetAsync(j.counters, 0, 64, st));
  uint64_t launches = 0;
  for (int i = 0; i < FB200_NUM_STAGES; i++) ctx->ev_used[i] = false;

  ctx->stage_begin(FB200_STAGE_SETUP);
  launch_count_blocks(j, st);
  launch_scan_u64(j.stream_blk0, j.stream_blk0, ns, st);
  launches += 2;
  CK(cudaMemcpyAsync(ctx->pinned, j.stream_blk0 + ns, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint64_t nb = ctx->pinned[0];
  if (nb > 0x7fffffffull) { ctx->err = "too many blocks in one call"; return FB200_ERR_ARG; }
  j.nblocks = nb;
  const uint64_t nbp = nb + 1;
  CK(ctx->blk_stream.ensure(nbp * 4));
  CK(ctx->blk_ntok.ensure(nbp * 4));
  CK(ctx->blk_kind.ensure(nbp));
  CK(ctx->blk_bits.ensure(nbp * 4));
  CK(ctx->blk_bit_start.ensure(nbp * 8));
  CK(ctx->blk_hdr_nbits.ensure(nbp * 4));
  CK(ctx->blk_hdr.ensure(nbp * kHdrWords * 4));
  CK(ctx->blk_freq.ensure(nbp * kFreqStride * 4));
  CK(ctx->blk_code.ensure(nbp * kFreqStride * 4));
  CK(ctx->tokens.ensure((n_total + 16) * 4));
  j.blk_stream = ctx->blk_stream.as<uint32_t>();
  j.blk_ntok = ctx->blk_ntok.as<uint32_t>();
  j.blk_kind = ctx->blk_kind.as<uint8_t>();
  j.blk_bits = ctx->blk_bits.as<uint32_t>();
  j.blk_bit_start = ctx->blk_bit_start.as<uint64_t>();
  j.blk_hdr_nbits = ctx->blk_hdr_nbits.as<uint32_t>();
  j.blk_hdr = ctx->blk_hdr.as<uint32_t>();
  j.blk_freq = ctx->blk_freq.as<uint32_t>();
  j.blk_code = ctx->blk_code.as<uint32_t>();
  j.tokens = ctx->tokens.as<uint32_t>();
  ctx->last_n_total = n_total;
  CK(cudaMemsetAsync(j.blk_ntok, 0, nbp * 4, st));
  CK(cudaMemsetAsync(j.blk_bits, 0, nbp * 4, st));

  launch_fill_blocks(j, st);
  ctx->stage_end(FB200_STAGE_SETUP);
  ctx->stage_begin(FB200_STAGE_PARSE);
  launch_parse(j, ctx->num_sms, st);
  ctx->stage_end(FB200_STAGE_PARSE);
  ctx->stage_begin(FB200_STAGE_HISTOGRAM);
  launch_histogram(j, st);
  ctx->stage_end(FB200_STAGE_HISTOGRAM);
  ctx->stage_begin(FB200_STAGE_BUILD);
  launch_build_codes(j, ctx->num_sms, st);
  ctx->stage_end(FB200_STAGE_BUILD);
  ctx->stage_begin(FB200_STAGE_LAYOUT);
  launch_layout(j, st);
  launch_scan_u64(j.stream_bytes, j.dst_off, ns, st);
  ctx->stage_end(FB200_STAGE_LAYOUT);
  launches += 7;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->pinned, j.dst_off + ns, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *total_out = ctx->pinned[0];
  ctx->stats.nblocks = nb;
  ctx->stats.kernel_launches = launches;
  return FB200_OK;
}

static int deflate_phase_b(fb200_ctx *ctx, uint8_t *d_dst, uint64_t total)
{
  cudaStream_t st = ctx->stream;
  DeflateJob &j = ctx->last;
  if ((reinterpret_cast<uintptr_t>(d_dst) & 3) != 0) { ctx->err = "device dst must be 4-byte aligned"; return FB200_ERR_ARG; }
  j.dst = d_dst;
  ctx->stage_begin(FB200_STAGE_PACK);
  CK(cudaMemsetAsync(d_dst, 0, (total + 3) & ~3ull, st));
  launch_pack(j, st);
  ctx->stage_end(FB200_STAGE_PACK);
  ctx->stats.kernel_launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->pinned, j.counters, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint32_t *c = reinterpret_cast<const uint32_t *>(ctx->pinned);
  if (c[4] != 0) {
    ctx->err = "internal error: packed block size differs from the computed layout";
    return FB200_ERR_CUDA;
  }
  return FB200_OK;
}

extern "C" int fb200_deflate_streams_dev(fb200_ctx *ctx, const uint8_t *d_src, const uint64_t *d_src_off,
                                         uint64_t nstreams, uint64_t n_total, uint8_t *d_dst, uint64_t dst_cap,
                                         uint64_t *d_dst_off, uint64_t *out_len)
{
  if (!ctx || !d_src_off || !d_dst_off || !out_len || (!d_src && n_total) || !d_dst) return FB200_ERR_ARG;
  uint64_t total = 0;
  int rc = deflate_phase_a(ctx, d_src, d_src_off, nstreams, n_total, d_dst_off, &total);
  if (rc != FB200_OK) return rc;
  *out_len = total;
  if (((total + 3) & ~3ull) > dst_cap) { ctx->err = "dst_cap too small"; return FB200_ERR_DST_TOO_SMALL; }
  return deflate_phase_b(ctx, d_dst, total);
}

extern "C" int fb200_deflate_segments_dev(fb200_ctx *ctx, const uint8_t *d_src, uint64_t n, uint64_t seg_size,
                                          uint8_t *d_dst, uint64_t dst_cap, uint64_t *d_seg_off, uint64_t *out_len)
{
  if (!ctx || seg_size == 0 || !d_seg_off) return FB200_ERR_ARG;
  const uint64_t nseg = (n + seg_size - 1) / seg_size;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->h_src_off.ensure((nseg + 1) * 8));
  launch_fill_seg_off(ctx->h_src_off.as<uint64_t>(), nseg, seg_size, n, ctx->stream);
  int rc = fb200_deflate_streams_dev(ctx, d_src, ctx->h_src_off.as<uint64_t>(), nseg, n, d_dst, dst_cap, d_seg_off,
                                     out_len);
  ctx->stats.kernel_launches += 1;
  return rc;
}

// Host-buffer deflate: the batch is cut into chunks of whole streams (~chunk_bytes each) that flow
// through a two-slot pipeline -- H2D of chunk c+1 (stream s_in) and D2H of chunk c-1 (s_out) run
// beside the kernels of chunk c (ctx->stream).  Streams are independent, so chunking never changes
// a byte; offsets are re-based per chunk on the device.
struct HostChunk {
  uint64_t a, b;        // streams [a, b)
  uint64_t byte0, bytes; // source byte range
};

static int deflate_host_common(fb200_ctx *ctx, const uint8_t *src, uint64_t n, const uint64_t *src_off, uint64_t ns,
                               uint64_t seg_size, uint8_t *dst, uint64_t dst_cap, uint64_t *dst_off, uint64_t *out_len)
{
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  *out_len = 0;
  if (ns == 0) {
    if (dst_off) dst_off[0] = 0;
    ctx->stats = fb200_stats{};
    ctx->last = DeflateJob{};
    return FB200_OK;
  }
  std::vector<HostChunk> chunks;
  {
    uint64_t a = 0;
    while (a < ns) {
      uint64_t b = a;
      const uint64_t byte0 = src_off ? src_off[a] : a * seg_size;
      uint64_t bytes = 0;
      if (src_off) {
        while (b < ns && (b == a || bytes < ctx->chunk_bytes)) { b++; bytes = src_off[b] - byte0; }
      } else {
        uint64_t cnt = ctx->chunk_bytes / seg_size;
        if (cnt == 0) cnt = 1;
        b = a + cnt < ns ? a + cnt : ns;
        const uint64_t end = b * seg_size < n ? b * seg_size : n;
        bytes = end - byte0;
      }
      chunks.push_back({a, b, byte0, bytes});
      a = b;
    }
  }
  uint64_t max_bytes = 0, max_cnt = 0;
  for (const HostChunk &c : chunks) {
    if (c.bytes > max_bytes) max_bytes = c.bytes;
    if (c.b - c.a > max_cnt) max_cnt = c.b - c.a;
  }
  for (int i = 0; i < 2; i++) {
    CK(ctx->p_in[i].ensure(max_bytes + 16));
    CK(ctx->p_off_in[i].ensure((max_cnt + 1) * 8));
    CK(ctx->p_off_out[i].ensure((max_cnt + 1) * 8));
    CK(ctx->p_off_abs[i].ensure((max_cnt + 1) * 8));
    if (chunks.size() > 1) CK(ctx->p_out[i].ensure(max_bytes / 2 + max_bytes / 8 + (1u << 20))); // grows if needed
  }
  if (src_off) {
    CK(ctx->all_off.ensure((ns + 1) * 8));
    CK(cudaMemcpyAsync(ctx->all_off.p, src_off, (ns + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
  }
  auto issue_h2d = [&](size_t c) -> int {
    const HostChunk &ch = chunks[c];
    const int slot = (int)(c & 1);
    if (c >= 2) CK(cudaStreamWaitEvent(ctx->s_in, ctx->e_comp[slot], 0));
    if (ch.bytes) CK(cudaMemcpyAsync(ctx->p_in[slot].p, src + ch.byte0, ch.bytes, cudaMemcpyHostToDevice, ctx->s_in));
    if (src_off)
      launch_affine_u64(ctx->p_off_in[slot].as<uint64_t>(), ctx->all_off.as<uint64_t>() + ch.a, ch.b - ch.a + 1,
                        0ull - ch.byte0, ctx->s_in);
    else
      launch_fill_seg_off(ctx->p_off_in[slot].as<uint64_t>(), ch.b - ch.a, seg_size, ch.bytes, ctx->s_in);
    CK(cudaEventRecord(ctx->e_in[slot], ctx->s_in));
    return FB200_OK;
  };
  int rc = issue_h2d(0);
  if (rc != FB200_OK) return rc;
  uint64_t base = 0, nblocks = 0, launches = 0;
  bool overflow = false;
  for (size_t c = 0; c < chunks.size(); c++) {
    const HostChunk &ch = chunks[c];
    const int slot = (int)(c & 1);
    if (c + 1 < chunks.size() && (rc = issue_h2d(c + 1)) != FB200_OK) return rc;
    CK(cudaStreamWaitEvent(st, ctx->e_in[slot], 0));
    uint64_t total = 0;
    rc = deflate_phase_a(ctx, ctx->p_in[slot].as<uint8_t>(), ctx->p_off_in[slot].as<uint64_t>(), ch.b - ch.a, ch.bytes,
                         ctx->p_off_out[slot].as<uint64_t>(), &total);
    if (rc != FB200_OK) return rc;
    nblocks += ctx->stats.nblocks;
    launches += ctx->stats.kernel_launches + 2;
    if (base + total > dst_cap) overflow = true;
    if (!overflow) {
      if (c >= 2) CK(cudaStreamWaitEvent(st, ctx->e_out[slot], 0));
      CK(ctx->p_out[slot].ensure(total + 16));
      rc = deflate_phase_b(ctx, ctx->p_out[slot].as<uint8_t>(), total);
      if (rc != FB200_OK) return rc;
      launches += 2;
      if (dst_off)
        launch_affine_u64(ctx->p_off_abs[slot].as<uint64_t>(), ctx->p_off_out[slot].as<uint64_t>(), ch.b - ch.a + 1,
                          base, st);
      CK(cudaEventRecord(ctx->e_comp[slot], st));
      CK(cudaStreamWaitEvent(ctx->s_out, ctx->e_comp[slot], 0));
      if (total) CK(cudaMemcpyAsync(dst + base, ctx->p_out[slot].p, total, cudaMemcpyDeviceToHost, ctx->s_out));
      if (dst_off)
        CK(cudaMemcpyAsync(dst_off + ch.a, ctx->p_off_abs[slot].p, (ch.b - ch.a + 1) * 8, cudaMemcpyDeviceToHost,
                           ctx->s_out));
      CK(cudaEventRecord(ctx->e_out[slot], ctx->s_out));
    } else {
      CK(cudaEventRecord(ctx->e_comp[slot], st));
    }
    base += total;
  }
  CK(cudaStreamSynchronize(ctx->s_in));
  CK(cudaStreamSynchronize(ctx->s_out));
  CK(cudaStreamSynchronize(st));
  *out_len = base; // on overflow: the capacity the call needs
  ctx->stats.nblocks = nblocks;
  ctx->stats.kernel_launches = launches;
  if (overflow) { ctx->err = "dst_cap too small"; return FB200_ERR_DST_TOO_SMALL; }
  return FB200_OK;
}

extern "C" int fb200_deflate_segments(fb200_ctx *ctx, const uint8_t *src, uint64_t n, uint64_t seg_size, uint8_t *dst,
                                      uint64_t dst_cap, uint64_t *seg_off, uint64_t *out_len)
{
  if (!ctx || seg_size == 0 || !out_len || (!src && n) || !dst) return FB200_ERR_ARG;
  const uint64_t nseg = (n + seg_size - 1) / seg_size;
  return deflate_host_common(ctx, src, n, nullptr, nseg, seg_size, dst, dst_cap, seg_off, out_len);
}

extern "C" int fb200_deflate_streams(fb200_ctx *ctx, const uint8_t *src, const uint64_t *src_off, uint64_t nstreams,
                                     uint8_t *dst, uint64_t dst_cap, uint64_t *dst_off, uint64_t *out_len)
{
  if (!ctx || !src_off || !out_len || !dst) return FB200_ERR_ARG;
  for (uint64_t i = 0; i < nstreams; i++)
    if (src_off[i + 1] < src_off[i]) { ctx->err = "src_off not monotone"; return FB200_ERR_ARG; }
  if (nstreams && src_off[0] != 0) { ctx->err = "src_off[0] must be 0"; return FB200_ERR_ARG; }
  const uint64_t n = nstreams ? src_off[nstreams] : 0;
  if (!src && n) return FB200_ERR_ARG;
  return deflate_host_common(ctx, src, n, src_off, nstreams, 0, dst, dst_cap, dst_off, out_len);
}

extern "C" int fb200_last_stats(const fb200_ctx *ctx, fb200_stats *out)
{
  if (!ctx || !out) return FB200_ERR_ARG;
  *out = ctx->stats;
  return FB200_OK;
}

extern "C" void *fb200_cuda_stream(const fb200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int fb200_last_stage_ms(const fb200_ctx *cctx, float *ms)
{
  fb200_ctx *ctx = const_cast<fb200_ctx *>(cctx);
  if (!ctx || !ms) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < FB200_NUM_STAGES; i++) {
    ms[i] = 0.f;
    if (ctx->ev_used[i]) CK(cudaEventElapsedTime(&ms[i], ctx->ev0[i], ctx->ev1[i]));
  }
  return FB200_OK;
}

extern "C" int fb200_last_blocks(const fb200_ctx *cctx, uint32_t *blk_ntok, uint8_t *blk_kind, uint32_t *blk_bits,
                                 uint64_t blk_cap, uint32_t *tokens, uint64_t tok_cap)
{
  fb200_ctx *ctx = const_cast<fb200_ctx *>(cctx);
  if (!ctx) return FB200_ERR_ARG;
  const DeflateJob &j = ctx->last;
  const uint64_t nb = j.nblocks;
  if (nb > blk_cap) return FB200_ERR_DST_TOO_SMALL;
  CK(cudaSetDevice(ctx->device));
  std::vector<uint32_t> ntok(nb), stream(nb);
  std::vector<uint64_t> blk0(j.nstreams + 1), soff(j.nstreams + 1);
  if (nb) {
    CK(cudaMemcpy(ntok.data(), j.blk_ntok, nb * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(stream.data(), j.blk_stream, nb * 4, cudaMemcpyDeviceToHost));
    if (blk_ntok) memcpy(blk_ntok, ntok.data(), nb * 4);
    if (blk_kind) CK(cudaMemcpy(blk_kind, j.blk_kind, nb, cudaMemcpyDeviceToHost));
    if (blk_bits) CK(cudaMemcpy(blk_bits, j.blk_bits, nb * 4, cudaMemcpyDeviceToHost));
  }
  if (tokens) {
    CK(cudaMemcpy(blk0.data(), j.stream_blk0, (j.nstreams + 1) * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(soff.data(), j.stream_off, (j.nstreams + 1) * 8, cudaMemcpyDeviceToHost));
    uint64_t o = 0;
    for (uint64_t b = 0; b < nb; b++) {
      if (!ntok[b]) continue;
      if (o + ntok[b] > tok_cap) return FB200_ERR_DST_TOO_SMALL;
      const uint32_t s = stream[b];
      const uint64_t src_off = soff[s] + (b - blk0[s]) * (uint64_t)kBlockSize;
      CK(cudaMemcpy(tokens + o, j.tokens + src_off, (size_t)ntok[b] * 4, cudaMemcpyDeviceToHost));
      o += ntok[b];
    }
  }
  return FB200_OK;
}

// ------------------------------------------------------------------
// inflate

extern "C" int fb200_inflate_batch_dev(fb200_ctx *ctx, const uint8_t *d_comp, const uint64_t *d_comp_off,
                                       uint64_t nstreams, uint8_t *d_out, const uint64_t *d_out_off,
                                       uint64_t *d_out_len, int32_t *d_status, int64_t *d_err_off,
                                       uint64_t *d_consumed)
{
  if (!ctx || !d_comp_off || !d_out_off || !d_out_len || !d_status || !d_err_off) return FB200_ERR_ARG;
  if (nstreams > 0xfffffff0ull) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CK(ctx->counters.ensure(64));
  CK(ctx->i_fallback.ensure((nstreams + 1) * 4));
  CK(cudaMemsetAsync(ctx->counters.p, 0, 64, st));
  InflateJob j{};
  j.comp = d_comp;
  j.comp_off = d_comp_off;
  j.nstreams = nstreams;
  j.out = d_out;
  j.out_off = d_out_off;
  j.out_len = d_out_len;
  j.status = d_status;
  j.err_off = d_err_off;
  j.consumed = d_consumed;
  j.counters = ctx->counters.as<uint32_t>();
  j.fallback = ctx->i_fallback.as<uint32_t>();
  for (int i = 0; i < FB200_NUM_STAGES; i++) ctx->ev_used[i] = false;
  uint64_t launches = nstreams ? 2 : 0;
  if (!ctx->inflate_v1 && nstreams) {
    // record areas: sized from the output capacity (two u64 read back; a match yields >= 3 bytes)
    CK(cudaMemcpyAsync(ctx->pinned, d_out_off, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->pinned + 1, d_out_off + nstreams, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t cap_total = ctx->pinned[1] - ctx->pinned[0];
    CK(ctx->i_rec_off.ensure((nstreams + 1) * 8));
    CK(ctx->i_nrec.ensure((nstreams + 1) * 4));
    CK(ctx->i_records.ensure((cap_total / 3 + 4 * nstreams + 8) * 8));
    j.rec_off = ctx->i_rec_off.as<uint64_t>();
    j.nrec = ctx->i_nrec.as<uint32_t>();
    j.records = ctx->i_records.as<uint2>();
    CK(ctx->i_order.ensure((nstreams + 1) * 4));
    CK(ctx->i_order_hist.ensure(1024 * 4));
    j.order = ctx->i_order.as<uint32_t>();
    launch_rec_off(d_out_off, ctx->i_rec_off.as<uint64_t>(), nstreams, st);
    launches += 5;
  }
  ctx->stage_begin(FB200_STAGE_INFLATE);
  if (!ctx->inflate_v1) launch_inflate2(j, ctx->num_sms, ctx->i_order_hist.as<uint32_t>(), st);
  launch_inflate(j, ctx->num_sms, ctx->inflate_v1, st);
  ctx->stage_end(FB200_STAGE_INFLATE);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->pinned, j.counters, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->stats = fb200_stats{};
  ctx->stats.kernel_launches = launches;
  ctx->stats.inflate_fallbacks = reinterpret_cast<const uint32_t *>(ctx->pinned)[2];
  return FB200_OK;
}

extern "C" int fb200_inflate_batch(fb200_ctx *ctx, const uint8_t *comp, const uint64_t *comp_off, uint64_t nstreams,
                                   uint8_t *out, const uint64_t *out_off, uint64_t *out_len, int32_t *status,
                                   int64_t *err_off, uint64_t *consumed)
{
  if (!ctx || !comp_off || !out_off || !out_len || !status || !err_off) return FB200_ERR_ARG;
  for (uint64_t i = 0; i < nstreams; i++)
    if (comp_off[i + 1] < comp_off[i] || out_off[i + 1] < out_off[i]) { ctx->err = "offsets not monotone"; return FB200_ERR_ARG; }
  const uint64_t nc = nstreams ? comp_off[nstreams] - comp_off[0] : 0;
  const uint64_t no = nstreams ? out_off[nstreams] - out_off[0] : 0;
  if ((!comp && nc) || (!out && no)) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->stats = fb200_stats{};
  if (nstreams == 0) return FB200_OK;
  cudaStream_t st = ctx->stream;
  // chunks of whole streams, two-slot pipeline as in deflate_host_common
  struct Chunk { uint64_t a, b; };
  std::vector<Chunk> chunks;
  uint64_t max_c = 0, max_o = 0, max_cnt = 0;
  for (uint64_t a = 0; a < nstreams;) {
    uint64_t b = a;
    while (b < nstreams && (b == a || (out_off[b] - out_off[a] < ctx->chunk_bytes && comp_off[b] - comp_off[a] < ctx->chunk_bytes)))
      b++;
    chunks.push_back({a, b});
    if (comp_off[b] - comp_off[a] > max_c) max_c = comp_off[b] - comp_off[a];
    if (out_off[b] - out_off[a] > max_o) max_o = out_off[b] - out_off[a];
    if (b - a > max_cnt) max_cnt = b - a;
    a = b;
  }
  for (int i = 0; i < 2; i++) {
    CK(ctx->p_in[i].ensure(max_c + 16));
    CK(ctx->p_out[i].ensure(max_o + 16));
    CK(ctx->p_off_in[i].ensure((max_cnt + 1) * 8));
    CK(ctx->p_off_out[i].ensure((max_cnt + 1) * 8));
    CK(ctx->p_len[i].ensure((max_cnt + 1) * 8));
    CK(ctx->p_status[i].ensure((max_cnt + 1) * 4));
    CK(ctx->p_eoff[i].ensure((max_cnt + 1) * 8));
    CK(ctx->p_cons[i].ensure((max_cnt + 1) * 8));
  }
  CK(ctx->all_off.ensure((nstreams + 1) * 8));
  CK(ctx->all_off2.ensure((nstreams + 1) * 8));
  CK(cudaMemcpyAsync(ctx->all_off.p, comp_off, (nstreams + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
  CK(cudaMemcpyAsync(ctx->all_off2.p, out_off, (nstreams + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
  auto issue_h2d = [&](size_t c) -> int {
    const Chunk &ch = chunks[c];
    const int slot = (int)(c & 1);
    const uint64_t cb = comp_off[ch.b] - comp_off[ch.a];
    if (c >= 2) CK(cudaStreamWaitEvent(ctx->s_in, ctx->e_comp[slot], 0));
    if (cb) CK(cudaMemcpyAsync(ctx->p_in[slot].p, comp + comp_off[ch.a], cb, cudaMemcpyHostToDevice, ctx->s_in));
    launch_affine_u64(ctx->p_off_in[slot].as<uint64_t>(), ctx->all_off.as<uint64_t>() + ch.a, ch.b - ch.a + 1,
                      0ull - comp_off[ch.a], ctx->s_in);
    launch_affine_u64(ctx->p_off_out[slot].as<uint64_t>(), ctx->all_off2.as<uint64_t>() + ch.a, ch.b - ch.a + 1,
                      0ull - out_off[ch.a], ctx->s_in);
    CK(cudaEventRecord(ctx->e_in[slot], ctx->s_in));
    return FB200_OK;
  };
  int rc = issue_h2d(0);
  if (rc != FB200_OK) return rc;
  uint64_t launches = 0, fallbacks = 0;
  for (size_t c = 0; c < chunks.size(); c++) {
    const Chunk &ch = chunks[c];
    const int slot = (int)(c & 1);
    const uint64_t cnt = ch.b - ch.a;
    if (c + 1 < chunks.size() && (rc = issue_h2d(c + 1)) != FB200_OK) return rc;
    CK(cudaStreamWaitEvent(st, ctx->e_in[slot], 0));
    if (c >= 2) CK(cudaStreamWaitEvent(st, ctx->e_out[slot], 0));
    rc = fb200_inflate_batch_dev(ctx, ctx->p_in[slot].as<uint8_t>(), ctx->p_off_in[slot].as<uint64_t>(), cnt,
                                 ctx->p_out[slot].as<uint8_t>(), ctx->p_off_out[slot].as<uint64_t>(),
                                 ctx->p_len[slot].as<uint64_t>(), ctx->p_status[slot].as<int32_t>(),
                                 ctx->p_eoff[slot].as<int64_t>(), ctx->p_cons[slot].as<uint64_t>());
    if (rc != FB200_OK) return rc;
    launches += ctx->stats.kernel_launches + 2;
    fallbacks += ctx->stats.inflate_fallbacks;
    CK(cudaEventRecord(ctx->e_comp[slot], st));
    CK(cudaStreamWaitEvent(ctx->s_out, ctx->e_comp[slot], 0));
    const uint64_t ob = out_off[ch.b] - out_off[ch.a];
    if (ob) CK(cudaMemcpyAsync(out + out_off[ch.a], ctx->p_out[slot].p, ob, cudaMemcpyDeviceToHost, ctx->s_out));
    CK(cudaMemcpyAsync(out_len + ch.a, ctx->p_len[slot].p, cnt * 8, cudaMemcpyDeviceToHost, ctx->s_out));
    CK(cudaMemcpyAsync(status + ch.a, ctx->p_status[slot].p, cnt * 4, cudaMemcpyDeviceToHost, ctx->s_out));
    CK(cudaMemcpyAsync(err_off + ch.a, ctx->p_eoff[slot].p, cnt * 8, cudaMemcpyDeviceToHost, ctx->s_out));
    if (consumed) CK(cudaMemcpyAsync(consumed + ch.a, ctx->p_cons[slot].p, cnt * 8, cudaMemcpyDeviceToHost, ctx->s_out));
    CK(cudaEventRecord(ctx->e_out[slot], ctx->s_out));
  }
  CK(cudaStreamSynchronize(ctx->s_in));
  CK(cudaStreamSynchronize(ctx->s_out));
  CK(cudaStreamSynchronize(st));
  ctx->stats.kernel_launches = launches;
  ctx->stats.inflate_fallbacks = fallbacks;
  return FB200_OK;
}

// ------------------------------------------------------------------
// Streaming Writer (writer.mbt:10-58 / deflate.mbt:157-183, :280-294).
// The compressed bytes are a pure function of the concatenated input (blocks
// are cut only when the 65535-byte window fills or at close), so the object
// buffers writes and runs the GPU path at close.

struct fb200_writer {
  fb200_ctx *ctx;
  fb200_sink_fn sink;
  void *user;
  std::vector<uint8_t> buf;
  bool closed = false;
  int sticky = 0;
};

extern "C" fb200_writer *fb200_writer_new(fb200_ctx *ctx, fb200_sink_fn sink, void *user)
{
  if (!ctx || !sink) return nullptr;
  fb200_writer *w = new (std::nothrow) fb200_writer();
  if (!w) return nullptr;
  w->ctx = ctx;
  w->sink = sink;
  w->user = user;
  return w;
}

extern "C" fb200_writer *fb200_writer_new_dict(fb200_ctx *ctx, fb200_sink_fn sink, void *user, const uint8_t *dict,
                                               uint64_t n)
{
  fb200_writer *w = fb200_writer_new(ctx, sink, user);
  if (!w) return nullptr;
  // fill_window keeps the last 32768 bytes of the dictionary in the input window (deflate.mbt:114-120)
  if (n > (uint64_t)kMaxMatchOffset) {
    dict += n - kMaxMatchOffset;
    n = kMaxMatchOffset;
  }
  w->buf.assign(dict, dict + n);
  return w;
}

extern "C" int64_t fb200_writer_write(fb200_writer *w, const uint8_t *data, uint64_t n)
{
  if (!w) return FB200_ERR_ARG;
  if (w->closed) return FB200_ERR_CLOSED; // writer_closed_error (deflate.mbt:154, :281-283)
  if (w->sticky) return w->sticky;
  if (n) w->buf.insert(w->buf.end(), data, data + n);
  return (int64_t)n;
}

extern "C" int fb200_writer_close(fb200_writer *w)
{
  if (!w) return FB200_ERR_ARG;
  if (w->closed) return FB200_OK; // deflate.mbt:158-160
  if (w->sticky) return w->sticky;
  const uint64_t n = w->buf.size();
  std::vector<uint8_t> out(fb200_deflate_stream_bound(n));
  uint64_t off[2] = {0, n}, doff[2], olen = 0;
  int rc = fb200_deflate_streams(w->ctx, w->buf.data(), off, 1, out.data(), out.size(), doff, &olen);
  if (rc != FB200_OK) { w->sticky = rc; return rc; }
  if (w->sink(w->user, out.data(), olen) != 0) { w->sticky = FB200_ERR_ARG; return w->sticky; }
  w->closed = true;
  std::vector<uint8_t>().swap(w->buf);
  return FB200_OK;
}

extern "C" void fb200_writer_free(fb200_writer *w) { delete w; }

// Streaming Decompressor (inflate.mbt:257-418).  The first read inflates the
// whole stream on the GPU; reads then hand the result out with the
// reference's granularity: one 32 KiB window flush at a time, the final
// status riding on the read that drains the last (partial) flush.
struct fb200_reader {
  fb200_ctx *ctx;
  const uint8_t *comp;
  uint64_t n;
  bool decoded = false;
  std::vector<uint8_t> out;
  uint64_t total = 0, pos = 0;
  int32_t status = -1;
  int64_t err_off = 0;
  int rc = FB200_OK;
};

extern "C" fb200_reader *fb200_reader_new(fb200_ctx *ctx, const uint8_t *comp, uint64_t n)
{
  if (!ctx || (!comp && n)) return nullptr;
  fb200_reader *r = new (std::nothrow) fb200_reader();
  if (!r) return nullptr;
  r->ctx = ctx;
  r->comp = comp;
  r->n = n;
  return r;
}

static void reader_decode(fb200_reader *r)
{
  r->decoded = true;
  uint64_t cap = r->n * 8 + 65536;
  for (;;) {
    r->out.resize(cap);
    uint64_t coff[2] = {0, r->n}, ooff[2] = {0, cap}, olen = 0, cons = 0;
    int32_t st = -1;
    int64_t eo = 0;
    r->rc = fb200_inflate_batch(r->ctx, r->comp, coff, 1, r->out.data(), ooff, &olen, &st, &eo, &cons);
    if (r->rc != FB200_OK) { r->status = FB200_ST_INTERNAL; r->total = 0; return; }
    if (st == FB200_ST_DST_TOO_SMALL && cap < r->n * 1040 + 65536) { cap *= 4; continue; }
    r->total = olen;
    r->status = st;
    r->err_off = eo;
    return;
  }
}

extern "C" uint64_t fb200_reader_read(fb200_reader *r, uint8_t *buf, uint64_t n, int32_t *status, int64_t *err_off)
{
  if (!r || !status) return 0;
  if (!r->decoded) reader_decode(r);
  if (r->pos == r->total) { // nothing buffered: the sticky error (inflate.mbt:398-401)
    *status = r->status;
    if (err_off) *err_off = r->err_off;
    return 0;
  }
  const uint64_t window = (uint64_t)kMaxMatchOffset;
  uint64_t chunk_end = (r->pos / window + 1) * window;
  if (chunk_end > r->total) chunk_end = r->total;
  uint64_t k = chunk_end - r->pos;
  if (k > n) k = n;
  memcpy(buf, r->out.data() + r->pos, k);
  r->pos += k;
  *status = -1;
  // the status rides on the read that drains the final, partial window flush (inflate.mbt:392-396)
  if (r->pos == r->total && (r->total % window) != 0) {
    *status = r->status;
    if (err_off) *err_off = r->err_off;
  }
  return k;
}

extern "C" int fb200_reader_close(fb200_reader *r)
{
  if (!r) return FB200_ERR_ARG;
  if (!r->decoded) return FB200_OK; // err is still None
  if (r->status == FB200_ST_EOF || r->status == FB200_ST_EOF_AT_REFILL || r->status < 0) return FB200_OK;
  return r->status;
}

extern "C" void fb200_reader_free(fb200_reader *r) { delete r; }
