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
#include <chrono>
#include <cstring>
#include <new>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
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

// Host-buffer calls of different contexts on one GPU (the asynchronous forms, or several host threads) are put in
// FIFO order, by the time of the call, twice over: their H2D input copies follow one another at full PCIe rate
// instead of sharing the link, and their kernel phases follow one another instead of fighting for the SMs (a
// persistent parse next to a persistent inflate leaves both crawling) -- while the copies of one call still run
// beside the kernels and the D2H copies of the others.  A call waits on the host only until its predecessor has
// QUEUED its phase; the ordering on the device is done with events.
struct DeviceQueue {
  std::mutex m;
  std::condition_variable cv;
  uint64_t next_ticket = 0, copy_turn = 0, compute_turn = 0;
  cudaEvent_t last_feed = nullptr, last_compute = nullptr;
  bool feed_recorded = false, compute_recorded = false;
};
static DeviceQueue g_queue[64];

static double now_ms()
{
  static const auto t0 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
static bool trace_calls()
{
  static const bool on = getenv("FB200_TRACE") && atoi(getenv("FB200_TRACE")) >= 2;
  return on;
}

struct QueueTurn { // one host-buffer call's place in its device's queue; releases what it still holds when it dies
  DeviceQueue *q;
  uint64_t ticket;
  bool copy_done = false, compute_done = false;
  double t_call = 0, t_copy = 0, t_copy_end = 0, t_compute = 0, t_compute_end = 0; // FB200_TRACE=2: host-side timeline of the call
  explicit QueueTurn(int device) : q(&g_queue[device >= 0 && device < 64 ? device : 0])
  {
    std::lock_guard<std::mutex> lk(q->m);
    ticket = q->next_ticket++;
    t_call = now_ms();
    if (!q->last_feed) {
      cudaEventCreateWithFlags(&q->last_feed, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&q->last_compute, cudaEventDisableTiming);
    }
  }
  // blocks until every earlier call has queued its input copies; s_in then waits for them on the device
  void begin_copy(cudaStream_t s_in)
  {
    std::unique_lock<std::mutex> lk(q->m);
    q->cv.wait(lk, [&] { return q->copy_turn == ticket; });
    if (q->feed_recorded) cudaStreamWaitEvent(s_in, q->last_feed, 0);
    t_copy = now_ms();
  }
  void end_copy(cudaStream_t s_in)
  {
    std::lock_guard<std::mutex> lk(q->m);
    if (copy_done) return;
    t_copy_end = now_ms();
    if (s_in) { cudaEventRecord(q->last_feed, s_in); q->feed_recorded = true; }
    copy_done = true;
    q->copy_turn = ticket + 1;
    q->cv.notify_all();
  }
  void begin_compute(cudaStream_t st)
  {
    std::unique_lock<std::mutex> lk(q->m);
    q->cv.wait(lk, [&] { return q->compute_turn == ticket; });
    if (q->compute_recorded) cudaStreamWaitEvent(st, q->last_compute, 0);
    t_compute = now_ms();
  }
  // st: the stream on which everything this call launched has been (or has been made to be) ordered
  void end_compute(cudaStream_t st)
  {
    std::lock_guard<std::mutex> lk(q->m);
    if (compute_done) return;
    t_compute_end = now_ms();
    if (st) { cudaEventRecord(q->last_compute, st); q->compute_recorded = true; }
    compute_done = true;
    q->compute_turn = ticket + 1;
    q->cv.notify_all();
  }
  void trace(const char *what, double t_kernels_done) const
  {
    if (trace_calls())
      fprintf(stderr, "[fb200] #%llu %-7s call %.2f | copies queued %.2f..%.2f | kernels queued %.2f..%.2f | kernels done %.2f | return %.2f ms\n",
              (unsigned long long)ticket, what, t_call, t_copy, t_copy_end, t_compute, t_compute_end, t_kernels_done, now_ms());
  }
  ~QueueTurn()
  {
    // error paths: pass the turns on (in order: a turn can only be passed once it has come)
    if (!copy_done) { { std::unique_lock<std::mutex> lk(q->m); q->cv.wait(lk, [&] { return q->copy_turn == ticket; }); } end_copy(nullptr); }
    if (!compute_done) { { std::unique_lock<std::mutex> lk(q->m); q->cv.wait(lk, [&] { return q->compute_turn == ticket; }); } end_compute(nullptr); }
  }
};

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
  uint32_t *d_wm = nullptr;        // [2] arrival watermarks (deflate, inflate) advanced by the H2D stream
  uint32_t *wm_vals = nullptr;     // pinned: the values the H2D stream copies into d_wm, one per chunk
  volatile uint32_t *h_flags = nullptr; // mapped pinned: inflate output groups finished on the device
  DevBuf group_done;
  static constexpr int kMaxChunks = 4096;
  cudaStream_t s_in = nullptr, s_out = nullptr, s_post = nullptr;
  cudaStream_t s_xfer = nullptr; // fb200_mg_put / _get: peer copies into / out of the frame (copy engines, beside the kernels)
  uint64_t *d_frame_res = nullptr; // [2] result of the frame-range kernel
  cudaEvent_t e_xfer = nullptr;
  cudaStream_t s_post2 = nullptr;
  // Group pipeline of the host-buffer deflate (FB200_DEFLATE_PIPELINE, default on): after the parse, K2..K4 run
  // group by group (two streams, so that K3 of one group overlaps K4 of the previous one) and the output of a
  // packed group is copied back while the later groups are still being packed.
  bool pipeline = true;
  uint64_t group_bytes = 64ull << 20;
  static constexpr int kMaxGroups = 256;
  // block-parallel parse of multi-block streams (FB200_PARSE_BLOCKPAR: 0 never, 1 when there are few such
  // streams (default), 2 always)
  int blockpar = 1;
  DevBuf bp_flags, bp_idx, bp_tabs, bp_state, bp_list;
  DevBuf parse_gtables; // hash tables of the parse warps that have no shared-memory table (per GPU)
  DevBuf d_group;                 // [kMaxGroups] u32: K3 work counters
  DevBuf d_group_bounds;          // [kMaxGroups + 1] u64: first block of every group
  uint64_t *h_gbounds = nullptr;  // pinned [kMaxGroups + 1] block bounds, [kMaxGroups + 1] output byte offsets
  std::vector<cudaEvent_t> e_gsize, e_gdone, e_gchain;
  cudaEvent_t e_bounds = nullptr, e_parsed = nullptr;
  cudaEvent_t e_in[2] = {}, e_comp[2] = {}, e_out[2] = {};
  uint64_t chunk_bytes = 32ull << 20;
  // Host-buffer calls overlap the H2D copy with the kernel that consumes it: the kernel waits on a device watermark
  // the copy stream advances.  All copies are queued BEFORE the kernel is launched, so a launch that blocks the
  // host (CUDA_LAUNCH_BLOCKING, a debugger) cannot deadlock; tools that replay kernels (ncu restores device memory,
  // the watermark included, between passes) need the copies to have finished before the launch, which is what
  // overlap_h2d == false selects: no watermark, the kernel is ordered behind the last chunk.
  bool overlap_h2d = true;
  // calls with at most this many streams inflate with one CTA per stream instead of one warp per stream
  // (FB200_INFLATE_CTA_STREAMS; 0 = always a warp per stream)
  uint32_t cta_streams = 296;
  double t_kernels_done = 0; // FB200_TRACE=2
  // fb200_*_async: the blocking call runs on a helper thread; fb200_wait joins it (one call in flight per context)
  std::thread worker;
  bool async_pending = false;
  int async_rc = FB200_OK;
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

extern char **environ;

// A tool that serialises kernel launches or replays kernels is attached to this process (Nsight Compute,
// compute-sanitizer, cuda-gdb inject through these variables), or launches were made blocking.
static bool serialising_tool_attached()
{
  if (const char *e = getenv("CUDA_LAUNCH_BLOCKING"))
    if (*e && strcmp(e, "0") != 0) return true;
  static const char *const prefixes[] = {"CUDA_INJECTION", "NV_COMPUTE_PROFILER_", "NV_NSIGHT_", "NV_TPS_LAUNCH",
                                         "NVTX_INJECTION", "NV_SANITIZER_", "COMPUTE_SANITIZER_", "CUDA_DEBUGGER_"};
  for (char **e = environ; e && *e; e++)
    for (const char *p : prefixes)
      if (strncmp(*e, p, strlen(p)) == 0) return true;
  return false;
}

// test hook (not part of the public header): the closed form the block-parallel parse uses for "this block starts
// with a cleared table" (deflate-fast.mbt:129-132)
extern "C" int fb200_debug_block_resets(uint64_t b) { return block_resets_table(b) ? 1 : 0; }

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
      cudaMallocHost((void **)&ctx->pinned, 64 * sizeof(uint64_t)) != cudaSuccess ||
      cudaMallocHost((void **)&ctx->h_gbounds, 2 * (fb200_ctx::kMaxGroups + 1) * sizeof(uint64_t)) != cudaSuccess) {
    cudaGetLastError();
    fb200_destroy(ctx); // releases whatever was created
    return FB200_ERR_CUDA;
  }
  for (int i = 0; i < FB200_NUM_STAGES; i++) {
    cudaEventCreate(&ctx->ev0[i]);
    cudaEventCreate(&ctx->ev1[i]);
  }
  cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->s_post, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->s_xfer, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->e_xfer, cudaEventDisableTiming);
  cudaStreamCreateWithFlags(&ctx->s_post2, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->e_bounds, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->e_parsed, cudaEventDisableTiming);
  // FB200_HOST_OVERLAP: 1 forces the watermark overlap on, 0 off; unset = on unless a serialising tool is attached
  ctx->overlap_h2d = !serialising_tool_attached();
  if (const char *e = getenv("FB200_HOST_OVERLAP")) ctx->overlap_h2d = atoi(e) != 0;
  if (const char *e = getenv("FB200_INFLATE_CTA_STREAMS")) ctx->cta_streams = (uint32_t)atol(e);
  if (const char *e = getenv("FB200_DEFLATE_PIPELINE")) ctx->pipeline = atoi(e) != 0;
  if (const char *e = getenv("FB200_PARSE_BLOCKPAR")) ctx->blockpar = atoi(e);
  if (const char *e = getenv("FB200_GROUP_MB")) {
    const long mb = atol(e);
    if (mb > 0) ctx->group_bytes = (uint64_t)mb << 20;
  }
  for (int i = 0; i < 2; i++) {
    cudaEventCreateWithFlags(&ctx->e_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->e_comp[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->e_out[i], cudaEventDisableTiming);
  }
  cudaMalloc((void **)&ctx->d_wm, 64);
  cudaMalloc((void **)&ctx->d_frame_res, 16);
  cudaMallocHost((void **)&ctx->wm_vals, fb200_ctx::kMaxChunks * sizeof(uint32_t));
  cudaHostAlloc((void **)&ctx->h_flags, fb200_ctx::kMaxChunks * sizeof(uint32_t), cudaHostAllocMapped);
  if (!ctx->d_wm || !ctx->d_frame_res || !ctx->wm_vals || !ctx->h_flags) { cudaGetLastError(); fb200_destroy(ctx); return FB200_ERR_CUDA; }
  if (const char *e = getenv("FB200_CHUNK_MB")) {
    const long mb = atol(e);
    if (mb > 0) ctx->chunk_bytes = (uint64_t)mb << 20;
  }
  preload_parse_kernels();
  preload_encode_kernels();
  preload_inflate_kernels();
  preload_inflate3_kernels();
  launch_init_tables(ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    cudaGetLastError();
    fb200_destroy(ctx);
    return FB200_ERR_CUDA;
  }
  *out = ctx;
  return FB200_OK;
}

extern "C" void fb200_destroy(fb200_ctx *ctx)
{
  if (!ctx) return;
  if (ctx->worker.joinable()) ctx->worker.join();
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
  ctx->group_done.release();
  ctx->d_group.release();
  for (DevBuf *b : {&ctx->bp_flags, &ctx->bp_idx, &ctx->bp_tabs, &ctx->bp_state, &ctx->bp_list, &ctx->parse_gtables}) b->release();
  ctx->d_group_bounds.release();
  for (auto *v : {&ctx->e_gsize, &ctx->e_gdone, &ctx->e_gchain})
    for (cudaEvent_t e : *v) cudaEventDestroy(e);
  if (ctx->e_bounds) cudaEventDestroy(ctx->e_bounds);
  if (ctx->e_parsed) cudaEventDestroy(ctx->e_parsed);
  if (ctx->h_gbounds) cudaFreeHost(ctx->h_gbounds);
  if (ctx->s_post) cudaStreamDestroy(ctx->s_post);
  if (ctx->s_post2) cudaStreamDestroy(ctx->s_post2);
  if (ctx->s_xfer) cudaStreamDestroy(ctx->s_xfer);
  if (ctx->e_xfer) cudaEventDestroy(ctx->e_xfer);
  if (ctx->d_wm) cudaFree(ctx->d_wm);
  if (ctx->d_frame_res) cudaFree(ctx->d_frame_res);
  if (ctx->wm_vals) cudaFreeHost(ctx->wm_vals);
  if (ctx->h_flags) cudaFreeHost((void *)ctx->h_flags);
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
// multi-GPU frame assembly: the frame lives on the assembling GPU, peers map it through
// CUDA IPC and put their payload into it with copy-engine peer copies over NVLink.
static_assert(sizeof(cudaIpcMemHandle_t) == FB200_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int fb200_mg_frame_alloc(fb200_ctx *ctx, uint64_t bytes, void **d_frame, uint8_t *handle)
{
  if (!ctx || !d_frame || !handle || bytes == 0) return FB200_ERR_ARG;
  *d_frame = nullptr;
  CK(cudaSetDevice(ctx->device));
  void *p = nullptr;
  CK(cudaMalloc(&p, bytes)); // a plain allocation of its own: IPC exports whole allocations
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    ctx->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return FB200_ERR_CUDA;
  }
  memcpy(handle, &h, sizeof h);
  *d_frame = p;
  return FB200_OK;
}

extern "C" int fb200_mg_frame_open(fb200_ctx *ctx, const uint8_t *handle, void **d_frame)
{
  if (!ctx || !d_frame || !handle) return FB200_ERR_ARG;
  *d_frame = nullptr;
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CK(cudaIpcOpenMemHandle(d_frame, h, cudaIpcMemLazyEnablePeerAccess));
  return FB200_OK;
}

extern "C" int fb200_mg_frame_close(fb200_ctx *ctx, void *d_frame, int owner)
{
  if (!ctx) return FB200_ERR_ARG;
  if (!d_frame) return FB200_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->s_xfer));
  if (owner) CK(cudaFree(d_frame));
  else CK(cudaIpcCloseMemHandle(d_frame));
  return FB200_OK;
}

extern "C" int fb200_mg_put(fb200_ctx *ctx, void *d_frame, uint64_t offset, const void *d_payload, uint64_t n)
{
  if (!ctx || !d_frame || (!d_payload && n)) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return FB200_OK;
  CK(cudaEventRecord(ctx->e_xfer, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->s_xfer, ctx->e_xfer, 0));
  CK(cudaMemcpyAsync(static_cast<uint8_t *>(d_frame) + offset, d_payload, n, cudaMemcpyDeviceToDevice, ctx->s_xfer));
  return FB200_OK;
}

extern "C" int fb200_mg_wait(fb200_ctx *ctx)
{
  if (!ctx) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->s_xfer));
  return FB200_OK;
}

// Frame reader: the compressed streams of segments [first, first + count) travel from the frame (on this GPU, or on
// the assembling GPU and mapped through CUDA IPC: a peer copy over NVLink) into d_comp, and their offsets
// (count + 1 values, starting at 0) are written to d_comp_off.  The sizes are read from the frame header by a
// kernel (peer loads when the frame is remote); the payload moves with the copy engines.
static int mg_get_impl(fb200_ctx *ctx, const void *d_frame, uint64_t frame_bytes, uint64_t first, uint64_t count,
                       uint8_t *d_comp, uint64_t comp_cap, uint64_t *d_comp_off, uint32_t *seg_size, uint64_t *nseg_total,
                       uint64_t *out_bytes, bool wait)
{
  if (!ctx || !d_frame || !d_comp_off || (!d_comp && comp_cap) || !out_bytes) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->s_xfer;
  *out_bytes = 0;
  if (frame_bytes < 16) { ctx->err = "frame shorter than its header"; return FB200_ERR_ARG; }
  uint32_t *h = reinterpret_cast<uint32_t *>(ctx->pinned + 48);
  CK(cudaMemcpyAsync(h, d_frame, 16, cudaMemcpyDefault, st));
  CK(cudaStreamSynchronize(st));
  const uint64_t nseg = (uint64_t)h[2] | ((uint64_t)h[3] << 32);
  if (h[0] != FB200_FRAME_MAGIC) { ctx->err = "not a frame (magic)"; return FB200_ERR_ARG; }
  if (fb200_frame_header_bytes(nseg) > frame_bytes || first > nseg || count > nseg - first) {
    ctx->err = "segment range outside the frame";
    return FB200_ERR_ARG;
  }
  if (seg_size) *seg_size = h[1];
  if (nseg_total) *nseg_total = nseg;
  uint64_t *res = ctx->pinned + 50; // {payload offset of segment `first` in the frame, bytes of the range}
  launch_frame_range(reinterpret_cast<const uint32_t *>(static_cast<const uint8_t *>(d_frame) + 16), first, count, d_comp_off,
                     ctx->d_frame_res, st);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(res, ctx->d_frame_res, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint64_t base = fb200_frame_header_bytes(nseg) + res[0], total = res[1];
  *out_bytes = total;
  if (base + total > frame_bytes) { ctx->err = "frame truncated: the streams of the range end beyond it"; return FB200_ERR_ARG; }
  if (total > comp_cap) { ctx->err = "comp_cap too small"; return FB200_ERR_DST_TOO_SMALL; }
  if (total) CK(cudaMemcpyAsync(d_comp, static_cast<const uint8_t *>(d_frame) + base, total, cudaMemcpyDefault, st));
  if (wait) CK(cudaStreamSynchronize(st));
  return FB200_OK;
}

extern "C" int fb200_mg_get(fb200_ctx *ctx, const void *d_frame, uint64_t frame_bytes, uint64_t first, uint64_t count,
                            uint8_t *d_comp, uint64_t comp_cap, uint64_t *d_comp_off, uint32_t *seg_size, uint64_t *nseg_total,
                            uint64_t *out_bytes)
{
  return mg_get_impl(ctx, d_frame, frame_bytes, first, count, d_comp, comp_cap, d_comp_off, seg_size, nseg_total, out_bytes,
                     true);
}

extern "C" int fb200_mg_get_async(fb200_ctx *ctx, const void *d_frame, uint64_t frame_bytes, uint64_t first, uint64_t count,
                                  uint8_t *d_comp, uint64_t comp_cap, uint64_t *d_comp_off, uint32_t *seg_size,
                                  uint64_t *nseg_total, uint64_t *out_bytes)
{
  return mg_get_impl(ctx, d_frame, frame_bytes, first, count, d_comp, comp_cap, d_comp_off, seg_size, nseg_total, out_bytes,
                     false);
}

// ------------------------------------------------------------------
// deflate core.  One launch of the parse over the whole batch, then K2, K3, layout and K4.  Device-buffer calls
// run them over the whole batch in stream order.  Host-buffer calls cut the streams into groups and run K2..K4
// group by group on two streams; the output of a packed group travels back to the host (copy engine) while
// the later groups are being packed.  (Running K2..K4 of finished groups beside the still running parse was
// built and measured: the parse leaves ~12 K registers and 3 KB of shared memory per SM, the co-resident CTAs
// crawl -- group 0's histogram + codes took 12 ms -- and making room for them, a 196 KB carve-out, costs the
// parse 1.1 ms of L1 hits.  The SMs have no idle capacity to give; the copy engines do.)

struct DeflateIo {
  const uint8_t *d_src = nullptr;
  const uint64_t *d_src_off = nullptr; // device [ns + 1]
  uint64_t ns = 0, n_total = 0;
  uint64_t *d_dst_off = nullptr;       // device [ns + 1]
  uint64_t nb_known = ~0ull;           // number of blocks if the caller knows it, else read back
  uint64_t n_multi = 0, nmb = 0;       // (with nb_known) streams of more than one parsed block, and their blocks
  const uint32_t *avail = nullptr;     // arrival watermark (host-buffer calls)
  uint8_t *d_dst = nullptr;            // device output (4-byte aligned)
  uint64_t d_cap = 0;                  // its capacity in bytes
  uint8_t *h_dst = nullptr;            // host-buffer calls: where finished groups are copied to
  uint64_t h_cap = 0;
  QueueTurn *turn = nullptr;           // host-buffer calls: place in the device's FIFO (kernel phase)
  // continuation of one stream across calls (DeflateJob::cont_*): the streaming Writer
  bool cont_prev = false, cont_open = false;
  uint32_t cont_start_bit = 0;
  uint64_t cont_block_base = 0;
  uint16_t *d_seed = nullptr;          // device [1 << 14]: end table of the previous call (in), of this call (out, cont_open)
};

static int deflate_run(fb200_ctx *ctx, const DeflateIo &io, uint64_t *total_out)
{
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint64_t ns = io.ns;
  ctx->stats = fb200_stats{};
  *total_out = 0;
  if (ns > 0xfffffff0ull) { ctx->err = "too many streams"; return FB200_ERR_ARG; }
  if ((reinterpret_cast<uintptr_t>(io.d_dst) & 3) != 0) { ctx->err = "device dst must be 4-byte aligned"; return FB200_ERR_ARG; }
  DeflateJob &j = ctx->last;
  j = DeflateJob{};
  j.src = io.d_src;
  j.stream_off = io.d_src_off;
  j.nstreams = ns;
  CK(ctx->stream_blk0.ensure((ns + 1) * 8));
  CK(ctx->stream_bytes.ensure((ns + 1) * 8));
  CK(ctx->stream_trailer.ensure((ns + 1) * 8));
  CK(ctx->counters.ensure(64));
  static_assert(kBuildCounterStride == fb200_ctx::kMaxGroups, "one K3 counter pair per group");
  CK(ctx->d_group.ensure(2 * fb200_ctx::kMaxGroups * 4));
  CK(ctx->d_group_bounds.ensure((fb200_ctx::kMaxGroups + 1) * 8));
  j.stream_blk0 = ctx->stream_blk0.as<uint64_t>();
  j.stream_bytes = ctx->stream_bytes.as<uint64_t>();
  j.stream_trailer_bit = ctx->stream_trailer.as<uint64_t>();
  j.dst_off = io.d_dst_off;
  j.counters = ctx->counters.as<uint32_t>();
  j.avail = io.avail;
  j.dst = io.d_dst;
  j.dst_cap = io.d_cap & ~3ull; // K4 clears and writes whole words
  j.cont_prev = io.cont_prev ? 1u : 0u;
  j.cont_open = io.cont_open ? 1u : 0u;
  j.cont_start_bit = io.cont_start_bit;
  j.cont_block_base = io.cont_block_base;
  if ((io.cont_prev || io.cont_open) && (ns != 1 || !io.d_seed)) { ctx->err = "a continued stream is a single stream"; return FB200_ERR_ARG; }
  CK(cudaMemsetAsync(j.counters, 0, 64, st));
  uint64_t launches = 0;
  for (int i = 0; i < FB200_NUM_STAGES; i++) ctx->ev_used[i] = false;

  // groups: ~group_bytes of input each
  // (groups for device-buffer calls too: measured, K2..K4 take 4.0-4.9 ms instead of 4.2 -- nothing to overlap with)
  const bool piped = ctx->pipeline && ns > 0 && io.h_dst != nullptr;
  uint64_t gs = ns ? ns : 1;
  if (piped) {
    const double avg = ns ? (double)io.n_total / (double)ns : 1.0;
    gs = (uint64_t)((double)ctx->group_bytes / (avg > 1.0 ? avg : 1.0));
    if (gs == 0) gs = 1;
    if (gs > ns) gs = ns;
    if ((ns + gs - 1) / gs > (uint64_t)fb200_ctx::kMaxGroups) gs = (ns + fb200_ctx::kMaxGroups - 1) / fb200_ctx::kMaxGroups;
  }
  const uint64_t ngroups = ns ? (ns + gs - 1) / gs : 0;
  uint32_t *d_k3cnt = ctx->d_group.as<uint32_t>();
  CK(cudaMemsetAsync(d_k3cnt, 0, 2 * fb200_ctx::kMaxGroups * 4, st));
  while (ctx->e_gsize.size() < ngroups) {
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); // (timed: FB200_TRACE prints the group timeline)
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    ctx->e_gsize.push_back(e0);
    ctx->e_gdone.push_back(e1);
    ctx->e_gchain.push_back(e2);
  }
  uint64_t *h_bb = ctx->h_gbounds, *h_goff = ctx->h_gbounds + fb200_ctx::kMaxGroups + 1;

  ctx->stage_begin(FB200_STAGE_SETUP);
  launch_count_blocks(j, st);
  launch_scan_u64(j.stream_blk0, j.stream_blk0, ns, st);
  // (the group bounds go straight into pinned host memory, written by the kernel: a small D2H copy would queue
  // behind the bulk copies of other calls in the copy engine, and the host waits for these few bytes below)
  launch_gather_u64(h_bb, j.stream_blk0, gs, ns, ngroups + 1, st);
  CK(cudaEventRecord(ctx->e_bounds, st));
  launches += 3;
  uint64_t nb = io.nb_known, n_multi = io.n_multi, nmb = io.nmb;
  if (nb == ~0ull) {
    launch_count_multi(j, st); // -> counters[12] streams, counters[13] blocks (saturating)
    CK(cudaMemcpyAsync(ctx->pinned + 16, j.counters + 12, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    nb = h_bb[ngroups];
    const uint32_t *cm = reinterpret_cast<const uint32_t *>(ctx->pinned + 16);
    n_multi = cm[0];
    nmb = cm[1];
  }
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
  CK(ctx->tokens.ensure((io.n_total + 16) * 4));
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
  ctx->last_n_total = io.n_total;
  CK(cudaMemsetAsync(j.blk_ntok, 0, nbp * 4, st));
  CK(cudaMemsetAsync(j.blk_bits, 0, nbp * 4, st));
  launch_fill_blocks(j, st);
  launches += 1;
  ctx->stage_end(FB200_STAGE_SETUP);

  CK(ctx->parse_gtables.ensure(parse_gtables_bytes(ctx->num_sms)));
  if (io.turn) io.turn->begin_compute(st);
  ctx->stage_begin(FB200_STAGE_PARSE);
  launch_parse_single(j, ctx->num_sms, ctx->parse_gtables.p, st);
  launches += 1;
  CK(cudaGetLastError());
  const bool blockpar = n_multi > 0 && (ctx->blockpar == 2 || (ctx->blockpar == 1 && n_multi < 4096) || io.cont_prev || io.cont_open);
  if (n_multi > 0 && !blockpar) {
    launch_parse_multi(j, ctx->num_sms, ctx->parse_gtables.p, st); // one warp per stream, its blocks in sequence
    launches += 1;
  } else if (blockpar) {
    // rounds over the blocks of the multi-block streams (BlockParJob, kernels.h)
    CK(ctx->bp_flags.ensure((nb + 1) * 8));
    CK(ctx->bp_idx.ensure((nb + 1) * 8));
    launch_bp_flags(j, ctx->bp_flags.as<uint64_t>(), st);
    launch_scan_u64(ctx->bp_flags.as<uint64_t>(), ctx->bp_idx.as<uint64_t>(), nb, st);
    CK(cudaMemcpyAsync(ctx->pinned + 17, ctx->bp_idx.as<uint64_t>() + nb, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    nmb = ctx->pinned[17];
    CK(ctx->bp_tabs.ensure(2 * nmb * (uint64_t)kTableSize * 2));
    CK(ctx->bp_state.ensure(4 * nmb + 16)); // lat[2][nmb], chg[2][nmb]
    CK(ctx->bp_list.ensure(nmb * 4 + 16));
    uint8_t *lat = ctx->bp_state.as<uint8_t>(), *chg = lat + 2 * nmb;
    CK(cudaMemsetAsync(lat, 0, 4 * nmb, st)); // round 1 reads "latest copy = 0" for every block
    BlockParJob bp{};
    bp.list = ctx->bp_list.as<uint32_t>();
    bp.mb_idx = ctx->bp_idx.as<uint64_t>();
    bp.tabs = ctx->bp_tabs.as<uint16_t>();
    bp.nmb = nmb;
    bp.cont_prev = j.cont_prev;
    bp.cont_open = j.cont_open;
    bp.cont_block_base = j.cont_block_base;
    if (io.cont_prev) // the stand-in block's end table = what the previous call left (copy 0 is every block's latest at first)
      CK(cudaMemcpyAsync(bp.tabs, io.d_seed, (size_t)kTableSize * 2, cudaMemcpyDeviceToDevice, st));
    uint32_t *h_nlist = reinterpret_cast<uint32_t *>(ctx->pinned + 18);
    uint64_t rounds = 0, parsed_blocks = 0;
    for (int round = 1;; round++) {
      bp.round = round;
      bp.lat_prev = lat + (size_t)((round - 1) & 1) * nmb;
      bp.lat_next = lat + (size_t)(round & 1) * nmb;
      bp.chg_next = chg + (size_t)(round & 1) * nmb;
      CK(cudaMemsetAsync(j.counters + 13, 0, 8, st)); // [13] list length, [14] work counter of the round
      launch_bp_round(j, bp, chg + (size_t)((round - 1) & 1) * nmb, ctx->bp_list.as<uint32_t>(), j.counters + 13, st);
      CK(cudaMemcpyAsync(h_nlist, j.counters + 13, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      launches += 1;
      if (*h_nlist == 0) break;
      bp.nlist = *h_nlist;
      launch_parse_blocks(j, bp, j.counters + 14, ctx->num_sms, ctx->parse_gtables.p, st);
      launches += 1;
      rounds++;
      parsed_blocks += bp.nlist;
      if ((uint64_t)round > nmb + 2) { ctx->err = "internal error: block-parallel parse does not converge"; return FB200_ERR_CUDA; }
    }
    if (io.cont_open) // the end table of the last block, for the next call (bp.lat_next: the round that found nothing to do)
      launch_bp_save_table(bp, nmb - 1, bp.lat_next, io.d_seed, st);
    CK(cudaGetLastError());
    if (getenv("FB200_TRACE"))
      fprintf(stderr, "[fb200] block-parallel parse: %llu streams, %llu blocks, %llu rounds, %llu block parses\n",
              (unsigned long long)n_multi, (unsigned long long)nmb, (unsigned long long)rounds, (unsigned long long)parsed_blocks);
  }
  ctx->stage_end(FB200_STAGE_PARSE);

  // K2 .. K4 of one group on stream sp
  auto launch_group = [&](uint64_t g, cudaStream_t sp, bool chain) -> int {
    DeflateJob jg = j;
    jg.st_begin = g * gs;
    jg.st_end = jg.st_begin + gs < ns ? jg.st_begin + gs : ns;
    jg.blk_begin = piped ? h_bb[g] : 0;
    jg.blk_end = piped ? h_bb[g + 1] : nb;
    jg.work_counter = d_k3cnt + g;
    if (!piped) ctx->stage_begin(FB200_STAGE_HISTOGRAM);
    launch_histogram(jg, sp);
    if (!piped) { ctx->stage_end(FB200_STAGE_HISTOGRAM); ctx->stage_begin(FB200_STAGE_BUILD); }
    launch_build_codes(jg, ctx->num_sms, sp);
    if (!piped) { ctx->stage_end(FB200_STAGE_BUILD); ctx->stage_begin(FB200_STAGE_LAYOUT); }
    if (chain && g > 0) CK(cudaStreamWaitEvent(sp, ctx->e_gchain[g - 1], 0)); // offsets + shared edge word of the previous group
    launch_layout(jg, sp);
    launch_scan_u64(j.stream_bytes + jg.st_begin, j.dst_off + jg.st_begin, jg.st_end - jg.st_begin, sp,
                    g ? j.dst_off + jg.st_begin : nullptr, h_goff + g + 1);
    CK(cudaEventRecord(ctx->e_gsize[g], sp));
    if (!piped) { ctx->stage_end(FB200_STAGE_LAYOUT); ctx->stage_begin(FB200_STAGE_PACK); }
    launch_zero_range(jg, sp);
    if (chain) CK(cudaEventRecord(ctx->e_gchain[g], sp));
    launch_pack(jg, sp);
    if (!piped) ctx->stage_end(FB200_STAGE_PACK);
    CK(cudaEventRecord(ctx->e_gdone[g], sp));
    launches += 7;
    CK(cudaGetLastError());
    return FB200_OK;
  };
  // output of group g back to the host (host-buffer calls)
  h_goff[0] = 0;
  bool overflow = false;
  auto copy_group = [&](uint64_t g) -> int {
    const uint64_t a = h_goff[g], b = h_goff[g + 1];
    if (b > io.h_cap || b > j.dst_cap) { overflow = true; return FB200_OK; }
    CK(cudaStreamWaitEvent(ctx->s_out, ctx->e_gdone[g], 0));
    if (b > a) CK(cudaMemcpyAsync(io.h_dst + a, io.d_dst + a, b - a, cudaMemcpyDeviceToHost, ctx->s_out));
    return FB200_OK;
  };

  if (!piped) {
    if (ngroups) {
      const int rc = launch_group(0, st, false);
      if (rc != FB200_OK) return rc;
      if (io.turn) io.turn->end_compute(st);
      CK(cudaStreamSynchronize(st));
      if (io.h_dst) {
        const int rc2 = copy_group(0);
        if (rc2 != FB200_OK) return rc2;
      }
    }
  } else {
    CK(cudaEventRecord(ctx->e_parsed, st));
    CK(cudaStreamWaitEvent(ctx->s_post, ctx->e_parsed, 0));
    CK(cudaStreamWaitEvent(ctx->s_post2, ctx->e_parsed, 0));
    CK(cudaEventSynchronize(ctx->e_bounds)); // block bounds of the groups are on the host (queued before the parse)
    for (uint64_t g = 0; g < ngroups; g++) {
      const int rc = launch_group(g, (g & 1) ? ctx->s_post2 : ctx->s_post, true);
      if (rc != FB200_OK) return rc;
    }
    if (io.turn) { // everything this call launches has been queued: the next call's kernels follow the last groups
      if (ngroups >= 1) CK(cudaStreamWaitEvent(st, ctx->e_gdone[ngroups - 1], 0));
      if (ngroups >= 2) CK(cudaStreamWaitEvent(st, ctx->e_gdone[ngroups - 2], 0));
      io.turn->end_compute(st);
    }
    for (uint64_t g = 0; g < ngroups && io.h_dst; g++) {
      CK(cudaEventSynchronize(ctx->e_gsize[g]));
      const int rc = copy_group(g);
      if (rc != FB200_OK) return rc;
    }
    // "pack" stage of the pipeline = K2 .. K4 of all groups
    CK(cudaEventRecord(ctx->ev0[FB200_STAGE_PACK], st)); // = end of the parse (stream order)
    CK(cudaStreamSynchronize(ctx->s_post));
    CK(cudaStreamSynchronize(ctx->s_post2));
    ctx->t_kernels_done = now_ms();
    CK(cudaEventRecord(ctx->ev1[FB200_STAGE_PACK], st));
    ctx->ev_used[FB200_STAGE_PACK] = true;
  }
  CK(cudaMemcpyAsync(ctx->pinned, j.counters, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (io.h_dst) CK(cudaStreamSynchronize(ctx->s_out));
  static const bool trace = getenv("FB200_TRACE") != nullptr;
  if (trace && piped) {
    float t_p0 = 0, t_p1 = 0;
    cudaEventElapsedTime(&t_p0, ctx->ev0[FB200_STAGE_SETUP], ctx->ev0[FB200_STAGE_PARSE]);
    cudaEventElapsedTime(&t_p1, ctx->ev0[FB200_STAGE_SETUP], ctx->ev1[FB200_STAGE_PARSE]);
    fprintf(stderr, "[fb200] deflate: parse %.2f .. %.2f ms; groups (layout done / packed):", t_p0, t_p1);
    for (uint64_t g = 0; g < ngroups; g++) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, ctx->ev0[FB200_STAGE_SETUP], ctx->e_gsize[g]);
      cudaEventElapsedTime(&b, ctx->ev0[FB200_STAGE_SETUP], ctx->e_gdone[g]);
      fprintf(stderr, " %.2f/%.2f", a, b);
    }
    fprintf(stderr, "\n");
  }
  const uint64_t total = ngroups ? h_goff[ngroups] : 0;
  *total_out = total;
  ctx->stats.nblocks = nb;
  ctx->stats.kernel_launches = launches;
  if (overflow || total > j.dst_cap || (io.h_dst && total > io.h_cap)) { ctx->err = "dst_cap too small"; return FB200_ERR_DST_TOO_SMALL; }
  if (reinterpret_cast<const uint32_t *>(ctx->pinned)[4] != 0) {
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
  DeflateIo io;
  io.d_src = d_src;
  io.d_src_off = d_src_off;
  io.ns = nstreams;
  io.n_total = n_total;
  io.d_dst_off = d_dst_off;
  io.d_dst = d_dst;
  io.d_cap = dst_cap;
  uint64_t total = 0;
  const int rc = deflate_run(ctx, io, &total);
  *out_len = total;
  return rc;
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

// Host-buffer deflate.  The kernels are launched once for the whole batch; the input is copied in chunks
// on a second stream, and after every chunk that stream advances a device watermark ("streams whose bytes
// have arrived") which the parse waits on, so the H2D copy runs beside the parse instead of in front of it.
// Chunk boundaries are rounded up to 128 bytes: a cache line never holds bytes of two different arrivals.
// Every copy and watermark update is queued before the parse is launched (see fb200_ctx::overlap_h2d).
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
  if (ns > 0xfffffff0ull) { ctx->err = "too many streams"; return FB200_ERR_ARG; }
  const uint64_t base0 = src_off ? src_off[0] : 0; // streams may start anywhere in src
  // chunks of whole streams + number of blocks (deflate.mbt:222-229: one block per 65535 bytes)
  struct Cut { uint64_t streams, bytes; }; // streams / bytes delivered once this chunk has arrived
  std::vector<Cut> cuts;
  uint64_t nb = 0, n_multi = 0, nmb = 0;
  uint64_t step = ctx->chunk_bytes;
  if (n / step + 2 > (uint64_t)fb200_ctx::kMaxChunks) step = n / (fb200_ctx::kMaxChunks - 2) + 1;
  if (src_off) {
    uint64_t next_cut = step;
    for (uint64_t i = 0; i < ns; i++) {
      const uint64_t len = src_off[i + 1] - src_off[i];
      nb += (len + kBlockSize - 1) / kBlockSize;
      if (len >= (uint64_t)kBlockSize + 128) { n_multi++; nmb += (len + kBlockSize - 1) / kBlockSize; }
      const uint64_t endb = src_off[i + 1] - base0;
      if (endb >= next_cut || i + 1 == ns) {
        cuts.push_back({i + 1, endb});
        next_cut = endb + step;
      }
    }
  } else {
    uint64_t per = step / seg_size;
    if (per == 0) per = 1;
    for (uint64_t a = 0; a < ns; a += per) {
      const uint64_t b = a + per < ns ? a + per : ns;
      cuts.push_back({b, b * seg_size < n ? b * seg_size : n});
    }
    const uint64_t full = n / seg_size, tail = n % seg_size;
    nb = full * ((seg_size + kBlockSize - 1) / kBlockSize) + (tail + kBlockSize - 1) / kBlockSize;
    if (seg_size >= (uint64_t)kBlockSize + 128) { n_multi += full; nmb += full * ((seg_size + kBlockSize - 1) / kBlockSize); }
    if (tail >= (uint64_t)kBlockSize + 128) { n_multi++; nmb += (tail + kBlockSize - 1) / kBlockSize; }
  }
  QueueTurn turn(ctx->device);
  CK(ctx->p_in[0].ensure(n + 256));
  CK(ctx->p_off_in[0].ensure((ns + 1) * 8));
  CK(ctx->p_off_out[0].ensure((ns + 1) * 8));
  uint8_t *d_src = ctx->p_in[0].as<uint8_t>();
  uint64_t *d_off = ctx->p_off_in[0].as<uint64_t>();
  // reset the watermark before the H2D stream may touch it
  CK(cudaMemsetAsync(ctx->d_wm, 0, 8, st));
  CK(cudaEventRecord(ctx->e_comp[0], st));
  CK(cudaStreamWaitEvent(ctx->s_in, ctx->e_comp[0], 0));
  turn.begin_copy(ctx->s_in);
  if (src_off) {
    CK(ctx->all_off.ensure((ns + 1) * 8));
    CK(cudaMemcpyAsync(ctx->all_off.p, src_off, (ns + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
    CK(cudaEventRecord(ctx->e_in[0], ctx->s_in));
    CK(cudaStreamWaitEvent(st, ctx->e_in[0], 0));
    launch_affine_u64(d_off, ctx->all_off.as<uint64_t>(), ns + 1, 0ull - base0, st);
  } else {
    launch_fill_seg_off(d_off, ns, seg_size, n, st);
  }
  // device output: the caller's capacity bounds it (a call that does not fit reports the need and writes nothing more)
  uint64_t dcap = 2 * n + 640 * (nb + ns) + 16 * ns + 64;
  if (dcap > ((dst_cap + 3) & ~3ull) + 4) dcap = ((dst_cap + 3) & ~3ull) + 4;
  CK(ctx->p_out[0].ensure(dcap + 16));
  DeflateIo io;
  io.d_src = d_src;
  io.d_src_off = d_off;
  io.ns = ns;
  io.n_total = n;
  io.d_dst_off = ctx->p_off_out[0].as<uint64_t>();
  io.nb_known = nb;
  io.n_multi = n_multi;
  io.nmb = nmb;
  io.avail = ctx->overlap_h2d ? ctx->d_wm : nullptr;
  io.d_dst = ctx->p_out[0].as<uint8_t>();
  io.d_cap = dcap;
  io.h_dst = dst;
  io.h_cap = dst_cap;
  // the feed: input chunks + watermark updates on the copy stream, all queued before any kernel that waits on them
  {
    uint64_t done = 0;
    for (size_t c = 0; c < cuts.size(); c++) {
      uint64_t upto = c + 1 == cuts.size() ? n : ((cuts[c].bytes + 127) & ~127ull);
      if (upto > n) upto = n;
      if (upto > done) {
        CK(cudaMemcpyAsync(d_src + done, src + base0 + done, upto - done, cudaMemcpyHostToDevice, ctx->s_in));
        done = upto;
      }
      if (ctx->overlap_h2d) {
        ctx->wm_vals[c] = (uint32_t)cuts[c].streams;
        CK(cudaMemcpyAsync(ctx->d_wm, &ctx->wm_vals[c], 4, cudaMemcpyHostToDevice, ctx->s_in));
      }
    }
    if (!ctx->overlap_h2d) { // copy, then launch
      CK(cudaEventRecord(ctx->e_in[1], ctx->s_in));
      CK(cudaStreamWaitEvent(st, ctx->e_in[1], 0));
    }
  }
  turn.end_copy(ctx->s_in);
  io.turn = &turn;
  uint64_t total = 0;
  const int rc = deflate_run(ctx, io, &total);
  *out_len = total;
  if (rc != FB200_OK) {
    cudaStreamSynchronize(ctx->s_in);
    return rc;
  }
  if (dst_off) CK(cudaMemcpyAsync(dst_off, ctx->p_off_out[0].p, (ns + 1) * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaStreamSynchronize(ctx->s_in));
  ctx->stats.kernel_launches += 1;
  turn.trace("deflate", ctx->t_kernels_done);
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

struct InflateHooks { // host-buffer calls only
  const uint32_t *hist0 = nullptr; // preset dictionaries (device array), see InflateJob::hist0
  const uint32_t *avail = nullptr;
  uint32_t *group_done = nullptr;
  volatile uint32_t *group_flag = nullptr;
  uint32_t group_streams = 0;
};

static int inflate_launch(fb200_ctx *ctx, const uint8_t *d_comp, const uint64_t *d_comp_off, uint64_t nstreams,
                          uint8_t *d_out, const uint64_t *d_out_off, uint64_t *d_out_len, int32_t *d_status,
                          int64_t *d_err_off, uint64_t *d_consumed, uint64_t cap_total_known, const InflateHooks &hk)
{
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
  j.avail = hk.avail;
  j.cta_streams = ctx->cta_streams;
  j.hist0 = hk.hist0;
  j.group_done = hk.group_done;
  j.group_flag = hk.group_flag;
  j.group_streams = hk.group_streams;
  for (int i = 0; i < FB200_NUM_STAGES; i++) ctx->ev_used[i] = false;
  uint64_t launches = nstreams ? 2 : 0;
  if (nstreams) {
    // record areas: sized from the output capacity (a match yields >= 3 bytes)
    uint64_t cap_total = cap_total_known;
    if (cap_total == ~0ull) {
      CK(cudaMemcpyAsync(ctx->pinned, d_out_off, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(ctx->pinned + 1, d_out_off + nstreams, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      cap_total = ctx->pinned[1] - ctx->pinned[0];
    }
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
    launches += 1;
    // longest streams first -- unless a host-buffer call wants output groups to finish in index order
    if (hk.group_done) j.order = nullptr; // (index order on the mixed corpus: 10.7 ms per GiB against 9.8)
    else { launch_stream_order(j, ctx->i_order_hist.as<uint32_t>(), st); launches += 3; }
  }
  ctx->stage_begin(FB200_STAGE_INFLATE);
  launch_inflate3(j, ctx->num_sms, st);
  launch_inflate_exact(j, ctx->num_sms, st);
  ctx->stage_end(FB200_STAGE_INFLATE);
  CK(cudaGetLastError());
  ctx->stats = fb200_stats{};
  ctx->stats.kernel_launches = launches;
  return FB200_OK;
}

static int inflate_finish(fb200_ctx *ctx)
{
  // (queued here, not behind the kernels in inflate_launch: a host-buffer call records its "kernels done" event for
  // the next call's kernels right after the launch, and this small copy waits behind bulk D2H copies in the copy engine)
  CK(cudaMemcpyAsync(ctx->pinned, ctx->counters.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stats.inflate_fallbacks = reinterpret_cast<const uint32_t *>(ctx->pinned)[2];
  if (getenv("FB200_TRACE"))
    fprintf(stderr, "[fb200] inflate: rounds=%u blocks=%u\n", reinterpret_cast<const uint32_t *>(ctx->pinned)[5],
            reinterpret_cast<const uint32_t *>(ctx->pinned)[6]);
  return FB200_OK;
}

extern "C" int fb200_inflate_batch_dev(fb200_ctx *ctx, const uint8_t *d_comp, const uint64_t *d_comp_off,
                                       uint64_t nstreams, uint8_t *d_out, const uint64_t *d_out_off,
                                       uint64_t *d_out_len, int32_t *d_status, int64_t *d_err_off,
                                       uint64_t *d_consumed)
{
  if (!ctx || !d_comp_off || !d_out_off || !d_out_len || !d_status || !d_err_off) return FB200_ERR_ARG;
  int rc = inflate_launch(ctx, d_comp, d_comp_off, nstreams, d_out, d_out_off, d_out_len, d_status, d_err_off,
                          d_consumed, ~0ull, InflateHooks{});
  if (rc != FB200_OK) return rc;
  return inflate_finish(ctx);
}

// Host-buffer inflate.  One launch for the whole batch; the compressed input is copied in chunks on a second
// stream that advances the arrival watermark the kernel waits on, and the kernel publishes finished output
// groups through host-visible flags, so the D2H copy of group g runs while later groups are still being
// decoded.  If any stream had to be re-decoded by the exact kernel, the output is copied once more at the end.
extern "C" int fb200_inflate_batch(fb200_ctx *ctx, const uint8_t *comp, const uint64_t *comp_off, uint64_t nstreams,
                                   uint8_t *out, const uint64_t *out_off, uint64_t *out_len, int32_t *status,
                                   int64_t *err_off, uint64_t *consumed)
{
  if (!ctx || !comp_off || !out_off || !out_len || !status || !err_off) return FB200_ERR_ARG;
  for (uint64_t i = 0; i < nstreams; i++)
    if (comp_off[i + 1] < comp_off[i] || out_off[i + 1] < out_off[i]) { ctx->err = "offsets not monotone"; return FB200_ERR_ARG; }
  const uint64_t c0 = nstreams ? comp_off[0] : 0, o0 = nstreams ? out_off[0] : 0;
  const uint64_t nc = nstreams ? comp_off[nstreams] - c0 : 0;
  const uint64_t no = nstreams ? out_off[nstreams] - o0 : 0;
  if ((!comp && nc) || (!out && no)) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->stats = fb200_stats{};
  if (nstreams == 0) return FB200_OK;
  if (nstreams > 0xfffffff0ull) return FB200_ERR_ARG;
  cudaStream_t st = ctx->stream;
  // output groups: ~chunk_bytes of output capacity each, at most kMaxChunks of them
  uint64_t gs = no ? (uint64_t)((double)ctx->chunk_bytes / ((double)no / (double)nstreams)) : nstreams;
  if (gs == 0) gs = 1;
  if ((nstreams + gs - 1) / gs > (uint64_t)fb200_ctx::kMaxChunks) gs = (nstreams + fb200_ctx::kMaxChunks - 1) / fb200_ctx::kMaxChunks;
  const uint64_t ngroups = (nstreams + gs - 1) / gs;
  CK(ctx->p_in[0].ensure(nc + 256));
  CK(ctx->p_out[0].ensure(no + 16));
  CK(ctx->p_off_in[0].ensure((nstreams + 1) * 8));
  CK(ctx->p_off_out[0].ensure((nstreams + 1) * 8));
  CK(ctx->p_len[0].ensure((nstreams + 1) * 8));
  CK(ctx->p_status[0].ensure((nstreams + 1) * 4));
  CK(ctx->p_eoff[0].ensure((nstreams + 1) * 8));
  CK(ctx->p_cons[0].ensure((nstreams + 1) * 8));
  CK(ctx->all_off.ensure((nstreams + 1) * 8));
  CK(ctx->all_off2.ensure((nstreams + 1) * 8));
  CK(ctx->group_done.ensure(ngroups * 4));
  uint8_t *d_comp = ctx->p_in[0].as<uint8_t>();
  uint8_t *d_out = ctx->p_out[0].as<uint8_t>();
  CK(cudaMemsetAsync(ctx->d_wm, 0, 8, st));
  CK(cudaMemsetAsync(ctx->group_done.p, 0, ngroups * 4, st));
  for (uint64_t g = 0; g < ngroups; g++) ctx->h_flags[g] = 0;
  CK(cudaEventRecord(ctx->e_comp[0], st));
  CK(cudaStreamWaitEvent(ctx->s_in, ctx->e_comp[0], 0));
  QueueTurn turn(ctx->device);
  turn.begin_copy(ctx->s_in);
  CK(cudaMemcpyAsync(ctx->all_off.p, comp_off, (nstreams + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
  CK(cudaMemcpyAsync(ctx->all_off2.p, out_off, (nstreams + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
  CK(cudaEventRecord(ctx->e_in[0], ctx->s_in));
  CK(cudaStreamWaitEvent(st, ctx->e_in[0], 0));
  launch_affine_u64(ctx->p_off_in[0].as<uint64_t>(), ctx->all_off.as<uint64_t>(), nstreams + 1, 0ull - c0, st);
  launch_affine_u64(ctx->p_off_out[0].as<uint64_t>(), ctx->all_off2.as<uint64_t>(), nstreams + 1, 0ull - o0, st);
  static const bool trace = getenv("FB200_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  double t_first = 0, t_last = 0, t_fed = 0;
  // the feed first: input chunks (boundaries rounded up to 128 bytes) + watermark updates on the copy stream,
  // all queued before the kernel that waits on them is launched (see fb200_ctx::overlap_h2d)
  {
    uint64_t step = ctx->chunk_bytes;
    if (nc / step + 2 > (uint64_t)fb200_ctx::kMaxChunks) step = nc / (fb200_ctx::kMaxChunks - 2) + 1;
    uint64_t done = 0, next_cut = step;
    size_t c = 0;
    for (uint64_t i = 0; i < nstreams; i++) {
      const uint64_t endb = comp_off[i + 1] - c0;
      if (endb >= next_cut || i + 1 == nstreams) {
        uint64_t upto = i + 1 == nstreams ? nc : ((endb + 127) & ~127ull);
        if (upto > nc) upto = nc;
        if (upto > done) {
          CK(cudaMemcpyAsync(d_comp + done, comp + c0 + done, upto - done, cudaMemcpyHostToDevice, ctx->s_in));
          done = upto;
        }
        if (ctx->overlap_h2d) {
          ctx->wm_vals[c] = (uint32_t)(i + 1);
          CK(cudaMemcpyAsync(ctx->d_wm + 1, &ctx->wm_vals[c], 4, cudaMemcpyHostToDevice, ctx->s_in));
        }
        c++;
        next_cut = endb + step;
      }
    }
    if (!ctx->overlap_h2d) { // copy, then launch
      CK(cudaEventRecord(ctx->e_in[1], ctx->s_in));
      CK(cudaStreamWaitEvent(st, ctx->e_in[1], 0));
    }
  }
  turn.end_copy(ctx->s_in);
  t_fed = now();
  turn.begin_compute(st);
  InflateHooks hk;
  hk.avail = ctx->overlap_h2d ? ctx->d_wm + 1 : nullptr;
  hk.group_done = ctx->group_done.as<uint32_t>();
  hk.group_flag = ctx->h_flags;
  hk.group_streams = (uint32_t)gs;
  int rc = inflate_launch(ctx, d_comp, ctx->p_off_in[0].as<uint64_t>(), nstreams, d_out, ctx->p_off_out[0].as<uint64_t>(),
                          ctx->p_len[0].as<uint64_t>(), ctx->p_status[0].as<int32_t>(), ctx->p_eoff[0].as<int64_t>(),
                          ctx->p_cons[0].as<uint64_t>(), no, hk);
  if (rc != FB200_OK) return rc;
  turn.end_compute(st);
  {
    // drain: copy every output group back as soon as the device reports it finished
    for (uint64_t g = 0; g < ngroups; g++) {
      unsigned spins = 0;
      while (ctx->h_flags[g] == 0) {
        if ((++spins & 0x3ff) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) break; // finished, or failed
      }
      if (ctx->h_flags[g] == 0) break; // the stream ended without the flag: error path below reports it
      const uint64_t a = g * gs, b = a + gs < nstreams ? a + gs : nstreams;
      const uint64_t ob = out_off[b] - out_off[a];
      if (ob) CK(cudaMemcpyAsync(out + out_off[a], d_out + (out_off[a] - o0), ob, cudaMemcpyDeviceToHost, ctx->s_out));
      if (g == 0) t_first = now();
      t_last = now();
    }
  }
  rc = inflate_finish(ctx);
  if (rc != FB200_OK) return rc;
  const double t_kdone = now();
  if (ctx->stats.inflate_fallbacks) { // re-copy the whole output: exact-kernel results came last
    CK(cudaStreamSynchronize(ctx->s_out));
    if (no) CK(cudaMemcpyAsync(out + o0, d_out, no, cudaMemcpyDeviceToHost, st));
  } else {
    for (uint64_t g = 0; g < ngroups; g++)
      if (ctx->h_flags[g] == 0) { ctx->err = "internal error: output group never completed"; return FB200_ERR_CUDA; }
  }
  CK(cudaMemcpyAsync(out_len, ctx->p_len[0].p, nstreams * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, ctx->p_status[0].p, nstreams * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(err_off, ctx->p_eoff[0].p, nstreams * 8, cudaMemcpyDeviceToHost, st));
  if (consumed) CK(cudaMemcpyAsync(consumed, ctx->p_cons[0].p, nstreams * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaStreamSynchronize(ctx->s_in));
  CK(cudaStreamSynchronize(ctx->s_out));
  ctx->stats.kernel_launches += 2;
  turn.trace("inflate", t_kdone - t0 + turn.t_call);
  if (trace)
    fprintf(stderr, "[fb200] inflate host: groups=%llu fed=%.2f first_flag=%.2f last_flag=%.2f kernels_done=%.2f end=%.2f ms\n",
            (unsigned long long)ngroups, t_fed - t0, t_first - t0, t_last - t0, t_kdone - t0, now() - t0);
  return FB200_OK;
}

// ------------------------------------------------------------------
// Asynchronous forms of the host-buffer calls.  Deflate and inflate of different batches overlap when they run
// on two contexts: the H2D copy of one call travels beside the D2H copy of the other (PCIe is full duplex) and
// the kernels of both share the SMs.  The blocking call needs the host while it runs (it issues the D2H copy of
// every output group as soon as the device reports its size), so the asynchronous form runs it on a helper
// thread of the context; buffers and result pointers must stay valid until fb200_wait returns.
#define ASYNC_BEGIN(ctx)                                                                     \
  do {                                                                                       \
    if (!(ctx)) return FB200_ERR_ARG;                                                        \
    if ((ctx)->async_pending) { (ctx)->err = "an asynchronous call is still in flight: fb200_wait first"; return FB200_ERR_ARG; } \
    if ((ctx)->worker.joinable()) (ctx)->worker.join();                                      \
    (ctx)->async_pending = true;                                                             \
  } while (0)

extern "C" int fb200_deflate_segments_async(fb200_ctx *ctx, const uint8_t *src, uint64_t n, uint64_t seg_size,
                                            uint8_t *dst, uint64_t dst_cap, uint64_t *seg_off, uint64_t *out_len)
{
  ASYNC_BEGIN(ctx);
  ctx->worker = std::thread([=] { ctx->async_rc = fb200_deflate_segments(ctx, src, n, seg_size, dst, dst_cap, seg_off, out_len); });
  return FB200_OK;
}

extern "C" int fb200_deflate_streams_async(fb200_ctx *ctx, const uint8_t *src, const uint64_t *src_off, uint64_t nstreams,
                                           uint8_t *dst, uint64_t dst_cap, uint64_t *dst_off, uint64_t *out_len)
{
  ASYNC_BEGIN(ctx);
  ctx->worker = std::thread([=] { ctx->async_rc = fb200_deflate_streams(ctx, src, src_off, nstreams, dst, dst_cap, dst_off, out_len); });
  return FB200_OK;
}

extern "C" int fb200_inflate_batch_async(fb200_ctx *ctx, const uint8_t *comp, const uint64_t *comp_off, uint64_t nstreams,
                                         uint8_t *out, const uint64_t *out_off, uint64_t *out_len, int32_t *status,
                                         int64_t *err_off, uint64_t *consumed)
{
  ASYNC_BEGIN(ctx);
  ctx->worker = std::thread(
      [=] { ctx->async_rc = fb200_inflate_batch(ctx, comp, comp_off, nstreams, out, out_off, out_len, status, err_off, consumed); });
  return FB200_OK;
}

extern "C" int fb200_wait(fb200_ctx *ctx)
{
  if (!ctx) return FB200_ERR_ARG;
  if (!ctx->async_pending) return FB200_OK;
  if (ctx->worker.joinable()) ctx->worker.join();
  ctx->async_pending = false;
  return ctx->async_rc;
}

// ------------------------------------------------------------------
// Streaming Writer (writer.mbt:10-58 / deflate.mbt:157-183, :280-294).
// Compressor::write copies its input into a 65535-byte window and encodes the window the moment it is full
// (fill_store / enc_speed, deflate.mbt:222-294); only the last, partial window waits for close.  The object does
// the same at the granularity of a write call: every write compresses the full windows it has completed and hands
// their bytes to the sink before it returns, so at most one window (+ the bits of an unfinished byte) stays
// buffered, whatever the length of the stream.  What carries over from call to call is what the reference's
// encoder carries over from block to block: the hash table (as the normalised end table of the block-parallel
// parse, on the device), the last 32768 bytes (a candidate in the previous block is verified on 4 bytes, D1), the
// number of blocks so far (table resets, deflate-fast.mbt:129-132) and the bit position inside the last byte.

struct fb200_writer {
  fb200_ctx *ctx;
  int device = 0;             // (kept here: the object may be freed after its context)
  fb200_sink_fn sink;
  void *user;
  std::vector<uint8_t> pend;  // bytes written and not yet compressed (< 65535 between calls)
  std::vector<uint8_t> hist;  // the last <= 32768 bytes already compressed
  std::vector<uint8_t> stage, out;
  std::vector<uint8_t> outq;  // sink == NULL: compressed bytes waiting for fb200_writer_take
  bool streaming = false;     // some windows have been compressed: the device holds the continuation state
  uint64_t blocks_done = 0;
  uint32_t carry_bits = 0;    // bits of `carry` that belong to the stream so far
  uint8_t carry = 0;
  uint16_t *d_seed = nullptr; // device: end table of the last compressed block
  bool closed = false;
  int sticky = 0;
};

static int writer_deliver(fb200_writer *w, const uint8_t *p, uint64_t n)
{
  if (!n) return 0;
  if (w->sink) return w->sink(w->user, p, n);
  w->outq.insert(w->outq.end(), p, p + n);
  return 0;
}

extern "C" fb200_writer *fb200_writer_new(fb200_ctx *ctx, fb200_sink_fn sink, void *user)
{
  if (!ctx) return nullptr;
  fb200_writer *w = new (std::nothrow) fb200_writer();
  if (!w) return nullptr;
  w->ctx = ctx;
  w->device = ctx->device;
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
  w->pend.assign(dict, dict + n);
  return w;
}

// Compresses data[0..n) as the continuation of the writer's stream: n is a multiple of 65535 unless `final`, which
// also writes the final empty stored block (Compressor::close, deflate.mbt:157-183).
static int writer_emit(fb200_writer *w, const uint8_t *data, uint64_t n, bool final)
{
  fb200_ctx *ctx = w->ctx;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!w->d_seed) {
    CK(cudaMalloc((void **)&w->d_seed, (size_t)kTableSize * 2));
    CK(cudaMemsetAsync(w->d_seed, 0, (size_t)kTableSize * 2, st)); // nothing in reach
  }
  // [stand-in block: 65535 bytes, the history at its end][data]
  const uint64_t total = (uint64_t)kBlockSize + n;
  w->stage.resize(total);
  memset(w->stage.data(), 0, (size_t)kBlockSize - w->hist.size());
  if (!w->hist.empty()) memcpy(w->stage.data() + kBlockSize - w->hist.size(), w->hist.data(), w->hist.size());
  if (n) memcpy(w->stage.data() + kBlockSize, data, n);
  const uint64_t dcap = ((fb200_deflate_stream_bound(n) + 64 + 3) & ~3ull);
  CK(ctx->p_in[1].ensure(total + 256));
  CK(ctx->p_out[1].ensure(dcap + 16));
  CK(ctx->p_off_in[1].ensure(64));
  CK(ctx->p_off_out[1].ensure(64));
  uint64_t *h = ctx->pinned + 56;
  h[0] = 0;
  h[1] = total;
  CK(cudaMemcpyAsync(ctx->p_in[1].p, w->stage.data(), total, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->p_off_in[1].p, h, 16, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st)); // (the staging vector is pageable)
  DeflateIo io;
  io.d_src = ctx->p_in[1].as<uint8_t>();
  io.d_src_off = ctx->p_off_in[1].as<uint64_t>();
  io.ns = 1;
  io.n_total = total;
  io.d_dst_off = ctx->p_off_out[1].as<uint64_t>();
  io.d_dst = ctx->p_out[1].as<uint8_t>();
  io.d_cap = dcap;
  io.cont_prev = true;
  io.cont_open = !final;
  io.cont_start_bit = w->carry_bits;
  io.cont_block_base = w->blocks_done;
  io.d_seed = w->d_seed;
  uint64_t olen = 0;
  const int rc = deflate_run(ctx, io, &olen);
  if (rc != FB200_OK) return rc;
  uint64_t bits_end = olen * 8;
  if (!final) { // where the last block ended inside the last byte
    CK(cudaMemcpyAsync(h + 2, ctx->last.stream_trailer_bit, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    bits_end = h[2];
  }
  w->out.resize(olen + 1);
  if (olen) CK(cudaMemcpy(w->out.data(), ctx->p_out[1].p, olen, cudaMemcpyDeviceToHost));
  if (w->carry_bits && olen) w->out[0] |= w->carry; // the bits the previous call left in this byte
  uint64_t emit = olen;
  if (!final) {
    w->carry_bits = (uint32_t)(bits_end & 7);
    if (w->carry_bits) { emit = olen - 1; w->carry = w->out[olen - 1]; }
    else w->carry = 0;
    w->blocks_done += n / kBlockSize;
    // history for the next call: the last 32768 bytes of everything compressed so far
    if (n >= (uint64_t)kMaxMatchOffset) w->hist.assign(data + n - kMaxMatchOffset, data + n);
    else {
      w->hist.insert(w->hist.end(), data, data + n);
      if (w->hist.size() > (size_t)kMaxMatchOffset) w->hist.erase(w->hist.begin(), w->hist.end() - kMaxMatchOffset);
    }
    w->streaming = true;
  }
  if (writer_deliver(w, w->out.data(), emit) != 0) return FB200_ERR_ARG;
  return FB200_OK;
}

extern "C" int64_t fb200_writer_write(fb200_writer *w, const uint8_t *data, uint64_t n)
{
  if (!w) return FB200_ERR_ARG;
  if (w->closed) return FB200_ERR_CLOSED; // writer_closed_error (deflate.mbt:154, :281-283)
  if (w->sticky) return w->sticky;
  if (n) w->pend.insert(w->pend.end(), data, data + n);
  const uint64_t full = w->pend.size() / kBlockSize * kBlockSize; // windows that are full now (deflate.mbt:284-291)
  if (full) {
    const int rc = writer_emit(w, w->pend.data(), full, false);
    if (rc != FB200_OK) { w->sticky = rc; return rc; }
    w->pend.erase(w->pend.begin(), w->pend.begin() + (ptrdiff_t)full);
  }
  return (int64_t)n;
}

extern "C" int fb200_writer_close(fb200_writer *w)
{
  if (!w) return FB200_ERR_ARG;
  if (w->closed) return FB200_OK; // deflate.mbt:158-160
  if (w->sticky) return w->sticky;
  int rc;
  if (!w->streaming) { // nothing compressed yet: the whole stream in one call
    const uint64_t n = w->pend.size();
    w->out.resize(fb200_deflate_stream_bound(n));
    uint64_t off[2] = {0, n}, doff[2], olen = 0;
    rc = fb200_deflate_streams(w->ctx, w->pend.data(), off, 1, w->out.data(), w->out.size(), doff, &olen);
    if (rc == FB200_OK && writer_deliver(w, w->out.data(), olen) != 0) rc = FB200_ERR_ARG;
  } else {
    rc = writer_emit(w, w->pend.data(), w->pend.size(), true);
  }
  if (rc != FB200_OK) { w->sticky = rc; return rc; }
  w->closed = true;
  std::vector<uint8_t>().swap(w->pend);
  std::vector<uint8_t>().swap(w->stage);
  std::vector<uint8_t>().swap(w->out);
  return FB200_OK;
}

extern "C" uint64_t fb200_writer_pending(const fb200_writer *w) { return w ? w->outq.size() : 0; }

extern "C" uint64_t fb200_writer_take(fb200_writer *w, uint8_t *dst, uint64_t cap)
{
  if (!w || (!dst && cap)) return 0;
  const uint64_t k = w->outq.size() < cap ? w->outq.size() : cap;
  if (k) {
    memcpy(dst, w->outq.data(), k);
    w->outq.erase(w->outq.begin(), w->outq.begin() + (ptrdiff_t)k);
  }
  return k;
}

extern "C" void fb200_writer_free(fb200_writer *w)
{
  if (!w) return;
  if (w->d_seed) {
    if (cudaSetDevice(w->device) != cudaSuccess || cudaFree(w->d_seed) != cudaSuccess) cudaGetLastError();
  }
  delete w;
}

// Streaming Decompressor (inflate.mbt:257-418).  The first read inflates the
// whole stream on the GPU; reads then hand the result out with the
// reference's granularity: one 32 KiB window flush at a time, the final
// status riding on the read that drains the last (partial) flush.
struct fb200_reader {
  fb200_ctx *ctx;
  const uint8_t *comp;
  uint64_t n;
  std::vector<uint8_t> dict; // &Reader::new_dict / Decompressor::reset: the last <= 32768 bytes of the dictionary
  bool decoded = false;
  std::unique_ptr<uint8_t[]> out;
  uint64_t total = 0, pos = 0, consumed = 0;
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

// One stream with a preset dictionary (&Reader::new_dict, inflate.mbt:310-317): the dictionary tail sits directly
// in front of the output slot on the device (InflateJob::hist0), exactly like DictDecoder::new pre-loads the
// window (dict-decoder.mbt:42-60).
extern "C" int fb200_inflate_dict(fb200_ctx *ctx, const uint8_t *comp, uint64_t n, const uint8_t *dict, uint64_t dict_len,
                                  uint8_t *out, uint64_t cap, uint64_t *out_len, int32_t *status, int64_t *err_off,
                                  uint64_t *consumed)
{
  if (!ctx || (!comp && n) || (!dict && dict_len) || (!out && cap) || !out_len || !status || !err_off) return FB200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (dict_len > (uint64_t)kMaxMatchOffset) { // only the last 32768 bytes matter (dict-decoder.mbt:49-52)
    dict += dict_len - kMaxMatchOffset;
    dict_len = kMaxMatchOffset;
  }
  const uint64_t D = dict_len;
  const uint64_t Dpad = (D + 15) & ~15ull; // keep the slot 16-byte aligned
  CK(ctx->p_in[1].ensure(n + 256));
  CK(ctx->p_out[1].ensure(Dpad + cap + 16));
  CK(ctx->p_off_in[1].ensure(64));
  CK(ctx->p_off_out[1].ensure(64));
  CK(ctx->p_len[1].ensure(64));
  CK(ctx->p_status[1].ensure(64));
  CK(ctx->p_eoff[1].ensure(64));
  CK(ctx->p_cons[1].ensure(64));
  uint8_t *d_out = ctx->p_out[1].as<uint8_t>();
  uint64_t *h = ctx->pinned + 40; // staging of the small arrays
  h[0] = 0; h[1] = n;                    // comp_off
  h[2] = Dpad; h[3] = Dpad + cap;        // out_off
  reinterpret_cast<uint32_t *>(h + 4)[0] = (uint32_t)D; // hist0
  if (n) CK(cudaMemcpyAsync(ctx->p_in[1].p, comp, n, cudaMemcpyHostToDevice, s));
  if (D) CK(cudaMemcpyAsync(d_out + Dpad - D, dict, D, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->p_off_in[1].p, h, 16, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->p_off_out[1].p, h + 2, 16, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->p_cons[1].as<uint8_t>() + 32, h + 4, 4, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s)); // (comp / dict may be pageable: the staging must not be reused before they are read)
  InflateHooks hk;
  hk.hist0 = reinterpret_cast<const uint32_t *>(ctx->p_cons[1].as<uint8_t>() + 32);
  int rc = inflate_launch(ctx, ctx->p_in[1].as<uint8_t>(), ctx->p_off_in[1].as<uint64_t>(), 1, d_out,
                          ctx->p_off_out[1].as<uint64_t>(), ctx->p_len[1].as<uint64_t>(), ctx->p_status[1].as<int32_t>(),
                          ctx->p_eoff[1].as<int64_t>(), ctx->p_cons[1].as<uint64_t>(), cap, hk);
  if (rc != FB200_OK) return rc;
  rc = inflate_finish(ctx);
  if (rc != FB200_OK) return rc;
  CK(cudaMemcpyAsync(h + 8, ctx->p_len[1].p, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h + 9, ctx->p_status[1].p, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h + 10, ctx->p_eoff[1].p, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h + 11, ctx->p_cons[1].p, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *out_len = h[8];
  *status = reinterpret_cast<int32_t *>(h + 9)[0];
  *err_off = reinterpret_cast<int64_t *>(h + 10)[0];
  if (consumed) *consumed = h[11];
  if (*out_len) CK(cudaMemcpy(out, d_out + Dpad, *out_len, cudaMemcpyDeviceToHost));
  return FB200_OK;
}

static void reader_decode(fb200_reader *r)
{
  r->decoded = true;
  // One stream through fb200_inflate_dict (with or without a dictionary): the output comes back in one copy of exactly
  // the bytes produced.  The capacity is a guess (a deflate stream expands at most 1032 : 1) that grows on demand;
  // the host buffer is left uninitialised, untouched pages cost nothing.
  uint64_t cap = r->n * 8 + 65536;
  for (;;) {
    r->out.reset(new (std::nothrow) uint8_t[cap]);
    if (!r->out) { r->rc = FB200_ERR_NOMEM; r->status = FB200_ST_INTERNAL; r->total = 0; return; }
    uint64_t olen = 0, cons = 0;
    int32_t st = -1;
    int64_t eo = 0;
    r->rc = fb200_inflate_dict(r->ctx, r->comp, r->n, r->dict.empty() ? nullptr : r->dict.data(), r->dict.size(), r->out.get(), cap,
                               &olen, &st, &eo, &cons);
    if (r->rc != FB200_OK) { r->status = FB200_ST_INTERNAL; r->total = 0; return; }
    if (st == FB200_ST_DST_TOO_SMALL && cap < r->n * 1040 + 65536) { cap *= 4; continue; }
    r->total = olen;
    r->status = st;
    r->err_off = eo;
    r->consumed = cons;
    return;
  }
}

static void reader_set_dict(fb200_reader *r, const uint8_t *dict, uint64_t n)
{
  // DictDecoder::new keeps the last `size` (32768) bytes of the dictionary (dict-decoder.mbt:49-52)
  if (n > (uint64_t)kMaxMatchOffset) {
    dict += n - kMaxMatchOffset;
    n = kMaxMatchOffset;
  }
  if (n) r->dict.assign(dict, dict + n);
  else r->dict.clear();
}

extern "C" fb200_reader *fb200_reader_new_dict(fb200_ctx *ctx, const uint8_t *comp, uint64_t n, const uint8_t *dict,
                                               uint64_t dict_len)
{
  if (!dict && dict_len) return nullptr;
  fb200_reader *r = fb200_reader_new(ctx, comp, n);
  if (r) reader_set_dict(r, dict, dict_len);
  return r;
}

extern "C" int fb200_reader_reset(fb200_reader *r, const uint8_t *comp, uint64_t n, const uint8_t *dict, uint64_t dict_len)
{
  if (!r || (!comp && n) || (!dict && dict_len)) return FB200_ERR_ARG;
  r->comp = comp;
  r->n = n;
  r->decoded = false;
  r->total = r->pos = r->consumed = 0;
  r->status = -1;
  r->err_off = 0;
  r->rc = FB200_OK;
  reader_set_dict(r, dict, dict_len);
  return FB200_OK;
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
  // DictDecoder::new starts with wr_pos = the dictionary length (mod the window, dict-decoder.mbt:56-62), so the
  // window flushes fall where (D + position) is a multiple of 32768
  const uint64_t window = (uint64_t)kMaxMatchOffset;
  const uint64_t D = r->dict.size() % window;
  uint64_t chunk_end = ((r->pos + D) / window + 1) * window - D;
  if (chunk_end > r->total) chunk_end = r->total;
  uint64_t k = chunk_end - r->pos;
  if (k > n) k = n;
  memcpy(buf, r->out.get() + r->pos, k);
  r->pos += k;
  *status = -1;
  // the status rides on the read that drains the final, partial window flush (inflate.mbt:392-396)
  if (r->pos == r->total && ((r->total + D) % window) != 0) {
    *status = r->status;
    if (err_off) *err_off = r->err_off;
  }
  return k;
}

extern "C" uint64_t fb200_reader_consumed(fb200_reader *r)
{
  if (!r) return 0;
  if (!r->decoded) reader_decode(r);
  return r->consumed;
}

extern "C" int fb200_reader_close(fb200_reader *r)
{
  if (!r) return FB200_ERR_ARG;
  if (!r->decoded) return FB200_OK; // err is still None
  if (r->status == FB200_ST_EOF || r->status == FB200_ST_EOF_AT_REFILL || r->status < 0) return FB200_OK;
  return r->status;
}

extern "C" void fb200_reader_free(fb200_reader *r) { delete r; }
