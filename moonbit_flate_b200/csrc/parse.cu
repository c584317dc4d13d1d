// parse.cu -- K1: exact greedy LZ77 parse of DeflateFast::encode
// (deflate-fast.mbt:123-342), one warp per independent stream.
//
// The reference walks one position at a time: probe the 1<<14-entry hash table,
// insert the current position, test the candidate, on a hit extend the match,
// then insert s-1 / probe s (deflate-fast.mbt:246-265).  The token sequence is
// a pure function of the bytes, so it can be reproduced 32 probes at a time:
//
//  * After a match (or at block start) the probe positions are a fixed
//    schedule: probe k of a probe loop sits at p0 + d_k with d_0 = 0,
//    d_{k+1} = d_k + 1 + (d_k >> 5)  (skip = 32 + d_k, step = skip >> 5,
//    deflate-fast.mbt:178-187).  Lane l of a batch takes probe k0 + l.
//  * A "post-match" batch folds the reference's insert(s-1) and probe(s) into
//    lanes 0 and 1 and starts the new probe loop at lane 2.
//  * All lanes read their bucket, then same-bucket dependencies inside the
//    batch are resolved with __match_any_sync: a lane's candidate is the
//    nearest lower lane with the same bucket, else the old table entry --
//    exactly what sequential execution would have read.
//  * __ballot_sync finds the first lane that hits (distance <= 32768 and the
//    four bytes equal, :195-196) or runs past s_limit (:188); lanes up to and
//    including a hit lane commit their insert (the reference writes the table
//    before testing, :191-193), the highest committed lane per bucket wins.
//  * match_len (:286-342) compares 32 bytes per step with ballot/ffs.  A
//    candidate in the previous block yields length 0 beyond the 4 hashed
//    bytes because the reference never populates `prev` (quirk D1).
//
// The table lives in shared memory and stores positions only: `val` of the
// reference's TableEntry is by construction load32(src, position), so it is
// re-read from the source (L1/L2 resident).  Streams with a single parsed
// block use uint16 positions (32 KiB per warp); longer streams, whose table
// persists across blocks, use uint32 absolute positions + 1 (0 = empty).
#include "common.cuh"
#include "kernels.h"

#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace fb {

constexpr int kSchedLen = 512;
__device__ uint32_t g_sched[kSchedLen];

__global__ void k_init_sched()
{
  uint32_t d = 0;
  for (int k = 0; k < kSchedLen; k++) {
    g_sched[k] = d;
    d = d + 1 + (d >> 5);
    if (d > (1u << 20)) d = 1u << 20;
  }
}

constexpr unsigned kFull = 0xffffffffu;
#ifndef FB_PARSE_SW
#define FB_PARSE_SW 5  // default warps per SM with a shared-memory table
#endif
#ifndef FB_PARSE_GW
#define FB_PARSE_GW 25 // default warps per SM with a global-memory table
#endif
// Look-ahead of the post-match batches of the global-table warps: the buckets the positions FB_PF_DIST bytes
// ahead hash to are prefetched into L2, so that the next batch's table read is not a trip to DRAM.
#ifndef FB_PF_DIST
#define FB_PF_DIST 32 // measured: 24 -> 18.77 ms, 32 -> 18.48, 48 -> 18.72, 64 -> 18.80
#endif

// match_len tail (deflate-fast.mbt:286-307): number of equal bytes of src[s2..] and src[t..] (t < s2), at most a,
// given that the first `from` already matched.  128 bytes per step: lane l compares the four bytes at offset
// off + 4 l of both sides as one word, so the longest match (258) takes two trips to memory and a match that ends
// in the first 128 bytes one.  n = bytes in the block (words that would reach beyond it are read byte by byte).
__device__ __forceinline__ int match_tail(const uint8_t *srcb, int s2, int t, int a, int from, int lane, int n)
{
  for (int off = from; off < a; off += 128) {
    const int i = off + 4 * lane;
    uint32_t x = 0;
    if (i < a) {
      if (s2 + i + 8 <= n) {
        x = ld32u(srcb + s2 + i) ^ ld32u(srcb + t + i);
        if (i + 4 > a) x &= (1u << (8 * (a - i))) - 1u;
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (i + k < a) x |= (uint32_t)(__ldg(srcb + s2 + i + k) ^ __ldg(srcb + t + i + k)) << (8 * k);
      }
    }
    const unsigned mm = __ballot_sync(kFull, i >= a || x != 0);
    if (mm) {
      const int l = __ffs(mm) - 1;
      const uint32_t xl = __shfl_sync(kFull, x, l);
      const int il = off + 4 * l;
      return xl ? il + ((__ffs(xl) - 1) >> 3) : a; // (xl == 0: lane l lies beyond the limit)
    }
  }
  return a;
}

template <bool MULTI, typename T, bool GTAB>
__device__ __forceinline__ void parse_worker(const DeflateJob &j, uint32_t *counter, T *table, const BlockParJob *bp = nullptr)
{
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1;
  const unsigned gt_mask = ~((2u << lane) - 1); // lanes above this one (0 for lane 31)

  for (;;) {
    uint32_t st32 = 0;
    if (lane == 0) st32 = atomicAdd(counter, 1u);
    st32 = __shfl_sync(kFull, st32, 0);
    uint64_t bp_gb = 0; // block-parallel rounds: the work item is one block of a multi-block stream
    if (MULTI && bp) {
      if (st32 >= bp->nlist) break;
      bp_gb = bp->list[st32];
      st32 = j.blk_stream[bp_gb];
    }
    if (st32 >= j.nstreams) break;
    if (j.avail) { // host-buffer call: wait until the H2D stream has delivered this stream's bytes
      if (lane == 0)
        while (*(volatile const uint32_t *)j.avail <= st32) __nanosleep(500);
      __syncwarp();
    }
    const uint64_t o0 = j.stream_off[st32];
    const uint64_t L = j.stream_off[st32 + 1] - o0;
    const bool is_multi = L >= (uint64_t)kBlockSize + 128;
    if (is_multi != MULTI) continue;
    if (L < 128) continue; // nothing to parse: at most one small block

    const uint64_t blk0 = j.stream_blk0[st32];
    const uint32_t nblk = (uint32_t)((L + kBlockSize - 1) / kBlockSize);
    const uint32_t bp_b = (MULTI && bp) ? (uint32_t)(bp_gb - blk0) : 0u;
    if (j.cont_prev && !(MULTI && bp)) continue; // a continued stream is parsed by the block-parallel rounds only
    // round 1: empty tables -- except behind the stand-in block of a continued stream, whose end table is given
    const uint64_t abs_b = (MULTI && bp && bp->cont_prev) ? bp->cont_block_base + bp_b - 1 : bp_b; // index in the whole stream
    const bool bp_seed = MULTI && bp && bp_b > 0 && !block_resets_table(abs_b) &&
                         (bp->round > 1 || (bp->cont_prev && bp_b == 1));
    if (!bp_seed) { // DeflateFast::new (:111-117): empty table
      const uint32_t fill = MULTI ? 0u : 0xffffffffu;
      uint4 *t4 = reinterpret_cast<uint4 *>(table);
      const int n4 = (int)(kTableSize * sizeof(T) / 16);
      for (int i = lane; i < n4; i += 32) t4[i] = make_uint4(fill, fill, fill, fill);
      __syncwarp();
    } else if (MULTI) {
      // block-parallel round: the table as the previous block left it, from that block's normalised end table
      // (distance of the last position of every bucket from the block end, 1..32768, 0 = none in reach).
      // Positions are kept relative to (block start - 32768): the entry is 32768 - distance + 1.
      const uint64_t mp = bp->mb_idx[bp_gb - 1];
      const uint16_t *seed = bp->tabs + ((size_t)bp->lat_prev[mp] * bp->nmb + mp) * kTableSize;
      for (int i = lane; i < kTableSize; i += 32) {
        const uint32_t d = seed[i];
        table[i] = (T)(d ? (uint32_t)kMaxMatchOffset - d + 1u : 0u);
      }
      __syncwarp();
    }
    // DeflateFast.cur (:95-117): 65535 at the start, + the block length after every block.  When it reaches
    // buffer_reset (about every 2 GiB of one stream) encode calls shift_offsets (:129-132), which -- `prev` being
    // always empty, D1 -- clears the table and restarts cur at 32769 (:366-374).  Table positions are kept
    // relative to the last such point, so they neither wrap nor collide with the "empty" value.
    int64_t ref_cur = kBlockSize;
    int64_t pos_base = (MULTI && bp) ? (int64_t)bp_b * kBlockSize - kMaxMatchOffset : 0;
    uint32_t b = bp_b;
    const uint32_t b_end = (MULTI && bp) ? bp_b + 1 : nblk;
    for (; b < b_end; b++) {
      const uint64_t boff = (uint64_t)b * kBlockSize;
      const int n = (int)((L - boff) < (uint64_t)kBlockSize ? (L - boff) : (uint64_t)kBlockSize);
      if (n < 128) break; // small tail: not parsed (deflate.mbt:244-257)
      if (MULTI && !bp && ref_cur >= (int64_t)kBufferReset) {
        uint4 *t4 = reinterpret_cast<uint4 *>(table);
        const int n4 = (int)(kTableSize * sizeof(T) / 16);
        for (int i = lane; i < n4; i += 32) t4[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        ref_cur = kMaxMatchOffset + 1;
        pos_base = (int64_t)boff;
      }
      ref_cur += n;
      const uint8_t *srcb = j.src + o0 + boff;
      uint32_t *tok = j.tokens + o0 + boff;
      const uint32_t S0 = (uint32_t)((int64_t)boff - pos_base); // block start relative to the last table reset (MULTI)
      const int s_limit = n - kInputMargin;

      int s = 0, next_emit = 0;
      uint32_t ntok = 0;
      bool modeM = false;
      int loop_p0 = 0, k0 = 0;

      for (;;) {
        // ---- fast path: post-match batch, 32 consecutive positions s-1 .. s+30 ----
        // lane 0 = insert(s-1), lane 1 = probe(s), lanes 2.. = probes s+1.. (step 1).
        // One round evaluates every lane: its 4 bytes, bucket, old entry, the 4-byte
        // verify and up to 8 bytes of speculative match extension.  A speculative
        // insert + read-back finds W, the width of the lane prefix in which no two
        // lanes share a bucket (in each sharing group exactly one lane wins the
        // store; W = lowest losing lane).  For lanes < W the old entry is what
        // sequential execution reads whichever earlier lanes end up inserted, so the
        // prefix is consumed match after match from registers: literals are the low
        // byte each lane already holds, the next probe lane is s' - base, lanes
        // skipped by a match are simply not inserted.  (Exactness argument and CPU
        // emulation: tests/hostmodel/hostmodel.cu, fbm_parse_stream_v2.)
        if (modeM && s + 31 <= s_limit) {
          const int base = s - 1;
          const int pos = base + lane;
          const uint32_t cv = ld32u(srcb + pos);
          // (never beyond the block: with host-buffer calls the bytes after it may not have arrived yet)
          if (lane == 0 && pos + 256 < n) asm volatile("prefetch.global.L1 [%0];" ::"l"(srcb + pos + 256));
          const uint32_t h = hash4(cv);
          T *slot = table + h;
          const T old = *slot; // (ld.global.cg for the global tables: measured, no difference)
          const T mine = MULTI ? (T)(S0 + (uint32_t)pos + 1u) : (T)pos;
          if (GTAB && !MULTI && pos + FB_PF_DIST + 4 <= n)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(table + hash4(ld32u(srcb + pos + FB_PF_DIST))));
          unsigned conf;
          if (GTAB) { // table in global memory: no dependent read-back, compare buckets across lanes instead
            conf = __ballot_sync(kFull, (__match_any_sync(kFull, h) & lt_mask) != 0);
          } else {
            __syncwarp();
            *slot = mine;
          }
          int cand;
          bool ok;
          if (MULTI) {
            const uint32_t D = (uint32_t)mine - (uint32_t)old;
            ok = (old != 0) && (D <= (uint32_t)kMaxMatchOffset);
            cand = pos - (int)D;
          } else {
            cand = (int)((uint32_t)old & 0xffffu);
            ok = (uint32_t)(pos - cand - 1) < (uint32_t)kMaxMatchOffset;
          }
          ok = ok && (lane != 0);
          // 12 bytes at the candidate (own position when there is none: harmless L1 hit)
          uint32_t c0, c1, c2;
          {
            const uint8_t *cp = srcb + (ok ? cand : pos);
            const uintptr_t ca = (uintptr_t)cp;
            const uint32_t *q = (const uint32_t *)(ca & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(ca & 3) * 8;
            const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
            c0 = __funnelshift_r(w0, w1, sh);
            c1 = __funnelshift_r(w1, w2, sh);
            c2 = __funnelshift_r(w2, w3, sh);
          }
          const uint32_t p1 = __shfl_down_sync(kFull, cv, 4); // bytes pos+4 .. pos+7   (lanes <= 27)
          const uint32_t p2 = __shfl_down_sync(kFull, cv, 8); // bytes pos+8 .. pos+11  (lanes <= 23)
          const bool hit = ok && (c0 == cv);
          int avail = lane <= 23 ? 8 : (lane <= 27 ? 4 : 0);
          int extl = 0;
          {
            const uint32_t x1 = c1 ^ p1, x2 = c2 ^ p2;
            const int e1 = x1 ? ((__ffs(x1) - 1) >> 3) : 4;
            const int e2 = x2 ? ((__ffs(x2) - 1) >> 3) : 4;
            if (avail >= 4) extl = (e1 < 4 || avail == 4) ? e1 : 4 + e2;
            if (cand + 4 < 0) { extl = 0; avail = 99; } // candidate in the previous block: match_len == 0 (D1)
          }
          if (!GTAB) {
            __syncwarp();
            const T rb = *slot;
            conf = __ballot_sync(kFull, rb != mine);
            // final table state is written below: restore every bucket first
            *slot = old;
          }
          unsigned hitm = __ballot_sync(kFull, hit);
          const int W = conf ? __ffs(conf) - 1 : 32;
          if (W >= 2) {
            const unsigned wmask = (W < 32 ? (1u << W) : 0u) - 1u; // lanes below W
            hitm &= wmask;
            // Walk the matches of the prefix.  Only masks are updated per match: `keep` = lanes whose
            // insert happens, `emit` = lanes that produce a token (literal lanes and match lanes);
            // the tokens themselves are stored once, after the walk, at ntok + rank within `emit`.
            unsigned keep = 0, emit = 0;
            int my_ext = extl;
            int cur = 1;
            bool block_done = false;
            const int packed_mine = (extl << 1) | (extl == avail ? 1 : 0);
            for (;;) {
              const unsigned below = (1u << cur) - 1u;
              const unsigned hm = hitm & ~below;
              if (hm == 0) { // no further hit below W: lanes cur..W-1 are literals, probing continues at lane W
                keep |= wmask & ~(below >> 1);
                emit |= wmask & ~below;
                next_emit = base + W;
                modeM = false; loop_p0 = base + cur + 1; k0 = W - 1 - cur;
                break;
              }
              const int m = __ffs(hm) - 1;
              const unsigned upto = (2u << m) - 1u;
              keep |= upto & ~(below >> 1);
              emit |= upto & ~below; // literals cur..m-1, match at m
              const int packed = __shfl_sync(kFull, packed_mine, m);
              int ext = packed >> 1;
              const int s2 = base + m + 4;
              if (packed & 1) { // the speculative bytes all matched: keep comparing
                const int t = __shfl_sync(kFull, cand, m) + 4;
                int s1 = s2 + kMaxMatchLength - 4;
                if (s1 > n) s1 = n;
                ext = match_tail(srcb, s2, t, s1 - s2, ext, lane, n);
                if (lane == m) my_ext = ext;
              }
              s = s2 + ext;
              next_emit = s;
              if (s >= s_limit) { block_done = true; break; } // :236-238
              const int ncur = s - base;
              if (ncur >= W) break; // next batch starts at s (still post-match mode)
              cur = ncur;
            }
            if ((emit >> lane) & 1u) {
              uint32_t t = cv & 0xffu; // emit_literal (:273-279)
              if ((hitm >> lane) & 1u) // match_token(l + 4 - 3, s - t - 1) (:228-233)
                t = kMatchType + ((uint32_t)(my_ext + 1) << kLengthShift) + (uint32_t)(pos - cand - 1);
              __stcs(&tok[ntok + (uint32_t)__popc(emit & lt_mask)], t);
            }
            ntok += (uint32_t)__popc(emit);
            if (!GTAB) __syncwarp();
            if ((keep >> lane) & 1u) *slot = mine; // kept lanes share no bucket
            __syncwarp();
            if (block_done) break;
            continue;
          }
          __syncwarp(); // bucket shared by lanes 0/1: resolve this batch exactly below
        }
        // ---- generic path: lane's table operation in this batch ----
        int pos, step;
        bool probe, loopprobe;
        if (modeM) {
          if (lane == 0) { // insert(s-1), :246-251
            pos = s - 1; step = 0; probe = false; loopprobe = false;
          } else if (lane == 1) { // probe(s), :252-260
            pos = s; step = 0; probe = true; loopprobe = false;
          } else { // new probe loop from s+1 with skip = 32, :261-265 -> :178-202
            pos = s + 1 + (lane - 2); step = 1; probe = true; loopprobe = true;
          }
        } else {
          const int k = k0 + lane;
          const uint32_t d = k < 32 ? (uint32_t)k : (k < kSchedLen ? g_sched[k] : (1u << 20));
          pos = loop_p0 + (int)d;
          step = 1 + (int)(d >> 5);
          probe = true; loopprobe = true;
        }
        const bool fail = loopprobe && (pos + step > s_limit); // :188
        const bool active = !fail;

        uint32_t cv = 0, h = 0x10000u | (uint32_t)lane;
        T old = 0;
        if (active) {
          cv = ld32u(srcb + pos);
          h = hash4(cv);
          old = table[h];
        }
        const unsigned peers = __match_any_sync(kFull, h);
        const unsigned lower = peers & lt_mask;
        const int srcl = lower ? 31 - __clz(lower) : lane;
        const int ppos = __shfl_sync(kFull, pos, srcl);
        int cand;
        bool ok;
        if (lower) {
          cand = ppos;
          ok = (pos - cand) <= kMaxMatchOffset;
        } else if (MULTI) {
          const uint32_t D = (S0 + (uint32_t)pos + 1u) - (uint32_t)old;
          ok = (old != 0) && (D <= (uint32_t)kMaxMatchOffset);
          cand = pos - (int)D;
        } else {
          cand = (int)((uint32_t)old & 0xffffu);
          const int D = pos - cand;
          ok = (D >= 1) && (D <= kMaxMatchOffset);
        }
        bool hit = false;
        if (active && probe && ok) hit = (ld32u(srcb + cand) == cv); // :196
        const unsigned hitm = __ballot_sync(kFull, hit);
        const unsigned evt = hitm | __ballot_sync(kFull, fail);
        const int m = evt ? __ffs(evt) - 1 : 32;
        const bool mhit = evt && ((hitm >> m) & 1u);
        const unsigned cmask = (m == 32) ? kFull : (mhit ? ((2u << m) - 1u) : ((1u << m) - 1u));
        const bool committed = active && ((cmask >> lane) & 1u);
        const unsigned cm = __ballot_sync(kFull, committed);
        if (committed && (peers & cm & gt_mask) == 0) // last writer of this bucket in program order
          table[h] = MULTI ? (T)(S0 + (uint32_t)pos + 1u) : (T)pos;
        __syncwarp();

        if (m == 32) { // 32 misses: keep probing (:198-199)
          if (modeM) { modeM = false; loop_p0 = s + 1; k0 = 30; }
          else k0 += 32;
          continue;
        }
        if (!mhit) break; // next_s > s_limit -> emit_remainder (:189)

        const int s_hit = __shfl_sync(kFull, pos, m);
        const int c = __shfl_sync(kFull, cand, m);
        // emit_literal(src[next_emit:s]) (:207)
        for (int i = next_emit + lane; i < s_hit; i += 32) __stcs(&tok[ntok + (uint32_t)(i - next_emit)], (uint32_t)__ldg(srcb + i));
        ntok += (uint32_t)(s_hit - next_emit);
        // match_len (:286-307); t < 0 -> 0 (D1, :310-313)
        const int s2 = s_hit + 4, t = c + 4;
        int ext = 0;
        if (t >= 0) {
          int s1 = s2 + kMaxMatchLength - 4;
          if (s1 > n) s1 = n;
          ext = match_tail(srcb, s2, t, s1 - s2, 0, lane, n);
        }
        if (lane == 0) // match_token(l + 4 - 3, s - t - 1) (:228-233)
          __stcs(&tok[ntok], kMatchType + ((uint32_t)(ext + 1) << kLengthShift) + (uint32_t)(s2 - t - 1));
        ntok++;
        s = s2 + ext;
        next_emit = s;
        if (s >= s_limit) break; // :236-238
        modeM = true;
      }
      // emit_remainder (:152-159)
      for (int i = next_emit + lane; i < n; i += 32) __stcs(&tok[ntok + (uint32_t)(i - next_emit)], (uint32_t)__ldg(srcb + i));
      ntok += (uint32_t)(n - next_emit);
      if (lane == 0) j.blk_ntok[blk0 + b] = ntok;
      __syncwarp();
      if (MULTI && bp && ((b + 1 < nblk && L - (uint64_t)(b + 1) * kBlockSize >= 128) || (bp->cont_open && b + 1 == nblk))) {
        // the next block is parsed too: leave it this block's end table, and note whether it differs from the
        // one the previous round left (only then the next block has to be parsed again)
        const uint64_t m = bp->mb_idx[bp_gb];
        const uint32_t old_buf = bp->lat_prev[m], new_buf = old_buf ^ 1u;
        const uint16_t *oldt = bp->tabs + ((size_t)old_buf * bp->nmb + m) * kTableSize;
        uint16_t *newt = bp->tabs + ((size_t)new_buf * bp->nmb + m) * kTableSize;
        const uint32_t end_rel = S0 + (uint32_t)n; // block end, relative like the table positions
        bool diff = false;
        for (int i = lane; i < kTableSize; i += 32) {
          const uint32_t e = (uint32_t)table[i];
          uint32_t d = e ? end_rel - (e - 1u) : 0u;
          if (d > (uint32_t)kMaxMatchOffset) d = 0;
          diff |= (bp->round == 1) || (oldt[i] != (uint16_t)d);
          newt[i] = (uint16_t)d;
        }
        diff = __any_sync(kFull, diff);
        if (lane == 0) {
          bp->lat_next[m] = (uint8_t)new_buf;
          bp->chg_next[m] = diff ? 1 : 0;
        }
        __syncwarp();
      }
    }
  }
}

// Warps [0, smem_warps) of a CTA keep their table in shared memory; the others
// (the kernel is latency bound and shared memory caps it at 7 tables per SM)
// keep theirs in a global scratch area that stays L2 resident.
template <bool MULTI>
__global__ void k_parse(DeflateJob j, uint32_t *counter, int smem_warps, void *gtables)
{
  using T = typename std::conditional<MULTI, uint32_t, uint16_t>::type;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  if (warp < smem_warps) {
    parse_worker<MULTI, T, false>(j, counter, reinterpret_cast<T *>(smem_raw) + (size_t)warp * kTableSize);
  } else {
    const int gw = (int)(blockDim.x >> 5) - smem_warps;
    T *table = reinterpret_cast<T *>(gtables) + ((size_t)blockIdx.x * gw + (warp - smem_warps)) * kTableSize;
    parse_worker<MULTI, T, true>(j, counter, table);
  }
}

// One round of the block-parallel parse of multi-block streams: the blocks on the list are parsed, each from the
// end table its predecessor has at the moment.
__global__ void k_parse_blocks(DeflateJob j, BlockParJob bp, uint32_t *counter, int smem_warps, void *gtables)
{
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  if (warp < smem_warps) {
    parse_worker<true, uint32_t, false>(j, counter, reinterpret_cast<uint32_t *>(smem_raw) + (size_t)warp * kTableSize, &bp);
  } else {
    const int gw = (int)(blockDim.x >> 5) - smem_warps;
    uint32_t *table = reinterpret_cast<uint32_t *>(gtables) + ((size_t)blockIdx.x * gw + (warp - smem_warps)) * kTableSize;
    parse_worker<true, uint32_t, true>(j, counter, table, &bp);
  }
}

__device__ __forceinline__ bool blk_of_multi_stream(const DeflateJob &j, uint64_t gb, uint32_t *b_out, bool *parsed)
{
  const uint32_t st = j.blk_stream[gb];
  const uint64_t L = j.stream_off[st + 1] - j.stream_off[st];
  const uint64_t b = gb - j.stream_blk0[st];
  *b_out = (uint32_t)b;
  *parsed = L - b * (uint64_t)kBlockSize >= 128;
  return L >= (uint64_t)kBlockSize + 128;
}

__global__ void k_bp_flags(DeflateJob j, uint64_t *flags)
{
  const uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= j.nblocks) return;
  uint32_t b;
  bool parsed;
  flags[gb] = blk_of_multi_stream(j, gb, &b, &parsed) ? 1 : 0;
}

// Work list of a round: every parsed block in round 1, afterwards the blocks whose predecessor's end table
// changed in the previous round.  Also carries the per-block state over to this round's arrays.
__global__ void k_bp_round(DeflateJob j, BlockParJob bp, const uint8_t *chg_prev, uint32_t *list, uint32_t *nlist)
{
  const uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= j.nblocks) return;
  uint32_t b;
  bool parsed;
  if (!blk_of_multi_stream(j, gb, &b, &parsed)) return;
  const uint64_t m = bp.mb_idx[gb];
  bp.lat_next[m] = bp.round == 1 ? 0 : bp.lat_prev[m];
  bp.chg_next[m] = 0;
  const bool need = parsed && !(bp.cont_prev && b == 0) &&
                    (bp.round == 1 || (b > 0 && !block_resets_table(bp.cont_prev ? bp.cont_block_base + b - 1 : b) && chg_prev[m - 1]));
  if (need) list[atomicAdd(nlist, 1u)] = (uint32_t)gb;
}

__global__ void k_bp_save_table(BlockParJob bp, uint64_t m, const uint8_t *lat, uint16_t *out)
{
  const uint16_t *t = bp.tabs + ((size_t)lat[m] * bp.nmb + m) * kTableSize;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kTableSize; i += gridDim.x * blockDim.x) out[i] = t[i];
}

void launch_bp_save_table(const BlockParJob &bp, uint64_t m, const uint8_t *lat, uint16_t *out, cudaStream_t st)
{
  k_bp_save_table<<<16, 256, 0, st>>>(bp, m, lat, out);
}

void launch_init_tables(cudaStream_t st) { k_init_sched<<<1, 1, 0, st>>>(); }

// Launch configuration: warps per CTA (one CTA per SM) with their table in shared memory / in global memory.
// FB200_PARSE_WARPS / FB200_PARSE_GWARPS override the defaults; whatever is asked for is cut down to what the
// registers of the kernel allow (cudaFuncAttributes::maxThreadsPerBlock).
static int g_parse_occ_single = 0, g_parse_occ_multi = 0, g_parse_gwarps = 0; // as configured
constexpr int kMaxDevices = 64;
struct ParseCfg { int sw = 0, gw = 0; };
static ParseCfg g_cfg_single[kMaxDevices], g_cfg_multi[kMaxDevices], g_cfg_blocks[kMaxDevices];

static int current_device()
{
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

template <typename K> static ParseCfg fit_cfg(K kernel, int sw, int gw)
{
  cudaFuncAttributes a{};
  int maxw = 32;
  if (cudaFuncGetAttributes(&a, kernel) == cudaSuccess && a.maxThreadsPerBlock >= 32) maxw = a.maxThreadsPerBlock / 32;
  else cudaGetLastError();
  ParseCfg c;
  c.sw = sw < maxw ? sw : maxw;
  c.gw = gw < maxw - c.sw ? gw : maxw - c.sw;
  if (c.sw + c.gw == 0) c.sw = 1;
  return c;
}

static void parse_init(int num_sms)
{
  (void)num_sms;
  static bool configured = false;
  if (!configured) { // warp mix: once per process
    const char *e = getenv("FB200_PARSE_WARPS");
    int w = e ? atoi(e) : FB_PARSE_SW;
    if (w < 0) w = 0;
    if (w > 7) w = 7;
    const char *g = getenv("FB200_PARSE_GWARPS");
    int gw = g ? atoi(g) : FB_PARSE_GW;
    if (gw < 0) gw = 0;
    if (gw > 32) gw = 32;
    if (w + gw == 0) w = 1;
    if (w + gw > 32) gw = 32 - w; // one CTA per SM, at most 1024 threads
    g_parse_occ_single = w;
    g_parse_occ_multi = w > 3 ? 3 : w;
    g_parse_gwarps = gw;
    configured = true;
  }
  static bool attr_set[kMaxDevices] = {}; // function attributes are per device
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaFuncSetAttribute(k_parse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_parse_occ_single * kTableSize * 2);
    cudaFuncSetAttribute(k_parse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_parse_occ_multi * kTableSize * 4);
    cudaFuncSetAttribute(k_parse_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, g_parse_occ_multi * kTableSize * 4);
    const char *co = getenv("FB200_PARSE_CARVEOUT");
    const int pct = co ? atoi(co) : -1;
    // Shared-memory carve-out: by default the smallest one that holds the tables (5 tables -> 164 KB, 92 KB of
    // L1).  Measured with 5 + 25 warps: 196 KB carve-out 19.8 ms per GiB, 228 KB 30.7 ms, default 18.6 ms.
    if (pct >= 0) {
      cudaFuncSetAttribute(k_parse<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      cudaFuncSetAttribute(k_parse<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    g_cfg_single[dev] = fit_cfg(k_parse<false>, g_parse_occ_single, g_parse_gwarps);
    g_cfg_multi[dev] = fit_cfg(k_parse<true>, g_parse_occ_multi, g_parse_gwarps);
    g_cfg_blocks[dev] = fit_cfg(k_parse_blocks, g_parse_occ_multi, g_parse_gwarps);
    attr_set[dev] = true;
  }
}

// bytes of global-table scratch launch_parse_* needs (owned by the context: one per GPU)
size_t parse_gtables_bytes(int num_sms)
{
  parse_init(num_sms);
  return (size_t)num_sms * (size_t)(g_parse_gwarps ? g_parse_gwarps : 1) * kTableSize * 4;
}

void launch_parse_single(const DeflateJob &j, int num_sms, void *gtables, cudaStream_t st)
{
  parse_init(num_sms);
  const ParseCfg c = g_cfg_single[current_device()];
  k_parse<false><<<num_sms, (c.sw + c.gw) * 32, c.sw * kTableSize * 2, st>>>(j, j.counters + 0, c.sw, gtables);
}

void launch_parse_multi(const DeflateJob &j, int num_sms, void *gtables, cudaStream_t st)
{
  parse_init(num_sms);
  const ParseCfg c = g_cfg_multi[current_device()];
  k_parse<true><<<num_sms, (c.sw + c.gw) * 32, c.sw * kTableSize * 4, st>>>(j, j.counters + 1, c.sw, gtables);
}

void launch_parse(const DeflateJob &j, int num_sms, void *gtables, cudaStream_t st)
{
  launch_parse_single(j, num_sms, gtables, st);
  launch_parse_multi(j, num_sms, gtables, st);
}

void launch_bp_flags(const DeflateJob &j, uint64_t *flags, cudaStream_t st)
{
  if (j.nblocks == 0) return;
  k_bp_flags<<<(unsigned)((j.nblocks + 255) / 256), 256, 0, st>>>(j, flags);
}

void launch_bp_round(const DeflateJob &j, const BlockParJob &bp, const uint8_t *chg_prev, uint32_t *list, uint32_t *nlist,
                     cudaStream_t st)
{
  if (j.nblocks == 0) return;
  k_bp_round<<<(unsigned)((j.nblocks + 255) / 256), 256, 0, st>>>(j, bp, chg_prev, list, nlist);
}

void launch_parse_blocks(const DeflateJob &j, const BlockParJob &bp, uint32_t *counter, int num_sms, void *gtables,
                         cudaStream_t st)
{
  parse_init(num_sms);
  ParseCfg c = g_cfg_blocks[current_device()];
  // few blocks on the list (one long stream): warps with shared-memory tables only -- a batch is one trip to L2
  // shorter, and a round of the fixpoint lasts as long as the parse of ONE block (1 MiB stream: 14.3 -> 10.3 ms)
  if (c.sw > 0 && (uint64_t)bp.nlist <= (uint64_t)num_sms * (uint64_t)c.sw) c.gw = 0;
  k_parse_blocks<<<num_sms, (c.sw + c.gw) * 32, c.sw * kTableSize * 4, st>>>(j, bp, counter, c.sw, gtables);
}

// ------------------------------------------------------------------
// setup kernels

__global__ void k_count_blocks(DeflateJob j, uint64_t *nblk_out)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nstreams) return;
  uint64_t L = j.stream_off[i + 1] - j.stream_off[i];
  nblk_out[i] = (L + kBlockSize - 1) / kBlockSize; // Compressor::write cuts at 65535 (deflate.mbt:222-229,:238)
}

// streams of more than one parsed block -> counters[12], their blocks -> counters[13] (saturating)
__global__ void k_count_multi(DeflateJob j)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nstreams) return;
  const uint64_t L = j.stream_off[i + 1] - j.stream_off[i];
  if (L >= (uint64_t)kBlockSize + 128) {
    atomicAdd(&j.counters[12], 1u);
    const uint64_t nb = (L + kBlockSize - 1) / kBlockSize;
    const uint32_t old = atomicAdd(&j.counters[13], (uint32_t)(nb > 0x3fffffffu ? 0x3fffffffu : nb));
    if (old > 0x7fffffffu) atomicExch(&j.counters[13], 0x7fffffffu);
  }
}

void launch_count_multi(const DeflateJob &j, cudaStream_t st)
{
  if (j.nstreams == 0) return;
  k_count_multi<<<(unsigned)((j.nstreams + 255) / 256), 256, 0, st>>>(j);
}

void launch_count_blocks(const DeflateJob &j, cudaStream_t st)
{
  if (j.nstreams == 0) return;
  unsigned g = (unsigned)((j.nstreams + 255) / 256);
  k_count_blocks<<<g, 256, 0, st>>>(j, j.stream_blk0);
}

__global__ void k_fill_blocks(DeflateJob j)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.nstreams) return;
  uint64_t b0 = j.stream_blk0[i], b1 = j.stream_blk0[i + 1];
  for (uint64_t b = b0; b < b1; b++) j.blk_stream[b] = (uint32_t)i;
}

void launch_fill_blocks(const DeflateJob &j, cudaStream_t st)
{
  if (j.nstreams == 0) return;
  unsigned g = (unsigned)((j.nstreams + 255) / 256);
  k_fill_blocks<<<g, 256, 0, st>>>(j);
}

__global__ void k_fill_seg_off(uint64_t *off, uint64_t nseg, uint64_t seg, uint64_t n)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nseg) return;
  uint64_t v = i * seg;
  off[i] = v < n ? v : n;
}

void launch_fill_seg_off(uint64_t *off, uint64_t nseg, uint64_t seg, uint64_t n, cudaStream_t st)
{
  unsigned g = (unsigned)((nseg + 1 + 255) / 256);
  k_fill_seg_off<<<g, 256, 0, st>>>(off, nseg, seg, n);
}

__global__ void k_affine_u64(uint64_t *out, const uint64_t *in, uint64_t cnt, uint64_t delta)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) out[i] = in[i] + delta;
}

void launch_affine_u64(uint64_t *out, const uint64_t *in, uint64_t cnt, uint64_t delta, cudaStream_t st)
{
  if (cnt == 0) return;
  k_affine_u64<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(out, in, cnt, delta);
}

// Single-CTA exclusive scan (metadata only: <= a few million entries).
__global__ void __launch_bounds__(1024) k_scan_u64(const uint64_t *in, uint64_t *out, uint64_t n, const uint64_t *carry_from,
                                                   uint64_t *host_total)
{
  __shared__ uint64_t wsum[32];
  __shared__ uint64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = carry_from ? *carry_from : 0;
  __syncthreads();
  for (uint64_t base = 0; base < n; base += 1024) {
    uint64_t i = base + tid;
    uint64_t v = i < n ? in[i] : 0;
    uint64_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint64_t w = wsum[lane];
      uint64_t xs = w;
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(kFull, xs, o);
        if (lane >= o) xs += y;
      }
      wsum[lane] = xs - w; // exclusive
    }
    __syncthreads();
    const uint64_t carry = carry_s;
    if (i < n) out[i] = carry + wsum[warp] + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31] + x;
    __syncthreads();
  }
  if (tid == 0) {
    out[n] = carry_s;
    if (host_total) { // pinned host memory: the host learns the total without queueing behind a bulk D2H copy
      *host_total = carry_s;
      __threadfence_system();
    }
  }
}

void launch_scan_u64(const uint64_t *in, uint64_t *out, uint64_t n, cudaStream_t st, const uint64_t *carry_from,
                     uint64_t *host_total)
{
  k_scan_u64<<<1, 1024, 0, st>>>(in, out, n, carry_from, host_total);
}

__global__ void k_gather_u64(uint64_t *out, const uint64_t *in, uint64_t stride, uint64_t n, uint64_t cnt)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const uint64_t k = i * stride;
  out[i] = in[k < n ? k : n];
}

void launch_gather_u64(uint64_t *out, const uint64_t *in, uint64_t stride, uint64_t n, uint64_t cnt, cudaStream_t st)
{
  if (cnt == 0) return;
  k_gather_u64<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(out, in, stride, n, cnt);
}

// Frame reader (single CTA: the header holds 4 bytes per segment): prefix of the sizes in front of `first`, then
// an exclusive scan of the sizes of the range.
__global__ void __launch_bounds__(1024) k_frame_range(const uint32_t *sizes, uint64_t first, uint64_t count,
                                                      uint64_t *comp_off, uint64_t *res)
{
  __shared__ uint64_t wsum[32];
  __shared__ uint64_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t acc = 0;
  for (uint64_t i = tid; i < first; i += 1024) acc += sizes[i];
  for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(kFull, acc, o);
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    uint64_t t = 0;
    for (int w = 0; w < 32; w++) t += wsum[w];
    res[0] = t;
    carry_s = 0;
  }
  __syncthreads();
  for (uint64_t base = 0; base < count; base += 1024) {
    const uint64_t i = base + tid;
    const uint64_t v = i < count ? sizes[first + i] : 0;
    uint64_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const uint64_t w = wsum[lane];
      uint64_t xs = w;
      for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(kFull, xs, o);
        if (lane >= o) xs += y;
      }
      wsum[lane] = xs - w;
    }
    __syncthreads();
    const uint64_t carry = carry_s;
    if (i < count) comp_off[i] = carry + wsum[warp] + x - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31] + x;
    __syncthreads();
  }
  if (tid == 0) {
    comp_off[count] = carry_s;
    res[1] = carry_s;
  }
}

void launch_frame_range(const uint32_t *sizes, uint64_t first, uint64_t count, uint64_t *comp_off, uint64_t *res, cudaStream_t st)
{
  k_frame_range<<<1, 1024, 0, st>>>(sizes, first, count, comp_off, res);
}

// CUDA loads kernels lazily, and loading one while another kernel spins on a host-fed watermark can
// deadlock: every kernel of this file is loaded when the context is created.
void preload_parse_kernels()
{
  parse_init(0); // process-wide launch configuration + per-device function attributes: before any concurrent use
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_init_sched);
  cudaFuncGetAttributes(&a, k_parse<false>);
  cudaFuncGetAttributes(&a, k_parse<true>);
  cudaFuncGetAttributes(&a, k_parse_blocks);
  cudaFuncGetAttributes(&a, k_bp_flags);
  cudaFuncGetAttributes(&a, k_bp_round);
  cudaFuncGetAttributes(&a, k_bp_save_table);
  cudaFuncGetAttributes(&a, k_count_blocks);
  cudaFuncGetAttributes(&a, k_count_multi);
  cudaFuncGetAttributes(&a, k_fill_blocks);
  cudaFuncGetAttributes(&a, k_fill_seg_off);
  cudaFuncGetAttributes(&a, k_affine_u64);
  cudaFuncGetAttributes(&a, k_scan_u64);
  cudaFuncGetAttributes(&a, k_gather_u64);
  cudaFuncGetAttributes(&a, k_frame_range);
}

} // namespace fb
