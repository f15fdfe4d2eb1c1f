// sort.cuh -- K3: stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass, plus
// the exclusive scan it and the binning stage share.
//
// Replaces the reference's `indices.sort_by(|a,b| z[a].partial_cmp(&z[b]))`
// (gaussians.rs:303, :314, :469 -- a stable merge sort on view-space z).  The frame's
// (tile | depth) ordering is produced as an LSD radix sort whose four depth-digit passes run
// on the N Gaussians *before* they are duplicated per tile (duplicates share the depth key,
// so sorting them after duplication would move I >= N items through the same four passes),
// and whose tile-digit passes run on the I tile instances.  Stability of every pass keeps
// equal depths in ascending Gaussian index, exactly like the reference's stable sort.
//
// Per pass: histogram per 4096-key block -> per-digit exclusive scan over the blocks ->
// scatter (3 launches).  The scatter ranks keys with warp match-any (stable), reorders the block in shared
// memory so that each digit's run is written with coalesced stores.
// Algorithmic bytes per pass: 4 (hist read) + 8 (read) + 8 (write) = 20 B per pair.
#pragma once
#include "common.cuh"

namespace splat {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_BLOCK = RS_THREADS * RS_ITEMS;   // 4096 pairs per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WARP_SPAN = RS_ITEMS * 32;       // 512 consecutive pairs per warp
#ifndef SPLAT_RS_MIN_BLOCKS
#define SPLAT_RS_MIN_BLOCKS 4
#endif
constexpr int RS_MIN_BLOCKS = SPLAT_RS_MIN_BLOCKS; // resident scatter CTAs per SM the register budget must allow

SPLAT_DEVINL uint32_t rs_count(const uint32_t *n_ptr, uint32_t n_fixed) {
  return n_ptr ? *n_ptr : n_fixed;
}

// All three kernels of a pass take the pair count from device memory (n_ptr, or n_fixed if null),
// so the LAUNCH never depends on a count the host has not seen yet: the frame is enqueued without
// a host round trip, with a grid sized from the previous frame (splat_api.cu).  One 4096-pair
// block per CTA; CTAs beyond the real count exit at once, and a count beyond the launched grid is
// caught on the device before any of these kernels runs (scan_partials_kernel: the frame is
// skipped and repeated).  A grid-stride loop instead costs the scatter 17 registers (spills).
//
// hist[d * nblk + blk] = number of keys of block blk whose digit is d (nblk = ceil(n / 4096)).
// All 16 keys of a thread are loaded first (four 16-byte loads in flight), then counted with
// shared-memory atomics.  The digit of the last, partial pass is masked to its `nbits` bits.
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ n_ptr, uint32_t n_fixed,
               int shift, uint32_t dmask, uint32_t *__restrict__ hist) {
  __shared__ uint32_t h[256];
  const uint32_t n = rs_count(n_ptr, n_fixed);
  const uint32_t nblk = (n + RS_BLOCK - 1) / RS_BLOCK;
  const uint32_t blk = blockIdx.x;
  if (blk >= nblk) return;
  {
    const uint32_t base = blk * RS_BLOCK;
    h[threadIdx.x] = 0;
    __syncthreads();
    if (base + RS_BLOCK <= n) {
      uint4 v[RS_ITEMS / 4];
#pragma unroll
      for (int k = 0; k < RS_ITEMS / 4; ++k)
        v[k] = __ldg(reinterpret_cast<const uint4 *>(keys + base) + k * RS_THREADS + threadIdx.x);
#pragma unroll
      for (int k = 0; k < RS_ITEMS / 4; ++k) {
        atomicAdd(&h[(v[k].x >> shift) & dmask], 1u);
        atomicAdd(&h[(v[k].y >> shift) & dmask], 1u);
        atomicAdd(&h[(v[k].z >> shift) & dmask], 1u);
        atomicAdd(&h[(v[k].w >> shift) & dmask], 1u);
      }
    } else {
#pragma unroll
      for (int k = 0; k < RS_ITEMS; ++k) {
        const uint32_t idx = base + k * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & dmask], 1u);
      }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blk] = h[threadIdx.x];
  }
}

// One CTA per digit d: exclusive scan of hist[d][0..nblk) in place (where block blk's keys with
// digit d start inside the digit's run) and the digit total into tot[d].  The scatter kernel adds
// the digit's global base itself (an exclusive scan of the 256 totals), so a pass is three
// launches: histogram, this, scatter.
constexpr int RW_THREADS = 1024;
__global__ void __launch_bounds__(RW_THREADS)
rs_rowscan_kernel(uint32_t *__restrict__ hist, const uint32_t *__restrict__ n_ptr, uint32_t n_fixed,
                  uint32_t *__restrict__ tot) {
  __shared__ uint32_t wsum[RW_THREADS / 32];
  __shared__ uint32_t carry_s;
  const uint32_t n = rs_count(n_ptr, n_fixed);
  const uint32_t nblk = (n + RS_BLOCK - 1) / RS_BLOCK;
  uint32_t *row = hist + (size_t)blockIdx.x * nblk;
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblk; base += RW_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = (i < nblk) ? row[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t wb = 0;
    for (uint32_t q = 0; q < w; ++q) wb += wsum[q];
    const uint32_t carry = carry_s;
    if (i < nblk) row[i] = carry + wb + incl - v;
    __syncthreads();
    if (threadIdx.x == RW_THREADS - 1) carry_s = carry + wb + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) tot[blockIdx.x] = carry_s;
}

// NB = 8: all eight digit bits (fully unrolled ranking); NB = 0: `nbits` < 8 bits at run time
// (the last pass of a key whose width is not a multiple of 8; the digit is masked to them).
template <int NB>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS)
rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                  const uint32_t *__restrict__ n_ptr, uint32_t n_fixed, int shift, int nbits,
                  const uint32_t *__restrict__ hist_scanned, const uint32_t *__restrict__ tot_g) {
  __shared__ uint32_t cnt[RS_WARPS][256];   // per-warp digit counts, then exclusive warp bases
  __shared__ uint32_t dbase[256];           // block-local start of each digit's run
  __shared__ uint32_t gofs[256];            // global offset of the run minus dbase
  __shared__ uint32_t skey[RS_BLOCK];
  __shared__ uint32_t sval[RS_BLOCK];
  __shared__ uint32_t wsum[RS_WARPS];

  const uint32_t n = rs_count(n_ptr, n_fixed);
  const uint32_t nblk = (n + RS_BLOCK - 1) / RS_BLOCK;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t dmask = NB ? 0xFFu : ((1u << nbits) - 1u);

  const uint32_t blk = blockIdx.x;
  if (blk >= nblk) return;
  const uint32_t base = blk * RS_BLOCK;
  const uint32_t nvalid = min((uint32_t)RS_BLOCK, n - base);

#pragma unroll
  for (int q = 0; q < RS_WARPS; ++q) cnt[q][tid] = 0;
  __syncthreads();

  // Stable ranking inside the warp.  The lanes that hold the same digit ("peers") are found with
  // one ballot per digit bit; match.any would do it in one instruction, but on sm_100 it keeps
  // the ADU pipe busy for ~64 cycles per warp instruction (ncu r1h: pipe_adu 76%, the limiter of
  // this kernel), a ballot for 2.  The digit's first lane bumps the warp's counter; rank = old
  // count + number of peers in lower lanes, so equal digits keep their input order.
  uint32_t key[RS_ITEMS];
  uint16_t rank[RS_ITEMS];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const uint32_t li = w * RS_WARP_SPAN + k * 32 + lane;   // index order == (warp, round, lane)
    key[k] = (li < nvalid) ? keys_in[base + li] : 0xFFFFFFFFu;
  }
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const uint32_t li = w * RS_WARP_SPAN + k * 32 + lane;
    const bool valid = li < nvalid;
    const uint32_t d = (key[k] >> shift) & dmask;
    // (mixing in a few match.any rounds to run on the ADU pipe beside the ballots was tried
    // and is slower: r1s)
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
    const int nb = NB ? NB : nbits;
#pragma unroll
    for (int b = 0; b < nb; ++b) {
      const bool bit = (d >> b) & 1u;
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
      peers &= bit ? bal : ~bal;
    }
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && (int)lane == leader) old = atomicAdd(&cnt[w][d], (uint32_t)__popc(peers));
    old = __shfl_sync(0xFFFFFFFFu, old, leader & 31);
    rank[k] = (uint16_t)(old + __popc(peers & lt_mask));
    __syncwarp();
  }
  __syncthreads();

  // digit tid: exclusive bases across warps + block total
  uint32_t tot = 0;
#pragma unroll
  for (int q = 0; q < RS_WARPS; ++q) {
    const uint32_t t = cnt[q][tid];
    cnt[q][tid] = tot;
    tot += t;
  }
  // exclusive scan of the 256 digit totals
  uint32_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int q = 0; q < RS_WARPS; ++q) wbase += (q < (int)w) ? wsum[q] : 0u;
  const uint32_t excl = wbase + incl - tot;
  dbase[tid] = excl;
  // global base of digit tid = exclusive scan of the 256 digit totals (rs_rowscan_kernel)
  const uint32_t gt = tot_g[tid];
  uint32_t gincl = gt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, gincl, o);
    if (lane >= (uint32_t)o) gincl += t;
  }
  __syncthreads();                 // wsum is re-used
  if (lane == 31) wsum[w] = gincl;
  __syncthreads();
  uint32_t gbase = 0;
#pragma unroll
  for (int q = 0; q < RS_WARPS; ++q) gbase += (q < (int)w) ? wsum[q] : 0u;
  gofs[tid] = (gbase + gincl - gt) + hist_scanned[(size_t)tid * nblk + blk] - excl;
  __syncthreads();

#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const uint32_t li = w * RS_WARP_SPAN + k * 32 + lane;
    if (li < nvalid) {
      const uint32_t d = (key[k] >> shift) & dmask;
      const uint32_t lp = dbase[d] + cnt[w][d] + rank[k];
      skey[lp] = key[k];
      sval[lp] = vals_in[base + li];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const uint32_t lp = k * RS_THREADS + tid;
    if (lp < nvalid) {
      const uint32_t kk = skey[lp];
      const uint32_t g = gofs[(kk >> shift) & dmask] + lp;
      keys_out[g] = kk;
      vals_out[g] = sval[lp];
    }
  }
}

// ---------------------------------------------------------------- exclusive scan (u32)
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_BLOCK = SC_THREADS * SC_ITEMS;   // 2048

SPLAT_DEVINL uint32_t block_reduce_sum(uint32_t v, uint32_t *sh /* 8 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
#pragma unroll
  for (int q = 0; q < SC_THREADS / 32; ++q) t += sh[q];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(SC_THREADS)
scan_reduce_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ partial, uint32_t n) {
  __shared__ uint32_t sh[SC_THREADS / 32];
  const uint32_t base = blockIdx.x * SC_BLOCK + threadIdx.x * SC_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) s += (base + k < n) ? in[base + k] : 0u;
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// single CTA: exclusive scan of the block partials (64-bit running sum for the grand total).
// With `st` set, the total is a count the next launches were sized for WITHOUT the host having
// seen it (a stripe's survivors, the frame's tile instances): it is checked against that bound ON
// THE DEVICE -- *eff_out = total if it fits, else 0 and the frame is flagged (st->overflow): every
// later kernel then has nothing to do, the target stays untouched, and the host repeats the frame.
__global__ void __launch_bounds__(1024)
scan_partials_kernel(uint32_t *__restrict__ partial, uint32_t np, unsigned long long *total_out,
                     FrameStatus *st = nullptr, unsigned long long cap = 0, unsigned int *eff_out = nullptr,
                     bool zero_total_on_overflow = false) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  for (uint32_t base = 0; base < np; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = (i < np) ? partial[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    unsigned long long wb = 0;
    for (uint32_t q = 0; q < w; ++q) wb += wsum[q];
    const unsigned long long carry = carry_s;
    if (i < np) partial[i] = (uint32_t)(carry + wb + incl - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wb + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    unsigned long long total = carry_s;
    if (st) {
      const bool already = st->overflow != 0u;      // an earlier stage of this frame did not fit
      const bool fits = total <= cap && !already;
      if (eff_out) *eff_out = fits ? (unsigned int)total : 0u;
      if (!fits && !already) { st->overflow = 1u; st->skipped += 1u; }
      if (!fits && zero_total_on_overflow) total = 0;
    }
    if (total_out) *total_out = total;
  }
}

__global__ void __launch_bounds__(SC_THREADS)
scan_apply_kernel(const uint32_t *in, uint32_t *out,   // in == out is allowed (in-place)
                  const uint32_t *__restrict__ partial, uint32_t n) {
  __shared__ uint32_t wsum[SC_THREADS / 32];
  const uint32_t base = blockIdx.x * SC_BLOCK + threadIdx.x * SC_ITEMS;
  uint32_t v[SC_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t wb = 0;
#pragma unroll
  for (int q = 0; q < SC_THREADS / 32; ++q) wb += (q < (int)w) ? wsum[q] : 0u;
  uint32_t run = partial[blockIdx.x] + wb + incl - s;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

}  // namespace splat
