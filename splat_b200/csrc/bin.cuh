// bin.cuh -- K2 (tile counts, duplication into (tile, Gaussian) instances) and K4 (tile ranges).
//
// Replaces the instance expansion of render_to_buffer (pipelines.rs:69-79, :263-273: six
// VertexInstances per Gaussian in sorted order) and euc's implicit "every primitive visits
// every covered pixel in submission order" loop: after these kernels each 16x16 tile owns the
// far -> near list of the Gaussians whose 3-sigma quad can touch it.
#pragma once
#include "common.cuh"
#include "project.cuh"

namespace splat {

SPLAT_DEVINL TileRect unpack_rect(uint2 r) {
  TileRect t;
  t.x0 = (uint16_t)(r.x & 0xFFFFu); t.y0 = (uint16_t)(r.x >> 16);
  t.x1 = (uint16_t)(r.y & 0xFFFFu); t.y1 = (uint16_t)(r.y >> 16);
  return t;
}

SPLAT_DEVINL TileRect clip_rect(TileRect t, const TileRect box) {
  if (t.x1 < t.x0 || t.y1 < t.y0) return t;     // empty stays empty
  t.x0 = (uint16_t)max((uint32_t)t.x0, (uint32_t)box.x0); t.y0 = (uint16_t)max((uint32_t)t.y0, (uint32_t)box.y0);
  t.x1 = (uint16_t)min((uint32_t)t.x1, (uint32_t)box.x1); t.y1 = (uint16_t)min((uint32_t)t.y1, (uint32_t)box.y1);
  return t;                                     // may have become empty (x1 < x0 or y1 < y0): count() == 0
}

// cnt[r] = number of tiles of the Gaussian at depth rank r (0 for culled ones, whose key
// 0xFFFFFFFF sorted them to the end).  Also counts the visible Gaussians.
__global__ void __launch_bounds__(256)
tile_count_kernel(const uint32_t *__restrict__ sorted_keys, const uint32_t *__restrict__ order,
                  const uint32_t *__restrict__ tcnt, uint32_t *__restrict__ cnt, uint32_t n,
                  const uint32_t *__restrict__ n_sorted, uint32_t rank_cut, FrameStatus *__restrict__ status,
                  const uint2 *__restrict__ rects, TileRect box) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  // stripe renders sort only *n_sorted pairs; the tail of the buffers is stale
  const uint32_t ns = n_sorted ? *n_sorted : n;
  bool vis = false;
  if (r < n) {
    vis = r < ns && sorted_keys[r] != KEY_CULLED;
    // near cut: ranks below rank_cut (the farthest Gaussians) get no instances this pass
    uint32_t c = 0;
    if (vis && r >= rank_cut) {
      if (rects) c = clip_rect(unpack_rect(__ldg(&rects[order[r]])), box).count();   // pass restricted to a box of tiles
      else c = __ldg(&tcnt[order[r]]);
    }
    cnt[r] = c;
  }
  // one atomic per CTA: 190k same-address atomics (one per warp) serialised in L2 and were the
  // whole cost of this kernel (r1h: 143 us at 7% issue, 14% of the DRAM peak)
  const int nv = __syncthreads_count(vis);
  if (threadIdx.x == 0 && nv) atomicAdd(&status->n_visible, (unsigned int)nv);
}

// ---------------------------------------------------------------- near cut (far coverage)
// The exact early termination of the blend kernel reads only the nearest few hundred entries of
// a tile list, so a frame first bins and sorts only the nearest Gaussians (depth ranks >=
// rank_cut).  To stay exact the blend must know, per tile, whether anything was cut away:
// far_cnt[t] = number of cut Gaussians whose quad touches tile t.  Each quad adds +1 on a
// rectangle of tiles; instead of one atomic per (Gaussian, tile) pair -- the very enumeration the
// cut avoids -- the four corners of a 2-D difference array are bumped in shared memory (a
// persistent grid, each CTA folding a slice of the ranks) and far_prefix_kernel integrates it.
constexpr int FC_THREADS = 1024;
constexpr int FC_UNROLL = 4;     // independent order -> rect gathers in flight per thread
__global__ void __launch_bounds__(FC_THREADS)
far_cover_kernel(const uint32_t *__restrict__ order, const uint2 *__restrict__ rects,
                 const uint32_t *__restrict__ n_sorted, uint32_t n, uint32_t rank_cut,
                 uint32_t tiles_x, uint32_t tiles_y, int *__restrict__ diff /* (tiles_y+1) x (tiles_x+1), zeroed */) {
  extern __shared__ int s_diff[];
  const uint32_t pitch = tiles_x + 1u, cells = pitch * (tiles_y + 1u);
  for (uint32_t i = threadIdx.x; i < cells; i += FC_THREADS) s_diff[i] = 0;
  __syncthreads();
  const uint32_t ns = n_sorted ? *n_sorted : n;
  const uint32_t m = min(rank_cut, ns);
  const uint32_t per = (m + gridDim.x - 1u) / gridDim.x;
  const uint32_t r0 = blockIdx.x * per, r1 = min(m, r0 + per);
  for (uint32_t rb = r0 + threadIdx.x; rb < r1; rb += FC_THREADS * FC_UNROLL) {
    uint32_t gi[FC_UNROLL];
    uint2 rc[FC_UNROLL];
#pragma unroll
    for (int u = 0; u < FC_UNROLL; ++u) {
      const uint32_t r = rb + u * FC_THREADS;
      gi[u] = (r < r1) ? __ldg(&order[r]) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int u = 0; u < FC_UNROLL; ++u) rc[u] = (gi[u] != 0xFFFFFFFFu) ? __ldg(&rects[gi[u]]) : make_uint2(1u, 0u);
#pragma unroll
    for (int u = 0; u < FC_UNROLL; ++u) {
      const uint32_t x0 = rc[u].x & 0xFFFFu, y0 = rc[u].x >> 16, x1 = rc[u].y & 0xFFFFu, y1 = rc[u].y >> 16;
      if (x1 >= x0 && y1 >= y0) {
        atomicAdd(&s_diff[y0 * pitch + x0], 1);
        atomicAdd(&s_diff[y0 * pitch + x1 + 1u], -1);
        atomicAdd(&s_diff[(y1 + 1u) * pitch + x0], -1);
        atomicAdd(&s_diff[(y1 + 1u) * pitch + x1 + 1u], 1);
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < cells; i += FC_THREADS) {
    const int v = s_diff[i];
    if (v) atomicAdd(&diff[i], v);
  }
}

// Single CTA: 2-D inclusive prefix sum of the difference array (in shared memory) ->
// far_cnt[ty * tiles_x + tx].
__global__ void __launch_bounds__(1024)
far_prefix_kernel(const int *__restrict__ diff, uint32_t tiles_x, uint32_t tiles_y, uint32_t *__restrict__ far_cnt,
                  FrameStatus *__restrict__ status) {
  extern __shared__ int s_diff[];
  const uint32_t pitch = tiles_x + 1u, cells = pitch * (tiles_y + 1u);
  for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) s_diff[i] = diff[i];
  __syncthreads();
  for (uint32_t y = threadIdx.x; y < tiles_y; y += blockDim.x) {     // along x, one row per thread
    int run = 0;
    for (uint32_t x = 0; x < tiles_x; ++x) { run += s_diff[y * pitch + x]; s_diff[y * pitch + x] = run; }
  }
  __syncthreads();
  for (uint32_t x = threadIdx.x; x < tiles_x; x += blockDim.x) {     // along y, one column per thread
    int run = 0;
    for (uint32_t y = 0; y < tiles_y; ++y) { run += s_diff[y * pitch + x]; s_diff[y * pitch + x] = run; }
  }
  __syncthreads();
  unsigned long long cut = 0;
  for (uint32_t i = threadIdx.x; i < tiles_x * tiles_y; i += blockDim.x) {
    const uint32_t v = (uint32_t)s_diff[(i / tiles_x) * pitch + (i % tiles_x)];
    far_cnt[i] = v;
    cut += v;
  }
  for (int o = 16; o > 0; o >>= 1) cut += __shfl_xor_sync(0xFFFFFFFFu, cut, o);
  if ((threadIdx.x & 31u) == 0 && cut) atomicAdd(&status->n_cut, cut);
}

// Duplication, load-balanced over OUTPUT positions: a CTA owns 256 consecutive depth ranks, whose
// instances form one contiguous output range [offs[r0], offs[r0+256)).  The CTA walks that range
// 256 positions at a time; thread j finds its rank with an 8-step binary search over the 256
// offsets in shared memory, so a quad that covers thousands of tiles is spread over the whole
// CTA, consecutive threads write consecutive positions (fully coalesced 4-byte stores), and no
// lane idles on small quads.  key = stripe-local tile id, value = Gaussian index; emission
// order = depth rank, which the stable tile sort preserves.
__global__ void __launch_bounds__(256)
emit_instances_kernel(const uint32_t *__restrict__ order, const uint2 *__restrict__ rects,
                      const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ offs,
                      uint32_t *__restrict__ inst_keys, uint32_t *__restrict__ inst_vals, uint32_t n,
                      uint32_t tiles_x, TileRect box, const uint32_t *__restrict__ n_eff) {
  __shared__ uint32_t s_off[257];
  if (*n_eff == 0u) return;     // nothing to emit, or the pairs do not fit the buffers (frame skipped)
  __shared__ uint32_t s_idx[256];
  __shared__ uint2 s_rect[256];
  const uint32_t tid = threadIdx.x, r0 = blockIdx.x * 256u, r = r0 + tid;
  const uint32_t last = min(n, r0 + 256u) - 1u;     // last valid rank of this CTA (r0 < n always)
  uint32_t c = 0, o = 0;
  if (r < n) {
    c = cnt[r];
    o = offs[r];
    s_off[tid] = o;
    if (c) {
      const uint32_t gi = order[r];
      s_idx[tid] = gi;
      const TileRect tr = clip_rect(unpack_rect(rects[gi]), box);
      s_rect[tid] = make_uint2((uint32_t)tr.x0 | ((uint32_t)tr.y0 << 16), (uint32_t)tr.x1 | ((uint32_t)tr.y1 << 16));
    }
    if (r == last) s_off[256] = o + c;
  }
  __syncthreads();
  const uint32_t begin = s_off[0], end = s_off[256];
  if (r >= n) s_off[tid] = end;     // only in the last CTA; read by the search below
  __syncthreads();
  for (uint32_t j = begin + tid; j < end; j += 256u) {
    // largest k in [0,255] with s_off[k] <= j  (zero-count ranks share their successor's offset,
    // so this lands on the rank that really owns position j)
    uint32_t k = 0;
#pragma unroll
    for (uint32_t step = 128u; step > 0u; step >>= 1)
      if (s_off[k + step] <= j) k += step;
    const uint32_t q = j - s_off[k];
    const uint2 rc = s_rect[k];
    const uint32_t x0 = rc.x & 0xFFFFu, y0 = rc.x >> 16, wdt = (rc.y & 0xFFFFu) - x0 + 1u;
    // q / wdt for q < 2^24, wdt < 2^16 through a float reciprocal, off by at most one
    uint32_t row = (uint32_t)(__uint2float_rz(q) * __frcp_rz(__uint2float_rz(wdt)));
    if ((row + 1u) * wdt <= q) row += 1u;
    const uint32_t col = q - row * wdt;
    inst_keys[j] = (y0 + row) * tiles_x + x0 + col;
    inst_vals[j] = s_idx[k];
  }
}

// ranges[t] = [start, end) of tile t in the tile-sorted instance list (zeroed beforehand).
// Four keys per thread (one 16-byte load) plus the two neighbours across the group boundary;
// grid-stride over a device-side count (see sort.cuh).
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t *__restrict__ sorted_tile_keys, const uint32_t *__restrict__ n_ptr,
                   uint2 *__restrict__ ranges) {
  const uint32_t n = *n_ptr;
  for (unsigned long long j64 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 4ull; j64 < n;
       j64 += (unsigned long long)gridDim.x * blockDim.x * 4ull) {
    const uint32_t j = (uint32_t)j64;
    uint32_t k[6];   // k[0] = key[j-1], k[1..4] = key[j..j+3], k[5] = key[j+4]
    if (j + 4u <= n) {
      const uint4 v = *reinterpret_cast<const uint4 *>(sorted_tile_keys + j);
      k[1] = v.x; k[2] = v.y; k[3] = v.z; k[4] = v.w;
    } else {
#pragma unroll
      for (uint32_t q = 0; q < 4u; ++q) k[1 + q] = (j + q < n) ? sorted_tile_keys[j + q] : 0xFFFFFFFFu;
    }
    k[0] = j ? sorted_tile_keys[j - 1u] : 0xFFFFFFFFu;
    k[5] = (j + 4u < n) ? sorted_tile_keys[j + 4u] : 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t q = 0; q < 4u; ++q) {
      if (j + q >= n) break;
      const uint32_t t = k[1 + q];
      if (j + q == 0u || k[q] != t) ranges[t].x = j + q;
      if (j + q + 1u == n || k[2 + q] != t) ranges[t].y = j + q + 1u;
    }
  }
}

}  // namespace splat
