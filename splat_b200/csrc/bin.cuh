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

// cnt[r] = number of tiles of the Gaussian at depth rank r (0 for culled ones, whose key
// 0xFFFFFFFF sorted them to the end).  Also counts the visible Gaussians.
__global__ void __launch_bounds__(256)
tile_count_kernel(const uint32_t *__restrict__ sorted_keys, const uint32_t *__restrict__ order,
                  const uint32_t *__restrict__ tcnt, uint32_t *__restrict__ cnt, uint32_t n,
                  const uint32_t *__restrict__ n_sorted, FrameStatus *__restrict__ status) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  // stripe renders sort only *n_sorted pairs; the tail of the buffers is stale
  const uint32_t ns = n_sorted ? *n_sorted : n;
  bool vis = false;
  if (r < n) {
    vis = r < ns && sorted_keys[r] != KEY_CULLED;
    cnt[r] = vis ? __ldg(&tcnt[order[r]]) : 0u;
  }
  // one atomic per CTA: 190k same-address atomics (one per warp) serialised in L2 and were the
  // whole cost of this kernel (r1h: 143 us at 7% issue, 14% of the DRAM peak)
  const int nv = __syncthreads_count(vis);
  if (threadIdx.x == 0 && nv) atomicAdd(&status->n_visible, (unsigned int)nv);
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
                      uint32_t tiles_x) {
  __shared__ uint32_t s_off[257];
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
      s_rect[tid] = rects[gi];
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
// Four keys per thread (one 16-byte load) plus the two neighbours across the group boundary.
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t *__restrict__ sorted_tile_keys, uint32_t n,
                   uint2 *__restrict__ ranges) {
  const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  if (j >= n) return;
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

}  // namespace splat
