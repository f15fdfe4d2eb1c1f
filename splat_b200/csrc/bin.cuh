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
                  const uint2 *__restrict__ rects, uint32_t *__restrict__ cnt, uint32_t n,
                  FrameStatus *__restrict__ status) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  bool vis = false;
  if (r < n) {
    vis = sorted_keys[r] != KEY_CULLED;
    cnt[r] = vis ? unpack_rect(rects[order[r]]).count() : 0u;
  }
  const uint32_t b = __ballot_sync(0xFFFFFFFFu, vis);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&status->n_visible, (unsigned int)__popc(b));
}

// Warp-cooperative duplication: the warp walks its 32 Gaussians one at a time and its lanes
// write that Gaussian's tiles, so big quads (thousands of tiles) are spread over 32 lanes and
// the stores of one Gaussian are contiguous.  key = stripe-local tile id, value = Gaussian
// index; emission order = depth rank, which the stable tile sort preserves.
__global__ void __launch_bounds__(256)
emit_instances_kernel(const uint32_t *__restrict__ order, const uint2 *__restrict__ rects,
                      const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ offs,
                      uint32_t *__restrict__ inst_keys, uint32_t *__restrict__ inst_vals, uint32_t n,
                      uint32_t tiles_x) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t my_cnt = 0, my_off = 0, my_idx = 0;
  uint2 my_rect = make_uint2(0, 0);
  if (r < n) {
    my_cnt = cnt[r];
    if (my_cnt) {
      my_off = offs[r];
      my_idx = order[r];
      my_rect = rects[my_idx];
    }
  }
  uint32_t todo = __ballot_sync(0xFFFFFFFFu, my_cnt != 0);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t c = __shfl_sync(0xFFFFFFFFu, my_cnt, src);
    const uint32_t o = __shfl_sync(0xFFFFFFFFu, my_off, src);
    const uint32_t g = __shfl_sync(0xFFFFFFFFu, my_idx, src);
    const uint32_t rx = __shfl_sync(0xFFFFFFFFu, my_rect.x, src);
    const uint32_t ry = __shfl_sync(0xFFFFFFFFu, my_rect.y, src);
    const uint32_t x0 = rx & 0xFFFFu, y0 = rx >> 16, x1 = ry & 0xFFFFu;
    const uint32_t wdt = x1 - x0 + 1;
    for (uint32_t k = lane; k < c; k += 32) {
      const uint32_t ty = y0 + k / wdt, tx = x0 + k % wdt;
      inst_keys[o + k] = ty * tiles_x + tx;
      inst_vals[o + k] = g;
    }
  }
}

// ranges[t] = [start, end) of tile t in the tile-sorted instance list (zeroed beforehand).
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t *__restrict__ sorted_tile_keys, uint32_t n,
                   uint2 *__restrict__ ranges) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t t = sorted_tile_keys[j];
  if (j == 0 || sorted_tile_keys[j - 1] != t) ranges[t].x = j;
  if (j + 1 == n || sorted_tile_keys[j + 1] != t) ranges[t].y = j + 1;
}

}  // namespace splat
