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

// number of FAILED tiles inside an inclusive tile rectangle, from the summed-area table of the
// failed-tile bitmap (pass_b_setup_kernel): sat[(y) * pitch + x] = failed tiles in [0,y) x [0,x)
SPLAT_DEVINL uint32_t failed_in_rect(const int *__restrict__ sat, uint32_t pitch, const TileRect t) {
  if (t.x1 < t.x0 || t.y1 < t.y0) return 0u;
  const uint32_t x0 = t.x0, y0 = t.y0, x1 = t.x1 + 1u, y1 = t.y1 + 1u;
  return (uint32_t)(sat[y1 * pitch + x1] - sat[y0 * pitch + x1] - sat[y1 * pitch + x0] + sat[y0 * pitch + x0]);
}

// cnt[r] = number of tiles of the Gaussian at depth rank r (0 for culled ones, whose key
// 0xFFFFFFFF sorted them to the end).  Also counts the visible Gaussians.
//   first / only pass : ranks below rank_cut (the farthest Gaussians) get instances only for the OPEN
//                       tiles (open_sat / rank_rects, see far_prefix_kernel); the others for every tile
//   second pass (sat) : every rank, but only the tiles the first pass marked failed
__global__ void __launch_bounds__(256)
tile_count_kernel(const uint32_t *__restrict__ sorted_keys, const uint32_t *__restrict__ order,
                  const uint32_t *__restrict__ tcnt, uint32_t *__restrict__ cnt, uint32_t n,
                  const uint32_t *__restrict__ n_sorted, uint32_t rank_cut, FrameStatus *__restrict__ status,
                  const uint2 *__restrict__ rects, const int *__restrict__ sat, uint32_t sat_pitch,
                  const uint2 *__restrict__ rank_rects, const int *__restrict__ open_sat) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  // stripe renders sort only *n_sorted pairs; the tail of the buffers is stale
  const uint32_t ns = n_sorted ? *n_sorted : n;
  bool vis = false;
  if (r < n) {
    vis = r < ns && sorted_keys[r] != KEY_CULLED;
    uint32_t c = 0;
    if (vis && r >= rank_cut) {
      if (sat) c = failed_in_rect(sat, sat_pitch, unpack_rect(__ldg(&rects[order[r]])));
      else c = __ldg(&tcnt[order[r]]);
    } else if (vis && open_sat) {
      c = failed_in_rect(open_sat, sat_pitch, unpack_rect(__ldg(&rank_rects[r])));   // cut Gaussian: open tiles only
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
                 uint32_t tiles_x, uint32_t tiles_y, int *__restrict__ diff /* (tiles_y+1) x (tiles_x+1), zeroed */,
                 uint2 *__restrict__ rank_rects /* rect of the Gaussian at depth rank r, r < rank_cut (coalesced for the next kernels) */) {
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
      if (rb + u * FC_THREADS < r1) rank_rects[rb + u * FC_THREADS] = rc[u];
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
// far_cnt[ty * tiles_x + tx], the number of cut Gaussians whose quad touches the tile.
//
// OPEN tiles.  A tile that only a few cut Gaussians touch (0 < far_cnt <= NEAR_OPEN_MAX_FAR) is
// where the near lists fail: measured on the bench orbit, every tile that did not converge had 1..5
// cut Gaussians and a short, sparse near list (profiles/r2_near_cut_failures.jsonl).  Binning those
// few Gaussians is nearly free, so such a tile is opened: the cut Gaussians are binned for it too
// (tile_count / emit_masked test the ranks below the cut against the open-tile bitmap), its list is
// complete again and far_cnt is reported as 0.  open_sat = summed-area table of the open bitmap
// (in global memory: two tables do not fit shared memory at 4K).
constexpr uint32_t NEAR_OPEN_MAX_FAR = 64;
__global__ void __launch_bounds__(1024)
far_prefix_kernel(const int *__restrict__ diff, uint32_t tiles_x, uint32_t tiles_y, uint32_t *__restrict__ far_cnt,
                  FrameStatus *__restrict__ status, uint32_t *__restrict__ tile_open, int *__restrict__ open_sat) {
  extern __shared__ int s_diff[];
  const uint32_t pitch = tiles_x + 1u, cells = pitch * (tiles_y + 1u);
  for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) { s_diff[i] = diff[i]; open_sat[i] = 0; }
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
    const uint32_t ty = i / tiles_x, tx = i % tiles_x;
    uint32_t v = (uint32_t)s_diff[ty * pitch + tx];
    const bool open = v != 0u && v <= NEAR_OPEN_MAX_FAR;
    tile_open[i] = open ? 1u : 0u;
    open_sat[(ty + 1u) * pitch + tx + 1u] = open ? 1 : 0;
    if (open) v = 0u;                  // its list will be complete
    far_cnt[i] = v;
    cut += v;
  }
  for (int o = 16; o > 0; o >>= 1) cut += __shfl_xor_sync(0xFFFFFFFFu, cut, o);
  if ((threadIdx.x & 31u) == 0 && cut) atomicAdd(&status->n_cut, cut);
  __syncthreads();
  for (uint32_t y = 1u + threadIdx.x; y <= tiles_y; y += blockDim.x) {
    int run = 0;
    for (uint32_t x = 1; x <= tiles_x; ++x) { run += open_sat[y * pitch + x]; open_sat[y * pitch + x] = run; }
  }
  __syncthreads();
  for (uint32_t x = 1u + threadIdx.x; x <= tiles_x; x += blockDim.x) {
    int run = 0;
    for (uint32_t y = 1; y <= tiles_y; ++y) { run += open_sat[y * pitch + x]; open_sat[y * pitch + x] = run; }
  }
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
                      uint32_t tiles_x, const uint32_t *__restrict__ n_eff, uint32_t rank_cut) {
  __shared__ uint32_t s_off[257];
  if (*n_eff == 0u) return;     // nothing to emit, or the pairs do not fit the buffers (frame skipped)
  __shared__ uint32_t s_idx[256];
  __shared__ uint2 s_rect[256];
  const uint32_t tid = threadIdx.x, r0 = blockIdx.x * 256u, r = r0 + tid;
  // near cut: the ranks below rank_cut own (few) instances in open tiles only -- written by
  // emit_masked_kernel; this kernel starts at the first rank at or above the cut
  if (r0 + 256u <= rank_cut) return;
  const uint32_t first = rank_cut > r0 ? rank_cut - r0 : 0u;
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
  const uint32_t begin = s_off[first], end = s_off[256];
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

// Second near-cut pass: only the tiles the first pass marked failed get instances.  Most depth ranks
// touch none of them (cnt == 0) and cost nothing; for the others the WARP walks the rank's tile
// rectangle together -- 32 tiles per step, rows without a failed tile skipped through the
// summed-area table -- so that a quad covering the whole screen is not one thread's 8,000-step loop
// (first version: 532 us of a 700 us pass, profiles/r2k).  The failed tiles of a step are written
// with a ballot-ranked, coalesced store.  Order inside one Gaussian's instances is irrelevant: the
// tile sort separates them.
// Used twice: for the ranks below the cut against the OPEN tiles (first pass; rank_rects holds their
// rectangles by rank), and for every rank against the FAILED tiles (second pass).
__global__ void __launch_bounds__(256)
emit_masked_kernel(const uint32_t *__restrict__ order, const uint2 *__restrict__ rects, const uint2 *__restrict__ rank_rects,
                   const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ offs,
                   uint32_t *__restrict__ inst_keys, uint32_t *__restrict__ inst_vals, uint32_t n,
                   uint32_t tiles_x, const uint32_t *__restrict__ tile_failed, const int *__restrict__ sat,
                   const uint32_t *__restrict__ n_eff) {
  if (*n_eff == 0u) return;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
  const uint32_t my_cnt = r < n ? cnt[r] : 0u;
  uint32_t todo = __ballot_sync(0xFFFFFFFFu, my_cnt != 0u);
  if (todo == 0u) return;
  const uint32_t my_gi = my_cnt ? order[r] : 0u;
  const uint2 my_rect = my_cnt ? (rank_rects ? rank_rects[r] : rects[my_gi]) : make_uint2(0u, 0u);
  const uint32_t my_off = my_cnt ? offs[r] : 0u;
  const uint32_t pitch = tiles_x + 1u, lt_mask = (1u << lane) - 1u;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1u;
    const uint32_t gi = __shfl_sync(0xFFFFFFFFu, my_gi, src);
    const uint32_t rx = __shfl_sync(0xFFFFFFFFu, my_rect.x, src), ry = __shfl_sync(0xFFFFFFFFu, my_rect.y, src);
    uint32_t o = __shfl_sync(0xFFFFFFFFu, my_off, src);
    const uint32_t x0 = rx & 0xFFFFu, y0 = rx >> 16, x1 = ry & 0xFFFFu, y1 = ry >> 16;
    for (uint32_t y = y0; y <= y1; ++y) {
      // failed tiles of this row inside [x0, x1], from the summed-area table
      const int in_row = sat[(y + 1u) * pitch + x1 + 1u] - sat[y * pitch + x1 + 1u] - sat[(y + 1u) * pitch + x0] + sat[y * pitch + x0];
      if (in_row == 0) continue;
      for (uint32_t xb = x0; xb <= x1; xb += 32u) {
        const uint32_t x = xb + lane;
        const uint32_t tile = y * tiles_x + x;
        const bool hit = x <= x1 && __ldg(&tile_failed[tile]) != 0u;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
        if (hit) {
          const uint32_t p = o + __popc(bal & lt_mask);
          inst_keys[p] = tile;
          inst_vals[p] = gi;
        }
        o += __popc(bal);
      }
    }
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
