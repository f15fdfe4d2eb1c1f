// splat_api.cu -- the C ABI of include/splat.h: context, scene upload, frame orchestration.
//
// Frame = render_to_buffer (pipelines.rs:66-86 / :260-280):
//   K1 project -> K3a depth radix sort (N keys) -> K2 tile count + scan -> [host reads the
//   instance count] -> K2 emit -> K3b tile radix sort (I keys) -> K4 ranges -> K5 blend.
// One host<->device round trip per frame (the instance count), everything else is enqueued
// asynchronously on one stream; the framebuffer upload runs on a second stream and is only
// waited for by the blend kernel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <new>
#include <string>
#include <thread>

#include "../../include/splat.h"
#include "bin.cuh"
#include "blend.cuh"
#include "blend_float.cuh"
#include "common.cuh"
#include "project.cuh"
#include "sort.cuh"

using namespace splat;

namespace {
constexpr uint64_t NEAR_CUT_MIN_VISIBLE = 200000;   // smaller scenes: the cut's own kernels cost more than they save
constexpr size_t FAR_SMEM_MAX = 200 * 1024;            // difference array of far_cover_kernel (shared memory)
constexpr uint32_t NEAR_CUT_DEFAULT = 128;             // 1/8 of the Gaussians
enum { EV_START = 0, EV_PROJECT, EV_DSORT, EV_COUNT, EV_EMIT, EV_TSORT, EV_RANGES, EV_BLEND,
       EV_H2D0, EV_H2D1, EV_D2H0, EV_D2H1, EV_COUNT_ };
}

struct splat_ctx {
  splat_config cfg{};
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev[EV_COUNT_] = {};
  cudaEvent_t status_ev = nullptr, h2d_done = nullptr;
  std::string err;

  uint32_t n = 0;
  float4 *scene = nullptr;
  Rec *recs = nullptr;
  uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr};
  uint2 *rects = nullptr;
  uint32_t *tcnt = nullptr;        // tiles per Gaussian (by Gaussian index)
  uint32_t *block_kept = nullptr;  // stripe renders: survivors per project CTA, then their exclusive scan
  uint32_t *cnt = nullptr, *offs = nullptr;
  uint32_t *hist = nullptr; size_t hist_cap = 0;
  uint32_t *tot = nullptr;          // 256 digit totals of the current radix pass
  uint32_t *partial = nullptr; size_t partial_cap = 0;
  uint64_t inst_cap = 0;
  uint32_t *ikeys[2] = {nullptr, nullptr}, *ivals[2] = {nullptr, nullptr};
  uint2 *ranges = nullptr; size_t ranges_cap = 0;
  uint2 *units = nullptr;          // blend work units, heaviest first (up to 4 per tile)
  uint32_t *far_cnt = nullptr;     // near cut: cut Gaussians per tile
  uint32_t *tile_failed = nullptr; // near cut: tiles that need the complete lists
  int *far_diff = nullptr;         // its 2-D difference array
  size_t far_cells_cap = 0;
  uint32_t cut_frac = 1024;        // Gaussians binned by the near-cut pass, in 1/1024 (1024 = no cut)
  uint32_t last_cut = 0;           // rank_cut of the last frame
  uint32_t last_failed = 0;        // groups / tiles that did not converge in its near-cut pass
  uint64_t last_cut_instances = 0; // (tile, Gaussian) pairs it did not bin
  uint32_t *n_units = nullptr;
  FrameStatus *d_status = nullptr, *h_status = nullptr;
  uint32_t *h_wd = nullptr, *d_wd = nullptr;   // blend watchdog record (mapped pinned host memory, blend.cuh)
  bool debug_sync = false;                     // SPLAT_DEBUG_SYNC=1: bounded wait after every launch, names the kernel that hangs
  double wait_limit_s = 30.0;                  // SPLAT_WAIT_LIMIT_S: bound of every host wait on the device
  uint32_t *d_fb = nullptr; size_t fb_cap = 0;

  float4 *d_tap = nullptr; size_t tap_cap = 0; bool want_tap = false;   // float mode: un-quantised result per pixel (tests)

  // last frame
  int order_buf = 0;          // which vals[] holds the depth order
  bool have_frame = false, host_copy = false;
  bool status_pending = false;   // a frame's status copy is in flight (status_ev)
  bool retry_pending = false;    // a frame was skipped on the device (instance buffers too small) and not yet repeated
  bool loads_valid = false;      // c->ranges describes the complete tile lists of the last frame
  uint32_t skipped_seen = 0;
  uint64_t frames_skipped = 0;
  uint32_t geom[4] = {0, 0, 0, 0};   // W, H, row0, row1 of the last frame
  FrameParams last_params{};
  uint32_t *last_fb = nullptr;
  cudaStream_t last_stream = nullptr;
  uint32_t retried = 0;
  uint64_t launches = 0, last_instances = 0, last_visible = 0, last_tiles = 0;
};

namespace {

int fail(splat_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
  }
  return code;
}

#define CU(expr)                                                              \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) return fail(c, e__ == cudaErrorMemoryAllocation ? SPLAT_ERR_NOMEM : SPLAT_ERR_CUDA, #expr, e__); \
  } while (0)

// Bounded host wait: polls instead of blocking, so that a kernel that never finishes becomes an
// error code with a message (which wait, and the blend watchdog's record if it tripped), not a hang.
int wait_done(splat_ctx *c, cudaEvent_t ev, cudaStream_t s, const char *what) {
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0;; ++spins) {
    const cudaError_t e = ev ? cudaEventQuery(ev) : cudaStreamQuery(s);
    if (e == cudaSuccess) return SPLAT_OK;
    if (e != cudaErrorNotReady) {
      int rc = fail(c, SPLAT_ERR_CUDA, what, e);
      if (c->h_wd && c->h_wd[0]) {
        char buf[160];
        std::snprintf(buf, sizeof buf, " [blend watchdog: block %u thread %u role %u slot %u chunk %u parity %u; state:", c->h_wd[1],
                      c->h_wd[2], c->h_wd[3] >> 28, (c->h_wd[3] >> 20) & 0xFFu, c->h_wd[3] & 0xFFFFFu, c->h_wd[4]);
        c->err += buf;
        for (uint32_t i = 0; i < c->h_wd[5] && i < (uint32_t)WD_WORDS - 8u; ++i) {
          std::snprintf(buf, sizeof buf, " %08x", c->h_wd[8 + i]);
          c->err += buf;
        }
        c->err += "]";
      }
      return rc;
    }
    if (spins > 2000) {
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt > c->wait_limit_s) {
        c->err = std::string("timeout waiting for the device: ") + what;
        return SPLAT_ERR_CUDA;
      }
      if (dt > 0.002) std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  }
}
#define LAUNCHED(name)                                                        \
  do {                                                                        \
    c->launches += 1;                                                         \
    if (c->debug_sync) {                                                      \
      int rc__ = wait_done(c, nullptr, s, name);                              \
      if (rc__) return rc__;                                                  \
    }                                                                         \
  } while (0)

template <typename T>
cudaError_t dev_alloc(T **p, size_t count) {
  return cudaMalloc(reinterpret_cast<void **>(p), std::max<size_t>(count, 1) * sizeof(T));
}
template <typename T>
void dev_free(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// stable LSD radix sort of (key, value) pairs on bits [0, bits); returns the buffer index
// (0/1) that holds the result, or a negative error code (SPLAT_DEBUG_SYNC runs).  The kernels take
// the pair count from *n_ptr (or n_fixed if n_ptr is null); `n_grid` -- the number of pairs the
// launch is sized for -- must be an upper bound of it (the caller guarantees that on the device).
int radix_sort(splat_ctx *c, cudaStream_t s, uint32_t *keys[2], uint32_t *vals[2], uint64_t n_grid, int bits,
               int cur, const uint32_t *n_ptr, uint32_t n_fixed) {
  const uint32_t grid = std::max(1u, std::min(cdiv(n_grid, RS_BLOCK), 1u << 20));
  for (int shift = 0; shift < bits; shift += 8) {
    const int nbits = std::min(8, bits - shift);
    rs_hist_kernel<<<grid, RS_THREADS, 0, s>>>(keys[cur], n_ptr, n_fixed, shift, (1u << nbits) - 1u, c->hist);
    rs_rowscan_kernel<<<256, RW_THREADS, 0, s>>>(c->hist, n_ptr, n_fixed, c->tot);
    if (nbits == 8)
      rs_scatter_kernel<8><<<grid, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                                                       n_ptr, n_fixed, shift, 8, c->hist, c->tot);
    else
      rs_scatter_kernel<0><<<grid, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                                                       n_ptr, n_fixed, shift, nbits, c->hist, c->tot);
    c->launches += 2;
    LAUNCHED("radix sort pass (hist, rowscan, scatter)");
    cur ^= 1;
  }
  return cur;
}

int ensure_scratch(splat_ctx *c, uint64_t sort_items) {
  const size_t need_hist = 256ull * cdiv(sort_items, RS_BLOCK);
  if (need_hist > c->hist_cap) {
    dev_free(c->hist);
    CU(dev_alloc(&c->hist, need_hist));
    c->hist_cap = need_hist;
  }
  const size_t need_part = std::max<size_t>(cdiv(need_hist, SC_BLOCK), cdiv(c->n, SC_BLOCK)) + 1;
  if (need_part > c->partial_cap) {
    dev_free(c->partial);
    CU(dev_alloc(&c->partial, need_part));
    c->partial_cap = need_part;
  }
  return SPLAT_OK;
}

int ensure_instances(splat_ctx *c, uint64_t want) {
  if (want <= c->inst_cap) return SPLAT_OK;
  if (want >= 0xFFFFFFFFull) return fail(c, SPLAT_ERR_UNSUPPORTED, "more than 2^32-1 tile instances in one stripe");
  uint64_t cap = std::min<uint64_t>(0xFFFFFFFEull, want + want / 4 + 4096);
  for (int k = 0; k < 2; ++k) { dev_free(c->ikeys[k]); dev_free(c->ivals[k]); }
  c->inst_cap = 0;
  for (int k = 0; k < 2; ++k) {
    CU(dev_alloc(&c->ikeys[k], cap));
    CU(dev_alloc(&c->ivals[k], cap));
  }
  c->inst_cap = cap;
  return ensure_scratch(c, std::max<uint64_t>(cap, c->n));
}

void free_scene(splat_ctx *c) {
  cudaDeviceSynchronize();   // frames may still be in flight on a caller's stream
  dev_free(c->scene); dev_free(c->recs); dev_free(c->rects); dev_free(c->tcnt); dev_free(c->block_kept); dev_free(c->cnt); dev_free(c->offs);
  for (int k = 0; k < 2; ++k) { dev_free(c->keys[k]); dev_free(c->vals[k]); }
  c->n = 0;
  c->have_frame = false;
  c->status_pending = false;
  c->retry_pending = false;
  c->loads_valid = false;
}

int alloc_scene(splat_ctx *c, uint64_t n) {
  if (n == 0 || n > 0x7FFFFFFFull) return fail(c, SPLAT_ERR_INVALID, "n must be in [1, 2^31)");
  free_scene(c);
  CU(dev_alloc(&c->scene, (size_t)SCENE_PLANES * n));
  CU(dev_alloc(&c->recs, n));
  CU(dev_alloc(&c->rects, n));
  CU(dev_alloc(&c->tcnt, n));
  CU(dev_alloc(&c->block_kept, cdiv(n, 256)));
  CU(dev_alloc(&c->cnt, n));
  CU(dev_alloc(&c->offs, n));
  for (int k = 0; k < 2; ++k) { CU(dev_alloc(&c->keys[k], n)); CU(dev_alloc(&c->vals[k], n)); }
  c->n = (uint32_t)n;
  int rc = ensure_scratch(c, n);
  if (rc) return rc;
  return ensure_instances(c, c->cfg.max_instances ? std::max<uint64_t>(c->cfg.max_instances, 4096u) : (1u << 20));
}

int make_params(splat_ctx *c, const splat_camera *cam, uint32_t W, uint32_t H, uint32_t row0,
                uint32_t row1, FrameParams *P) {
  if (!cam) return fail(c, SPLAT_ERR_INVALID, "camera is null");
  if (W == 0 || H == 0 || W > 65535u * TILE || H > 65535u * TILE) return fail(c, SPLAT_ERR_INVALID, "bad target size");
  if (!(cam->w == (float)W && cam->h == (float)H))
    return fail(c, SPLAT_ERR_UNSUPPORTED, "camera.w/h must equal the target size");
  if (row0 >= row1 || row1 > H || row0 % TILE != 0 || (row1 % TILE != 0 && row1 != H))
    return fail(c, SPLAT_ERR_INVALID, "stripe [row0,row1) must be non-empty, inside the image and tile aligned");
  std::memcpy(P->view, cam->view, sizeof(P->view));
  std::memcpy(P->proj, cam->proj, sizeof(P->proj));
  std::memcpy(P->cam_pos, cam->position, sizeof(P->cam_pos));
  P->focal = cam->focal; P->htanx = cam->htanx; P->htany = cam->htany;
  P->lowpass = c->cfg.lowpass;
  P->sample_off = c->cfg.sample_offset;
  P->ysign = c->cfg.y_down ? 1.0f : -1.0f;
  P->zclip_mode = c->cfg.zclip_mode;
  P->stripe_cull = (row0 != 0 || row1 != H) ? 1 : 0;
  P->W = W; P->H = H; P->row0 = row0; P->row1 = row1;
  P->tiles_x = cdiv(W, TILE);
  P->tile_y0 = row0 / TILE;
  P->tiles_y = cdiv(row1, TILE) - P->tile_y0;
  P->n = c->n;
  P->nz2 = 0x8000000080000000ull;
  if ((uint64_t)P->tiles_x * P->tiles_y >= (1ull << 24))
    return fail(c, SPLAT_ERR_UNSUPPORTED, "more than 2^24 tiles in one stripe");
  return SPLAT_OK;
}

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t *p, uint32_t v, size_t n) {
  const size_t i0 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i0 + k < n) p[i0 + k] = v;
}

int ilog2_ceil(uint32_t v) {
  int b = 0;
  while ((1ull << b) < v) ++b;
  return std::max(b, 1);
}

// What the host learnt from a finished frame's status block (copied to pinned memory behind the
// blend kernel).  Called wherever the host has waited for the frame anyway, or finds it finished.
void absorb_status(splat_ctx *c) {
  const FrameStatus &fs = *c->h_status;
  c->last_instances = fs.n_instances;
  c->last_visible = fs.n_visible;
  if (fs.skipped != c->skipped_seen) {          // a frame (or several) wanted more pairs than the buffers hold
    c->frames_skipped += fs.skipped - c->skipped_seen;
    c->skipped_seen = fs.skipped;
    c->retry_pending = true;
  }
  c->status_pending = false;
}
void poll_status(splat_ctx *c) {
  if (c->status_pending && cudaEventQuery(c->status_ev) == cudaSuccess) absorb_status(c);
}

struct Pass {
  uint32_t rank_cut = 0;            // near cut: depth ranks below this get no instances (0 = complete lists)
  const TileRect *only_box = nullptr;
  bool only_failed = false;
  bool sync_count = true;           // host reads the instance count mid-frame (exact grids, buffers grown on the spot)
};

// Second half of a frame on `s`: bin the Gaussians of depth rank >= rank_cut into tiles, sort the
// instances by tile, blend.
//   sync_count = true : the host waits for the instance count (one round trip), grows the buffers if
//                       needed and sizes the launches exactly.  First frame of a target geometry,
//                       near-cut frames, and the repeat of a skipped frame.
//   sync_count = false: nothing blocks.  Launches are sized from the previous frame (+12.5%); the
//                       kernels read the count from device memory, and a frame whose pairs do not
//                       fit that bound blends NOTHING (status.overflow) and is repeated by the host.
int bin_sort_blend(splat_ctx *c, const FrameParams &P, uint32_t *fb_rows_dev, cudaStream_t s, cudaEvent_t wait_ev,
                   int cur, const uint32_t *n_sorted, const Pass &pass) {
  TileRect box;                      // tiles this pass bins into (stripe-local tile coordinates)
  box.x0 = 0; box.y0 = 0; box.x1 = 0xFFFF; box.y1 = 0xFFFF;
  if (pass.only_box) box = *pass.only_box;
  const uint32_t rank_cut = pass.rank_cut;
  const uint32_t n = c->n;
  const uint32_t T = P.tiles_x * P.tiles_y;
  const bool flt = c->cfg.blend_mode == SPLAT_BLEND_FLOAT;
  if (T > c->ranges_cap) {
    CU(cudaStreamSynchronize(s));
    dev_free(c->ranges);
    dev_free(c->units);
    dev_free(c->far_cnt);
    dev_free(c->tile_failed);
    CU(dev_alloc(&c->ranges, T));
    CU(dev_alloc(&c->units, (size_t)4 * T));
    CU(dev_alloc(&c->far_cnt, T));
    CU(dev_alloc(&c->tile_failed, T));
    c->ranges_cap = T;
  }
  {
    const size_t cells = (size_t)(P.tiles_x + 1) * (P.tiles_y + 1);   // not a function of T alone
    if (cells > c->far_cells_cap) {
      CU(cudaStreamSynchronize(s));
      dev_free(c->far_diff);
      CU(dev_alloc(&c->far_diff, cells));
      c->far_cells_cap = cells;
    }
  }
  // n_instances, n_visible, n_failed, fail box, n_cut (the fields behind them belong to the frame)
  CU(cudaMemsetAsync(c->d_status, 0, offsetof(FrameStatus, n_sort), s));
  const uint32_t *far = nullptr;
  if (rank_cut) {
    const size_t cells = (size_t)(P.tiles_x + 1) * (P.tiles_y + 1);
    CU(cudaMemsetAsync(c->far_diff, 0, cells * sizeof(int), s));
    CU(cudaMemsetAsync(c->tile_failed, 0, (size_t)T * sizeof(uint32_t), s));
    far_cover_kernel<<<148, FC_THREADS, cells * sizeof(int), s>>>(c->vals[cur], c->rects, n_sorted, n, rank_cut,
                                                                  P.tiles_x, P.tiles_y, c->far_diff);
    far_prefix_kernel<<<1, 1024, cells * sizeof(int), s>>>(c->far_diff, P.tiles_x, P.tiles_y, c->far_cnt, c->d_status);
    c->launches += 1;
    LAUNCHED("far_cover / far_prefix");
    far = c->far_cnt;
  }
  // no-round-trip frames: launches are sized for 1.125x the last count the host has seen (+64k); a
  // frame that wants more is skipped on the device (overflow) and repeated with a round trip
  const uint64_t n_bound = std::min<uint64_t>(c->inst_cap, c->last_instances + c->last_instances / 8 + 65536u);
  tile_count_kernel<<<cdiv(n, 256), 256, 0, s>>>(c->keys[cur], c->vals[cur], c->tcnt, c->cnt, n, n_sorted, rank_cut, c->d_status,
                                                  pass.only_box ? c->rects : nullptr, box);
  LAUNCHED("tile_count_kernel");
  // exclusive scan cnt -> offs; grand total -> status.n_instances, checked against the buffer
  // capacity on the device (status.n_inst_eff / overflow)
  {
    const uint32_t np = std::max(1u, cdiv(n, SC_BLOCK));
    scan_reduce_kernel<<<np, SC_THREADS, 0, s>>>(c->cnt, c->partial, n);
    scan_partials_kernel<<<1, 1024, 0, s>>>(c->partial, np, &c->d_status->n_instances, c->d_status,
                                            pass.sync_count ? 0xFFFFFFFEull : (unsigned long long)n_bound);
    scan_apply_kernel<<<np, SC_THREADS, 0, s>>>(c->cnt, c->offs, c->partial, n);
    c->launches += 2;
    LAUNCHED("scan (tile counts)");
  }
  CU(cudaEventRecord(c->ev[EV_COUNT], s));
  uint64_t n_grid;                   // instances the launches below are sized for
  if (pass.sync_count) {
    CU(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(FrameStatus), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(c->status_ev, s));
    { int rcw = wait_done(c, c->status_ev, s, "tile count (first half of the frame)"); if (rcw) return rcw; }
    const uint64_t I = c->h_status->n_instances;
    c->last_instances = I;
    c->last_visible = c->h_status->n_visible;
    if (I >= 0xFFFFFFFEull) return fail(c, SPLAT_ERR_UNSUPPORTED, "more than 2^32-2 tile instances in one stripe");
    if (I > c->inst_cap) {
      int rc = ensure_instances(c, I);
      if (rc) return rc;
      c->retried += 1;
    }
    n_grid = I;
  } else {
    n_grid = n_bound;
  }
  c->last_tiles = (uint64_t)T;
  const uint32_t *n_eff = &c->d_status->n_inst_eff;
  emit_instances_kernel<<<cdiv(n, 256), 256, 0, s>>>(c->vals[cur], c->rects, c->cnt, c->offs,
                                                      c->ikeys[0], c->ivals[0], n, P.tiles_x, box, n_eff);
  LAUNCHED("emit_instances_kernel");
  CU(cudaEventRecord(c->ev[EV_EMIT], s));
  int icur = 0;
  if (n_grid > 0) icur = radix_sort(c, s, c->ikeys, c->ivals, n_grid, ilog2_ceil(T), 0, n_eff, 0);
  if (icur < 0) return icur;
  CU(cudaEventRecord(c->ev[EV_TSORT], s));
  CU(cudaMemsetAsync(c->ranges, 0, (size_t)T * sizeof(uint2), s));
  tile_ranges_kernel<<<std::max(1u, std::min(cdiv(n_grid, 1024), 1u << 20)), 256, 0, s>>>(c->ikeys[icur], n_eff, c->ranges);
  LAUNCHED("tile_ranges_kernel");
  unit_order_kernel<<<1, 1024, 0, s>>>(c->ranges, T, c->units, c->n_units, &c->d_status->n_instances, far, c->d_status, P.tiles_x,
                                       c->tile_failed, pass.only_failed ? 1 : 0, flt ? 1 : 0);   // heaviest first
  LAUNCHED("unit_order_kernel");
  CU(cudaEventRecord(c->ev[EV_RANGES], s));
  if (wait_ev) CU(cudaStreamWaitEvent(s, wait_ev, 0));
  if (flt) {
    blend_float_kernel<<<T, BF_THREADS, 0, s>>>(c->ranges, c->units, c->n_units, c->ivals[icur], c->recs, fb_rows_dev, P,
                                                c->want_tap ? c->d_tap : nullptr, c->d_wd);
    LAUNCHED("blend_float_kernel");
  } else {
    blend_kernel<<<4 * T, BL_THREADS, BL_SMEM_BYTES, s>>>(c->ranges, c->units, c->n_units, c->ivals[icur], c->recs, fb_rows_dev, P,
                                                          far, c->d_status, c->tile_failed, c->d_wd);
    LAUNCHED("blend_kernel");
  }
  CU(cudaEventRecord(c->ev[EV_BLEND], s));
  CU(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(FrameStatus), cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(c->status_ev, s));
  c->status_pending = true;
  if (rank_cut) {
    { int rcw = wait_done(c, c->status_ev, s, "near-cut pass (bin, sort, blend)"); if (rcw) return rcw; }   // did every pixel converge on the near lists?
    c->status_pending = false;
  }
  return SPLAT_OK;
}

// Enqueue one frame on `s`, writing rows [row0,row1) into fb_rows_dev.  If wait_ev is set the
// blend kernel waits for it (framebuffer upload on the copy stream).  force_sync: use the
// synchronous count path whatever the history says (the repeat of a skipped frame).
int render_frame(splat_ctx *c, const FrameParams &P, uint32_t *fb_rows_dev, cudaStream_t s, cudaEvent_t wait_ev,
                 bool force_sync = false) {
  const uint32_t n = c->n;
  c->launches = 0;
  c->retried = 0;
  poll_status(c);
  // Same target geometry as the last frame and its count known: nothing has to block.
  const bool same_geom = c->have_frame && c->geom[0] == P.W && c->geom[1] == P.H && c->geom[2] == P.row0 && c->geom[3] == P.row1;
  bool async = same_geom && !force_sync && !c->retry_pending && c->cut_frac >= 1024u && !c->cfg.sync_frames;
  if (async && c->last_instances + c->last_instances / 8 + 65536u > c->inst_cap) {
    // cudaFree / cudaMalloc wait for the frames in flight; rare (the buffers are grown with 25% headroom)
    int rc = ensure_instances(c, c->last_instances + c->last_instances / 8 + 65536u);
    if (rc) return rc;
  }
  CU(cudaEventRecord(c->ev[EV_START], s));
  CU(cudaMemsetAsync(c->d_status, 0, offsetof(FrameStatus, skipped), s));
  project_kernel<<<cdiv(n, 256), 256, 0, s>>>(c->scene, P, c->recs, c->keys[0], c->vals[0], c->rects, c->tcnt, c->block_kept);
  LAUNCHED("project_kernel");
  CU(cudaEventRecord(c->ev[EV_PROJECT], s));
  int cur = 0;
  const uint32_t *n_sorted = nullptr;
  if (P.stripe_cull) {
    // squeeze out what the stripe cannot see, then sort only the survivors (device-side count)
    const uint32_t nb = cdiv(n, 256), np = std::max(1u, cdiv(nb, SC_BLOCK));
    scan_reduce_kernel<<<np, SC_THREADS, 0, s>>>(c->block_kept, c->partial, nb);
    scan_partials_kernel<<<1, 1024, 0, s>>>(c->partial, np, &c->d_status->n_sort);
    scan_apply_kernel<<<np, SC_THREADS, 0, s>>>(c->block_kept, c->block_kept, c->partial, nb);
    compact_pairs_kernel<<<nb, 256, 0, s>>>(c->keys[0], c->vals[0], c->keys[1], c->vals[1], c->block_kept, n);
    c->launches += 3;
    LAUNCHED("stripe compaction");
    cur = 1;
    n_sorted = reinterpret_cast<const uint32_t *>(&c->d_status->n_sort);   // low word (n < 2^31)
  }
  // (a stripe sorts only its survivors -- a device-side count; the CTAs beyond it exit at once)
  cur = radix_sort(c, s, c->keys, c->vals, n, 32, cur, n_sorted, n);
  if (cur < 0) return cur;
  c->order_buf = cur;
  CU(cudaEventRecord(c->ev[EV_DSORT], s));

  // Near cut: the blend reads only the nearest few hundred entries of every tile list (exact early
  // termination), so first bin + sort only the nearest cut_frac/1024 of the Gaussians -- the
  // depth ranks the previous frame makes us expect at the top -- and fall back to the complete
  // lists only if some pixel did not converge on them (then keep twice as many from now on).
  uint32_t rank_cut = 0;
  const size_t far_smem = (size_t)(P.tiles_x + 1) * (P.tiles_y + 1) * sizeof(int);
  const uint64_t min_visible = c->cfg.near_cut > 0 ? 1u : NEAR_CUT_MIN_VISIBLE;   // a fixed fraction is honoured on any scene (tests)
  if (c->cut_frac < 1024u && c->have_frame && c->last_visible >= min_visible && far_smem <= FAR_SMEM_MAX) {
    const uint64_t keep = (c->last_visible * c->cut_frac + 1023u) / 1024u;
    if (keep < c->last_visible) rank_cut = (uint32_t)(c->last_visible - keep);
  }
  Pass pass;
  pass.rank_cut = rank_cut;
  pass.sync_count = !async;
  int rc = bin_sort_blend(c, P, fb_rows_dev, s, wait_ev, cur, n_sorted, pass);
  if (rc) return rc;
  c->last_cut = rank_cut;
  c->last_failed = rank_cut ? c->h_status->n_failed : 0u;
  c->last_cut_instances = rank_cut ? c->h_status->n_cut : 0ull;
  c->loads_valid = true;
  if (rank_cut && (c->h_status->n_failed != 0 || c->h_status->n_instances == 0)) {
    // The near lists were not enough for this view: bin + sort + blend again with ALL Gaussians,
    // restricted to the bounding box of the tiles that did not converge when that box is small
    // (typically a strip along one screen edge).  Groups that already converged wrote final values
    // and, if they are blended again, converge to them again without reading the framebuffer;
    // groups that did not converge left their pixels untouched.
    const FrameStatus fs = *c->h_status;
    TileRect fbx;
    fbx.x0 = (uint16_t)~fs.fail_ix0; fbx.y0 = (uint16_t)~fs.fail_iy0;
    fbx.x1 = (uint16_t)fs.fail_x1;   fbx.y1 = (uint16_t)fs.fail_y1;
    const bool have_box = fs.n_failed != 0 && fs.n_instances != 0 && fbx.x1 >= fbx.x0 && fbx.y1 >= fbx.y0;
    const uint64_t area = have_box ? (uint64_t)(fbx.x1 - fbx.x0 + 1) * (fbx.y1 - fbx.y0 + 1) : ~0ull;
    const bool partial = have_box && area * 2u <= (uint64_t)P.tiles_x * P.tiles_y;
    if (!partial && c->cfg.near_cut < 0) c->cut_frac = std::min<uint32_t>(1024u, c->cut_frac * 2u);
    c->retried += 1;
    const uint64_t near_instances = fs.n_instances;
    Pass again;
    again.only_box = partial ? &fbx : nullptr;
    again.only_failed = fs.n_instances != 0;
    rc = bin_sort_blend(c, P, fb_rows_dev, s, wait_ev, cur, n_sorted, again);
    if (rc) return rc;
    if (partial) {
      c->last_instances += near_instances;     // both passes' instances were binned and sorted
      c->loads_valid = false;                   // c->ranges holds the box-restricted lists only
    }
  }
  CU(cudaGetLastError());
  c->have_frame = true;
  c->geom[0] = P.W; c->geom[1] = P.H; c->geom[2] = P.row0; c->geom[3] = P.row1;
  c->last_params = P; c->last_fb = fb_rows_dev; c->last_stream = s;
  return SPLAT_OK;
}

// Wait for the last enqueued frame; if it was skipped (its tile instances did not fit the buffers:
// only possible on the no-round-trip path, when the count more than doubled in one frame), grow
// the buffers and render it again on the synchronous path.  The skipped frame blended nothing, so
// the target still holds what the caller put there.
int finish_frame(splat_ctx *c) {
  if (!c->have_frame) return SPLAT_OK;
  if (c->status_pending) {
    int rc = wait_done(c, c->status_ev, c->last_stream, "frame");
    if (rc) return rc;
    absorb_status(c);
  }
  for (int tries = 0; c->retry_pending && tries < 3; ++tries) {
    c->retry_pending = false;
    c->retried += 1;
    int rc = ensure_instances(c, c->last_instances + c->last_instances / 4 + 65536u);
    if (rc) return rc;
    rc = render_frame(c, c->last_params, c->last_fb, c->last_stream, nullptr, true);
    if (rc) return rc;
    rc = wait_done(c, c->status_ev, c->last_stream, "frame (repeat)");
    if (rc) return rc;
    absorb_status(c);
  }
  return SPLAT_OK;
}

int upload_common(splat_ctx *c, uint64_t n) {
  CU(cudaSetDevice(c->cfg.device));
  return alloc_scene(c, n);
}

}  // namespace

thread_local char g_create_error[256] = "";

extern "C" {

uint32_t splat_abi_version(void) { return SPLAT_ABI_VERSION; }

void splat_config_default(splat_config *cfg) {
  if (!cfg) return;
  cfg->device = 0;
  cfg->lowpass = 0.3f;        // Pipeline02 (gaussians.rs:517-518)
  cfg->y_down = 0;            // E1: pinned by the reference's own images (tests/test_reference_images.py)
  cfg->zclip_mode = 1;
  cfg->sample_offset = 0.5f;
  cfg->tile = TILE;
  cfg->max_instances = 0;
  cfg->blend_mode = SPLAT_BLEND_REFERENCE;
  cfg->near_cut = 0;
  cfg->sync_frames = 0;
  cfg->reserved = 0;
}

const char *splat_create_error(void) { return g_create_error; }

int splat_create(splat_ctx **out, const splat_config *cfg) {
  if (!out) return SPLAT_ERR_INVALID;
  *out = nullptr;
  splat_ctx *c = new (std::nothrow) splat_ctx();
  if (!c) return SPLAT_ERR_NOMEM;
  if (cfg) c->cfg = *cfg; else splat_config_default(&c->cfg);
  g_create_error[0] = 0;
  auto bail_msg = [&](int code, const char *what) {
    const cudaError_t e = cudaGetLastError();
    std::snprintf(g_create_error, sizeof g_create_error, "%s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
    splat_destroy(c);
    return code;
  };
  auto bail = [&](int code) { return bail_msg(code, code == SPLAT_ERR_NOMEM ? "allocation failed" : "CUDA runtime call failed (is this an sm_100a device?)"); };
  if (c->cfg.tile != (uint32_t)TILE) return bail_msg(SPLAT_ERR_UNSUPPORTED, "only 16-pixel tiles are built");
  if (c->cfg.blend_mode != SPLAT_BLEND_REFERENCE && c->cfg.blend_mode != SPLAT_BLEND_FLOAT)
    return bail_msg(SPLAT_ERR_UNSUPPORTED, "blend_mode must be SPLAT_BLEND_REFERENCE or SPLAT_BLEND_FLOAT");
  if (c->cfg.near_cut < -1 || c->cfg.near_cut > 1024) return bail_msg(SPLAT_ERR_INVALID, "near_cut must be -1, 0 or 1..1024");
  if (c->cfg.near_cut != 0 && c->cfg.blend_mode != SPLAT_BLEND_REFERENCE)
    return bail_msg(SPLAT_ERR_UNSUPPORTED, "the near cut belongs to the reference blend (the float blend terminates early by itself)");
  // 0 = off (default: see DESIGN.md, known issue), -1 = automatic, 1..1024 = fixed fraction
  c->cut_frac = c->cfg.near_cut == 0 ? 1024u : (c->cfg.near_cut < 0 ? NEAR_CUT_DEFAULT : (uint32_t)c->cfg.near_cut);
  if (!(c->cfg.lowpass >= 0.0f) || !std::isfinite(c->cfg.sample_offset)) return bail_msg(SPLAT_ERR_INVALID, "lowpass must be >= 0 and sample_offset finite");
  if (cudaSetDevice(c->cfg.device) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  for (int i = 0; i < EV_COUNT_; ++i)
    if (cudaEventCreate(&c->ev[i]) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (cudaEventCreateWithFlags(&c->status_ev, cudaEventDisableTiming) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (cudaEventCreateWithFlags(&c->h2d_done, cudaEventDisableTiming) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (cudaFuncSetAttribute(blend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BL_SMEM_BYTES) != cudaSuccess)
    return bail(SPLAT_ERR_CUDA);
  if (cudaFuncSetAttribute(far_cover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAR_SMEM_MAX) != cudaSuccess)
    return bail(SPLAT_ERR_CUDA);
  if (cudaFuncSetAttribute(far_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAR_SMEM_MAX) != cudaSuccess)
    return bail(SPLAT_ERR_CUDA);
  if (dev_alloc(&c->d_status, 1) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  if (cudaMemset(c->d_status, 0, sizeof(FrameStatus)) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (dev_alloc(&c->n_units, 1) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  if (dev_alloc(&c->tot, 256) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  if (cudaMallocHost(reinterpret_cast<void **>(&c->h_status), sizeof(FrameStatus)) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  std::memset(c->h_status, 0, sizeof(FrameStatus));
  if (cudaHostAlloc(reinterpret_cast<void **>(&c->h_wd), WD_WORDS * sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  std::memset(c->h_wd, 0, WD_WORDS * sizeof(uint32_t));
  if (cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->d_wd), c->h_wd, 0) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (const char *e = std::getenv("SPLAT_DEBUG_SYNC")) c->debug_sync = e[0] == '1';
  if (const char *e = std::getenv("SPLAT_WAIT_LIMIT_S")) { const double v = std::atof(e); if (v > 0.0) c->wait_limit_s = v; }
  *out = c;
  return SPLAT_OK;
}

void splat_destroy(splat_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_scene(c);
  dev_free(c->hist); dev_free(c->tot); dev_free(c->partial); dev_free(c->ranges); dev_free(c->units); dev_free(c->far_cnt); dev_free(c->far_diff); dev_free(c->tile_failed); dev_free(c->n_units); dev_free(c->d_status); dev_free(c->d_fb); dev_free(c->d_tap);
  for (int k = 0; k < 2; ++k) { dev_free(c->ikeys[k]); dev_free(c->ivals[k]); }
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_wd) cudaFreeHost(c->h_wd);
  for (int i = 0; i < EV_COUNT_; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->status_ev) cudaEventDestroy(c->status_ev);
  if (c->h2d_done) cudaEventDestroy(c->h2d_done);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
}

const char *splat_last_error(const splat_ctx *c) { return c ? c->err.c_str() : "null context"; }

int splat_upload_soa(splat_ctx *c, const float *pos4, const float *scale3, const float *opacity,
                     const float *rot_xyzw, const float *sh48, uint64_t n) {
  if (!c) return SPLAT_ERR_INVALID;
  if (!pos4 || !scale3 || !opacity || !rot_xyzw || !sh48) return fail(c, SPLAT_ERR_INVALID, "null scene array");
  int rc = upload_common(c, n);
  if (rc) return rc;
  float *raw = nullptr;   // staging: pos4 | rot | scale3 | opacity | sh48
  const size_t total = (size_t)n * (4 + 4 + 3 + 1 + 48);
  CU(dev_alloc(&raw, total));
  float *d_pos = raw, *d_rot = raw + 4 * n, *d_scale = raw + 8 * n, *d_op = raw + 11 * n, *d_sh = raw + 12 * n;
  cudaError_t e = cudaMemcpyAsync(d_pos, pos4, n * 16, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rot, rot_xyzw, n * 16, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_scale, scale3, n * 12, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_op, opacity, n * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_sh, sh48, n * 192, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    pack_scene_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(reinterpret_cast<const float4 *>(d_pos), d_scale, d_op,
                                                           reinterpret_cast<const float4 *>(d_rot), d_sh, c->scene, (uint32_t)n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(raw);
  if (e != cudaSuccess) { free_scene(c); return fail(c, SPLAT_ERR_CUDA, "scene upload", e); }
  return SPLAT_OK;
}

int splat_upload_aos(splat_ctx *c, const float *g59, uint64_t n) {
  if (!c) return SPLAT_ERR_INVALID;
  if (!g59) return fail(c, SPLAT_ERR_INVALID, "null scene array");
  int rc = upload_common(c, n);
  if (rc) return rc;
  float *raw = nullptr;
  CU(dev_alloc(&raw, (size_t)n * 59));
  cudaError_t e = cudaMemcpyAsync(raw, g59, n * 59 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    pack_scene_aos_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(raw, c->scene, (uint32_t)n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(raw);
  if (e != cudaSuccess) { free_scene(c); return fail(c, SPLAT_ERR_CUDA, "scene upload", e); }
  return SPLAT_OK;
}

// a frame skipped on the device cannot be repeated behind the caller's back when the target is the
// caller's device buffer (it may already have been consumed): report it, once, on the next call
static int report_skipped(splat_ctx *c) {
  poll_status(c);
  if (!c->retry_pending) return SPLAT_OK;
  c->retry_pending = false;
  int rc = ensure_instances(c, c->last_instances + c->last_instances / 4 + 65536u);
  if (rc) return rc;
  return fail(c, SPLAT_ERR_RETRY, "the previous splat_render_device frame was not rendered (its tile instances did not fit the "
                                  "buffers, which have now been grown; its target is untouched): render it again");
}

int splat_render_device(splat_ctx *c, const splat_camera *cam, void *fb_rows_dev, uint32_t W, uint32_t H,
                        uint32_t row0, uint32_t row1, void *stream) {
  if (!c) return SPLAT_ERR_INVALID;
  if (!c->n) return fail(c, SPLAT_ERR_STATE, "no scene uploaded");
  if (!fb_rows_dev) return fail(c, SPLAT_ERR_INVALID, "framebuffer is null");
  FrameParams P;
  int rc = make_params(c, cam, W, H, row0, row1, &P);
  if (rc) return rc;
  CU(cudaSetDevice(c->cfg.device));
  rc = report_skipped(c);
  if (rc) return rc;
  c->host_copy = false;
  return render_frame(c, P, static_cast<uint32_t *>(fb_rows_dev), stream ? static_cast<cudaStream_t>(stream) : c->stream, nullptr);
}

// host-buffer frames: the call waits for the frame anyway, so a skipped frame is simply repeated
static int host_frame(splat_ctx *c, const FrameParams &P, uint32_t *host_fb, size_t px, cudaEvent_t wait_ev) {
  int rc = render_frame(c, P, c->d_fb, c->stream, wait_ev);
  if (rc) { cudaStreamSynchronize(c->copy_stream); return rc; }
  for (int attempt = 0;; ++attempt) {
    CU(cudaEventRecord(c->ev[EV_D2H0], c->stream));
    CU(cudaMemcpyAsync(host_fb, c->d_fb, px * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev[EV_D2H1], c->stream));
    rc = wait_done(c, nullptr, c->stream, "frame (kernels + framebuffer download)");
    if (rc) return rc;
    if (c->status_pending) absorb_status(c);
    if (!c->retry_pending || attempt >= 2) break;
    rc = finish_frame(c);          // grows the buffers, renders the frame again (d_fb was left untouched)
    if (rc) return rc;
  }
  c->host_copy = true;
  return SPLAT_OK;
}

int splat_render_rows(splat_ctx *c, const splat_camera *cam, uint32_t *fb_rows, uint32_t W, uint32_t H,
                      uint32_t row0, uint32_t row1) {
  if (!c) return SPLAT_ERR_INVALID;
  if (!c->n) return fail(c, SPLAT_ERR_STATE, "no scene uploaded");
  if (!fb_rows) return fail(c, SPLAT_ERR_INVALID, "framebuffer is null");
  FrameParams P;
  int rc = make_params(c, cam, W, H, row0, row1, &P);
  if (rc) return rc;
  CU(cudaSetDevice(c->cfg.device));
  rc = report_skipped(c);
  if (rc) return rc;
  const size_t px = (size_t)(row1 - row0) * W;
  if (px > c->fb_cap) {
    CU(cudaStreamSynchronize(c->stream));
    dev_free(c->d_fb);
    CU(dev_alloc(&c->d_fb, px));
    c->fb_cap = px;
  }
  // upload on the copy stream: overlaps project/sort/binning, only blend waits for it
  CU(cudaEventRecord(c->ev[EV_H2D0], c->copy_stream));
  CU(cudaMemcpyAsync(c->d_fb, fb_rows, px * 4, cudaMemcpyHostToDevice, c->copy_stream));
  CU(cudaEventRecord(c->ev[EV_H2D1], c->copy_stream));
  CU(cudaEventRecord(c->h2d_done, c->copy_stream));
  return host_frame(c, P, fb_rows, px, c->h2d_done);
}

int splat_render(splat_ctx *c, const splat_camera *cam, uint32_t *fb, uint32_t W, uint32_t H) {
  return splat_render_rows(c, cam, fb, W, H, 0, H);
}

int splat_render_cleared(splat_ctx *c, const splat_camera *cam, uint32_t *fb_out, uint32_t W, uint32_t H,
                         uint32_t clear) {
  if (!c) return SPLAT_ERR_INVALID;
  if (!c->n) return fail(c, SPLAT_ERR_STATE, "no scene uploaded");
  if (!fb_out) return fail(c, SPLAT_ERR_INVALID, "framebuffer is null");
  FrameParams P;
  int rc = make_params(c, cam, W, H, 0, H, &P);
  if (rc) return rc;
  CU(cudaSetDevice(c->cfg.device));
  rc = report_skipped(c);
  if (rc) return rc;
  const size_t px = (size_t)W * H;
  if (px > c->fb_cap) {
    CU(cudaStreamSynchronize(c->stream));
    dev_free(c->d_fb);
    CU(dev_alloc(&c->d_fb, px));
    c->fb_cap = px;
  }
  CU(cudaEventRecord(c->ev[EV_H2D0], c->stream));
  if (clear == 0u || ((clear & 0xFFu) * 0x01010101u) == clear) {
    CU(cudaMemsetAsync(c->d_fb, (int)(clear & 0xFFu), px * 4, c->stream));
  } else {
    fill_u32_kernel<<<cdiv(px, 1024), 256, 0, c->stream>>>(c->d_fb, clear, px);
  }
  CU(cudaEventRecord(c->ev[EV_H2D1], c->stream));
  return host_frame(c, P, fb_out, px, nullptr);
}

int splat_debug_render_float(splat_ctx *c, const splat_camera *cam, uint32_t *fb_inout, uint32_t W, uint32_t H, float *rgba) {
  if (!c || !rgba) return SPLAT_ERR_INVALID;
  if (c->cfg.blend_mode != SPLAT_BLEND_FLOAT) return fail(c, SPLAT_ERR_STATE, "context was not created with SPLAT_BLEND_FLOAT");
  CU(cudaSetDevice(c->cfg.device));
  const size_t px = (size_t)W * H;
  if (px > c->tap_cap) {
    CU(cudaDeviceSynchronize());
    dev_free(c->d_tap);
    CU(dev_alloc(&c->d_tap, px));
    c->tap_cap = px;
  }
  CU(cudaMemset(c->d_tap, 0xFF, px * sizeof(float4)));   // NaN pattern = "pixel not touched"
  c->want_tap = true;
  const int rc = splat_render(c, cam, fb_inout, W, H);
  c->want_tap = false;
  if (rc) return rc;
  CU(cudaMemcpy(rgba, c->d_tap, px * sizeof(float4), cudaMemcpyDeviceToHost));
  return SPLAT_OK;
}

int splat_get_timings(splat_ctx *c, splat_timings *t) {
  if (!c || !t) return SPLAT_ERR_INVALID;
  if (!c->have_frame) return fail(c, SPLAT_ERR_STATE, "no frame rendered yet");
  CU(cudaSetDevice(c->cfg.device));
  if (c->status_pending) {
    int rcw = wait_done(c, c->status_ev, c->last_stream, "frame");
    if (rcw) return rcw;
    absorb_status(c);
  } else {
    int rcw = wait_done(c, c->ev[EV_BLEND], c->last_stream, "frame");
    if (rcw) return rcw;
  }
  { int rcs = report_skipped(c); if (rcs) return rcs; }
  std::memset(t, 0, sizeof(*t));
  auto el = [&](int a, int b) { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]); return ms; };
  t->project_ms = el(EV_START, EV_PROJECT);
  t->sort_ms = el(EV_PROJECT, EV_DSORT) + el(EV_EMIT, EV_TSORT);
  t->bin_ms = el(EV_DSORT, EV_COUNT) + el(EV_COUNT, EV_EMIT) + el(EV_TSORT, EV_RANGES);
  t->blend_ms = el(EV_RANGES, EV_BLEND);
  t->total_ms = el(EV_START, EV_BLEND);
  if (c->host_copy) {
    CU(cudaEventSynchronize(c->ev[EV_D2H1]));
    t->h2d_ms = el(EV_H2D0, EV_H2D1);
    t->d2h_ms = el(EV_D2H0, EV_D2H1);
  }
  t->frames_retried = c->retried;
  t->n_gaussians = c->n;
  t->n_visible = c->last_visible;
  t->n_instances = c->last_instances;
  t->n_tiles = c->last_tiles;
  t->kernel_launches = c->launches;
  t->near_cut_rank = c->last_cut;
  t->near_cut_failed = c->last_failed;
  t->near_cut_instances = c->last_cut_instances;
  t->frames_skipped = c->frames_skipped;
  return SPLAT_OK;
}

int splat_get_tile_loads(splat_ctx *c, uint32_t *per_tile, uint64_t cap, uint64_t *n_tiles) {
  if (!c || !per_tile || !n_tiles) return SPLAT_ERR_INVALID;
  if (!c->have_frame) return fail(c, SPLAT_ERR_STATE, "no frame rendered yet");
  if (!c->loads_valid) return fail(c, SPLAT_ERR_STATE, "the last frame's second pass binned only part of the screen (near cut): no complete tile lists");
  CU(cudaSetDevice(c->cfg.device));
  { int rcw = wait_done(c, c->ev[EV_BLEND], c->last_stream, "frame"); if (rcw) return rcw; }
  if (c->status_pending && cudaEventQuery(c->status_ev) == cudaSuccess) absorb_status(c);
  const uint64_t T = c->last_tiles, m = std::min<uint64_t>(cap, T);
  *n_tiles = T;
  if (c->last_instances == 0) { std::memset(per_tile, 0, m * sizeof(uint32_t)); return SPLAT_OK; }
  uint2 *h = static_cast<uint2 *>(std::malloc(std::max<uint64_t>(m, 1) * sizeof(uint2)));
  if (!h) return fail(c, SPLAT_ERR_NOMEM, "host allocation");
  cudaError_t e = cudaMemcpy(h, c->ranges, m * sizeof(uint2), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) for (uint64_t t = 0; t < m; ++t) per_tile[t] = h[t].y - h[t].x;
  std::free(h);
  if (e != cudaSuccess) return fail(c, SPLAT_ERR_CUDA, "tile loads", e);
  return SPLAT_OK;
}

int splat_debug_project(splat_ctx *c, const splat_camera *cam, uint32_t W, uint32_t H, float *records12,
                        uint32_t *depth_keys, uint32_t *tile_rects4) {
  if (!c) return SPLAT_ERR_INVALID;
  if (!c->n) return fail(c, SPLAT_ERR_STATE, "no scene uploaded");
  FrameParams P;
  int rc = make_params(c, cam, W, H, 0, H, &P);
  if (rc) return rc;
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaMemsetAsync(c->recs, 0, (size_t)c->n * sizeof(Rec), c->stream));
  project_kernel<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->scene, P, c->recs, c->keys[0], c->vals[0], c->rects, c->tcnt, c->block_kept);
  CU(cudaGetLastError());
  if (records12) CU(cudaMemcpyAsync(records12, c->recs, (size_t)c->n * sizeof(Rec), cudaMemcpyDeviceToHost, c->stream));
  if (depth_keys) CU(cudaMemcpyAsync(depth_keys, c->keys[0], (size_t)c->n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (tile_rects4) CU(cudaMemcpyAsync(tile_rects4, c->rects, (size_t)c->n * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->have_frame = false;
  return SPLAT_OK;
}

int splat_debug_read_order(splat_ctx *c, uint32_t *order, uint64_t cap, uint64_t *n_visible) {
  if (!c || !order || !n_visible) return SPLAT_ERR_INVALID;
  if (!c->have_frame) return fail(c, SPLAT_ERR_STATE, "no frame rendered yet");
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaDeviceSynchronize());
  const uint64_t m = std::min<uint64_t>(cap, c->last_visible);
  CU(cudaMemcpy(order, c->vals[c->order_buf], m * 4, cudaMemcpyDeviceToHost));
  *n_visible = c->last_visible;
  return SPLAT_OK;
}

int splat_debug_sort_pairs(splat_ctx *c, uint32_t *keys, uint32_t *vals, uint64_t n, int bits) {
  // standalone exercise of the radix sort (tests): sorts host arrays in place
  if (!c || !keys || !vals || n == 0 || n >= 0xFFFFFFFFull || bits < 1 || bits > 32) return SPLAT_ERR_INVALID;
  CU(cudaSetDevice(c->cfg.device));
  uint32_t *k[2] = {nullptr, nullptr}, *v[2] = {nullptr, nullptr};
  int rc = SPLAT_OK;
  const uint32_t saved_n = c->n;
  for (int i = 0; i < 2; ++i) {
    if (dev_alloc(&k[i], n) != cudaSuccess || dev_alloc(&v[i], n) != cudaSuccess) rc = SPLAT_ERR_NOMEM;
  }
  if (!rc) rc = ensure_scratch(c, n);
  if (!rc) {
    cudaMemcpy(k[0], keys, n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(v[0], vals, n * 4, cudaMemcpyHostToDevice);
    const int cur = radix_sort(c, c->stream, k, v, n, bits, 0, nullptr, (uint32_t)n);
    if (cur < 0) rc = cur;
    else {
      cudaStreamSynchronize(c->stream);
      cudaMemcpy(keys, k[cur], n * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(vals, v[cur], n * 4, cudaMemcpyDeviceToHost);
      if (cudaGetLastError() != cudaSuccess) rc = fail(c, SPLAT_ERR_CUDA, "debug sort");
    }
  }
  for (int i = 0; i < 2; ++i) { dev_free(k[i]); dev_free(v[i]); }
  c->n = saved_n;
  return rc;
}

int splat_debug_blend_stats(splat_ctx *c, uint64_t *out8, int reset) {
  if (!c || !out8) return SPLAT_ERR_INVALID;
#ifdef SPLAT_STATS
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out8, g_blend_stats, 8 * sizeof(uint64_t)));
  if (reset) {
    const uint64_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyToSymbol(g_blend_stats, z, sizeof(z)));
  }
  return SPLAT_OK;
#else
  (void)reset;
  return fail(c, SPLAT_ERR_UNSUPPORTED, "library built without -DSPLAT_STATS");
#endif
}

int splat_pin_host(void *p, uint64_t bytes) {
  return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? SPLAT_OK : SPLAT_ERR_CUDA;
}
int splat_unpin_host(void *p) { return cudaHostUnregister(p) == cudaSuccess ? SPLAT_OK : SPLAT_ERR_CUDA; }

}  // extern "C"
