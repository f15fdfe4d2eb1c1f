// splat_api.cu -- the C ABI of include/splat.h: context, scene upload, frame orchestration.
//
// Frame = render_to_buffer (pipelines.rs:66-86 / :260-280):
//   K1 project -> K3a depth radix sort (N keys) -> [near cut: K4b far_cover / far_prefix] -> K2 tile
//   count + scan -> K2 emit -> K3b tile radix sort (I keys) -> K4 ranges + unit order -> K5 blend
//   -> [near cut: one kernel that checks on the device whether the near lists sufficed].
// Only the first frame of a target geometry (and the repeat of an abandoned frame) reads the
// tile-instance count on the host; every other frame is enqueued without any host wait: launches are
// sized from the previous frame, the kernels read the real counts from device memory, and a frame
// that outgrows its bounds is abandoned on the device and repeated (render_frame, finish_frame).
// Everything runs on one stream; the framebuffer upload runs on a second stream and is only waited
// for by the blend kernel.  Multi-GPU (group contexts, per-rank communicators) is at the end.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/splat.h"
#include "bin.cuh"
#include "blend.cuh"
#include "blend_float.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "ingest.cuh"
#include "project.cuh"
#include "sort.cuh"

using namespace splat;

namespace {
constexpr uint64_t NEAR_CUT_MIN_VISIBLE = 200000;   // smaller scenes: the cut's own kernels cost more than they save
constexpr size_t FAR_SMEM_MAX = 200 * 1024;            // difference array of far_cover_kernel (shared memory)
constexpr uint32_t NEAR_CUT_DEFAULT = 128;             // 1/8 of the Gaussians
enum { EV_START = 0, EV_PROJECT, EV_DSORT, EV_COUNT, EV_EMIT, EV_TSORT, EV_RANGES, EV_BLEND,
       EV_B_END,                  // near-cut frames: the second pass
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
  uint32_t *tile_open = nullptr;   // near cut: tiles that so few cut Gaussians touch that they are binned for them after all
  int *open_sat = nullptr;         // summed-area table of tile_open
  uint2 *rank_rects = nullptr;     // near cut: tile rectangle of the Gaussian at depth rank r (r below the cut)
  uint32_t *tile_failed = nullptr; // near cut: tiles that need the complete lists
  int *far_diff = nullptr;         // its 2-D difference array
  size_t far_cells_cap = 0;
  uint32_t cut_frac = 1024;        // Gaussians binned by the near-cut pass, in 1/1024 (1024 = no cut)
  uint32_t frame_cut = 0;          // rank_cut of the frame enqueued last
  uint32_t frame_cut_seen = 0;     // ... of the frame whose status was absorbed last
  uint32_t last_cut = 0;           // rank_cut of the last frame whose status the host has seen
  uint32_t last_failed = 0;        // groups / tiles that did not converge in its near-cut pass
  uint64_t last_second_instances = 0, near_cut_fallbacks = 0, last_full_instances = 0;
  uint64_t last_cut_instances = 0; // (tile, Gaussian) pairs it did not bin
  uint32_t *n_units = nullptr;
  FrameStatus *d_status = nullptr, *h_status = nullptr;
  uint32_t *h_wd = nullptr, *d_wd = nullptr;   // blend watchdog record (mapped pinned host memory, blend.cuh)
  bool debug_sync = false;                     // SPLAT_DEBUG_SYNC=1: bounded wait after every launch, names the kernel that hangs
  double wait_limit_s = 30.0;                  // SPLAT_WAIT_LIMIT_S: bound of every host wait on the device
  uint32_t *d_fb = nullptr, *d_fb_bak = nullptr; size_t fb_cap = 0;

  float4 *d_tap = nullptr; size_t tap_cap = 0; bool want_tap = false;   // float mode: un-quantised result per pixel (tests)

  // last frame
  int order_buf = 0;          // which vals[] holds the depth order
  bool have_frame = false, host_copy = false;
  bool status_pending = false;   // some frame's end-of-frame status has not been absorbed yet
  // End-of-frame statuses land in a small ring (the host may run several frames ahead of the device;
  // one event re-recorded every frame would never be seen complete while frames keep coming)
  static constexpr int RING = 4;
  FrameStatus *h_ring = nullptr;            // RING pinned copies
  cudaEvent_t ring_ev[RING] = {};
  bool ring_pending[RING] = {};
  uint32_t ring_cut[RING] = {};             // rank_cut the frame in that slot was rendered with
  uint64_t seq = 0;                         // frames enqueued
  bool retry_pending = false;    // a frame was skipped on the device (instance buffers too small) and not yet repeated
  bool loads_valid = false;      // c->ranges describes the complete tile lists of the last frame
  uint32_t skipped_seen = 0;
  uint64_t frames_skipped = 0;
  uint32_t geom[4] = {0, 0, 0, 0};   // W, H, row0, row1 of the last frame
  FrameParams last_params{};
  uint32_t *last_fb = nullptr;
  cudaStream_t last_stream = nullptr;
  uint32_t retried = 0;
  bool empty_scene = false;         // an upload of zero Gaussians: a valid scene that renders nothing (the reference's empty Vec)
  uint64_t launches = 0, last_instances = 0, last_visible = 0, last_tiles = 0, last_sort = 0;

  // ---- multi-GPU (comm.cuh).  One process per GPU: this context joined a communicator
  // (splat_comm_init_rank).  One process, several GPUs: this is a GROUP context
  // (splat_create_multi) that owns one ordinary context per device in `members`.
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  bool owns_comm = false;
  std::vector<splat_ctx *> members;
  std::vector<uint32_t> bounds;        // group: stripe rows [2 * members], tile aligned
  uint32_t bounds_geom[2] = {0, 0};    // W, H the bounds were made for
  uint32_t *d_frame = nullptr;         // member: full W x H frame on its device (gather source / target)
  uint32_t *d_frame_bak = nullptr;     // member: the rows it uploaded, should its stripe have to be repeated
  size_t frame_cap = 0;
  uint32_t frames_since_rebalance = 0;
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
    uint32_t *ki = keys[cur], *vi = vals[cur], *ko = keys[cur ^ 1], *vo = vals[cur ^ 1];
    if (nbits == 8) rs_scatter_kernel<8><<<grid, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n_ptr, n_fixed, shift, 8, c->hist, c->tot);
    else rs_scatter_kernel<0><<<grid, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n_ptr, n_fixed, shift, nbits, c->hist, c->tot);
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
  dev_free(c->rank_rects);
  dev_free(c->scene); dev_free(c->recs); dev_free(c->rects); dev_free(c->tcnt); dev_free(c->block_kept); dev_free(c->cnt); dev_free(c->offs);
  for (int k = 0; k < 2; ++k) { dev_free(c->keys[k]); dev_free(c->vals[k]); }
  c->n = 0;
  c->have_frame = false;
  c->status_pending = false;
  for (int i = 0; i < splat_ctx::RING; ++i) c->ring_pending[i] = false;
  c->retry_pending = false;
  c->loads_valid = false;
}

int alloc_scene(splat_ctx *c, uint64_t n) {
  if (n == 0 || n > 0x7FFFFFFFull) return fail(c, SPLAT_ERR_INVALID, "n must be in [1, 2^31)");
  free_scene(c);
  c->empty_scene = false;
  CU(dev_alloc(&c->scene, (size_t)SCENE_PLANES * n));
  CU(dev_alloc(&c->recs, n));
  CU(dev_alloc(&c->rects, n));
  CU(dev_alloc(&c->rank_rects, n));
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

// ---------------------------------------------------------------- near cut: the check after the near pass, on the device
// After the first (near) pass of a near-cut frame the device knows which tiles did not converge.
// Shipped build (SPLAT_CDP = 0): pass_b_setup_kernel only records that (FrameStatus::overflow /
// skipped) and the frame is repeated without the cut.  `make cdp` (SPLAT_CDP = 1, -rdc=true, measured
// 4-6% slower overall: profiles/r2p_ab_cdp_rdc.txt) instead runs a second pass from the device:
// pass_b_setup_kernel (bin.cuh) is the only thing the host enqueues for the second pass; if there
// is work it starts this chain with CUDA dynamic parallelism -- tail launches, which run in order
// after the launching grid and before the next kernel of the host's stream -- and every launch is
// sized from the counts the device already holds.  A frame whose near lists sufficed pays one
// small kernel; nothing is gated, nothing is over-launched, the host never waits.
struct PassBArgs {
  FrameStatus *st; uint32_t *tile_failed; int *sat;
  const uint32_t *sorted_keys, *order, *tcnt, *n_sorted; const uint2 *rects;
  uint32_t *cnt, *offs, *partial; uint32_t n;
  uint32_t *ikeys[2], *ivals[2]; uint32_t *hist, *tot; int tile_bits; unsigned long long cap;
  uint2 *ranges, *units; uint32_t *n_units; uint32_t T;
  const Rec *recs; uint32_t *fb_rows; uint32_t *wd; float4 *tap; int flt;
  int diagnose_only;     // SPLAT_NO_SECOND_PASS=1: decide, but launch nothing (tools/near_cut_failures.py; wrong pixels)
  FrameParams P;
};

#ifndef SPLAT_CDP
// 0 (shipped): no device-side launches.  A near-cut frame whose near lists do not suffice is abandoned like one
// that outgrew its buffers and repeated without the cut -- with OPEN tiles (bin.cuh) that is a rarity (0 of 50
// frames on the bench orbit).  1 (`make cdp`, needs -rdc=true): the chain below repairs such a frame on the
// device; -rdc costs every kernel 4-6% (profiles/r2p_ab_cdp_rdc.txt), more than the repair ever saves.
#define SPLAT_CDP 0
#endif
#if SPLAT_CDP
__global__ void pass_b_stage3_kernel(PassBArgs a, int icur) {
  const uint32_t nu = *a.n_units;
  if (nu == 0u) return;
  if (a.flt)
    blend_float_kernel<<<nu, BF_THREADS, 0, cudaStreamTailLaunch>>>(a.ranges, a.units, a.n_units, a.ivals[icur], a.recs, a.fb_rows, a.P, a.tap, a.wd);
  else
    blend_kernel<<<nu, BL_THREADS, BL_SMEM_BYTES, cudaStreamTailLaunch>>>(a.ranges, a.units, a.n_units, a.ivals[icur], a.recs, a.fb_rows, a.P,
                                                                         nullptr, a.st, a.tile_failed, a.wd);
}

__global__ void pass_b_stage2_kernel(PassBArgs a) {
  const uint32_t I = a.st->n_inst_eff;       // pairs of the failed tiles' complete lists (0: none, or they do not fit)
  if (I == 0u) return;
  const uint32_t n = a.n;
  const cudaStream_t tl = cudaStreamTailLaunch;
  emit_masked_kernel<<<(n + 255u) / 256u, 256, 0, tl>>>(a.order, a.rects, nullptr, a.cnt, a.offs, a.ikeys[0], a.ivals[0], n, a.P.tiles_x,
                                                        a.tile_failed, a.sat, &a.st->n_inst_eff);
  const uint32_t nblk = (I + RS_BLOCK - 1u) / RS_BLOCK;
  const uint32_t *n_eff = &a.st->n_inst_eff;
  int cur = 0;
  for (int shift = 0; shift < a.tile_bits; shift += 8) {
    const int nbits = min(8, a.tile_bits - shift);
    rs_hist_kernel<<<nblk, RS_THREADS, 0, tl>>>(a.ikeys[cur], n_eff, 0u, shift, (1u << nbits) - 1u, a.hist);
    rs_rowscan_kernel<<<256, RW_THREADS, 0, tl>>>(a.hist, n_eff, 0u, a.tot);
    if (nbits == 8) rs_scatter_kernel<8><<<nblk, RS_THREADS, 0, tl>>>(a.ikeys[cur], a.ivals[cur], a.ikeys[cur ^ 1], a.ivals[cur ^ 1], n_eff, 0u, shift, 8, a.hist, a.tot);
    else rs_scatter_kernel<0><<<nblk, RS_THREADS, 0, tl>>>(a.ikeys[cur], a.ivals[cur], a.ikeys[cur ^ 1], a.ivals[cur ^ 1], n_eff, 0u, shift, nbits, a.hist, a.tot);
    cur ^= 1;
  }
  fill_u32_kernel<<<(2u * a.T + 1023u) / 1024u, 256, 0, tl>>>(reinterpret_cast<uint32_t *>(a.ranges), 0u, (size_t)2u * a.T);
  tile_ranges_kernel<<<(I + 1023u) / 1024u, 256, 0, tl>>>(a.ikeys[cur], n_eff, a.ranges);
  unit_order_kernel<<<1, 1024, 0, tl>>>(a.ranges, a.T, a.units, a.n_units, &a.st->n_instances, nullptr, a.st, a.P.tiles_x, a.tile_failed, 1, a.flt);
  pass_b_stage3_kernel<<<1, 1, 0, tl>>>(a, cur);
}

// launched by the LAST thread of pass_b_setup_kernel when the second pass has work
__device__ void pass_b_launch(const PassBArgs &a) {
  const uint32_t n = a.n;
  const cudaStream_t tl = cudaStreamTailLaunch;
  tile_count_kernel<<<(n + 255u) / 256u, 256, 0, tl>>>(a.sorted_keys, a.order, a.tcnt, a.cnt, n, a.n_sorted, 0u, a.st, a.rects, a.sat, a.P.tiles_x + 1u,
                                                       nullptr, nullptr);
  const uint32_t np = max(1u, (n + SC_BLOCK - 1u) / SC_BLOCK);
  scan_reduce_kernel<<<np, SC_THREADS, 0, tl>>>(a.cnt, a.partial, n);
  scan_partials_kernel<<<1, 1024, 0, tl>>>(a.partial, np, &a.st->n_instances, a.st, a.cap, &a.st->n_inst_eff, false);
  scan_apply_kernel<<<np, SC_THREADS, 0, tl>>>(a.cnt, a.offs, a.partial, n);
  pass_b_stage2_kernel<<<1, 1, 0, tl>>>(a);
}
#endif

// Between the two passes of a near-cut frame, on the device (one CTA): did the near lists suffice?
// Saves the first pass's counters and, if some tiles did not converge, builds the summed-area table
// of the failed-tile bitmap that the second pass bins against and launches that pass.  If the
// first pass binned nothing at all, every tile is redone.
__global__ void __launch_bounds__(1024)
pass_b_setup_kernel(PassBArgs a, uint32_t tag) {
  extern __shared__ int s_sat[];
  FrameStatus *st = a.st;
  const uint32_t tiles_x = a.P.tiles_x, tiles_y = a.P.tiles_y;
  const uint32_t pitch = tiles_x + 1u, cells = pitch * (tiles_y + 1u), T = tiles_x * tiles_y;
  const bool nothing = st->n_instances == 0ull;
  const bool active = (st->n_failed != 0u || nothing) && st->overflow == 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    st->a_instances = st->n_instances;
    st->a_failed = nothing ? T : st->n_failed;
    st->a_visible = st->n_visible;
    st->a_tag = tag;
    st->b_active = active ? 1u : 0u;
    st->n_instances = 0ull;
    st->n_visible = 0u;
    st->n_failed = 0u;
    st->n_inst_eff = 0u;
  }
  if (!active) return;
  for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) s_sat[i] = 0;
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
    if (nothing) a.tile_failed[t] = 1u;
    s_sat[(t / tiles_x + 1u) * pitch + (t % tiles_x) + 1u] = (nothing || a.tile_failed[t]) ? 1 : 0;
  }
  __syncthreads();
  for (uint32_t y = 1u + threadIdx.x; y <= tiles_y; y += blockDim.x) {     // along x, one row per thread
    int run = 0;
    for (uint32_t x = 1; x <= tiles_x; ++x) { run += s_sat[y * pitch + x]; s_sat[y * pitch + x] = run; }
  }
  __syncthreads();
  for (uint32_t x = 1u + threadIdx.x; x <= tiles_x; x += blockDim.x) {     // along y, one column per thread
    int run = 0;
    for (uint32_t y = 1; y <= tiles_y; ++y) { run += s_sat[y * pitch + x]; s_sat[y * pitch + x] = run; }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) a.sat[i] = s_sat[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    st->a_failed_tiles = (uint32_t)s_sat[tiles_y * pitch + tiles_x];
#if SPLAT_CDP
    if (!a.diagnose_only) pass_b_launch(a);
#else
    if (!a.diagnose_only) { st->overflow = 1u; st->skipped += 1u; }
#endif
  }
}

int ilog2_ceil(uint32_t v) {
  int b = 0;
  while ((1ull << b) < v) ++b;
  return std::max(b, 1);
}

// What the host learnt from a finished frame's status block (copied to pinned memory behind the
// blend kernel).  Called wherever the host has waited for the frame anyway, or finds it finished.
void absorb_slot(splat_ctx *c, int slot) {
  const FrameStatus &fs = c->h_ring[slot];
  c->frame_cut_seen = c->ring_cut[slot];
  c->ring_pending[slot] = false;
  const bool cut = fs.a_tag != 0u;               // a near-cut frame: the first pass's counters were saved on the device
  c->last_instances = cut ? fs.a_instances : fs.n_instances;
  c->last_visible = cut ? fs.a_visible : fs.n_visible;
  c->last_sort = fs.n_sort;
  c->last_cut = cut ? c->frame_cut_seen : 0u;
  c->last_failed = cut ? fs.a_failed : 0u;
  c->last_cut_instances = cut ? fs.n_cut : 0ull;
  c->last_second_instances = cut ? fs.n_instances : 0ull;
  if (cut && fs.b_active) c->near_cut_fallbacks += 1;
  c->last_full_instances = cut ? fs.a_instances + fs.n_cut : fs.n_instances;   // pairs of the complete lists
  // automatic mode: the near lists did not suffice (a repeated frame costs more than the cut saves; with the
  // device-launched second pass: it had to bin more than an eighth of what the cut saved, or more than half
  // of the tiles) -> keep twice as many Gaussians in the first pass from now on
  if (cut && c->cfg.near_cut < 0 && fs.a_tag == c->cut_frac &&
      (SPLAT_CDP ? ((uint64_t)fs.a_failed_tiles * 2u > c->last_tiles || fs.n_instances * 8ull > fs.n_cut) : fs.b_active != 0u))
    c->cut_frac = std::min<uint32_t>(1024u, c->cut_frac * 2u);
  if (fs.skipped != c->skipped_seen) {          // a frame (or several) wanted more pairs than the buffers hold
    c->frames_skipped += fs.skipped - c->skipped_seen;
    c->skipped_seen = fs.skipped;
    c->retry_pending = true;
  }
}
// absorb, oldest first, every status whose copy has landed (all = true: the caller has waited for the
// last frame, so all of them have)
void absorb_status(splat_ctx *c, bool all = true) {
  bool any = false;
  for (uint64_t q = c->seq >= (uint64_t)splat_ctx::RING ? c->seq - splat_ctx::RING : 0; q < c->seq; ++q) {
    const int slot = (int)(q % splat_ctx::RING);
    if (!c->ring_pending[slot]) continue;
    if (all || cudaEventQuery(c->ring_ev[slot]) == cudaSuccess) absorb_slot(c, slot);
    else any = true;
  }
  c->status_pending = any;
}
void poll_status(splat_ctx *c) {
  if (c->status_pending) absorb_status(c, false);
}
cudaEvent_t last_status_event(splat_ctx *c) { return c->ring_ev[(c->seq + splat_ctx::RING - 1) % splat_ctx::RING]; }

// capacity to grow the instance buffers to after a frame did not fit: the largest count any pass of it wanted, +50%
uint64_t grow_target(const splat_ctx *c) {
  const uint64_t w = std::max(std::max(c->last_instances, c->last_full_instances), c->last_second_instances);
  return w + w / 2 + 65536u;
}

// pairs the depth sort of a frame is launched for: all Gaussians, or (stripe frames the host has
// history for) the same bound the dense projection was launched with
uint64_t sort_bound_all(const splat_ctx *c, const FrameParams &, bool async, uint32_t n) {
  return async ? std::min<uint64_t>(n, c->last_sort + c->last_sort / 8 + 65536u) : n;
}

struct Pass {
  uint32_t rank_cut = 0;            // near cut: depth ranks below this get no instances (0 = complete lists)
  bool sync_count = true;           // host reads the instance count mid-frame (exact grids, buffers grown on the spot)
};

// Second half of a frame on `s`: bin the Gaussians into tiles, sort the instances by tile, blend.
//   sync_count = true : the host waits for the instance count (one round trip), grows the buffers if
//                       needed and sizes the launches exactly.  First frame of a target geometry and
//                       the repeat of a skipped frame.
//   sync_count = false: nothing blocks.  Launches are sized from the previous frame (+12.5%); the
//                       kernels read the count from device memory, and a frame whose pairs do not
//                       fit that bound blends NOTHING (status.overflow) and is repeated by the host.
// With rank_cut this is the first pass of a near-cut frame; enqueue_second_pass follows it.
int bin_sort_blend(splat_ctx *c, const FrameParams &P, uint32_t *fb_rows_dev, cudaStream_t s, cudaEvent_t wait_ev,
                   int cur, const uint32_t *n_sorted, const Pass &pass) {
  const uint32_t rank_cut = pass.rank_cut;
  const uint32_t n = c->n;
  const uint32_t T = P.tiles_x * P.tiles_y;
  const bool flt = c->cfg.blend_mode == SPLAT_BLEND_FLOAT;
  const size_t cells = (size_t)(P.tiles_x + 1) * (P.tiles_y + 1);   // not a function of T alone
  if (T > c->ranges_cap) {
    CU(cudaStreamSynchronize(s));
    dev_free(c->ranges);
    dev_free(c->units);
    dev_free(c->far_cnt);
    dev_free(c->tile_failed);
    dev_free(c->tile_open);
    CU(dev_alloc(&c->tile_open, T));
    CU(dev_alloc(&c->ranges, T));
    CU(dev_alloc(&c->units, (size_t)4 * T));
    CU(dev_alloc(&c->far_cnt, T));
    CU(dev_alloc(&c->tile_failed, T));
    c->ranges_cap = T;
  }
  if (cells > c->far_cells_cap) {
    CU(cudaStreamSynchronize(s));
    dev_free(c->far_diff);
    dev_free(c->open_sat);
    CU(dev_alloc(&c->far_diff, cells));
    CU(dev_alloc(&c->open_sat, cells));
    c->far_cells_cap = cells;
  }
  // n_instances, n_visible, n_failed, fail box, n_cut (the fields behind them belong to the frame)
  CU(cudaMemsetAsync(c->d_status, 0, offsetof(FrameStatus, n_sort), s));
  const uint32_t *far = nullptr;
  if (rank_cut) {
    CU(cudaMemsetAsync(c->far_diff, 0, cells * sizeof(int), s));
    CU(cudaMemsetAsync(c->tile_failed, 0, (size_t)T * sizeof(uint32_t), s));
    far_cover_kernel<<<148, FC_THREADS, cells * sizeof(int), s>>>(c->vals[cur], c->rects, n_sorted, n, rank_cut,
                                                                  P.tiles_x, P.tiles_y, c->far_diff, c->rank_rects);
    far_prefix_kernel<<<1, 1024, cells * sizeof(int), s>>>(c->far_diff, P.tiles_x, P.tiles_y, c->far_cnt, c->d_status,
                                                           c->tile_open, c->open_sat);
    c->launches += 1;
    LAUNCHED("far_cover / far_prefix");
    far = c->far_cnt;
  }
  // no-round-trip frames: launches are sized for 1.125x the last count the host has seen (+64k); a
  // frame that wants more is skipped on the device (overflow) and repeated with a round trip
  const uint64_t n_bound = std::min<uint64_t>(c->inst_cap, c->last_instances + c->last_instances / 8 + 65536u);
  tile_count_kernel<<<cdiv(n, 256), 256, 0, s>>>(c->keys[cur], c->vals[cur], c->tcnt, c->cnt, n, n_sorted, rank_cut, c->d_status,
                                                  nullptr, nullptr, P.tiles_x + 1u, rank_cut ? c->rank_rects : nullptr,
                                                  rank_cut ? c->open_sat : nullptr);
  LAUNCHED("tile_count_kernel");
  // exclusive scan cnt -> offs; grand total -> status.n_instances, checked against the bound on the
  // device (status.n_inst_eff / overflow)
  {
    const uint32_t np = std::max(1u, cdiv(n, SC_BLOCK));
    scan_reduce_kernel<<<np, SC_THREADS, 0, s>>>(c->cnt, c->partial, n);
    scan_partials_kernel<<<1, 1024, 0, s>>>(c->partial, np, &c->d_status->n_instances, c->d_status,
                                            pass.sync_count ? 0xFFFFFFFEull : (unsigned long long)n_bound, &c->d_status->n_inst_eff);
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
    c->last_full_instances = I;
    c->last_visible = c->h_status->n_visible;
    c->last_sort = c->h_status->n_sort;
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
                                                      c->ikeys[0], c->ivals[0], n, P.tiles_x, n_eff, rank_cut);
  if (rank_cut)   // the cut Gaussians' few instances in the open tiles
    emit_masked_kernel<<<cdiv(rank_cut, 256), 256, 0, s>>>(c->vals[cur], c->rects, c->rank_rects, c->cnt, c->offs, c->ikeys[0], c->ivals[0],
                                                           rank_cut, P.tiles_x, c->tile_open, c->open_sat, n_eff);
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
                                       c->tile_failed, 0, flt ? 1 : 0);   // heaviest first
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
  return SPLAT_OK;
}

// Near-cut frames: ONE kernel decides on the device whether the near lists sufficed and, if not,
// launches the second pass itself (pass_b_setup_kernel above).
int enqueue_second_pass(splat_ctx *c, const FrameParams &P, uint32_t *fb_rows_dev, cudaStream_t s, int cur, const uint32_t *n_sorted) {
  PassBArgs a;
  a.st = c->d_status; a.tile_failed = c->tile_failed; a.sat = c->far_diff;
  a.sorted_keys = c->keys[cur]; a.order = c->vals[cur]; a.tcnt = c->tcnt; a.n_sorted = n_sorted; a.rects = c->rects;
  a.cnt = c->cnt; a.offs = c->offs; a.partial = c->partial; a.n = c->n;
  for (int k = 0; k < 2; ++k) { a.ikeys[k] = c->ikeys[k]; a.ivals[k] = c->ivals[k]; }
  a.hist = c->hist; a.tot = c->tot;
  a.T = P.tiles_x * P.tiles_y;
  a.tile_bits = ilog2_ceil(a.T);
  a.cap = c->inst_cap;
  a.ranges = c->ranges; a.units = c->units; a.n_units = c->n_units;
  a.recs = c->recs; a.fb_rows = fb_rows_dev; a.wd = c->d_wd; a.tap = nullptr;
  a.flt = c->cfg.blend_mode == SPLAT_BLEND_FLOAT ? 1 : 0;
  a.diagnose_only = std::getenv("SPLAT_NO_SECOND_PASS") ? 1 : 0;
  a.P = P;
  const size_t cells = (size_t)(P.tiles_x + 1) * (P.tiles_y + 1);
  pass_b_setup_kernel<<<1, 1024, cells * sizeof(int), s>>>(a, c->cut_frac);
  LAUNCHED("pass_b_setup_kernel (+ the second pass it launches)");
  CU(cudaEventRecord(c->ev[EV_B_END], s));
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
  // Near cut: the blend reads only the nearest few hundred entries of every tile list (exact early
  // termination), so first bin + sort only the nearest cut_frac/1024 of the Gaussians -- the
  // depth ranks the previous frame makes us expect at the top -- and fall back to the complete
  // lists only if some pixel did not converge on them (then keep twice as many from now on).
  uint32_t rank_cut = 0;
  const size_t far_smem = (size_t)(P.tiles_x + 1) * (P.tiles_y + 1) * sizeof(int);
  const uint64_t min_visible = c->cfg.near_cut > 0 ? 1u : NEAR_CUT_MIN_VISIBLE;   // a fixed fraction is honoured on any scene (tests)
  // automatic mode only where it pays and is safe: tile lists of >= 2048 entries on average, so that the
  // nearest eighth still holds a few hundred entries per tile
  const bool long_lists = c->cfg.near_cut > 0 || c->last_full_instances >= 2048ull * P.tiles_x * P.tiles_y;
  if (c->cut_frac < 1024u && c->have_frame && c->last_visible >= min_visible && long_lists && far_smem <= FAR_SMEM_MAX) {
    const uint64_t keep = (c->last_visible * c->cut_frac + 1023u) / 1024u;
    if (keep < c->last_visible) rank_cut = (uint32_t)(c->last_visible - keep);
  }
  // a frame with a host round trip (first of a geometry, repeat of a skipped one) bins everything
  bool async = same_geom && !force_sync && !c->retry_pending && !c->cfg.sync_frames;
  if (!async) rank_cut = 0;
  {
    // the buffers must hold this frame's pairs: the first pass's (+12.5%), and -- near-cut frames --
    // whatever the second pass may want, up to the complete lists.  cudaFree / cudaMalloc wait for the
    // frames in flight; rare (the buffers are grown with 25% headroom)
    const uint64_t want = rank_cut ? c->last_full_instances + c->last_full_instances / 4 : c->last_instances;
    if (async && want + want / 8 + 65536u > c->inst_cap) {
      int rc = ensure_instances(c, want + want / 8 + 65536u);
      if (rc) return rc;
    }
  }
  CU(cudaEventRecord(c->ev[EV_START], s));
  CU(cudaMemsetAsync(c->d_status, 0, offsetof(FrameStatus, skipped), s));
  const uint32_t *n_sorted = nullptr;
  int cur0 = 0;                 // which keys[] / vals[] buffer holds the pairs that enter the depth sort
  bool dense_stripe = false;
  if (!P.stripe_cull) {
    project_kernel<false><<<cdiv(n, 256), 256, 0, s>>>(c->scene, P, c->recs, c->keys[0], c->vals[0], c->rects, c->tcnt, nullptr, nullptr, nullptr);
    LAUNCHED("project_kernel");
  } else if (same_geom && !force_sync && c->last_sort * 10u > (uint64_t)n * 3u) {
    // Dense stripe (the last frame kept more than 30% of the Gaussians: with 16-byte planes nearly every
    // sector would be read anyway and the pre-pass only adds work): full projection over all Gaussians,
    // those that cannot reach the stripe culled inside, then the kept pairs squeezed to the front.
    const uint32_t nb = cdiv(n, 256), np = std::max(1u, cdiv(nb, SC_BLOCK));
    project_kernel<false><<<nb, 256, 0, s>>>(c->scene, P, c->recs, c->keys[0], c->vals[0], c->rects, c->tcnt, nullptr, nullptr, c->block_kept);
    scan_reduce_kernel<<<np, SC_THREADS, 0, s>>>(c->block_kept, c->partial, nb);
    scan_partials_kernel<<<1, 1024, 0, s>>>(c->partial, np, &c->d_status->n_sort);
    scan_apply_kernel<<<np, SC_THREADS, 0, s>>>(c->block_kept, c->block_kept, c->partial, nb);
    compact_pairs_kernel<<<nb, 256, 0, s>>>(c->keys[0], c->vals[0], c->keys[1], c->vals[1], c->block_kept, n);
    n_sorted = reinterpret_cast<const uint32_t *>(&c->d_status->n_sort);
    cur0 = 1;
    dense_stripe = true;
    c->launches += 4;
    LAUNCHED("project_kernel + pair compaction (dense stripe)");
  } else {
    // Stripe (multi-GPU) frames: a 48 B/Gaussian pre-pass votes which Gaussians can reach the rows
    // of this stripe; the survivors' indices are compacted in index order and the projection runs
    // densely over them -- the O(N) work every rank repeats shrinks to the pre-pass.
    uint32_t *mask = c->cnt, *surv = c->offs;          // both idle until the binning stage
    const uint32_t nb = cdiv(n, 256), np = std::max(1u, cdiv(nb, SC_BLOCK));
    // launch bound of the dense projection: the last survivor count the host has seen, +12.5%
    const uint64_t sort_bound = async ? std::min<uint64_t>(n, c->last_sort + c->last_sort / 8 + 65536u) : n;
    stripe_cull_kernel<<<nb, 256, 0, s>>>(c->scene, P, mask, c->block_kept);
    scan_reduce_kernel<<<np, SC_THREADS, 0, s>>>(c->block_kept, c->partial, nb);
    scan_partials_kernel<<<1, 1024, 0, s>>>(c->partial, np, &c->d_status->n_sort, c->d_status, sort_bound, nullptr, true);
    scan_apply_kernel<<<np, SC_THREADS, 0, s>>>(c->block_kept, c->block_kept, c->partial, nb);
    compact_idx_kernel<<<nb, 256, 0, s>>>(mask, c->block_kept, surv, n);
    n_sorted = reinterpret_cast<const uint32_t *>(&c->d_status->n_sort);   // low word (n < 2^31)
    project_kernel<true><<<std::max(1u, cdiv(sort_bound, 256)), 256, 0, s>>>(c->scene, P, c->recs, c->keys[0], c->vals[0], c->rects, c->tcnt,
                                                                          surv, n_sorted, nullptr);
    c->launches += 5;
    LAUNCHED("stripe pre-pass + project_kernel");
  }
  CU(cudaEventRecord(c->ev[EV_PROJECT], s));
  int cur = cur0;
  // (a stripe sorts only its survivors -- a device-side count; the CTAs beyond it exit at once)
  cur = radix_sort(c, s, c->keys, c->vals, (P.stripe_cull && !dense_stripe) ? sort_bound_all(c, P, async, n) : n, 32, cur, n_sorted, n);
  if (cur < 0) return cur;
  c->order_buf = cur;
  CU(cudaEventRecord(c->ev[EV_DSORT], s));

  Pass pass;
  pass.rank_cut = rank_cut;
  pass.sync_count = !async;
  int rc = bin_sort_blend(c, P, fb_rows_dev, s, wait_ev, cur, n_sorted, pass);
  if (rc) return rc;
  c->loads_valid = rank_cut == 0;
  if (rank_cut) {
    // The second pass decides ON THE DEVICE whether it has work (and launches itself): if some tiles
    // did not converge on the near lists, they -- and only they -- are binned, sorted and blended
    // again with ALL Gaussians.  Groups that already converged wrote final values and, blended again,
    // converge to them again without reading the framebuffer; groups that did not converge left
    // their pixels untouched.
    rc = enqueue_second_pass(c, P, fb_rows_dev, s, cur, n_sorted);
    if (rc) return rc;
  }
  c->frame_cut = rank_cut;
  {
    const int slot = (int)(c->seq % splat_ctx::RING);
    if (c->ring_pending[slot] && cudaEventQuery(c->ring_ev[slot]) == cudaSuccess) absorb_slot(c, slot);   // else: dropped
    CU(cudaMemcpyAsync(&c->h_ring[slot], c->d_status, sizeof(FrameStatus), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(c->ring_ev[slot], s));
    c->ring_pending[slot] = true;
    c->ring_cut[slot] = rank_cut;
    c->seq += 1;
    c->status_pending = true;
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
    int rc = wait_done(c, last_status_event(c), c->last_stream, "frame");
    if (rc) return rc;
    absorb_status(c);
  }
  for (int tries = 0; c->retry_pending && tries < 3; ++tries) {
    c->retry_pending = false;
    int rc = ensure_instances(c, grow_target(c));
    if (rc) return rc;
    const uint32_t retried = c->retried;
    rc = render_frame(c, c->last_params, c->last_fb, c->last_stream, nullptr, true);
    if (rc) return rc;
    c->retried += retried + 1;         // render_frame starts its count at 0: this frame is a repeat
    rc = wait_done(c, last_status_event(c), c->last_stream, "frame (repeat)");
    if (rc) return rc;
    absorb_status(c);
  }
  return SPLAT_OK;
}

// An empty scene (an empty Vec<Gaussian> / a PLY with no vertices) is valid input for render_to_buffer: nothing is drawn.
// No device memory, no kernel: the render entry points return before touching the device.
int upload_empty(splat_ctx *c) {
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaStreamSynchronize(c->stream));
  free_scene(c);
  c->n = 0;
  c->empty_scene = true;
  for (splat_ctx *m : c->members) {
    int rc = upload_empty(m);
    if (rc) { c->err = m->err; return rc; }
  }
  return SPLAT_OK;
}

int upload_common(splat_ctx *c, uint64_t n) {
  CU(cudaSetDevice(c->cfg.device));
  return alloc_scene(c, n);
}


// ---------------------------------------------------------------- multi-GPU: group contexts
#define NC(expr)                                                                              \
  do {                                                                                        \
    ncclResult_t r__ = (expr);                                                                \
    if (r__ != ncclSuccess) {                                                                 \
      c->err = std::string(#expr) + ": " + (nccl_api() ? nccl_api()->GetErrorString(r__) : "NCCL"); \
      return SPLAT_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

// equal numbers of tile rows (remainder to the first members)
void equal_bounds(std::vector<uint32_t> &b, uint32_t H, int parts) {
  const uint32_t tr = cdiv(H, TILE);
  b.assign((size_t)2 * parts, 0u);
  uint32_t r = 0;
  for (int k = 0; k < parts; ++k) {
    const uint32_t rows = tr / parts + ((uint32_t)k < tr % parts ? 1u : 0u);
    b[2 * k] = std::min(r * TILE, H);
    r += rows;
    b[2 * k + 1] = std::min(r * TILE, H);
  }
}

// New stripe boundaries from the members' measured frame times: the cost of a tile row is taken as
// uniform inside the stripe that rendered it, and the rows are re-cut so that the heaviest stripe is
// as light as possible (greedy fill against a bisected bottleneck).  SURVEY H6.
void rebalance_bounds(std::vector<uint32_t> &b, const std::vector<float> &ms, uint32_t H) {
  const int parts = (int)ms.size();
  const uint32_t tr = cdiv(H, TILE);
  std::vector<double> w(tr, 0.0);
  for (int k = 0; k < parts; ++k) {
    const uint32_t t0 = b[2 * k] / TILE, t1 = cdiv(b[2 * k + 1], TILE);
    for (uint32_t t = t0; t < t1 && t < tr; ++t) w[t] = std::max(1e-4, (double)ms[k]) / std::max(1u, t1 - t0);
  }
  double lo = 0.0, hi = 0.0;
  for (double v : w) { lo = std::max(lo, v); hi += v; }
  auto cuts_for = [&](double cap, std::vector<uint32_t> *out) {
    uint32_t t = 0;
    if (out) out->assign((size_t)2 * parts, 0u);
    for (int k = 0; k < parts; ++k) {
      const uint32_t start = t;
      double acc = 0.0;
      while (t < tr && acc + w[t] <= cap) acc += w[t++];
      if (out) { (*out)[2 * k] = std::min(start * TILE, H); (*out)[2 * k + 1] = std::min(t * TILE, H); }
    }
    return t >= tr;
  };
  for (int it = 0; it < 50; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (cuts_for(mid, nullptr)) hi = mid; else lo = mid;
  }
  std::vector<uint32_t> nb;
  if (cuts_for(hi * (1.0 + 1e-9) + 1e-12, &nb)) { nb[(size_t)2 * parts - 1] = H; b = nb; }
}

int ensure_frame(splat_ctx *m, size_t px) {
  splat_ctx *c = m;
  if (px <= m->frame_cap) return SPLAT_OK;
  CU(cudaDeviceSynchronize());
  dev_free(m->d_frame);
  dev_free(m->d_frame_bak);
  CU(dev_alloc(&m->d_frame, px));
  CU(dev_alloc(&m->d_frame_bak, px));
  m->frame_cap = px;
  return SPLAT_OK;
}

// C0 for a group: member 0 holds the packed device scene; the others receive it (ncclBroadcast).
int group_broadcast_scene(splat_ctx *g) {
  splat_ctx *c = g;
  NcclApi *N = nccl_api();
  if (!N) return fail(g, SPLAT_ERR_UNSUPPORTED, "NCCL is not available");
  splat_ctx *root = g->members[0];
  const uint64_t n = root->n;
  for (size_t i = 1; i < g->members.size(); ++i) {
    splat_ctx *m = g->members[i];
    CU(cudaSetDevice(m->cfg.device));
    int rc = alloc_scene(m, n);
    if (rc) { g->err = m->err; return rc; }
  }
  NC(N->GroupStart());
  for (splat_ctx *m : g->members) {
    CU(cudaSetDevice(m->cfg.device));
    NC(N->Broadcast(root->scene, m->scene, (size_t)SCENE_PLANES * n * 4u, ncclFloat, 0, m->comm, m->stream));
  }
  NC(N->GroupEnd());
  for (splat_ctx *m : g->members) {
    CU(cudaSetDevice(m->cfg.device));
    CU(cudaStreamSynchronize(m->stream));
  }
  g->n = (uint32_t)n;
  return SPLAT_OK;
}

// One frame on every member: each renders its stripe of rows into its own W x H device frame, the
// stripes are gathered into member 0's frame with one grouped send/recv, member 0 returns the frame.
// clear_value < 0: fb_inout is blended onto (each member uploads its own rows); else cleared on the device.
int group_render(splat_ctx *g, const splat_camera *cam, uint32_t *fb, uint32_t W, uint32_t H, long long clear_value) {
  splat_ctx *c = g;
  NcclApi *N = nccl_api();
  if (!N) return fail(g, SPLAT_ERR_UNSUPPORTED, "NCCL is not available");
  const int G = (int)g->members.size();
  if (!g->n) return fail(g, SPLAT_ERR_STATE, "no scene uploaded");
  if (!fb) return fail(g, SPLAT_ERR_INVALID, "framebuffer is null");
  if (g->bounds.size() != (size_t)2 * G || g->bounds_geom[0] != W || g->bounds_geom[1] != H) {
    equal_bounds(g->bounds, H, G);
    g->bounds_geom[0] = W; g->bounds_geom[1] = H;
    g->frames_since_rebalance = 0;
  }
  const size_t px = (size_t)W * H;
  std::vector<char> todo((size_t)G, 1);     // members whose stripe still has to be rendered
  for (int attempt = 0; attempt < 3; ++attempt) {
    // enqueue: upload / clear, kernels (a repeated stripe starts from the caller's pixels again)
    for (int k = 0; k < G; ++k) {
      splat_ctx *m = g->members[k];
      const uint32_t r0 = g->bounds[2 * k], r1 = g->bounds[2 * k + 1];
      CU(cudaSetDevice(m->cfg.device));
      int rc = ensure_frame(m, px);
      if (rc) { g->err = m->err; return rc; }
      if (r1 <= r0 || !todo[k]) continue;
      FrameParams P;
      rc = make_params(m, cam, W, H, r0, r1, &P);
      if (rc) { g->err = m->err; return rc; }
      uint32_t *rows = m->d_frame + (size_t)r0 * W;
      const size_t bytes = (size_t)(r1 - r0) * W * 4u;
      cudaEvent_t wait_ev = nullptr;
      CU(cudaEventRecord(m->ev[EV_H2D0], m->copy_stream));
      if (clear_value < 0 && attempt > 0) {
        // the download overwrote the caller's pixels: the member kept a copy of what it uploaded
        CU(cudaMemcpyAsync(rows, m->d_frame_bak + (size_t)r0 * W, bytes, cudaMemcpyDeviceToDevice, m->copy_stream));
      } else if (clear_value < 0) {
        CU(cudaMemcpyAsync(rows, fb + (size_t)r0 * W, bytes, cudaMemcpyHostToDevice, m->copy_stream));
        CU(cudaMemcpyAsync(m->d_frame_bak + (size_t)r0 * W, rows, bytes, cudaMemcpyDeviceToDevice, m->copy_stream));
      } else if (clear_value == 0 || (((uint32_t)clear_value & 0xFFu) * 0x01010101u) == (uint32_t)clear_value) {
        CU(cudaMemsetAsync(rows, (int)((uint32_t)clear_value & 0xFFu), bytes, m->copy_stream));
      } else {
        fill_u32_kernel<<<cdiv(bytes / 4u, 1024), 256, 0, m->copy_stream>>>(rows, (uint32_t)clear_value, bytes / 4u);
      }
      CU(cudaEventRecord(m->ev[EV_H2D1], m->copy_stream));
      CU(cudaEventRecord(m->h2d_done, m->copy_stream));
      wait_ev = m->h2d_done;
      rc = render_frame(m, P, rows, m->stream, wait_ev, attempt > 0);
      if (rc) { g->err = m->err; return rc; }
    }
    // C1: one grouped gather, enqueued on every member's stream right behind its blend kernel
    NC(N->GroupStart());
    for (int k = 0; k < G; ++k) {
      splat_ctx *m = g->members[k];
      CU(cudaSetDevice(m->cfg.device));
      NC(gather_stripes(N, m->comm, G, k, 0, m->d_frame, W, g->bounds.data(), m->stream));
    }
    NC(N->GroupEnd());
    splat_ctx *root = g->members[0];
    CU(cudaSetDevice(root->cfg.device));
    if (attempt == 0 && clear_value < 0) {
      // the members members' uploads read the caller's buffer: they must be done before the download writes it
      for (int k = 0; k < G; ++k) {
        CU(cudaSetDevice(g->members[k]->cfg.device));
        CU(cudaStreamSynchronize(g->members[k]->copy_stream));
      }
      CU(cudaSetDevice(root->cfg.device));
    }
    CU(cudaEventRecord(root->ev[EV_D2H0], root->stream));
    CU(cudaMemcpyAsync(fb, root->d_frame, px * 4u, cudaMemcpyDeviceToHost, root->stream));
    CU(cudaEventRecord(root->ev[EV_D2H1], root->stream));
    bool again = false;
    for (int k = 0; k < G; ++k) {
      splat_ctx *m = g->members[k];
      if (g->bounds[2 * k + 1] <= g->bounds[2 * k] && k != 0) continue;
      CU(cudaSetDevice(m->cfg.device));
      int rc = wait_done(m, nullptr, m->stream, "frame (group member)");
      if (rc) { g->err = m->err; return rc; }
      if (m->status_pending) absorb_status(m);
      todo[k] = 0;
      if (m->retry_pending) {          // this member's stripe was skipped on the device: grow, render it again
        m->retry_pending = false;
        rc = ensure_instances(m, grow_target(m));
        if (rc) { g->err = m->err; return rc; }
        todo[k] = 1;
        again = true;
      }
    }
    root->host_copy = true;
    if (!again) break;
  }
  g->have_frame = true;
  // re-cut the stripes from the measured times when they drift apart (at most every 8th frame: a
  // new cut costs each member one frame with a host round trip)
  g->frames_since_rebalance += 1;
  if (G > 1 && g->cfg.reserved == 0 && (g->frames_since_rebalance == 2 || g->frames_since_rebalance % 8 == 0)) {
    std::vector<float> ms((size_t)G, 0.f);
    float mx = 0.f, mn = 1e30f;
    for (int k = 0; k < G; ++k) {
      splat_ctx *m = g->members[k];
      if (g->bounds[2 * k + 1] <= g->bounds[2 * k]) { ms[k] = 0.f; mn = 0.f; continue; }
      CU(cudaSetDevice(m->cfg.device));
      float v = 0.f;
      cudaEventElapsedTime(&v, m->ev[EV_START], m->ev[EV_BLEND]);
      ms[k] = v; mx = std::max(mx, v); mn = std::min(mn, v);
    }
    if (mx > 1.15f * mn + 0.02f) rebalance_bounds(g->bounds, ms, H);
  }
  return SPLAT_OK;
}

int group_timings(splat_ctx *g, splat_timings *t) {
  std::memset(t, 0, sizeof(*t));
  for (size_t k = 0; k < g->members.size(); ++k) {
    splat_ctx *m = g->members[k];
    if (!m->have_frame) continue;
    splat_timings tm;
    int rc = splat_get_timings(m, &tm);
    if (rc) { g->err = m->err; return rc; }
    t->project_ms = std::max(t->project_ms, tm.project_ms); t->sort_ms = std::max(t->sort_ms, tm.sort_ms);
    t->bin_ms = std::max(t->bin_ms, tm.bin_ms); t->blend_ms = std::max(t->blend_ms, tm.blend_ms);
    t->total_ms = std::max(t->total_ms, tm.total_ms);
    t->h2d_ms = std::max(t->h2d_ms, tm.h2d_ms);
    if (k == 0) t->d2h_ms = tm.d2h_ms;
    t->frames_retried += tm.frames_retried; t->frames_skipped += tm.frames_skipped;
    t->n_visible += tm.n_visible; t->n_instances += tm.n_instances; t->n_tiles += tm.n_tiles;
    t->kernel_launches += tm.kernel_launches;
  }
  t->n_gaussians = g->n;
  return SPLAT_OK;
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
  cfg->near_cut = -1;         // automatic (exact either way; 0 switches it off)
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
  if (c->cfg.near_cut > 0 && c->cfg.blend_mode != SPLAT_BLEND_REFERENCE)
    return bail_msg(SPLAT_ERR_UNSUPPORTED, "the near cut belongs to the reference blend (the float blend terminates early by itself)");
  if (c->cfg.blend_mode != SPLAT_BLEND_REFERENCE) c->cfg.near_cut = 0;   // "automatic" means none there
  // 0 = off, -1 = automatic (default), 1..1024 = fixed fraction
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
  if (cudaFuncSetAttribute(pass_b_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAR_SMEM_MAX) != cudaSuccess)
    return bail(SPLAT_ERR_CUDA);
  if (dev_alloc(&c->d_status, 1) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  if (cudaMemset(c->d_status, 0, sizeof(FrameStatus)) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
  if (dev_alloc(&c->n_units, 1) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  if (dev_alloc(&c->tot, 256) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  if (cudaMallocHost(reinterpret_cast<void **>(&c->h_status), sizeof(FrameStatus)) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  std::memset(c->h_status, 0, sizeof(FrameStatus));
  if (cudaMallocHost(reinterpret_cast<void **>(&c->h_ring), splat_ctx::RING * sizeof(FrameStatus)) != cudaSuccess) return bail(SPLAT_ERR_NOMEM);
  std::memset(c->h_ring, 0, splat_ctx::RING * sizeof(FrameStatus));
  for (int i = 0; i < splat_ctx::RING; ++i)
    if (cudaEventCreateWithFlags(&c->ring_ev[i], cudaEventDisableTiming) != cudaSuccess) return bail(SPLAT_ERR_CUDA);
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
  if (!c->members.empty()) {           // group context: it owns nothing on a device itself
    for (splat_ctx *m : c->members) splat_destroy(m);
    delete c;
    return;
  }
  cudaSetDevice(c->cfg.device);
  if (c->comm && c->owns_comm) {
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (NcclApi *N = nccl_api()) N->CommDestroy(c->comm);
    c->comm = nullptr;
  }
  dev_free(c->d_frame);
  dev_free(c->d_frame_bak);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_scene(c);
  dev_free(c->hist); dev_free(c->tot); dev_free(c->partial); dev_free(c->ranges); dev_free(c->units); dev_free(c->far_cnt); dev_free(c->far_diff); dev_free(c->tile_failed); dev_free(c->tile_open); dev_free(c->open_sat); dev_free(c->n_units); dev_free(c->d_status); dev_free(c->d_fb); dev_free(c->d_fb_bak); dev_free(c->d_tap);
  for (int k = 0; k < 2; ++k) { dev_free(c->ikeys[k]); dev_free(c->ivals[k]); }
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->h_ring) cudaFreeHost(c->h_ring);
  for (int i = 0; i < splat_ctx::RING; ++i) if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]);
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
  if (n == 0) return upload_empty(c);
  if (!c->members.empty()) {
    int rc = splat_upload_soa(c->members[0], pos4, scale3, opacity, rot_xyzw, sh48, n);
    if (rc) { c->err = c->members[0]->err; return rc; }
    return group_broadcast_scene(c);
  }
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

int splat_upload_ply_raw(splat_ctx *c, const void *vertex_rows, uint64_t n, uint32_t stride_floats, float *activated60) {
  if (!c) return SPLAT_ERR_INVALID;
  if (n == 0) return upload_empty(c);
  if (!vertex_rows) return fail(c, SPLAT_ERR_INVALID, "null vertex payload");
  if (stride_floats < (uint32_t)PLY_FLOATS) return fail(c, SPLAT_ERR_INVALID, "the INRIA vertex layout has 62 floats per vertex");
  if (!c->members.empty()) {
    int rc = splat_upload_ply_raw(c->members[0], vertex_rows, n, stride_floats, activated60);
    if (rc) { c->err = c->members[0]->err; return rc; }
    return group_broadcast_scene(c);
  }
  int rc = upload_common(c, n);
  if (rc) return rc;
  float *raw = nullptr, *act = nullptr, *mean3 = nullptr;
  const size_t raw_floats = (size_t)n * stride_floats, act_floats = (size_t)n * 60;
  cudaError_t e = dev_alloc(&raw, raw_floats);
  if (e == cudaSuccess) e = dev_alloc(&act, act_floats);
  if (e == cudaSuccess) e = dev_alloc(&mean3, 4);
  float *d_pos = act, *d_rot = act + 4 * n, *d_scale = act + 8 * n, *d_op = act + 11 * n, *d_sh = act + 12 * n;
  if (e == cudaSuccess) e = cudaMemcpyAsync(raw, vertex_rows, raw_floats * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    ply_mean_kernel<<<1, 1024, 0, c->stream>>>(raw, stride_floats, n, mean3);
    ply_activate_kernel<<<cdiv(n, PLY_ROWS), PLY_ROWS, 0, c->stream>>>(raw, stride_floats, (uint32_t)n, mean3, reinterpret_cast<float4 *>(d_pos),
                                                                      d_scale, d_op, reinterpret_cast<float4 *>(d_rot), d_sh);
    pack_scene_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(reinterpret_cast<const float4 *>(d_pos), d_scale, d_op,
                                                           reinterpret_cast<const float4 *>(d_rot), d_sh, c->scene, (uint32_t)n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && activated60) e = cudaMemcpyAsync(activated60, act, act_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(raw); cudaFree(act); cudaFree(mean3);
  if (e != cudaSuccess) { free_scene(c); return fail(c, e == cudaErrorMemoryAllocation ? SPLAT_ERR_NOMEM : SPLAT_ERR_CUDA, "PLY ingest", e); }
  return SPLAT_OK;
}

int splat_upload_aos(splat_ctx *c, const float *g59, uint64_t n) {
  if (!c) return SPLAT_ERR_INVALID;
  if (n == 0) return upload_empty(c);
  if (!c->members.empty()) {
    int rc = splat_upload_aos(c->members[0], g59, n);
    if (rc) { c->err = c->members[0]->err; return rc; }
    return group_broadcast_scene(c);
  }
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

// a frame of an EMPTY scene: the arguments are checked like those of any frame, nothing is drawn.  clear < 0: the
// target keeps its contents (render_to_buffer blends onto it); otherwise the host target is filled with `clear`.
static int render_nothing(splat_ctx *c, const splat_camera *cam, uint32_t *host_fb, uint32_t W, uint32_t H, uint32_t row0, uint32_t row1,
                          long long clear) {
  splat_ctx *m = c->members.empty() ? c : c->members[0];
  FrameParams P;
  const int rc = make_params(m, cam, W, H, row0, row1, &P);
  if (rc) { c->err = m->err; return rc; }
  if (host_fb && clear >= 0) std::fill(host_fb, host_fb + (size_t)(row1 - row0) * W, (uint32_t)clear);
  return SPLAT_OK;
}

// a frame skipped on the device cannot be repeated behind the caller's back when the target is the
// caller's device buffer (it may already have been consumed): report it, once, on the next call
static int report_skipped(splat_ctx *c) {
  poll_status(c);
  if (!c->retry_pending) return SPLAT_OK;
  c->retry_pending = false;
  int rc = ensure_instances(c, grow_target(c));
  if (rc) return rc;
  return fail(c, SPLAT_ERR_RETRY, "the previous splat_render_device frame was not rendered (its tile instances did not fit the "
                                  "buffers, which have now been grown; its target is untouched): render it again");
}

int splat_render_device(splat_ctx *c, const splat_camera *cam, void *fb_rows_dev, uint32_t W, uint32_t H,
                        uint32_t row0, uint32_t row1, void *stream) {
  if (!c) return SPLAT_ERR_INVALID;
  if (!c->members.empty()) return fail(c, SPLAT_ERR_UNSUPPORTED, "a group context renders through the host-buffer entry points");
  if (c->empty_scene) return fb_rows_dev ? render_nothing(c, cam, nullptr, W, H, row0, row1, -1) : fail(c, SPLAT_ERR_INVALID, "framebuffer is null");
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
// `restore` puts the frame's initial contents back into d_fb (on c->stream): a skipped frame blended
// nothing if it had no near cut, but the first pass of a near-cut frame may already have written tiles.
static int host_frame(splat_ctx *c, const FrameParams &P, uint32_t *host_fb, size_t px, cudaEvent_t wait_ev,
                      const std::function<int()> &restore) {
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
    rc = restore();
    if (rc) return rc;
    rc = finish_frame(c);          // grows the buffers, renders the frame again with a host round trip
    if (rc) return rc;
  }
  c->host_copy = true;
  return SPLAT_OK;
}

int splat_render_rows(splat_ctx *c, const splat_camera *cam, uint32_t *fb_rows, uint32_t W, uint32_t H,
                      uint32_t row0, uint32_t row1) {
  if (!c) return SPLAT_ERR_INVALID;
  if (c->empty_scene) return fb_rows ? render_nothing(c, cam, fb_rows, W, H, row0, row1, -1) : fail(c, SPLAT_ERR_INVALID, "framebuffer is null");
  if (!c->members.empty()) {
    if (row0 != 0 || row1 != H) return fail(c, SPLAT_ERR_UNSUPPORTED, "a group context renders whole frames (it cuts the stripes itself)");
    return group_render(c, cam, fb_rows, W, H, -1);
  }
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
    dev_free(c->d_fb_bak);
    CU(dev_alloc(&c->d_fb, px));
    CU(dev_alloc(&c->d_fb_bak, px));
    c->fb_cap = px;
  }
  // upload on the copy stream: overlaps project/sort/binning, only blend waits for it; a device-side
  // copy of the uploaded pixels is kept in case the frame has to be repeated (the host buffer is
  // overwritten by the download)
  CU(cudaEventRecord(c->ev[EV_H2D0], c->copy_stream));
  CU(cudaMemcpyAsync(c->d_fb, fb_rows, px * 4, cudaMemcpyHostToDevice, c->copy_stream));
  CU(cudaEventRecord(c->ev[EV_H2D1], c->copy_stream));
  CU(cudaMemcpyAsync(c->d_fb_bak, c->d_fb, px * 4, cudaMemcpyDeviceToDevice, c->copy_stream));
  CU(cudaEventRecord(c->h2d_done, c->copy_stream));
  return host_frame(c, P, fb_rows, px, c->h2d_done, [c, px]() -> int {
    CU(cudaMemcpyAsync(c->d_fb, c->d_fb_bak, px * 4, cudaMemcpyDeviceToDevice, c->stream));
    return SPLAT_OK;
  });
}

int splat_render(splat_ctx *c, const splat_camera *cam, uint32_t *fb, uint32_t W, uint32_t H) {
  return splat_render_rows(c, cam, fb, W, H, 0, H);
}

int splat_render_cleared(splat_ctx *c, const splat_camera *cam, uint32_t *fb_out, uint32_t W, uint32_t H,
                         uint32_t clear) {
  if (!c) return SPLAT_ERR_INVALID;
  if (c->empty_scene) return fb_out ? render_nothing(c, cam, fb_out, W, H, 0, H, (long long)clear) : fail(c, SPLAT_ERR_INVALID, "framebuffer is null");
  if (!c->members.empty()) return group_render(c, cam, fb_out, W, H, (long long)clear);
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
    dev_free(c->d_fb_bak);
    CU(dev_alloc(&c->d_fb, px));
    CU(dev_alloc(&c->d_fb_bak, px));
    c->fb_cap = px;
  }
  auto clear_fb = [c, px, clear]() -> int {
    if (clear == 0u || ((clear & 0xFFu) * 0x01010101u) == clear) {
      CU(cudaMemsetAsync(c->d_fb, (int)(clear & 0xFFu), px * 4, c->stream));
    } else {
      fill_u32_kernel<<<cdiv(px, 1024), 256, 0, c->stream>>>(c->d_fb, clear, px);
    }
    return SPLAT_OK;
  };
  CU(cudaEventRecord(c->ev[EV_H2D0], c->stream));
  rc = clear_fb();
  if (rc) return rc;
  CU(cudaEventRecord(c->ev[EV_H2D1], c->stream));
  return host_frame(c, P, fb_out, px, nullptr, clear_fb);
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
  if (c->empty_scene) { std::memset(t, 0, sizeof(*t)); return SPLAT_OK; }     // frames of an empty scene do no work
  if (!c->members.empty()) {
    if (!c->have_frame) return fail(c, SPLAT_ERR_STATE, "no frame rendered yet");
    return group_timings(c, t);
  }
  if (!c->have_frame) return fail(c, SPLAT_ERR_STATE, "no frame rendered yet");
  CU(cudaSetDevice(c->cfg.device));
  if (c->status_pending) {
    int rcw = wait_done(c, last_status_event(c), c->last_stream, "frame");
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
  if (c->last_cut) {
    t->second_pass_ms = el(EV_BLEND, EV_B_END);
    t->total_ms += t->second_pass_ms;
  }
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
  t->second_pass_instances = c->last_second_instances;
  t->near_cut_fallbacks = c->near_cut_fallbacks;
  return SPLAT_OK;
}

int splat_get_tile_loads(splat_ctx *c, uint32_t *per_tile, uint64_t cap, uint64_t *n_tiles) {
  if (!c || !per_tile || !n_tiles) return SPLAT_ERR_INVALID;
  if (!c->members.empty()) return fail(c, SPLAT_ERR_UNSUPPORTED, "ask the member that rendered the stripe");
  if (!c->have_frame) return fail(c, SPLAT_ERR_STATE, "no frame rendered yet");
  if (!c->loads_valid) return fail(c, SPLAT_ERR_STATE, "the last frame's second pass binned only part of the screen (near cut): no complete tile lists");
  CU(cudaSetDevice(c->cfg.device));
  { int rcw = wait_done(c, c->ev[EV_BLEND], c->last_stream, "frame"); if (rcw) return rcw; }
  poll_status(c);
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
  project_kernel<false><<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->scene, P, c->recs, c->keys[0], c->vals[0], c->rects, c->tcnt, nullptr, nullptr, nullptr);
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

int splat_debug_partition(uint32_t *bounds, const float *ms, int32_t parts, uint32_t H) {
  if (!bounds || parts < 1 || H == 0) return SPLAT_ERR_INVALID;
  std::vector<uint32_t> b;
  if (!ms) {
    equal_bounds(b, H, parts);
  } else {
    b.assign(bounds, bounds + (size_t)2 * parts);
    rebalance_bounds(b, std::vector<float>(ms, ms + parts), H);
  }
  std::copy(b.begin(), b.end(), bounds);
  return SPLAT_OK;
}

int splat_debug_read_tiles(splat_ctx *c, int which, uint32_t *out, uint64_t cap_words, uint64_t *n_words) {
  // tests / diagnostics: per-tile arrays of the last frame.  which: 0 = list ranges (2 words per tile),
  // 1 = far_cnt (cut Gaussians per tile), 2 = tile_failed
  if (!c || !out || !n_words || which < 0 || which > 2) return SPLAT_ERR_INVALID;
  if (!c->have_frame || !c->members.empty()) return fail(c, SPLAT_ERR_STATE, "no frame rendered yet");
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaDeviceSynchronize());
  const uint64_t T = c->last_tiles, words = which == 0 ? 2 * T : T;
  *n_words = words;
  const void *src = which == 0 ? (const void *)c->ranges : (which == 1 ? (const void *)c->far_cnt : (const void *)c->tile_failed);
  CU(cudaMemcpy(out, src, std::min(cap_words, words) * 4u, cudaMemcpyDeviceToHost));
  return SPLAT_OK;
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

int splat_create_multi(splat_ctx **out, const splat_config *cfg, const int32_t *devices, int32_t n_devices) {
  if (!out) return SPLAT_ERR_INVALID;
  *out = nullptr;
  g_create_error[0] = 0;
  if (!devices || n_devices < 1 || n_devices > 64) {
    std::snprintf(g_create_error, sizeof g_create_error, "devices must list 1..64 CUDA device ordinals");
    return SPLAT_ERR_INVALID;
  }
  NcclApi *N = nccl_api();
  if (!N) {
    std::snprintf(g_create_error, sizeof g_create_error, "NCCL is not available: %s", NcclApi().error.c_str());
    return SPLAT_ERR_UNSUPPORTED;
  }
  splat_ctx *g = new (std::nothrow) splat_ctx();
  if (!g) return SPLAT_ERR_NOMEM;
  if (cfg) g->cfg = *cfg; else splat_config_default(&g->cfg);
  g->cfg.device = devices[0];
  for (int k = 0; k < n_devices; ++k) {
    splat_config mc = g->cfg;
    mc.device = devices[k];
    mc.near_cut = 0;                 // stripes use the no-round-trip path
    splat_ctx *m = nullptr;
    const int rc = splat_create(&m, &mc);
    if (rc) { splat_destroy(g); return rc; }
    g->members.push_back(m);
  }
  std::vector<ncclComm_t> comms((size_t)n_devices, nullptr);
  std::vector<int> devs(devices, devices + n_devices);
  const ncclResult_t r = N->CommInitAll(comms.data(), n_devices, devs.data());
  if (r != ncclSuccess) {
    std::snprintf(g_create_error, sizeof g_create_error, "ncclCommInitAll: %s", N->GetErrorString(r));
    splat_destroy(g);
    return SPLAT_ERR_CUDA;
  }
  for (int k = 0; k < n_devices; ++k) {
    g->members[k]->comm = comms[k];
    g->members[k]->owns_comm = true;
    g->members[k]->n_ranks = n_devices;
    g->members[k]->rank = k;
  }
  g->n_ranks = n_devices;
  *out = g;
  return SPLAT_OK;
}

int splat_comm_unique_id(void *id128) {
  if (!id128) return SPLAT_ERR_INVALID;
  NcclApi *N = nccl_api();
  if (!N) return SPLAT_ERR_UNSUPPORTED;
  static_assert(sizeof(ncclUniqueId) == SPLAT_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  return N->GetUniqueId(static_cast<ncclUniqueId *>(id128)) == ncclSuccess ? SPLAT_OK : SPLAT_ERR_CUDA;
}

int splat_comm_init_rank(splat_ctx *c, const void *id128, int32_t n_ranks, int32_t rank) {
  if (!c || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return SPLAT_ERR_INVALID;
  if (!c->members.empty()) return fail(c, SPLAT_ERR_STATE, "a group context owns its communicators already");
  if (c->comm) return fail(c, SPLAT_ERR_STATE, "the context already joined a communicator");
  NcclApi *N = nccl_api();
  if (!N) return fail(c, SPLAT_ERR_UNSUPPORTED, "NCCL is not available");
  CU(cudaSetDevice(c->cfg.device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  NC(N->CommInitRank(&c->comm, n_ranks, id, rank));
  c->owns_comm = true;
  c->n_ranks = n_ranks;
  c->rank = rank;
  return SPLAT_OK;
}

int splat_comm_broadcast_scene(splat_ctx *c, int32_t root, uint64_t n) {
  if (!c || !c->comm) return c ? fail(c, SPLAT_ERR_STATE, "no communicator") : SPLAT_ERR_INVALID;
  if (root < 0 || root >= c->n_ranks) return fail(c, SPLAT_ERR_INVALID, "bad root");
  NcclApi *N = nccl_api();
  CU(cudaSetDevice(c->cfg.device));
  if (c->rank == root) {
    if (!c->n || c->n != n) return fail(c, SPLAT_ERR_STATE, "root: upload the scene first (n must be its size)");
  } else {
    int rc = alloc_scene(c, n);
    if (rc) return rc;
  }
  NC(N->Broadcast(c->scene, c->scene, (size_t)SCENE_PLANES * n * 4u, ncclFloat, root, c->comm, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SPLAT_OK;
}

int splat_gather_stripes(splat_ctx *c, void *fb_dev, uint32_t W, uint32_t H, const uint32_t *bounds, int32_t root, void *stream) {
  if (!c || !fb_dev || !bounds) return SPLAT_ERR_INVALID;
  if (c->n_ranks <= 1) return SPLAT_OK;
  if (!c->comm) return fail(c, SPLAT_ERR_STATE, "no communicator");
  if (root < 0 || root >= c->n_ranks) return fail(c, SPLAT_ERR_INVALID, "bad root");
  uint32_t prev = 0;
  for (int r = 0; r < c->n_ranks; ++r) {
    if (bounds[2 * r] != prev || bounds[2 * r + 1] < bounds[2 * r]) return fail(c, SPLAT_ERR_INVALID, "stripes must be contiguous and ordered");
    prev = bounds[2 * r + 1];
  }
  if (prev != H) return fail(c, SPLAT_ERR_INVALID, "stripes must cover the image");
  NcclApi *N = nccl_api();
  CU(cudaSetDevice(c->cfg.device));
  NC(gather_stripes(N, c->comm, c->n_ranks, c->rank, root, static_cast<uint32_t *>(fb_dev), W, bounds,
                    stream ? static_cast<cudaStream_t>(stream) : c->stream));
  return SPLAT_OK;
}

int splat_group_get_bounds(splat_ctx *c, uint32_t *bounds, int32_t cap_ranks, int32_t *n_ranks) {
  if (!c || !n_ranks) return SPLAT_ERR_INVALID;
  *n_ranks = (int32_t)c->members.size();
  if (c->members.empty()) return fail(c, SPLAT_ERR_STATE, "not a group context");
  if (bounds)
    for (int k = 0; k < *n_ranks && k < cap_ranks && (size_t)(2 * k + 1) < c->bounds.size(); ++k) {
      bounds[2 * k] = c->bounds[2 * k];
      bounds[2 * k + 1] = c->bounds[2 * k + 1];
    }
  return SPLAT_OK;
}

// (a refused registration is not sticky, but it would be the "last error" the next frame's launch check reads)
int splat_pin_host(void *p, uint64_t bytes) {
  if (!p || bytes == 0) return SPLAT_ERR_INVALID;
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) return SPLAT_OK;
  (void)cudaGetLastError();
  return SPLAT_ERR_CUDA;
}
int splat_unpin_host(void *p) {
  if (!p) return SPLAT_ERR_INVALID;
  if (cudaHostUnregister(p) == cudaSuccess) return SPLAT_OK;
  (void)cudaGetLastError();
  return SPLAT_ERR_CUDA;
}

}  // extern "C"
