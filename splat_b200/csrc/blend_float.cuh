// blend_float.cuh -- K5f: SPLAT_BLEND_FLOAT, the un-quantised front-to-back compositor
// (BASELINE north star: "tiled per-pixel front-to-back alpha compositing ...
// TMA staging of each tile's sorted Gaussian list into shared memory ... running transmittance").
//
// NOT the reference's pixels: the reference truncates to u8 after every Gaussian
// (pipelines.rs:147-168), which is what blend.cuh reproduces bit for bit.  This mode is
// the standard 3DGS formulation of the same image: per pixel, nearest first,
//     C += T * alpha_i * colour_i ;  T *= (1 - alpha_i)
// with fragment() exactly as the reference has it (pipelines.rs:127-145: d = pixel centre - centre,
// power, alpha = min(0.99, opacity * exp(power)), dropped when power > 0 or alpha < 1/255; the same
// pinned exp as the parity path), then ONE quantisation at the end:
//     out = C + T * old_pixel/255 ;  byte = trunc_sat(out * 255) ;  alpha byte = trunc_sat((1 - T) * 255).
// Pixels no fragment contributes to are left untouched.  The only float compositor the reference
// tree holds is the prototype's plot_opacity (notebook cell 3); the CPU checker restates it
// (orc_render_float mode 1, pinned to the executed cell) and this kernel is compared with the same
// recurrence under the Rust fragment() rules (mode 0) to <= 1e-4 RMSE (tests/test_gpu_float.py).
//
// Structure: one CTA per 16x16 tile, one thread per pixel, a warp = an 8x4 pixel block.  The
// tile's list (indices into the 48-byte records, far -> near) is walked from its END in batches of
// BF_BATCH entries.  Staging is a TMA gather: thread j reads list index (end-1-j) and issues ONE
// cp.async.bulk of that 48-byte record into shared-memory slot j, all copies of a batch completing
// on one mbarrier (expect_tx = 48 * entries); two stages, so the copies of batch b+1 fly while
// batch b is composited.  Early out: a pixel is finished when T < 2^-16 (its remaining
// contribution is below 1.6e-5 of full scale); a warp skips entries once all its pixels are
// finished or when the entry's 3-sigma rectangle misses its 8x4 block (warp votes), and the CTA
// stops staging when every pixel of the tile is finished (__syncthreads_and).
#pragma once
#include "blend.cuh"
#include "common.cuh"

namespace splat {

constexpr int BF_THREADS = 256;
constexpr int BF_BATCH = 128;
constexpr int BF_STAGES = 2;
constexpr float BF_T_MIN = 1.0f / 65536.0f;

struct BlendFloatSmem {
  Rec rec[BF_STAGES][BF_BATCH];     // 48-byte records, filled by cp.async.bulk
  uint64_t full[BF_STAGES];         // mbarriers: "stage holds batch b"
};

__global__ void __launch_bounds__(BF_THREADS)
blend_float_kernel(const uint2 *__restrict__ ranges, const uint2 *__restrict__ units, const uint32_t *__restrict__ n_units,
                   const uint32_t *__restrict__ inst_vals, const Rec *__restrict__ recs, uint32_t *__restrict__ fb_rows,
                   const __grid_constant__ FrameParams P, float4 *__restrict__ tap, uint32_t *__restrict__ wd) {
  __shared__ __align__(128) BlendFloatSmem S;
  if (blockIdx.x >= *n_units) return;
  const uint32_t tile = units[blockIdx.x].x;
  const uint2 range = ranges[tile];
  const uint32_t len = range.y - range.x;
  const uint32_t tile_x = tile % P.tiles_x, tile_y = tile / P.tiles_x;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  // warp w = 8x4 block (w & 1, w >> 1) of the tile; lane = (x & 7) + 8 * (y & 3)
  const uint32_t bx0 = tile_x * TILE + 8u * (w & 1u), by0 = (P.tile_y0 + tile_y) * TILE + 4u * (w >> 1);
  const uint32_t px = bx0 + (lane & 7u), py = by0 + (lane >> 3);
  const bool inside = px < P.W && py < P.row1;
  const float sx = (float)px + P.sample_off, sy = (float)py + P.sample_off;
  const float bxl = (float)bx0 + P.sample_off, bxh = (float)(bx0 + 7u) + P.sample_off;
  const float byl = (float)by0 + P.sample_off, byh = (float)(by0 + 3u) + P.sample_off;

  if (tid < BF_STAGES) mbar_init(&S.full[tid], 1);
  if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // inits visible to the async proxy
  __syncthreads();

  const uint32_t nbatch = (len + BF_BATCH - 1) / BF_BATCH;
  // entry j of batch b is list position range.y - 1 - (b * BF_BATCH + j): nearest first
  auto issue = [&](uint32_t b) {
    const uint32_t first = b * BF_BATCH, nb = min((uint32_t)BF_BATCH, len - first);
    uint64_t *bar = &S.full[b % BF_STAGES];
    if (tid == 0) mbar_expect_tx(bar, nb * (uint32_t)sizeof(Rec));
    if (tid < nb) {
      const uint32_t gi = __ldg(&inst_vals[range.y - 1u - first - tid]);
      tma_load_1d(&S.rec[b % BF_STAGES][tid], recs + gi, (uint32_t)sizeof(Rec), bar);
    }
  };
  issue(0);
  if (nbatch > 1) issue(1);

  float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
  bool done = !inside, touched = false;
  uint32_t b = 0;
  bool all_done = false;
  for (; b < nbatch; ++b) {
    const uint32_t st = b % BF_STAGES, nb = min((uint32_t)BF_BATCH, len - b * BF_BATCH);
    mbar_wait<3>(&S.full[st], (b / BF_STAGES) & 1u, nullptr, wd, (4u << 28) | (st << 20) | (b & 0xFFFFFu));
    if (!__all_sync(0xFFFFFFFFu, done)) {
      for (uint32_t j = 0; j < nb; ++j) {
        const float4 a = S.rec[st][j].a, bb = S.rec[st][j].b;
        // warp-uniform: does the 3-sigma rectangle reach this warp's 8x4 block at all?
        if (fmaxf(fmaxf(bxl - a.x, a.x - bxh), 0.0f) > bb.z || fmaxf(fmaxf(byl - a.y, a.y - byh), 0.0f) > bb.w) continue;
        const float dx = sx - a.x, dy = sy - a.y;
        // pipelines.rs:134, left to right, no FMA (file is compiled --fmad=false)
        const float power = -0.5f * (a.z * dx * dx + bb.x * dy * dy) - a.w * dx * dy;
        const float4 c = S.rec[st][j].c;
        if (done || !(fabsf(dx) <= bb.z) || !(fabsf(dy) <= bb.w) || power > 0.0f || !(power >= c.w)) continue;
        const float alpha = fminf(0.99f, bb.y * expf_pinned(power));       // pipelines.rs:139 (power >= pth >= -87)
        if (alpha < (1.0f / 255.0f)) continue;                              // pipelines.rs:140
        const float wgt = T * alpha;
        cr = fmaf(wgt, c.x, cr); cg = fmaf(wgt, c.y, cg); cb = fmaf(wgt, c.z, cb);
        T = T - wgt;
        touched = true;
        done = T < BF_T_MIN;
      }
    }
    // every pixel finished -> nothing further down the list can change the tile; also the
    // barrier that lets stage `st` be refilled
    if (__syncthreads_and(done)) { all_done = true; break; }
    if (b + BF_STAGES < nbatch) issue(b + BF_STAGES);
  }
  // a CTA must not exit with bulk copies still landing in its shared memory: batch b+1 was issued
  if (all_done && b + 1u < nbatch)
    mbar_wait<3>(&S.full[(b + 1u) % BF_STAGES], ((b + 1u) / BF_STAGES) & 1u, nullptr, wd, (5u << 28) | ((b + 1u) & 0xFFFFFu));

  if (inside && touched) {
    uint32_t *pix = fb_rows + (size_t)(py - P.row0) * P.W + px;
    const uint32_t old = *pix;
    const float orr = div255((float)((old >> 16) & 0xFFu)), og = div255((float)((old >> 8) & 0xFFu)), ob = div255((float)(old & 0xFFu));
    const float r = fmaf(T, orr, cr), g = fmaf(T, og, cg), bl = fmaf(T, ob, cb), al = 1.0f - T;
    // `as u8`: truncate, saturate, NaN -> 0
    const uint32_t rb = (uint32_t)(__saturatef(r) * 255.0f), gb = (uint32_t)(__saturatef(g) * 255.0f);
    const uint32_t bbv = (uint32_t)(__saturatef(bl) * 255.0f), ab = (uint32_t)(__saturatef(al) * 255.0f);
    *pix = bbv | (gb << 8) | (rb << 16) | (ab << 24);
    if (tap) tap[(size_t)(py - P.row0) * P.W + px] = make_float4(r, g, bl, al);
  }
}

}  // namespace splat
