// blend.cuh -- K5: per-tile, per-pixel far->near compositing with the reference's quantised
// blend.  Replaces, for every pixel a quad covers,
//   fragment()  pipelines.rs:127-145 (=:215-233): Gaussian falloff, alpha clamp / early-outs
//   blend()     pipelines.rs:147-168 (=:235-256): decode u8 pixel, "over", truncate back to u8
// and euc's coverage / interpolation loop.
//
// The reference truncates the pixel to 8 bits after EVERY Gaussian, so compositing is not
// associative: no transmittance prefix scan, no early termination -- each pixel walks every
// covering Gaussian in order and must reproduce the f32 operation sequence exactly:
//   old = byte / 255           (IEEE division; here a 2-op fmaf form, exact for all 256 bytes)
//   out = (1-a)*old + a*new    (two products, one add, no FMA)
//   byte' = trunc_sat(out*255) (clamp + round-toward-zero add of 2^23)
// exp() is the pinned "splat_expf v1" sequence shared with the oracle (oracle/splat_oracle.c).
//
// Work decomposition: one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel sub-tile.
// The tile's sorted list is staged through shared memory 256 entries at a time; while
// staging, each thread also tests its entry's 3-sigma rectangle against the 8 sub-tiles and
// stores an 8-bit overlap mask, so a warp only iterates (ballot + ffs) over the entries that
// can touch its 32 pixels.
#pragma once
#include "common.cuh"

namespace splat {

constexpr int BL_THREADS = 256;
constexpr int BL_BATCH = 256;

// splat_expf v1 on its hot domain [-87, 0] (callers guarantee the domain).
SPLAT_DEVINL float expf_pinned(float x) {
  const float MAGIC = 12582912.0f;
  const float tm = __fmaf_rn(x, 0x1.715476p+0f, MAGIC);
  const float n = __fsub_rn(tm, MAGIC);
  float r = __fmaf_rn(n, -0x1.62e4p-1f, x);
  r = __fmaf_rn(n, -0x1.7f7d1cp-20f, r);
  float p = 0x1.687b46p-10f;
  p = __fmaf_rn(p, r, 0x1.123bdcp-7f);
  p = __fmaf_rn(p, r, 0x1.555b5cp-5f);
  p = __fmaf_rn(p, r, 0x1.55548ep-3f);
  p = __fmaf_rn(p, r, 0x1.fffff8p-2f);
  p = __fmaf_rn(p, r, 1.0f);
  p = __fmaf_rn(p, r, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(tm) << 23));
}

// byte / 255.0f, correctly rounded for every integer 0..255 (checked exhaustively in
// tests/test_host_math.py): q = fma(n, RN(1/255), n * (1/255 - RN(1/255))).
SPLAT_DEVINL float div255(float n) {
  return __fmaf_rn(n, 0x1.010102p-8f, __fmul_rn(n, -0x1.fdfdfep-33f));
}

// One channel of blend(): returns the new channel state (= new byte / 255).
SPLAT_DEVINL float blend_channel(float c_old, float om, float u) {
  const float out = __fadd_rn(__fmul_rn(om, c_old), u);
  float v = __fmul_rn(out, 255.0f);
  v = fminf(fmaxf(v, 0.0f), 255.0f);                                   // `as u8` saturates, NaN -> 0
  const float byte = __fsub_rn(__fadd_rz(v, 8388608.0f), 8388608.0f);  // truncate toward zero
  return div255(byte);
}

__global__ void __launch_bounds__(BL_THREADS)
blend_kernel(const uint2 *__restrict__ ranges, const uint32_t *__restrict__ inst_vals,
             const Rec *__restrict__ recs, uint32_t *__restrict__ fb_rows,
             const __grid_constant__ FrameParams P) {
  __shared__ float4 sa[BL_BATCH], sb[BL_BATCH], sc[BL_BATCH];
  __shared__ uint8_t smask[BL_BATCH];

  const uint32_t tile = blockIdx.y * P.tiles_x + blockIdx.x;
  const uint2 range = ranges[tile];
  if (range.y <= range.x) return;   // nothing touches this tile: pixels stay as they are

  const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t tx0 = blockIdx.x * TILE, ty0 = (P.tile_y0 + blockIdx.y) * TILE;
  const uint32_t px = tx0 + 8u * (w & 1u) + (lane & 7u);
  const uint32_t py = ty0 + 4u * (w >> 1) + (lane >> 3);
  const bool inside = px < P.W && py < P.row1;
  const float sx = (float)px + P.sample_off, sy = (float)py + P.sample_off;

  uint32_t *pix = fb_rows + (size_t)(py - P.row0) * P.W + px;
  uint32_t old = 0;
  if (inside) old = *pix;
  float cr = div255((float)((old >> 16) & 0xFFu));
  float cg = div255((float)((old >> 8) & 0xFFu));
  float cb = div255((float)(old & 0xFFu));
  float last_alpha = -1.0f;   // < 0: no quad covered this pixel yet

  // sub-tile sample intervals for the overlap masks (same (float)p + off as sx/sy above)
  float sxl[2], sxh[2], syl[4], syh[4];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    sxl[c] = (float)(tx0 + 8u * c) + P.sample_off;
    sxh[c] = (float)(tx0 + 8u * c + 7u) + P.sample_off;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    syl[r] = (float)(ty0 + 4u * r) + P.sample_off;
    syh[r] = (float)(ty0 + 4u * r + 3u) + P.sample_off;
  }

  for (uint32_t base = range.x; base < range.y; base += BL_BATCH) {
    const uint32_t nb = min((uint32_t)BL_BATCH, range.y - base);
    __syncthreads();   // previous batch fully consumed
    if (tid < nb) {
      const uint32_t g = __ldg(&inst_vals[base + tid]);
      const float4 *rp = reinterpret_cast<const float4 *>(recs + g);
      const float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
      sa[tid] = a; sb[tid] = b; sc[tid] = c;
      // |RN(s - cxp)| >= RN(dist(cxp, [lo,hi])) for every sample s in [lo,hi] (rounding is
      // monotone), so "dist > h" proves that no pixel of the sub-tile passes |dx| <= h.
      uint32_t ox = 0, oy = 0;
#pragma unroll
      for (int q = 0; q < 2; ++q)
        ox |= (fmaxf(fmaxf(sxl[q] - a.x, a.x - sxh[q]), 0.0f) <= b.z) ? (1u << q) : 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        oy |= (fmaxf(fmaxf(syl[q] - a.y, a.y - syh[q]), 0.0f) <= b.w) ? (1u << q) : 0u;
      uint32_t m = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) m |= (((ox >> (q & 1)) & (oy >> (q >> 1))) & 1u) << q;
      smask[tid] = (uint8_t)m;
    }
    __syncthreads();

    for (uint32_t c0 = 0; c0 < nb; c0 += 32) {
      const uint32_t e = c0 + lane;
      const uint32_t mine = (e < nb) ? ((smask[e] >> w) & 1u) : 0u;
      uint32_t todo = __ballot_sync(0xFFFFFFFFu, mine);
      while (todo) {
        const uint32_t j = c0 + (uint32_t)__ffs(todo) - 1u;
        todo &= todo - 1u;
        const float4 a = sa[j], b = sb[j], c = sc[j];
        const float dx = sx - a.x, dy = sy - a.y;
        const bool inr = (fabsf(dx) <= b.z) && (fabsf(dy) <= b.w);
        // pipelines.rs:134, left to right, no FMA
        const float q1 = __fmul_rn(__fmul_rn(a.z, dx), dx);
        const float q2 = __fmul_rn(__fmul_rn(b.x, dy), dy);
        const float q3 = __fmul_rn(__fmul_rn(a.w, dx), dy);
        const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(q1, q2)), q3);
        const bool cand = inr && !(power > 0.0f) && (power >= c.w);
        bool contrib = false;
        float al = 0.0f;
        if (__any_sync(0xFFFFFFFFu, cand)) {
          const float ex = expf_pinned(power);
          al = fminf(0.99f, __fmul_rn(b.y, ex));          // pipelines.rs:139
          contrib = cand && !(al < (1.0f / 255.0f));      // pipelines.rs:140
          if (contrib) {
            const float om = __fsub_rn(1.0f, al);
            cr = blend_channel(cr, om, __fmul_rn(al, c.x));
            cg = blend_channel(cg, om, __fmul_rn(al, c.y));
            cb = blend_channel(cb, om, __fmul_rn(al, c.z));
          }
        }
        // E7: a zero fragment is still blended -- RGB unchanged, alpha byte reset to 0
        if (inr) last_alpha = contrib ? al : 0.0f;
      }
    }
  }

  if (inside && last_alpha >= 0.0f) {
    const uint32_t r = (uint32_t)__fmul_rn(cr, 255.0f), g = (uint32_t)__fmul_rn(cg, 255.0f);
    const uint32_t bl = (uint32_t)__fmul_rn(cb, 255.0f), a = (uint32_t)__fmul_rn(last_alpha, 255.0f);
    *pix = bl | (g << 8) | (r << 16) | (a << 24);
  }
}

}  // namespace splat
