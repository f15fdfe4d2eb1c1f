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
//   byte' = trunc_sat(out*255) (saturating add + round-toward-zero add of 2^23)
// exp() is the pinned "splat_expf v1" sequence shared with the oracle (oracle/splat_oracle.c).
//
// Work decomposition (r1b; the first version -- one 8-warp CTA per tile, every warp doing
// alpha AND blend -- left the SMs 45% idle because tile lists are extremely skewed: the
// heaviest tile holds 275k of 48M instances and its 8 sequential warp streams were the frame's
// critical path):
//   * work unit = half a 16x16 tile (16x8 pixels); units are issued heaviest-first
//     (tile_order_kernel) so the long units start at t=0 and short ones fill the tail;
//   * per 8x4-pixel group a TEAM of three warps: two PRODUCER warps evaluate
//     fragment() -- coverage, power, exp, alpha -- for alternating chunks of 8 list entries and
//     publish one alpha per pixel (+ the entry's colour) into a shared-memory ring; one
//     CONSUMER warp runs only the strictly sequential blend() chain.  Producers and consumer
//     hand chunks over through mbarriers (full/empty per ring slot), so the per-pixel chain is
//     ~36 instructions per entry instead of ~80 and the independent part runs ahead of it;
//   * the unit's sorted list is staged through shared memory 256 entries at a time by the 8
//     producer warps (one entry per thread), which also compact, per group, the indices of
//     the entries whose 3-sigma rectangle can touch that group's 32 pixels.
#pragma once
#include "common.cuh"

namespace splat {

constexpr int BL_GROUPS = 4;                       // 8x4-pixel groups per CTA (16x8 pixels)
constexpr int BL_PRODUCER_THREADS = 256;           // 8 producer warps: team = warp>>1, p = warp&1
constexpr int BL_THREADS = BL_PRODUCER_THREADS + 32 * BL_GROUPS;   // + one consumer warp per team
constexpr int BL_BATCH = 256;                      // list entries staged per round
constexpr int BL_CH = 8;                           // entries per ring chunk
constexpr int BL_D = 4;                            // ring depth in chunks (2 per producer)
constexpr int BL_SLOT_F = 36;                      // floats per ring entry: 32 alphas + rgb + pad

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
// `as u8` saturates and maps NaN to 0: trunc_sat(out*255) == trunc(sat01(out)*255), because
// out >= 1 gives 255 either way and out <= 0 / NaN give 0; the saturation rides on the FADD.
SPLAT_DEVINL float blend_channel(float c_old, float om, float u) {
  const float out = __saturatef(__fadd_rn(__fmul_rn(om, c_old), u));
  const float v = __fmul_rn(out, 255.0f);
  const float byte = __fsub_rn(__fadd_rz(v, 8388608.0f), 8388608.0f);  // truncate toward zero
  return div255(byte);
}

// ---------------------------------------------------------------- mbarrier / named barrier
SPLAT_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SPLAT_DEVINL void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SPLAT_DEVINL void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
SPLAT_DEVINL void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware instead of polling
      : "memory");
}
SPLAT_DEVINL void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(BL_PRODUCER_THREADS) : "memory"); }

// ---------------------------------------------------------------- heaviest-first unit order
// order[rank] = tile id, tiles sorted by descending list length (bucketed on 16*log2(len)):
// single CTA counting sort.  Also publishes nothing else; empty tiles end up last.
__global__ void __launch_bounds__(1024)
tile_order_kernel(const uint2 *__restrict__ ranges, uint32_t T, uint32_t *__restrict__ order) {
  constexpr int NB = 512;
  __shared__ uint32_t hist[NB];
  for (int i = threadIdx.x; i < NB; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  auto bucket = [](uint2 r) -> uint32_t {
    const uint32_t len = r.y > r.x ? r.y - r.x : 0u;
    if (len == 0) return NB - 1;
    const int b = (int)(16.0f * __log2f((float)len));   // 0 .. 16*32-1
    return (uint32_t)max(0, NB - 2 - b);
  };
  for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) atomicAdd(&hist[bucket(ranges[t])], 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < NB; ++i) { const uint32_t c = hist[i]; hist[i] = run; run += c; }
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) order[atomicAdd(&hist[bucket(ranges[t])], 1u)] = t;
}

// ---------------------------------------------------------------- K5
struct BlendSmem {
  float4 sa[BL_BATCH], sb[BL_BATCH], sc[BL_BATCH];        // staged records (see Rec)
  float ring[BL_GROUPS][BL_D][BL_CH][BL_SLOT_F];          // per team: alpha per pixel + colour
  uint64_t full[BL_GROUPS][BL_D], empty[BL_GROUPS][BL_D]; // mbarriers
  uint32_t hdr[BL_GROUPS][BL_D];                          // entries in the chunk | last << 8
  uint32_t wcount[BL_GROUPS][BL_PRODUCER_THREADS / 32];   // per staging warp, per group
  uint8_t list[BL_GROUPS][BL_BATCH];                      // compacted entry indices per group
};

__global__ void __launch_bounds__(BL_THREADS, 3)
blend_kernel(const uint2 *__restrict__ ranges, const uint32_t *__restrict__ order,
             const uint32_t *__restrict__ inst_vals, const Rec *__restrict__ recs,
             uint32_t *__restrict__ fb_rows, const __grid_constant__ FrameParams P) {
  __shared__ BlendSmem S;

  const uint32_t tile = order[blockIdx.x >> 1], half = blockIdx.x & 1u;
  const uint2 range = ranges[tile];
  if (range.y <= range.x) return;   // nothing touches this tile: pixels stay as they are
  const uint32_t tile_x = tile % P.tiles_x, tile_y = tile / P.tiles_x;
  const uint32_t tx0 = tile_x * TILE, ty0 = (P.tile_y0 + tile_y) * TILE + 8u * half;
  if (ty0 >= P.row1) return;        // lower half of a ragged last tile row

  const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  if (tid < BL_GROUPS * BL_D) {
    mbar_init(&S.full[tid / BL_D][tid % BL_D], 1);
    mbar_init(&S.empty[tid / BL_D][tid % BL_D], 1);
  }
  __syncthreads();

  if (w < BL_PRODUCER_THREADS / 32) {
    // ============================== PRODUCERS ==============================
    const uint32_t g = w >> 1, p = w & 1u;
    const float sx = (float)(tx0 + 8u * (g & 1u) + (lane & 7u)) + P.sample_off;
    const float sy = (float)(ty0 + 4u * (g >> 1) + (lane >> 3)) + P.sample_off;
    // group sample intervals for the overlap test (same (float)p + off as sx/sy)
    float sxl[2], sxh[2], syl[2], syh[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      sxl[q] = (float)(tx0 + 8u * q) + P.sample_off;
      sxh[q] = (float)(tx0 + 8u * q + 7u) + P.sample_off;
      syl[q] = (float)(ty0 + 4u * q) + P.sample_off;
      syh[q] = (float)(ty0 + 4u * q + 3u) + P.sample_off;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t seq = 0;   // entries of this team published so far (both producers count alike)

    for (uint32_t base = range.x; base < range.y; base += BL_BATCH) {
      const uint32_t nb = min((uint32_t)BL_BATCH, range.y - base);
      producers_sync();   // previous batch no longer read by any producer
      uint32_t bits = 0;
      if (tid < nb) {
        const uint32_t gi = __ldg(&inst_vals[base + tid]);
        const float4 *rp = reinterpret_cast<const float4 *>(recs + gi);
        const float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
        S.sa[tid] = a; S.sb[tid] = b; S.sc[tid] = c;
        // |RN(s - cxp)| >= RN(dist(cxp, [lo,hi])) for every sample s in [lo,hi] (rounding is
        // monotone), so "dist > h" proves that no pixel of the group passes |dx| <= h.
        uint32_t ox = 0, oy = 0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          ox |= (fmaxf(fmaxf(sxl[q] - a.x, a.x - sxh[q]), 0.0f) <= b.z) ? (1u << q) : 0u;
          oy |= (fmaxf(fmaxf(syl[q] - a.y, a.y - syh[q]), 0.0f) <= b.w) ? (1u << q) : 0u;
        }
#pragma unroll
        for (int q = 0; q < BL_GROUPS; ++q) bits |= (((ox >> (q & 1)) & (oy >> (q >> 1))) & 1u) << q;
        // Second, tighter test: can ANY sample of the group reach power >= pth?  power = -q/2
        // with q(dx,dy) = A dx^2 + 2B dx dy + C dy^2; for a positive-definite conic the minimum
        // of q over the group's sample rectangle is 0 if the centre is inside, else it lies on
        // one of the four edges (1-D clamped minimisation).  The bound is evaluated in f32
        // with a relative slack 100x above its rounding error, so it never removes a pair
        // that the exact per-pixel test would accept.  (The alpha byte, which also depends
        // on non-contributing covered entries, is resolved separately by the consumer.)
        const float cA = a.z, cB = a.w, cC = b.x, pth = c.w;
        if (pth > 0.0f) bits = 0;   // opacity < 1/255: power <= 0 < pth can never pass
        if (bits && cA > 0.0f && cC > 0.0f && cA * cC - cB * cB > 0.0f) {
          const float invA = __fdiv_rn(1.0f, cA), invC = __fdiv_rn(1.0f, cC);
#pragma unroll
          for (int q = 0; q < BL_GROUPS; ++q) {
            if (!((bits >> q) & 1u)) continue;
            const float xl = sxl[q & 1] - a.x, xh = sxh[q & 1] - a.x;
            const float yl = syl[q >> 1] - a.y, yh = syh[q >> 1] - a.y;
            if (xl <= 0.0f && xh >= 0.0f && yl <= 0.0f && yh >= 0.0f) continue;   // centre inside
            float qmin = 3.0e38f;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float x0 = e ? xh : xl;
              const float ty = fminf(fmaxf(-cB * x0 * invC, yl), yh);
              qmin = fminf(qmin, cA * x0 * x0 + (2.0f * cB * x0 + cC * ty) * ty);
              const float y0 = e ? yh : yl;
              const float tx = fminf(fmaxf(-cB * y0 * invA, xl), xh);
              qmin = fminf(qmin, cC * y0 * y0 + (2.0f * cB * y0 + cA * tx) * tx);
            }
            const float xm = fmaxf(fabsf(xl), fabsf(xh)), ym = fmaxf(fabsf(yl), fabsf(yh));
            const float slack = 1e-5f * (cA * xm * xm + cC * ym * ym + 2.0f * fabsf(cB) * xm * ym) + 1e-4f;
            if (!(0.5f * qmin <= -pth + slack)) bits &= ~(1u << q);
          }
        }
      }
      uint32_t rank[BL_GROUPS];
#pragma unroll
      for (int q = 0; q < BL_GROUPS; ++q) {
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (bits >> q) & 1u);
        rank[q] = __popc(bal & lt_mask);
        if (lane == 0) S.wcount[q][w] = __popc(bal);
      }
      producers_sync();
      uint32_t n_mine = 0;
#pragma unroll
      for (int q = 0; q < BL_GROUPS; ++q) {
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int ww = 0; ww < BL_PRODUCER_THREADS / 32; ++ww) {
          const uint32_t c = S.wcount[q][ww];
          before += (ww < (int)w) ? c : 0u;
          total += c;
        }
        if ((bits >> q) & 1u) S.list[q][before + rank[q]] = (uint8_t)tid;
        if (q == (int)g) n_mine = total;
      }
      producers_sync();

      uint32_t s = seq;
      const uint32_t s_end = seq + n_mine;
      while (s < s_end) {
        const uint32_t chunk = s / BL_CH;
        const uint32_t chunk_end = min(s_end, (chunk + 1) * BL_CH);
        if ((chunk & 1u) != p) { s = chunk_end; continue; }
        const uint32_t slot = chunk % BL_D;
        if (s % BL_CH == 0) mbar_wait(&S.empty[g][slot], ((chunk / BL_D) & 1u) ^ 1u);
        float *slotp = &S.ring[g][slot][s % BL_CH][0];
        for (; s < chunk_end; ++s) {
          const uint32_t j = S.list[g][s - seq];
          const float4 a = S.sa[j], b = S.sb[j], c = S.sc[j];
          const float dx = sx - a.x, dy = sy - a.y;
          const bool inr = (fabsf(dx) <= b.z) && (fabsf(dy) <= b.w);
          // pipelines.rs:134, left to right, no FMA
          const float q1 = __fmul_rn(__fmul_rn(a.z, dx), dx);
          const float q2 = __fmul_rn(__fmul_rn(b.x, dy), dy);
          const float q3 = __fmul_rn(__fmul_rn(a.w, dx), dy);
          const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(q1, q2)), q3);
          const bool cand = inr && !(power > 0.0f) && (power >= c.w);
          float val = 0.0f;   // zero fragment: RGB unchanged (its alpha-byte effect: see consumer)
          if (__any_sync(0xFFFFFFFFu, cand)) {
            const float ex = expf_pinned(power);
            const float al = fminf(0.99f, __fmul_rn(b.y, ex));     // pipelines.rs:139
            if (cand && !(al < (1.0f / 255.0f))) val = al;          // pipelines.rs:140
          }
          slotp[lane] = val;
          if (lane == 0) *reinterpret_cast<float4 *>(slotp + 32) = c;   // rgb (+ threshold, unused)
          slotp += BL_SLOT_F;
        }
        if (chunk_end % BL_CH == 0) {
          __syncwarp();
          if (lane == 0) {
            S.hdr[g][slot] = BL_CH;
            mbar_arrive(&S.full[g][slot]);
          }
        }
      }
      seq = s_end;
    }
    // terminator: the partial last chunk, or an empty extra chunk, carries the `last` flag
    {
      const uint32_t chunk = seq / BL_CH, rem = seq % BL_CH;
      if ((chunk & 1u) == p) {
        const uint32_t slot = chunk % BL_D;
        if (rem == 0) mbar_wait(&S.empty[g][slot], ((chunk / BL_D) & 1u) ^ 1u);
        __syncwarp();
        if (lane == 0) {
          S.hdr[g][slot] = rem | 0x100u;
          mbar_arrive(&S.full[g][slot]);
        }
      }
    }
  } else {
    // ============================== CONSUMERS ==============================
    const uint32_t g = w - BL_PRODUCER_THREADS / 32;
    const uint32_t px = tx0 + 8u * (g & 1u) + (lane & 7u);
    const uint32_t py = ty0 + 4u * (g >> 1) + (lane >> 3);
    const bool inside = px < P.W && py < P.row1;
    uint32_t *pix = fb_rows + (size_t)(py - P.row0) * P.W + px;
    uint32_t old = 0;
    if (inside) old = *pix;
    float cr = div255((float)((old >> 16) & 0xFFu));
    float cg = div255((float)((old >> 8) & 0xFFu));
    float cb = div255((float)(old & 0xFFu));
    // E7 (alpha byte).  blend() stores the CURRENT fragment's alpha, and euc calls it for every
    // covered pixel, so the byte a pixel ends up with belongs to the LAST entry of the list
    // whose 3-sigma rectangle covers it -- 0 if that fragment was zero.  That entry is found
    // here by walking the list backwards (typically a few dozen entries) while the producers
    // fill the ring; the main loop then only has to handle entries that change RGB.
    const float sx = (float)px + P.sample_off, sy = (float)py + P.sample_off;
    uint32_t last_g = 0xFFFFFFFFu;
    bool found = !inside;
    for (uint32_t end = range.y; end > range.x && __any_sync(0xFFFFFFFFu, !found); end -= min(32u, end - range.x)) {
      const uint32_t cntb = min(32u, end - range.x);
      uint32_t gi = 0;
      float ecx = 0.f, ecy = 0.f, ehx = -1.f, ehy = -1.f;
      if (lane < cntb) {
        gi = __ldg(&inst_vals[end - 1u - lane]);
        const float4 *rp = reinterpret_cast<const float4 *>(recs + gi);
        const float4 a = __ldg(rp), b = __ldg(rp + 1);
        ecx = a.x; ecy = a.y; ehx = b.z; ehy = b.w;
      }
      for (uint32_t k = 0; k < cntb; ++k) {
        const float kx = __shfl_sync(0xFFFFFFFFu, ecx, k), ky = __shfl_sync(0xFFFFFFFFu, ecy, k);
        const float khx = __shfl_sync(0xFFFFFFFFu, ehx, k), khy = __shfl_sync(0xFFFFFFFFu, ehy, k);
        const uint32_t kg = __shfl_sync(0xFFFFFFFFu, gi, k);
        if (!found && fabsf(sx - kx) <= khx && fabsf(sy - ky) <= khy) { found = true; last_g = kg; }
        if (!__any_sync(0xFFFFFFFFu, !found)) break;
      }
    }

    for (uint32_t chunk = 0;; ++chunk) {
      const uint32_t slot = chunk % BL_D;
      mbar_wait(&S.full[g][slot], (chunk / BL_D) & 1u);
      const uint32_t h = S.hdr[g][slot];
      const uint32_t n = h & 0xFFu;
      for (uint32_t e = 0; e < n; ++e) {
        const float *slotp = &S.ring[g][slot][e][0];
        const float al = slotp[lane];
        if (__any_sync(0xFFFFFFFFu, al > 0.0f)) {
          const float4 col = *reinterpret_cast<const float4 *>(slotp + 32);
          if (al > 0.0f) {
            const float om = __fsub_rn(1.0f, al);
            cr = blend_channel(cr, om, __fmul_rn(al, col.x));
            cg = blend_channel(cg, om, __fmul_rn(al, col.y));
            cb = blend_channel(cb, om, __fmul_rn(al, col.z));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[g][slot]);
      if (h & 0x100u) break;
    }

    if (inside && last_g != 0xFFFFFFFFu) {
      // fragment() of the last covering entry for this pixel (pipelines.rs:127-145)
      const float4 *rp = reinterpret_cast<const float4 *>(recs + last_g);
      const float4 a = __ldg(rp), b = __ldg(rp + 1);
      const float dx = sx - a.x, dy = sy - a.y;
      const float q1 = __fmul_rn(__fmul_rn(a.z, dx), dx);
      const float q2 = __fmul_rn(__fmul_rn(b.x, dy), dy);
      const float q3 = __fmul_rn(__fmul_rn(a.w, dx), dy);
      const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(q1, q2)), q3);
      float last_alpha = 0.0f;
      if (!(power > 0.0f)) {
        const float ex = (power >= -87.0f) ? expf_pinned(power) : 0.0f;   // splat_expf flushes below -87
        const float t = fminf(0.99f, __fmul_rn(b.y, ex));
        if (!(t < (1.0f / 255.0f))) last_alpha = t;
      }
      const uint32_t r = (uint32_t)__fmul_rn(cr, 255.0f), gg = (uint32_t)__fmul_rn(cg, 255.0f);
      const uint32_t bl = (uint32_t)__fmul_rn(cb, 255.0f), av = (uint32_t)__fmul_rn(last_alpha, 255.0f);
      *pix = bl | (gg << 8) | (r << 16) | (av << 24);
    }
  }
}

}  // namespace splat
