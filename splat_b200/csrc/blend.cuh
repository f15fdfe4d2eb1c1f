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
//   byte' = trunc_sat(out*255) (round-toward-zero add of 2^23)
// exp() is the pinned "splat_expf v1" operation sequence, restated independently by the CPU checker.
//
// The kernel is bound by the FP32 pipe (about 45 IEEE operations per pixel-Gaussian pair, 3.8e9
// contributing pairs per 1080p frame of the 6.1M scene), not by HBM, so the design minimises
// issue slots and FMA-pipe cycles per pair:
//   * every lane owns TWO pixels (x, y) and (x, y+4) and evaluates them with Blackwell's packed
//     f32x2 instructions (FFMA2 / FMUL2 / FADD2: two IEEE-rounded results per lane per issue
//     slot; measured in tools/microbench: same FLOP rate as scalar, half the instructions);
//   * a 16x16 tile = four 8x8-pixel groups.  Tile lists are extremely skewed (in the 6.1M scene
//     119 of 8160 tiles hold 42% of all tile instances and the longest list is 0.6% of the
//     frame, about what one of the 148 SMs can process in the whole frame time), so a work
//     unit is a tile, half a tile or a single group depending on the list length
//     (unit_order_kernel); units are issued heaviest-first so the long lists start at t=0 and
//     short ones fill the tail;
//   * per group a TEAM: 2, 4 or 8 PRODUCER warps (8 per CTA, shared among the unit's groups)
//     evaluate fragment() -- coverage, power, exp, alpha -- for alternating chunks of 8 list
//     entries and publish the entries that change at least one pixel (two alphas per lane +
//     the colour) into a shared-memory ring; one CONSUMER warp runs only the strictly
//     sequential blend() chain (28 FMA-pipe instructions per entry).  Hand-over through
//     mbarriers (full/empty per ring slot);
//   * the tile's sorted list is staged through shared memory 256 entries at a time by the 8
//     producer warps (one entry per thread), which also compact, per group, the indices of
//     the entries whose 3-sigma rectangle and alpha >= 1/255 ellipse can touch the group.
//
// Exact early termination.  The recurrence cannot be cut short by a transmittance threshold, but
// it has a property that gives the same saving without changing a single bit: for a fixed entry
// (alpha, colour) the map old byte -> new byte is monotone non-decreasing (every step -- the two
// correctly rounded products, the rounded sum, the saturation, the rounded x255, the truncation,
// the correctly rounded /255 -- is monotone in `old`), and so is any composition of such maps.
// Hence if a run of entries sends BOTH byte 0 and byte 255 to the same byte v, it sends every
// byte to v: the pixel's final value is independent of everything before that run and of the
// framebuffer contents.  The kernel therefore composites only a SUFFIX of the tile list, carrying
// two states per pixel (started at 0 and at 255) until they coincide for all pixels of the group
// and one state afterwards; if some pixel has not converged when the list ends the attempt is
// repeated with a 2x longer suffix, and an attempt that reaches the head of the list starts from
// the real framebuffer bytes (the plain algorithm).  On the 6.1M-Gaussian bench scene a pixel is
// covered by ~1,800 contributing entries of which the last ~100 decide its value.
//
// ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even under --fmad=false (seen in SASS), which
// would change roundings.  Every packed product that feeds an addition is therefore written as
// fma(a, b, nz) with nz = (-0.0, -0.0) read from the kernel parameters at run time: a*b + (-0)
// is exactly RN(a*b), and ptxas can neither fold the unknown addend nor fuse an FMA with an add.
#pragma once
#include "common.cuh"

namespace splat {

constexpr int BL_GROUPS = 4;                       // 8x8-pixel groups per CTA (one 16x16 tile)
constexpr int BL_PRODUCER_THREADS = 256;           // 8 producer warps: team = warp>>1, p = warp&1
constexpr int BL_THREADS = BL_PRODUCER_THREADS + 32 * BL_GROUPS;   // + one consumer warp per team
constexpr int BL_BATCH = 256;                      // list entries staged per round
constexpr int BL_CH = 8;                           // list entries per ring chunk
constexpr int BL_SLOTS = 16;                       // ring chunk slots per CTA, split among the unit's groups
#ifndef SPLAT_SUFFIX0
#define SPLAT_SUFFIX0 192
#endif
#ifndef SPLAT_SUFFIX_GROWTH
#define SPLAT_SUFFIX_GROWTH 2
#endif
constexpr uint32_t BL_SUFFIX0 = SPLAT_SUFFIX0;              // list entries composited by the first suffix attempt (r1q/r1r sweeps)
constexpr uint32_t BL_SUFFIX_GROWTH = SPLAT_SUFFIX_GROWTH;  // growth factor per failed attempt
constexpr uint32_t BL_SUFFIX_MIN_LEN = 3 * BL_SUFFIX0 / 2;  // shorter lists are composited whole, straight away

// ---------------------------------------------------------------- packed f32x2 helpers
typedef unsigned long long f32x2;   // two IEEE binary32 values in one 64-bit register pair

SPLAT_DEVINL f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
SPLAT_DEVINL f32x2 pk1(float v) { return pk(v, v); }
SPLAT_DEVINL void upk(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
SPLAT_DEVINL f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// product that does NOT feed an addition (safe to leave as FMUL2)
SPLAT_DEVINL f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// product that feeds an addition: a*b + (-0) == RN(a*b), unfusable (see header)
SPLAT_DEVINL f32x2 mul2x(f32x2 a, f32x2 b, f32x2 nz) { return fma2(a, b, nz); }
SPLAT_DEVINL f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
SPLAT_DEVINL f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
SPLAT_DEVINL f32x2 add2_rz(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// splat_expf v1 on its hot domain [-87, 0] (callers guarantee the domain), scalar and packed.
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
SPLAT_DEVINL void expf2_pinned(f32x2 x, float &e0, float &e1) {
  const f32x2 MAGIC = pk1(12582912.0f);
  const f32x2 tm = fma2(x, pk1(0x1.715476p+0f), MAGIC);
  const f32x2 n = sub2(tm, MAGIC);
  f32x2 r = fma2(n, pk1(-0x1.62e4p-1f), x);
  r = fma2(n, pk1(-0x1.7f7d1cp-20f), r);
  f32x2 p = pk1(0x1.687b46p-10f);
  p = fma2(p, r, pk1(0x1.123bdcp-7f));
  p = fma2(p, r, pk1(0x1.555b5cp-5f));
  p = fma2(p, r, pk1(0x1.55548ep-3f));
  p = fma2(p, r, pk1(0x1.fffff8p-2f));
  p = fma2(p, r, pk1(1.0f));
  p = fma2(p, r, pk1(1.0f));
  float p0, p1, t0, t1;
  upk(p, p0, p1);
  upk(tm, t0, t1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// byte / 255.0f, correctly rounded for every integer 0..255 (checked exhaustively in
// tests/test_host_math.py): q = fma(n, RN(1/255), n * (1/255 - RN(1/255))).
constexpr float DIV255_HI = 0x1.010102p-8f, DIV255_LO = -0x1.fdfdfep-33f;
SPLAT_DEVINL float div255(float n) { return __fmaf_rn(n, DIV255_HI, __fmul_rn(n, DIV255_LO)); }

// One channel of blend(): returns the new channel state (= new byte / 255).
// `as u8` saturates and maps NaN to 0: trunc_sat(out*255) == trunc(sat01(out)*255), because
// out >= 1 gives 255 either way and out <= 0 / NaN give 0; the saturation rides on the FADD.
SPLAT_DEVINL float blend_channel(float c_old, float om, float u) {
  const float out = __saturatef(__fadd_rn(__fmul_rn(om, c_old), u));
  const float v = __fmul_rn(out, 255.0f);
  const float byte = __fsub_rn(__fadd_rz(v, 8388608.0f), 8388608.0f);  // truncate toward zero
  return div255(byte);
}
// Two pixels of one channel.  Products and the truncation chain are packed; the addition is
// issued as two scalar FADD.SAT because f32x2 has no saturating form (same FMA-pipe cycles as
// one FADD2, one more issue slot) -- colours are not clamped by the reference (gaussians.rs:97),
// so out can leave [0,1] and `as u8` saturates.
SPLAT_DEVINL f32x2 blend_channel2(f32x2 c_old, f32x2 om, f32x2 al, float col, f32x2 nz) {
  float t0, t1, u0, u1;
  upk(mul2x(om, c_old, nz), t0, t1);
  upk(mul2x(al, pk1(col), nz), u0, u1);
  const f32x2 out = pk(__saturatef(__fadd_rn(t0, u0)), __saturatef(__fadd_rn(t1, u1)));
  const f32x2 v = mul2x(out, pk1(255.0f), nz);
  const f32x2 n = sub2(add2_rz(v, pk1(8388608.0f)), pk1(8388608.0f));
  return fma2(n, pk1(DIV255_HI), mul2(n, pk1(DIV255_LO)));
}

// ---------------------------------------------------------------- mbarrier / named barrier
#ifndef SPLAT_SLEEP_NS
#define SPLAT_SLEEP_NS 256
#endif
#ifndef SPLAT_SLEEP_NS_CONSUMER
#define SPLAT_SLEEP_NS_CONSUMER 32
#endif
#ifndef SPLAT_TMA_STAGE
// 1: stage each batch of the tile list with per-record cp.async.bulk copies (TMA gather) onto one
// mbarrier; 0: one thread per entry, __ldg + st.shared.  Measured A/B on the bench frame
// (profiles/r2k_ab_tma_staging.txt): the TMA gather is 3% SLOWER here (blend 0.887 vs 0.863 ms) --
// a 48-byte copy per instruction keeps the copy engine no busier than the LSU was, and the batch
// still waits for its slowest record -- so the parity kernel ships the __ldg path; the float
// compositor (blend_float.cuh), which double-buffers whole batches, uses the TMA gather.
#define SPLAT_TMA_STAGE 0
#endif
#ifndef SPLAT_MAX_PPG_LOG
#define SPLAT_MAX_PPG_LOG 3   // at most 2^this producer warps evaluate for one group; the rest only stage
#endif
SPLAT_DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SPLAT_DEVINL void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SPLAT_DEVINL void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
// Blocking wait on an mbarrier phase.  MODE 0: test_wait polling (lowest wake-up latency, but a
// polling warp keeps taking issue slots from the warps that do the work).  MODE 1: try_wait with
// a suspend-time hint (measured on B200: the suspended warp is woken by unrelated barrier
// traffic and re-polls ~10^9 times per frame, 21% of all issued instructions in r1d).
// MODE 2: test_wait, then nanosleep with a short back-off -- a sleeping warp issues nothing.
SPLAT_DEVINL bool mbar_test(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
constexpr uint32_t WD_POLLS = 1u << 22;
constexpr int WD_WORDS = 128;
// wd: WD_WORDS words of mapped pinned host memory (splat_ctx::h_wd): [0] tripped, [1] block,
// [2] thread, [3] tag = role << 28 | slot << 20 | chunk (low 20 bits), [4] parity waited for,
// [5] number of dump words, [8..] the CTA's synchronisation state (see WdDump)
struct WdDump { const uint32_t *words; uint32_t n; };
__device__ __noinline__ void watchdog_trip(uint32_t *wd, uint32_t tag, uint32_t parity, WdDump dump) {
  if (wd && atomicCAS(&wd[0], 0u, 1u) == 0u) {
    wd[1] = blockIdx.x; wd[2] = threadIdx.x; wd[3] = tag; wd[4] = parity;
    const uint32_t n = min(dump.n, (uint32_t)WD_WORDS - 8u);
    wd[5] = n;
    for (uint32_t i = 0; i < n; ++i) wd[8 + i] = dump.words[i];
    __threadfence_system();
  }
  __trap();
}
template <int MODE>
SPLAT_DEVINL void mbar_wait(uint64_t *bar, uint32_t parity, const uint32_t *gword = nullptr, uint32_t *wd = nullptr,
                            uint32_t tag = 0, WdDump dump = WdDump{nullptr, 0u}) {
  if (MODE == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
  } else if (MODE == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
  } else if (MODE == 4) {
    // Poll, and between polls park the warp on the scoreboard: a volatile global load (an L2 hit,
    // several hundred cycles) whose result is folded into the next poll's address.  One issue
    // slot per poll instead of one every few cycles -- nanosleep returns almost immediately
    // here (r1h/r1l: the same ~5e7 polls per frame at 64 ns and at 256 ns).
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t, a;\n\t"
        "mov.b32 a, %0;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [a], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "ld.volatile.global.u32 t, [%2];\n\t"
        "and.b32 t, t, 0;\n\t"
        "add.u32 a, %0, t;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "l"(gword)
        : "memory");
  } else {
    // test_wait, then nanosleep with a short back-off -- a sleeping warp issues nothing.  The poll
    // count doubles as a watchdog: no legitimate wait in this kernel lasts longer than a few hundred
    // microseconds, so after WD_POLLS polls (>= 0.1 s) the warp records who was waiting for what in
    // host-visible memory and traps; the host then reports SPLAT_ERR_CUDA instead of hanging.
    uint32_t polls = 0;
    while (!mbar_test(bar, parity)) {
      __nanosleep(MODE == 2 ? SPLAT_SLEEP_NS : SPLAT_SLEEP_NS_CONSUMER);
      if (++polls > WD_POLLS) watchdog_trip(wd, tag, parity, dump);
    }
  }
}
#ifndef SPLAT_WAIT_CONSUMER
#define SPLAT_WAIT_CONSUMER 3
#endif
#ifndef SPLAT_WAIT_PRODUCER
#define SPLAT_WAIT_PRODUCER 2
#endif
SPLAT_DEVINL void mbar_inval(uint64_t *bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
SPLAT_DEVINL void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// one 1-D bulk copy global -> shared (the TMA engine, no tensor map needed), completion counted in
// bytes on an mbarrier.  16-byte aligned addresses, size a multiple of 16.
SPLAT_DEVINL void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// all threads that take part in a unit (the 8 producer warps + its consumer warps)
SPLAT_DEVINL void unit_sync(uint32_t nthreads) { asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); }
SPLAT_DEVINL void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(BL_PRODUCER_THREADS) : "memory"); }

// ---------------------------------------------------------------- heaviest-first unit order
// A unit is (tile, first group, number of groups): 4 groups = whole tile, 2 = upper / lower half,
// 1 = a single 8x8 group.  unit.y = ngroups | first_group << 8.  Tiles are ordered by descending
// list length (counting sort on 16*log2(len) buckets, single CTA); a tile whose list is longer
// than 1/3200 (1/800) of the frame's instances is split into 2 (4) units, which then run on
// different SMs with 4 (8) producer warps per group.  Empty tiles produce no unit.
SPLAT_DEVINL uint32_t unit_split(uint32_t len, uint32_t t2, uint32_t t4) { return len > t4 ? 4u : (len > t2 ? 2u : 1u); }

// near cut: grow the bounding box (in stripe-local tile coordinates) of the tiles that need the
// complete lists
SPLAT_DEVINL void mark_failed_tile(FrameStatus *status, uint32_t *tile_failed, uint32_t tile, uint32_t tx, uint32_t ty) {
  tile_failed[tile] = 1u;
  atomicMax(&status->fail_ix0, ~tx);
  atomicMax(&status->fail_iy0, ~ty);
  atomicMax(&status->fail_x1, tx);
  atomicMax(&status->fail_y1, ty);
}

__global__ void __launch_bounds__(1024)
unit_order_kernel(const uint2 *__restrict__ ranges, uint32_t T, uint2 *__restrict__ units,
                  uint32_t *__restrict__ n_units, const unsigned long long *__restrict__ n_instances,
                  const uint32_t *__restrict__ far_cnt, FrameStatus *__restrict__ status, uint32_t tiles_x,
                  uint32_t *__restrict__ tile_failed, int only_failed, int no_split) {
  // only_failed: the pass after a near-cut pass blends nothing but the tiles that pass marked --
  // every other tile is final, and one that was composited from the framebuffer bytes must not
  // be composited onto its own output again
  constexpr int NB = 512;
  __shared__ uint32_t hist[NB];
  for (int i = threadIdx.x; i < NB; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const unsigned long long I = *n_instances;
#ifdef SPLAT_NO_SPLIT
  const uint32_t t4 = 0xFFFFFFFFu, t2 = 0xFFFFFFFFu;
  (void)I;
#else
  // no_split: whole tiles only (the float compositor runs one CTA per tile)
  const uint32_t t4 = no_split ? 0xFFFFFFFFu : (uint32_t)max(4096ull, I / 800ull), t2 = no_split ? 0xFFFFFFFFu : (uint32_t)max(2048ull, I / 3200ull);
#endif
  auto bucket = [](uint32_t len) -> uint32_t {
    const int b = (int)(16.0f * __log2f((float)len));   // 0 .. 16*32-1
    return (uint32_t)max(0, NB - 1 - b);
  };
  uint32_t lost = 0;   // near cut: tiles whose whole list was cut away get no unit -- the frame needs the full lists
  for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
    const uint2 r = ranges[t];
    if (only_failed && !tile_failed[t]) continue;
    if (r.y > r.x) atomicAdd(&hist[bucket(r.y - r.x)], unit_split(r.y - r.x, t2, t4));
    else if (far_cnt && far_cnt[t]) {
      lost += 1;
      mark_failed_tile(status, tile_failed, t, t % tiles_x, t / tiles_x);
    }
  }
  if (lost) atomicAdd(&status->n_failed, lost);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < NB; ++i) { const uint32_t c = hist[i]; hist[i] = run; run += c; }
    *n_units = run;
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
    const uint2 r = ranges[t];
    if (r.y <= r.x) continue;
    if (only_failed && !tile_failed[t]) continue;
    const uint32_t nu = unit_split(r.y - r.x, t2, t4), ng = 4u / nu;
    const uint32_t pos = atomicAdd(&hist[bucket(r.y - r.x)], nu);
    for (uint32_t k = 0; k < nu; ++k) units[pos + k] = make_uint2(t, ng | ((k * ng) << 8));
  }
}

// ---------------------------------------------------------------- K5
#ifdef SPLAT_STATS
// instrumented build only (tools/run_stats.sh): work counters of the blend kernel
//  [0] group-entries evaluated by producers  [1] group-entries handed to consumers
//  [2] pixels with alpha > 0                 [3] list entries staged (per unit)
//  [4] candidate pixels (rect + power tests) [5] lanes (pixel pairs) with alpha > 0
__device__ unsigned long long g_blend_stats[8];
#define STAT_ADD(i, v) atomicAdd(&g_blend_stats[i], (unsigned long long)(v))
#endif
#ifdef SPLAT_WD_TRACE
#define WD_TRACE(code, v) do { if (lane == 0) S.trace[w][0] = ((uint32_t)(code) << 28) | ((uint32_t)(v) & 0x0FFFFFFFu); } while (0)
#define WD_COUNT(v) do { if (lane == 0) S.trace[w][1] = (uint32_t)(v); } while (0)
#else
#define WD_TRACE(code, v) do { } while (0)
#define WD_COUNT(v) do { } while (0)
#endif
struct RingEntry {
  float2 al[32];   // per lane: alpha of pixel (x, y) and of pixel (x, y+4); 0 = no change
  float4 col;      // r, g, b (+ the power threshold, unused by the consumer)
};
struct BlendSmem {
  Rec srec[BL_BATCH];                                     // staged records: the tile list's next BL_BATCH entries
  RingEntry ring[BL_SLOTS][BL_CH];                        // chunk slots, BL_SLOTS / ngroups per team
  // ---- the block from here to `trace` is what the watchdog dumps (contiguous on purpose)
  uint64_t full[BL_SLOTS], empty[BL_SLOTS];               // mbarriers
  uint32_t hdr[BL_SLOTS];                                 // entries in the chunk | last << 8
  uint32_t info[8];                                       // tile, ng | g0 << 8, len, start - range.x, whole | truncated << 1, suffix
  uint32_t trace[BL_THREADS / 32][2];                     // -DSPLAT_WD_TRACE: last checkpoint of every warp
  uint32_t wcount[BL_GROUPS][BL_PRODUCER_THREADS / 32];   // per staging warp, per group
  uint8_t list[BL_GROUPS][BL_BATCH];                      // compacted entry indices per group
  uint32_t fail[BL_GROUPS];                               // per team: the suffix attempt did not converge
  uint32_t giveup;                                        // truncated list with a pixel nothing covers
  uint64_t stage_bar;                                     // mbarrier the TMA gather of a batch completes on
};
constexpr size_t BL_SMEM_BYTES = sizeof(BlendSmem);

// fragment() for one pixel against one record (pipelines.rs:127-145), used by the epilogue
SPLAT_DEVINL float fragment_alpha(float sx, float sy, const float4 a, const float4 b) {
  const float dx = sx - a.x, dy = sy - a.y;
  const float q1 = __fmul_rn(__fmul_rn(a.z, dx), dx);
  const float q2 = __fmul_rn(__fmul_rn(b.x, dy), dy);
  const float q3 = __fmul_rn(__fmul_rn(a.w, dx), dy);
  const float power = __fsub_rn(__fmul_rn(-0.5f, __fadd_rn(q1, q2)), q3);
  float alpha = 0.0f;
  if (!(power > 0.0f)) {
    const float ex = (power >= -87.0f) ? expf_pinned(power) : 0.0f;   // splat_expf flushes below -87
    const float t = fminf(0.99f, __fmul_rn(b.y, ex));
    if (!(t < (1.0f / 255.0f))) alpha = t;
  }
  return alpha;
}

__global__ void __launch_bounds__(BL_THREADS, 3)
blend_kernel(const uint2 *__restrict__ ranges, const uint2 *__restrict__ units,
             const uint32_t *__restrict__ n_units, const uint32_t *__restrict__ inst_vals, const Rec *__restrict__ recs,
             uint32_t *__restrict__ fb_rows, const __grid_constant__ FrameParams P,
             const uint32_t *__restrict__ far_cnt, FrameStatus *__restrict__ status, uint32_t *__restrict__ tile_failed,
             uint32_t *__restrict__ wd) {
  extern __shared__ __align__(16) unsigned char blend_smem_raw[];
  BlendSmem &S = *reinterpret_cast<BlendSmem *>(blend_smem_raw);

  if (blockIdx.x >= *n_units) return;   // tiles nothing touches have no unit: pixels stay as they are
  const uint2 unit = units[blockIdx.x];
  const uint32_t tile = unit.x, ng = unit.y & 0xFFu, g0 = unit.y >> 8;     // ng in {1, 2, 4}
  const uint32_t lg_ng = ng >> 1;                                          // log2(ng)
  const uint32_t ppg_log = 3u - lg_ng, ppg_mask = (1u << ppg_log) - 1u;    // producers per group
  const uint32_t d_log = 4u - lg_ng, d_mask = (1u << d_log) - 1u;          // chunk slots per group
  const uint2 range = ranges[tile];
  const uint32_t tile_x = tile % P.tiles_x, tile_y = tile / P.tiles_x;
  const uint32_t tx0 = tile_x * TILE, ty0 = (P.tile_y0 + tile_y) * TILE;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  // The mbarriers are initialised ONCE per CTA and made visible by a CTA-wide barrier in which every
  // thread takes part, before the spare warps of a split unit leave.  Later suffix attempts do not
  // touch them again: chunk numbers -- hence slots, owners and phase parities -- simply run on
  // (cbase).  Invalidating and re-initialising the barriers between attempts, with other warps
  // already polling for the next attempt's barrier, dead-locked about one unit in 10^4 on frames
  // with thousands of short truncated lists (round-2 hunt, profiles/r2_near_cut_hang.txt).
  if (tid < BL_SLOTS) {
    mbar_init(&S.full[tid], 1);
    mbar_init(&S.empty[tid], 1);
  }
  if (tid == 0) {
    S.giveup = 0;
    mbar_init(&S.stage_bar, 1);
  }
  __syncthreads();
  [[maybe_unused]] uint32_t stage_par = 0;   // phase parity of stage_bar: one phase per staged batch, all attempts
  if (w >= BL_PRODUCER_THREADS / 32 + ng) return;   // split units: the spare consumer warps have nothing to do
  const uint32_t nsync = BL_PRODUCER_THREADS + 32u * ng;   // threads that take part in this unit
  const f32x2 NZ = P.nz2;
  const uint32_t len = range.y - range.x;
  uint32_t suffix = BL_SUFFIX0;
  // near cut (bin.cuh): Gaussians farther than everything in this list were not binned this pass
  // and some of them touch this tile.  The list is then only a suffix of the real one: it can
  // never be composited "exactly from its head"; if its pixels do not converge the frame is
  // repeated with the full lists (splat_api.cu).
  const bool truncated = far_cnt != nullptr && far_cnt[tile] != 0u;
  bool gave_up = false;
  uint32_t cbase = 0;   // chunks this thread's team handed over in earlier attempts

  // ---- suffix attempts (see "Exact early termination" in the header) ----
  // Attempt k composites only the last `suffix` list entries, starting every pixel from BOTH
  // extreme states (byte 0 and byte 255).  Each entry's byte -> byte map is monotone, so once
  // the two trajectories of a pixel coincide the result no longer depends on anything earlier
  // in the list -- or on the framebuffer contents.  If some pixel has not converged when the
  // list ends, the attempt is repeated with a 2x longer suffix; an attempt that reaches the
  // head of the list starts from the real framebuffer bytes and is exact by construction.
  uint32_t start = range.x;
  bool exact = true;
  // consumer state that survives the attempt loop
  f32x2 cr = 0, cg = 0, cb = 0;
  uint32_t last0 = 0xFFFFFFFFu, last1 = 0xFFFFFFFFu;
  bool alpha_done = false;

  for (;;) {
  const bool whole = (len <= BL_SUFFIX_MIN_LEN) || (suffix >= len);
  exact = whole && !truncated;
  start = whole ? range.x : range.y - suffix;
  if (tid == 0) {
    S.info[0] = tile; S.info[1] = ng | (g0 << 8); S.info[2] = len; S.info[3] = start - range.x;
    S.info[4] = (whole ? 1u : 0u) | (truncated ? 2u : 0u); S.info[5] = suffix;
  }
  const WdDump wdd{reinterpret_cast<const uint32_t *>(&S.full[0]),
                   (uint32_t)((sizeof(S.full) + sizeof(S.empty) + sizeof(S.hdr) + sizeof(S.info) + sizeof(S.trace)) / 4)};
  unit_sync(nsync);

  if (w < BL_PRODUCER_THREADS / 32) {
    // ============================== PRODUCERS ==============================
    const uint32_t gl = w >> ppg_log, p = w & ppg_mask, g = g0 + gl;   // local team, rank in team, group
    // SPLAT_MAX_PPG_LOG < 3 lets the surplus producer warps of a split unit only stage (measured
    // r1e: slower -- the heavy units are the critical path and want every producer they can get).
    const uint32_t ev_mask = min(ppg_mask, (1u << SPLAT_MAX_PPG_LOG) - 1u);
    const bool evaluates = p <= ev_mask;
    const uint32_t unit_mask = ((1u << ng) - 1u) << g0;
    const float sx = (float)(tx0 + 8u * (g & 1u) + (lane & 7u)) + P.sample_off;
    const float sy0 = (float)(ty0 + 8u * (g >> 1) + (lane >> 3)) + P.sample_off;
    const float sy1 = (float)(ty0 + 8u * (g >> 1) + (lane >> 3) + 4u) + P.sample_off;
    const f32x2 sy2 = pk(sy0, sy1);
    // group sample intervals for the overlap test (same (float)p + off as sx/sy)
    float sxl[2], sxh[2], syl[2], syh[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      sxl[q] = (float)(tx0 + 8u * q) + P.sample_off;
      sxh[q] = (float)(tx0 + 8u * q + 7u) + P.sample_off;
      syl[q] = (float)(ty0 + 8u * q) + P.sample_off;
      syh[q] = (float)(ty0 + 8u * q + 7u) + P.sample_off;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t seq = cbase * BL_CH;   // list entries of this team handed out so far, counted on from the earlier
                                    // attempts in whole chunks (both producers count alike)
    uint32_t nout = 0;   // entries written into the chunk this producer is filling
#ifdef SPLAT_STATS
    uint32_t st_eval = 0, st_out = 0, st_pix = 0, st_cand = 0, st_lanes = 0;
#endif

    for (uint32_t base = start; base < range.y; base += BL_BATCH) {
      const uint32_t nb = min((uint32_t)BL_BATCH, range.y - base);
      WD_TRACE(1, base - start);
      producers_sync();   // previous batch no longer read by any producer
      uint32_t bits = 0;
#if SPLAT_TMA_STAGE
      // TMA gather of the batch: thread j reads list index j and issues ONE bulk copy of that 48-byte
      // record into shared-memory slot j; all copies complete on one mbarrier (expect_tx = 48 * nb)
      if (tid == 0) mbar_expect_tx(&S.stage_bar, nb * (uint32_t)sizeof(Rec));
      if (tid < nb) tma_load_1d(&S.srec[tid], recs + __ldg(&inst_vals[base + tid]), (uint32_t)sizeof(Rec), &S.stage_bar);
      mbar_wait<SPLAT_WAIT_PRODUCER>(&S.stage_bar, stage_par, n_units, wd, (6u << 28) | ((base - start) & 0xFFFFFu), wdd);
      stage_par ^= 1u;
#endif
      if (tid < nb) {
#if SPLAT_TMA_STAGE
        const float4 a = S.srec[tid].a, b = S.srec[tid].b, c = S.srec[tid].c;
#else
        const uint32_t gi = __ldg(&inst_vals[base + tid]);
        const float4 *rp = reinterpret_cast<const float4 *>(recs + gi);
        const float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
        S.srec[tid].a = a; S.srec[tid].b = b; S.srec[tid].c = c;
#endif
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
        bits &= unit_mask;
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
        rank[q] = 0;
        if (!((unit_mask >> q) & 1u)) continue;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (bits >> q) & 1u);
        rank[q] = __popc(bal & lt_mask);
        if (lane == 0) S.wcount[q][w] = __popc(bal);
      }
      producers_sync();
      uint32_t n_mine = 0;
#pragma unroll
      for (int q = 0; q < BL_GROUPS; ++q) {
        if (!((unit_mask >> q) & 1u)) continue;
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

      // chunks of BL_CH consecutive list entries of the team are dealt round-robin to its
      // evaluating producers: chunk c belongs to producer (c & ev_mask)
      const uint32_t s_end = seq + n_mine;
      WD_TRACE(2, s_end);
      WD_COUNT(seq | (n_mine << 16));
      uint32_t chunk = seq / BL_CH;
      chunk += (p - chunk) & ev_mask;                      // first chunk >= seq/BL_CH owned by p
      for (; evaluates && chunk * BL_CH < s_end; chunk += ev_mask + 1u) {
        uint32_t s = max(seq, chunk * BL_CH);
        const uint32_t chunk_end = min(s_end, (chunk + 1) * BL_CH);
        const uint32_t slot = (gl << d_log) + (chunk & d_mask);
        if (s % BL_CH == 0) {
          WD_TRACE(3, chunk);
          mbar_wait<SPLAT_WAIT_PRODUCER>(&S.empty[slot], ((chunk >> d_log) & 1u) ^ 1u, n_units, wd,
                                         (1u << 28) | (slot << 20) | (chunk & 0xFFFFFu), wdd);
          nout = 0;
        }
        RingEntry *slotp = &S.ring[slot][nout];
#pragma unroll 2
        for (; s < chunk_end; ++s) {
          const uint32_t j = S.list[g][s - seq];
          const float4 a = S.srec[j].a, b = S.srec[j].b, c = S.srec[j].c;
          const float dx = sx - a.x;
          const f32x2 dy2 = sub2(sy2, pk1(a.y));
          float dy0, dy1;
          upk(dy2, dy0, dy1);
          const bool inx = fabsf(dx) <= b.z;
          // pipelines.rs:134, left to right, no FMA: (A*dx)*dx, (C*dy)*dy, (B*dx)*dy
          const float adx = __fmul_rn(a.z, dx), bdx = __fmul_rn(a.w, dx);
          const float q1 = __fmul_rn(adx, dx);
          const f32x2 q2 = mul2x(mul2(pk1(b.x), dy2), dy2, NZ);
          const f32x2 q3 = mul2x(pk1(bdx), dy2, NZ);
          const f32x2 pw2 = sub2(mul2x(add2(pk1(q1), q2), pk1(-0.5f), NZ), q3);
          float pw0, pw1;
          upk(pw2, pw0, pw1);
          const bool cand0 = inx && (fabsf(dy0) <= b.w) && !(pw0 > 0.0f) && (pw0 >= c.w);
          const bool cand1 = inx && (fabsf(dy1) <= b.w) && !(pw1 > 0.0f) && (pw1 >= c.w);
          // straight-line on purpose (no warp-uniform early-outs): lets the compiler interleave
          // the dependent chains of consecutive entries.  Non-candidate lanes compute a garbage
          // exp that the selects below discard.
          float e0, e1;
          expf2_pinned(pw2, e0, e1);
          const float al0 = fminf(0.99f, __fmul_rn(b.y, e0));     // pipelines.rs:139
          const float al1 = fminf(0.99f, __fmul_rn(b.y, e1));
          // zero fragment: RGB unchanged (alpha byte: see consumer)         pipelines.rs:140
          const float v0 = (cand0 && !(al0 < (1.0f / 255.0f))) ? al0 : 0.0f;
          const float v1 = (cand1 && !(al1 < (1.0f / 255.0f))) ? al1 : 0.0f;
          slotp->al[lane] = make_float2(v0, v1);
          if (lane == 0) slotp->col = c;
          // entries that change no pixel of the group are overwritten by the next one
          const uint32_t adv = __any_sync(0xFFFFFFFFu, (v0 > 0.0f) || (v1 > 0.0f)) ? 1u : 0u;
#ifdef SPLAT_STATS
          st_eval += 1; st_out += adv;
          st_pix += (v0 > 0.0f) + (v1 > 0.0f); st_cand += (uint32_t)cand0 + (uint32_t)cand1;
          st_lanes += ((v0 > 0.0f) || (v1 > 0.0f)) ? 1u : 0u;
#endif
          slotp += adv;
          nout += adv;
        }
        if (chunk_end % BL_CH == 0) {
          __syncwarp();
          if (lane == 0) {
            S.hdr[slot] = nout;
            mbar_arrive(&S.full[slot]);
          }
          WD_TRACE(4, chunk);
        }
      }
      seq = s_end;
#ifdef SPLAT_STATS
      if (tid == 0) STAT_ADD(3, nb);
#endif
    }
#ifdef SPLAT_STATS
    {
      uint32_t a = st_pix, b = st_cand, c = st_lanes;
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
      }
      if (lane == 0) { STAT_ADD(0, st_eval); STAT_ADD(1, st_out); STAT_ADD(2, a); STAT_ADD(4, b); STAT_ADD(5, c); }
    }
#endif
    // terminator: the partial last chunk, or an empty extra chunk, carries the `last` flag
    {
      const uint32_t chunk = seq / BL_CH, rem = seq % BL_CH;
      if (evaluates && (chunk & ev_mask) == p) {
        const uint32_t slot = (gl << d_log) + (chunk & d_mask);
        if (rem == 0) {
          WD_TRACE(5, chunk);
          mbar_wait<SPLAT_WAIT_PRODUCER>(&S.empty[slot], ((chunk >> d_log) & 1u) ^ 1u, n_units, wd,
                                         (2u << 28) | (slot << 20) | (chunk & 0xFFFFFu), wdd);
          nout = 0;
        }
        __syncwarp();
        if (lane == 0) {
          S.hdr[slot] = nout | 0x100u;
          mbar_arrive(&S.full[slot]);
        }
      }
      WD_TRACE(6, seq);
      cbase = chunk + 1u;
    }
  } else {
    // ============================== CONSUMERS ==============================
    const uint32_t gl = w - BL_PRODUCER_THREADS / 32;
    const uint32_t g = g0 + gl;
    const uint32_t px = tx0 + 8u * (g & 1u) + (lane & 7u);
    const uint32_t py0 = ty0 + 8u * (g >> 1) + (lane >> 3), py1 = py0 + 4u;
    const bool inside0 = px < P.W && py0 < P.row1, inside1 = px < P.W && py1 < P.row1;
    const float sx = (float)px + P.sample_off;
    const float sy0 = (float)py0 + P.sample_off, sy1 = (float)py1 + P.sample_off;
    if (!alpha_done) {
      // E7 (alpha byte).  blend() stores the CURRENT fragment's alpha, and euc calls it for every
      // covered pixel, so the byte a pixel ends up with belongs to the LAST entry of the list
      // whose 3-sigma rectangle covers it -- 0 if that fragment was zero.  That entry is found
      // here by walking the list backwards (typically a few dozen entries) while the producers
      // fill the ring; the main loop then only has to handle entries that change RGB.
      alpha_done = true;
      bool found0 = !inside0, found1 = !inside1;
      for (uint32_t end = range.y; end > range.x && __any_sync(0xFFFFFFFFu, !(found0 && found1));
           end -= min(32u, end - range.x)) {
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
          const bool cx_in = fabsf(sx - kx) <= khx;
          if (!found0 && cx_in && fabsf(sy0 - ky) <= khy) { found0 = true; last0 = kg; }
          if (!found1 && cx_in && fabsf(sy1 - ky) <= khy) { found1 = true; last1 = kg; }
          if (!__any_sync(0xFFFFFFFFu, !(found0 && found1))) break;
        }
      }
    }
    // pixels whose value matters: inside the image and covered by at least one quad
    const bool need0 = inside0 && last0 != 0xFFFFFFFFu, need1 = inside1 && last1 != 0xFFFFFFFFu;

    // state: lo trajectory in cr/cg/cb, hi trajectory in hr/hg/hb (only while `dual`)
    f32x2 hr, hg, hb;
    bool dual = !exact;
    if (exact) {
      const uint32_t *pix0 = fb_rows + (size_t)(py0 - P.row0) * P.W + px;
      uint32_t old0 = 0, old1 = 0;
      if (inside0) old0 = *pix0;
      if (inside1) old1 = *(pix0 + (size_t)4u * P.W);
      cr = pk(div255((float)((old0 >> 16) & 0xFFu)), div255((float)((old1 >> 16) & 0xFFu)));
      cg = pk(div255((float)((old0 >> 8) & 0xFFu)), div255((float)((old1 >> 8) & 0xFFu)));
      cb = pk(div255((float)(old0 & 0xFFu)), div255((float)(old1 & 0xFFu)));
      hr = cr; hg = cg; hb = cb;
    } else {
      cr = cg = cb = pk1(0.0f);            // byte 0
      hr = hg = hb = pk1(1.0f);            // byte 255: div255(255) == 1
    }

    for (uint32_t chunk = cbase;; ++chunk) {
      const uint32_t slot = (gl << d_log) + (chunk & d_mask);
      WD_TRACE(7, chunk);
      mbar_wait<SPLAT_WAIT_CONSUMER>(&S.full[slot], (chunk >> d_log) & 1u, n_units, wd,
                                     (3u << 28) | (slot << 20) | (chunk & 0xFFFFFu), wdd);
      const uint32_t h = S.hdr[slot];
      const uint32_t n = h & 0xFFu;
      const RingEntry *ep = &S.ring[slot][0];
      if (dual) {
#pragma unroll 2
        for (uint32_t e = 0; e < n; ++e, ++ep) {
          const float2 al = ep->al[lane];
          const float4 col = ep->col;
          const f32x2 al2 = pk(al.x, al.y);
          const f32x2 om = sub2(pk1(1.0f), al2);
          cr = blend_channel2(cr, om, al2, col.x, NZ);
          hr = blend_channel2(hr, om, al2, col.x, NZ);
          cg = blend_channel2(cg, om, al2, col.y, NZ);
          hg = blend_channel2(hg, om, al2, col.y, NZ);
          cb = blend_channel2(cb, om, al2, col.z, NZ);
          hb = blend_channel2(hb, om, al2, col.z, NZ);
        }
        float l0, l1, u0, u1;
        bool m0 = true, m1 = true;
        upk(cr, l0, l1); upk(hr, u0, u1); m0 = m0 && (l0 == u0); m1 = m1 && (l1 == u1);
        upk(cg, l0, l1); upk(hg, u0, u1); m0 = m0 && (l0 == u0); m1 = m1 && (l1 == u1);
        upk(cb, l0, l1); upk(hb, u0, u1); m0 = m0 && (l0 == u0); m1 = m1 && (l1 == u1);
        if (__all_sync(0xFFFFFFFFu, (m0 || !need0) && (m1 || !need1))) dual = false;
      } else {
        for (uint32_t e = 0; e < n; ++e, ++ep) {
          const float2 al = ep->al[lane];
          const float4 col = ep->col;
          const f32x2 al2 = pk(al.x, al.y);
          const f32x2 om = sub2(pk1(1.0f), al2);
          // alpha == 0 is a natural no-op: out = 1*c + 0 and (x/255)*255 truncates back to x
          cr = blend_channel2(cr, om, al2, col.x, NZ);
          cg = blend_channel2(cg, om, al2, col.y, NZ);
          cb = blend_channel2(cb, om, al2, col.z, NZ);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[slot]);
      if (h & 0x100u) { cbase = chunk + 1u; break; }
    }
    // truncated list: an inside pixel that none of its quads covers may be covered by a cut one
    const bool unknown = truncated && __any_sync(0xFFFFFFFFu, (inside0 && last0 == 0xFFFFFFFFu) ||
                                                                 (inside1 && last1 == 0xFFFFFFFFu));
    if (lane == 0) {
      S.fail[gl] = (dual || unknown) ? 1u : 0u;
      if (unknown) S.giveup = 1u;
    }
#ifdef SPLAT_STATS
    if (lane == 0) { STAT_ADD(6, 1); STAT_ADD(7, dual ? 1 : 0); }
#endif
  }

  unit_sync(nsync);
  uint32_t any_fail = 0;
  for (uint32_t q = 0; q < ng; ++q) any_fail |= S.fail[q];
  if (!any_fail) break;
  if (whole || S.giveup) { gave_up = true; break; }     // only possible for a truncated list
  suffix = (suffix > 0x10000000u) ? 0xFFFFFFFFu : suffix * BL_SUFFIX_GROWTH;
  }   // attempts

  if (gave_up && tid == 0) {
    atomicAdd(&status->n_failed, 1u);
    mark_failed_tile(status, tile_failed, tile, tile_x, tile_y);
  }
  if (w >= BL_PRODUCER_THREADS / 32 && !(gave_up && S.fail[w - BL_PRODUCER_THREADS / 32])) {
    // epilogue: pack RGB, resolve the alpha byte (E7), write the pixel
    // (a group of a given-up unit that did converge still writes: its values are final)
    const uint32_t g = g0 + (w - BL_PRODUCER_THREADS / 32);
    const uint32_t px = tx0 + 8u * (g & 1u) + (lane & 7u);
    const uint32_t py0 = ty0 + 8u * (g >> 1) + (lane >> 3), py1 = py0 + 4u;
    const bool inside0 = px < P.W && py0 < P.row1, inside1 = px < P.W && py1 < P.row1;
    const float sx = (float)px + P.sample_off;
    const float sy0 = (float)py0 + P.sample_off, sy1 = (float)py1 + P.sample_off;
    uint32_t *pix0 = fb_rows + (size_t)(py0 - P.row0) * P.W + px;
    uint32_t *pix1 = pix0 + (size_t)4u * P.W;
    float r0, r1, g0f, g1f, b0, b1;
    upk(cr, r0, r1); upk(cg, g0f, g1f); upk(cb, b0, b1);
    if (inside0 && last0 != 0xFFFFFFFFu) {
      // fragment() of the last covering entry for this pixel (pipelines.rs:127-145)
      const float4 *rp = reinterpret_cast<const float4 *>(recs + last0);
      const float la = fragment_alpha(sx, sy0, __ldg(rp), __ldg(rp + 1));
      const uint32_t r = (uint32_t)__fmul_rn(r0, 255.0f), gg = (uint32_t)__fmul_rn(g0f, 255.0f);
      const uint32_t bl = (uint32_t)__fmul_rn(b0, 255.0f), av = (uint32_t)__fmul_rn(la, 255.0f);
      *pix0 = bl | (gg << 8) | (r << 16) | (av << 24);
    }
    if (inside1 && last1 != 0xFFFFFFFFu) {
      const float4 *rp = reinterpret_cast<const float4 *>(recs + last1);
      const float la = fragment_alpha(sx, sy1, __ldg(rp), __ldg(rp + 1));
      const uint32_t r = (uint32_t)__fmul_rn(r1, 255.0f), gg = (uint32_t)__fmul_rn(g1f, 255.0f);
      const uint32_t bl = (uint32_t)__fmul_rn(b1, 255.0f), av = (uint32_t)__fmul_rn(la, 255.0f);
      *pix1 = bl | (gg << 8) | (r << 16) | (av << 24);
    }
  }
}

}  // namespace splat
