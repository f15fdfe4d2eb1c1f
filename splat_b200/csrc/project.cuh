// project.cuh -- K0 scene packing (once per upload) and K1 per-Gaussian projection (per frame).
//
// K1 replaces the reference's vertex stage, which euc calls six times per Gaussian:
//   vertex()                 pipelines.rs:96-125 (Pipeline01), :184-213 (Pipeline02)
//   gaussian_vertex_shader   pipelines.rs:17-51
//   project_cov3d_to_screen  gaussians.rs:114-161, :473-522
//   eval_spherical_harmonics gaussians.rs:41-99 (sh_dim = 15: degrees 0..2)
// plus the depth computation of sort_gaussians (gaussians.rs:297-303).
//
// All arithmetic is IEEE binary32 in the reference's evaluation order (nalgebra 0.32.3 gemv
// order: ((a0*b0 + a1*b1) + a2*b2), no FMA).  The file is compiled with --fmad=false and
// IEEE division / square root, so the source order below is the evaluation order.  Products
// with the structural zeros of J and diag(scale^2) are dropped: x + (+-0) == x, so the
// results are value-identical to the full 3x3 products the reference performs.
#pragma once
#include "common.cuh"

namespace splat {

// ---------------------------------------------------------------- K0: pack + cov3d
// compute_cov3d, gaussians.rs:101-113 / :446-462:  R * diag(s^2) * R^T with R from the
// normalised quaternion (UnitQuaternion::from_quaternion -> to_rotation_matrix).
SPLAT_DEVINL void cov3d_from_rot_scale(const float4 q /* i,j,k,w */, const float s0, const float s1,
                                       const float s2, float C[9]) {
  float i = q.x, j = q.y, k = q.z, w = q.w;
  float a = i * i, b = j * j, c = k * k, d = w * w;   // nalgebra 4-vector dot: (x0y0+x2y2)+(x1y1+x3y3)
  a += c;
  b += d;
  const float nrm = __fsqrt_rn(a + b);
  i = __fdiv_rn(i, nrm); j = __fdiv_rn(j, nrm); k = __fdiv_rn(k, nrm); w = __fdiv_rn(w, nrm);
  const float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
  const float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f;
  const float ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
  const float R[3][3] = {{ww + ii - jj - kk, ij - wk, wj + ik},
                         {wk + ij, ww - ii + jj - kk, jk - wi},
                         {ik - wj, wi + jk, ww - ii - jj + kk}};
  const float s2v[3] = {s0 * s0, s1 * s1, s2 * s2};
  float RS[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cidx = 0; cidx < 3; ++cidx) RS[r][cidx] = R[r][cidx] * s2v[cidx];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int qd = 0; qd < 3; ++qd)
      C[r * 3 + qd] = (RS[r][0] * R[qd][0] + RS[r][1] * R[qd][1]) + RS[r][2] * R[qd][2];
}

// One thread per Gaussian.  Inputs are the raw GaussianList arrays (gaussians.rs:408-416)
// already on the device; output is the 10-plane float4 scene (common.cuh).
__global__ void __launch_bounds__(256)
pack_scene_kernel(const float4 *__restrict__ pos4, const float *__restrict__ scale3,
                  const float *__restrict__ opacity, const float4 *__restrict__ rot,
                  const float *__restrict__ sh48, float4 *__restrict__ scene, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pos4[i];
  float f[36];
  cov3d_from_rot_scale(rot[i], scale3[3 * (size_t)i + 0], scale3[3 * (size_t)i + 1],
                       scale3[3 * (size_t)i + 2], f);
  const float *sh = sh48 + 48 * (size_t)i;
#pragma unroll
  for (int k = 0; k < 27; ++k) f[9 + k] = sh[k];
  scene[i] = make_float4(p.x, p.y, p.z, opacity[i]);
#pragma unroll
  for (int k = 0; k < 9; ++k)
    scene[(size_t)(k + 1) * n + i] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
}

// AoS variant for Pipeline01's Vec<Gaussian>: 59 floats per Gaussian (include/splat.h).
__global__ void __launch_bounds__(256)
pack_scene_aos_kernel(const float *__restrict__ g59, float4 *__restrict__ scene, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *g = g59 + 59 * (size_t)i;
  float f[36];
  cov3d_from_rot_scale(make_float4(g[7], g[8], g[9], g[10]), g[3], g[4], g[5], f);
#pragma unroll
  for (int k = 0; k < 27; ++k) f[9 + k] = g[11 + k];
  scene[i] = make_float4(g[0], g[1], g[2], g[6]);
#pragma unroll
  for (int k = 0; k < 9; ++k)
    scene[(size_t)(k + 1) * n + i] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
}

// ---------------------------------------------------------------- K1: project
SPLAT_DEVINL bool finitef(float x) { return fabsf(x) <= 3.402823466e38f; }

SPLAT_DEVINL uint32_t depth_key(float z) {
  // ascending order of f32 as ascending u32; -0.0 and +0.0 must compare equal (Rust partial_cmp)
  z = z + 0.0f;
  const uint32_t u = __float_as_uint(z);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

SPLAT_DEVINL int clamp_d2i(double v, int lo, int hi) {
  if (!(v > (double)lo)) return lo;
  if (v > (double)hi) return hi;
  return (int)v;
}

// Conservative inclusive tile rectangle of the 3-sigma quad inside the stripe; the exact
// per-pixel coverage test is repeated in blend_kernel, so this only has to be a superset.
struct TileRect {
  uint16_t x0, y0, x1, y1;   // tile coordinates; y relative to the stripe's first tile row
  SPLAT_DEVINL uint32_t count() const { return (x1 >= x0 && y1 >= y0) ? (uint32_t)(x1 - x0 + 1) * (y1 - y0 + 1) : 0u; }
};
static_assert(sizeof(TileRect) == 8, "TileRect is packed as uint2");

// One thread per Gaussian.  Geometry phase: 4 coalesced float4 loads (position/opacity, cov3d),
// view transform, cov2d, conic, 3-sigma extents, z clip, tile rectangle.  Colour phase: 6 more
// float4 loads (SH degrees 0..2) and the SH evaluation.  Out: 48 B record + 4 B key + 4 B index
// + 8 B rect + 4 B tile count.  Algorithmic bytes: 160 in + 68 out = 228 B per Gaussian.
//
// Stripe renders (multi-GPU, P.stripe_cull): a Gaussian whose quad cannot touch this rank's
// stripe is dropped after the geometry phase -- no SH loads, key = KEY_CULLED -- and
// block_kept[] counts the survivors per CTA so that compact_pairs_kernel can squeeze the
// (key, index) pairs before the depth sort: each rank then sorts only what its stripe sees.
__global__ void __launch_bounds__(256)
project_kernel(const float4 *__restrict__ scene, const __grid_constant__ FrameParams P,
               Rec *__restrict__ recs, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
               uint2 *__restrict__ rects, uint32_t *__restrict__ tcnt, uint32_t *__restrict__ block_kept) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = P.n;
  const bool valid = i < n;
  const uint32_t ii = valid ? i : n - 1u;     // out-of-range threads recompute the last Gaussian, store nothing
  const float4 p0 = __ldg(&scene[ii]);
  float f[36];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 v = __ldg(&scene[(size_t)(k + 1) * n + ii]);
    f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
  }
  if (!P.stripe_cull) {   // full-frame renders need the colour of (almost) everything: all loads in flight at once
#pragma unroll
    for (int k = 3; k < 9; ++k) {
      const float4 v = __ldg(&scene[(size_t)(k + 1) * n + ii]);
      f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
    }
  }
  const float *C3 = f;        // cov3d row-major (f[0..8]); f[9..11] = sh[0..2]
  const float *sh = f + 9;    // sh[0..26]
  const float *V = P.view;

  // gaussians.rs:119-121  pos_cam = view * (x, y, z, 1)
  float pc[4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    pc[r] = ((V[0 * 4 + r] * p0.x + V[1 * 4 + r] * p0.y) + V[2 * 4 + r] * p0.z) + V[3 * 4 + r];
  const float zv = pc[2];

  // gaussians.rs:143-151.  J = [[f/tz,0,*],[0,f/tz,*],[0,0,0]]; only T's first two columns
  // reach the kept 2x2 block: T[k][c] = W[k][c] * J[c][c], W[k][c] = view(c,k) = V[k*4+c].
  const float jd = __fdiv_rn(P.focal, zv);
  float T[3][2];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T[k][0] = V[k * 4 + 0] * jd;
    T[k][1] = V[k * 4 + 1] * jd;
  }
  // X = T^T * cov3d^T : X[r][k] = (T[0][r]*C3[k][0] + T[1][r]*C3[k][1]) + T[2][r]*C3[k][2]
  float X[2][3];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      X[r][k] = (T[0][r] * C3[k * 3 + 0] + T[1][r] * C3[k * 3 + 1]) + T[2][r] * C3[k * 3 + 2];
  // cov = X * T : cov[r][c] = (X[r][0]*T[0][c] + X[r][1]*T[1][c]) + X[r][2]*T[2][c]
  float cv[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) cv[r][c] = (X[r][0] * T[0][c] + X[r][1] * T[1][c]) + X[r][2] * T[2][c];
  const float m11 = cv[0][0] + P.lowpass, m12 = cv[0][1], m21 = cv[1][0], m22 = cv[1][1] + P.lowpass;

  // pipelines.rs:21-26  2x2 inverse by division, 3-sigma half extents
  const float det = m11 * m22 - m21 * m12;
  const float cA = __fdiv_rn(m22, det), cB = __fdiv_rn(-m12, det), cC = __fdiv_rn(m11, det);
  const float hx = 3.0f * __fsqrt_rn(m11), hy = 3.0f * __fsqrt_rn(m22);

  // pipelines.rs:36-42  centre in NDC: proj * pos_cam, then / w
  const float *Pm = P.proj;
  float ps[4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    ps[r] = ((Pm[0 * 4 + r] * pc[0] + Pm[1 * 4 + r] * pc[1]) + Pm[2 * 4 + r] * pc[2]) + Pm[3 * 4 + r] * pc[3];
  const float ndx = __fdiv_rn(ps[0], ps[3]), ndy = __fdiv_rn(ps[1], ps[3]), ndz = __fdiv_rn(ps[2], ps[3]);
  const float cxp = (ndx * 0.5f + 0.5f) * (float)P.W;
  const float cyp = ((P.ysign * ndy) * 0.5f + 0.5f) * (float)P.H;
  const float op = p0.w;

  // visibility: euc's z clip on the centre (all four corners share z) + degeneracy guard
  // (everything except the colour, which the second phase adds)
  bool zok;
  if (P.zclip_mode == 0) zok = (ndz >= 0.0f && ndz < 1.0f);
  else if (P.zclip_mode == 1) zok = (ndz >= -1.0f && ndz < 1.0f);
  else zok = true;
  const bool visg = zok && (det != 0.0f) && finitef(zv) && finitef(cA) && finitef(cB) && finitef(cC) &&
                    finitef(hx) && finitef(hy) && finitef(op) && finitef(cxp) && finitef(cyp) &&
                    // degeneracy guard (keeps `power` finite for every on-screen pixel)
                    fabsf(cA) <= 1e18f && fabsf(cB) <= 1e18f && fabsf(cC) <= 1e18f &&
                    fabsf(cxp) <= 1e9f && fabsf(cyp) <= 1e9f;

  TileRect tr;
  tr.x0 = 1; tr.x1 = 0; tr.y0 = 1; tr.y1 = 0;   // empty
  if (visg) {
    const double off = (double)P.sample_off;
    const double slx = 1.0 + 1e-6 * (fabs((double)cxp) + (double)hx);
    const double sly = 1.0 + 1e-6 * (fabs((double)cyp) + (double)hy);
    const int px0 = clamp_d2i(floor((double)cxp - (double)hx - off - slx), 0, (int)P.W);
    const int px1 = clamp_d2i(ceil((double)cxp + (double)hx - off + slx), -1, (int)P.W - 1);
    const int py0 = clamp_d2i(floor((double)cyp - (double)hy - off - sly), (int)P.row0, (int)P.row1);
    const int py1 = clamp_d2i(ceil((double)cyp + (double)hy - off + sly), (int)P.row0 - 1, (int)P.row1 - 1);
    if (px0 <= px1 && py0 <= py1) {
      tr.x0 = (uint16_t)(px0 / TILE);
      tr.x1 = (uint16_t)(px1 / TILE);
      tr.y0 = (uint16_t)(py0 / TILE - (int)P.tile_y0);
      tr.y1 = (uint16_t)(py1 / TILE - (int)P.tile_y0);
    }
  }

  // ---- colour phase (skipped in stripe renders for Gaussians that cannot touch the stripe)
  bool vis = visg;
  const bool need_colour = visg && !(P.stripe_cull && tr.count() == 0u);
  float col[3] = {0.f, 0.f, 0.f};
  if (need_colour) {
    if (P.stripe_cull) {
#pragma unroll
      for (int k = 3; k < 9; ++k) {
        const float4 v = __ldg(&scene[(size_t)(k + 1) * n + ii]);
        f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
      }
    }
    // pipelines.rs:99-100 + gaussians.rs:41-99  SH colour along normalize(position - camera.position)
    const float d0 = p0.x - P.cam_pos[0], d1 = p0.y - P.cam_pos[1], d2 = p0.z - P.cam_pos[2];
    const float dn = __fsqrt_rn(d0 * d0 + d1 * d1 + d2 * d2);
    const float x = __fdiv_rn(d0, dn), y = __fdiv_rn(d1, dn), z = __fdiv_rn(d2, dn);
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const float k1y = 0.4886025119029199f * y, k1z = 0.4886025119029199f * z, k1x = 0.4886025119029199f * x;
    const float k4 = 1.0925484305920792f * xy, k5 = -1.0925484305920792f * yz;
    const float k6 = 0.31539156525252005f * (2.0f * zz - xx - yy);
    const float k7 = -1.0925484305920792f * xz, k8 = 0.5462742152960396f * (xx - yy);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = 0.28209479177387814f * sh[c];
      v = v - k1y * sh[3 + c] + k1z * sh[6 + c] - k1x * sh[9 + c];
      v = v + k4 * sh[12 + c] + k5 * sh[15 + c] + k6 * sh[18 + c] + k7 * sh[21 + c] + k8 * sh[24 + c];
      col[c] = v + 0.5f;
    }
    vis = finitef(col[0]) && finitef(col[1]) && finitef(col[2]);
  } else if (P.stripe_cull) {
    vis = false;
  }
  if (!vis) { tr.x0 = 1; tr.x1 = 0; tr.y0 = 1; tr.y1 = 0; }

  // power threshold: alpha = min(0.99, op*exp(power)) < 1/255 is certain below pth
  // (0.002 of slack in the exponent against the 1-ulp error of the pinned exp); exp is
  // flushed to zero below -87, so pth never needs to go lower.  (double log: only for
  // Gaussians that are kept)
  float pth = __int_as_float(0x7f800000);   // never contributes
  if (vis && op > 0.0f) {
    const double t = log(1.0 / (255.0 * (double)op)) - 2e-3;   // float rounding of t << 2e-3
    pth = (t < -87.0) ? -87.0f : (float)t;
  }

  if (valid) {
    if (vis) {
      Rec r;
      r.a = make_float4(cxp, cyp, cA, P.ysign * cB);
      r.b = make_float4(cC, op, hx, hy);
      r.c = make_float4(col[0], col[1], col[2], pth);
      recs[i] = r;
    }
    keys[i] = vis ? depth_key(zv) : KEY_CULLED;
    vals[i] = i;
    rects[i] = make_uint2((uint32_t)tr.x0 | ((uint32_t)tr.y0 << 16), (uint32_t)tr.x1 | ((uint32_t)tr.y1 << 16));
    tcnt[i] = tr.count();   // 4-byte gather target for tile_count_kernel (8 per sector, stays in L2)
  }
  if (P.stripe_cull) {
    const int kept = __syncthreads_count(valid && vis);
    if (threadIdx.x == 0) block_kept[blockIdx.x] = (uint32_t)kept;
  }
}

// Stripe renders: squeeze the (depth key, index) pairs of the Gaussians that survived the stripe
// cull to the front, in index order (the stable sort's tie-break is the input order).  CTA b owns
// the 256 pairs project_kernel's CTA b wrote; block_off = exclusive scan of block_kept.
__global__ void __launch_bounds__(256)
compact_pairs_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                     const uint32_t *__restrict__ block_off, uint32_t n) {
  __shared__ uint32_t wcnt[8];
  const uint32_t i = blockIdx.x * 256u + threadIdx.x, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  uint32_t k = KEY_CULLED, v = 0;
  if (i < n) { k = keys_in[i]; v = vals_in[i]; }
  const bool keep = k != KEY_CULLED;
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
  if (lane == 0) wcnt[w] = __popc(bal);
  __syncthreads();
  uint32_t before = 0;
#pragma unroll
  for (uint32_t q = 0; q < 8u; ++q) before += (q < w) ? wcnt[q] : 0u;
  if (keep) {
    const uint32_t o = block_off[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1u));
    keys_out[o] = k;
    vals_out[o] = v;
  }
}

}  // namespace splat
