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

// Device scene layout (common.cuh): plane 0 = (x, y, z, cov3d[8]), planes 1-2 = cov3d[0..7],
// plane 3 = (opacity, sh[0..2]), planes 4-9 = sh[3..26].  The first three planes (48 B) are all
// the geometry the stripe pre-pass needs.  f[0..8] = cov3d row-major, f[9..35] = sh[0..26].
SPLAT_DEVINL void store_scene(float4 *__restrict__ scene, uint32_t n, uint32_t i, float x, float y, float z, float op,
                              const float f[36]) {
  scene[i] = make_float4(x, y, z, f[8]);
  scene[(size_t)1 * n + i] = make_float4(f[0], f[1], f[2], f[3]);
  scene[(size_t)2 * n + i] = make_float4(f[4], f[5], f[6], f[7]);
  scene[(size_t)3 * n + i] = make_float4(op, f[9], f[10], f[11]);
#pragma unroll
  for (int k = 0; k < 6; ++k)
    scene[(size_t)(k + 4) * n + i] = make_float4(f[12 + 4 * k], f[13 + 4 * k], f[14 + 4 * k], f[15 + 4 * k]);
}
// the reverse: all ten planes of Gaussian ii (ten independent 16-byte loads in flight)
SPLAT_DEVINL void load_scene(const float4 *__restrict__ scene, uint32_t n, uint32_t ii, float4 &p0, float &op, float f[36]) {
  float4 v[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) v[k] = __ldg(&scene[(size_t)k * n + ii]);
  p0 = v[0];
  f[0] = v[1].x; f[1] = v[1].y; f[2] = v[1].z; f[3] = v[1].w;
  f[4] = v[2].x; f[5] = v[2].y; f[6] = v[2].z; f[7] = v[2].w;
  f[8] = v[0].w;
  op = v[3].x; f[9] = v[3].y; f[10] = v[3].z; f[11] = v[3].w;
#pragma unroll
  for (int k = 0; k < 6; ++k) { f[12 + 4 * k] = v[4 + k].x; f[13 + 4 * k] = v[4 + k].y; f[14 + 4 * k] = v[4 + k].z; f[15 + 4 * k] = v[4 + k].w; }
}

// One thread per Gaussian.  Inputs are the raw GaussianList arrays (gaussians.rs:408-416)
// already on the device; output is the 10-plane float4 scene.
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
  store_scene(scene, n, i, p.x, p.y, p.z, opacity[i], f);
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
  store_scene(scene, n, i, g[0], g[1], g[2], g[6], f);
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

// ---------------------------------------------------------------- stripe pre-pass (multi-GPU)
// A rank that renders one stripe of the screen needs only the Gaussians whose 3-sigma quad can
// reach its rows.  One thread per Gaussian reads the 48 B of geometry (planes 0-2), repeats the
// y half of the projection -- view transform, the yy entry of the 2D covariance, NDC y -- in
// exactly the operations of project_kernel, and votes: bit = 1 if rows [row0, row1) widened by one
// pixel can intersect the quad.  A superset by construction (same arithmetic plus the margin); the
// exact decision is made by project_kernel, which then runs densely over the survivors only.
// Out: one ballot word per warp and the survivor count per CTA.
__global__ void __launch_bounds__(256)
stripe_cull_kernel(const float4 *__restrict__ scene, const __grid_constant__ FrameParams P,
                   uint32_t *__restrict__ mask, uint32_t *__restrict__ block_kept) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = P.n;
  const bool valid = i < n;
  const uint32_t ii = valid ? i : n - 1u;
  const float4 p0 = __ldg(&scene[ii]);
  const float4 c0 = __ldg(&scene[(size_t)1 * n + ii]), c1 = __ldg(&scene[(size_t)2 * n + ii]);
  const float C3[9] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, p0.w};
  const float *V = P.view;
  float pc[4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    pc[r] = ((V[0 * 4 + r] * p0.x + V[1 * 4 + r] * p0.y) + V[2 * 4 + r] * p0.z) + V[3 * 4 + r];
  const float zv = pc[2];
  const float jd = __fdiv_rn(P.focal, zv);
  const float T1[3] = {V[0 * 4 + 1] * jd, V[1 * 4 + 1] * jd, V[2 * 4 + 1] * jd};
  float X1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) X1[k] = (T1[0] * C3[k * 3 + 0] + T1[1] * C3[k * 3 + 1]) + T1[2] * C3[k * 3 + 2];
  const float m22 = ((X1[0] * T1[0] + X1[1] * T1[1]) + X1[2] * T1[2]) + P.lowpass;
  const float hy = 3.0f * __fsqrt_rn(m22);
  const float *Pm = P.proj;
  const float ps1 = ((Pm[0 * 4 + 1] * pc[0] + Pm[1 * 4 + 1] * pc[1]) + Pm[2 * 4 + 1] * pc[2]) + Pm[3 * 4 + 1] * pc[3];
  const float ps2 = ((Pm[0 * 4 + 2] * pc[0] + Pm[1 * 4 + 2] * pc[1]) + Pm[2 * 4 + 2] * pc[2]) + Pm[3 * 4 + 2] * pc[3];
  const float ps3 = ((Pm[0 * 4 + 3] * pc[0] + Pm[1 * 4 + 3] * pc[1]) + Pm[2 * 4 + 3] * pc[2]) + Pm[3 * 4 + 3] * pc[3];
  const float ndy = __fdiv_rn(ps1, ps3), ndz = __fdiv_rn(ps2, ps3);
  const float cyp = ((P.ysign * ndy) * 0.5f + 0.5f) * (float)P.H;
  bool zok;
  if (P.zclip_mode == 0) zok = (ndz >= 0.0f && ndz < 1.0f);
  else if (P.zclip_mode == 1) zok = (ndz >= -1.0f && ndz < 1.0f);
  else zok = true;
  bool keep = valid && zok && finitef(zv) && finitef(hy) && finitef(cyp) && fabsf(cyp) <= 1e9f;
  if (keep) {
    const double off = (double)P.sample_off;
    const double sly = 2.0 + 1e-6 * (fabs((double)cyp) + (double)hy);     // project_kernel's slack + one pixel
    const double y0 = floor((double)cyp - (double)hy - off - sly), y1 = ceil((double)cyp + (double)hy - off + sly);
    keep = y1 >= (double)P.row0 && y0 < (double)P.row1;
  }
  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep);
  if ((threadIdx.x & 31u) == 0u && (i >> 5) < (n + 31u) / 32u) mask[i >> 5] = bal;
  const int kept = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_kept[blockIdx.x] = (uint32_t)kept;
}

// surv[block_off[b] + rank] = i for every voted Gaussian, in index order (the stable depth sort's
// tie-break is the input order).  CTA b owns the 256 Gaussians stripe_cull_kernel's CTA b voted on.
__global__ void __launch_bounds__(256)
compact_idx_kernel(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ block_off,
                   uint32_t *__restrict__ surv, uint32_t n) {
  __shared__ uint32_t wcnt[8];
  const uint32_t i = blockIdx.x * 256u + threadIdx.x, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t bal = (i - lane) < n ? mask[i >> 5] : 0u;
  if (lane == 0) wcnt[w] = __popc(bal);
  __syncthreads();
  uint32_t before = 0;
#pragma unroll
  for (uint32_t q = 0; q < 8u; ++q) before += (q < w) ? wcnt[q] : 0u;
  if ((bal >> lane) & 1u) surv[block_off[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1u))] = i;
}

// K1.  One thread per Gaussian (SURV = false: all of them; SURV = true: the j-th survivor of the
// stripe pre-pass, `n_surv` of them, a device-side count).  Ten coalesced (or, for survivors,
// gathered) float4 loads, view transform, cov2d, conic, 3-sigma extents, z clip, tile rectangle,
// SH colour.  Out: 48 B record + 8 B tile rect + 4 B tile count (by Gaussian index) and the
// (depth key, index) pair (by thread index: a stripe's pairs are dense and in index order).
// Algorithmic bytes: 160 in + 68 out = 228 B per Gaussian.
template <bool SURV>
__global__ void __launch_bounds__(256)
project_kernel(const float4 *__restrict__ scene, const __grid_constant__ FrameParams P,
               Rec *__restrict__ recs, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
               uint2 *__restrict__ rects, uint32_t *__restrict__ tcnt,
               const uint32_t *__restrict__ surv, const uint32_t *__restrict__ n_surv,
               uint32_t *__restrict__ block_kept /* !SURV stripe frames: kept Gaussians per CTA (compact_pairs_kernel) */) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = P.n;
  const uint32_t count = SURV ? *n_surv : n;
  const bool valid = j < count;
  if (SURV && blockIdx.x * blockDim.x >= count) return;
  const uint32_t i = valid ? (SURV ? __ldg(&surv[j]) : j) : n - 1u;   // out-of-range threads recompute the last Gaussian, store nothing
  float4 p0;
  float op;
  float f[36];
  load_scene(scene, n, i, p0, op, f);
  const float *C3 = f;        // cov3d row-major (f[0..8])
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

  // visibility: euc's z clip on the centre (all four corners share z) + degeneracy guard
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

  // pipelines.rs:99-100 + gaussians.rs:41-99  SH colour along normalize(position - camera.position)
  float col[3];
  {
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
  }
  // a stripe keeps only what can touch its rows (the others never reach the sort)
  const bool vis = visg && finitef(col[0]) && finitef(col[1]) && finitef(col[2]) && !(P.stripe_cull && tr.count() == 0u);
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
    keys[j] = vis ? depth_key(zv) : KEY_CULLED;
    vals[j] = i;
    rects[i] = make_uint2((uint32_t)tr.x0 | ((uint32_t)tr.y0 << 16), (uint32_t)tr.x1 | ((uint32_t)tr.y1 << 16));
    tcnt[i] = tr.count();   // 4-byte gather target for tile_count_kernel (8 per sector, stays in L2)
  }
  if (!SURV && block_kept) {
    const int kept = __syncthreads_count(valid && vis);
    if (threadIdx.x == 0) block_kept[blockIdx.x] = (uint32_t)kept;
  }
}

// Dense stripes (most Gaussians can reach the stripe: the pre-pass would only add work): the full
// projection ran over all Gaussians; squeeze the (depth key, index) pairs of the kept ones to the
// front, in index order (the stable sort's tie-break is the input order).  CTA b owns the 256 pairs
// project_kernel's CTA b wrote; block_off = exclusive scan of block_kept.
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
