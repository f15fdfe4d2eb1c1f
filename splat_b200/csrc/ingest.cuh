// ingest.cuh -- f-1: PLY ingest on the device.  Replaces load_from_ply (gaussians.rs:375-405) and
// PropertyAccess::set_property (:258-282) for the INRIA 3DGS vertex layout -- 62 little-endian f32
// per vertex: x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..2 rot_0..3 (notes.md:1-8,
// notes/util_gau.py:66-96).  The caller hands over the raw vertex payload (mmap'ed or read); the
// device activates it:
//   scale_i -> exp(scale_i)                 :265-267
//   opacity -> 1 / (1 + exp(-opacity))      :268
//   rot_0 -> w, rot_1..3 -> i, j, k         :269-272   (the device scene keeps nalgebra's i, j, k, w order)
//   f_dc_k -> sh[k], f_rest_i -> sh[3 + i]  :273-279   (no channel transpose)
//   position -= mean position, the mean accumulated SEQUENTIALLY in f32 in file order, one
//   division by n at the end                :394-402   (order-dependent rounding, restated exactly)
// exp policy: Rust's f32::exp is the platform libm; this path uses the same pinned routine as the
// blend ("splat_expf v1", <= 1 ulp from glibc expf), so activated scales / opacities can differ
// from the reference's by one ulp -- stated here and in DESIGN.md.
#pragma once
#include "blend.cuh"
#include "common.cuh"

namespace splat {

constexpr int PLY_FLOATS = 62;   // x y z | nx ny nz | f_dc 3 | f_rest 45 | opacity | scale 3 | rot 4
constexpr int PLY_X = 0, PLY_DC = 6, PLY_REST = 9, PLY_OPACITY = 54, PLY_SCALE = 55, PLY_ROT = 58;

// splat_expf v1 on its whole domain (oracle: orc_expf)
SPLAT_DEVINL float expf_pinned_full(float x) {
  if (!(x >= -87.0f)) return (x != x) ? x : 0.0f;
  if (x > 88.0f) return __int_as_float(0x7f800000);
  return expf_pinned(x);
}

// mean3 = (sum of positions, sequentially in f32, in file order) / n.   One CTA: 1024 threads stage
// 1024 rows' x, y, z in shared memory, thread 0 adds them in order (three independent chains).
__global__ void __launch_bounds__(1024)
ply_mean_kernel(const float *__restrict__ rows, uint32_t stride, unsigned long long n, float *__restrict__ mean3) {
  __shared__ float sx[1024], sy[1024], sz[1024];
  float ax = 0.0f, ay = 0.0f, az = 0.0f;
  for (unsigned long long base = 0; base < n; base += 1024ull) {
    const unsigned long long i = base + threadIdx.x;
    if (i < n) {
      const float *r = rows + i * stride;
      sx[threadIdx.x] = r[PLY_X]; sy[threadIdx.x] = r[PLY_X + 1]; sz[threadIdx.x] = r[PLY_X + 2];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t m = (uint32_t)min(1024ull, n - base);
      for (uint32_t k = 0; k < m; ++k) { ax = __fadd_rn(ax, sx[k]); ay = __fadd_rn(ay, sy[k]); az = __fadd_rn(az, sz[k]); }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float fn = (float)n;   // `gaussians.len() as f32`
    mean3[0] = __fdiv_rn(ax, fn); mean3[1] = __fdiv_rn(ay, fn); mean3[2] = __fdiv_rn(az, fn);
  }
}

// One thread per vertex; a CTA stages its 128 rows in shared memory with coalesced loads.  Out: the
// GaussianList arrays (gaussians.rs:408-416) that splat_upload_soa also stages on the device.
constexpr int PLY_ROWS = 128;
__global__ void __launch_bounds__(PLY_ROWS)
ply_activate_kernel(const float *__restrict__ rows, uint32_t stride, uint32_t n, const float *__restrict__ mean3,
                    float4 *__restrict__ pos4, float *__restrict__ scale3, float *__restrict__ opacity,
                    float4 *__restrict__ rot, float *__restrict__ sh48) {
  __shared__ float s[PLY_ROWS * PLY_FLOATS];
  const uint32_t row0 = blockIdx.x * PLY_ROWS, m = min((uint32_t)PLY_ROWS, n - row0);
  for (uint32_t k = threadIdx.x; k < m * PLY_FLOATS; k += PLY_ROWS) {
    const uint32_t r = k / PLY_FLOATS, f = k - r * PLY_FLOATS;
    s[k] = rows[(size_t)(row0 + r) * stride + f];
  }
  __syncthreads();
  if (threadIdx.x >= m) return;
  const uint32_t i = row0 + threadIdx.x;
  const float *v = s + threadIdx.x * PLY_FLOATS;
  pos4[i] = make_float4(__fsub_rn(v[PLY_X], mean3[0]), __fsub_rn(v[PLY_X + 1], mean3[1]), __fsub_rn(v[PLY_X + 2], mean3[2]), 1.0f);
#pragma unroll
  for (int k = 0; k < 3; ++k) scale3[3 * (size_t)i + k] = expf_pinned_full(v[PLY_SCALE + k]);
  opacity[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_pinned_full(-v[PLY_OPACITY])));
  rot[i] = make_float4(v[PLY_ROT + 1], v[PLY_ROT + 2], v[PLY_ROT + 3], v[PLY_ROT]);
  float *sh = sh48 + 48 * (size_t)i;
#pragma unroll
  for (int k = 0; k < 3; ++k) sh[k] = v[PLY_DC + k];
#pragma unroll
  for (int k = 0; k < 45; ++k) sh[3 + k] = v[PLY_REST + k];
}

}  // namespace splat
