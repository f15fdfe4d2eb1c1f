// common.cuh -- shared declarations of the sm_100a splat rasteriser (product code).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace splat {

constexpr int TILE = 16;                 // screen tile edge (pixels)
constexpr int SCENE_PLANES = 10;         // float4 planes per Gaussian in the device scene
constexpr uint32_t KEY_CULLED = 0xFFFFFFFFu;

// Device scene layout (HBM, written once per upload by pack_scene_kernel), 160 B per Gaussian:
//   plane 0      : (x, y, z, cov3d[8])
//   planes 1..2  : cov3d[0..7] (row-major)        -- planes 0-2 are all the stripe pre-pass reads
//   plane 3      : (opacity, sh[0], sh[1], sh[2])
//   planes 4..9  : sh[3..26] (SH degrees 1..2)
// Each plane is a dense float4[N] array, so a warp reading plane k for 32 consecutive
// Gaussians issues one fully coalesced 512-byte request.

// Everything the kernels need to know about one frame; passed by value (__grid_constant__).
struct FrameParams {
  float view[16];      // column-major, camera.get_view_matrix()
  float proj[16];      // column-major, camera.get_project_matrix()
  float cam_pos[3];    // camera.position
  float focal, htanx, htany;
  float lowpass;
  float sample_off;
  float ysign;         // +1: y_down, -1: y-up
  int zclip_mode;
  int stripe_cull;     // 1: drop Gaussians that cannot touch the stripe before the colour phase and the depth sort
  uint32_t W, H;       // full image size in pixels
  uint32_t row0, row1; // rendered stripe [row0,row1)
  uint32_t tiles_x;    // ceil(W/16)
  uint32_t tile_y0;    // first tile row of the stripe
  uint32_t tiles_y;    // tile rows in the stripe
  uint32_t n;          // Gaussians
  unsigned long long nz2;  // (-0.0f, -0.0f): run-time addend that keeps packed products unfused (blend.cuh)
};

// Splat record produced by project_kernel and consumed by blend_kernel: 3 x float4 = 48 B.
//   a = (cxp, cyp, A, B)   pixel-space centre, conic A and B (B carries the y-axis sign)
//   b = (C, opacity, hx, hy) conic C, opacity, 3-sigma half extents in pixels
//   c = (r, g, b, pth)     SH colour (unclamped), conservative power threshold below which
//                          alpha < 1/255 is certain
struct Rec {
  float4 a, b, c;
};

// Frame status written by the device, copied to pinned host memory at the end of every frame.
struct FrameStatus {
  unsigned long long n_instances;  // total (tile, Gaussian) pairs wanted
  unsigned int n_visible;
  unsigned int n_failed;           // near-cut frames: groups / tiles whose pixels need Gaussians that were cut
  unsigned int fail_ix0, fail_iy0; // bounding box of those tiles: ~min x, ~min y (kept as maxima so that 0 = none)
  unsigned int fail_x1, fail_y1;   //                             max x, max y
  unsigned long long n_cut;        // near-cut frames: (tile, Gaussian) pairs that were not binned
  // ---- everything above is zeroed at the start of every binning pass
  unsigned long long n_sort;       // stripe renders: (key, index) pairs that enter the depth sort
  unsigned int n_inst_eff;         // pairs the kernels behind the count process: n_instances, or 0 if they do not fit
  unsigned int overflow;           // this frame wanted more pairs than the instance buffers hold: nothing was blended
  // near-cut frames, written between the two passes (pass_b_setup_kernel)
  unsigned long long a_instances;  // pairs the first (near) pass binned
  unsigned int a_failed;           // units / tiles that did not converge on the near lists
  unsigned int a_visible;
  unsigned int a_tag;              // cut_frac the frame was rendered with (0: no cut)
  unsigned int b_active;           // 1: the second pass (complete lists of the failed tiles) has work
  unsigned int a_failed_tiles;     // tiles the second pass redoes
  // ---- everything above is zeroed at the start of every frame
  unsigned int skipped;            // frames skipped that way since the context was created (never zeroed)
};
static_assert(sizeof(FrameStatus) % 8 == 0, "FrameStatus is copied as a block");

#define SPLAT_DEVINL __device__ __forceinline__

}  // namespace splat
