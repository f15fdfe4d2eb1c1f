/*
 * splat.h -- C ABI of libsplat_b200: a B200 (sm_100a) Gaussian-splat rasteriser that drops in
 * behind thomasantony/splat's render entry point.
 *
 * Replaces (reference paths relative to /root/reference/src):
 *   GaussianSplatPipeline01::render_to_buffer   pipelines.rs:66-86
 *   GaussianSplatPipeline02::render_to_buffer   pipelines.rs:260-280
 * and everything those call per frame: sort_gaussians / GaussianList::sort
 * (gaussians.rs:297-306, :464-471), the euc Pipeline callbacks vertex / fragment / blend
 * (pipelines.rs:96-168, :184-256), gaussian_vertex_shader (:17-51), project_cov3d_to_screen
 * (gaussians.rs:114-161, :473-522), eval_spherical_harmonics (:41-99), compute_cov3d
 * (:101-113, :446-462), and the euc 0.6.0 rasteriser loop itself (Cargo.lock:221-229).
 *
 * PARITY is to the CPU restatement of that path in oracle/ (the reference itself cannot be built
 * here: no Rust toolchain, euc un-vendored): bit-identical pixels for SPLAT_BLEND_REFERENCE.  Known
 * distances from the real Rust/euc binary: (1) exp -- Rust calls the platform libm, this library one
 * pinned operation sequence within 1 ulp of glibc's expf (about 1 pixel in 64,000 differs by one
 * step); (2) euc's per-pixel coverage and attribute interpolation are modelled (inclusive
 * axis-aligned quad, sampled at pixel centres, d = sample - centre), not reproduced from source;
 * (3) Gaussians with a singular 2-D covariance or non-finite values are skipped where the reference
 * panics or blends NaN.  Orientation (NDC +y = top row) is pinned to the reference's own images.
 *
 * Plain C: opaque context, plain pointers and sizes, int error codes; nothing throws or aborts
 * across this boundary.  The Rust side binds it with bindgen; tests and
 * bench.py bind it with ctypes (splat_b200/_lib.py).
 *
 * Threading: one render at a time per context (not re-entrant); distinct contexts are
 * independent.  All host pointers are borrowed for the duration of the call only.
 */
#ifndef SPLAT_B200_H
#define SPLAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPLAT_ABI_VERSION 2u

typedef struct splat_ctx splat_ctx;

enum {
  SPLAT_OK = 0,
  SPLAT_ERR_INVALID = -1,     /* bad argument (null pointer, zero size, misaligned stripe ...)  */
  SPLAT_ERR_CUDA = -2,        /* a CUDA runtime call failed; see splat_last_error               */
  SPLAT_ERR_NOMEM = -3,       /* host or device allocation failed                               */
  SPLAT_ERR_UNSUPPORTED = -4, /* e.g. camera.w/h differ from the target size, tile != 16        */
  SPLAT_ERR_STATE = -5,       /* render before upload                                           */
  SPLAT_ERR_RETRY = -6        /* the previous splat_render_device frame was skipped on the device (see there);
                               * the buffers have been grown: render that frame again                */
};

/* blend_mode values.  Only SPLAT_BLEND_REFERENCE reproduces the reference's pixels
 * (pipelines.rs:147-168: u8 truncation after every Gaussian, far -> near).
 * SPLAT_BLEND_FLOAT is the standard un-quantised formulation of the same image: per pixel, nearest
 * first, C += T*alpha*colour, T *= 1-alpha with the reference's fragment() (pipelines.rs:127-145)
 * unchanged, stopped when T < 2^-16, then ONE quantisation: byte = (C + T*old/255)*255 truncated,
 * alpha byte = (1-T)*255; pixels no fragment contributes to are left untouched.  It differs from
 * the reference by that blend's own per-layer truncation bias (6-9e-3 RMSE) and is
 * checked against a float CPU restatement (<= 1e-4 RMSE) that is itself pinned to the reference's
 * Python prototype (notebook cell 3). */
enum {
  SPLAT_BLEND_REFERENCE = 0,
  SPLAT_BLEND_FLOAT = 1
};

/* Behaviour switches.  lowpass selects which reference pipeline is reproduced; the three
 * euc-semantics switches mirror the assumptions E1/E3 of SURVEY.md section 8c and must match
 * the oracle's orc_config for parity. */
typedef struct {
  int32_t  device;         /* CUDA device ordinal                                                */
  float    lowpass;        /* 0.01 = Pipeline01 (gaussians.rs:156-157), 0.3 = Pipeline02 (:517-518) */
  int32_t  y_down;         /* 0 (default): NDC +y is the TOP row -- what the reference's own images show
                            * (notes/screenshot.png and notebook cells 3/6, pinned by
                            * tests/test_reference_images.py); 1: NDC +y is the bottom row          */
  int32_t  zclip_mode;     /* 0: keep 0<=z<1, 1 (default, goes with y-up): keep -1<=z<1, 2: no z clip */
  float    sample_offset;  /* pixel sample point (x+off, y+off); 0.5                             */
  uint32_t tile;           /* screen tile edge in pixels; only 16 is built                       */
  uint64_t max_instances;  /* initial capacity of the tile-instance buffers (0 = auto, grows)    */
  int32_t  blend_mode;     /* SPLAT_BLEND_*: 0 = the reference's quantised far->near blend (parity) */
  int32_t  near_cut;       /* The reference blend reads only the nearest few hundred entries of a tile list (exact
                            * early termination), so a frame first bins + sorts only the nearest k/1024 of the
                            * Gaussians; tiles that only a few of the cut Gaussians touch get those binned after all
                            * ("open" tiles), and a frame whose near lists still do not suffice is abandoned on the
                            * device and repeated without the cut (host-buffer calls do that themselves;
                            * splat_render_device reports SPLAT_ERR_RETRY).  Pixels are identical either way.
                            * -1 (default) = automatic: 1/8 once a scene shows >= 200k visible Gaussians and tile
                            * lists of >= 2048 entries on average, doubled after any failure; 0 = off;
                            * 1..1024 = fixed fraction.  Nothing is read on the host during a frame.           */
  int32_t  sync_frames;    /* 1: read the tile-instance count on the host in the middle of every frame (exact
                            * launch sizes, one round trip per frame).  0 (default): only the first frame of a
                            * target geometry does; later frames are enqueued without any host wait            */
  int32_t  reserved;       /* 0 */
} splat_config;

/* What the kernels need from `Camera` (camera.rs:4-19): the two matrices exactly as nalgebra
 * stores them (column-major), the *field* `position` (camera.rs:10 -- never updated by
 * orbiting, and the SH view direction uses it, pipelines.rs:99), w/h and
 * get_htanfovxy_focal() (camera.rs:84-89). */
typedef struct {
  float view[16];     /* camera.get_view_matrix().as_slice()    */
  float proj[16];     /* camera.get_project_matrix().as_slice() */
  float position[3];  /* camera.position                        */
  float w, h;         /* camera.w, camera.h                     */
  float htanx, htany, focal; /* camera.get_htanfovxy_focal()    */
} splat_camera;

/* Per-stage device times of the last completed render (CUDA events) and its work counters. */
typedef struct {
  float    project_ms;   /* K1 project                                         */
  float    sort_ms;      /* K3 depth radix sort + tile radix sort              */
  float    bin_ms;       /* K2 count/scan/emit + K4 tile ranges                */
  float    blend_ms;     /* K5 blend                                           */
  float    total_ms;     /* first kernel to last kernel                        */
  float    h2d_ms;       /* framebuffer upload (host-buffer entry points only) */
  float    d2h_ms;       /* framebuffer download                               */
  uint32_t frames_retried; /* renders (partly) repeated: instance buffers had to grow, or a near-cut pass did not converge */
  uint64_t n_gaussians;
  uint64_t n_visible;    /* Gaussians that pass the z clip and the degeneracy guard (stripe renders:
                          * and whose quad can touch the stripe -- the others are dropped before the sort) */
  uint64_t n_instances;  /* (tile, Gaussian) pairs                                  */
  uint64_t n_tiles;      /* tiles in the rendered stripe                            */
  uint64_t kernel_launches; /* kernels launched by the last render                  */
  uint64_t near_cut_rank;   /* depth ranks below this were left out of the first binning pass (0 = none) */
  uint64_t near_cut_failed; /* pixel groups / tiles that did not converge in that pass (0 = it was enough) */
  uint64_t near_cut_instances; /* (tile, Gaussian) pairs that pass did not have to bin and sort */
  uint64_t frames_skipped;  /* since the context was created: frames abandoned on the device on the no-round-trip
                             * path -- their tile instances outgrew the launch bounds, or their near lists did
                             * not suffice (host-buffer calls repeat them themselves)                          */
  uint64_t second_pass_instances; /* `make cdp` build only (a second pass launched from the device instead of a
                                   * repeated frame): pairs that pass binned for the tiles that did not converge
                                   * on the near lists; 0 in the shipped build                                  */
  uint64_t near_cut_fallbacks;    /* since the context was created: near-cut frames whose near lists did not suffice */
  float    second_pass_ms;  /* the device-side check of the near pass (and, `make cdp`, the second pass); in total_ms */
  float    reserved_;
} splat_timings;

uint32_t    splat_abi_version(void);
void        splat_config_default(splat_config *cfg);
int         splat_create(splat_ctx **out, const splat_config *cfg);
/* why the last splat_create on this thread failed ("" if it did not): there is no context to ask */
const char *splat_create_error(void);
void        splat_destroy(splat_ctx *ctx);
const char *splat_last_error(const splat_ctx *ctx);

/* Scene upload, GaussianList layout (gaussians.rs:408-416; from_vec :419-440): each array is
 * the contiguous column-major nalgebra matrix -- pos4 4xN (x,y,z,1), scale3 3xN (already
 * exp'd), opacity N (already sigmoid'd), rot_xyzw 4xN (nalgebra coords order i,j,k,w,
 * un-normalised is fine), sh48 48xN.  Host pointers.  Copies to the device, computes cov3d
 * there (compute_cov3d, gaussians.rs:446-462) and repacks to the device layout.
 * n = 0 (every upload call) is a valid, empty scene: its frames check their arguments and draw nothing, like
 * render_to_buffer over an empty Vec<Gaussian>. */
int splat_upload_soa(splat_ctx *ctx, const float *pos4, const float *scale3, const float *opacity,
                     const float *rot_xyzw, const float *sh48, uint64_t n);

/* Scene upload for Pipeline01's Vec<Gaussian> (gaussians.rs:31-38).  The Rust struct is not
 * repr(C), so the shim copies each Gaussian into 59 consecutive floats:
 * position[3] scale[3] opacity rotation_xyzw[4] sh[48]. */
int splat_upload_aos(splat_ctx *ctx, const float *gaussians59, uint64_t n);

/* Scene upload straight from a PLY file's vertex payload -- load_from_ply (gaussians.rs:375-405) with
 * set_property (:258-282) on the device.  vertex_rows: host pointer to n vertices of the INRIA 3DGS
 * binary_little_endian layout (62 f32: x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..2
 * rot_0..3), stride_floats apart (62 for a plain file).  The device applies exp / sigmoid / the
 * rot_0 -> w and f_rest_i -> sh[3+i] mappings and subtracts the mean position, accumulated
 * sequentially in f32 in file order exactly like the reference.  exp is the library's pinned routine
 * (<= 1 ulp from libm's expf, which Rust calls): activated scales / opacities may differ from the
 * reference's by one ulp.  activated60 (NULL to skip): receives the activated GaussianList arrays,
 * n each of pos4 | rot_xyzw | scale3 | opacity | sh48 back to back (tests). */
int splat_upload_ply_raw(splat_ctx *ctx, const void *vertex_rows, uint64_t n, uint32_t stride_floats, float *activated60);

/* render_to_buffer.  fb_inout: W*H pixels, row-major, 0xAARRGGBB (euc::Buffer<u32,2>::raw(),
 * main.rs:79), blended onto (the caller clears it, main.rs:73) and overwritten.  Synchronous:
 * host->device copy of fb, all kernels, device->host copy.  Requires camera.w == W and
 * camera.h == H (true for every caller in the reference). */
int splat_render(splat_ctx *ctx, const splat_camera *cam, uint32_t *fb_inout, uint32_t W, uint32_t H);

/* `color = Buffer2d::fill([W, H], clear); render_to_buffer(&mut color)` in one call (what
 * main.rs:73-74 does every frame with clear = 0): the frame is cleared on the device, so the
 * caller neither fills the host buffer nor pays its upload; fb_out is only written. */
int splat_render_cleared(splat_ctx *ctx, const splat_camera *cam, uint32_t *fb_out, uint32_t W, uint32_t H,
                         uint32_t clear);

/* Same for a horizontal stripe of rows [row0,row1) of the W x H image (multi-GPU sharding by
 * screen-tile stripes): fb_rows points at row row0 and holds (row1-row0)*W pixels.  row0 must
 * be a multiple of the tile size; row1 a multiple of it or == H. */
int splat_render_rows(splat_ctx *ctx, const splat_camera *cam, uint32_t *fb_rows, uint32_t W,
                      uint32_t H, uint32_t row0, uint32_t row1);

/* Device-buffer variant: fb_rows_dev is a device pointer on cfg.device (e.g. the slice of an
 * NCCL-registered gather buffer); work is enqueued on `stream` (a cudaStream_t, NULL = the
 * context's own stream) and the call returns without waiting for it.
 * The first frame of a target geometry (W, H, row0, row1) blocks once mid-frame to read the
 * tile-instance count and size the buffers.  Every later frame is enqueued with NO host wait: its
 * launches are sized from the previous frame (+12.5%), the kernels read the real counts on the
 * device, and a near-cut frame decides on the device whether its second pass has work.  Should a
 * frame nevertheless outgrow those bounds, the device abandons it: without a near cut nothing was
 * blended and the target is exactly as the caller provided it; with a near cut the tiles that had
 * already converged hold their final pixels and all others are untouched.  The next call on this
 * context (render or splat_get_timings) grows the buffers and returns SPLAT_ERR_RETRY once: put the
 * target's previous contents back (e.g. clear it again) and render that frame again.  The
 * host-buffer entry points do all of that themselves. */
int splat_render_device(splat_ctx *ctx, const splat_camera *cam, void *fb_rows_dev, uint32_t W,
                        uint32_t H, uint32_t row0, uint32_t row1, void *stream);

/* Blocks until the last enqueued render finished, then reports its stage times (all fields describe
 * that render; frames_skipped is cumulative). */
int splat_get_timings(splat_ctx *ctx, splat_timings *out);

/* Tile-list lengths of the last completed render: per_tile[ty * tiles_x + tx] = number of
 * (tile, Gaussian) instances of that tile of the rendered stripe (tiles_x = ceil(W/16)).  The
 * multi-GPU driver sums them per tile row to place stripe boundaries.  Writes
 * min(cap, n_tiles) entries and returns the stripe's tile count in *n_tiles. */
int splat_get_tile_loads(splat_ctx *ctx, uint32_t *per_tile, uint64_t cap, uint64_t *n_tiles);

/* ---- Multi-GPU.  The frame shards by horizontal stripes of whole tile rows; the scene
 * is replicated (one broadcast at upload); the only per-frame exchange is ONE gather of the
 * stripes' rows into the root's frame (grouped ncclSend / ncclRecv over NVLink, straight from and
 * into the render targets).  The reference has no counterpart: its only parallelism is euc's
 * row-group threading.  libnccl.so.2 is loaded on first use; single-GPU callers never need it.
 *
 * (1) One process, several GPUs -- what a Rust caller of render_to_buffer gets by listing devices:
 * splat_create_multi returns a GROUP context that owns one ordinary context per device and their
 * communicators (ncclCommInitAll).  splat_upload_soa/aos upload to devices[0] and broadcast the
 * packed scene; splat_render / splat_render_cleared render every member's stripe concurrently,
 * gather, and return the frame; the stripe boundaries are re-cut from the members' measured frame
 * times; splat_get_timings reports the slowest member per stage.  Pixels are byte-identical to a
 * single-device render for every partition.  cfg.reserved = 1 keeps equal stripes. */
int splat_create_multi(splat_ctx **out, const splat_config *cfg, const int32_t *devices, int32_t n_devices);
/* stripe rows [bounds[2k], bounds[2k+1]) of member k for the next frame */
int splat_group_get_bounds(splat_ctx *group, uint32_t *bounds, int32_t cap_ranks, int32_t *n_ranks);

/* (2) One process per GPU (a torchrun-style launcher): rank 0 makes the id, the launcher carries its
 * 128 bytes to the other ranks, every rank joins with its own ordinary context. */
#define SPLAT_UNIQUE_ID_BYTES 128
int splat_comm_unique_id(void *id128);
int splat_comm_init_rank(splat_ctx *ctx, const void *id128, int32_t n_ranks, int32_t rank);
/* C0: `root` has uploaded n Gaussians; every other rank receives the packed device scene. */
int splat_comm_broadcast_scene(splat_ctx *ctx, int32_t root, uint64_t n);
/* C1: rows [bounds[2r], bounds[2r+1]) of rank r's W x H device frame -> the same rows of root's frame,
 * enqueued on `stream` (NULL = the context's) behind the render that wrote them.  bounds: 2 * n_ranks
 * entries, contiguous, ordered, covering [0, H). */
int splat_gather_stripes(splat_ctx *ctx, void *fb_dev, uint32_t W, uint32_t H, const uint32_t *bounds, int32_t root,
                         void *stream);

/* Pin / unpin a caller-owned host buffer (cudaHostRegister) so that the framebuffer copies of
 * splat_render run at full PCIe rate; e.g. the Rust shim pins `color.raw_mut()` once. */
int splat_pin_host(void *p, uint64_t bytes);
int splat_unpin_host(void *p);

/* Debug / parity taps (tests only).
 * splat_debug_project runs K1 alone for the full frame and copies out, per Gaussian:
 *   records12: 12 floats (cxp cyp A B | C opacity hx hy | r g b power_threshold), zeros if culled;
 *   depth_keys: u32 sort key (0xFFFFFFFF = culled); tile_rects4: 4 x u16 (x0 y0 x1 y1, inclusive).
 * splat_debug_read_order returns the depth order (Gaussian indices far -> near) of the last render
 *   (n_visible entries; for a stripe render only the Gaussians that can touch the stripe).
 * splat_debug_sort_pairs sorts host (key,value) arrays in place on bits [0,bits) with the device
 * radix sort. */
int splat_debug_project(splat_ctx *ctx, const splat_camera *cam, uint32_t W, uint32_t H,
                        float *records12, uint32_t *depth_keys, uint32_t *tile_rects4);
int splat_debug_read_order(splat_ctx *ctx, uint32_t *order, uint64_t cap, uint64_t *n_visible);
int splat_debug_sort_pairs(splat_ctx *ctx, uint32_t *keys, uint32_t *vals, uint64_t n, int bits);
/* Work counters of the blend kernel accumulated since the last reset (instrumented builds,
 * -DSPLAT_STATS, only; SPLAT_ERR_UNSUPPORTED otherwise): out8[0] group-entries evaluated,
 * [1] group-entries that changed a pixel, [2] (pixel, Gaussian) pairs with alpha > 0,
 * [3] tile-list entries staged, [4] candidate pairs, [5] pixel-pair lanes with alpha > 0. */
int splat_debug_blend_stats(splat_ctx *ctx, uint64_t *out8, int reset);
/* per-tile arrays of the last frame: which = 0 list ranges (2 words per tile), 1 cut Gaussians per tile
 * (near-cut frames), 2 tiles marked for the second pass */
int splat_debug_read_tiles(splat_ctx *ctx, int which, uint32_t *out, uint64_t cap_words, uint64_t *n_words);
/* SPLAT_BLEND_FLOAT contexts: splat_render, plus the un-quantised result of every pixel a fragment
 * contributed to: rgba[4*(y*W+x)] = r, g, b, 1-T (NaN where the pixel was not touched). */
int splat_debug_render_float(splat_ctx *ctx, const splat_camera *cam, uint32_t *fb_inout, uint32_t W, uint32_t H,
                             float *rgba);
/* Host only (no device, no context): the stripe partition rules of a group context.  ms == NULL: the
 * initial cut, equal numbers of tile rows.  Otherwise bounds[2*parts] holds the current [row0,row1) per
 * member and ms[parts] the members' measured frame times; bounds is replaced by the re-cut partition. */
int splat_debug_partition(uint32_t *bounds, const float *ms, int32_t parts, uint32_t H);

#ifdef __cplusplus
}
#endif
#endif /* SPLAT_B200_H */
