/* TEST INFRASTRUCTURE, not product code.  A stand-in for libsplat_b200.so that implements the handful of
 * C-ABI entry points include/splat_pipeline.hpp calls on top of the CPU oracle (liboracle.so), so that the
 * C++ host side -- marshalling of the camera, the scene arrays, the colour buffer, the upload-once rule,
 * error propagation -- can be exercised end to end on a box without a GPU (tests/test_cpp_host.py builds
 * it into a temporary directory and puts that first on LD_LIBRARY_PATH).  It is never built into, shipped
 * with or loaded by the product. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "splat.h"

typedef struct { float lowpass; int y_down; int zclip_mode; float sample_offset; int exp_mode; int nthreads; } orc_config;
void orc_compute_cov3d_all(const float *rot_xyzw, const float *scale3, uint64_t n, float *cov3d9);
int orc_render(const float *pos4, const float *cov3d9, const float *opacity, const float *sh48, uint64_t n, const void *cam,
               const orc_config *cfg, uint32_t *fb, uint32_t W, uint32_t H, uint32_t row0, uint32_t row1, void *st);

struct splat_ctx {
  splat_config cfg;
  float *pos4, *cov, *op, *sh;
  uint64_t n;
  int uploads, renders;
  char err[256];
};

static char g_create_error[256] = "";
static int g_pins = 0, g_pins_total = 0;

uint32_t splat_abi_version(void) { return SPLAT_ABI_VERSION; }
void splat_config_default(splat_config *c) {
  memset(c, 0, sizeof *c);
  c->lowpass = 0.3f; c->y_down = 0; c->zclip_mode = 1; c->sample_offset = 0.5f; c->tile = 16; c->near_cut = -1;
}
int splat_create(splat_ctx **out, const splat_config *cfg) {
  if (getenv("FAKE_SPLAT_FAIL_CREATE")) { snprintf(g_create_error, sizeof g_create_error, "no device (fake)"); return SPLAT_ERR_CUDA; }
  splat_ctx *c = (splat_ctx *)calloc(1, sizeof *c);
  c->cfg = *cfg;
  *out = c;
  return SPLAT_OK;
}
int splat_create_multi(splat_ctx **out, const splat_config *cfg, const int32_t *devices, int32_t n) {
  int rc = splat_create(out, cfg);
  if (rc) return rc;
  const char *log = getenv("FAKE_SPLAT_LOG");
  if (log) { FILE *f = fopen(log, "a"); if (f) { fprintf(f, "group of %d:", n); for (int i = 0; i < n; ++i) fprintf(f, " %d", devices[i]); fprintf(f, "\n"); fclose(f); } }
  return SPLAT_OK;
}
const char *splat_create_error(void) { return g_create_error; }
void splat_destroy(splat_ctx *c) {
  if (!c) return;
  const char *log = getenv("FAKE_SPLAT_LOG");
  if (log) { FILE *f = fopen(log, "a"); if (f) { fprintf(f, "uploads=%d renders=%d lowpass=%.2f pinned=%d still=%d\n", c->uploads, c->renders, c->cfg.lowpass, g_pins_total, g_pins); fclose(f); } }
  free(c->pos4); free(c->cov); free(c->op); free(c->sh); free(c);
}
const char *splat_last_error(const splat_ctx *c) { return c ? c->err : "null context"; }

static int store(splat_ctx *c, const float *pos4, const float *scale3, const float *op, const float *rot, const float *sh, uint64_t n) {
  free(c->pos4); free(c->cov); free(c->op); free(c->sh);
  const uint64_t m = n ? n : 1;
  c->pos4 = (float *)malloc(m * 16); c->cov = (float *)malloc(m * 36); c->op = (float *)malloc(m * 4); c->sh = (float *)malloc(m * 192);
  memcpy(c->pos4, pos4, n * 16); memcpy(c->op, op, n * 4); memcpy(c->sh, sh, n * 192);
  orc_compute_cov3d_all(rot, scale3, n, c->cov);
  c->n = n;
  c->uploads += 1;
  return SPLAT_OK;
}
int splat_upload_soa(splat_ctx *c, const float *pos4, const float *scale3, const float *opacity, const float *rot_xyzw, const float *sh48, uint64_t n) {
  if (!c || (n && (!pos4 || !scale3 || !opacity || !rot_xyzw || !sh48))) return SPLAT_ERR_INVALID;   /* n = 0: an empty scene */
  return store(c, pos4, scale3, opacity, rot_xyzw, sh48, n);
}
int splat_upload_aos(splat_ctx *c, const float *g59, uint64_t n) {
  if (!c || (n && !g59)) return SPLAT_ERR_INVALID;
  const uint64_t m = n ? n : 1;
  float *pos = (float *)malloc(m * 16), *sc = (float *)malloc(m * 12), *op = (float *)malloc(m * 4), *rot = (float *)malloc(m * 16), *sh = (float *)malloc(m * 192);
  for (uint64_t i = 0; i < n; ++i) {
    const float *g = g59 + 59 * i;
    memcpy(pos + 4 * i, g, 12); pos[4 * i + 3] = 1.0f;
    memcpy(sc + 3 * i, g + 3, 12);
    op[i] = g[6];
    memcpy(rot + 4 * i, g + 7, 16);
    memcpy(sh + 48 * i, g + 11, 192);
  }
  int rc = store(c, pos, sc, op, rot, sh, n);
  free(pos); free(sc); free(op); free(rot); free(sh);
  return rc;
}
int splat_render(splat_ctx *c, const splat_camera *cam, uint32_t *fb, uint32_t W, uint32_t H) {
  if (!c || !cam || !fb) return SPLAT_ERR_INVALID;
  if (!c->pos4) { snprintf(c->err, sizeof c->err, "render before upload"); return SPLAT_ERR_STATE; }
  if (cam->w != (float)W || cam->h != (float)H) { snprintf(c->err, sizeof c->err, "camera.w/h differ from the target size"); return SPLAT_ERR_UNSUPPORTED; }
  orc_config oc = {c->cfg.lowpass, c->cfg.y_down, c->cfg.zclip_mode, c->cfg.sample_offset, 0, 2};
  c->renders += 1;
  return orc_render(c->pos4, c->cov, c->op, c->sh, c->n, cam, &oc, fb, W, H, 0, H, NULL) ? SPLAT_ERR_CUDA : SPLAT_OK;
}
int splat_render_cleared(splat_ctx *c, const splat_camera *cam, uint32_t *fb, uint32_t W, uint32_t H, uint32_t clear) {
  if (!fb) return SPLAT_ERR_INVALID;
  for (uint64_t i = 0; i < (uint64_t)W * H; ++i) fb[i] = clear;
  return splat_render(c, cam, fb, W, H);
}
int splat_pin_host(void *p, uint64_t bytes) { (void)bytes; if (!p) return SPLAT_ERR_INVALID; g_pins += 1; g_pins_total += 1; return SPLAT_OK; }
int splat_unpin_host(void *p) { (void)p; g_pins -= 1; return SPLAT_OK; }
int splat_get_timings(splat_ctx *c, splat_timings *t) {
  if (!c || !t) return SPLAT_ERR_INVALID;
  memset(t, 0, sizeof *t);
  t->n_gaussians = c->n;
  return SPLAT_OK;
}
