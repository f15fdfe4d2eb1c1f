#!/usr/bin/env python3
"""AddressSanitizer pass over the host-emulated library (tests/cuda_emu): near-cut frames (open tiles and fall-back), stripe frames
(pre-pass and dense paths, with and without host round trips), the float blend, device PLY ingest -- each compared with the
oracle.  Run as:  EMU_ASAN=1 LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/emu_asan_cases.py
(about half an hour on 8 cores; the sanitizer stops at the first bad access)."""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests', 'cuda_emu'), os.path.join(ROOT, 'tests')]
import numpy as np
import emu_build
from splat_b200 import _lib
_lib.LIB_PATH = emu_build.build(); _lib._lib = None
from oracle import oracle as orc
from splat_b200.camera import Camera
from splat_b200.gaussians import synthetic_scene, save_ply
orc.build()
def cam_(W, H, z, yaw=0.0):
    c = Camera(H, W, (0.0, 0.0, z)); c.update_yaw_angle(yaw); c.update_camera_pose(); return c
t0 = time.time()
# 1. near cut with open tiles + fallback, several frames, ragged size
W, H = 150, 90
scene = synthetic_scene(6000, seed=0x5EED0040, log_scale_mean=-2.8)
for frac in (16, 1):
    ctx = _lib.Context(device=0, near_cut=frac); ctx.upload(scene)
    for k, (z, yaw) in enumerate([(4.0, 0.0), (4.0, 0.3), (2.0, 0.6)]):
        cam = cam_(W, H, z, yaw)
        fb = np.zeros((H, W), np.uint32); ctx.render(_lib.camera_struct(cam), fb)
        ref = np.zeros((H, W), np.uint32); orc.render(scene, orc.camera_from(cam), orc.make_config(), ref)
        print("near_cut", frac, k, int(np.count_nonzero(fb != ref)), f"{time.time()-t0:.0f}s", flush=True)
    ctx.close()
# 2. stripes (pre-pass and dense paths), sync and async
for sync in (0, 1):
    ctx = _lib.Context(device=0, sync_frames=sync, near_cut=0); ctx.upload(scene)
    for k in range(3):
        cam = cam_(W, H, 3.0, 0.2 * k)
        ref = np.zeros((H, W), np.uint32); orc.render(scene, orc.camera_from(cam), orc.make_config(), ref)
        for r0, r1 in ((32, 64), (0, 16)):
            part = np.zeros((r1 - r0, W), np.uint32); ctx.render(_lib.camera_struct(cam), part, r0, r1)
            print("stripe", sync, k, (r0, r1), int(np.count_nonzero(part != ref[r0:r1])), f"{time.time()-t0:.0f}s", flush=True)
    ctx.close()
# 3. float mode
ctx = _lib.Context(device=0, blend_mode=_lib.SPLAT_BLEND_FLOAT); ctx.upload(scene)
fb = np.zeros((H, W), np.uint32); rgba = ctx.render_float(_lib.camera_struct(cam_(W, H, 4.0)), fb); ctx.close()
print("float touched", int((~np.isnan(rgba[..., 3])).sum()), f"{time.time()-t0:.0f}s", flush=True)
# 4. device PLY ingest
rng = np.random.default_rng(1); n = 3000
raw = {k: rng.normal(0, 0.5, n).astype(np.float32) for k in ("x", "y", "z", "opacity", "rot_0", "rot_1", "rot_2", "rot_3")}
for i in range(3): raw[f"scale_{i}"] = rng.normal(-3.0, 0.5, n).astype(np.float32); raw[f"f_dc_{i}"] = rng.normal(0, 1, n).astype(np.float32)
save_ply("/tmp/asan_scene.ply", raw)
ctx = _lib.Context(device=0); dev = ctx.upload_ply("/tmp/asan_scene.ply", want_activated=True)
fb = np.zeros((H, W), np.uint32); ctx.render(_lib.camera_struct(cam_(W, H, 3.0)), fb); ctx.close()
print("ply", dev.num_gaussians, int(np.count_nonzero(fb)), f"{time.time()-t0:.0f}s", flush=True)
print("ASAN CASES DONE")
