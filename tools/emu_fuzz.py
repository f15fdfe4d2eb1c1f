#!/usr/bin/env python3
"""Randomised parity runs of the library's host-emulated build (tests/cuda_emu) against the oracle: random image
sizes (ragged tiles), cameras (far, near, inside the cloud, yawed / pitched), scene statistics, pipelines, euc
switches, near-cut fractions, stripe partitions, frames blended onto noise, several frames per context so that the
no-round-trip path and its launch bounds are exercised.  No GPU needed.  Usage: tools/emu_fuzz.py [first_seed] [count] [reference|float|group]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))


def one_case(seed, lib, orc):
    from splat_b200.camera import Camera
    from splat_b200.gaussians import synthetic_scene

    rng = np.random.default_rng(seed)
    big = os.environ.get("EMU_FUZZ_BIG") == "1"          # larger images and scenes: minutes per case
    W, H = (int(rng.integers(300, 1300)), int(rng.integers(200, 760))) if big else (int(rng.integers(17, 420)), int(rng.integers(17, 300)))
    n = int(rng.choice([20000, 80000, 250000, 600000] if big else [1, 7, 300, 2000, 9000, 30000]))
    lsm = float(rng.uniform(-4.5, -2.6) if big else rng.uniform(-4.5, -1.8))
    scene = synthetic_scene(n, seed=0x5EED0000 + seed, log_scale_mean=lsm)
    if rng.random() < 0.3:
        scene.opacities[:] = rng.uniform(0.005, 0.08)                     # faint: pixels that never converge
    if rng.random() < 0.2 and n > 10:
        bad = rng.integers(0, n, size=max(1, n // 50))
        scene.scales[bad] = 0.0                                            # degenerate covariances
    lowpass = float(rng.choice([0.3, 0.01]))
    y_down, zclip = int(rng.integers(0, 2)), int(rng.integers(0, 3))
    near_cut = int(rng.choice([-1, 0, 0, 1, 16, 128, 700]))
    sync = int(rng.random() < 0.25)
    ctx = lib.Context(device=0, lowpass=lowpass, y_down=y_down, zclip_mode=zclip, near_cut=near_cut, sync_frames=sync)
    ctx.upload(scene)
    cfg = orc.make_config(lowpass=lowpass, y_down=y_down, zclip_mode=zclip, nthreads=4)
    # stripes: tile-aligned partition of [0, H)
    trows = (H + 15) // 16
    if rng.random() < 0.5 and trows > 1:
        cuts = sorted(set(rng.integers(1, trows, size=int(rng.integers(1, 4))).tolist()))
        rows = [(a * 16, min(b * 16, H)) for a, b in zip([0] + cuts, cuts + [trows])]
    else:
        rows = [(0, H)]
    desc = dict(seed=seed, W=W, H=H, n=n, lsm=round(lsm, 2), lowpass=lowpass, y_down=y_down, zclip=zclip, near_cut=near_cut, sync=sync, rows=rows)
    frames = int(rng.integers(1, 5))
    for k in range(frames):
        dist = float(rng.choice([0.3, 1.0, 2.5, 5.0, 12.0]))
        cam = Camera(H, W, (float(rng.normal(0, 0.3)), float(rng.normal(0, 0.3)), dist))
        cam.update_yaw_angle(float(rng.uniform(-3, 3)))
        cam.update_pitch_angle(float(rng.uniform(-0.6, 0.6)))
        cam.update_camera_pose()
        fb0 = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) if rng.random() < 0.4 else np.zeros((H, W), np.uint32)
        want = fb0.copy()
        orc.render(scene, orc.camera_from(cam), cfg, want)
        got = fb0.copy()
        cs = lib.camera_struct(cam)
        for r0, r1 in rows:
            part = np.ascontiguousarray(got[r0:r1])
            ctx.render(cs, part, r0, r1)
            got[r0:r1] = part
        bad = int(np.count_nonzero(got != want))
        if bad:
            ctx.close()
            return False, dict(desc, frame=k, dist=dist, mismatching=bad)
    ctx.close()
    return True, desc


def one_float_case(seed, lib, orc):
    """SPLAT_BLEND_FLOAT against the float oracle (mode 0): same touched pixels, RMSE <= 1e-4"""
    from splat_b200.camera import Camera
    from splat_b200.gaussians import synthetic_scene

    rng = np.random.default_rng(10_000 + seed)
    W, H = int(rng.integers(17, 360)), int(rng.integers(17, 260))
    n = int(rng.choice([5, 400, 3000, 20000]))
    lsm = float(rng.uniform(-4.2, -2.0))
    scene = synthetic_scene(n, seed=0x5EED1000 + seed, log_scale_mean=lsm)
    if rng.random() < 0.3:
        scene.opacities[:] = rng.uniform(0.005, 0.08)
    cam = Camera(H, W, (float(rng.normal(0, 0.3)), float(rng.normal(0, 0.3)), float(rng.choice([1.0, 2.5, 5.0, 12.0]))))
    cam.update_yaw_angle(float(rng.uniform(-3, 3)))
    cam.update_camera_pose()
    fb0 = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) if rng.random() < 0.5 else np.zeros((H, W), np.uint32)
    desc = dict(mode="float", seed=seed, W=W, H=H, n=n, lsm=round(lsm, 2))

    def decode(fb):
        return np.stack([((fb >> sh) & 0xFF).astype(np.float32) / np.float32(255.0) for sh in (16, 8, 0)], axis=-1)

    ctx = lib.Context(device=0, blend_mode=lib.SPLAT_BLEND_FLOAT)
    ctx.upload(scene)
    fb = fb0.copy()
    rgba = ctx.render_float(lib.camera_struct(cam), fb)
    ctx.close()
    cfg = orc.make_config()
    sp = orc.project(scene, orc.camera_from(cam), cfg, W, H)
    order = orc.sort_visible(sp)
    want = decode(fb0)
    acc = np.zeros((H, W), np.float32)
    orc.render_float(sp, order, cfg, want, acc, mode=0)
    touched = ~np.isnan(rgba[..., 3])
    if not np.array_equal(touched, acc > 0):
        return False, dict(desc, why="touched sets differ", n_px=int(np.count_nonzero(touched != (acc > 0))))
    got = np.where(touched[..., None], rgba[..., :3], decode(fb0))
    rmse = float(np.sqrt(((got - want) ** 2).mean(axis=(0, 1))).max())
    if rmse >= 1e-4 or not np.array_equal(fb[~touched], fb0[~touched]):
        return False, dict(desc, why="rmse / untouched pixels", rmse=rmse)
    return True, desc


def one_group_case(seed, lib, orc):
    """a group context over 2..4 emulated devices against a single device, several frames (re-cut stripes)"""
    from splat_b200.camera import Camera
    from splat_b200.gaussians import synthetic_scene

    rng = np.random.default_rng(20_000 + seed)
    W, H = int(rng.integers(40, 300)), int(rng.integers(40, 240))
    n = int(rng.choice([50, 2000, 12000]))
    G = int(rng.integers(2, 5))
    scene = synthetic_scene(n, seed=0x5EED2000 + seed, log_scale_mean=float(rng.uniform(-4.0, -2.5)))
    desc = dict(mode="group", seed=seed, W=W, H=H, n=n, G=G)
    grp = lib.Context(devices=list(range(G)), equal_stripes=bool(rng.random() < 0.3))
    one = lib.Context(device=0)
    grp.upload(scene)
    one.upload(scene)
    for k in range(int(rng.integers(2, 6))):
        cam = Camera(H, W, (0.0, 0.0, float(rng.choice([1.5, 4.0, 9.0]))))
        cam.update_yaw_angle(float(rng.uniform(-3, 3)))
        cam.update_camera_pose()
        fb0 = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) if rng.random() < 0.4 else np.zeros((H, W), np.uint32)
        want, got = fb0.copy(), fb0.copy()
        one.render(lib.camera_struct(cam), want)
        grp.render(lib.camera_struct(cam), got)
        if not np.array_equal(got, want):
            return False, dict(desc, frame=k, mismatching=int(np.count_nonzero(got != want)), bounds=grp.group_bounds())
    grp.close()
    one.close()
    return True, desc


def main():
    import emu_build
    from oracle import oracle as orc
    from splat_b200 import _lib

    orc.build()
    _lib.LIB_PATH, _lib._lib = emu_build.build(), None
    _lib.load()
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    case = {"reference": one_case, "float": one_float_case, "group": one_group_case}[sys.argv[3] if len(sys.argv) > 3 else "reference"]
    t0, fails = time.time(), 0
    for seed in range(first, first + count):
        ok, d = case(seed, _lib, orc)
        if not ok:
            fails += 1
            print("MISMATCH", d, flush=True)
        elif seed % 10 == 0:
            print(f"seed {seed} ok ({time.time() - t0:.0f}s)", flush=True)
    print(f"{count} cases, {fails} failed, {time.time() - t0:.0f}s")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
