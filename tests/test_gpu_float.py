"""SPLAT_BLEND_FLOAT: the un-quantised front-to-back compositor of the CUDA
library against the float CPU restatement (oracle.render_float mode 0 -- the same "over"
recurrence as the prototype's plot_opacity, to which mode 1 is pinned in
tests/test_reference_images.py, with the Rust fragment() rules).  Tolerance: 1e-4 RMSE per channel
on the un-quantised values (BASELINE north star); the 8-bit output may differ from the quantised
oracle image by one step where the float result sits on a truncation boundary."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import os

    from splat_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    _lib.load()
    return _lib


def _camera(W, H, pos, yaw=0.0):
    from splat_b200.camera import Camera

    cam = Camera(H, W, pos)
    cam.update_yaw_angle(yaw)
    cam.update_camera_pose()
    return cam


def _decode(fb):
    return np.stack([(fb >> 16) & 0xFF, (fb >> 8) & 0xFF, fb & 0xFF], axis=-1).astype(np.float32) / 255.0


CASES = [
    # name, n, seed, W, H, camera, yaw, log_scale_mean, noise framebuffer
    ("small_256", 1_000, 0x5EED0061, 256, 256, (0.0, 0.0, 5.0), 0.0, -3.0, False),
    ("ragged_onto_noise", 20_000, 0x5EED0062, 250, 130, (0.0, 0.0, 4.0), 0.4, -3.2, True),
    ("deep_lists", 150_000, 0x5EED0063, 160, 96, (0.0, 0.0, 3.0), 0.0, -2.6, True),
    ("demo_cam_100k_720p", 100_000, 0x5EED0064, 1280, 720, (-0.57651054, 2.99040512, -0.03924271), 0.0, -4.0, False),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_float_blend_matches_float_oracle(lib, orc, case):
    from splat_b200.gaussians import synthetic_scene

    name, n, seed, W, H, pos, yaw, lsm, noise = case
    scene = synthetic_scene(n, seed=seed, log_scale_mean=lsm)
    cam = _camera(W, H, pos, yaw)
    fb0 = np.zeros((H, W), np.uint32)
    if noise:
        fb0 = np.random.default_rng(seed).integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32)
    ctx = lib.Context(device=0, blend_mode=lib.SPLAT_BLEND_FLOAT)
    ctx.upload(scene)
    fb = fb0.copy()
    rgba = ctx.render_float(lib.camera_struct(cam), fb)
    t = ctx.timings()
    ctx.close()

    cfg = orc.make_config()
    sp = orc.project(scene, orc.camera_from(cam), cfg, W, H)
    order = orc.sort_visible(sp)
    assert t["n_visible"] == len(order)
    want = _decode(fb0)                              # composited onto the existing pixels
    acc = np.zeros((H, W), np.float32)
    orc.render_float(sp, order, cfg, want, acc, mode=0)

    touched = ~np.isnan(rgba[..., 3])
    assert np.array_equal(touched, acc > 0)          # exactly the pixels some fragment contributed to
    assert touched.sum() > W * H // 20
    got = np.where(touched[..., None], rgba[..., :3], _decode(fb0))
    rmse = np.sqrt(((got - want) ** 2).mean(axis=(0, 1)))
    assert rmse.max() < 1e-4, rmse                   # the north-star tolerance; measured ~1e-6
    assert np.abs(got - want).max() < 1e-3
    assert np.abs(rgba[..., 3][touched] - acc[touched]).max() < 1e-3
    # untouched pixels keep all four bytes; touched ones hold trunc(255 * value), alpha = trunc(255 * (1 - T))
    assert np.array_equal(fb[~touched], fb0[~touched])
    q = np.floor(np.clip(want, 0.0, 1.0) * 255.0)
    d = np.abs(_decode(fb) * 255.0 - q)[touched]
    assert d.max() <= 1.0 and (d > 0).mean() < 0.02
    a8 = (fb >> 24).astype(np.float32)[touched]
    assert np.abs(a8 - np.floor(np.clip(acc[touched], 0, 1) * 255.0)).max() <= 1.0


def test_float_and_reference_blends_differ_only_by_the_truncation_bias(lib, orc):
    """Same scene through both blend modes: the reference blend truncates after every Gaussian
    (6-9e-3 RMSE darker than float compositing); nothing else distinguishes them."""
    from splat_b200.gaussians import synthetic_scene

    W, H = 400, 300
    scene = synthetic_scene(30_000, seed=0x5EED0065, log_scale_mean=-3.3)
    cam = _camera(W, H, (0.0, 0.0, 4.0), 0.2)
    out = {}
    for mode in (lib.SPLAT_BLEND_REFERENCE, lib.SPLAT_BLEND_FLOAT):
        ctx = lib.Context(device=0, blend_mode=mode)
        ctx.upload(scene)
        fb = np.zeros((H, W), np.uint32)
        ctx.render(lib.camera_struct(cam), fb)
        out[mode] = _decode(fb)
        ctx.close()
    diff = out[lib.SPLAT_BLEND_FLOAT] - out[lib.SPLAT_BLEND_REFERENCE]
    rmse = float(np.sqrt((diff ** 2).mean()))
    assert 1e-4 < rmse < 2.5e-2, rmse
    assert diff.mean() > 0          # the per-layer truncation only ever loses light


def test_float_mode_rejects_the_near_cut(lib):
    with pytest.raises(lib.SplatError) as e:
        lib.Context(device=0, blend_mode=lib.SPLAT_BLEND_FLOAT, near_cut=128)
    assert e.value.code == -4 and "near cut" in str(e.value)
