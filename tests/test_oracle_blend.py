"""Closed-form checks of the oracle's fragment()+blend() restatement (pipelines.rs:127-168)
-- the known answers listed in SURVEY.md section 8c -- and of its rasteriser conventions
(E3/E4/E7/E8) on images small enough to reason about by hand."""
import types

import numpy as np


def _one(orc, old, alpha_opacity, rgb, dx=0.0, dy=0.0, conic=(1.0, 0.0, 1.0)):
    return orc.shade_blend(old, conic[0], conic[1], conic[2], dx, dy, alpha_opacity, rgb)


def test_blend_onto_black_is_truncated_product(orc):
    # known answer 3: byte = trunc(alpha * c * 255), alpha byte = trunc(alpha * 255)
    for a in (0.25, 0.5, 0.731, 0.99):
        for c in (0.1, 0.5, 0.999, 1.0):
            px = _one(orc, 0, a, (c, c / 2, 0.0))
            al = np.float32(a)
            want_r = int(np.float32(np.float32(al * np.float32(c))) * np.float32(255.0))
            want_g = int(np.float32(al * np.float32(c / 2)) * np.float32(255.0))
            assert (px >> 16) & 0xFF == want_r and (px >> 8) & 0xFF == want_g and px & 0xFF == 0
            assert px >> 24 == int(al * np.float32(255.0))


def test_opacity_is_clamped_to_099(orc):
    px = _one(orc, 0, 1.0, (1.0, 1.0, 1.0))
    assert px >> 24 == int(np.float32(0.99) * np.float32(255.0)) == 252
    assert (px >> 16) & 0xFF == 252


def test_zero_fragment_keeps_rgb_and_resets_alpha(orc):
    # E7: alpha below 1/255 -> Fragment::zeros() is still blended
    old = 0xAB123456
    px = _one(orc, old, 0.5, (1.0, 1.0, 1.0), dx=10.0)      # exp(-50) -> alpha < 1/255
    assert px == 0x00123456
    px = _one(orc, old, 0.003, (1.0, 1.0, 1.0))             # opacity itself below 1/255
    assert px == 0x00123456
    # power > 0 (indefinite conic) -> zeros as well (pipelines.rs:135)
    px = _one(orc, old, 0.9, (1.0, 1.0, 1.0), dx=1.0, dy=1.0, conic=(0.1, -5.0, 0.1))
    assert px == 0x00123456


def test_saturating_cast_for_out_of_gamut_colour(orc):
    # colour is not clamped (gaussians.rs:97); `as u8` saturates and maps negatives to 0
    px = _one(orc, 0x00808080, 0.9, (3.0, -2.0, 1.0000086))
    assert (px >> 16) & 0xFF == 255 and (px >> 8) & 0xFF == 0
    px = _one(orc, 0, 0.99, (float("nan"), 0.5, 0.5))
    assert (px >> 16) & 0xFF == 0                             # NaN as u8 == 0


def test_per_layer_truncation_differs_from_float_compositing(orc):
    """F4: truncating after every layer biases the result low -- the reason the CUDA kernel
    reproduces the recurrence instead of a transmittance scan."""
    rng = np.random.default_rng(5)
    err = []
    for _ in range(200):
        px, acc = 0, np.zeros(3)
        for _ in range(20):
            a, c = rng.uniform(0.05, 0.6), rng.uniform(0, 1, 3)
            px = _one(orc, px, a, tuple(c))
            acc = (1 - a) * acc + a * c
        got = np.array([(px >> 16) & 0xFF, (px >> 8) & 0xFF, px & 0xFF]) / 255.0
        err.append(got - acc)
    err = np.array(err)
    assert err.mean() < -0.003 and np.sqrt((err ** 2).mean()) > 0.004


def _single(pos, scale, opacity=1.0, rgb=(1.0, 1.0, 1.0)):
    p = np.ones((1, 4), np.float32); p[0, :3] = pos
    sh = np.zeros((1, 48), np.float32)
    sh[0, :3] = (np.array(rgb, np.float32) - np.float32(0.5)) / np.float32(0.28209479177387814)
    return types.SimpleNamespace(positions=p, scales=np.full((1, 3), scale, np.float32),
                                 opacities=np.array([opacity], np.float32),
                                 rotations=np.array([[0, 0, 0, 1]], np.float32), sh=sh)


def _cam(orc, W, H, pos=(0.0, 0.0, 5.0)):
    from splat_b200.camera import Camera

    cam = Camera(H, W, pos)
    cam.update_camera_pose()
    return orc.camera_from(cam)


def test_single_gaussian_image_by_hand(orc):
    """One isotropic Gaussian at the origin seen from (0,0,5) at 64x64: centre at (32,32),
    cov2d = (focal/z)^2 s^2 + lowpass with focal = 32; the quad is the axis-aligned 3-sigma
    rect sampled at pixel centres (E3/E4), rows are y*W+x (E8)."""
    W = H = 64
    s, lowpass = 0.5, 0.3
    fb = np.zeros((H, W), np.uint32)
    st = orc.render(_single((0, 0, 0), s), _cam(orc, W, H), orc.make_config(lowpass=lowpass), fb)
    var = np.float32((32.0 / 5.0) ** 2 * s * s + lowpass)
    hx = 3.0 * np.sqrt(var)
    ys, xs = np.nonzero(fb >> 24 | (fb & 0xFFFFFF))
    assert st.n_visible == 1
    # covered pixels: |x + 0.5 - 32| <= hx
    inside = np.abs(np.arange(W) + 0.5 - 32.0) <= hx
    assert st.pairs_in_rect == int(inside.sum()) ** 2
    # symmetric about the centre, peak at the four centre pixels
    assert np.array_equal(fb, fb[::-1, :]) and np.array_equal(fb, fb[:, ::-1])
    d2 = 0.5 * 0.5 * 2
    a = min(0.99, float(np.exp(-0.5 * d2 / var)))
    assert abs(int(fb[32, 32] >> 24) - int(a * 255)) <= 1
    assert xs.min() >= 32 - np.ceil(hx) - 1 and xs.max() <= 32 + np.ceil(hx)


def test_far_to_near_order_and_view_depth(orc):
    # known answer 5: view z of the origin from (0,0,5) is -5; nearer Gaussians are drawn later
    W = H = 64
    cam = _cam(orc, W, H)
    red, green = _single((0, 0, 0), 0.3, 1.0, (1, 0, 0)), _single((0, 0, 2), 0.3, 1.0, (0, 1, 0))
    sc = types.SimpleNamespace(**{k: np.concatenate([getattr(green, k), getattr(red, k)])
                                  for k in ("positions", "scales", "opacities", "rotations", "sh")})
    sp = orc.project(sc, cam, orc.make_config(), W, H)
    assert np.isclose(sp["z_view"][1], -5.0) and np.isclose(sp["z_view"][0], -3.0)
    assert list(orc.sort_visible(sp)) == [1, 0]            # far (red) first, near (green) last
    fb = np.zeros((H, W), np.uint32)
    orc.render(sc, cam, orc.make_config(), fb)
    r, g = (fb[32, 32] >> 16) & 0xFF, (fb[32, 32] >> 8) & 0xFF
    assert g > 200 and r < 10                               # green (near) over red (far)


def test_naive_scene_colours_decode(orc):
    # known answer 6: (c-0.5)/0.28209 with SH_C0 = 0.2820948 decodes 1.0 to just above 1
    from splat_b200.gaussians import GaussianList

    sc = GaussianList.naive_gaussians()
    sp = orc.project(sc, _cam(orc, 1280, 720, (0, 0, 3)), orc.make_config(), 1280, 720)
    want = np.array([[1, 0, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    assert np.allclose(sp["color"], want, atol=2e-5)
    assert (sp["color"][want == 1] > 1.0).all()


def test_glibc_exp_variant_bounds_the_pinned_exp(orc):
    """exp_mode=1 renders with glibc expf instead of the pinned routine: the images differ in
    well under 1% of the pixels, by one LSB -- the distance between two correct libms, which
    is all that separates the pinned exp from the platform exp Rust would call."""
    from splat_b200.gaussians import synthetic_scene

    W, H = 320, 240
    sc = synthetic_scene(3000, seed=0x5EED0040, log_scale_mean=-3.0)
    cam = _cam(orc, W, H)
    a, b = np.zeros((H, W), np.uint32), np.zeros((H, W), np.uint32)
    orc.render(sc, cam, orc.make_config(exp_mode=0), a)
    orc.render(sc, cam, orc.make_config(exp_mode=1), b)
    ch = lambda f: np.stack([(f >> 16) & 0xFF, (f >> 8) & 0xFF, f & 0xFF], -1).astype(np.int64)
    d = ch(a) - ch(b)
    assert np.abs(d).max() <= 1
    rmse = np.sqrt((d.astype(np.float64) ** 2).mean()) / 255.0
    assert rmse < 1e-4, rmse        # the north-star bar (measured: 0 .. 1.6e-5)
