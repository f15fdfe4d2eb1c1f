"""GPU edge cases added after round 2's GPU budget had ended: an empty scene, and a scene of which nothing is visible.
They have run on the host-emulated build of the library (tests/test_emu_parity.py) but not yet on a device, which is
why they live in a file that sorts last: everything that HAS run on a B200 is tried before them."""
import numpy as np
import pytest

from test_gpu_parity import _camera, _scene, lib  # noqa: F401

pytestmark = pytest.mark.gpu


def test_empty_scene_renders_nothing(lib, orc):
    """An empty Vec<Gaussian> / GaussianList is valid input for render_to_buffer (pipelines.rs:66-86: sort of nothing,
    no instances): the buffer keeps its contents.  Uploading zero Gaussians is therefore not an error; frames of
    an empty scene check their arguments and draw nothing, and a later upload on the same context renders normally."""
    from splat_b200.gaussians import GaussianList

    W, H = 200, 120
    cam = _camera(W, H, (0.0, 0.0, 5.0))
    cs = lib.camera_struct(cam)
    empty = GaussianList(np.zeros((0, 4), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.float32),
                         np.zeros((0, 4), np.float32), np.zeros((0, 48), np.float32))
    noise = np.random.default_rng(9).integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32)
    ctx = lib.Context(device=0)
    ctx.upload(empty)
    fb = noise.copy()
    ctx.render(cs, fb)
    ref = noise.copy()
    orc.render(empty, orc.camera_from(cam), orc.make_config(), ref)
    assert np.array_equal(ref, noise) and np.array_equal(fb, noise)
    part = np.ascontiguousarray(noise[16:48])
    ctx.render(cs, part, 16, 48)                                   # a stripe
    assert np.array_equal(part, noise[16:48])
    fb = noise.copy()
    ctx._check(ctx.L.splat_render_cleared(ctx.h, cs, fb.ctypes.data, W, H, 0x11223344))
    assert np.all(fb == 0x11223344)
    t = ctx.timings()
    assert t["n_gaussians"] == 0 and t["n_instances"] == 0
    with pytest.raises(lib.SplatError) as e:                       # arguments are still checked
        ctx._check(ctx.L.splat_render(ctx.h, lib.camera_struct(_camera(64, 64, (0, 0, 5))), fb.ctypes.data, W, H))
    assert e.value.code == -4
    ctx.upload_aos(np.zeros((0, 59), np.float32))                  # Pipeline01's Vec<Gaussian>, empty
    fb = noise.copy()
    ctx.render(cs, fb)
    assert np.array_equal(fb, noise)
    scene = _scene(500, 0x5EED0095, -3.0)                          # the context is not stuck in the empty state
    ctx.upload(scene)
    fb = np.zeros((H, W), np.uint32)
    ctx.render(cs, fb)
    ref = np.zeros((H, W), np.uint32)
    orc.render(scene, orc.camera_from(cam), orc.make_config(), ref)
    assert np.count_nonzero(ref) > 0 and np.array_equal(fb, ref)
    ctx.upload(empty)                                              # ... and back
    fb = noise.copy()
    ctx.render(cs, fb)
    assert np.array_equal(fb, noise)
    ctx.close()


@pytest.mark.parametrize("near_cut,sync_frames", [(0, 0), (-1, 0), (64, 0), (0, 1)])
def test_everything_culled_leaves_the_buffer_untouched(lib, orc, near_cut, sync_frames):
    """A scene entirely behind the camera: no visible Gaussian, no tile instance, on the first frame of a geometry
    and on the frames after it (launches sized from a count of zero), whole frames and stripes."""
    W, H = 200, 120
    scene = _scene(3000, 0x5EED0096, -3.0)
    scene.positions[:, 2] += 20.0
    noise = np.random.default_rng(1).integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32)
    ctx = lib.Context(device=0, near_cut=near_cut, sync_frames=sync_frames)
    ctx.upload(scene)
    for k in range(3):
        cam = _camera(W, H, (0.0, 0.0, 5.0), yaw=0.1 * k)
        ref = noise.copy()
        st = orc.render(scene, orc.camera_from(cam), orc.make_config(), ref)
        assert st.n_visible == 0 and np.array_equal(ref, noise)
        fb = noise.copy()
        ctx.render(lib.camera_struct(cam), fb)
        assert np.array_equal(fb, noise)
        t = ctx.timings()
        assert t["n_visible"] == 0 and t["n_instances"] == 0
        part = np.ascontiguousarray(noise[32:64])
        ctx.render(lib.camera_struct(cam), part, 32, 64)
        assert np.array_equal(part, noise[32:64])
    ctx.close()
