"""Multi-GPU behind the C ABI (include/splat.h "Multi-GPU"): one process driving several
devices through a GROUP context (splat_create_multi: ncclCommInitAll, scene broadcast, per-frame
stripe gather).  Needs >= 2 GPUs: skipped on a one-GPU box (run under `gpurun --gpus 2`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import os

    import torch

    from splat_b200 import _lib

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
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


@pytest.mark.parametrize("equal", [True, False], ids=["equal_stripes", "rebalanced"])
def test_group_context_equals_single_device_and_oracle(lib, orc, equal):
    import torch

    from splat_b200.gaussians import synthetic_scene

    G = min(torch.cuda.device_count(), 8)
    W, H = 640, 360
    scene = synthetic_scene(120_000, seed=0x5EED0081, log_scale_mean=-3.6)
    grp = lib.Context(devices=list(range(G)), equal_stripes=equal)
    grp.upload(scene)
    one = lib.Context(device=0)
    one.upload(scene)
    cfg = orc.make_config()
    rng = np.random.default_rng(81)
    for k, yaw in enumerate(np.linspace(0.0, 1.2, 12)):
        cam = _camera(W, H, (0.0, 0.0, 4.0), yaw=float(yaw))
        fb0 = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) if k % 3 == 2 else np.zeros((H, W), np.uint32)
        want = fb0.copy()
        one.render(lib.camera_struct(cam), want)
        got = fb0.copy()
        grp.render(lib.camera_struct(cam), got)
        assert np.array_equal(got, want), (k, int(np.count_nonzero(got != want)))
        if k in (0, 7):
            ref = fb0.copy()
            orc.render(scene, orc.camera_from(cam), cfg, ref)
            assert np.array_equal(got, ref)
        # the fused clear + render entry point
        got2 = np.full((H, W), 0xDEADBEEF, np.uint32)
        grp.render_cleared(lib.camera_struct(cam), got2, 0)
        want2 = np.zeros((H, W), np.uint32)
        one.render(lib.camera_struct(cam), want2)
        assert np.array_equal(got2, want2), k
    b = grp.group_bounds()
    assert b[0][0] == 0 and b[-1][1] == H and all(b[i][1] == b[i + 1][0] for i in range(G - 1))
    if equal:
        trows = [(r1 + 15) // 16 - r0 // 16 for r0, r1 in b]
        assert max(trows) - min(trows) <= 1
    t = grp.timings()
    assert t["n_instances"] > 0 and t["n_gaussians"] == scene.num_gaussians
    grp.close()
    one.close()


def test_rank_contexts_gather_in_one_process(lib, orc):
    """The one-process-per-GPU entry points (splat_comm_init_rank / splat_comm_broadcast_scene /
    splat_gather_stripes) cannot be driven by a single thread for several ranks (ncclCommInitRank blocks
    until all ranks joined), so here each rank gets a thread -- the same calls a torchrun launcher makes."""
    import threading

    import torch

    from splat_b200 import stripes
    from splat_b200.gaussians import synthetic_scene

    G = 2
    W, H = 500, 330
    scene = synthetic_scene(30_000, seed=0x5EED0082, log_scale_mean=-3.4)
    cam = _camera(W, H, (0.0, 0.0, 4.0), yaw=0.3)
    bounds = stripes.stripe_bounds(H, G)
    ctxs = [lib.Context(device=r, near_cut=0) for r in range(G)]
    uid = ctxs[0].unique_id()
    frames = [None] * G
    errs = []

    def run(r):
        try:
            torch.cuda.set_device(r)
            c = ctxs[r]
            c.comm_init(uid, G, r)
            if r == 0:
                c.upload(scene)
            c.broadcast_scene(0, scene.num_gaussians)
            fb = torch.zeros((H, W), dtype=torch.int32, device=torch.device("cuda", r))
            torch.cuda.synchronize()
            r0, r1 = bounds[r]
            for _ in range(3):
                fb.zero_()
                torch.cuda.synchronize()
                c.render_device(lib.camera_struct(cam), fb[r0:r1].data_ptr(), W, H, r0, r1)
                c.gather_stripes(fb.data_ptr(), W, H, bounds, 0)
                c.timings()
            torch.cuda.synchronize()
            frames[r] = fb.cpu().numpy().view(np.uint32)
        except Exception as e:   # noqa: BLE001
            errs.append((r, repr(e)))

    th = [threading.Thread(target=run, args=(r,)) for r in range(G)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert not errs, errs
    ref = np.zeros((H, W), np.uint32)
    orc.render(scene, orc.camera_from(cam), orc.make_config(), ref)
    assert np.array_equal(frames[0], ref), int(np.count_nonzero(frames[0] != ref))
    for c in ctxs:
        c.close()
