"""PLY ingest: the loader mirrors load_from_ply / set_property
(gaussians.rs:258-282, :375-405) -- activations, quaternion component order, the
un-transposed f_rest mapping and the sequential-f32 mean recentering."""
import numpy as np
import pytest

from splat_b200.gaussians import GaussianList, load_from_ply, load_ply_soa, save_ply, _PLY_PROPS


def _raw(n, seed=0):
    rng = np.random.default_rng(seed)
    raw = {p: rng.normal(size=n).astype(np.float32) for p in _PLY_PROPS}
    raw["x"] += 3.0
    return raw


def test_schema_is_the_inria_62_float_layout():
    assert len(_PLY_PROPS) == 62 and _PLY_PROPS[:3] == ["x", "y", "z"] and _PLY_PROPS[-4:] == ["rot_0", "rot_1", "rot_2", "rot_3"]


def test_roundtrip_activations_and_layout(tmp_path):
    n = 257
    raw = _raw(n)
    path = str(tmp_path / "pc.ply")
    save_ply(path, raw)
    g = load_ply_soa(path)
    assert isinstance(g, GaussianList) and g.num_gaussians == n
    for i in range(3):
        assert np.array_equal(g.scales[:, i], np.exp(raw[f"scale_{i}"]).astype(np.float32))      # :265-267
    assert np.allclose(g.opacities, 1 / (1 + np.exp(-raw["opacity"].astype(np.float64))), rtol=1e-6)  # :268
    # rot_0 is w and lands in the LAST slot (nalgebra coords i,j,k,w) :269-272
    assert np.array_equal(g.rotations[:, 3], raw["rot_0"]) and np.array_equal(g.rotations[:, 0], raw["rot_1"])
    # f_dc -> sh[0..3], f_rest_i -> sh[3+i], no channel-major transpose :276-279
    assert np.array_equal(g.sh[:, 1], raw["f_dc_1"]) and np.array_equal(g.sh[:, 3 + 17], raw["f_rest_17"])
    assert np.all(g.positions[:, 3] == 1.0)


def test_recentring_is_a_sequential_f32_sum(tmp_path):
    n = 5000
    raw = _raw(n, 1)
    path = str(tmp_path / "pc.ply")
    save_ply(path, raw)
    g = load_ply_soa(path)
    acc = np.float32(0.0)
    for v in raw["x"]:
        acc = np.float32(acc + v)                        # gaussians.rs:395-397
    avg = np.float32(acc / np.float32(n))
    assert np.array_equal(g.positions[:, 0], raw["x"] - avg)


def test_aos_view_and_ascii(tmp_path):
    raw = _raw(3, 2)
    path = str(tmp_path / "pc.ply")
    save_ply(path, raw)
    gs = load_from_ply(path)
    assert len(gs) == 3 and gs[0].sh.shape == (48,) and gs[0].rotation.shape == (4,)
    header = "ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\nproperty float opacity\nend_header\n"
    p2 = tmp_path / "a.ply"
    p2.write_text(header + "1 2 3 0\n3 2 1 0\n")
    g = load_ply_soa(str(p2))
    assert np.allclose(g.positions[:, :3], [[-1, 0, 1], [1, 0, -1]]) and np.allclose(g.opacities, 0.5)
    assert np.array_equal(g.rotations, [[0, 0, 0, 1], [0, 0, 0, 1]])   # Quaternion::identity()


def test_unexpected_element_is_an_error(tmp_path):
    p = tmp_path / "bad.ply"
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement face 0\nend_header\n")
    with pytest.raises(ValueError):
        load_ply_soa(str(p))                             # gaussians.rs:390 panics


def test_trim_keeps_the_first_vertices(tmp_path):
    """`trim` (00_ply_load.rs): first three vertices, same schema."""
    from splat_b200.gaussians import load_ply_soa, save_ply, trim_ply

    raw = _raw(10, seed=3)
    src, dst = str(tmp_path / "a.ply"), str(tmp_path / "b.ply")
    save_ply(src, raw)
    assert trim_ply(src, dst, 3) == 3
    a, b = load_ply_soa(src), load_ply_soa(dst)
    assert b.num_gaussians == 3
    # activations are per vertex; only the recentring (mean over the file) differs
    assert np.array_equal(a.scales[:3], b.scales[:3]) and np.array_equal(a.sh[:3], b.sh[:3])
    assert np.array_equal(a.opacities[:3], b.opacities[:3]) and np.array_equal(a.rotations[:3], b.rotations[:3])


# --------------------------------------------------------------------------- device ingest (f-1)
@pytest.mark.gpu
def test_device_ply_ingest_matches_host_loader_and_renders_identically(tmp_path):
    """splat_upload_ply_raw: the raw vertex payload is activated ON THE DEVICE (exp / sigmoid /
    rot_0 -> w / f_rest_i -> sh[3+i]) and recentred with the reference's sequential f32 mean
    (gaussians.rs:258-282, :394-402).  Against the numpy host loader: positions, rotations and SH
    are bit-identical; scales and opacities agree to two ulps (the device uses the library's pinned
    exp, numpy its own -- neither is Rust's libm, include/splat.h says so).  Uploading the activated
    arrays through splat_upload_soa renders the same frame, which equals the oracle's."""
    from oracle import oracle as orc
    from splat_b200 import _lib
    from splat_b200.camera import Camera
    from splat_b200.gaussians import load_ply_soa, save_ply

    orc.build()
    rng = np.random.default_rng(11)
    n = 20_000
    raw = {"x": rng.normal(2.0, 1.0, n), "y": rng.normal(-1.0, 0.7, n), "z": rng.normal(0.5, 1.2, n),
           "opacity": rng.normal(0.5, 2.0, n), "rot_0": rng.normal(size=n), "rot_1": rng.normal(size=n),
           "rot_2": rng.normal(size=n), "rot_3": rng.normal(size=n)}
    for i in range(3):
        raw[f"scale_{i}"] = rng.normal(-3.2, 0.7, n)
        raw[f"f_dc_{i}"] = rng.normal(0.0, 1.2, n)
    for i in range(45):
        raw[f"f_rest_{i}"] = rng.normal(0.0, 0.15, n)
    path = str(tmp_path / "scene.ply")
    save_ply(path, raw)
    host = load_ply_soa(path)

    ctx = _lib.Context(device=0)
    dev = ctx.upload_ply(path, want_activated=True)
    assert np.array_equal(dev.positions, host.positions)          # incl. the sequential-f32 mean
    assert np.array_equal(dev.rotations, host.rotations) and np.array_equal(dev.sh, host.sh)

    def ulps(a, b):
        return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max()

    assert ulps(dev.scales, host.scales) <= 2 and np.allclose(dev.opacities, host.opacities, rtol=1e-6, atol=0)
    W, H = 400, 300
    cam = Camera(H, W, (0.0, 0.0, 4.0))
    cam.update_camera_pose()
    cs = _lib.camera_struct(cam)
    a = np.zeros((H, W), np.uint32)
    ctx.render(cs, a)
    ctx2 = _lib.Context(device=0)
    ctx2.upload(dev)
    b = np.zeros((H, W), np.uint32)
    ctx2.render(cs, b)
    ref = np.zeros((H, W), np.uint32)
    orc.render(dev, orc.camera_from(cam), orc.make_config(), ref)
    assert np.count_nonzero(ref) > W * H // 10
    assert np.array_equal(a, b) and np.array_equal(a, ref)
    ctx.close()
    ctx2.close()
