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
