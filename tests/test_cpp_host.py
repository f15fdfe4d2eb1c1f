"""The C++ host side above the C ABI (include/splat_pipeline.hpp: Camera, Gaussian, GaussianList,
load_from_ply, GaussianSplatPipeline01/02::render_to_buffer -- the reference's interface for this path in
compiled code) through its demo driver examples/cpp_host/splat_demo.

CPU: the camera maths and the marshalled splat_camera against the Python mirror; load_from_ply against the
Python loader; the whole render loop against the oracle, with a stand-in library (tests/fake_splat) that
implements the C-ABI calls on top of the oracle -- this checks every byte the C++ side hands across the
boundary; and that with the REAL library and no GPU the program fails loudly (no CPU path).
GPU (tests/test_zz_cpp_host_gpu.py): the same loop on the real library, bit-exact against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from splat_b200 import _lib
from splat_b200.camera import Camera
from splat_b200.gaussians import GaussianList, load_ply_soa, naive_gaussians, save_ply

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAM_BYTES = ctypes.sizeof(_lib.SplatCamera)


@pytest.fixture(scope="module")
def demo(cpp_demo):
    return cpp_demo


def run(demo, *args, env=None, check=True):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([demo, *[str(a) for a in args]], capture_output=True, text=True, env=e, timeout=300)
    if check:
        assert p.returncode == 0, p.stderr
    return p


def cam_floats(cs) -> np.ndarray:
    return np.frombuffer(bytes(cs), np.float32).copy()


def read_list(path) -> GaussianList:
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw, np.uint64, 1)[0])
    a = np.frombuffer(raw, np.float32, offset=8)
    assert a.size == 60 * n
    o = np.cumsum([0, 4 * n, 3 * n, n, 4 * n, 48 * n])
    return GaussianList(a[o[0]:o[1]].reshape(n, 4), a[o[1]:o[2]].reshape(n, 3), a[o[2]:o[3]], a[o[3]:o[4]].reshape(n, 4),
                        a[o[4]:o[5]].reshape(n, 48))


def raw_scene(n, seed=11):
    """pre-activation PLY properties of a small blob in front of the camera"""
    rng = np.random.default_rng(seed)
    raw = {"x": rng.normal(0.3, 0.5, n), "y": rng.normal(-0.2, 0.5, n), "z": rng.normal(0.1, 0.5, n),
           "opacity": rng.normal(0.5, 2.0, n)}
    for i in range(3):
        raw[f"scale_{i}"] = rng.normal(-3.0, 0.6, n)
        raw[f"f_dc_{i}"] = rng.normal(0.0, 1.2, n)
    for i in range(4):
        raw[f"rot_{i}"] = rng.normal(0.0, 1.0, n)
    for i in range(24):
        raw[f"f_rest_{i}"] = rng.normal(0.0, 0.15, n)
    return {k: v.astype(np.float32) for k, v in raw.items()}


def read_frames(path, W, H):
    raw = open(path, "rb").read()
    per = CAM_BYTES + 4 * W * H
    assert len(raw) % per == 0 and len(raw) > 0
    out = []
    for k in range(len(raw) // per):
        cs = _lib.SplatCamera.from_buffer_copy(raw[k * per:k * per + CAM_BYTES])
        fb = np.frombuffer(raw, np.uint32, W * H, offset=k * per + CAM_BYTES).reshape(H, W).copy()
        out.append((cs, fb))
    return out


def oracle_frame(orc, scene: GaussianList, cs, W, H, lowpass):
    view = np.array(cs.view, np.float32).reshape(4, 4).T
    proj = np.array(cs.proj, np.float32).reshape(4, 4).T
    cam = orc.make_camera(view, proj, list(cs.position), cs.w, cs.h, cs.htanx, cs.htany, cs.focal)
    fb = np.zeros((H, W), np.uint32)
    orc.render(scene, cam, orc.make_config(lowpass=lowpass, nthreads=2), fb)
    return fb


@pytest.mark.parametrize("pose", [(96, 160, 0.0, 0.0, 5.0, 0.0, 0.0), (240, 320, 0.5, -0.3, 4.0, 0.7, 0.2), (1080, 1920, 0.0, 0.0, 3.0, -2.4, -0.4)])
def test_camera_matches_the_python_mirror(demo, tmp_path, pose):
    """camera.rs:41-89 in C++ against splat_b200/camera.py (itself checked against the prototype's matrices in
    tests/test_camera.py): same struct, field by field; sin/cos/tan come from different libms, hence 2e-6"""
    H, W, x, y, z, yaw, pitch = pose
    out = tmp_path / "cam.bin"
    run(demo, "camera", H, W, x, y, z, yaw, pitch, out)
    got = np.frombuffer(out.read_bytes(), np.float32)
    cam = Camera(H, W, (x, y, z))
    cam.update_yaw_angle(yaw)
    cam.update_pitch_angle(pitch)
    cam.update_camera_pose()
    want = cam_floats(_lib.camera_struct(cam))
    assert got.size == want.size == CAM_BYTES // 4
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    assert got[35] == np.float32(W) and got[36] == np.float32(H)            # w, h
    np.testing.assert_array_equal(got[32:35], np.float32([x, y, z]))        # position: the FIELD, never the orbited eye


def test_naive_scene_and_ply_loader_match_the_python_mirror(demo, tmp_path):
    out = tmp_path / "naive.bin"
    run(demo, "naive", out)
    got, want = read_list(out), GaussianList.from_vec(naive_gaussians())
    for name in ("positions", "scales", "opacities", "rotations", "sh"):
        np.testing.assert_array_equal(getattr(got, name), getattr(want, name), err_msg=name)

    ply = tmp_path / "scene.ply"
    save_ply(str(ply), raw_scene(500))
    run(demo, "ply", ply, out)
    got, want = read_list(out), load_ply_soa(str(ply))
    assert got.num_gaussians == 500
    np.testing.assert_array_equal(got.positions, want.positions)           # sequential f32 mean: exact
    np.testing.assert_array_equal(got.rotations, want.rotations)           # rot_0 -> w
    np.testing.assert_array_equal(got.sh, want.sh)                         # f_rest_i -> sh[3+i]
    np.testing.assert_allclose(got.scales, want.scales, rtol=3e-7)         # exp from two libms
    np.testing.assert_allclose(got.opacities, want.opacities, rtol=1e-6)


def test_ply_with_another_element_is_the_reference_panic(demo, tmp_path):
    ply = tmp_path / "bad.ply"
    ply.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 0\nproperty float x\nelement face 0\nend_header\n")
    p = run(demo, "ply", ply, tmp_path / "o.bin", check=False)
    assert p.returncode == 3 and "Unexpected element!" in p.stderr
    p = run(demo, "ply", tmp_path / "missing.ply", tmp_path / "o.bin", check=False)
    assert p.returncode == 3 and "cannot open" in p.stderr


def test_trim_writes_the_same_file_as_the_python_mirror(demo, tmp_path):
    from splat_b200.gaussians import trim_ply

    src, a, b = tmp_path / "s.ply", tmp_path / "a.ply", tmp_path / "b.ply"
    save_ply(str(src), raw_scene(40))
    for count in (3, 40, 100):
        p = run(demo, "trim", src, a, count)
        assert int(p.stdout) == trim_ply(str(src), str(b), count) == min(count, 40)
        assert load_ply_soa(str(a)).num_gaussians == min(count, 40)
        assert a.read_bytes().split(b"end_header\n", 1)[1] == b.read_bytes().split(b"end_header\n", 1)[1]
        assert a.read_bytes().split(b"end_header\n", 1)[0].split() == b.read_bytes().split(b"end_header\n", 1)[0].split()


@pytest.fixture(scope="module")
def fake_lib_dir(tmp_path_factory, orc):
    d = tmp_path_factory.mktemp("fake_splat")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", str(d / "libsplat_b200.so"),
                           os.path.join(ROOT, "tests", "fake_splat", "fake_splat.c"), "-L", odir, "-loracle", f"-Wl,-rpath,{odir}"])
    return str(d)


@pytest.mark.parametrize("which,cleared", [(2, 0), (2, 1), (1, 0)])
def test_render_loop_hands_the_library_exactly_the_right_bytes(demo, tmp_path, orc, fake_lib_dir, which, cleared):
    """main.rs's loop (yaw, update_camera_pose, fill(0), render_to_buffer) in C++ over a stand-in library that renders
    with the oracle: every frame equals the oracle's frame for the camera the program reports, the reported cameras
    follow the orbit, and the scene went across once."""
    W, H, frames, step = 160, 96, 3, 0.35
    ply, scene_dump, out, log = tmp_path / "s.ply", tmp_path / "s.bin", tmp_path / "frames.bin", tmp_path / "fake.log"
    save_ply(str(ply), raw_scene(400))
    run(demo, "ply", ply, scene_dump)
    scene = read_list(scene_dump)
    env = {"LD_LIBRARY_PATH": fake_lib_dir + ":" + os.path.join(ROOT, "oracle"), "FAKE_SPLAT_LOG": str(log)}
    run(demo, "render", ply, H, W, 0.0, 0.0, 3.0, frames, step, which, cleared, out, env=env)
    got = read_frames(out, W, H)
    assert len(got) == frames
    lowpass = 0.01 if which == 1 else 0.3
    cam = Camera(H, W, (0.0, 0.0, 3.0))
    seen = set()
    for cs, fb in got:
        cam.update_yaw_angle(step)
        cam.update_camera_pose()
        np.testing.assert_allclose(cam_floats(cs), cam_floats(_lib.camera_struct(cam)), rtol=2e-6, atol=2e-6)
        ref = oracle_frame(orc, scene, cs, W, H, lowpass)
        assert np.count_nonzero(ref) > 500
        assert np.array_equal(fb, ref), f"{np.count_nonzero(fb != ref)} pixels differ"
        seen.add(fb.tobytes())
    assert len(seen) == frames                                              # the camera really moved
    assert log.read_text().strip() == f"uploads=1 renders={frames} lowpass={lowpass:.2f} pinned=1 still=0"   # the colour buffer: pinned once, unpinned before destroy


def test_a_ply_without_vertices_is_an_empty_scene(demo, tmp_path, fake_lib_dir):
    """load_from_ply on `element vertex 0` gives an empty Vec (no division by zero in the recentring), and
    render_to_buffer over it leaves the cleared buffer cleared"""
    W, H = 64, 48
    ply, dump, out = tmp_path / "e.ply", tmp_path / "e.bin", tmp_path / "f.bin"
    save_ply(str(ply), {"x": np.zeros(0, np.float32)})
    run(demo, "ply", ply, dump)
    assert read_list(dump).num_gaussians == 0
    env = {"LD_LIBRARY_PATH": fake_lib_dir + ":" + os.path.join(ROOT, "oracle")}
    run(demo, "render", ply, H, W, 0.0, 0.0, 3.0, 2, 0.2, 2, 0, out, env=env)
    for _, fb in read_frames(out, W, H):
        assert not fb.any()


def test_listing_devices_asks_for_a_group_context(demo, tmp_path, fake_lib_dir):
    log = tmp_path / "fake.log"
    env = {"LD_LIBRARY_PATH": fake_lib_dir + ":" + os.path.join(ROOT, "oracle"), "FAKE_SPLAT_LOG": str(log), "SPLAT_DEMO_DEVICES": "0,2,3"}
    run(demo, "render", "naive", 96, 160, 0, 0, 5, 2, 0.1, 2, 0, tmp_path / "o.bin", env=env)
    assert log.read_text().splitlines() == ["group of 3: 0 2 3", "uploads=1 renders=2 lowpass=0.30 pinned=1 still=0"]


def test_library_errors_surface_as_exceptions_with_the_library_message(demo, tmp_path, fake_lib_dir):
    env = {"LD_LIBRARY_PATH": fake_lib_dir + ":" + os.path.join(ROOT, "oracle"), "FAKE_SPLAT_FAIL_CREATE": "1"}
    p = run(demo, "render", "naive", 96, 160, 0, 0, 5, 1, 0.1, 2, 0, tmp_path / "o.bin", env=env, check=False)
    assert p.returncode == 3 and "splat_create: no device (fake)" in p.stderr


def test_without_a_gpu_the_real_library_fails_loudly(demo, tmp_path):
    """no CPU path behind the C++ host side either"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = tmp_path / "o.bin"
    p = run(demo, "render", "naive", 96, 160, 0, 0, 5, 1, 0.1, 2, 0, out, check=False)
    assert p.returncode == 3 and "splat_create" in p.stderr
    assert not out.exists() or out.stat().st_size == 0
