"""Host-side Camera mirror (splat_b200/camera.py) against closed forms of what camera.rs
computes with nalgebra-glm: look_at_rh, perspective_rh_no, axis-angle rotation, the focal
formula, and the two quirks the drop-in must keep."""
import numpy as np

from splat_b200.camera import Camera, look_at, perspective, rotation


def test_constants_match_camera_rs():
    cam = Camera(720, 1280)
    assert cam.znear == np.float32(0.01) and cam.zfar == np.float32(100.0)      # camera.rs:24-25
    assert cam.fovy == np.float32(np.pi / 2)                                     # :28
    assert list(cam.up) == [0.0, -1.0, 0.0] and list(cam.position) == [0.0, 0.0, 3.0]
    htanx, htany, focal = cam.get_htanfovxy_focal()                              # :84-89
    assert np.isclose(htany, 1.0) and np.isclose(htanx, 1280 / 720) and np.isclose(focal, 360.0)


def test_look_at_is_right_handed():
    v = look_at(np.array([0, 0, 5], np.float32), np.zeros(3, np.float32), np.array([0, -1, 0], np.float32))
    # origin lands at z = -5 in view space (visible points have negative z)
    assert np.allclose(v @ np.array([0, 0, 0, 1], np.float32), [0, 0, -5, 1])
    # rotation block orthonormal, up = (0,-1,0) flips x and y
    assert np.allclose(v[:3, :3] @ v[:3, :3].T, np.eye(3), atol=1e-6)
    assert np.allclose(v[:3, :3], np.diag([-1.0, -1.0, 1.0]))


def test_perspective_is_rh_negative_one_to_one():
    p = perspective(1280 / 720, np.pi / 2, 0.01, 100.0)
    assert np.isclose(p[0, 0], 720 / 1280) and np.isclose(p[1, 1], 1.0) and p[3, 2] == -1.0
    near = p @ np.array([0, 0, -0.01, 1], np.float32)
    far = p @ np.array([0, 0, -100.0, 1], np.float32)
    assert np.isclose(near[2] / near[3], -1.0, atol=1e-5) and np.isclose(far[2] / far[3], 1.0, atol=1e-5)


def test_rotation_axis_angle():
    r = rotation(np.pi / 2, np.array([0, 0, 2], np.float32))
    assert np.allclose(r @ np.array([1, 0, 0, 1], np.float32), [0, 1, 0, 1], atol=1e-6)
    assert np.array_equal(rotation(0.3, np.zeros(3, np.float32)), np.eye(4, dtype=np.float32))


def test_orbit_keeps_position_field_and_distance():
    """camera.rs:103-126 never writes self.position: the SH view direction keeps using the
    start position (pipelines.rs:99), while the view matrix orbits at constant distance."""
    cam = Camera(600, 800, (0.0, 0.0, 5.0))
    cam.update_camera_pose()
    v0 = cam.get_view_matrix().copy()
    for _ in range(9):
        cam.update_yaw_angle(10 * np.pi / 180)          # main.rs:53-60
    cam.update_camera_pose()
    assert list(cam.position) == [0.0, 0.0, 5.0] and not cam.is_pose_dirty
    v = cam.get_view_matrix()
    eye = -v[:3, :3].T @ v[:3, 3]
    assert np.isclose(np.linalg.norm(eye), 5.0, atol=1e-5)
    assert np.isclose(abs(eye[0]), 5.0, atol=1e-4)      # a quarter turn about the y axis
    assert not np.allclose(v, v0)
    for _ in range(27):
        cam.update_yaw_angle(10 * np.pi / 180)
    cam.update_camera_pose()
    assert np.allclose(cam.get_view_matrix(), v0, atol=2e-5)   # full turn


def test_pitch_uses_unrotated_right_vector():
    # camera.rs:61: right = cross(up, self.position) with the *unrotated* position
    cam = Camera(600, 800, (0.0, 0.0, 5.0))
    cam.update_yaw_angle(np.pi / 2)
    cam.update_pitch_angle(0.4)
    cam.update_camera_pose()
    v = cam.get_view_matrix()
    eye = -v[:3, :3].T @ v[:3, 3]
    # yaw moved the eye onto the x axis; pitching about right=(-5,0,0)||x leaves it there
    assert np.allclose(np.abs(eye), [5.0, 0.0, 0.0], atol=1e-4)


def test_matrices_are_float32_and_marshal_column_major():
    from splat_b200._lib import camera_struct

    cam = Camera(480, 640, (0.3, -0.2, 2.0))
    cam.update_camera_pose()
    s = camera_struct(cam)
    v = cam.get_view_matrix()
    assert v.dtype == np.float32
    assert np.array_equal(np.array(s.view[:]).reshape(4, 4).T.astype(np.float32), v)   # nalgebra storage
    assert s.w == 640.0 and s.h == 480.0 and np.isclose(s.focal, 240.0)


def test_camera_matrices_match_the_executed_prototype():
    """The view / projection matrices and the focal triple of the host mirror against the ones the
    reference's Python prototype produced (notes/util.py Camera with PyGLM semantics, executed by
    tools/make_golden_from_notebook.py): same look_at_rh, perspective_rh_no, constants."""
    import json
    import os

    from conftest import GOLDEN

    doc = json.load(open(os.path.join(GOLDEN, "notebook_projection.json")))
    assert len(doc["cases"]) >= 4
    for case in doc["cases"]:
        cam = Camera(case["h"], case["w"], tuple(case["cam_pos"]))
        cam.update_camera_pose()
        assert np.allclose(cam.get_view_matrix(), np.array(case["view"]), atol=2e-6), case["name"]
        assert np.allclose(cam.get_project_matrix(), np.array(case["proj"]), rtol=1e-6, atol=1e-6), case["name"]
        assert np.allclose(cam.get_htanfovxy_focal(), case["htanfovxy_focal"], rtol=1e-6), case["name"]
