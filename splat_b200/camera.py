"""Host-side mirror of the reference's orbit camera (reference: src/camera.rs).

Same field names, constants and quirks as `camera.rs:4-127`; all arithmetic in float32.
Only the two matrices, `position`, `w`, `h` and the focal length feed the CUDA path
(`splat_camera` in include/splat.h); everything here runs once per frame on the host.

The matrix builders restate nalgebra-glm 0.18.0 (`Cargo.lock:633`, not vendored):
`glm::look_at` = right-handed look-at, `glm::perspective(aspect, fovy, near, far)` =
`perspective_rh_no` (NDC z in [-1, 1]), `glm::rotation(angle, axis)` = axis-angle
(Rodrigues) as a homogeneous 4x4.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _normalize(v: np.ndarray) -> np.ndarray:
    v = v.astype(f32)
    n = f32(np.sqrt(f32(v[0] * v[0] + v[1] * v[1]) + f32(v[2] * v[2])))
    return (v / n).astype(f32)


def _cross(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = a.astype(f32)
    b = b.astype(f32)
    return np.array(
        [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=f32
    )


def look_at(eye: np.ndarray, center: np.ndarray, up: np.ndarray) -> np.ndarray:
    """`glm::look_at` (right-handed), used at camera.rs:65.  Returns a 4x4 float32 matrix
    indexed [row, col]."""
    eye = eye.astype(f32)
    z = _normalize(eye - center.astype(f32))  # camera looks down -z
    x = _normalize(_cross(up, z))
    y = _normalize(_cross(z, x))
    m = np.eye(4, dtype=f32)
    m[0, :3], m[1, :3], m[2, :3] = x, y, z
    neg = (-eye).astype(f32)
    for r, ax in enumerate((x, y, z)):
        m[r, 3] = f32(f32(ax[0] * neg[0] + ax[1] * neg[1]) + ax[2] * neg[2])
    return m


def perspective(aspect: float, fovy: float, near: float, far: float) -> np.ndarray:
    """`glm::perspective(aspect, fovy, near, far)` = perspective_rh_no, used at camera.rs:67."""
    aspect, fovy, near, far = f32(aspect), f32(fovy), f32(near), f32(far)
    t = f32(np.tan(f32(fovy / f32(2.0))))
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0] = f32(1.0) / f32(aspect * t)
    m[1, 1] = f32(1.0) / t
    m[2, 2] = -f32(far + near) / f32(far - near)
    m[2, 3] = -f32(f32(f32(2.0) * far) * near) / f32(far - near)
    m[3, 2] = f32(-1.0)
    return m


def rotation(angle: float, axis: np.ndarray) -> np.ndarray:
    """`glm::rotation(angle, axis)`: axis-angle rotation as a 4x4 (camera.rs:57, :62).
    A zero axis yields the identity (nalgebra's `Unit::try_new` fallback is not reachable
    from the viewer; NaNs are returned as-is)."""
    a = axis.astype(f32)
    n = f32(np.sqrt(f32(a[0] * a[0] + a[1] * a[1]) + f32(a[2] * a[2])))
    m = np.eye(4, dtype=f32)
    if n == 0:
        return m
    ux, uy, uz = (a / n).astype(f32)
    s, c = f32(np.sin(f32(angle))), f32(np.cos(f32(angle)))
    k = f32(1.0) - c
    m[:3, :3] = np.array(
        [
            [ux * ux * k + c, ux * uy * k - uz * s, ux * uz * k + uy * s],
            [ux * uy * k + uz * s, uy * uy * k + c, uy * uz * k - ux * s],
            [ux * uz * k - uy * s, uy * uz * k + ux * s, uz * uz * k + c],
        ],
        dtype=f32,
    )
    return m


class Camera:
    """Mirror of `struct Camera` (camera.rs:4-19) and its impl (:21-127)."""

    def __init__(self, h: float, w: float, start_position=None):
        # camera.rs:22-39
        self.znear = f32(0.01)
        self.zfar = f32(100.0)
        self.h = f32(h)
        self.w = f32(w)
        self.fovy = f32(np.pi / 2.0)
        pos = (0.0, 0.0, 3.0) if start_position is None else start_position
        self.position = np.array(pos, dtype=f32)
        self.target = np.zeros(3, dtype=f32)
        self.up = np.array([0.0, -1.0, 0.0], dtype=f32)
        self.yaw = f32(0.0)
        self.pitch = f32(0.0)
        self.is_pose_dirty = True
        self.is_intrin_dirty = True
        self.view_matrix = np.eye(4, dtype=f32)
        self.projection_matrix = np.eye(4, dtype=f32)

    def compute_matrices(self) -> None:
        # camera.rs:41-68
        position = np.append(self.position, f32(1.0)).astype(f32)
        pivot = np.append(self.target, f32(1.0)).astype(f32)
        viewdir = _normalize(self.position - self.target)
        cos_angle = f32(np.dot(viewdir, self.up))
        if cos_angle * np.sign(self.pitch) > f32(0.99):
            self.pitch = f32(0.0)
        rotation_x = rotation(self.yaw, self.up)
        position = (rotation_x @ (position - pivot)).astype(f32) + pivot
        # camera.rs:61: uses the *unrotated* self.position
        right = _cross(self.up, self.position)
        rotation_y = rotation(self.pitch, right)
        final_position = (rotation_y @ (position - pivot)).astype(f32) + pivot
        self.view_matrix = look_at(final_position[:3], self.target, self.up)
        self.projection_matrix = perspective(self.w / self.h, self.fovy, self.znear, self.zfar)

    def get_view_matrix(self) -> np.ndarray:
        return self.view_matrix

    def get_project_matrix(self) -> np.ndarray:
        return self.projection_matrix

    def update_resolution(self, height: float, width: float) -> None:
        self.h, self.w = f32(height), f32(width)
        self.is_intrin_dirty = True

    def get_htanfovxy_focal(self) -> np.ndarray:
        # camera.rs:84-89
        htany = f32(np.tan(f32(self.fovy / f32(2.0))))
        htanx = f32(f32(htany / self.h) * self.w)
        focal = f32(self.h / f32(f32(2.0) * htany))
        return np.array([htanx, htany, focal], dtype=f32)

    def get_focal(self) -> np.float32:
        return self.get_htanfovxy_focal()[2]

    def update_pitch_angle(self, delta: float) -> None:
        self.pitch = f32(self.pitch + f32(delta))
        self.is_pose_dirty = True

    def update_yaw_angle(self, delta: float) -> None:
        self.yaw = f32(self.yaw + f32(delta))
        self.is_pose_dirty = True

    def update_camera_pose(self) -> None:
        # camera.rs:103-126; note self.position is never updated
        self.compute_matrices()
        self.is_pose_dirty = False
