"""ctypes binding of oracle/liboracle.so (the CPU restatement in splat_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the splat_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


class OrcCamera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("position", C.c_float * 3),
                ("w", C.c_float), ("h", C.c_float),
                ("htanx", C.c_float), ("htany", C.c_float), ("focal", C.c_float)]


class OrcConfig(C.Structure):
    _fields_ = [("lowpass", C.c_float), ("y_down", C.c_int), ("zclip_mode", C.c_int),
                ("sample_offset", C.c_float), ("exp_mode", C.c_int), ("nthreads", C.c_int)]


class OrcSplat(C.Structure):
    _fields_ = [("z_view", C.c_float), ("cov2d", C.c_float * 4), ("conic", C.c_float * 3),
                ("bbox", C.c_float * 2), ("ndc", C.c_float * 4), ("color", C.c_float * 3),
                ("opacity", C.c_float), ("cxp", C.c_float), ("cyp", C.c_float),
                ("conic_b_px", C.c_float), ("visible", C.c_int)]


class OrcStats(C.Structure):
    _fields_ = [("n_visible", C.c_uint64), ("pairs_in_rect", C.c_uint64),
                ("pairs_contributing", C.c_uint64)]


SPLAT_DTYPE = np.dtype([("z_view", "f4"), ("cov2d", "f4", 4), ("conic", "f4", 3), ("bbox", "f4", 2),
                        ("ndc", "f4", 4), ("color", "f4", 3), ("opacity", "f4"), ("cxp", "f4"),
                        ("cyp", "f4"), ("conic_b_px", "f4"), ("visible", "i4")])


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc; a few seconds)."""
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.orc_expf.restype = C.c_float
        L.orc_expf.argtypes = [C.c_float]
        L.orc_compute_cov3d_all.argtypes = [fp, fp, C.c_uint64, fp]
        L.orc_project.argtypes = [fp, fp, fp, fp, C.c_uint64, C.POINTER(OrcCamera),
                                  C.POINTER(OrcConfig), C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_sort_visible.restype = C.c_uint64
        L.orc_sort_visible.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_rasterize.restype = C.c_int
        L.orc_rasterize.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(OrcConfig),
                                    C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.POINTER(OrcStats)]
        L.orc_rasterize_rowlist.restype = C.c_int
        L.orc_rasterize_rowlist.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(OrcConfig),
                                            C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                                            C.POINTER(OrcStats)]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [fp, fp, fp, fp, C.c_uint64, C.POINTER(OrcCamera),
                                 C.POINTER(OrcConfig), C.c_void_p, C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_uint32, C.POINTER(OrcStats)]
        L.orc_shade_blend.restype = C.c_uint32
        L.orc_shade_blend.argtypes = [C.c_uint32] + [C.c_float] * 6 + [fp, C.c_int]
        L.orc_render_float.restype = C.c_int
        L.orc_render_float.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(OrcConfig), C.c_void_p,
                                       C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_sizeof_splat.restype = C.c_uint32
        assert L.orc_sizeof_splat() == SPLAT_DTYPE.itemsize == C.sizeof(OrcSplat)
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_camera(view, proj, position, w, h, htanx, htany, focal) -> OrcCamera:
    """view/proj: 4x4 arrays indexed [row, col]; stored column-major like nalgebra."""
    cam = OrcCamera()
    v = np.asarray(view, np.float32).T.reshape(-1)
    p = np.asarray(proj, np.float32).T.reshape(-1)
    for i in range(16):
        cam.view[i] = float(v[i])
        cam.proj[i] = float(p[i])
    for i in range(3):
        cam.position[i] = float(position[i])
    cam.w, cam.h = float(w), float(h)
    cam.htanx, cam.htany, cam.focal = float(htanx), float(htany), float(focal)
    return cam


def camera_from(camera) -> OrcCamera:
    """From a splat_b200.camera.Camera-shaped object (duck typed: no import of the product)."""
    hf = camera.get_htanfovxy_focal()
    return make_camera(camera.get_view_matrix(), camera.get_project_matrix(), camera.position,
                       camera.w, camera.h, hf[0], hf[1], hf[2])


def make_config(lowpass=0.3, y_down=0, zclip_mode=1, sample_offset=0.5, exp_mode=0,
                nthreads=None) -> OrcConfig:
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    return OrcConfig(lowpass, y_down, zclip_mode, sample_offset, exp_mode, nthreads)


def expf(x: float) -> float:
    return float(lib().orc_expf(float(x)))


def compute_cov3d(rotations: np.ndarray, scales: np.ndarray) -> np.ndarray:
    n = rotations.shape[0]
    out = np.zeros((n, 9), np.float32)
    lib().orc_compute_cov3d_all(_fp(np.ascontiguousarray(rotations, np.float32)),
                                _fp(np.ascontiguousarray(scales, np.float32)), n, _fp(out))
    return out


def project(scene, cam: OrcCamera, cfg: OrcConfig, W: int, H: int, cov3d=None) -> np.ndarray:
    """scene: GaussianList-shaped (positions (N,4), scales, opacities, rotations, sh (N,48))."""
    n = scene.positions.shape[0]
    if cov3d is None:
        cov3d = compute_cov3d(scene.rotations, scene.scales)
    out = np.zeros(n, SPLAT_DTYPE)
    lib().orc_project(_fp(scene.positions), _fp(cov3d), _fp(scene.opacities), _fp(scene.sh), n,
                      C.byref(cam), C.byref(cfg), W, H, out.ctypes.data)
    return out


def sort_visible(splats: np.ndarray) -> np.ndarray:
    order = np.zeros(max(len(splats), 1), np.uint32)
    m = lib().orc_sort_visible(splats.ctypes.data, len(splats), order.ctypes.data)
    return order[:m].copy()


def rasterize(splats, order, cfg, fb: np.ndarray, row0=0, row1=None):
    H, W = fb.shape
    assert fb.dtype == np.uint32 and fb.flags["C_CONTIGUOUS"]
    st = OrcStats()
    order = np.ascontiguousarray(order, np.uint32)
    rc = lib().orc_rasterize(splats.ctypes.data, order.ctypes.data, len(order), C.byref(cfg),
                             fb.ctypes.data, W, H, row0, H if row1 is None else row1, C.byref(st))
    if rc:
        raise RuntimeError(f"orc_rasterize failed: {rc}")
    return st


def rasterize_rows(splats, order, cfg, fb: np.ndarray, rows):
    """Blend onto an explicit ascending list of rows (bench.py's bounded CPU sample)."""
    H, W = fb.shape
    assert fb.dtype == np.uint32 and fb.flags["C_CONTIGUOUS"]
    st = OrcStats()
    order = np.ascontiguousarray(order, np.uint32)
    rows = np.ascontiguousarray(rows, np.int32)
    rc = lib().orc_rasterize_rowlist(splats.ctypes.data, order.ctypes.data, len(order), C.byref(cfg),
                                     fb.ctypes.data, W, H, rows.ctypes.data, len(rows), C.byref(st))
    if rc:
        raise RuntimeError(f"orc_rasterize_rowlist failed: {rc}")
    return st


def render(scene, cam: OrcCamera, cfg: OrcConfig, fb: np.ndarray, row0=0, row1=None, cov3d=None):
    """render_to_buffer restatement: blends the scene onto fb (uint32 (H, W)) in place."""
    H, W = fb.shape
    assert fb.dtype == np.uint32 and fb.flags["C_CONTIGUOUS"]
    n = scene.positions.shape[0]
    if cov3d is None:
        cov3d = compute_cov3d(scene.rotations, scene.scales)
    st = OrcStats()
    rc = lib().orc_render(_fp(scene.positions), _fp(cov3d), _fp(scene.opacities), _fp(scene.sh), n,
                          C.byref(cam), C.byref(cfg), fb.ctypes.data, W, H, row0,
                          H if row1 is None else row1, C.byref(st))
    if rc:
        raise RuntimeError(f"orc_render failed: {rc}")
    return st


def render_float(splats, order, cfg: OrcConfig, rgb: np.ndarray, acc_alpha: np.ndarray | None = None, mode: int = 0):
    """Un-quantised back-to-front "over" compositing onto rgb ((H, W, 3) f32) in place.
    mode 0 = the Rust fragment() semantics (pixel-centre sampling, 1/255 alpha cut, unclamped colour);
    mode 1 = the prototype's plot_opacity (notebook cell 3) restated line by line."""
    H, W, ch = rgb.shape
    assert ch == 3 and rgb.dtype == np.float32 and rgb.flags["C_CONTIGUOUS"]
    if acc_alpha is not None:
        assert acc_alpha.shape == (H, W) and acc_alpha.dtype == np.float32 and acc_alpha.flags["C_CONTIGUOUS"]
    order = np.ascontiguousarray(order, np.uint32)
    rc = lib().orc_render_float(splats.ctypes.data, order.ctypes.data, len(order), C.byref(cfg), rgb.ctypes.data,
                                acc_alpha.ctypes.data if acc_alpha is not None else None, W, H, int(mode))
    if rc:
        raise RuntimeError(f"orc_render_float failed: {rc}")


def shade_blend(old: int, A, B, Cc, dx, dy, opacity, rgb, exp_mode=0) -> int:
    arr = (C.c_float * 3)(*[float(v) for v in rgb])
    return int(lib().orc_shade_blend(int(old), float(A), float(B), float(Cc), float(dx), float(dy),
                                     float(opacity), arr, int(exp_mode)))
