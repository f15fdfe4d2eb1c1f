"""ctypes binding of libsplat_b200.so -- the same C ABI (include/splat.h) the Rust shim binds
with bindgen.  There is no fallback: if the CUDA library is missing or fails to load, importing
a render entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPLAT_B200_LIB overrides the path (kernel experiments build variant libraries side by side)
LIB_PATH = os.environ.get("SPLAT_B200_LIB") or os.path.join(_HERE, "libsplat_b200.so")

SPLAT_OK = 0
ERRORS = {-1: "SPLAT_ERR_INVALID", -2: "SPLAT_ERR_CUDA", -3: "SPLAT_ERR_NOMEM",
          -4: "SPLAT_ERR_UNSUPPORTED", -5: "SPLAT_ERR_STATE", -6: "SPLAT_ERR_RETRY"}
SPLAT_BLEND_REFERENCE, SPLAT_BLEND_FLOAT = 0, 1

# every symbol include/splat.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = ["splat_abi_version", "splat_config_default", "splat_create", "splat_create_error", "splat_destroy",
           "splat_last_error", "splat_upload_soa", "splat_upload_aos", "splat_upload_ply_raw", "splat_render", "splat_render_cleared",
           "splat_render_rows", "splat_render_device", "splat_get_timings", "splat_get_tile_loads", "splat_pin_host",
           "splat_unpin_host", "splat_debug_project", "splat_debug_read_order",
           "splat_debug_sort_pairs", "splat_debug_blend_stats", "splat_debug_render_float", "splat_debug_read_tiles",
           "splat_debug_partition",
           "splat_create_multi", "splat_group_get_bounds", "splat_comm_unique_id", "splat_comm_init_rank",
           "splat_comm_broadcast_scene", "splat_gather_stripes"]


class SplatConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("lowpass", C.c_float), ("y_down", C.c_int32),
                ("zclip_mode", C.c_int32), ("sample_offset", C.c_float), ("tile", C.c_uint32),
                ("max_instances", C.c_uint64), ("blend_mode", C.c_int32), ("near_cut", C.c_int32),
                ("sync_frames", C.c_int32), ("reserved", C.c_int32)]


class SplatCamera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("position", C.c_float * 3),
                ("w", C.c_float), ("h", C.c_float),
                ("htanx", C.c_float), ("htany", C.c_float), ("focal", C.c_float)]


class SplatTimings(C.Structure):
    _fields_ = [("project_ms", C.c_float), ("sort_ms", C.c_float), ("bin_ms", C.c_float),
                ("blend_ms", C.c_float), ("total_ms", C.c_float), ("h2d_ms", C.c_float),
                ("d2h_ms", C.c_float), ("frames_retried", C.c_uint32),
                ("n_gaussians", C.c_uint64), ("n_visible", C.c_uint64), ("n_instances", C.c_uint64),
                ("n_tiles", C.c_uint64), ("kernel_launches", C.c_uint64), ("near_cut_rank", C.c_uint64), ("near_cut_failed", C.c_uint64), ("near_cut_instances", C.c_uint64),
                ("frames_skipped", C.c_uint64), ("second_pass_instances", C.c_uint64), ("near_cut_fallbacks", C.c_uint64),
                ("second_pass_ms", C.c_float), ("reserved_", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SplatError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    fp, u32p, vp = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_void_p
    L.splat_abi_version.restype = C.c_uint32
    L.splat_config_default.argtypes = [C.POINTER(SplatConfig)]
    L.splat_config_default.restype = None
    L.splat_create.argtypes = [C.POINTER(vp), C.POINTER(SplatConfig)]
    L.splat_create_error.restype = C.c_char_p
    L.splat_destroy.argtypes = [vp]
    L.splat_destroy.restype = None
    L.splat_last_error.argtypes = [vp]
    L.splat_last_error.restype = C.c_char_p
    L.splat_upload_soa.argtypes = [vp, fp, fp, fp, fp, fp, C.c_uint64]
    L.splat_upload_aos.argtypes = [vp, fp, C.c_uint64]
    L.splat_upload_ply_raw.argtypes = [vp, vp, C.c_uint64, C.c_uint32, vp]
    L.splat_render.argtypes = [vp, C.POINTER(SplatCamera), vp, C.c_uint32, C.c_uint32]
    L.splat_render_cleared.argtypes = [vp, C.POINTER(SplatCamera), vp, C.c_uint32, C.c_uint32, C.c_uint32]
    L.splat_render_rows.argtypes = [vp, C.POINTER(SplatCamera), vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.splat_render_device.argtypes = [vp, C.POINTER(SplatCamera), vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.splat_get_timings.argtypes = [vp, C.POINTER(SplatTimings)]
    L.splat_get_tile_loads.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.splat_pin_host.argtypes = [vp, C.c_uint64]
    L.splat_unpin_host.argtypes = [vp]
    L.splat_debug_project.argtypes = [vp, C.POINTER(SplatCamera), C.c_uint32, C.c_uint32, vp, vp, vp]
    L.splat_debug_read_order.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.splat_debug_sort_pairs.argtypes = [vp, vp, vp, C.c_uint64, C.c_int]
    L.splat_debug_blend_stats.argtypes = [vp, vp, C.c_int]
    L.splat_debug_render_float.argtypes = [vp, C.POINTER(SplatCamera), vp, C.c_uint32, C.c_uint32, vp]
    L.splat_debug_read_tiles.argtypes = [vp, C.c_int, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.splat_debug_partition.argtypes = [vp, vp, C.c_int32, C.c_uint32]
    L.splat_create_multi.argtypes = [C.POINTER(vp), C.POINTER(SplatConfig), C.POINTER(C.c_int32), C.c_int32]
    L.splat_group_get_bounds.argtypes = [vp, vp, C.c_int32, C.POINTER(C.c_int32)]
    L.splat_comm_unique_id.argtypes = [vp]
    L.splat_comm_init_rank.argtypes = [vp, vp, C.c_int32, C.c_int32]
    L.splat_comm_broadcast_scene.argtypes = [vp, C.c_int32, C.c_uint64]
    L.splat_gather_stripes.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_int32, vp]
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if the .so does not export it
    _lib = L
    return L


def group_partition(H: int, parts: int, bounds=None, times_ms=None):
    """The library's own stripe partition rules, on the host (splat_debug_partition): the initial equal cut,
    or -- given the current bounds and the members' measured frame times -- the re-cut one."""
    L = load()
    flat = (C.c_uint32 * (2 * parts))()
    ms = None
    if times_ms is not None:
        flat[:] = [int(v) for b in bounds for v in b]
        ms = (C.c_float * parts)(*[float(t) for t in times_ms])
    rc = L.splat_debug_partition(flat, ms, parts, H)
    if rc:
        raise SplatError(rc, "splat_debug_partition")
    return [(int(flat[2 * k]), int(flat[2 * k + 1])) for k in range(parts)]


def camera_struct(camera) -> SplatCamera:
    """Marshal a splat_b200.camera.Camera exactly like the Rust shim does."""
    cam = SplatCamera()
    v = np.asarray(camera.get_view_matrix(), np.float32).T.reshape(-1)   # column-major
    p = np.asarray(camera.get_project_matrix(), np.float32).T.reshape(-1)
    hf = camera.get_htanfovxy_focal()
    for i in range(16):
        cam.view[i] = float(v[i])
        cam.proj[i] = float(p[i])
    for i in range(3):
        cam.position[i] = float(camera.position[i])
    cam.w, cam.h = float(camera.w), float(camera.h)
    cam.htanx, cam.htany, cam.focal = float(hf[0]), float(hf[1]), float(hf[2])
    return cam


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Context:
    """Owns one splat_ctx (one GPU)."""

    def __init__(self, device=0, lowpass=0.3, y_down=0, zclip_mode=1, sample_offset=0.5, max_instances=0, near_cut=-1,
                 blend_mode=SPLAT_BLEND_REFERENCE, sync_frames=0, devices=None, equal_stripes=False):
        """devices: a list of CUDA ordinals makes this a GROUP context (splat_create_multi): one process,
        several GPUs, stripes cut and gathered inside the library."""
        self.L = load()
        cfg = SplatConfig()
        self.L.splat_config_default(C.byref(cfg))
        cfg.device, cfg.lowpass, cfg.y_down = device, lowpass, y_down
        cfg.zclip_mode, cfg.sample_offset, cfg.max_instances = zclip_mode, sample_offset, max_instances
        cfg.near_cut, cfg.blend_mode, cfg.sync_frames = (near_cut if blend_mode == SPLAT_BLEND_REFERENCE or near_cut > 0 else 0), blend_mode, sync_frames
        cfg.reserved = 1 if equal_stripes else 0
        self.cfg = cfg
        self.h = C.c_void_p()
        self.devices = list(devices) if devices is not None else None
        if self.devices is None:
            rc = self.L.splat_create(C.byref(self.h), C.byref(cfg))
        else:
            arr = (C.c_int32 * len(self.devices))(*self.devices)
            rc = self.L.splat_create_multi(C.byref(self.h), C.byref(cfg), arr, len(self.devices))
        if rc:
            raise SplatError(rc, "splat_create failed: " + self.L.splat_create_error().decode())
        self.n = 0

    # ---- one process per GPU: communicator owned by the context (include/splat.h, multi-GPU (2))
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.L.splat_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int):
        self._check(self.L.splat_comm_init_rank(self.h, C.create_string_buffer(unique_id, 128), n_ranks, rank))

    def broadcast_scene(self, root: int, n: int):
        self._check(self.L.splat_comm_broadcast_scene(self.h, root, n))
        self.n = n

    def gather_stripes(self, fb_dev_ptr: int, W: int, H: int, bounds, root=0, stream=0):
        flat = (C.c_uint32 * (2 * len(bounds)))(*[int(v) for b in bounds for v in b])
        self._check(self.L.splat_gather_stripes(self.h, fb_dev_ptr, W, H, flat, root, stream or None))

    def group_bounds(self):
        n = C.c_int32()
        self._check(self.L.splat_group_get_bounds(self.h, None, 0, C.byref(n)))
        arr = (C.c_uint32 * (2 * n.value))()
        self._check(self.L.splat_group_get_bounds(self.h, arr, n.value, C.byref(n)))
        return [(int(arr[2 * k]), int(arr[2 * k + 1])) for k in range(n.value)]

    def _check(self, rc):
        if rc:
            raise SplatError(rc, self.L.splat_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.splat_destroy(self.h)
            self.h = None

    __del__ = close

    def upload(self, g):
        """g: GaussianList-shaped (positions, scales, opacities, rotations, sh)."""
        self._check(self.L.splat_upload_soa(self.h, _fp(g.positions), _fp(g.scales), _fp(g.opacities),
                                            _fp(g.rotations), _fp(g.sh), g.positions.shape[0]))
        self.n = g.positions.shape[0]

    def upload_ply(self, filename: str, want_activated: bool = False):
        """load_from_ply on the device: the vertex payload of an INRIA-layout binary PLY goes to the GPU as
        it lies in the file (memory-mapped); activation and recentring happen there.  Returns the activated
        GaussianList when asked (tests), else None."""
        from .gaussians import GaussianList, ply_vertex_payload

        rows, n = ply_vertex_payload(filename)          # (n, 62) float32 view of the file
        act = np.zeros(60 * n, np.float32) if want_activated else None
        self._check(self.L.splat_upload_ply_raw(self.h, rows.ctypes.data, n, rows.shape[1],
                                                act.ctypes.data if act is not None else None))
        self.n = n
        if act is None:
            return None
        pos, rot, sc, op, sh = np.split(act, [4 * n, 8 * n, 11 * n, 12 * n])
        return GaussianList(pos.reshape(n, 4), sc.reshape(n, 3), op, rot.reshape(n, 4), sh.reshape(n, 48))

    def upload_aos(self, g59: np.ndarray):
        self._check(self.L.splat_upload_aos(self.h, _fp(g59), g59.shape[0]))
        self.n = g59.shape[0]

    def render(self, cam: SplatCamera, fb: np.ndarray, row0=0, row1=None):
        """fb: uint32 (rows, W) host array holding rows [row0,row1) of the H x W image."""
        assert fb.dtype == np.uint32 and fb.flags["C_CONTIGUOUS"]
        H, W = int(cam.h), int(cam.w)
        row1 = H if row1 is None else row1
        assert fb.shape == (row1 - row0, W)
        self._check(self.L.splat_render_rows(self.h, C.byref(cam), fb.ctypes.data, W, H, row0, row1))

    def render_cleared(self, cam: SplatCamera, fb: np.ndarray, clear: int = 0):
        """clear + render_to_buffer in one call; fb (H, W) uint32 is only written."""
        assert fb.dtype == np.uint32 and fb.flags["C_CONTIGUOUS"]
        H, W = int(cam.h), int(cam.w)
        assert fb.shape == (H, W)
        self._check(self.L.splat_render_cleared(self.h, C.byref(cam), fb.ctypes.data, W, H, clear))

    def render_ptr(self, cam: SplatCamera, host_ptr: int, W: int, H: int, row0=0, row1=None):
        self._check(self.L.splat_render_rows(self.h, C.byref(cam), host_ptr, W, H, row0, H if row1 is None else row1))

    def render_device(self, cam: SplatCamera, dev_ptr: int, W: int, H: int, row0=0, row1=None, stream=0):
        self._check(self.L.splat_render_device(self.h, C.byref(cam), dev_ptr, W, H, row0,
                                               H if row1 is None else row1, stream or None))

    def render_float(self, cam: SplatCamera, fb: np.ndarray) -> np.ndarray:
        """SPLAT_BLEND_FLOAT contexts: render onto fb in place and return the un-quantised (H, W, 4) f32
        r, g, b, 1-T of every touched pixel (NaN elsewhere)."""
        assert fb.dtype == np.uint32 and fb.flags["C_CONTIGUOUS"]
        H, W = int(cam.h), int(cam.w)
        assert fb.shape == (H, W)
        rgba = np.zeros((H, W, 4), np.float32)
        self._check(self.L.splat_debug_render_float(self.h, C.byref(cam), fb.ctypes.data, W, H, rgba.ctypes.data))
        return rgba

    def timings(self) -> dict:
        t = SplatTimings()
        self._check(self.L.splat_get_timings(self.h, C.byref(t)))
        return t.as_dict()

    def tile_loads(self, tiles_x: int) -> np.ndarray:
        """(tile_rows, tiles_x) instance counts of the last render's stripe."""
        nt = C.c_uint64()
        t = self.timings()
        out = np.zeros(max(int(t["n_tiles"]), 1), np.uint32)
        self._check(self.L.splat_get_tile_loads(self.h, out.ctypes.data, len(out), C.byref(nt)))
        return out[: nt.value].reshape(-1, tiles_x)

    def debug_project(self, cam: SplatCamera):
        H, W = int(cam.h), int(cam.w)
        rec = np.zeros((self.n, 12), np.float32)
        keys = np.zeros(self.n, np.uint32)
        rects = np.zeros((self.n, 4), np.uint16)
        self._check(self.L.splat_debug_project(self.h, C.byref(cam), W, H, rec.ctypes.data,
                                               keys.ctypes.data, rects.ctypes.data))
        return rec, keys, rects

    def debug_order(self) -> np.ndarray:
        order = np.zeros(max(self.n, 1), np.uint32)
        nv = C.c_uint64()
        self._check(self.L.splat_debug_read_order(self.h, order.ctypes.data, len(order), C.byref(nv)))
        return order[: nv.value].copy()

    def debug_tiles(self, which: int) -> np.ndarray:
        """per-tile arrays of the last frame (0: ranges (T, 2), 1: far_cnt, 2: tile_failed)"""
        n = C.c_uint64()
        t = int(self.timings()["n_tiles"])
        out = np.zeros(2 * t, np.uint32)
        self._check(self.L.splat_debug_read_tiles(self.h, which, out.ctypes.data, len(out), C.byref(n)))
        out = out[: n.value]
        return out.reshape(-1, 2) if which == 0 else out

    def debug_blend_stats(self, reset=True) -> np.ndarray:
        out = np.zeros(8, np.uint64)
        self._check(self.L.splat_debug_blend_stats(self.h, out.ctypes.data, int(reset)))
        return out

    def debug_sort_pairs(self, keys: np.ndarray, vals: np.ndarray, bits=32):
        assert keys.dtype == np.uint32 and vals.dtype == np.uint32
        self._check(self.L.splat_debug_sort_pairs(self.h, keys.ctypes.data, vals.ctypes.data, len(keys), bits))
