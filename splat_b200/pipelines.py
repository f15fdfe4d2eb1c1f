"""Host-side mirror of the reference's render entry points (reference: src/pipelines.rs).

`GaussianSplatPipeline01` (pipelines.rs:54-57, AoS `Vec<Gaussian>`, low-pass 0.01) and
`GaussianSplatPipeline02` (:172-175, SoA `GaussianList`, low-pass 0.3) keep the reference's
public fields (`gaussians`, `camera`) and its one method, `render_to_buffer(color)`, where
`color` plays the role of `euc::Buffer<u32, 2>`: a (H, W) uint32 array, row-major, 0xAARRGGBB,
blended onto and overwritten.  The body is what the Rust shim's body becomes
(INTEGRATION.md): marshal -> splat_upload_* once -> splat_render.  All rendering maths runs in
libsplat_b200.so; there is no CPU path here.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import _lib
from .camera import Camera
from .gaussians import Gaussian, GaussianList


class _PipelineBase:
    LOWPASS = 0.3

    def __init__(self, gaussians, camera: Camera, device: int = 0, y_down: int = 0,
                 zclip_mode: int = 1, sample_offset: float = 0.5):
        self.gaussians = gaussians
        self.camera = camera
        self._ctx = _lib.Context(device=device, lowpass=self.LOWPASS, y_down=y_down,
                                 zclip_mode=zclip_mode, sample_offset=sample_offset)
        self._uploaded = None

    def _upload(self):
        raise NotImplementedError

    def render_to_buffer(self, color: np.ndarray) -> None:
        """pipelines.rs:66-86 / :260-280.  `color` is modified in place."""
        if color.dtype != np.uint32 or color.ndim != 2 or not color.flags["C_CONTIGUOUS"]:
            raise TypeError("color must be a C-contiguous (H, W) uint32 array")
        if self._uploaded is not self.gaussians:
            self._upload()
            self._uploaded = self.gaussians
        H, W = color.shape
        cam = _lib.camera_struct(self.camera)
        self._ctx._check(self._ctx.L.splat_render(self._ctx.h, cam, color.ctypes.data, W, H))

    def render_cleared_to_buffer(self, color: np.ndarray, clear: int = 0) -> None:
        """`color.fill(clear); render_to_buffer(color)` (main.rs:73-74) without the host fill and
        its upload: `color` is only written."""
        if color.dtype != np.uint32 or color.ndim != 2 or not color.flags["C_CONTIGUOUS"]:
            raise TypeError("color must be a C-contiguous (H, W) uint32 array")
        if self._uploaded is not self.gaussians:
            self._upload()
            self._uploaded = self.gaussians
        H, W = color.shape
        cam = _lib.camera_struct(self.camera)
        self._ctx._check(self._ctx.L.splat_render_cleared(self._ctx.h, cam, color.ctypes.data, W, H, clear))

    def timings(self) -> dict:
        return self._ctx.timings()


class GaussianSplatPipeline01(_PipelineBase):
    """pipelines.rs:54-169: `gaussians: Vec<Gaussian>`, low-pass +0.01 (gaussians.rs:156-157)."""

    LOWPASS = 0.01

    def _upload(self):
        gs: List[Gaussian] = self.gaussians
        g59 = np.zeros((len(gs), 59), np.float32)
        for i, g in enumerate(gs):
            g59[i, 0:3] = g.position
            g59[i, 3:6] = g.scale
            g59[i, 6] = g.opacity
            g59[i, 7:11] = g.rotation
            g59[i, 11:59] = g.sh
        self._ctx.upload_aos(g59)


class GaussianSplatPipeline02(_PipelineBase):
    """pipelines.rs:172-281: `gaussians: GaussianList`, low-pass +0.3 (gaussians.rs:517-518)."""

    LOWPASS = 0.3

    def _upload(self):
        self._ctx.upload(self.gaussians)
