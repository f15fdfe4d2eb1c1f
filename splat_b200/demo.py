"""Headless counterpart of the reference's viewers (src/main.rs, src/bin/01_naive_gaussian.rs, src/bin/02_ply_demo.rs):
the same loop -- orbit the camera, clear, render_to_buffer -- with the frames written as PNG files instead of shown
in a minifb window (the window itself is out of scope).

    python -m splat_b200.demo point_cloud.ply --frames 36 --out frames/          # 02_ply_demo.rs: 1280x720, its camera
    python -m splat_b200.demo naive --width 640 --height 480 --frames 4           # 01_naive_gaussian.rs's four Gaussians

The scene is activated and recentred on the device when the file has the INRIA layout (splat_upload_ply_raw),
otherwise by the host loader.  Needs libsplat_b200.so and a B200; there is no CPU path."""
from __future__ import annotations

import argparse
import os
import struct
import sys
import zlib

import numpy as np


def write_png(path: str, argb: np.ndarray) -> None:
    """(H, W) uint32 0xAARRGGBB (euc::Buffer<u32, 2>, main.rs:79) -> an 8-bit RGB PNG (alpha dropped, as on the minifb window)"""
    H, W = argb.shape
    rgb = np.empty((H, W, 3), np.uint8)
    rgb[..., 0], rgb[..., 1], rgb[..., 2] = (argb >> 16) & 0xFF, (argb >> 8) & 0xFF, argb & 0xFF
    raw = np.concatenate([np.zeros((H, 1), np.uint8), rgb.reshape(H, W * 3)], axis=1).tobytes()     # filter type 0 per row

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m splat_b200.demo", description=__doc__.split("\n\n")[0])
    ap.add_argument("scene", help="a 3DGS .ply file, or `naive` for the reference's 4-Gaussian test scene")
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--camera", type=float, nargs=3, default=None, help="start position (default: the demo's, 02_ply_demo.rs:22 / (0,0,3) for naive)")
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--yaw-step", type=float, default=0.1, help="radians per frame (the arrow keys of the viewer, main.rs:50-66)")
    ap.add_argument("--pipeline", type=int, choices=(1, 2), default=2, help="GaussianSplatPipeline01 (low-pass 0.01) or 02 (0.3)")
    ap.add_argument("--devices", type=int, nargs="+", default=[0], help="several ordinals = screen-tile stripes across those GPUs")
    ap.add_argument("--out", default="frames")
    args = ap.parse_args(argv)

    from . import _lib
    from .camera import Camera
    from .gaussians import GaussianList, load_ply_soa, naive_gaussians, ply_vertex_payload

    W, H = args.width, args.height
    naive = args.scene == "naive"
    pos = tuple(args.camera) if args.camera else ((0.0, 0.0, 3.0) if naive else (-0.57651054, 2.99040512, -0.03924271))
    camera = Camera(H, W, pos)
    ctx = _lib.Context(device=args.devices[0], lowpass=0.01 if args.pipeline == 1 else 0.3,
                       devices=args.devices if len(args.devices) > 1 else None)
    if naive:
        ctx.upload(GaussianList.from_vec(naive_gaussians()))
    else:
        try:
            ply_vertex_payload(args.scene)                     # INRIA layout: activate + recentre on the device
            ctx.upload_ply(args.scene)
        except ValueError:
            ctx.upload(load_ply_soa(args.scene))               # any other property order / ascii: the host loader
    os.makedirs(args.out, exist_ok=True)
    color = np.zeros((H, W), np.uint32)
    for i in range(args.frames):
        camera.update_yaw_angle(args.yaw_step if i else 0.0)
        camera.update_camera_pose()                            # main.rs:70
        ctx.render_cleared(_lib.camera_struct(camera), color, 0)    # main.rs:73-74: fill(0) + render_to_buffer
        write_png(os.path.join(args.out, f"frame_{i:04d}.png"), color)
    t = ctx.timings()
    print(f"{args.frames} frame(s) of {t['n_gaussians']} Gaussians at {W}x{H} -> {args.out}/  (last frame: {t['total_ms']:.3f} ms on the device)",
          file=sys.stderr)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
