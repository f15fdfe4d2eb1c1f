#!/usr/bin/env python
"""BASELINE.json configs 2-5 on one GPU: frames/s, per-stage times and achieved GB/s per stage,
plus size-independent parity properties at full size (the oracle is too slow there):

  * a frame assembled from tile-row stripes is byte-identical to the full-frame render
    (per-pixel lists do not depend on the partition; exercises the stripe cull + compaction path),
  * rendering onto a noise framebuffer: pixels that keep the noise are exactly pixels that stay 0
    when rendering onto zeros (no quad covers them); the share of pixels whose value does not
    depend on the input at all (exact early termination converged) is reported,
  * rendering twice gives identical bytes.

  python tools/configs_report.py [--quick] > gpurun_out/configs.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from splat_b200 import _lib  # noqa: E402
from splat_b200.camera import Camera  # noqa: E402
from splat_b200.gaussians import synthetic_scene  # noqa: E402

DEMO_CAM = (-0.57651054, 2.99040512, -0.03924271)   # 02_ply_demo.rs:22
NEAR_CUT = int(os.environ.get("SPLAT_NEAR_CUT", "0"))   # splat_config.near_cut for every context (0 = off)


def run(name, n, seed, W, H, campos, frames, props):
    import torch

    scene = synthetic_scene(n, seed=seed)
    ctx = _lib.Context(device=0, lowpass=0.3, near_cut=NEAR_CUT)
    ctx.upload(scene)
    cam = Camera(H, W, campos)
    cams = []
    for _ in range(frames + 3):
        cam.update_yaw_angle(bench.YAW_STEP)
        cam.update_camera_pose()
        cams.append(_lib.camera_struct(cam))
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    fb = torch.zeros((H, W), dtype=torch.int32, device=dev)
    for i in range(3):
        fb.zero_()
        ctx.render_device(cams[i], fb.data_ptr(), W, H, 0, H, stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(3, 3 + frames):
        fb.zero_()
        ctx.render_device(cams[i], fb.data_ptr(), W, H, 0, H, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    t = ctx.timings()
    full = fb.cpu().numpy().view(np.uint32).copy()
    T = ((W + 15) // 16) * ((H + 15) // 16)
    I = t["n_instances"]
    passes = (max(1, int(np.ceil(np.log2(max(T, 2))))) + 7) // 8
    peak, _ = bench.measured_peak()
    gb = lambda b, m: b / (m * 1e-3) / 1e9 if m > 0 else 0.0
    out = {"config": name, "n_gaussians": n, "width": W, "height": H, "fps": 1000.0 / ms, "ms_per_frame": ms,
           "n_instances": int(I), "n_visible": int(t["n_visible"]),
           "stages_ms": {k: t[k] for k in ("project_ms", "sort_ms", "bin_ms", "blend_ms", "total_ms")},
           "project_GBps": gb(228.0 * n, t["project_ms"]), "sort_GBps": gb(20.0 * (4 * n + passes * I), t["sort_ms"]),
           "blend_equiv_GBps": gb(52.0 * I + 8.0 * T + 8.0 * W * H, t["blend_ms"]), "hbm_peak_GBps": peak,
           "checksum": int(full.astype(np.uint64).sum())}
    if props:
        last = cams[2 + frames]
        # (1) stripes == full frame
        fb.zero_()
        trows = (H + 15) // 16
        cuts = [0, (trows // 3) * 16, (2 * trows // 3) * 16, H]
        for r0, r1 in zip(cuts[:-1], cuts[1:]):
            ctx.render_device(last, fb[r0:r1].data_ptr(), W, H, r0, r1, stream.cuda_stream)
        torch.cuda.synchronize()
        out["stripes_equal_full"] = bool(np.array_equal(fb.cpu().numpy().view(np.uint32), full))
        # (2) twice == once
        fb.zero_()
        ctx.render_device(last, fb.data_ptr(), W, H, 0, H, stream.cuda_stream)
        torch.cuda.synchronize()
        out["repeatable"] = bool(np.array_equal(fb.cpu().numpy().view(np.uint32), full))
        # (3) onto noise: pixels never covered keep the noise; the rest mostly equals the render onto zeros
        noise = np.random.default_rng(1).integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32)
        fb.copy_(torch.from_numpy(noise.view(np.int32)).to(dev))
        ctx.render_device(last, fb.data_ptr(), W, H, 0, H, stream.cuda_stream)
        torch.cuda.synchronize()
        on_noise = fb.cpu().numpy().view(np.uint32)
        kept = on_noise == noise          # pixels no quad covers keep the input; rendered onto zeros they are 0
        out["pixels_keeping_input"] = int(np.count_nonzero(kept))
        out["kept_pixels_are_uncovered"] = bool(np.all(full[kept] == 0))
        out["pixels_independent_of_input_pct"] = float(100.0 * np.count_nonzero(on_noise == full) / full.size)
    ctx.close()
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="skip the 4K config and the upper half of the sweep")
    a = ap.parse_args()
    run("C2 plush-sized synthetic, demo camera", 281_498, 0x5EED0002, 1280, 720, DEMO_CAM, 20, True)
    run("C3 bicycle-sized synthetic", 6_100_000, 0x5EED0003, 1920, 1080, (0.0, 0.0, 5.0), 20, True)
    if not a.quick:
        run("C4 garden-sized synthetic, 4K, one GPU", 5_800_000, 0x5EED0004, 3840, 2160, (0.0, 0.0, 5.0), 10, True)
    for n in ([10_000, 100_000, 1_000_000] if a.quick else [10_000, 30_000, 100_000, 300_000, 1_000_000, 3_000_000, 10_000_000]):
        run(f"C5 sweep N={n}", n, 0x5EED0005, 1920, 1080, (0.0, 0.0, 5.0), 20, False)


if __name__ == "__main__":
    main()
