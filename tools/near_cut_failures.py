#!/usr/bin/env python
"""Diagnostics: which tiles fail the first (near) pass of a near-cut frame?  Renders the bench orbit with
SPLAT_NO_SECOND_PASS=1 (pixels are wrong then; only the per-tile arrays are looked at) and prints, per frame,
the failed tiles' near-list lengths and cut counts."""
import json
import os
import sys

import numpy as np

os.environ["SPLAT_NO_SECOND_PASS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from splat_b200 import _lib  # noqa: E402


def main():
    n, W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 6_100_000, 1920, 1080
    sc = bench.make_scene(n)
    ctx = _lib.Context(device=0, lowpass=0.3, near_cut=-1)
    ctx.upload(sc)
    cams = [_lib.camera_struct(bench._CamView(c)) for c in bench.orbit_cameras(W, H, 24)]
    fb = np.zeros((H, W), np.uint32)
    for i, cam in enumerate(cams):
        fb[:] = 0
        ctx.render(cam, fb)
        t = ctx.timings()
        if not t["near_cut_rank"]:
            continue
        rng = ctx.debug_tiles(0).astype(np.int64)
        near = rng[:, 1] - rng[:, 0]
        far = ctx.debug_tiles(1).astype(np.int64)
        failed = ctx.debug_tiles(2) != 0
        q = lambda a: [int(v) for v in np.percentile(a, [0, 25, 50, 75, 100])] if len(a) else []
        print(json.dumps({"frame": i, "failed_tiles": int(failed.sum()), "near_len_failed": q(near[failed]), "far_cnt_failed": q(far[failed]),
                          "tiles_far_gt0": int((far > 0).sum()), "near_len_all": q(near), "far_all": q(far),
                          "short_near_tiles(<128)&far>0": int(((near < 128) & (far > 0)).sum()),
                          "failed_with_near>=128": int((failed & (near >= 128)).sum()),
                          "failed_with_near>=256": int((failed & (near >= 256)).sum()),
                          "far_sum_short_near": int(far[(near < 128) & (far > 0)].sum())}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
