#!/usr/bin/env python
"""One GPU rendering ONE stripe of the bench frame the way rank k of G would (for ncu / stage timing without
a multi-GPU box):  python tools/stripe_profile.py G k [frames]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from splat_b200 import _lib, stripes  # noqa: E402


def main():
    import torch

    G, k = int(sys.argv[1]), int(sys.argv[2])
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    n, W, H = 6_100_000, 1920, 1080
    sc = bench.make_scene(n)
    ctx = _lib.Context(device=0, lowpass=0.3)
    ctx.upload(sc)
    cams = [_lib.camera_struct(bench._CamView(c)) for c in bench.orbit_cameras(W, H, frames + 4)]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    fb = torch.zeros((H, W), dtype=torch.int32, device=dev)
    # bounds like the bench: probe frame + measured cost model would need all ranks; equal tile rows are enough here
    r0, r1 = stripes.stripe_bounds(H, G)[k]
    acc = {}
    for i, cam in enumerate(cams):
        fb[r0:r1].zero_()
        ctx.render_device(cam, fb[r0:r1].data_ptr(), W, H, r0, r1, stream.cuda_stream)
        t = ctx.timings()
        if i >= 4:
            for key in ("project_ms", "sort_ms", "bin_ms", "blend_ms", "second_pass_ms", "total_ms", "n_visible", "n_instances"):
                acc[key] = acc.get(key, 0.0) + t[key] / frames
    print(json.dumps({"G": G, "rank": k, "rows": [r0, r1], **{a: round(b, 4) for a, b in acc.items()}}))
    ctx.close()


if __name__ == "__main__":
    main()
