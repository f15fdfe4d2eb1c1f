#!/usr/bin/env python
"""Work counters of the blend kernel on the bench workload (instrumented library variant).

  SPLAT_B200_LIB=$PWD/splat_b200/libsplat_b200_stats.so python tools/blend_stats.py [--n N]

Build the variant with:  nvcc <flags of splat_b200/csrc/Makefile> -DSPLAT_STATS -shared \
                         -o splat_b200/libsplat_b200_stats.so splat_b200/csrc/splat_api.cu
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from splat_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=6_100_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--near-cut", type=int, default=0)
    a = ap.parse_args()
    sc = bench.make_scene(a.n)
    ctx = _lib.Context(device=0, lowpass=bench.LOWPASS, near_cut=a.near_cut)
    ctx.upload(sc)
    cams = bench.orbit_cameras(a.width, a.height, a.frames)
    fb = np.zeros((a.height, a.width), np.uint32)
    names = ["group_entries_evaluated", "group_entries_to_consumer", "pixel_pairs_alpha_gt0",
             "list_entries_staged", "candidate_pairs", "lanes_alpha_gt0", "suffix_attempts", "suffix_attempts_failed"]
    ctx.debug_blend_stats(reset=True)
    for i, c in enumerate(cams):
        fb[:] = 0
        ctx.render(_lib.camera_struct(bench._CamView(c)), fb)
        t = ctx.timings()
        st = ctx.debug_blend_stats(reset=True)
        d = {k: int(v) for k, v in zip(names, st)}
        d.update(frame=i, n_instances=int(t["n_instances"]), n_visible=int(t["n_visible"]), blend_ms=t["blend_ms"],
                 near_cut_rank=int(t["near_cut_rank"]), near_cut_failed=int(t["near_cut_failed"]), frames_retried=int(t["frames_retried"]), total_ms=t["total_ms"])
        d["lane_fill_consumer"] = d["lanes_alpha_gt0"] / max(1, 32 * d["group_entries_to_consumer"])
        d["lane_fill_producer"] = d["lanes_alpha_gt0"] / max(1, 32 * d["group_entries_evaluated"])
        print(json.dumps(d), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
