"""The product's own CUDA sources on the CPU: splat_api.cu and the *.cuh kernels compiled for the host against
tests/cuda_emu (threads of a block = fibers, inline PTX mapped to functions of the same meaning, CUDA runtime =
synchronous host calls) and driven through the same C ABI, by the SAME test functions as the GPU parity suite.
What this checks without a GPU: the kernels' logic and the frame orchestration (sorts, binning, near cut and
open tiles, abandoned frames, stripes, the blend's producer/consumer rings and exact early termination) give
bit-identical pixels to the oracle.  What it cannot check: speed, and hardware-level races.  The emulated
library is test infrastructure; the product never loads it (tests/test_abi.py::test_no_cpu_fallback...)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))


@pytest.fixture(scope="module")
def lib():
    import emu_build
    from splat_b200 import _lib

    path = emu_build.build()
    saved = (_lib.LIB_PATH, _lib._lib)
    _lib.LIB_PATH, _lib._lib = path, None
    _lib.load()
    yield _lib
    _lib.LIB_PATH, _lib._lib = saved


@pytest.fixture(autouse=True)
def emulated_device_memory(monkeypatch):
    """tests that hand the library torch CUDA tensors: on the emulated device, device memory is host memory"""
    import contextlib

    import torch

    import _fake_gpu

    real_device = torch.device
    monkeypatch.setattr(torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(torch.cuda, "Stream", _fake_gpu.FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)


# the GPU suite's functions, collected here against the emulated library (sizes the emulator finishes in seconds)
from test_gpu_parity import (  # noqa: E402,F401
    test_blends_onto_existing_contents,
    test_degenerate_inputs_are_skipped,
    test_depth_order_matches_stable_sort,
    test_errors_not_crashes,
    test_euc_switches,
    test_frames_without_a_host_round_trip,
    test_near_cut_is_exact,
    test_near_cut_stripes_and_empty_regions,
    test_radix_sort_matches_stable_argsort,
    test_render_cleared_equals_fill_then_render,
    test_render_device_stripes_into_one_device_frame,
    test_skipped_frame_is_repeated_or_reported,
    test_stripes_equal_full_frame,
)
from test_gpu_parity import (  # noqa: E402,F401
    test_deep_lists_exact_early_termination,
    test_framebuffer_bit_exact,
    test_pipeline_mirrors,
    test_projection_records_match_oracle,
)
from test_gpu_float import (  # noqa: E402,F401
    test_float_and_reference_blends_differ_only_by_the_truncation_bias,
    test_float_blend_matches_float_oracle,
    test_float_mode_rejects_the_near_cut,
)
from test_ply import test_device_ply_ingest_matches_host_loader_and_renders_identically  # noqa: E402,F401


def test_stripe_frames_with_a_host_round_trip_in_every_frame(lib, orc):
    """What `bench.py --gpus N` does by default (sync_frames = 1, near cut off): each rank renders ITS stripe of
    every frame of an orbit, repeatedly with the same geometry, so a dense stripe switches from the stripe
    pre-pass to the full projection + pair compaction from its second frame on -- a combination the GPU runs of
    round 2 never exercised.  Every stripe of every frame must equal the oracle's rows."""
    import numpy as np

    from test_gpu_parity import _camera, _scene

    W, H = 320, 208
    scene = _scene(30_000, 0x5EED0091, -3.6)
    cfg = orc.make_config()
    bounds = [(0, 48), (48, 112), (112, 160), (160, 208)]             # the centre stripes keep > 30% of the Gaussians
    ctxs = [lib.Context(device=0, sync_frames=1, near_cut=0) for _ in bounds]
    for c in ctxs:
        c.upload(scene)
    kept = []
    for k, (pos, yaw) in enumerate([((0.0, 0.0, 5.0), 0.0), ((0.0, 0.0, 5.0), 0.2), ((0.0, 0.0, 2.5), 0.4), ((0.0, 0.0, 6.0), 0.6)]):
        cam = _camera(W, H, pos, yaw=yaw)
        want = np.zeros((H, W), np.uint32)
        orc.render(scene, orc.camera_from(cam), cfg, want)
        got = np.zeros((H, W), np.uint32)
        for c, (r0, r1) in zip(ctxs, bounds):
            part = np.ascontiguousarray(got[r0:r1])
            c.render(lib.camera_struct(cam), part, r0, r1)
            got[r0:r1] = part
            kept.append(c.timings()["n_visible"] / scene.num_gaussians)
        assert np.array_equal(got, want), (k, int(np.count_nonzero(got != want)))
    assert max(kept) > 0.3 and min(kept) < 0.3                        # both projection paths were taken
    for c in ctxs:
        c.close()
