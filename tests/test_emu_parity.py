"""The product's own CUDA sources on the CPU: splat_api.cu and the *.cuh kernels compiled for the host against
tests/cuda_emu (threads of a block = fibers, inline PTX mapped to functions of the same meaning, CUDA runtime =
synchronous host calls) and driven through the same C ABI, by the SAME test functions as the GPU parity suite.
What this checks without a GPU: the kernels' logic and the frame orchestration (sorts, binning, near cut and
open tiles, abandoned frames, stripes, the blend's producer/consumer rings and exact early termination) give
bit-identical pixels to the oracle.  What it cannot check: speed, and hardware-level races.  The runs that take
more than a few seconds each are skipped unless SPLAT_EMU_FULL=1 (tests/conftest.py: EMU_ON_REQUEST).  The emulated
library is test infrastructure; the product never loads it (tests/test_abi.py::test_no_cpu_fallback...)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))


@pytest.fixture(scope="module")
def lib():
    import emu_build
    from splat_b200 import _lib

    import subprocess

    try:
        path = emu_build.build()
    except (subprocess.CalledProcessError, OSError) as e:      # no C++20 host compiler here: an infrastructure gap, not a product failure
        pytest.skip(f"cannot build the host-emulated library: {e}")
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
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 3)       # the emulated box: ordinals only


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
    test_config2_plush_sized_complete_frame,
    test_deep_lists_exact_early_termination,
    test_full_size_configs_sampled_stripes,
    test_framebuffer_bit_exact,
    test_pipeline_mirrors,
    test_projection_records_match_oracle,
)
from test_zz_gpu_edge_cases import test_empty_scene_renders_nothing, test_everything_culled_leaves_the_buffer_untouched  # noqa: E402,F401
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


@pytest.mark.parametrize("which,cleared,n,W,H", [(2, 0, 3000, 320, 240), (2, 1, 3000, 320, 240), (1, 0, 1500, 200, 150)])
def test_cpp_viewer_loop_on_the_emulated_library(lib, cpp_demo, tmp_path, orc, which, cleared, n, W, H):
    """tests/test_zz_cpp_host_gpu.py without the GPU: examples/cpp_host/splat_demo (the C++ mirror of the reference's
    interface) bound at load time to the emulated build of the real library instead of the CUDA one"""
    import numpy as np

    from test_cpp_host import oracle_frame, raw_scene, read_frames, read_list, run

    from splat_b200.gaussians import save_ply

    libdir = tmp_path / "lib"
    libdir.mkdir()
    os.symlink(lib.LIB_PATH, libdir / "libsplat_b200.so")
    env = {"LD_LIBRARY_PATH": str(libdir)}
    frames, step = 3, 0.35
    ply, scene_dump, out = tmp_path / "s.ply", tmp_path / "s.bin", tmp_path / "frames.bin"
    save_ply(str(ply), raw_scene(n))
    run(cpp_demo, "ply", ply, scene_dump)
    scene = read_list(scene_dump)
    run(cpp_demo, "render", ply, H, W, 0.0, 0.0, 3.0, frames, step, which, cleared, out, env=env)
    got = read_frames(out, W, H)
    assert len(got) == frames
    for k, (cs, fb) in enumerate(got):
        ref = oracle_frame(orc, scene, cs, W, H, 0.01 if which == 1 else 0.3)
        assert np.count_nonzero(ref) > 1000
        assert np.array_equal(fb, ref), f"frame {k}: {np.count_nonzero(fb != ref)} of {W * H} pixels differ"


# ---------------------------------------------------------------------------------------------------------------
# bench.py on the real library (emulated): the harness with the product's own frame, count, timing and retry
# semantics -- tests/test_bench_flow.py runs the same flows on a stand-in that renders with the oracle
def test_bench_single_rank_on_the_emulated_library(lib, monkeypatch, orc, capfd):
    import importlib
    import json

    import _fake_gpu

    bench = importlib.import_module("bench")
    _fake_gpu.install_emulated(monkeypatch.setattr, lib.LIB_PATH, make_scene=bench.make_scene)
    _fake_gpu.quiet_bench(monkeypatch.setattr, bench)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gaussians", "1500", "--width", "160", "--height", "96", "--steps", "2", "--warmup", "3"])
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    assert bench.main() == 0
    out = capfd.readouterr().out.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["parity"]["mismatching_pixels"] == 0 and line["parity"]["rows"] == 96 and line["parity"]["covered_pixels"] > 1000
    assert line["gpu_launches"] >= 2 * 20 and line["instances_per_frame"] > 0
    assert line["stages_ms"]["blend_ms"] > 0 and line["e2e"]["value"] > 0 and line["e2e_cleared"]["value"] > 0
    assert line["frames_repeated"] == 0


def _emu_rank_main(rank, world, port, mode, emu_path, q):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      EMU_WORKERS="3")
    import _fake_gpu
    import bench

    log = _fake_gpu.install_emulated(setattr, emu_path, make_scene=bench.make_scene)
    lines = []
    _fake_gpu.quiet_bench(setattr, bench, sink=lines.append)
    sys.argv = ["bench.py", "--gpus", str(world), "--gaussians", "1500", "--width", "160", "--height", "128", "--steps", "2", "--warmup", "3",
                "--stripe-mode", mode, "--rebalance-rounds", "1"]
    rc = bench.main()
    ctx = log.contexts[0]
    q.put({"rank": rank, "rc": rc, "lines": lines, "gathers": log.gathers, "sync_frames": ctx.kw.get("sync_frames")})


@pytest.mark.parametrize("mode", ["sync", "async"])
def test_bench_two_ranks_on_the_emulated_library(lib, orc, mode):
    """two processes, each with the real library (emulated) rendering its stripe, stripes gathered over gloo: the
    frame that reaches rank 0's host buffer is the oracle's whole frame"""
    import torch.multiprocessing as mp

    from test_bench_flow import _check_line, _expected_checksum, _free_port

    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_emu_rank_main, args=(r, 2, port, mode, lib.LIB_PATH, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted((q.get(timeout=900) for _ in procs), key=lambda r: r["rank"])
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:            # never leave a rank behind, whatever happened
            if p.is_alive():
                p.kill()
    r0, r1 = res
    assert r0["rc"] == 0 and r1["rc"] == 0 and len(r0["lines"]) == 1 and r1["lines"] == []
    line = r0["lines"][0]
    _check_line(line, 2, 2, 3)
    assert r0["gathers"] == r1["gathers"] > 0
    assert r0["sync_frames"] == (1 if mode == "sync" else 0)
    assert line["frame_checksum"] == _expected_checksum(orc, n=1500, W=160, H=128)


def test_a_repeated_frame_is_counted_as_retried(lib):
    """found with the emulator: finish_frame's increment of frames_retried was wiped by the repeat's own reset"""
    import numpy as np

    from test_gpu_parity import _camera, _scene

    W, H = 320, 200
    scene = _scene(40_000, 0x5EED0072, -3.4)
    ctx = lib.Context(device=0, max_instances=1)
    ctx.upload(scene)
    for pos in ((0.0, 0.0, 40.0), (0.0, 0.0, 40.0), (0.0, 0.0, 1.5)):
        fb = np.zeros((H, W), np.uint32)
        ctx.render(lib.camera_struct(_camera(W, H, pos)), fb)
    t = ctx.timings()
    assert t["frames_skipped"] == 1 and t["frames_retried"] == 1
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------
# multi-device code of the library on the emulated box ("devices" are ordinals, NCCL is tests/cuda_emu/fake_nccl.cpp)
from test_gpu_multi import test_rank_contexts_gather_in_one_process  # noqa: E402,F401


@pytest.mark.parametrize("equal", [True, False], ids=["equal_stripes", "rebalanced"])
def test_group_context_equals_single_device_and_oracle_emulated(lib, orc, equal):
    """tests/test_gpu_multi.py::test_group_context_equals_single_device_and_oracle at a size the emulator finishes
    in seconds: splat_create_multi over three devices, scene broadcast, stripes rendered by the members, gathered,
    re-cut from the members' measured times -- every frame equals the single-device frame and the oracle."""
    import numpy as np

    from test_gpu_multi import _camera

    from splat_b200.gaussians import synthetic_scene

    G, W, H = 3, 256, 160
    scene = synthetic_scene(10_000, seed=0x5EED0081, log_scale_mean=-3.4)
    grp = lib.Context(devices=list(range(G)), equal_stripes=equal)
    grp.upload(scene)
    one = lib.Context(device=0)
    one.upload(scene)
    cfg = orc.make_config()
    rng = np.random.default_rng(81)
    seen_bounds = set()
    for k, yaw in enumerate(np.linspace(0.0, 1.2, 4)):
        cam = _camera(W, H, (0.0, 0.0, 4.0), yaw=float(yaw))
        fb0 = rng.integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) if k % 3 == 2 else np.zeros((H, W), np.uint32)
        want = fb0.copy()
        one.render(lib.camera_struct(cam), want)
        got = fb0.copy()
        grp.render(lib.camera_struct(cam), got)
        assert np.array_equal(got, want), (k, int(np.count_nonzero(got != want)))
        if k in (0, 2):
            ref = fb0.copy()
            orc.render(scene, orc.camera_from(cam), cfg, ref)
            assert np.array_equal(got, ref)
        got2 = np.full((H, W), 0xDEADBEEF, np.uint32)
        grp.render_cleared(lib.camera_struct(cam), got2, 0)
        want2 = np.zeros((H, W), np.uint32)
        one.render(lib.camera_struct(cam), want2)
        assert np.array_equal(got2, want2), k
        seen_bounds.add(tuple(grp.group_bounds()))
    b = grp.group_bounds()
    assert b[0][0] == 0 and b[-1][1] == H and all(b[i][1] == b[i + 1][0] for i in range(G - 1))
    if equal:
        trows = [(r1 + 15) // 16 - r0 // 16 for r0, r1 in b]
        assert max(trows) - min(trows) <= 1 and len(seen_bounds) == 1
    t = grp.timings()
    assert t["n_instances"] > 0 and t["n_gaussians"] == scene.num_gaussians
    grp.close()
    one.close()


def test_stripe_harness_repeats_an_abandoned_frame_locally(lib, orc):
    """stripes.render_with_retry / timings_with_retry (what bench.py's ranks use) against the REAL retry contract of
    splat_render_device: a stripe frame that outgrows its launch bounds is abandoned on the device, the next call on
    that context reports SPLAT_ERR_RETRY once, the helper renders the abandoned frame again and carries on -- and
    every frame a rank ends up with equals the oracle's rows."""
    import numpy as np

    from test_gpu_parity import _camera, _scene

    from splat_b200 import stripes

    W, H = 320, 208
    scene = _scene(40_000, 0x5EED0072, -3.4)
    cams = [_camera(W, H, (0.0, 0.0, z)) for z in (40.0, 40.0, 1.5, 1.6)]      # the third frame wants > 10x the instances
    cfg = orc.make_config()
    want = []
    for cam in cams:
        ref = np.zeros((H, W), np.uint32)
        orc.render(scene, orc.camera_from(cam), cfg, ref)
        want.append(ref)
    bounds = [(0, 96), (96, 208)]
    repeated = 0
    for r0, r1 in bounds:
        ctx = lib.Context(device=0, near_cut=0, max_instances=1)
        ctx.upload(scene)
        fb = np.zeros((H, W), np.uint32)

        def render(i):
            fb[r0:r1] = 0
            ctx.render_device(lib.camera_struct(cams[i]), fb[r0:r1].ctypes.data, W, H, r0, r1)

        def is_retry(e):
            return isinstance(e, lib.SplatError) and e.code == stripes.RETRY

        for i in range(len(cams)):
            repeated += stripes.render_with_retry(render, i, is_retry)
            if i == len(cams) - 1 or i == 1:
                _, rep = stripes.timings_with_retry(ctx.timings, render, i, is_retry)     # waits for frame i (repeats it if abandoned)
                repeated += rep
                assert np.array_equal(fb[r0:r1], want[i][r0:r1]), (r0, i, int(np.count_nonzero(fb[r0:r1] != want[i][r0:r1])))
        assert ctx.timings()["frames_skipped"] >= 1
        ctx.close()
    assert repeated >= 2                                                         # each stripe had its near frame abandoned once


def test_cpp_host_renders_an_empty_ply_on_the_emulated_library(lib, cpp_demo, tmp_path):
    import numpy as np

    from test_cpp_host import read_frames, run

    from splat_b200.gaussians import save_ply

    libdir = tmp_path / "lib"
    libdir.mkdir()
    os.symlink(lib.LIB_PATH, libdir / "libsplat_b200.so")
    ply, out = tmp_path / "e.ply", tmp_path / "f.bin"
    save_ply(str(ply), {"x": np.zeros(0, np.float32)})
    for which in (1, 2):
        run(cpp_demo, "render", ply, 48, 64, 0.0, 0.0, 3.0, 2, 0.2, which, 0, out, env={"LD_LIBRARY_PATH": str(libdir)})
        assert all(not fb.any() for _, fb in read_frames(out, 64, 48))


def test_group_context_repeats_abandoned_member_stripes(lib, orc):
    """a camera that jumps from far to near: the members' stripes outgrow their launch bounds, are abandoned on the
    "device" and repeated inside splat_render of the group -- frames blended onto noise and the fused clear stay exact"""
    import numpy as np

    from test_gpu_multi import _camera

    from splat_b200.gaussians import synthetic_scene

    W, H = 320, 208
    scene = synthetic_scene(40_000, seed=0x5EED0072, log_scale_mean=-3.4)
    cfg = orc.make_config()
    grp = lib.Context(devices=[0, 1, 2], max_instances=1)
    grp.upload(scene)
    skipped = 0
    for k, z in enumerate([40.0, 40.0, 1.5, 40.0, 1.4]):
        cam = _camera(W, H, (0.0, 0.0, z))
        fb0 = np.random.default_rng(k).integers(0, 2 ** 32, size=(H, W), dtype=np.uint64).astype(np.uint32) if k % 2 else np.zeros((H, W), np.uint32)
        got, want = fb0.copy(), fb0.copy()
        grp.render(lib.camera_struct(cam), got)
        orc.render(scene, orc.camera_from(cam), cfg, want)
        assert np.array_equal(got, want), (k, int(np.count_nonzero(got != want)))
        skipped = grp.timings()["frames_skipped"]
        got2, want2 = np.full((H, W), 7, np.uint32), np.zeros((H, W), np.uint32)
        grp.render_cleared(lib.camera_struct(cam), got2, 0)
        orc.render(scene, orc.camera_from(cam), cfg, want2)
        assert np.array_equal(got2, want2), k
    assert skipped >= 1
    grp.close()


def test_headless_demo_writes_the_frames_the_oracle_renders(lib, orc, tmp_path):
    """python -m splat_b200.demo (the reference's viewer loop with PNG files instead of a window), on the emulated library:
    the PNGs decode to the oracle's RGB bytes, for a PLY (device ingest) and for the 4-Gaussian scene"""
    import struct
    import zlib

    import numpy as np

    from test_cpp_host import raw_scene

    from splat_b200 import demo
    from splat_b200.camera import Camera
    from splat_b200.gaussians import GaussianList, naive_gaussians, save_ply

    def read_png(path):
        data = open(path, "rb").read()
        assert data[:8] == b"\x89PNG\r\n\x1a\n"
        pos, idat, W, H = 8, b"", 0, 0
        while pos < len(data):
            n, tag = struct.unpack(">I4s", data[pos:pos + 8])
            body = data[pos + 8:pos + 8 + n]
            assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xFFFFFFFF
            if tag == b"IHDR":
                W, H, depth, ctype = struct.unpack(">IIBB", body[:10])
                assert (depth, ctype) == (8, 2)
            elif tag == b"IDAT":
                idat += body
            pos += 12 + n
        rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(H, 1 + 3 * W)
        assert not rows[:, 0].any()
        return rows[:, 1:].reshape(H, W, 3)

    W, H = 160, 96
    ply = tmp_path / "s.ply"
    save_ply(str(ply), raw_scene(1200))
    out = tmp_path / "frames"
    assert demo.main([str(ply), "--width", str(W), "--height", str(H), "--camera", "0", "0", "3", "--frames", "2", "--yaw-step", "0.3", "--out", str(out)]) == 0
    ctx = lib.Context(device=0)
    scene = ctx.upload_ply(str(ply), want_activated=True)          # the floats the device activated
    ctx.close()
    cam = Camera(H, W, (0.0, 0.0, 3.0))
    for i in range(2):
        cam.update_yaw_angle(0.3 if i else 0.0)
        cam.update_camera_pose()
        ref = np.zeros((H, W), np.uint32)
        orc.render(scene, orc.camera_from(cam), orc.make_config(), ref)
        rgb = read_png(out / f"frame_{i:04d}.png")
        want = np.stack([(ref >> 16) & 0xFF, (ref >> 8) & 0xFF, ref & 0xFF], axis=-1).astype(np.uint8)
        assert np.count_nonzero(ref) > 500 and np.array_equal(rgb, want)
    out2 = tmp_path / "naive"
    assert demo.main(["naive", "--width", "128", "--height", "80", "--pipeline", "1", "--out", str(out2)]) == 0
    cam = Camera(80, 128, (0.0, 0.0, 3.0))
    cam.update_camera_pose()
    ref = np.zeros((80, 128), np.uint32)
    orc.render(GaussianList.from_vec(naive_gaussians()), orc.camera_from(cam), orc.make_config(lowpass=0.01), ref)
    rgb = read_png(out2 / "frame_0000.png")
    assert np.array_equal(rgb, np.stack([(ref >> 16) & 0xFF, (ref >> 8) & 0xFF, ref & 0xFF], axis=-1).astype(np.uint8))
