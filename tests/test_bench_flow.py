"""bench.py's control flow and JSON contract, exercised on the CPU: the CUDA library and torch.cuda are
replaced by stand-ins (tests/_fake_gpu.py -- the stand-in "GPU" renders with the oracle, which tests may
call), so that a slip in the harness cannot first show up on the GPU box.  What is checked is the harness:
the single JSON line and its keys, the parity leg, and above all that a frame abandoned on one rank is
repeated with local work only, every rank issuing exactly one gather per frame (round 2's 8-GPU run hung
on precisely that).  No number in the line means anything here."""
import importlib
import json
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                 "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "frames_repeated", "stage_rooflines")


def _expected_checksum(orc, n=1500, W=160, H=96, steps=2, warmup=3):
    """the last e2e frame of a bench run (camera 2*(warmup+steps)-1 of the orbit), rendered by the oracle"""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    from splat_b200 import _lib

    cams = bench.orbit_cameras(W, H, 2 * (warmup + steps))
    cs = _lib.camera_struct(bench._CamView(cams[-1]))
    view = np.array(cs.view, np.float32).reshape(4, 4).T
    proj = np.array(cs.proj, np.float32).reshape(4, 4).T
    cam = orc.make_camera(view, proj, list(cs.position), cs.w, cs.h, cs.htanx, cs.htany, cs.focal)
    fb = np.zeros((H, W), np.uint32)
    orc.render(bench.make_scene(n), cam, orc.make_config(lowpass=0.3, nthreads=2), fb)
    assert np.count_nonzero(fb) > 0
    return int(fb.astype(np.uint64).sum())


def _check_line(line, n_gpus, steps, warmup):
    for key in CONTRACT_KEYS:
        assert key in line, key
    assert line["n_gpus"] == n_gpus and line["steps"] == steps and line["warmup"] == warmup and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] > 0
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_single_gpu_flow_emits_one_valid_json_line(monkeypatch, orc, capfd):
    import _fake_gpu

    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    log = _fake_gpu.install(monkeypatch.setattr, orc, abandon_at=(3,))
    _fake_gpu.quiet_bench(monkeypatch.setattr, bench)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gaussians", "1500", "--width", "160", "--height", "96", "--steps", "2", "--warmup", "3"])
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    assert bench.main() == 0
    out = capfd.readouterr().out.strip().splitlines()
    assert len(out) == 1                                      # exactly one JSON line on stdout
    line = json.loads(out[0])
    _check_line(line, 1, 2, 3)
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert line["parity"]["mismatching_pixels"] == 0 and line["parity"]["rows"] == 96      # the stand-in GPU is the oracle
    assert "pixels_differing" in line["exp_variant"]
    assert line["frames_repeated"] == 1 and log.retries_reported == 1    # the abandoned frame was repeated, once
    assert log.gathers == 0
    assert line["frame_checksum"] == _expected_checksum(orc)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    import _fake_gpu
    import bench
    from oracle import oracle as orc

    orc.build()
    orc.lib()
    # async stripes: rank 1 has two frames abandoned (one found by the next render, one by the timings read of the
    # very frame); rank 0 none.  sync stripes (the default): a frame is never abandoned.
    abandon = ((5, 16) if rank == 1 else ()) if mode == "async" else ()
    log = _fake_gpu.install(setattr, orc, abandon_at=abandon, make_scene=bench.make_scene)
    lines = []
    _fake_gpu.quiet_bench(setattr, bench, sink=lines.append)
    sys.argv = ["bench.py", "--gpus", str(world), "--gaussians", "1500", "--width", "160", "--height", "96", "--steps", "2", "--warmup", "3",
                "--stripe-mode", mode]
    rc = bench.main()
    ctx = log.contexts[0]
    q.put({"rank": rank, "rc": rc, "lines": lines, "gathers": log.gathers, "renders": log.renders, "retries": log.retries_reported,
           "sync_frames": ctx.kw.get("sync_frames"), "near_cut": ctx.kw.get("near_cut")})


@pytest.mark.parametrize("mode", ["sync", "async"])
def test_two_rank_flow_gathers_once_per_frame_whatever_happens_to_a_rank(mode, orc):
    import torch.multiprocessing as mp

    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_rank_main, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r["rank"])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:            # never leave a rank behind, whatever happened
            if p.is_alive():
                p.kill()
    r0, r1 = res
    assert r0["rc"] == 0 and r1["rc"] == 0
    assert len(r0["lines"]) == 1 and r1["lines"] == []         # rank 0 alone prints, one line
    line = r0["lines"][0]
    _check_line(line, 2, 2, 3)
    assert "cpu_baseline" not in line                          # N > 1: no CPU leg
    assert r0["gathers"] == r1["gathers"] > 0                  # the same number of collectives on every rank
    if mode == "sync":
        assert r0["sync_frames"] == 1 and r0["near_cut"] == 0 and r1["retries"] == 0
    else:
        assert r0["sync_frames"] == 0 and r1["retries"] == 2 and r1["renders"] > r0["renders"]
    assert line["frame_checksum"] == _expected_checksum(orc)   # the gathered frame in rank 0's host buffer is the whole frame


def test_reference_arm_times_the_cpu_restatement_and_other_ranks_do_no_work(monkeypatch, orc, capfd):
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "capture_stdout", lambda: None)
    argv = ["bench.py", "--impl", "reference", "--gaussians", "1500", "--width", "160", "--height", "96", "--steps", "2", "--warmup", "3"]
    monkeypatch.setattr(sys, "argv", argv)
    monkeypatch.setenv("RANK", "1")
    assert bench.main() == 0
    assert capfd.readouterr().out.strip() == ""                # ranks other than 0 print nothing
    monkeypatch.setenv("RANK", "0")
    assert bench.main() == 0
    out = capfd.readouterr().out.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["steps"] == 2 and line["warmup"] == 3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # both arms describe the workload with the same words (the driver compares them)
    ours = importlib.import_module("bench").workload_config(type("A", (), {"width": 160, "height": 96, "n": 1500})())
    assert line["config"] == ours
