"""Multi-GPU host logic without a GPU: stripe partition + grouped send/recv gather at world
size 2 over gloo.  Each rank "renders" its stripe with the CPU oracle (the checker standing in
for the device here) into its rows of a full frame; after the gather rank 0 must hold exactly
the single-process full-frame render."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from splat_b200 import stripes  # noqa: E402


def test_equal_bounds_cover_and_align():
    for H in (1080, 130, 16, 17, 2160):
        for world in (1, 2, 3, 4, 8):
            b = stripes.stripe_bounds(H, world)
            assert len(b) == world
            stripes.check_bounds(b, H)
    # more ranks than tile rows: the surplus ranks get empty stripes
    b = stripes.stripe_bounds(40, 8)
    stripes.check_bounds(b, 40)
    assert sum(1 for r0, r1 in b if r1 > r0) == 3


def test_balanced_bounds_minimise_the_heaviest_stripe():
    rng = np.random.default_rng(7)
    for H, world in ((1080, 2), (1080, 4), (1080, 8), (720, 3), (330, 4)):
        tr = stripes.tile_rows(H)
        w = rng.gamma(0.5, 1.0, tr) * np.exp(-((np.arange(tr) - tr / 2.0) / (tr / 6.0)) ** 2)  # centre-heavy
        b = stripes.stripe_bounds(H, world, w)
        stripes.check_bounds(b, H)
        loads = [w[r0 // 16:(r1 + 15) // 16].sum() for r0, r1 in b]
        # brute-force optimum by dynamic programming
        pre = np.concatenate([[0.0], np.cumsum(w)])
        best = np.full((world + 1, tr + 1), np.inf)
        best[0, 0] = 0.0
        for k in range(1, world + 1):
            for e in range(tr + 1):
                for s in range(e + 1):
                    best[k, e] = min(best[k, e], max(best[k - 1, s], pre[e] - pre[s]))
        assert max(loads) <= best[world, tr] * (1 + 1e-9) + 1e-9
        eq = stripes.stripe_bounds(H, world)
        assert max(loads) <= max(w[r0 // 16:(r1 + 15) // 16].sum() for r0, r1 in eq) + 1e-12


def test_bad_bounds_are_rejected():
    with pytest.raises(ValueError):
        stripes.check_bounds([(0, 100), (100, 330)], 330)       # not tile aligned
    with pytest.raises(ValueError):
        stripes.check_bounds([(0, 160), (176, 330)], 330)       # gap
    with pytest.raises(ValueError):
        stripes.stripe_bounds(330, 2, [1.0, 2.0])               # wrong number of rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, balanced, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from oracle import oracle as orc
    from splat_b200 import stripes as st
    from splat_b200.camera import Camera
    from splat_b200.gaussians import synthetic_scene

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H = 200, 150
        dev = torch.device("cpu")
        scene = synthetic_scene(1500, seed=0x5EED0041, log_scale_mean=-3.0) if rank == 0 else None
        scene = st.broadcast_scene(scene, rank, dev)                 # C0
        cam = Camera(H, W, (0.0, 0.0, 4.0))
        cam.update_camera_pose()
        ocam, cfg = orc.camera_from(cam), orc.make_config(lowpass=0.3, nthreads=2)
        sp = orc.project(scene, ocam, cfg, W, H)
        order = orc.sort_visible(sp)
        bounds = None
        if rank == 0:
            row_load = None
            if balanced:   # stand-in for Context.tile_loads: 3-sigma rects per tile row
                cy, hy = sp["cyp"], sp["bbox"][:, 1]
                vis = sp["visible"] != 0
                row_load = np.zeros(st.tile_rows(H))
                for t in range(len(row_load)):
                    row_load[t] = np.count_nonzero(vis & (cy + hy >= 16 * t) & (cy - hy < 16 * (t + 1)))
            bounds = st.stripe_bounds(H, world, row_load)
        bounds = st.broadcast_bounds(bounds, world, dev)
        st.check_bounds(bounds, H)
        r0, r1 = bounds[rank]
        fb = np.full((H, W), 0x00102030, np.uint32)                  # blended onto, not cleared
        if rank != 0:
            fb[:r0] = 0xDEADBEEF                                     # rows a rank does not own are garbage
            fb[r1:] = 0xDEADBEEF
        if r1 > r0:
            orc.rasterize_rows(sp, order, cfg, fb, np.arange(r0, r1))
        t = torch.from_numpy(fb.view(np.int32))
        st.gather_stripes(t, bounds, rank)                           # C1
        if rank == 0:
            # rank 0's own rows outside its stripe were still the initial value before the gather
            ref = np.full((H, W), 0x00102030, np.uint32)
            orc.rasterize_rows(sp, order, cfg, ref, np.arange(0, H))
            q.put((bool(np.array_equal(fb, ref)), int(np.count_nonzero(fb != ref)), bounds))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balanced", [False, True])
def test_world_size_2_gather_equals_full_frame(balanced):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, balanced, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        ok, bad, bounds = q.get(timeout=240)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    assert ok, f"{bad} pixels differ after the gather (bounds {bounds})"
    assert bounds[0][1] == bounds[1][0] and bounds[1][1] == 150


def test_rebalance_from_measured_times_moves_rows_to_the_fast_ranks():
    """stripes.rebalance: the cost of a tile row is taken as uniform inside the stripe that rendered
    it; re-cutting gives the slow rank fewer rows, keeps the cover tile aligned and gap free, and is a
    fixed point when all ranks take the same time."""
    from splat_b200 import stripes

    H = 1080
    b0 = stripes.stripe_bounds(H, 4)
    same = stripes.rebalance(b0, [1.0, 1.0, 1.0, 1.0], H)
    stripes.check_bounds(same, H)
    rows = lambda b: [r1 - r0 for r0, r1 in b]
    assert max(rows(same)) - min(rows(same)) <= 16 * 2
    b1 = stripes.rebalance(b0, [0.5, 2.0, 2.0, 0.5], H)      # the two centre stripes are 4x as slow
    stripes.check_bounds(b1, H)
    assert rows(b1)[1] < rows(b0)[1] and rows(b1)[2] < rows(b0)[2]
    assert rows(b1)[0] > rows(b0)[0] and rows(b1)[3] > rows(b0)[3]
    # predicted times under the uniform-within-stripe model are closer together than before
    def predict(bounds, old_bounds, t):
        dens = np.zeros((H + 15) // 16)
        for (r0, r1), tt in zip(old_bounds, t):
            t0, t1 = r0 // 16, (r1 + 15) // 16
            dens[t0:t1] = tt / max(t1 - t0, 1)
        return [dens[r0 // 16:(r1 + 15) // 16].sum() for r0, r1 in bounds]
    p = predict(b1, b0, [0.5, 2.0, 2.0, 0.5])
    assert max(p) < 2.0 * 0.8 and max(p) / min(p) < 1.35
    # an empty stripe in the input (more ranks than work) is tolerated
    b2 = stripes.rebalance([(0, 0), (0, 544), (544, 1080)], [0.0, 1.0, 1.0], H)
    stripes.check_bounds(b2, H)


def test_retries_are_local_and_every_frame_is_gathered_exactly_once():
    """The multi-rank harness logic that went wrong in round 2: a rank whose frame is abandoned on the
    device must repeat it with local work only.  A fake context abandons some frames (the NEXT call
    reports it, like splat_render_device / splat_get_timings do); whatever happens, the loop issues
    one gather per frame, every frame ends up rendered, and a non-retry error still propagates."""
    from splat_b200 import stripes

    class Retry(Exception):
        pass

    class FakeCtx:
        def __init__(self, abandon):
            self.abandon, self.pending, self.done, self.calls = set(abandon), False, [], 0

        def render(self, i):
            self.calls += 1
            if self.pending:                 # the previous frame was abandoned: reported once, nothing enqueued
                self.pending = False
                raise Retry()
            if i in self.abandon:
                self.abandon.discard(i)      # it will fit next time (buffers grown)
                self.pending = True
                return
            self.done.append(i)

        def timings(self):
            if self.pending:
                self.pending = False
                raise Retry()
            return {"total_ms": 1.0}

    is_retry = lambda e: isinstance(e, Retry)
    for abandon in ([], [3], [0, 1, 2], [5, 6, 9]):
        ctx, gathers, repeated = FakeCtx(abandon), 0, 0
        for i in range(10):
            repeated += stripes.render_with_retry(ctx.render, i, is_retry)
            gathers += 1                                     # exactly one collective per frame
            if i % 2:
                tm, rep = stripes.timings_with_retry(ctx.timings, ctx.render, i, is_retry)
                repeated += rep
                assert tm["total_ms"] == 1.0
        stripes.timings_with_retry(ctx.timings, ctx.render, 9, is_retry)
        assert gathers == 10
        assert set(ctx.done) == set(range(10)), (abandon, ctx.done)       # every frame was rendered in the end
        assert (repeated > 0) == bool(abandon)

    def boom(i):
        raise ValueError("not a retry")
    with pytest.raises(ValueError):
        stripes.render_with_retry(boom, 0, is_retry)


def test_library_partition_rules_agree_with_the_python_ones():
    """splat_api.cu's equal_bounds / rebalance_bounds (what a group context uses) through the host-only
    export splat_debug_partition, against stripes.stripe_bounds / stripes.rebalance: the same cover rules
    and, for the re-cut, the same bottleneck (the cuts themselves may differ where several are optimal)."""
    from splat_b200 import _lib

    rng = np.random.default_rng(7)
    for H in (16, 17, 150, 720, 1080, 2160):
        for parts in (1, 2, 3, 4, 8):
            eq = _lib.group_partition(H, parts)
            stripes.check_bounds(eq, H)
            assert eq == stripes.stripe_bounds(H, parts)
            for _ in range(6):
                times = rng.uniform(0.2, 3.0, parts)
                times[[k for k, (a, b) in enumerate(eq) if b == a]] = 0.0      # an empty stripe takes no time
                lib_b = _lib.group_partition(H, parts, eq, times)
                py_b = stripes.rebalance(eq, times, H)
                stripes.check_bounds(lib_b, H)

                # predicted time of the heaviest stripe under the model both sides use (uniform cost per tile row)
                tr = stripes.tile_rows(H)
                w = np.zeros(tr)
                for (a, b), t in zip(eq, times):
                    t0, t1 = a // 16, (b + 15) // 16
                    if t1 > t0:
                        w[t0:t1] = max(float(np.float32(t)), 1e-4) / (t1 - t0)

                def worst(bounds):
                    return max(w[a // 16:(b + 15) // 16].sum() for a, b in bounds)

                assert worst(lib_b) <= worst(py_b) * (1 + 1e-5) + 1e-9, (H, parts, times, lib_b, py_b)
                assert worst(py_b) <= worst(lib_b) * (1 + 1e-5) + 1e-9, (H, parts, times, lib_b, py_b)
                assert worst(lib_b) <= worst(eq) * (1 + 1e-6)                  # never worse than what it started from
