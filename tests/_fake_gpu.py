"""Stand-ins that let bench.py's harness run on a box without a GPU (tests/test_bench_flow.py): torch.cuda's
streams/events become no-ops and wall-clock timers, `torch.device` always answers "cpu", the process group
is gloo, and splat_b200._lib.Context is replaced by a class with the same methods that renders with the CPU
oracle (test infrastructure may call the oracle).  Nothing here is reachable from the product path."""
import ctypes
import time
import types

import numpy as np


class FakeStream:
    cuda_stream = 1

    def __init__(self, *a, **k):
        pass


class FakeEvent:
    def __init__(self, *a, **k):
        self.t = 0.0

    def record(self, *a):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


def _rows(ptr, rows, W):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(rows, W))


def install_emulated(setattr_, emu_path, make_scene=None):
    """Like install(), but the library is the REAL one -- splat_api.cu and the kernels compiled for the host
    (tests/cuda_emu) -- so frames, counts, timings and retry semantics are the product's own.  Only what needs NCCL is
    replaced: the scene broadcast (every rank generates the same seeded scene) and the stripe gather (gloo)."""
    import torch
    import torch.distributed as dist

    from splat_b200 import _lib, stripes

    _patch_torch(setattr_, torch, dist)
    setattr_(_lib, "LIB_PATH", emu_path)
    setattr_(_lib, "_lib", None)
    _lib.load()
    log = types.SimpleNamespace(gathers=0, contexts=[])
    Real = _lib.Context

    class EmuContext(Real):
        def __init__(self, device=0, **kw):
            super().__init__(device=0, **kw)              # the emulated box has one device
            self.kw, self.rank = dict(kw), 0
            log.contexts.append(self)

        def unique_id(self):
            return bytes(128)

        def comm_init(self, uid, n_ranks, rank):
            self.rank = rank

        def broadcast_scene(self, root, n):
            if self.rank != root:
                self.upload(make_scene(n))
            self.n = n

        def gather_stripes(self, ptr, W, H, bounds, root=0, stream=0):
            log.gathers += 1
            t = torch.from_numpy(_rows(ptr, H, W).view(np.int32))
            stripes.gather_stripes(t, [tuple(b) for b in bounds], self.rank, root)

    setattr_(_lib, "Context", EmuContext)
    return log


def _patch_torch(setattr_, torch, dist):
    real_device = torch.device
    real_init = dist.init_process_group
    setattr_(torch, "device", lambda *a, **k: real_device("cpu"))
    setattr_(torch.cuda, "set_device", lambda *a, **k: None)
    setattr_(torch.cuda, "Stream", FakeStream)
    setattr_(torch.cuda, "set_stream", lambda *a, **k: None)
    setattr_(torch.cuda, "Event", FakeEvent)
    setattr_(torch.cuda, "synchronize", lambda *a, **k: None)
    setattr_(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    setattr_(dist, "init_process_group", lambda backend=None, **k: real_init("gloo"))


def install(setattr_, orc, abandon_at=(3,), make_scene=None):
    """setattr_(obj, name, value) applies one patch (pytest's monkeypatch.setattr, or plain setattr in a spawned
    worker).  `abandon_at`: ordinal numbers of render_device calls after which the NEXT call (render or
    timings) reports SPLAT_ERR_RETRY once, the way the library reports a frame abandoned on the device."""
    import torch
    import torch.distributed as dist

    from splat_b200 import _lib, stripes

    _patch_torch(setattr_, torch, dist)
    log = types.SimpleNamespace(gathers=0, renders=0, retries_reported=0, contexts=[])

    class FakeContext:
        def __init__(self, **kw):
            self.kw, self.scene, self.n = kw, None, 0
            self.pending = False
            self.rank, self.world = 0, 1
            self.h, self.L = None, types.SimpleNamespace(splat_render_cleared=self._cleared)
            log.contexts.append(self)

        # ---- scene
        def upload(self, sc):
            self.scene, self.n = sc, sc.num_gaussians

        def unique_id(self):
            return bytes(128)

        def comm_init(self, uid, n_ranks, rank):
            assert len(uid) == 128
            self.world, self.rank = n_ranks, rank

        def broadcast_scene(self, root, n):
            if self.scene is None:
                self.scene = make_scene(n)        # the same seeded scene the root generated
            self.n = n

        # ---- frames
        def _cam(self, cs):
            view = np.array(cs.view, np.float32).reshape(4, 4).T
            proj = np.array(cs.proj, np.float32).reshape(4, 4).T
            return orc.make_camera(view, proj, list(cs.position), cs.w, cs.h, cs.htanx, cs.htany, cs.focal)

        def _frame(self, cs, W, H):
            fb = np.zeros((H, W), np.uint32)
            orc.render(self.scene, self._cam(cs), orc.make_config(lowpass=0.3, nthreads=2), fb)
            return fb

        def _maybe_retry(self):
            if self.pending:
                self.pending = False
                log.retries_reported += 1
                raise _lib.SplatError(-6, "a frame was abandoned on the device")

        def render_device(self, cs, ptr, W, H, r0=0, r1=None, stream=0):
            r1 = H if r1 is None else r1
            self.last_rows = r1 - r0
            self._maybe_retry()
            log.renders += 1
            if log.renders in abandon_at:
                assert not self.kw.get("sync_frames"), "a frame that reads its counts on the host is never abandoned"
                self.pending = True
                return                                # its pixels are never written
            _rows(ptr, r1 - r0, W)[:] = self._frame(cs, W, H)[r0:r1]     # the target rows arrive cleared

        def render_ptr(self, cs, host_ptr, W, H, row0=0, row1=None):
            _rows(host_ptr, H, W)[:] = self._frame(cs, W, H)

        def render(self, cs, fb, row0=0, row1=None):
            fb[:] = self._frame(cs, fb.shape[1], fb.shape[0])

        def _cleared(self, h, cs, host_ptr, W, H, clear):
            assert clear == 0
            _rows(host_ptr, H, W)[:] = self._frame(cs, W, H)
            return 0

        def _check(self, rc):
            assert rc == 0

        def gather_stripes(self, ptr, W, H, bounds, root=0, stream=0):
            log.gathers += 1
            t = torch.from_numpy(_rows(ptr, H, W).view(np.int32))
            stripes.gather_stripes(t, [tuple(b) for b in bounds], self.rank, root)

        def tile_loads(self, tiles_x):
            return np.ones(((self.last_rows + 15) // 16, tiles_x), np.uint32)      # of the last rendered stripe

        def timings(self):
            self._maybe_retry()
            return {"project_ms": 0.1, "sort_ms": 0.1, "bin_ms": 0.1, "blend_ms": 0.2, "second_pass_ms": 0.0, "total_ms": 0.5 + 0.25 * self.rank,
                    "n_instances": 1000, "near_cut_instances": 0, "second_pass_instances": 0, "near_cut_failed": 0,
                    "kernel_launches": 20, "frames_skipped": log.retries_reported, "n_tiles": 60, "n_visible": 1000}

        def close(self):
            pass

    setattr_(_lib, "Context", FakeContext)
    setattr_(_lib, "load", lambda: None)
    setattr_(_lib, "LIB_PATH", __file__)       # "exists": no build attempt
    return log


def quiet_bench(setattr_, bench, sink=None):
    """bench.py without the nvidia-smi sampler and the fd-level stdout juggling; `sink(line)` receives the JSON line."""
    setattr_(bench, "capture_stdout", lambda: None)
    setattr_(bench.ClockSampler, "start", lambda self: None)
    setattr_(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1.0, "sm_max_mhz": 1.0, "reasons": []})
    if sink is not None:
        setattr_(bench, "emit", sink)
