"""Screen-tile stripe sharding for N GPUs: one process per GPU, every rank owns a
horizontal stripe of whole 16-pixel tile rows, renders it straight into its rows of a
full-frame buffer, and one grouped send/recv gathers the stripes into rank 0's frame.

The reference has no multi-device path; its only parallelism is euc's row-group threading
(SURVEY 8c E6), which this mirrors across GPUs.  Pixels are independent given the global
far -> near order, so the gathered frame is byte-identical for every partition.

Only torch.distributed plumbing lives here (works with NCCL on CUDA tensors and with gloo on
CPU tensors -- the latter is how tests/test_stripes_gloo.py covers it without a GPU).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

TILE = 16

Bounds = List[Tuple[int, int]]


def tile_rows(H: int) -> int:
    return (H + TILE - 1) // TILE


def stripe_bounds(H: int, world: int, row_load: Optional[Sequence[float]] = None) -> Bounds:
    """Pixel-row ranges [row0, row1) per rank, aligned to tile rows and covering [0, H).

    Without `row_load`: equal numbers of tile rows (remainder to the first ranks).  With
    `row_load` (one weight per tile row, e.g. tile instances per row from
    `Context.tile_loads`): contiguous partition that minimises the heaviest stripe (H6: the
    object is off-centre in every demo camera, so equal stripes are badly unbalanced).  Ranks
    beyond the number of tile rows get empty stripes (row0 == row1)."""
    tr = tile_rows(H)
    if world < 1:
        raise ValueError("world must be >= 1")
    if row_load is None:
        base, rem = divmod(tr, world)
        cuts, r = [0], 0
        for k in range(world):
            r += base + (1 if k < rem else 0)
            cuts.append(r)
    else:
        w = np.asarray(row_load, dtype=np.float64).reshape(-1)
        if len(w) != tr:
            raise ValueError(f"row_load has {len(w)} entries, image has {tr} tile rows")
        if np.any(w < 0) or not np.all(np.isfinite(w)):
            raise ValueError("row_load must be finite and non-negative")
        cuts = _min_max_partition(w, world)
    return [(min(cuts[k] * TILE, H), min(cuts[k + 1] * TILE, H)) for k in range(world)]


def _min_max_partition(w: np.ndarray, parts: int) -> List[int]:
    """Cut indices (len parts+1) of the contiguous partition of w into `parts` pieces with the
    smallest possible maximum piece sum (binary search on the bottleneck + greedy fill)."""
    n = len(w)
    pre = np.concatenate([[0.0], np.cumsum(w)])

    def cuts_for(cap: float):
        cuts, start = [0], 0
        for _ in range(parts):
            # furthest end with sum(w[start:end]) <= cap
            end = int(np.searchsorted(pre, pre[start] + cap, side="right")) - 1
            end = max(end, start)
            if end == start and start < n:
                return None          # a single row exceeds cap
            cuts.append(min(end, n))
            start = cuts[-1]
        return cuts if cuts[-1] >= n else None

    lo, hi = float(w.max(initial=0.0)), float(pre[-1])
    if hi == 0.0:
        return stripe_cuts_equal(n, parts)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if cuts_for(mid) is None:
            lo = mid
        else:
            hi = mid
    cuts = cuts_for(hi * (1.0 + 1e-12) + 1e-9)
    assert cuts is not None
    cuts[-1] = n
    return cuts


def rebalance(bounds: Bounds, times_ms: Sequence[float], H: int) -> Bounds:
    """New stripe boundaries from the ranks' measured frame times (the same rule as the library's
    group contexts, splat_api.cu rebalance_bounds): the cost of a tile row is taken as uniform inside
    the stripe that rendered it, and the rows are re-cut with the min-max partition."""
    tr = tile_rows(H)
    w = np.zeros(tr, np.float64)
    for (r0, r1), t in zip(bounds, times_ms):
        t0, t1 = r0 // TILE, (r1 + TILE - 1) // TILE
        if t1 > t0:
            w[t0:t1] = max(float(t), 1e-4) / (t1 - t0)
    return stripe_bounds(H, len(bounds), w)


RETRY = -6   # SPLAT_ERR_RETRY (include/splat.h)


def render_with_retry(render, i: int, is_retry, max_attempts: int = 6) -> int:
    """One rank's stripe of frame i on the no-host-wait path.  `render(k)` enqueues frame k and may
    raise an error for which `is_retry(err)` is true: the library then says that the PREVIOUS frame
    was abandoned on the device (it outgrew its launch bounds or its near lists) and has grown its
    buffers.  That is handled here with LOCAL work only -- render the abandoned frame again, then
    frame i -- never with a collective, so that every rank still issues exactly one gather per frame
    whatever happened to it (round 2: an unhandled retry on one rank left seven others waiting in the
    gather).  Returns the number of frames that had to be repeated."""
    repeated = 0
    for attempt in range(max_attempts):
        try:
            render(i)
            return repeated
        except Exception as e:   # noqa: BLE001
            if not is_retry(e) or attempt == max_attempts - 1:
                raise
            repeated += 1
            try:
                render(max(i - 1, 0))
            except Exception as e2:   # noqa: BLE001
                if not is_retry(e2):
                    raise
    return repeated


def timings_with_retry(timings, render, i: int, is_retry, max_attempts: int = 6):
    """Stage times of this rank's stripe of frame i (`timings()` waits for it).  If that very frame was
    abandoned on the device it is rendered again, locally.  Returns (timings, frames repeated)."""
    repeated = 0
    for attempt in range(max_attempts):
        try:
            return timings(), repeated
        except Exception as e:   # noqa: BLE001
            if not is_retry(e) or attempt == max_attempts - 1:
                raise
            repeated += 1 + render_with_retry(render, i, is_retry, max_attempts)
    raise AssertionError("unreachable")


def stripe_cuts_equal(n: int, parts: int) -> List[int]:
    base, rem = divmod(n, parts)
    cuts, r = [0], 0
    for k in range(parts):
        r += base + (1 if k < rem else 0)
        cuts.append(r)
    return cuts


def check_bounds(bounds: Bounds, H: int) -> None:
    """Raises unless `bounds` is a tile-aligned, ordered, gap-free cover of [0, H)."""
    prev = 0
    for r0, r1 in bounds:
        if r0 != prev or r1 < r0:
            raise ValueError(f"stripes must be contiguous and ordered: {bounds}")
        if r1 > r0 and (r0 % TILE != 0 or (r1 % TILE != 0 and r1 != H)):
            raise ValueError(f"stripe [{r0},{r1}) is not tile aligned")
        prev = r1
    if prev != H:
        raise ValueError(f"stripes end at row {prev}, image has {H} rows")


def gather_stripes(fb, bounds: Bounds, rank: int, root: int = 0) -> None:
    """C1: every rank's rows fb[row0:row1] -> the same rows of `root`'s fb, in place, no staging.

    fb: a (H, W) torch tensor on every rank (int32 view of the 0xAARRGGBB pixels).  Ranks with an
    empty stripe neither send nor receive.  With NCCL the transfers are enqueued on the current
    stream right behind the blend kernel that wrote the rows."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    if rank == root:
        ops = [dist.P2POp(dist.irecv, fb[b0:b1], k) for k, (b0, b1) in enumerate(bounds) if k != root and b1 > b0]
    else:
        r0, r1 = bounds[rank]
        ops = [dist.P2POp(dist.isend, fb[r0:r1], root)] if r1 > r0 else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def broadcast_bounds(bounds: Optional[Bounds], world: int, device, root: int = 0) -> Bounds:
    """Rank `root` decides the partition (from its tile-load probe); everyone adopts it."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(2 * world, dtype=torch.int64, device=device)
    if bounds is not None:
        t.copy_(torch.tensor([v for b in bounds for v in b], dtype=torch.int64))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, root)
    v = [int(x) for x in t.cpu().tolist()]
    return [(v[2 * k], v[2 * k + 1]) for k in range(world)]


def broadcast_scene(scene, rank: int, device, root: int = 0):
    """C0: the five GaussianList arrays, generated/loaded on `root`, broadcast once."""
    import torch
    import torch.distributed as dist

    from .gaussians import GaussianList

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return scene
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == root:
        n[0] = scene.num_gaussians
    dist.broadcast(n, root)
    n = int(n.item())
    shapes = [(n, 4), (n, 3), (n,), (n, 4), (n, 48)]
    host = [scene.positions, scene.scales, scene.opacities, scene.rotations, scene.sh] if rank == root else None
    out = []
    for i, shp in enumerate(shapes):
        t = torch.from_numpy(host[i]).to(device) if rank == root else torch.empty(shp, dtype=torch.float32, device=device)
        dist.broadcast(t, root)
        out.append(t.cpu().numpy())
        del t
    return GaussianList(*out)
