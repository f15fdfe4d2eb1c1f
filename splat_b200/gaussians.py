"""Host-side mirror of the reference's scene structs (reference: src/gaussians.rs).

`Gaussian` (gaussians.rs:31-38) and `GaussianList` (:408-416) keep the reference's field
names and memory layout: every per-Gaussian attribute vector is contiguous, i.e. a numpy
array of shape (N, k) in C order is byte-identical to nalgebra's column-major k x N matrix,
so the same pointers can be handed to the C ABI (`splat_upload_soa`) from Rust or Python.

No rendering maths lives here: cov3d, projection, SH colour, sort and blend all run in the
CUDA library.  This module only builds / loads scenes.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import List

import numpy as np

f32 = np.float32


@dataclasses.dataclass
class Gaussian:
    """AoS scene element, gaussians.rs:31-38 (`cov3d` is derived; the device recomputes it)."""

    position: np.ndarray  # (3,)
    scale: np.ndarray  # (3,) already exp()'d
    opacity: float  # already sigmoid()'d
    rotation: np.ndarray  # (4,) nalgebra coords order (i, j, k, w)
    sh: np.ndarray  # (48,) f_dc_0..2 then f_rest_0..44, as stored

    @staticmethod
    def new() -> "Gaussian":
        # PropertyAccess::new, gaussians.rs:247-256
        return Gaussian(
            position=np.zeros(3, f32),
            scale=np.zeros(3, f32),
            opacity=0.0,
            rotation=np.array([0, 0, 0, 1], f32),
            sh=np.zeros(48, f32),
        )


class GaussianList:
    """SoA scene, gaussians.rs:408-416.  Public arrays: positions (N,4) xyz1, scales (N,3),
    opacities (N,), rotations (N,4) ijkw, sh (N,48)."""

    def __init__(self, positions, scales, opacities, rotations, sh):
        self.positions = np.ascontiguousarray(positions, dtype=f32)
        self.scales = np.ascontiguousarray(scales, dtype=f32)
        self.opacities = np.ascontiguousarray(opacities, dtype=f32).reshape(-1)
        self.rotations = np.ascontiguousarray(rotations, dtype=f32)
        self.sh = np.ascontiguousarray(sh, dtype=f32)
        self.num_gaussians = int(self.positions.shape[0])
        n = self.num_gaussians
        assert self.positions.shape == (n, 4) and self.scales.shape == (n, 3)
        assert self.opacities.shape == (n,) and self.rotations.shape == (n, 4)
        assert self.sh.shape == (n, 48)

    @staticmethod
    def from_vec(gaussians: List[Gaussian]) -> "GaussianList":
        # gaussians.rs:419-440
        n = len(gaussians)
        pos = np.ones((n, 4), f32)
        sc = np.zeros((n, 3), f32)
        op = np.zeros(n, f32)
        rot = np.zeros((n, 4), f32)
        sh = np.zeros((n, 48), f32)
        for i, g in enumerate(gaussians):
            pos[i, :3] = g.position
            sc[i] = g.scale
            op[i] = g.opacity
            rot[i] = g.rotation
            sh[i] = g.sh
        return GaussianList(pos, sc, op, rot, sh)

    @staticmethod
    def naive_gaussians() -> "GaussianList":
        # gaussians.rs:442-445
        return GaussianList.from_vec(naive_gaussians())

    def to_vec(self) -> List[Gaussian]:
        return [
            Gaussian(self.positions[i, :3].copy(), self.scales[i].copy(), float(self.opacities[i]),
                     self.rotations[i].copy(), self.sh[i].copy())
            for i in range(self.num_gaussians)
        ]

    def subset(self, idx) -> "GaussianList":
        return GaussianList(self.positions[idx], self.scales[idx], self.opacities[idx],
                            self.rotations[idx], self.sh[idx])


def naive_gaussians() -> List[Gaussian]:
    """The 4-Gaussian test scene, gaussians.rs:319-374 (the 0.28209 literal at :330 is kept)."""
    out = []
    specs = [
        ((0.0, 0.0, 0.0), (0.03, 0.03, 0.03), (1.0, 0.0, 1.0)),
        ((1.0, 0.0, 0.0), (0.2, 0.03, 0.03), (1.0, 0.0, 0.0)),
        ((0.0, 1.0, 0.0), (0.03, 0.2, 0.03), (0.0, 1.0, 0.0)),
        ((0.0, 0.0, 1.0), (0.03, 0.03, 0.2), (0.0, 0.0, 1.0)),
    ]
    for pos, scale, color in specs:
        g = Gaussian.new()
        g.position = np.array(pos, f32)
        g.scale = np.array(scale, f32)
        g.opacity = 1.0
        g.rotation = np.array([0.0, 0.0, 0.0, 1.0], f32)  # Quaternion::new(w=1, 0, 0, 0)
        c = ((np.array(color, f32) - f32(0.5)) / f32(0.28209)).astype(f32)
        g.sh[:3] = c
        out.append(g)
    return out


# --------------------------------------------------------------------------- synthetic scenes

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def _splitmix64(seed: int, stream: int, a: int, b: int) -> np.ndarray:
    """outputs a+1 .. b of splitmix64 started at seed + stream*2^40 (counter based)."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed) + np.uint64(stream) * np.uint64(1 << 40)
        z = base + (np.arange(a + 1, b + 1, dtype=np.uint64)) * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform(seed, stream, a, b):
    return ((_splitmix64(seed, stream, a, b) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / (1 << 53))


def _normal(seed, stream, a, b):
    u1 = _uniform(seed, 2 * stream, a, b)
    u2 = _uniform(seed, 2 * stream + 1, a, b)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def _synthetic_range(seed: int, log_scale_mean: float, a: int, b: int, out) -> None:
    """Gaussians [a, b) of synthetic_scene, written into the preallocated f32 arrays `out`.  The
    generator is counter based, so any split into ranges gives the same bytes."""
    n = b - a
    s = 0
    def nrm(k=1):
        nonlocal s
        cols = []
        for _ in range(k):
            cols.append(_normal(seed, s, a, b)); s += 1
        return np.stack(cols, axis=1) if k > 1 else cols[0]
    def uni(k=1):
        nonlocal s
        cols = []
        for _ in range(k):
            cols.append(_uniform(seed, 1000 + s, a, b)); s += 1
        return np.stack(cols, axis=1) if k > 1 else cols[0]

    pos, scales, opac, rot, sh = out
    is_obj = uni() < 0.7
    obj = nrm(3) * 0.6
    bg = uni(3) * 8.0 - 4.0
    pos[a:b, :3] = np.where(is_obj[:, None], obj, bg)
    pos[a:b, 3] = 1.0
    base = nrm() * 0.8 + log_scale_mean
    scales[a:b] = np.exp(base[:, None] + nrm(3) * 0.5)
    q = nrm(4)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rot[a:b] = q
    opac[a:b] = 1.0 / (1.0 + np.exp(-(nrm() * 2.0 + 0.5)))
    sh[a:b, 0:3] = (nrm(3) * 0.35 + 0.5 - 0.5) / 0.28209479177387814
    sh[a:b, 3:27] = nrm(24) * 0.15
    assert n >= 0


def synthetic_scene(n: int, seed: int = 0x5EED0000, log_scale_mean: float = -4.0) -> GaussianList:
    """Deterministic stand-in for a trained scene (real .ply files are not available offline,
    SURVEY 8d): 70% of the Gaussians form an "object" blob N(0, 0.6^2 I), 30% a "background"
    U([-4,4]^3); log-scale N(log_scale_mean, 0.8^2) per Gaussian plus N(0, 0.5^2) per axis;
    random unit rotations; opacity sigmoid(N(0.5, 2^2)); DC colour N(0.5, 0.35^2) (a few
    percent outside [0,1] to exercise the unclamped-colour / saturating-cast path); 24
    higher-order SH coefficients N(0, 0.15^2); degree-3 coefficients zero.  Large scenes are
    generated in ranges on a thread pool (numpy releases the GIL); the bytes do not depend on
    the split."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    out = (np.zeros((n, 4), f32), np.zeros((n, 3), f32), np.zeros(n, f32), np.zeros((n, 4), f32), np.zeros((n, 48), f32))
    step = 1 << 17
    ranges = [(a, min(a + step, n)) for a in range(0, n, step)]
    workers = min(len(ranges), os.cpu_count() or 1)
    if workers <= 1:
        for a, b in ranges:
            _synthetic_range(seed, log_scale_mean, a, b, out)
    else:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(lambda r: _synthetic_range(seed, log_scale_mean, r[0], r[1], out), ranges))
    return GaussianList(*out)


# --------------------------------------------------------------------------- PLY (f-1)

_PLY_PROPS = (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
              + [f"f_rest_{i}" for i in range(45)] + ["opacity", "scale_0", "scale_1", "scale_2",
                                                       "rot_0", "rot_1", "rot_2", "rot_3"])


def save_ply(path: str, raw: dict) -> None:
    """Write an INRIA-3DGS binary_little_endian PLY (62 float properties, notes.md:1-8) from
    *raw* (pre-activation) property arrays; missing properties are written as zeros."""
    n = len(raw["x"])
    arr = np.zeros((n, len(_PLY_PROPS)), dtype="<f4")
    for k, name in enumerate(_PLY_PROPS):
        if name in raw:
            arr[:, k] = raw[name]
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n
    header += "".join(f"property float {p}\n" for p in _PLY_PROPS) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(arr.tobytes())


def load_from_ply(filename: str) -> List[Gaussian]:
    """`load_from_ply`, gaussians.rs:375-405 with `set_property` :258-282: scale -> exp,
    opacity -> 1/(1+exp(-v)), rot_0 -> w, f_rest_i -> sh[3+i] (no channel transpose), then the
    mean position (sequential f32 sum, :394-399) is subtracted."""
    return load_ply_soa(filename).to_vec()


def load_ply_soa(filename: str) -> GaussianList:
    with open(filename, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").split("\n")
    if lines[0].strip() != "ply":
        raise ValueError("not a PLY file")
    fmt, n, props, in_vertex = None, 0, [], False
    sizes = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1",
             "uint8": "u1", "char": "i1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4"}
    for ln in lines[1:]:
        t = ln.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            if t[1] != "vertex":
                raise ValueError("Unexpected element!")  # gaussians.rs:390 panics
            n, in_vertex = int(t[2]), True
        elif t[0] == "property" and in_vertex:
            props.append((t[2], sizes[t[1]]))
    if fmt == "binary_little_endian":
        dt = np.dtype([(nm, "<" + ty) for nm, ty in props])
        rec = np.frombuffer(data, dtype=dt, count=n, offset=end)
        col = {nm: rec[nm].astype(f32) for nm, _ in props}
    elif fmt == "ascii":
        vals = np.array(data[end:].split()[: n * len(props)], dtype=np.float64).reshape(n, len(props))
        col = {nm: vals[:, k].astype(f32) for k, (nm, _) in enumerate(props)}
    else:
        raise ValueError(f"unsupported PLY format {fmt}")
    zero = np.zeros(n, f32)
    g = lambda nm: col.get(nm, zero)
    pos = np.ones((n, 4), f32)
    pos[:, 0], pos[:, 1], pos[:, 2] = g("x"), g("y"), g("z")
    scales = np.stack([np.exp(g(f"scale_{i}")) for i in range(3)], axis=1).astype(f32)
    opac = (f32(1.0) / (f32(1.0) + np.exp(-g("opacity")))).astype(f32)
    # rot_0 -> rotation[3] (w), rot_1..3 -> i, j, k  (gaussians.rs:269-272)
    rot = np.stack([g("rot_1"), g("rot_2"), g("rot_3"), g("rot_0")], axis=1).astype(f32)
    if "rot_0" not in col:
        rot[:, 3] = 1.0  # Quaternion::identity()
    sh = np.zeros((n, 48), f32)
    for i in range(3):
        sh[:, i] = g(f"f_dc_{i}")
    for i in range(45):
        sh[:, 3 + i] = g(f"f_rest_{i}")
    if n:
        # sequential f32 accumulation, then one division (gaussians.rs:395-399)
        avg = np.array([np.cumsum(pos[:, k], dtype=f32)[-1] for k in range(3)], f32) / f32(n)
        pos[:, :3] = pos[:, :3] - avg
    return GaussianList(pos, scales, opac, rot, sh)


def ply_vertex_payload(filename: str):
    """(rows, n): the vertex payload of a binary_little_endian PLY in the INRIA 3DGS layout as an
    (n, 62) float32 memory map of the file -- what `splat_upload_ply_raw` takes.  Raises when the file
    has another element, property order or format (use load_ply_soa for those)."""
    with open(filename, "rb") as f:
        head = f.read(1 << 16)
    end = head.index(b"end_header\n") + len(b"end_header\n")
    lines = head[:end].decode("ascii").split("\n")
    n, props, fmt = 0, [], None
    for ln in lines[1:]:
        t = ln.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            if t[1] != "vertex":
                raise ValueError("Unexpected element!")
            n = int(t[2])
        elif t[0] == "property":
            props.append((t[2], t[1]))
    if fmt != "binary_little_endian" or [p for p, _ in props] != _PLY_PROPS or any(ty not in ("float", "float32") for _, ty in props):
        raise ValueError("not the INRIA 62-float binary_little_endian vertex layout")
    rows = np.memmap(filename, dtype="<f4", mode="r", offset=end, shape=(n, len(_PLY_PROPS)))
    return rows, n


def trim_ply(src: str, dst: str, count: int = 3) -> int:
    """`trim` (src/bin/00_ply_load.rs:9-63): copy the first `count` vertices of a binary
    little-endian PLY into a new PLY with the same header otherwise (tiny test scenes).  Returns
    the number of vertices written."""
    with open(src, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").split("\n")
    sizes = {"float": 4, "float32": 4, "double": 8, "float64": 8, "uchar": 1, "uint8": 1, "char": 1,
             "short": 2, "ushort": 2, "int": 4, "uint": 4}
    n, stride, out_lines = 0, 0, []
    for ln in lines:
        t = ln.split()
        if len(t) == 3 and t[0] == "element":
            if t[1] != "vertex":
                raise ValueError("Unexpected element!")
            n = int(t[2])
            ln = f"element vertex {min(n, count)}"
        elif len(t) == 3 and t[0] == "property":
            stride += sizes[t[1]]
        elif len(t) >= 2 and t[0] == "format" and t[1] != "binary_little_endian":
            raise ValueError("trim_ply handles binary_little_endian files")
        out_lines.append(ln)
    m = min(n, count)
    with open(dst, "wb") as f:
        f.write("\n".join(out_lines).encode("ascii"))
        f.write(data[end:end + m * stride])
    return m
