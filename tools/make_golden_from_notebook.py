#!/usr/bin/env python
"""Generate tests/golden/notebook_projection.json by EXECUTING the reference's own Python
prototype (reference: notes/00_Gaussian_Projection.ipynb cell 1, notes/util.py,
notes/util_gau.py) on seeded inputs.  Runs only in the build container where
/root/reference is mounted; the JSON it writes is committed so the tests never need it.

Nothing is copied out of the reference: the notebook cell and the two helper modules are
imported / exec'd from where they lie.  Their missing third-party imports (PyGLM, PyOpenGL,
plyfile -- not installed, no network) are satisfied by tiny stand-in modules defined below;
the `glm` stand-in is validated by the notebook's own stored cell-2 output (the four conics
must come out as 0.07541478 / 0.00173521 / 0.03394433, which this script asserts).

Two deliberate adaptations, both documented in the emitted JSON:
  * the prototype builds cov3D = R^T S R (notebook cell 1, compute_cov3d) while the Rust
    renderer builds R S R^T (gaussians.rs:111); the script feeds the prototype the CONJUGATE
    quaternion so that both describe the same covariance;
  * the prototype evaluates SH degree 3 when handed 48 coefficients and clips colours to
    [0,1]; the Rust renderer stops at degree 2 (sh_dim = 15, pipelines.rs:100) and does not
    clip.  The script hands the prototype the first 27 coefficients and records the clipped
    colour; the test clips the oracle's colour before comparing.
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "notebook_projection.json")


def _install_standins():
    glm = types.ModuleType("glm")

    def _n(v):
        return v / np.linalg.norm(v)

    def lookAt(eye, center, up):  # right-handed, like GLM's default
        eye, center, up = (np.asarray(a, np.float64) for a in (eye, center, up))
        f = _n(center - eye)
        s = _n(np.cross(f, up))
        u = np.cross(s, f)
        m = np.eye(4)
        m[0, :3], m[1, :3], m[2, :3] = s, u, -f
        m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, f @ eye
        return m

    def perspective(fovy, aspect, near, far):  # GLM default: RH, NDC z in [-1,1]
        t = np.tan(fovy / 2.0)
        m = np.zeros((4, 4))
        m[0, 0] = 1.0 / (aspect * t)
        m[1, 1] = 1.0 / t
        m[2, 2] = -(far + near) / (far - near)
        m[2, 3] = -(2.0 * far * near) / (far - near)
        m[3, 2] = -1.0
        return m

    glm.lookAt, glm.perspective = lookAt, perspective
    sys.modules["glm"] = glm
    ogl, gl, sh = types.ModuleType("OpenGL"), types.ModuleType("OpenGL.GL"), types.ModuleType("OpenGL.GL.shaders")
    ogl.GL, gl.shaders = gl, sh
    sys.modules.update({"OpenGL": ogl, "OpenGL.GL": gl, "OpenGL.GL.shaders": sh})
    ply = types.ModuleType("plyfile")
    ply.PlyData = object
    sys.modules["plyfile"] = ply


def main():
    _install_standins()
    sys.path.insert(0, os.path.join(REF, "notes"))
    import util  # noqa: reference module, unmodified
    import util_gau  # noqa: reference module, unmodified

    nb = json.load(open(os.path.join(REF, "notes", "00_Gaussian_Projection.ipynb")))
    cell1 = "".join(nb["cells"][1]["source"])
    ns = {"np": np, "sp": __import__("scipy"), "util": util, "Camera": util.Camera,
          "naive_gaussian": util_gau.naive_gaussian, "GaussianData": util_gau.GaussianData}
    import scipy.spatial.transform  # noqa: make sp.spatial.transform resolvable
    exec(compile(cell1, "notebook_cell_1", "exec"), ns)
    NBGaussian = ns["Gaussian"]

    def run_case(name, xyz, rot_wxyz, scale, opacity, sh27, cam_pos, w, h):
        cam = util.Camera(h, w, position=cam_pos)
        rec = {"name": name, "w": w, "h": h, "cam_pos": list(map(float, cam_pos)),
               "view": np.asarray(cam.get_view_matrix(), np.float64).tolist(),
               "proj": np.asarray(cam.get_projection_matrix(), np.float64).tolist(),
               "htanfovxy_focal": [float(v) for v in cam.get_htanfovxy_focal()],
               "xyz": xyz.tolist(), "rot_wxyz": rot_wxyz.tolist(), "scale": scale.tolist(),
               "opacity": opacity.tolist(), "sh27": sh27.tolist(), "out": []}
        for i in range(len(xyz)):
            q = rot_wxyz[i].astype(np.float64)
            q_conj = np.array([q[0], -q[1], -q[2], -q[3]])  # see module docstring
            g = NBGaussian(xyz[i], scale[i], q_conj, opacity[i:i + 1], sh27[i])
            cov2d = g.get_cov2d(cam)
            res = g.get_conic_and_bb(cam)
            d = np.asarray(xyz[i], np.float64) - cam.position
            d = d / np.linalg.norm(d)
            o = {"cov3d": np.asarray(g.cov3D, np.float64).tolist(),
                 "cov2d": np.asarray(cov2d, np.float64).tolist(),
                 "depth": float(g.get_depth(cam)),
                 "color_clipped": np.asarray(g.get_color(d), np.float64).tolist()}
            if res is not None:
                conic, bbox_cam, bbox_ndc = res
                o["conic"] = np.asarray(conic, np.float64).tolist()
                o["bbox_cam"] = np.abs(np.asarray(bbox_cam, np.float64)[0]).tolist()
                # centre in NDC = mean of the four corners; z,w are shared
                o["ndc"] = np.asarray(bbox_ndc, np.float64).mean(axis=0).tolist()
            rec["out"].append(o)
        return rec

    cases = []
    # (1) the reference's 4-Gaussian scene with the notebook's camera: Camera(720, 1280)
    nv = util_gau.naive_gaussian()
    sh27 = np.zeros((4, 27), np.float32)
    sh27[:, :3] = nv.sh
    c = run_case("naive_cam003", nv.xyz, nv.rot, nv.scale, nv.opacity[:, 0], sh27, (0.0, 0.0, 3.0), 1280, 720)
    stored = [[0.07541478, 0.0, 0.07541478], [0.00173521, 0.0, 0.07541478],
              [0.07541478, 0.0, 0.00173521], [0.03394433, 0.0, 0.03394433]]  # cell-2 stored output
    got = np.array([o["conic"] for o in c["out"]])
    assert np.allclose(got, stored, rtol=2e-6, atol=1e-9), (got, stored)
    c["stored_cell2_conics"] = stored
    cases.append(c)
    # (2) seeded random Gaussians, three cameras (incl. the demo camera of 02_ply_demo.rs:22)
    rng = np.random.default_rng(20261017)
    n = 48
    xyz = (rng.normal(size=(n, 3)) * 0.8).astype(np.float32)
    rot = rng.normal(size=(n, 4)).astype(np.float32)  # un-normalised on purpose (w, x, y, z)
    scale = np.exp(rng.normal(size=(n, 3)) * 0.7 - 3.0).astype(np.float32)
    opac = (1 / (1 + np.exp(-rng.normal(size=n) * 2))).astype(np.float32)
    sh = (rng.normal(size=(n, 27)) * 0.3).astype(np.float32)
    sh[:, :3] = ((rng.random((n, 3)) - 0.5) / 0.28209479).astype(np.float32)
    for nm, pos, (w, h) in [("rand_cam005", (0.0, 0.0, 5.0), (800, 600)),
                            ("rand_demo_cam", (-0.57651054, 2.99040512, -0.03924271), (1280, 720)),
                            ("rand_oblique", (2.5, -1.5, 3.0), (1920, 1080))]:
        # scipy normalises the quaternion internally, like UnitQuaternion::from_quaternion
        cases.append(run_case(nm, xyz, rot, scale, opac, sh, pos, w, h))
    doc = {"generator": "tools/make_golden_from_notebook.py",
           "source": "notes/00_Gaussian_Projection.ipynb cell 1 (class Gaussian), notes/util.py, notes/util_gau.py @ 0d856a6",
           "adaptations": ["prototype fed the conjugate quaternion (its cov3D = R^T S R vs Rust R S R^T)",
                           "prototype fed 27 SH coefficients (Rust sh_dim=15 stops at degree 2); colour is clipped to [0,1] by the prototype",
                           "rot is (w,x,y,z) as in the PLY / util_gau; the Rust structs store (x,y,z,w)"],
           "cases": cases}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(doc, f)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
