#!/usr/bin/env python
"""Generate the IMAGE fixtures that pin the oracle's rasteriser semantics (SURVEY 8a row a-7,
8c E1) to artefacts the reference tree itself holds.  Runs only in the build container where
/root/reference is mounted; what it writes under tests/golden/ is committed, so the tests never
need the reference.

Two fixture files:

tests/golden/reference_images.npz -- stored images, decoded, nothing rendered by us:
  cell3_png      the stored output of notebook cell 3 (the 4-Gaussian scene drawn by the
                 prototype's plot_opacity into a 2560x1440 bitmap, camera Camera(720, 1280))
  cell6_gray     the stored output of cell 6 (the plush scene, prototype), image area only,
                 grey, resized to 320x180
  shot_gray      notes/screenshot.png (the Rust / euc window, same scene, same camera:
                 02_ply_demo.rs:22 == cell 6), window content only, grey, resized to 320x180

tests/golden/prototype_images.npz -- images produced by EXECUTING the prototype's own
  plot_opacity / plot_model (notebook cells 3 and 4, exec'd from where they lie, with the same
  stand-in modules as tools/make_golden_from_notebook.py) on
  (a) naive_gaussian() at 1280x720, camera (0,0,3)   [the scene of cell 3 at the camera's own size]
  (b) 160 seeded random Gaussians at 320x180, camera (0,0,3)
  together with the inputs, so that tests/test_reference_images.py can feed the same scene to
  the oracle.

Adaptations (same as make_golden_from_notebook.py, stated in the npz `notes` entry): the
prototype is fed the conjugate quaternion (its cov3D = R^T S R vs Rust's R S R^T) and 27 SH
coefficients (Rust stops at degree 2).
"""
import ast
import base64
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_from_notebook import REF, _install_standins  # noqa: E402

GOLD = os.path.normpath(os.path.join(HERE, "..", "tests", "golden"))


def dark_box(rgb, frac=0.3, thr=60):
    """bounding box (y0, y1, x0, x1) of the rows / columns that are mostly dark: the imshow area
    of a matplotlib figure on a white page, or the window content of the screenshot"""
    dark = rgb[..., :3].astype(np.int32).sum(axis=2) < thr
    rows = np.where(dark.mean(axis=1) > frac)[0]
    cols = np.where(dark[rows.min():rows.max() + 1].mean(axis=0) > frac)[0]
    return int(rows.min()), int(rows.max()) + 1, int(cols.min()), int(cols.max()) + 1


def stored_images():
    from PIL import Image

    nb = json.load(open(os.path.join(REF, "notes", "00_Gaussian_Projection.ipynb")))

    def cell_png(i):
        for o in nb["cells"][i]["outputs"]:
            if "data" in o and "image/png" in o["data"]:
                return Image.open(io.BytesIO(base64.b64decode(o["data"]["image/png"]))).convert("RGB")
        raise KeyError(i)

    c3 = np.asarray(cell_png(3))
    c6 = cell_png(6)
    a6 = np.asarray(c6)
    y0, y1, x0, x1 = dark_box(a6)
    c6g = np.asarray(c6.crop((x0, y0, x1, y1)).convert("L").resize((320, 180), Image.BILINEAR))
    shot = Image.open(os.path.join(REF, "notes", "screenshot.png"))
    a = np.asarray(shot.convert("RGBA")).copy()
    a[a[..., 3] < 255, :3] = 255          # the transparent margin / drop shadow around the window is not content
    shot = shot.convert("RGB")
    sy0, sy1, sx0, sx1 = dark_box(a)
    shg = np.asarray(shot.crop((sx0, sy0, sx1, sy1)).convert("L").resize((320, 180), Image.BILINEAR))
    out = os.path.join(GOLD, "reference_images.npz")
    np.savez_compressed(out, cell3_png=c3, cell6_gray=c6g, shot_gray=shg,
                        cell6_box=np.array([y0, y1, x0, x1]), shot_box=np.array([sy0, sy1, sx0, sx1]),
                        notes=np.array("decoded from notes/00_Gaussian_Projection.ipynb (stored outputs of cells 3, 6) and "
                                       "notes/screenshot.png @ 0d856a6 by tools/make_golden_images.py"))
    print("wrote", out, os.path.getsize(out), "bytes; cell6 box", (y0, y1, x0, x1), "shot box", (sy0, sy1, sx0, sx1))


def prototype_images():
    import scipy  # noqa
    import scipy.spatial.transform  # noqa

    _install_standins()
    sys.path.insert(0, os.path.join(REF, "notes"))
    import util  # noqa: reference module, unmodified
    import util_gau  # noqa: reference module, unmodified

    nb = json.load(open(os.path.join(REF, "notes", "00_Gaussian_Projection.ipynb")))
    ns = {"np": np, "sp": scipy, "util": util, "Camera": util.Camera, "tqdm": lambda x: x,
          "naive_gaussian": util_gau.naive_gaussian, "GaussianData": util_gau.GaussianData}
    exec(compile("".join(nb["cells"][1]["source"]), "notebook_cell_1", "exec"), ns)   # class Gaussian
    # cells 3 and 4 also draw with matplotlib at cell level: take only the two function definitions
    for ci, fn in ((3, "plot_opacity"), (4, "plot_model")):
        tree = ast.parse("".join(nb["cells"][ci]["source"]))
        node = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == fn][0]
        exec(compile(ast.Module([node], []), f"notebook_cell_{ci}", "exec"), ns)
    NBGaussian, plot_model = ns["Gaussian"], ns["plot_model"]

    def run(xyz, rot_wxyz, scale, opacity, sh27, cam_pos, w, h):
        objs = []
        for i in range(len(xyz)):
            q = rot_wxyz[i].astype(np.float64)
            objs.append(NBGaussian(xyz[i], scale[i], np.array([q[0], -q[1], -q[2], -q[3]]), opacity[i:i + 1], sh27[i]))
        cam = util.Camera(h, w, position=cam_pos)
        ns["h"], ns["w"] = h, w              # plot_model reads the notebook's globals h, w
        ns["opacity"] = opacity[-1:]         # ... and plot_opacity the stale loop variable `opacity` (cell 3)
        bitmap = plot_model(cam, objs)
        hf = cam.get_htanfovxy_focal()
        return bitmap, {"view": np.asarray(cam.get_view_matrix(), np.float64),
                        "proj": np.asarray(cam.get_projection_matrix(), np.float64),
                        "hf": np.array(hf, np.float64), "cam_pos": np.array(cam_pos, np.float64)}

    out = {}
    nv = util_gau.naive_gaussian()
    sh27 = np.zeros((4, 27), np.float32)
    sh27[:, :3] = nv.sh
    img, cam = run(nv.xyz, nv.rot, nv.scale, nv.opacity[:, 0], sh27, (0.0, 0.0, 3.0), 1280, 720)
    out.update(naive_img=img, naive_xyz=nv.xyz, naive_rot_wxyz=nv.rot, naive_scale=nv.scale,
               naive_opacity=nv.opacity[:, 0], naive_sh27=sh27, naive_wh=np.array([1280, 720]),
               **{"naive_" + k: v for k, v in cam.items()})
    rng = np.random.default_rng(20261018)
    n = 160
    xyz = (rng.normal(size=(n, 3)) * np.array([1.6, 0.9, 0.6])).astype(np.float32)
    rot = rng.normal(size=(n, 4)).astype(np.float32)
    scale = np.exp(rng.normal(size=(n, 3)) * 0.6 - 2.2).astype(np.float32)
    opac = (1 / (1 + np.exp(-rng.normal(size=n) * 2))).astype(np.float32)
    sh = (rng.normal(size=(n, 27)) * 0.25).astype(np.float32)
    sh[:, :3] = ((rng.random((n, 3)) - 0.5) / 0.28209479).astype(np.float32)
    img, cam = run(xyz, rot, scale, opac, sh, (0.0, 0.0, 3.0), 320, 180)
    out.update(rand_img=img, rand_xyz=xyz, rand_rot_wxyz=rot, rand_scale=scale, rand_opacity=opac, rand_sh27=sh,
               rand_wh=np.array([320, 180]), **{"rand_" + k: v for k, v in cam.items()})
    out["notes"] = np.array("produced by executing plot_model / plot_opacity of notes/00_Gaussian_Projection.ipynb (cells 3, 4) "
                            "@ 0d856a6 with tools/make_golden_images.py; prototype fed the conjugate quaternion and 27 SH "
                            "coefficients; rot is (w,x,y,z)")
    path = os.path.join(GOLD, "prototype_images.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    stored_images()
    prototype_images()
