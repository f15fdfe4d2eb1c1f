"""The drop-in boundary without a GPU: libsplat_b200.so loads, exports every symbol that
include/splat.h declares (and nothing is declared that the ctypes binding does not know),
struct layouts agree between the header, the ctypes mirror and the library, and the render
entry points fail loudly -- never fall back -- when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "splat.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(splat_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from splat_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib


def test_header_and_binding_list_the_same_symbols(lib):
    assert _declared_functions() == sorted(lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    L = lib.load()
    for name in _declared_functions():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (splat_[a-z0-9_]+)", out))
    assert set(_declared_functions()) <= exported


def test_abi_version_and_defaults(lib):
    L = lib.load()
    assert L.splat_abi_version() == int(re.search(r"SPLAT_ABI_VERSION\s+(\d+)u", open(HEADER).read()).group(1))
    cfg = lib.SplatConfig()
    L.splat_config_default(C.byref(cfg))
    assert cfg.device == 0 and cfg.tile == 16 and cfg.y_down == 0 and cfg.zclip_mode == 1
    assert np.float32(cfg.lowpass) == np.float32(0.3) and cfg.sample_offset == 0.5   # Pipeline02


def test_struct_layouts_match_a_c_compiler(lib, tmp_path):
    """sizeof/offsetof of the public structs as gcc sees the header == the ctypes mirror."""
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "splat.h"\n'
        "int main(void){\n"
        'printf("%zu %zu %zu %zu ", sizeof(splat_config), offsetof(splat_config, sample_offset), offsetof(splat_config, max_instances), offsetof(splat_config, blend_mode));\n'
        'printf("%zu %zu %zu ", sizeof(splat_camera), offsetof(splat_camera, position), offsetof(splat_camera, focal));\n'
        'printf("%zu %zu %zu\\n", sizeof(splat_timings), offsetof(splat_timings, n_gaussians), offsetof(splat_timings, kernel_launches));\n'
        "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(lib.SplatConfig), lib.SplatConfig.sample_offset.offset, lib.SplatConfig.max_instances.offset,
            lib.SplatConfig.blend_mode.offset,
            C.sizeof(lib.SplatCamera), lib.SplatCamera.position.offset, lib.SplatCamera.focal.offset,
            C.sizeof(lib.SplatTimings), lib.SplatTimings.n_gaussians.offset, lib.SplatTimings.kernel_launches.offset]
    assert got == want


def test_no_cpu_fallback_without_a_device(lib):
    """On a box without a GPU the context cannot be created: the product path raises."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(lib.SplatError) as e:
        lib.Context(device=0)
    assert e.value.code in (-2, -3)
    from splat_b200.gaussians import GaussianList
    from splat_b200.camera import Camera
    from splat_b200.pipelines import GaussianSplatPipeline02

    with pytest.raises(lib.SplatError):
        GaussianSplatPipeline02(GaussianList.naive_gaussians(), Camera(64, 64))


def test_null_arguments_return_error_codes(lib):
    L = lib.load()
    assert L.splat_create(None, None) == -1
    assert L.splat_get_timings(None, None) == -1
    assert L.splat_last_error(None) == b"null context"
    L.splat_destroy(None)   # no-op


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under splat_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "splat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|splat_oracle", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_docs_only_name_entry_points_the_header_declares():
    """INTEGRATION.md (the binding a maintainer would add), README.md and DESIGN.md must not drift from include/splat.h"""
    import re

    hdr = open(os.path.join(ROOT, "include", "splat.h")).read()
    declared = set(re.findall(r"\bsplat_[a-z0-9_]+\b", hdr))
    not_symbols = {"splat_", "splat_comm_", "splat_demo", "splat_expf", "splat_oracle", "splat_pipeline", "splat_sys", "splat_b200",
                   "splat_api", "splat_upload_"}
    for doc in ("INTEGRATION.md", "README.md", "DESIGN.md"):
        names = set(re.findall(r"\bsplat_[a-z0-9_]*", open(os.path.join(ROOT, doc)).read()))
        unknown = sorted(n for n in names - declared - not_symbols if not n.endswith("_"))
        assert not unknown, f"{doc} names {unknown}, which include/splat.h does not declare"
