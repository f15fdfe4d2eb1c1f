import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# tests/test_emu_parity.py: the emulated runs that take more than a few seconds each only run on request, so that the
# default CPU suite stays within a couple of minutes (SPLAT_EMU_FULL=1 runs all of them: about twelve minutes)
EMU_ON_REQUEST = (
    "test_frames_without_a_host_round_trip", "test_near_cut_stripes_and_empty_regions", "test_near_cut_is_exact[128-None]",
    "test_near_cut_is_exact[512-False]", "test_a_repeated_frame_is_counted_as_retried", "test_framebuffer_bit_exact[inside_cloud]",
    "test_framebuffer_bit_exact[demo_cam_100k]", "test_float_blend_matches_float_oracle[demo_cam_100k_720p]",
    "test_render_device_stripes_into_one_device_frame", "test_stripes_equal_full_frame",
    "test_deep_lists_exact_early_termination[deep_150k_128x96]", "test_deep_lists_exact_early_termination[deep_mixed_opacity]",
    "test_deep_lists_exact_early_termination[deep_onto_noise]", "test_bench_two_ranks_on_the_emulated_library[async]",
    "test_group_context_equals_single_device_and_oracle_emulated[equal_stripes]",
    "test_group_context_repeats_abandoned_member_stripes",
    # BASELINE's full sizes (281 k @720p complete; 6.1 M @1080p and 5.8 M @4K on sampled stripes): about six minutes
    "test_config2_plush_sized_complete_frame", "test_full_size_configs_sampled_stripes[config3_bicycle_sized_1080p]",
    "test_full_size_configs_sampled_stripes[config4_garden_sized_4k]",
    "test_cpp_viewer_loop_on_the_emulated_library[2-1-3000-320-240]", "test_cpp_viewer_loop_on_the_emulated_library[1-0-1500-200-150]",
)


def pytest_collection_modifyitems(config, items):
    if os.environ.get("SPLAT_EMU_FULL") == "1":
        return
    skip = pytest.mark.skip(reason="emulated run on request only: SPLAT_EMU_FULL=1")
    for item in items:
        if item.nodeid.startswith("tests/test_emu_parity.py::") and item.name in EMU_ON_REQUEST:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/liboracle.so), built on demand with gcc."""
    from oracle import oracle

    oracle.build()
    oracle.lib()
    return oracle


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cpp_demo():
    """examples/cpp_host/splat_demo, the driver of the C++ host side (include/splat_pipeline.hpp).  Rebuilt with
    g++ when stale (seconds); a binary that travelled with the snapshot is used as it is if that fails."""
    import subprocess

    d = os.path.join(ROOT, "examples", "cpp_host")
    exe = os.path.join(d, "splat_demo")
    if not os.path.exists(os.path.join(ROOT, "splat_b200", "libsplat_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "splat_b200", "csrc")])
    r = subprocess.run(["make", "-s", "-C", d], capture_output=True, text=True)
    if r.returncode != 0 and not os.path.exists(exe):
        pytest.fail("cannot build examples/cpp_host/splat_demo:\n" + r.stderr)
    return exe
