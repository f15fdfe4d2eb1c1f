"""The C++ host side (include/splat_pipeline.hpp) on the real library and a real B200: the viewer loop of
examples/cpp_host/splat_demo -- load_from_ply, orbit, fill(0), render_to_buffer -- bit-exact against the
oracle for both pipelines and for the fused clear.  Runs last (the file name): it is the only GPU test that
goes through a second host language."""
import numpy as np
import pytest

from test_cpp_host import oracle_frame, raw_scene, read_frames, read_list, run

from splat_b200.gaussians import save_ply

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which,cleared,n,W,H", [(2, 0, 6000, 320, 240), (2, 1, 6000, 320, 240), (1, 0, 1500, 200, 150)])
def test_cpp_viewer_loop_matches_the_oracle(cpp_demo, tmp_path, orc, which, cleared, n, W, H):
    frames, step = 3, 0.35
    ply, scene_dump, out = tmp_path / "s.ply", tmp_path / "s.bin", tmp_path / "frames.bin"
    save_ply(str(ply), raw_scene(n))
    run(cpp_demo, "ply", ply, scene_dump)
    scene = read_list(scene_dump)                    # what the C++ loader produced: the oracle renders the same floats
    p = run(cpp_demo, "render", ply, H, W, 0.0, 0.0, 3.0, frames, step, which, cleared, out)
    got = read_frames(out, W, H)
    assert len(got) == frames
    for k, (cs, fb) in enumerate(got):
        ref = oracle_frame(orc, scene, cs, W, H, 0.01 if which == 1 else 0.3)
        assert np.count_nonzero(ref) > 1000
        bad = int(np.count_nonzero(fb != ref))
        assert bad == 0, f"frame {k}: {bad} of {W * H} pixels differ from the oracle\n{p.stderr}"
