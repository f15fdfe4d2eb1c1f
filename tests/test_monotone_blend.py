"""The property the exact early termination of the blend kernel rests on,
checked on the CPU against the oracle's own blend step (pipelines.rs:147-168):

  for a fixed fragment (alpha, colour) the map  old byte -> new byte  of one channel is monotone
  non-decreasing, hence so is any composition; therefore, once a run of fragments sends both
  byte 0 and byte 255 to the same byte, it sends EVERY byte there -- the pixel has forgotten
  what was underneath.
"""
import numpy as np

f32 = np.float32


def step(b, alpha, col):
    """One channel of blend() for an array of old bytes, in the reference's f32 order:
    trunc_sat(((1-a) * (b/255) + a*c) * 255)."""
    old = b.astype(f32) / f32(255.0)
    om = f32(1.0) - f32(alpha)
    out = (om * old).astype(f32) + (f32(alpha) * f32(col)).astype(f32)
    v = (out.astype(f32) * f32(255.0)).astype(f32)
    v = np.where(np.isnan(v), f32(0.0), v)
    return np.clip(np.trunc(v), 0, 255).astype(np.int64)


def test_numpy_step_is_the_oracles_blend(orc):
    rng = np.random.default_rng(11)
    for _ in range(300):
        alpha = float(f32(rng.uniform(1.0 / 255.0, 0.99)))
        col = [float(f32(c)) for c in rng.uniform(-0.5, 1.5, 3)]
        b = int(rng.integers(0, 256))
        old = b | (b << 8) | (b << 16)
        # dx = dy = 0: power = 0, exp = 1, alpha = min(0.99, opacity)
        px = orc.shade_blend(old, 1.0, 0.0, 1.0, 0.0, 0.0, alpha, col)
        want = [int(step(np.array([b]), alpha, c)[0]) for c in col]
        assert [(px >> 16) & 255, (px >> 8) & 255, px & 255] == want


def test_every_step_is_monotone_in_the_old_byte():
    rng = np.random.default_rng(12)
    b = np.arange(256)
    alphas = np.concatenate([[1.0 / 255.0, 0.99, 0.5], rng.uniform(1.0 / 255.0, 0.99, 4000)]).astype(f32)
    cols = np.concatenate([[0.0, 1.0, -0.3, 1.7, 0.5], rng.uniform(-0.5, 1.5, 4000)]).astype(f32)
    for a in alphas[:400]:
        for c in cols[:40]:
            y = step(b, a, c)
            assert np.all(np.diff(y) >= 0), (float(a), float(c))
    for a, c in zip(alphas, cols[: len(alphas)]):
        assert np.all(np.diff(step(b, a, c)) >= 0)


def test_merged_extremes_mean_every_start_byte_merged():
    rng = np.random.default_rng(13)
    merged_after = []
    for trial in range(200):
        n = 400
        # alphas as the renderer produces them: opacity * falloff, clamped, >= 1/255
        alpha = np.clip(rng.uniform(0.0, 1.0, n) ** 2 * rng.uniform(0.05, 1.0, n), 1.0 / 255.0, 0.99).astype(f32)
        col = rng.normal(0.5, 0.4, n).astype(f32)
        state = np.arange(256)
        k_merge = None
        for k in range(n):
            state = step(state, alpha[k], col[k])
            assert np.all(np.diff(state) >= 0)                # compositions stay monotone
            if state[0] == state[255]:
                assert np.all(state == state[0])              # ... so the sandwich closes everything
                k_merge = k + 1
                break
        assert k_merge is not None, "400 fragments did not make the pixel forget its start"
        merged_after.append(k_merge)
    # with these alphas a pixel forgets its start after a few dozen fragments (the kernel's first
    # suffix attempt is 192 list entries per tile)
    assert np.median(merged_after) < 100


def test_a_one_level_gap_can_persist_under_faint_fragments():
    """Why the kernel tracks both extremes instead of trusting a transmittance bound: with tiny
    alphas the truncation keeps two neighbouring states apart indefinitely."""
    lo, hi = np.array([100]), np.array([101])
    for _ in range(2000):
        lo, hi = step(lo, 1.0 / 255.0, 0.4), step(hi, 1.0 / 255.0, 0.4)
    # whether they have met or not, their order is preserved -- the scheme assumes nothing more
    assert lo[0] <= hi[0]
