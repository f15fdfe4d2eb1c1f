"""Exhaustive / closed-form checks of the arithmetic identities the CUDA blend kernel relies on
(splat_b200/csrc/blend.cuh) and of the pinned exp shared by oracle and kernel.  Pure host
arithmetic with exact rationals: no GPU, no oracle compute beyond orc_expf."""
import math
from fractions import Fraction

import numpy as np


def rn_f32(x: Fraction) -> float:
    """Round an exact rational to the nearest binary32 (ties to even), returned as a Python float."""
    if x == 0:
        return 0.0
    s = -1 if x < 0 else 1
    x = abs(x)
    e = math.floor(math.log2(float(x)))
    while Fraction(2) ** e > x:
        e -= 1
    while Fraction(2) ** (e + 1) <= x:
        e += 1
    e = max(e, -126)
    ulp = Fraction(2) ** (e - 23)
    q = x / ulp
    n = q.numerator // q.denominator
    rem = q - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n % 2 == 1):
        n += 1
    return s * float(n * ulp)


K1 = float.fromhex("0x1.010102p-8")     # RN(1/255)
K2 = float.fromhex("-0x1.fdfdfep-33")   # RN(1/255 - K1)


def test_div255_two_op_form_is_correctly_rounded_for_every_byte():
    """blend.cuh div255: fma(n, K1, n*K2) == RN(n/255) for n = 0..255 (what `byte as f32 / 255.0`
    gives in the reference, pipelines.rs:150-152)."""
    assert K1 == rn_f32(Fraction(1, 255))
    for n in range(256):
        t = rn_f32(Fraction(n) * Fraction(K2))
        got = rn_f32(Fraction(n) * Fraction(K1) + Fraction(t))
        want = rn_f32(Fraction(n, 255))
        assert got == want, n
        assert np.float32(n) / np.float32(255.0) == np.float32(want)


def test_byte_roundtrip_identity():
    """(x/255)*255 truncates back to x for all 256 bytes (SURVEY 8c known answer 2): a zero
    fragment leaves RGB unchanged."""
    for n in range(256):
        v = np.float32(np.float32(n) / np.float32(255.0)) * np.float32(255.0)
        assert int(v) == n


def test_magic_number_truncation():
    """trunc(v) for 0 <= v < 2^23 == RZ(v + 2^23) - 2^23 (blend.cuh blend_channel)."""
    rng = np.random.default_rng(0)
    vs = np.concatenate([rng.uniform(0, 255.0, 20000), np.arange(0, 256), np.arange(0, 256) - 2 ** -18,
                         np.arange(0, 256) + 2 ** -17]).astype(np.float32)
    vs = vs[vs >= 0]
    for v in vs[:3000]:
        exact = Fraction(float(v)) + 2 ** 23
        rz = Fraction(math.floor(exact))          # ulp(2^23) == 1: RZ keeps the integer part
        assert int(rz - 2 ** 23) == int(math.floor(float(v)))


def test_saturating_cast_equals_clamp_then_truncate():
    """`(out*255) as u8` (saturating, NaN -> 0) == trunc(sat01(out)*255) for finite out: the
    kernel saturates on the add (FADD.SAT) instead of on the cast."""
    rng = np.random.default_rng(1)
    outs = np.concatenate([rng.uniform(-2, 3, 50000), [0.0, 1.0, -0.0, 0.9999999, 1.0000001, 255.0 / 255.0]]).astype(np.float32)
    v = outs * np.float32(255.0)
    as_u8 = np.where(v >= 255, 255, np.where(v > 0, np.floor(np.minimum(v, 255)), 0)).astype(np.int64)
    sat = np.clip(outs, np.float32(0), np.float32(1)) * np.float32(255.0)
    assert np.array_equal(as_u8, np.floor(sat).astype(np.int64))


def test_pinned_exp_accuracy(orc):
    """splat_expf v1 stays within 1 ulp of the true exponential on its domain and flushes below
    -87; exp(0) == 1 exactly (alpha of a centred sample is exactly min(0.99, opacity))."""
    assert orc.expf(0.0) == 1.0
    assert orc.expf(-87.5) == 0.0 and orc.expf(-1e30) == 0.0
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(-87, 0, 20000), -np.logspace(-30, 1.9, 2000)]).astype(np.float32)
    worst = 0.0
    for x in xs:
        got = orc.expf(float(x))
        want = math.exp(float(x))
        ulp = math.ldexp(1.0, max(math.frexp(want)[1] - 1, -126) - 23)
        worst = max(worst, abs(got - want) / ulp)
    assert worst <= 1.0, worst
