"""not gpu: the float32 double-float translation of the AlignTraj pass (csrc/qcp.cu df_translate2).

The reference moves float32 positions by a float64 vector (``positions += vector``,
core/groups.py:1480): the float64 sum rounded once to float32.  The kernel computes it without
leaving the FP32 pipe: Knuth's TwoSum of x + t_hi, then + t_lo.  This restates that sequence in
NumPy float32 arithmetic (IEEE round-to-nearest, like FADD2) and pins the claim made in the
kernel's comment: identical to the float64 result except about one coordinate in 10^7, and then
the neighbouring float32."""
import numpy as np


def df_translate(x, t):
    f = np.float32
    th = t.astype(f)
    tl = (t - th.astype(np.float64)).astype(f)
    s = x + th
    bb = s - x
    e = (x - (s - bb)) + (th - bb)
    return s + (e + tl)


def test_df_translate_is_the_rounded_float64_sum():
    rng = np.random.default_rng(1)
    n = 4_000_000
    worst = 0
    for scale, shift in ((50.0, 37.123456789), (5.0, 0.5), (1000.0, -812.3456), (0.01, 100.0), (30.0, -1e-3)):
        x = (rng.standard_normal(n) * scale).astype(np.float32)
        t = np.float64(shift) + rng.standard_normal(n) * 1e-3
        want = (x.astype(np.float64) + t).astype(np.float32)
        got = df_translate(x, t)
        bad = np.flatnonzero(got != want)
        assert len(bad) <= 8, (scale, shift, len(bad))              # ~2e-7 of the coordinates at most
        if len(bad):
            ulp = np.abs(np.spacing(want[bad])).astype(np.float64)
            assert np.all(np.abs(got[bad].astype(np.float64) - want[bad].astype(np.float64)) <= ulp)
        worst = max(worst, len(bad))
    # exact cases: zero shift, x = 0, huge cancellation
    x = np.array([0.0, 1.5, -2.25, 1e-30, 3.0e7], dtype=np.float32)
    assert np.array_equal(df_translate(x, np.zeros(5)), x)
    t = -x.astype(np.float64)
    assert np.array_equal(df_translate(x, t), np.zeros(5, dtype=np.float32))
