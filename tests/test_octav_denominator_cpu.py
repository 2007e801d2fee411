"""The OCTAV kernels' fp32 shortcut for the update's denominator (csrc/octav.cu, octav_update /
make_const), checked exhaustively on the CPU.

The reference computes `(1 - s) * count + s * N` with `count` and `(1 - s)` in fp32 and `s * N` in
float64 (N is an np.int64 scalar, octav.py:105-106), then the quotient in fp32, i.e.
den = f32(f64(f32(count * (1 - s))) + f64(s) * N).  The kernels replace the float64 round trip by
ONE fp32 add whenever `s * N` is an fp32 value and the two terms are within 29 binary orders (then
the float64 sum of two fp32 values is exact, and rounding it once to fp32 is the fp32 sum).  This
test replays both expressions for every count on every shape class the shortcut is enabled for.
"""
import numpy as np
import pytest


def _const(bits, divisor, n):
  s = np.float32(4.0 ** (-bits) / divisor)
  one_m_s = np.float32(1.0) - s
  s_n = np.float64(s) * np.float64(n)
  f = np.float32(s_n)
  exact = (np.float64(f) == s_n) and 1 <= bits <= 8 and 1.0 <= divisor <= 64.0 and 1 <= n <= (1 << 24)
  return s, one_m_s, s_n, (f if exact else None)


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("n", [32, 64, 128, 256, 1024, 4096, 16384, 1 << 20])
def test_fp32_add_equals_float64_round_trip(bits, n):
  s, one_m_s, s_n, f = _const(bits, 3.0, n)
  assert f is not None, "powers of two keep s * N an fp32 value"
  counts = np.arange(0, n + 1, dtype=np.float32) if n <= 16384 else np.unique(
      np.concatenate([np.arange(0, 4097), np.random.default_rng(0).integers(0, n + 1, 200000), [n]])).astype(np.float32)
  den0 = counts * one_m_s                                   # fp32 product, as in the kernel
  want = (den0.astype(np.float64) + s_n).astype(np.float32)  # the reference's dtype flow
  got = den0 + f                                            # the kernel's single FADD
  np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("n", [11008, 5120, 14336, 2560, 3, 1000])
def test_shortcut_is_off_when_s_n_is_not_fp32(n):
  """Row lengths that are not powers of two: s * N generally needs more than 24 bits, the kernel
  keeps the float64 path (s_n_f32 = NaN)."""
  for bits in (4, 8):
    s, _, s_n, f = _const(bits, 3.0, n)
    if f is not None:  # exact by accident: then the shortcut must still agree
      counts = np.arange(0, n + 1, dtype=np.float32)
      den0 = counts * (np.float32(1.0) - s)
      np.testing.assert_array_equal((den0 + f).view(np.uint32),
                                    (den0.astype(np.float64) + s_n).astype(np.float32).view(np.uint32))
    else:
      assert np.float64(np.float32(s_n)) != s_n
