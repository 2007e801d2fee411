"""Pins oracle/aeq_oracle.py: (1) the reference's own literal test vectors
(file:line cited per test), (2) tests/golden/*.npz produced from
the unmodified reference by tests/golden/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest

from oracle import aeq_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
  return np.load(os.path.join(GOLD, name + ".npz"))


def _gk(key):
  return int(key.rsplit("_g", 1)[1])


# ---------------------------------------------------------------- literal vectors
# uniform_quantize_tensor_test.py:120-169
@pytest.mark.parametrize("bits,sym,scale,zp,expect", [
    (8, False, 0.12598425, [-127], None),
    (8, True, 0.12598425, [0], [-24, 10, 19, 127]),
    (4, False, 1.2666667, [-6], [-8, -5, -4, 7]),
    (4, True, 1.2666667, [-6], [-8, -5, -4, 7]),
])
def test_quantize_literals(bits, sym, scale, zp, expect):
  x = np.array([-3.0, 1.3, 2.4, 16.0], np.float32)
  q = O.quantize(x, np.array([scale], np.float32), np.array(zp, np.int8), bits, sym)
  if expect is not None:
    assert q.tolist() == expect
  assert q.dtype == np.int8


def test_quantize_narrow_range_literal():
  # uniform_quantize_tensor_test.py:138-144: sym 8-bit clips at -127
  x = np.array([-16.0, 1.3, 2.4, 16.0], np.float32)
  q = O.quantize(x, np.array([0.12598425], np.float32), np.array([0], np.int8), 8, True)
  assert q.tolist() == [-127, 10, 19, 127]


def test_ranges():
  # uniform_quantize_tensor_test.py:39-50
  assert O.qrange(8) == (-128.0, 127.0)
  assert O.qrange(4) == (-8.0, 7.0)
  assert O.qrange(2) == (-2.0, 1.0)


def test_dequantize_literals():
  # uniform_quantize_tensor_test.py:228-266
  d = O.dequantize(np.array([-24, 10, 19, 127]), np.array([0.12598425]), np.array([0]))
  np.testing.assert_allclose(d, [-3.023622, 1.2598425, 2.3937008, 16.0], atol=1e-4)
  d = O.dequantize(np.array([-8, -5, -4, 7]), np.array([1.2666667]), np.array([-6]))
  np.testing.assert_allclose(d, [-2.5333335, 1.2666668, 2.5333335, 16.466667], atol=1e-4)


def test_dequantize_blockwise_literal():
  # uniform_quantize_tensor_test.py:303-325
  q = np.array([[-8, -5, -4, 7], [-4, 7, -8, -5]])
  d = O.dequantize(q, np.full((2, 2), 1.2666667), np.zeros((2, 2), np.int64), block=2)
  np.testing.assert_allclose(
      d, [[-10.1333336, -6.3333335, -5.0666668, 8.8666669],
          [-5.0666668, 8.8666669, -10.1333336, -6.3333335]], atol=1e-4)


def test_scale_zp_literals():
  # uniform_quantize_tensor_test.py:459-515 (8-bit asym / sym, with clipping)
  mn, mx = np.array([[-3.0]], np.float32), np.array([[16.0]], np.float32)
  zp, sc = O.scale_zp(mn, mx, 8, False, False)
  np.testing.assert_allclose(sc, [[19.0 / 255.0]], rtol=1e-6)
  assert zp.tolist() == [[-88]]
  zp, sc = O.scale_zp(mn, mx, 8, True, False)
  np.testing.assert_allclose(sc, [[16.0 / 127.0]], rtol=1e-6)
  assert zp.tolist() == [[0]]
  clip = np.array([[4.0]], np.float32)
  _, sc = O.scale_zp(mn, mx, 8, True, False, clip)
  assert sc[0, 0] == np.float32(4.0) / np.float32(127.0)


def test_blockwise_minmax_literal():
  # naive_min_max_quantize_test.py:162-205: seed 666 uniform(-10, 10) [4, 32]
  rng = np.random.default_rng(666)
  w = rng.uniform(-10, 10, (4, 32)).astype(np.float32)
  r = O.minmax_requant(w, 4, True, block=32)
  absmax = np.abs(w).max(axis=1, keepdims=True)
  want = (absmax / 7).astype(np.float32)
  import ml_dtypes
  want = want.astype(ml_dtypes.bfloat16).astype(np.float16).astype(np.float32)
  np.testing.assert_allclose(r["scale"], want, atol=1e-5)
  assert r["scale"].shape == (4, 1) and not r["zero_point"].any()


def test_calibration_filter_literal():
  # naive_min_max_quantize_test.py:207-240 / common_quantize_test.py:100-118
  x = np.array([[-np.inf, -3.39e38, -1.0, 2.0, 3.39e38, np.inf]], np.float32)
  mn, mx = O.activation_minmax(x)
  assert mn.shape == (1, 1) and mn.item() == -1.0 and mx.item() == 2.0
  x = np.array([3.2e38, 3.3e38], np.float32)  # nothing passes the max filter
  mn, mx = O.activation_minmax(x)
  assert mx.item() == np.float32(3.3e38) and mn.item() == np.float32(3.2e38)


def test_ema_literals():
  # qsv_utils_test.py:25-66
  q = O.ema_update({}, {"min": np.array([[-1.0]]), "max": np.array([[1.0]])})
  assert q["min"][0, 0] == -1.0
  q = O.ema_update({"min": np.array([[-10.0]]), "max": np.array([[10.0]])},
                   {"min": np.array([[-20.0]]), "max": np.array([[20.0]])})
  np.testing.assert_allclose(q["min"], [[-10.5]])
  np.testing.assert_allclose(q["max"], [[10.5]])
  u = O.minmax_union({"min": np.array([-1.0]), "max": np.array([1.0])},
                     {"min": np.array([-0.5]), "max": np.array([2.0])})
  assert u["min"][0] == -1.0 and u["max"][0] == 2.0


def test_gptq_hessian_merge_literal():
  # qsv_utils_test.py:111-180: sample-weighted running mean
  a = {"min": np.array([0.0]), "max": np.array([1.0]), "hessian": np.eye(2) * 2.0, "num_samples": 2}
  b = {"min": np.array([0.0]), "max": np.array([1.0]), "hessian": np.eye(2) * 8.0, "num_samples": 6}
  m = O.gptq_update(a, b)
  np.testing.assert_allclose(m["hessian"], np.eye(2) * 6.5)
  assert m["num_samples"] == 8


def test_pack_literals():
  # quantize_tensor_test.py:298-341 (INT4, odd length 15) and :258-296 (INT2)
  v = np.arange(15, dtype=np.int8)
  assert O.pack_bits(4, v).tolist() == [0x10, 0x32, 0x54, 0x76, 0x98, 0xBA, 0xDC, 0x0E]
  v = np.array([0, 1, 2, 3, 3, 2, 1, 0, 2, 3], np.int8)
  assert O.pack_bits(2, v).tolist() == [0b11100100, 0b00011011, 0b00001110]
  v5 = np.arange(5, dtype=np.int8)
  assert O.pack_bits(5, v5).tolist() == v5.tolist()  # :343-383 other widths untouched


def test_hadamard_literals():
  # hadamard_rotation_test.py:274-351
  r = O.hadamard_requant(np.ones((6, 6), np.float32), 8)
  assert r["hadamard_size"] == 2
  np.testing.assert_array_equal(r["q"], np.tile([127, 0], (6, 3)))
  w = np.tile(np.array([[1, 2], [3, 4]], np.float32), (3, 3))
  r = O.hadamard_requant(w, 8)
  np.testing.assert_array_equal(r["q"], np.tile([[127, -42], [127, -18]], (3, 3)))


def test_hadamard_sizes():
  # hadamard_rotation_test.py:223-272
  assert O.hadamard_size(32) == 32
  assert O.hadamard_size(32, 16) == 16
  assert O.hadamard_size(11008) == 256
  assert O.hadamard_size(96, 100) == 32


def test_gptq_literals():
  # gptq_test.py:214-256 and :258-299: 3x3 Hessian, int8 per-channel
  w = np.array([[1.0, 2.0, 0.55], [-1.0, 0.1, -0.55]], np.float32)
  h = np.array([[2.0, 0.5, 0.1], [0.5, 2.0, 0.3], [0.1, 0.3, 2.0]], np.float32)
  r = O.gptq_requant(w, h, 8)
  assert r["q"].shape == (2, 3) and r["q"].dtype == np.int8
  assert np.abs(r["q"].astype(int) - O.minmax_requant(w, 8)["q"].astype(int)).max() <= 2


def test_block_divisibility_error():
  # uniform_quantize_tensor_test.py:171-226 message text
  with pytest.raises(ValueError, match="is not divisible by block size 32"):
    O.weight_minmax(np.zeros((4, 48), np.float32), block=32)


# ---------------------------------------------------------------- reference fixtures
def test_minmax_fixtures_bit_exact():
  g = gold("minmax")
  for key in g["cases"]:
    key = str(key)
    w = g[key.split("_")[0]]
    bits = int(key.split("_b")[1].split("_")[0])
    sym = key.split("_s")[1][0] == "1"
    gk = _gk(key)
    with np.errstate(all="ignore"):
      r = O.minmax_requant(w, bits, sym, block=max(gk, 0), per_channel=(gk == 0))
    np.testing.assert_array_equal(r["scale"], g[key + "_scale"], err_msg=key)
    np.testing.assert_array_equal(r["zero_point"], g[key + "_zp"], err_msg=key)
    np.testing.assert_array_equal(r["q"], g[key + "_q"], err_msg=key)
    assert r["q"].dtype == g[key + "_q"].dtype and r["zero_point"].dtype == g[key + "_zp"].dtype


def test_octav_fixtures_bit_exact():
  g = gold("octav")
  for key in g["cases"]:
    key = str(key)
    w = g[key.split("_")[0]]
    bits = int(key.split("_b")[1].split("_")[0])
    gk = _gk(key)
    with np.errstate(all="ignore"):
      r = O.octav_requant(w, bits, block=max(gk, 0), per_channel=(gk == 0))
    np.testing.assert_array_equal(r["clip"].reshape(-1), g[key + "_clip"].reshape(-1), err_msg=key)
    np.testing.assert_array_equal(r["scale"], g[key + "_scale"], err_msg=key)
    np.testing.assert_array_equal(r["q"], g[key + "_q"], err_msg=key)


def test_mse_fixtures_bit_exact():
  g = gold("mse")
  for key in g["cases"]:
    key = str(key)
    w = g[key.split("_")[0]]
    bits = int(key.split("_b")[1].split("_")[0])
    r = O.mse_requant(w, bits, per_channel=(_gk(key) == 0))
    np.testing.assert_array_equal(r["scale"], g[key + "_scale"], err_msg=key)
    np.testing.assert_array_equal(r["q"], g[key + "_q"], err_msg=key)
    assert r["zero_point"].dtype == np.int32


def test_hadamard_fixtures():
  """sgemm summation order is BLAS-build dependent: rotation within 1e-6 of the row
  scale, integers may differ by one step at rounding boundaries."""
  g = gold("hadamard")
  for key in g["cases"]:
    key = str(key)
    wi = key.split("_")[0]
    w = g[wi]
    cap = int(g[wi + "_cap"])
    bits = int(key.split("_b")[1])
    r = O.hadamard_requant(w, bits, None if cap < 0 else cap)
    assert r["hadamard_size"] == int(g[key + "_hsize"])
    ref_rot = g[wi + "_rot"]
    tol = 1e-6 * np.abs(ref_rot).max()
    np.testing.assert_allclose(r["rotated"], ref_rot, atol=tol, rtol=0)
    np.testing.assert_allclose(r["scale"], g[key + "_scale"], rtol=2e-6)
    dq = np.abs(r["q"].astype(int) - g[key + "_q"].astype(int))
    assert dq.max() <= 1 and (dq != 0).mean() < 2e-3, key


def test_gptq_fixtures():
  """LAPACK/BLAS dependent: Hessian inverse to 1e-4 relative, integers may move by a
  step where the propagated error crosses a rounding boundary."""
  g = gold("gptq")
  for i in range(3):
    x = g[f"x{i}"]
    np.testing.assert_allclose(O.gptq_hessian(x), g[f"h{i}"], rtol=1e-5, atol=1e-5)
    hinv = O.gptq_hessian_inverse(g[f"h{i}"])
    np.testing.assert_allclose(hinv, g[f"hinv{i}"], rtol=1e-4, atol=1e-7)
    h = g[f"h{i}"].copy()
    O.gptq_hessian_inverse(h, mutate=True)  # damping stays in the caller's array
    np.testing.assert_array_equal(np.diag(h), g[f"hdiag_after{i}"])
  for key in g["cases"]:
    key = str(key)
    i = int(key[1])
    bits = int(key.split("_b")[1].split("_")[0])
    gk = _gk(key)
    r = O.gptq_requant(g[f"w{i}"], g[f"h{i}"], bits, True, block=gk, per_channel=True)
    np.testing.assert_array_equal(r["scale"], g[key + "_scale"], err_msg=key)
    dq = np.abs(r["q"].astype(int) - g[key + "_q"].astype(int))
    assert dq.max() <= 1 and (dq != 0).mean() < 5e-3, key


def test_calibration_fixtures_bit_exact():
  g = gold("calibration")
  mins, maxs = [], []
  for j in range(6):
    mn, mx = O.activation_minmax(g[f"a{j}"])
    np.testing.assert_array_equal(mn, g[f"a{j}_min"])
    np.testing.assert_array_equal(mx, g[f"a{j}_max"])
    mins.append(mn)
    maxs.append(mx)
  emn, emx = O.ema_sequence(mins, maxs)
  np.testing.assert_array_equal(emn, g["ema_min"])
  np.testing.assert_array_equal(emx, g["ema_max"])


def test_pack_fixtures_bit_exact():
  g = gold("pack")
  for bits, n in ((4, 15), (4, 4096), (2, 13), (2, 1024)):
    np.testing.assert_array_equal(O.pack_bits(bits, g[f"b{bits}_n{n}_in"]), g[f"b{bits}_n{n}_out"])


def test_recovery_fixtures_bit_exact():
  """dequantized_weight_recovery + float_casting fixtures (tests/golden/recovery.npz)."""
  z = np.load(os.path.join(GOLD, "recovery.npz"))
  for i, (bits, gk) in enumerate(z["cases"]):
    o = O.dwr_requant(z[f"w{i}"], int(bits), block=max(int(gk), 0), per_channel=(gk == 0))
    np.testing.assert_array_equal(o["scale"], z[f"scale{i}"])
    assert o["scale"].dtype == z[f"scale{i}"].dtype
    np.testing.assert_array_equal(o["q"], z[f"q{i}"])
  with np.errstate(all="ignore"):
    np.testing.assert_array_equal(O.float_cast(z["cast_in"]).view(np.uint16), z["cast_out"].view(np.uint16))


def test_oscar_fixtures_bit_exact():
  """OSCAR fixtures (tests/golden/oscar.npz): the oracle restates the reference's float64
  NumPy expressions, so everything is bit-identical."""
  z = np.load(os.path.join(GOLD, "oscar.npz"))
  for i, (bits, gk, with_mu2) in enumerate(z["cases"]):
    mu2 = z[f"mu2_{i}"] if with_mu2 else None
    with np.errstate(all="ignore"):
      o = O.oscar_requant(z[f"w{i}"], mu2, int(bits), block=max(int(gk), 0), per_channel=(gk == 0))
    np.testing.assert_array_equal(o["scale"], z[f"scale{i}"])
    np.testing.assert_array_equal(o["q"], z[f"q{i}"])
    np.testing.assert_array_equal(o["multiplier"], z[f"mult{i}"])
  np.testing.assert_array_equal(O.oscar_mu2(z["act"]), z["act_mu2"])
