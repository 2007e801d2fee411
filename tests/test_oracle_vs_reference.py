"""Cross-checks oracle/aeq_oracle.py against the UNMODIFIED reference on fresh seeds.

Only runs where /root/reference is mounted (the build container); on the GPU box
these tests skip and parity rests on tests/golden/*.npz (generated here)."""
import numpy as np
import pytest

from oracle import aeq_oracle as O
from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def R():
  import types
  ns = types.SimpleNamespace()
  ns.q = refshim.ref("qtyping")
  ns.nmm = refshim.ref("algorithms.uniform_quantize.naive_min_max_quantize")
  ns.octav = refshim.ref("algorithms.uniform_quantize.octav")
  ns.mse = refshim.ref("algorithms.uniform_quantize.mse")
  ns.had = refshim.ref("algorithms.uniform_quantize.hadamard_rotation")
  ns.gptq = refshim.ref("algorithms.uniform_quantize.gptq")
  ns.uqt = refshim.ref("algorithms.uniform_quantize.uniform_quantize_tensor")
  ns.cq = refshim.ref("algorithms.uniform_quantize.common_quantize")
  ns.qsv = refshim.ref("utils.qsv_utils")
  ns.tu = refshim.ref("transformations.transformation_utils")
  G = ns.q.QuantGranularity
  ns.gran = {0: G.CHANNELWISE, -1: G.TENSORWISE, 32: G.BLOCKWISE_32, 64: G.BLOCKWISE_64,
             128: G.BLOCKWISE_128, 256: G.BLOCKWISE_256}
  return ns


def _run(R, mod, w, bits, sym, gk, qsv=None, **params):
  cfg = R.q.TensorQuantizationConfig(num_bits=bits, symmetric=sym, granularity=R.gran[gk],
                                     algorithm_params=params)
  with np.errstate(all="ignore"):
    return mod.get_tensor_quant_params(refshim.fc_op_info(cfg), cfg, w, qsv)


@pytest.mark.parametrize("shape", [(32, 64), (9, 2048), (3, 16384)])
@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("gk", [0, -1, 32, 256])
def test_minmax(R, shape, bits, sym, gk):
  if gk > 0 and (not sym or shape[1] % gk):
    pytest.skip("not a valid blockwise config")
  w = O.synthetic_weight(*shape, index=bits * 7 + gk % 5)
  w[0, :] = 0
  r = _run(R, R.nmm, w, bits, sym, gk)
  with np.errstate(all="ignore"):
    o = O.minmax_requant(w, bits, sym, block=max(gk, 0), per_channel=(gk == 0))
  np.testing.assert_array_equal(o["scale"], r.scale)
  np.testing.assert_array_equal(o["zero_point"], r.zero_point)
  np.testing.assert_array_equal(o["q"], r.quantized_data)


def test_minmax_over_32mib_chunked_path(R):
  """> 32 MiB triggers the reference's row-chunk loop (uqt:323-354)."""
  w = O.synthetic_weight(2304, 4096, 99)
  r = _run(R, R.nmm, w, 8, True, 0)
  o = O.minmax_requant(w, 8, True)
  np.testing.assert_array_equal(o["q"], r.quantized_data)


@pytest.mark.parametrize("shape", [(16, 512), (5, 4096)])
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("gk", [0, -1, 32])
def test_octav(R, shape, bits, gk):
  w = O.synthetic_weight(*shape, index=3 + bits)
  r = _run(R, R.octav, w, bits, True, gk)
  with np.errstate(all="ignore"):
    o = O.octav_requant(w, bits, block=max(gk, 0), per_channel=(gk == 0))
  np.testing.assert_array_equal(o["scale"], r.scale)
  np.testing.assert_array_equal(o["q"], r.quantized_data)


@pytest.mark.parametrize("bits", [4, 8])
def test_mse(R, bits):
  w = O.synthetic_weight(12, 1024, 5)
  r = _run(R, R.mse, w, bits, True, 0)
  o = O.mse_requant(w, bits)
  np.testing.assert_array_equal(o["scale"], r.scale)
  np.testing.assert_array_equal(o["q"], r.quantized_data)
  assert r.zero_point.dtype == o["zero_point"].dtype


@pytest.mark.parametrize("cols,cap", [(512, None), (1024, 128), (2752, None)])
def test_hadamard(R, cols, cap):
  w = O.synthetic_weight(6, cols, 8)
  params = {} if cap is None else {"max_hadamard_size": cap}
  r = _run(R, R.had, w, 4, True, 0, **params)
  o = O.hadamard_requant(w, 4, cap)
  assert o["hadamard_size"] == r.hadamard.hadamard_size
  np.testing.assert_array_equal(o["scale"], r.scale)  # same BLAS in-process
  np.testing.assert_array_equal(o["q"], r.quantized_data)


@pytest.mark.parametrize("gk", [0, 32])
def test_gptq(R, gk):
  w = O.synthetic_weight(10, 160, 9)
  x = O.synthetic_activation((3, 70, 160), 9)
  h = O.gptq_hessian(x)
  qsv = {"activation_tensor_qsv": {"hessian": h.copy(), "num_samples": 3}}
  r = _run(R, R.gptq, w, 4, True, gk, qsv)
  o = O.gptq_requant(w, h, 4, True, block=gk)
  np.testing.assert_array_equal(o["scale"], r.scale)
  np.testing.assert_array_equal(o["q"], r.quantized_data)


def test_activation_minmax_and_ema(R):
  q_ref, q_or = {}, {}
  for j in range(5):
    a = O.synthetic_activation((2, 17, 33), 100 + j)
    if j == 2:
      a[0, 0, 0] = -np.inf
    ref = R.cq.get_activation_min_max(a, -3e38, 3e38)
    mn, mx = O.activation_minmax(a)
    np.testing.assert_array_equal(mn, ref["min"])
    np.testing.assert_array_equal(mx, ref["max"])
    q_ref = R.qsv.moving_average_update(q_ref, ref)
    q_or = O.ema_update(q_or, {"min": mn, "max": mx})
  np.testing.assert_array_equal(q_or["min"], q_ref["min"])
  np.testing.assert_array_equal(q_or["max"], q_ref["max"])


@pytest.mark.parametrize("bits,n", [(4, 31), (2, 30), (4, 64)])
def test_pack(R, bits, n):
  half = 2 ** (bits - 1)
  v = np.random.default_rng(n).integers(-half, half, n, dtype=np.int8)
  np.testing.assert_array_equal(O.pack_bits(bits, v), R.tu.pack_data(bits, v.view(np.uint8).copy()))


# ---- §8(f) row 3: dequantized_weight_recovery / float_casting -------------------------------
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("gk", [0, -1, 32, 128])
def test_dequantized_weight_recovery(R, bits, gk):
  dwr = refshim.ref("algorithms.uniform_quantize.dequantized_weight_recovery")
  w = O.fake_quantized_weight(24, 256, bits, block=max(gk, 0), index=bits + gk % 7, per_channel=(gk == 0))
  w[1, :] = 0.0          # all equal -> min_scale
  w[2, :] = w[2, 0]      # a single non-zero value: its distance to the appended 0
  r = _run(R, dwr, w, bits, True, gk)
  o = O.dwr_requant(w, bits, block=max(gk, 0), per_channel=(gk == 0))
  np.testing.assert_array_equal(o["scale"], r.scale)
  assert o["scale"].dtype == r.scale.dtype and o["scale"].shape == r.scale.shape
  np.testing.assert_array_equal(o["zero_point"], r.zero_point)
  np.testing.assert_array_equal(o["q"], r.quantized_data)


def test_dequantized_weight_recovery_literal(R):
  """dequantized_weight_recovery_test.py: the per-tensor / per-channel literal vectors."""
  dwr = refshim.ref("algorithms.uniform_quantize.dequantized_weight_recovery")
  deq = np.array([[-0.5, 0.25, 1.0], [0.75, -1.25, 0.5]], dtype=np.float32)
  zp, sc = dwr.get_zp_scale_from_dequantized_symmetric_weights(deq, None)
  np.testing.assert_array_equal(O.dwr_requant(deq, 8, per_channel=False)["scale"], sc)
  zp, sc = dwr.get_zp_scale_from_dequantized_symmetric_weights(deq, 0)
  np.testing.assert_array_equal(O.dwr_group_scales(deq).reshape(2, 1), sc)


# ---- §8(f) row 3: OSCAR ------------------------------------------------------------------
def _mu2(d, seed):
  rng = np.random.default_rng(seed)
  mu2 = rng.standard_normal(d) ** 2 * 0.5 + 0.01
  mu2[::17] *= 40.0   # a few hot input channels: this is what makes scaling pay
  mu2[3] = 0.0        # a dead channel (floored)
  return mu2


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("gk", [0, -1, 32, 128])
@pytest.mark.parametrize("with_mu2", [True, False])
def test_oscar(R, bits, gk, with_mu2):
  osc = refshim.ref("algorithms.uniform_quantize.oscar")
  w = O.synthetic_weight(24, 256, bits + gk % 5)
  w[1, :] = 0.0
  mu2 = _mu2(256, bits) if with_mu2 else None
  r = _run(R, osc, w, bits, True, gk, qsv=({"mu2": mu2} if with_mu2 else None))
  with np.errstate(all="ignore"):
    o = O.oscar_requant(w, mu2, bits, block=max(gk, 0), per_channel=(gk == 0))
  np.testing.assert_array_equal(o["scale"], r.scale)
  assert o["scale"].dtype == r.scale.dtype and o["scale"].shape == r.scale.shape
  np.testing.assert_array_equal(o["zero_point"], r.zero_point)
  assert o["zero_point"].dtype == r.zero_point.dtype
  np.testing.assert_array_equal(o["q"], r.quantized_data)
  np.testing.assert_array_equal(o["multiplier"], r.custom_algorithm_param["multiplier"])
  if with_mu2:
    assert not np.all(o["channel_scale"] == 1.0), "the fixture must exercise channel scaling"


def test_oscar_mu2_and_merge(R):
  osc = refshim.ref("algorithms.uniform_quantize.oscar")
  x = O.synthetic_activation((3, 50, 64), 4)
  np.testing.assert_array_equal(O.oscar_mu2(x), np.mean(np.asarray(x, np.float64).reshape(-1, 64) ** 2, axis=0))
  np.testing.assert_array_equal(O.oscar_floor(_mu2(64, 1)), osc._floor_positive(_mu2(64, 1)))
