"""GPU parity of dequantized_weight_recovery and float_casting against the
reference-generated fixtures (tests/golden/recovery.npz) and the oracle.

Bar: bit-exact.  The recovered scale is the minimum over exact fp32 differences of sorted
values (order-free), the re-quantisation is one IEEE divide + rint per element, the fp16 cast
is one RNE conversion.
"""
import os

import numpy as np
import pytest

from oracle import aeq_oracle as O
from tests import synthetic_graph as sg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "recovery.npz")


def _cfg(bits, gk):
  from aeq_b200 import qtyping
  G = qtyping.QuantGranularity
  gran = {0: G.CHANNELWISE, -1: G.TENSORWISE, 32: G.BLOCKWISE_32, 64: G.BLOCKWISE_64,
          128: G.BLOCKWISE_128, 256: G.BLOCKWISE_256}[gk]
  return qtyping.TensorQuantizationConfig(num_bits=bits, symmetric=True, granularity=gran)


def _run(w, cfg, **op_kw):
  from aeq_b200.algorithms.uniform_quantize import dequantized_weight_recovery as dwr
  op, _ = sg.fc_graph(w)
  return dwr.get_tensor_quant_params(sg.op_info(op, cfg, **op_kw), cfg, w, None)


def test_recovery_golden(cuda):
  z = np.load(GOLD)
  for i, (bits, gk) in enumerate(z["cases"]):
    r = _run(z[f"w{i}"], _cfg(int(bits), int(gk)))
    np.testing.assert_array_equal(r.scale, z[f"scale{i}"])
    assert r.scale.dtype == z[f"scale{i}"].dtype and r.scale.shape == z[f"scale{i}"].shape
    assert r.zero_point.dtype == np.int32 and not r.zero_point.any()
    np.testing.assert_array_equal(r.quantized_data, z[f"q{i}"])
    assert r.quantized_data.dtype == z[f"q{i}"].dtype


@pytest.mark.parametrize("shape,bits,gk", [
    ((64, 4096), 4, 0), ((64, 4096), 8, 32), ((3, 16384), 8, 0), ((5, 1000), 4, 0),
    ((7, 48), 8, 0), ((1, 1), 8, 0), ((128, 2048), 4, 64), ((16, 4096), 8, 256),
    ((256, 4096), 4, -1), ((3, 7), 8, -1)])
def test_recovery_vs_oracle(cuda, shape, bits, gk):
  """Ragged rows (+inf padded segments), one-value rows, every block size, the radix-sort path."""
  w = O.fake_quantized_weight(shape[0], shape[1], bits, block=max(gk, 0), index=shape[1] % 13,
                              per_channel=(gk == 0))
  if shape[0] > 2:
    w[1, :] = 0.0
    w[2, :] = w[2, 0]
  r = _run(w, _cfg(bits, gk), skip_checks=True)
  o = O.dwr_requant(w, bits, block=max(gk, 0), per_channel=(gk == 0))
  np.testing.assert_array_equal(r.scale, o["scale"])
  np.testing.assert_array_equal(r.quantized_data, o["q"])
  # The recovery check (reference :36-63) must agree with the oracle's verdict: a 32-wide block
  # of 8-bit values often has no two neighbours one step apart, and then the reference raises.
  rec = O.dequantize(o["q"], o["scale"].astype(np.float32), o["zero_point"], block=max(gk, 0))
  if np.abs(rec - w).max() > 1e-4:
    with pytest.raises(RuntimeError, match="Failed to recover weights"):
      _run(w, _cfg(bits, gk))
  else:
    np.testing.assert_array_equal(_run(w, _cfg(bits, gk)).quantized_data, o["q"])


def test_recovery_rejects_float_weights(cuda):
  """A non-QAT tensor fails the recovery check with the reference's diagnosis
  (dequantized_weight_recovery.py:264-303); skip_checks bypasses it."""
  w = O.synthetic_weight(8, 512, 3) * 50
  with pytest.raises(RuntimeError, match="exceeds the limit of 16 for 4-bit"):
    _run(w, _cfg(4, 0))
  r = _run(w, _cfg(4, 0), skip_checks=True)
  assert r.quantized_data.shape == w.shape
  from aeq_b200 import qtyping
  cfg = qtyping.TensorQuantizationConfig(num_bits=8, symmetric=False,
                                         granularity=qtyping.QuantGranularity.CHANNELWISE)
  with pytest.raises(ValueError, match="Only symmetric weights"):
    _run(w, cfg)


def test_max_abs_diff_and_dwr_scales_raw(cuda):
  import torch
  from aeq_b200 import device
  a = torch.from_numpy(O.synthetic_weight(300, 1000, 1)).to(cuda)
  b = a.clone()
  b.view(-1)[123457] += 0.5
  assert device.max_abs_diff(a, b).item() == pytest.approx(0.5, abs=1e-6)
  b.view(-1)[5] = float("nan")
  assert np.isnan(device.max_abs_diff(a, b).item())
  x = a.reshape(-1)[: 64 * 4096].reshape(64, 4096).contiguous()
  np.testing.assert_array_equal(device.dwr_scales(x, 64, 4096).cpu().numpy(),
                                O.dwr_group_scales(x.cpu().numpy()))


def test_float_casting(cuda):
  """float_casting: golden cast (overflow, RNE boundary, subnormal) + the materialiser's shape."""
  import torch
  from aeq_b200 import device, qtyping
  from aeq_b200.algorithms.nonlinear_quantize import float_casting as fc
  from aeq_b200.algorithms.utils import common_utils
  z = np.load(GOLD)
  got = device.cast_f16(torch.from_numpy(z["cast_in"]).to(cuda)).cpu().numpy()
  np.testing.assert_array_equal(got.view(np.uint16), z["cast_out"].view(np.uint16))
  w = O.synthetic_weight(33, 77, 9)  # odd count: scalar tail
  np.testing.assert_array_equal(
      device.cast_f16(torch.from_numpy(w).to(cuda)).cpu().numpy().view(np.uint16),
      O.float_cast(w).view(np.uint16))
  op, graph = sg.fc_graph(w, bias=True)
  cfg = qtyping.TensorQuantizationConfig(num_bits=16, dtype=qtyping.TensorDataType.FLOAT)
  info = sg.op_info(op, cfg)
  fc.check_op_quantization_config(info.op_name, info.op_quant_config)
  cache = common_utils.TensorQuantParamsCache()
  params = fc.materialize_weight_op(info, graph, {}, cache)
  by_name = {p.tensor_name: p for p in params}
  wp = by_name["weight"].consumers[0]
  assert wp.transformations == [qtyping.QuantTransformation.ADD_DEQUANTIZE]
  assert isinstance(wp.parameters, qtyping.NonLinearQuantParams) and wp.parameters.num_bits == 16
  np.testing.assert_array_equal(wp.parameters.quantized_data.view(np.uint16), O.float_cast(w).view(np.uint16))
  for name in ("input", "bias"):
    assert by_name[name].consumers[0].transformations == [qtyping.QuantTransformation.NO_QUANTIZE]
  assert by_name["output"].producer.transformations == [qtyping.QuantTransformation.NO_QUANTIZE]
  assert fc.materialize_weight_op(info, graph, {}, cache)[1].consumers[0].parameters is wp.parameters
  with pytest.raises(ValueError, match="number of bits to be set as 16"):
    bad = qtyping.OpQuantizationConfig(weight_tensor_config=qtyping.TensorQuantizationConfig(num_bits=8))
    fc.check_op_quantization_config(info.op_name, bad)
