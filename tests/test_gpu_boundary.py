"""GPU parity through the plug-in boundary: NumPy in -> registry / get_tensor_quant_params ->
C ABI -> NumPy out, compared with fixtures generated from the unmodified reference
(tests/golden/minmax.npz) and with the oracle.  Bit-exact."""
import os

import numpy as np
import pytest

from oracle import aeq_oracle as O
from tests import synthetic_graph as sg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cfg(bits, sym, gk):
  from aeq_b200 import qtyping
  G = qtyping.QuantGranularity
  gran = {0: G.CHANNELWISE, -1: G.TENSORWISE, 32: G.BLOCKWISE_32, 64: G.BLOCKWISE_64,
          128: G.BLOCKWISE_128, 256: G.BLOCKWISE_256}[gk]
  return qtyping.TensorQuantizationConfig(bits, sym, gran)


def test_minmax_fixtures_through_get_tensor_quant_params(cuda):
  from aeq_b200.algorithms.uniform_quantize import naive_min_max_quantize as nmm
  g = np.load(os.path.join(GOLD, "minmax.npz"))
  for key in g["cases"]:
    key = str(key)
    w = g[key.split("_")[0]]
    bits = int(key.split("_b")[1].split("_")[0])
    sym = key.split("_s")[1][0] == "1"
    gk = int(key.rsplit("_g", 1)[1])
    cfg = _cfg(bits, sym, gk)
    op, _ = sg.fc_graph(w)
    w_ro = w.copy()
    w_ro.setflags(write=False)  # inputs are read-only mmap views in the reference
    r = nmm.get_tensor_quant_params(sg.op_info(op, cfg), cfg, w_ro, None)
    np.testing.assert_array_equal(r.scale, g[key + "_scale"], err_msg=key)
    np.testing.assert_array_equal(r.zero_point, g[key + "_zp"], err_msg=key)
    np.testing.assert_array_equal(r.quantized_data, g[key + "_q"], err_msg=key)
    assert r.scale.dtype == np.float32 and r.scale.shape == g[key + "_scale"].shape
    assert r.zero_point.dtype == g[key + "_zp"].dtype and r.quantized_data.dtype == g[key + "_q"].dtype
    assert r.num_bits == bits and r.symmetric == sym and r.block_size == max(gk, 0)
    assert r.quantized_dimension == (0 if gk == 0 else (1 if gk > 0 else None))


def test_materialize_through_registry_uses_cache(cuda):
  from aeq_b200 import algorithm_manager as am
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.utils import common_utils
  w = O.synthetic_weight(32, 256, 1)
  op, graph = sg.fc_graph(w)
  cfg = _cfg(8, True, 0)
  info = sg.op_info(op, cfg, compute_precision=qtyping.ComputePrecision.INTEGER)  # DRQ
  fn = am.get_quantization_func(am.AlgorithmName.MIN_MAX_UNIFORM_QUANT, qtyping.TFLOperationName.FULLY_CONNECTED,
                                qtyping.QuantizeMode.MATERIALIZE)
  cache = common_utils.TensorQuantParamsCache()
  out = fn(op_info=info, graph_info=graph, tensor_name_to_qsv={}, tensor_quant_params_cache=cache)
  by_name = {t.tensor_name: t for t in out}
  wp = by_name["weight"].consumers[0]
  assert wp.transformations == [qtyping.QuantTransformation.QUANTIZE_TENSOR]
  ref = O.minmax_requant(w, 8, True)
  np.testing.assert_array_equal(wp.parameters.quantized_data, ref["q"])
  np.testing.assert_array_equal(wp.parameters.scale, ref["scale"])
  assert by_name["input"].consumers[0].transformations == [qtyping.QuantTransformation.NO_QUANTIZE]
  assert by_name["output"].producer.transformations == [qtyping.QuantTransformation.NO_QUANTIZE]
  assert len(cache) == 1
  again = fn(op_info=info, graph_info=graph, tensor_name_to_qsv={}, tensor_quant_params_cache=cache)
  assert again[1].consumers[0].parameters is wp.parameters  # cache hit: same object
  # weight-only => ADD_DEQUANTIZE
  info2 = sg.op_info(op, cfg, explicit_dequantize=True)
  out2 = fn(op_info=info2, graph_info=graph, tensor_name_to_qsv={}, tensor_quant_params_cache=cache)
  assert out2[1].consumers[0].transformations == [qtyping.QuantTransformation.ADD_DEQUANTIZE]


def test_qsv_min_max_path(cuda):
  """Activation tensors: scale / zp from calibrated min/max only (naive_min_max_quantize.py:52-75)."""
  from aeq_b200.algorithms.uniform_quantize import naive_min_max_quantize as nmm
  cfg = _cfg(8, False, -1)
  op, _ = sg.fc_graph(np.zeros((4, 8), np.float32))
  qsv = {"min": np.array([[-3.0]], np.float32), "max": np.array([[16.0]], np.float32)}
  r = nmm.get_tensor_quant_params(sg.op_info(op, cfg), cfg, None, qsv)
  zp, sc = O.scale_zp(qsv["min"], qsv["max"], 8, False, False)
  np.testing.assert_array_equal(r.scale, sc)
  np.testing.assert_array_equal(r.zero_point, zp)
  assert r.quantized_data is None and r.zero_point.dtype == np.int8
  # weights with a QSV: parameters from the QSV, data quantised with them
  w = O.synthetic_weight(8, 64, 2)
  cfgw = _cfg(8, True, 0)
  qsvw = {"min": w.min(axis=1, keepdims=True) * 0.5, "max": w.max(axis=1, keepdims=True) * 0.5}
  opw, _ = sg.fc_graph(w)
  r = nmm.get_tensor_quant_params(sg.op_info(opw, cfgw), cfgw, w, qsvw)
  zp, sc = O.scale_zp(qsvw["min"], qsvw["max"], 8, True, False)
  np.testing.assert_array_equal(r.scale, sc)
  np.testing.assert_array_equal(r.quantized_data, O.quantize(w, sc, zp, 8, True))


def test_uniform_quantize_tensor_literals(cuda):
  """uniform_quantize_tensor_test.py:120-169, :228-266, :303-325, :459-515 on the device."""
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import uniform_quantize_tensor as uqt
  x = np.array([-3.0, 1.3, 2.4, 16.0], np.float32)
  for bits, sym, scale, zp, want in [(8, True, 0.12598425, 0, [-24, 10, 19, 127]),
                                     (4, False, 1.2666667, -6, [-8, -5, -4, 7]),
                                     (4, True, 1.2666667, -6, [-8, -5, -4, 7])]:
    p = qtyping.UniformQuantParams(bits, None, np.array([scale], np.float32), np.array([zp], np.int8), sym)
    assert uqt.uniform_quantize(x, p).tolist() == want
  p = qtyping.UniformQuantParams(8, None, np.array([0.12598425], np.float32), np.array([0], np.int8), True)
  assert uqt.uniform_quantize(np.array([-16.0, 1.3, 2.4, 16.0], np.float32), p).tolist() == [-127, 10, 19, 127]
  d = uqt.uniform_dequantize(np.array([-8, -5, -4, 7], np.int8),
                             qtyping.UniformQuantParams(4, 0, np.array([1.2666667]), np.array([-6]), False))
  np.testing.assert_allclose(d, [-2.5333335, 1.2666668, 2.5333335, 16.466667], atol=1e-4)
  d = uqt.uniform_dequantize(
      np.array([[-8, -5, -4, 7], [-4, 7, -8, -5]], np.int8),
      qtyping.UniformQuantParams(4, 0, np.full((1, 2, 2), 1.2666667), np.array([[0]]), True, block_size=2)
      if False else qtyping.UniformQuantParams(4, 1, np.full((2, 2), 1.2666667, np.float32),
                                               np.zeros((2, 2), np.int8), True, block_size=2))
  np.testing.assert_allclose(d, [[-10.1333336, -6.3333335, -5.0666668, 8.8666669],
                                 [-5.0666668, 8.8666669, -10.1333336, -6.3333335]], atol=1e-4)
  zp, sc = uqt.tensor_zp_scale_from_min_max(np.array([[-3.0]]), np.array([[16.0]]), 8, False,
                                            qtyping.QuantGranularity.TENSORWISE)
  ozp, osc = O.scale_zp(np.array([[-3.0]], np.float32), np.array([[16.0]], np.float32), 8, False, False)
  np.testing.assert_array_equal(sc, osc)
  np.testing.assert_array_equal(zp, ozp)
  _, sc = uqt.tensor_zp_scale_from_min_max(np.array([[-3.0]]), np.array([[16.0]]), 8, True,
                                           qtyping.QuantGranularity.TENSORWISE, np.array([[4.0]], np.float32))
  assert sc[0, 0] == np.float32(4.0) / np.float32(127.0)


def test_error_wrapping_names_the_tensor(cuda):
  from aeq_b200 import algorithm_manager as am
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.utils import common_utils
  w = O.synthetic_weight(8, 48, 1)  # 48 % 32 != 0
  op, graph = sg.fc_graph(w)
  info = sg.op_info(op, _cfg(4, True, 32), compute_precision=qtyping.ComputePrecision.INTEGER)
  fn = am.get_quantization_func("min_max_uniform_quantize", qtyping.TFLOperationName.FULLY_CONNECTED,
                                qtyping.QuantizeMode.MATERIALIZE)
  with pytest.raises(ValueError, match=r"Failed to get quantization parameters for tensor: weight\. Error: .*not"
                                        r" divisible by block size 32"):
    fn(op_info=info, graph_info=graph, tensor_name_to_qsv={},
       tensor_quant_params_cache=common_utils.TensorQuantParamsCache())
