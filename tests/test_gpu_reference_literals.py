"""The reference's own literal test vectors, restated against the CUDA path.

Each test cites the reference test it restates; inputs are cast to float32 (the only dtype the
reference ever materialises, common_utils.py:914-922) — verified in SURVEY.md Appendix A to give
the same expected integers.
"""
import numpy as np
import pytest

from tests import synthetic_graph as sg

pytestmark = pytest.mark.gpu


def _cfg(bits, gran, sym=True, **params):
  from aeq_b200 import qtyping
  return qtyping.TensorQuantizationConfig(num_bits=bits, symmetric=sym,
                                          granularity=getattr(qtyping.QuantGranularity, gran),
                                          algorithm_params=params)


def _info(w, cfg):
  op, _ = sg.fc_graph(w.reshape(w.shape[0], -1))
  return sg.op_info(op, cfg)


def test_pack_literals(cuda):
  """quantize_tensor_test.py:258-296 (INT2), :298-341 (INT4, odd length 15 -> last byte 0x0E)."""
  import torch
  from aeq_b200 import device
  v = torch.arange(15, dtype=torch.int8, device=cuda)
  assert device.pack_bits(v, 4).cpu().tolist() == [0x10, 0x32, 0x54, 0x76, 0x98, 0xBA, 0xDC, 0x0E]
  v = torch.tensor([0, 1, 2, 3, 3, 2, 1, 0, 2, 3], dtype=torch.int8, device=cuda)
  assert device.pack_bits(v, 2).cpu().tolist() == [0b11100100, 0b00011011, 0b00001110]


def test_dequantize_literals(cuda):
  """uniform_quantize_tensor_test.py:228-266 and the blockwise case :303-325."""
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import uniform_quantize_tensor as uqt
  p = qtyping.UniformQuantParams(8, None, np.array([0.12598425], np.float32), np.array([0], np.int8))
  d = uqt.uniform_dequantize(np.array([-24, 10, 19, 127], np.int8), p)
  np.testing.assert_allclose(d, [-3.023622, 1.2598425, 2.3937008, 16.0], atol=1e-6)
  p = qtyping.UniformQuantParams(4, None, np.array([1.2666667], np.float32), np.array([-6], np.int8))
  d = uqt.uniform_dequantize(np.array([-8, -5, -4, 7], np.int8), p)
  np.testing.assert_allclose(d, [-2.5333335, 1.2666668, 2.5333335, 16.466667], atol=1e-5)


def test_scale_zp_literals(cuda):
  """uniform_quantize_tensor_test.py:459-515 incl. clipping -> scale == clip / qmax exactly."""
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import uniform_quantize_tensor as uqt
  G = qtyping.QuantGranularity
  mn, mx = np.array([[-3.0]], np.float32), np.array([[16.0]], np.float32)
  zp, sc = uqt.tensor_zp_scale_from_min_max(mn, mx, 8, False, G.TENSORWISE)
  np.testing.assert_allclose(sc, [[19.0 / 255.0]], rtol=1e-6)
  assert zp.tolist() == [[-88]] and zp.dtype == np.int8
  zp, sc = uqt.tensor_zp_scale_from_min_max(mn, mx, 8, True, G.TENSORWISE)
  assert sc[0, 0] == np.float32(16.0) / np.float32(127.0) and zp.tolist() == [[0]]
  _, sc = uqt.tensor_zp_scale_from_min_max(mn, mx, 8, True, G.TENSORWISE, np.array([[4.0]], np.float32))
  assert sc[0, 0] == np.float32(4.0) / np.float32(127.0)


def test_blockwise_minmax_seed_666(cuda):
  """naive_min_max_quantize_test.py:162-205: uniform(-10, 10) [4, 32], scale = bf16 -> fp16 of absmax / 7."""
  import ml_dtypes
  from aeq_b200.algorithms.uniform_quantize import naive_min_max_quantize as nmm
  w = np.random.default_rng(666).uniform(-10, 10, (4, 32)).astype(np.float32)
  cfg = _cfg(4, "BLOCKWISE_32")
  r = nmm.get_tensor_quant_params(_info(w, cfg), cfg, w, None)
  want = (np.abs(w).max(axis=1, keepdims=True) / 7).astype(np.float32)
  want = want.astype(ml_dtypes.bfloat16).astype(np.float16).astype(np.float32)
  np.testing.assert_allclose(r.scale, want, atol=1e-5)
  np.testing.assert_array_equal(r.scale, want)
  assert r.scale.shape == (4, 1) and not r.zero_point.any() and r.block_size == 32
  assert r.quantized_dimension == 1 and r.quantized_data.dtype == np.int8


def test_calibration_filter_literals(cuda):
  """naive_min_max_quantize_test.py:207-240 and common_quantize_test.py:100-118."""
  from aeq_b200.algorithms.uniform_quantize import common_quantize as cq
  x = np.array([[-np.inf, -3.39e38, -1.0, 2.0, 3.39e38, np.inf]], np.float32)
  mm = cq.get_activation_min_max(x, -3e38, 3e38)
  assert mm["min"].shape == (1, 1) and mm["min"].item() == -1.0 and mm["max"].item() == 2.0
  raw = cq.get_activation_min_max(np.array([3.2e38, 3.3e38], np.float32), -3e38, 3e38)
  assert raw["max"].item() == np.float32(3.3e38) and raw["min"].item() == np.float32(3.2e38)


def test_hadamard_rank3_literal(cuda):
  """hadamard_rotation_test.py:318-351 (golden 3): rank-3 tensor, per-channel over dim 0."""
  from aeq_b200.algorithms.uniform_quantize import hadamard_rotation as had
  data = np.reshape(np.tile([[1, 2], [3, 4]], [3, 3]), (2, 3, 6)).astype(np.float32)
  want = np.reshape(np.tile([[54, -18], [127, -18]], [3, 3]), (2, 3, 6))
  cfg = _cfg(8, "CHANNELWISE")
  r = had.get_tensor_quant_params(_info(data, cfg), cfg, data, None)
  np.testing.assert_array_equal(r.quantized_data, want)
  assert r.hadamard.hadamard_size == 2 and r.quantized_data.shape == (2, 3, 6)


def test_hadamard_size_rules(cuda):
  """hadamard_rotation_test.py:223-272: 32 -> 32; capped at 16; Llama 11008 -> 256 tiled 43 x."""
  from aeq_b200.algorithms.uniform_quantize import hadamard_rotation as had
  assert had.hadamard_size_for(32) == 32 and had.hadamard_size_for(32, 16) == 16
  assert had.hadamard_size_for(11008) == 256 and had.hadamard_size_for(96, 100) == 32
  for bad_cfg, msg in ((dict(content=None), "only supported for weight tensors"),):
    with pytest.raises(ValueError, match=msg):
      had.get_tensor_quant_params(None, _cfg(8, "CHANNELWISE"), None, None)
  w = np.ones((4, 8), np.float32)
  with pytest.raises(ValueError, match="not supported for static quantization"):
    had.get_tensor_quant_params(_info(w, _cfg(8, "CHANNELWISE")), _cfg(8, "CHANNELWISE"), w, {"min": 0})
  with pytest.raises(ValueError, match="rank >= 2"):
    had.get_tensor_quant_params(None, _cfg(8, "CHANNELWISE"), np.ones(8, np.float32), None)


def test_gptq_blockwise_picks_the_column_scale(cuda):
  """gptq_test.py:331-398 (the reference patches the block size to 2 on a [2, 4] weight; the
  same property at the real block size 32): identity Hessian, per-block scales k/127 ->
  every value quantises to 10 when each column uses ITS block's scale."""
  import ml_dtypes
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import gptq
  w = (np.repeat(np.array([[10, 20], [30, 40]], np.float32), 32, axis=1) / 127).astype(np.float32)
  qsv = {"min": np.array([[-1.0, -2.0], [-3.0, -4.0]]), "max": np.array([[1.0, 2.0], [3.0, 4.0]]),
         "activation_tensor_qsv": {"hessian": np.eye(64, dtype=np.float32), "num_samples": 1}}
  cfg = _cfg(8, "BLOCKWISE_32")
  info = qtyping.OpInfo(op=sg.fc_graph(w)[0], op_name=qtyping.TFLOperationName.FULLY_CONNECTED,
                        subgraph_op_index=-1, op_quant_config=qtyping.OpQuantizationConfig())
  r = gptq.get_tensor_quant_params(info, cfg, w, qsv)
  want = np.array([[1 / 127, 2 / 127], [3 / 127, 4 / 127]], np.float32)
  want = want.astype(ml_dtypes.bfloat16).astype(np.float16).astype(np.float32)
  np.testing.assert_array_equal(r.scale, want)
  assert not r.zero_point.any() and r.quantized_data.dtype == np.int8
  np.testing.assert_array_equal(r.quantized_data, np.full_like(w, 10, dtype=np.int8))


def test_gptq_hessian_literal(cuda):
  """gptq_test.py:50-114: +-1e39-like values are filtered from min / max but NOT from the Hessian."""
  import types
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import gptq
  x = np.array([[[1.0, 2.0], [3.0, 4.0]], [[-3.2e38, 1.0], [0.5, 3.3e38]]], np.float32)
  op = types.SimpleNamespace(inputs=[0, 1, -1], outputs=[2])
  graph = qtyping.GraphInfo(
      subgraph_tensors=[sg.tensor("in", x.shape, 0), sg.tensor("w", (2, 2), 1), sg.tensor("out", (2, 2, 2), 0)],
      buffers=[types.SimpleNamespace(data=None), types.SimpleNamespace(data=np.zeros((2, 2), np.float32).tobytes())])
  q = gptq.calibrate(op, graph, {"in": x, "out": np.zeros((2, 2, 2), np.float32)})["in"]
  assert q["min"].item() == 0.5 and q["max"].item() == 4.0 and int(q["num_samples"]) == 2
  x2 = x.reshape(-1, 2)
  with np.errstate(all="ignore"):
    want = (2.0 / np.array(2)) * x2.T.dot(x2)
  np.testing.assert_allclose(q["hessian"], want, rtol=1e-6)
