"""End-to-end `aeq_b200.Quantizer` on synthetic TFLite models (BASELINE.json configs[0]:
dynamic_wi8_afp32 on a tiny FULLY_CONNECTED model, plumbing + arithmetic): read -> recipe ->
device kernels -> QUANTIZE_TENSOR / ADD_DEQUANTIZE -> write -> read back, checked against the
oracle.  Bar: bit-exact integers, packed bytes and scales (min-max is order-free)."""
import numpy as np
import pytest

from oracle import aeq_oracle as O
from tests import tfl_fixtures

pytestmark = pytest.mark.gpu


def _weights():
  return [O.synthetic_weight(48, 64, 1), O.synthetic_weight(32, 48, 2), O.synthetic_weight(64, 32, 3)]


def _tensor(g, name):
  return next(t for t in g.tensors if t.name == name)


def test_dynamic_wi8_afp32_end_to_end(cuda, tmp_path):
  from aeq_b200 import quantizer, recipe
  from aeq_b200.utils import tfl_flatbuffer_utils as fu
  from aeq_b200.utils import tfl_model as T
  ws = _weights()
  bias = [np.linspace(-1, 1, w.shape[0]).astype(np.float32) for w in ws]
  path = tmp_path / "float.tflite"
  T.write_model(tfl_fixtures.fc_stack(ws, biases=bias), str(path))
  qz = quantizer.Quantizer(str(path), recipe.dynamic_wi8_afp32())
  assert not qz.need_calibration()
  result = qz.quantize()
  assert result.recipe == recipe.dynamic_wi8_afp32()
  result.save(str(tmp_path / "out"), "model")
  with pytest.raises(ValueError, match="already exists"):
    result.export_model(str(tmp_path / "out" / "model.tflite"))
  m = T.read_model(str(tmp_path / "out" / "model.tflite"))
  g = m.subgraphs[0]
  assert len(g.operators) == 3 and len(m.buffers) == 1 + 6
  for i, w in enumerate(ws):
    want = O.minmax_requant(w, 8, True)
    t = _tensor(g, b"layer%d/w" % i)
    assert t.type == T.TensorType.INT8 and list(t.shape) == list(w.shape)
    q = t.quantization
    np.testing.assert_array_equal(q.scale, want["scale"].ravel())
    assert q.scale.dtype == np.float32 and q.zeroPoint.dtype == np.int64 and not q.zeroPoint.any()
    assert q.quantizedDimension == 0 and q.details is None
    np.testing.assert_array_equal(fu.get_tensor_data(t, m.buffers), want["q"])
    b = _tensor(g, b"layer%d/b" % i)   # DRQ leaves the bias float
    assert b.type == T.TensorType.FLOAT32 and b.quantization is None
    np.testing.assert_array_equal(fu.get_tensor_data(b, m.buffers), bias[i])
  assert len(result.quantized_model) < path.stat().st_size * 0.45  # fp32 -> int8 weights


def test_dynamic_wi4_blockwise_packs_and_adds_scale_tensor(cuda):
  from aeq_b200 import quantizer, recipe
  from aeq_b200.utils import tfl_model as T
  ws = [O.synthetic_weight(64, 64, 1), O.synthetic_weight(32, 64, 2), O.synthetic_weight(96, 32, 3)]
  data = T.write_model_to_bytes(tfl_fixtures.fc_stack(ws))
  m = T.read_model_from_bytes(quantizer.Quantizer(data, recipe.dynamic_wi4b32_afp32()).quantize().quantized_model)
  g = m.subgraphs[0]
  for i, w in enumerate(ws):
    want = O.minmax_requant(w, 4, True, block=32)
    t = _tensor(g, b"layer%d/w" % i)
    assert t.type == T.TensorType.INT4
    np.testing.assert_array_equal(np.asarray(m.buffers[t.buffer].data), O.pack_bits(4, want["q"]))
    q = t.quantization
    assert q.detailsType == T.QuantizationDetails.BlockwiseQuantization and q.quantizedDimension == 0
    assert q.details.blockSize == 32 and q.details.zeroPoints == -1 and q.scale is None
    st = g.tensors[q.details.scales]
    assert st.name == b"layer%d/w_scales" % i and st.type == T.TensorType.FLOAT16
    assert list(st.shape) == [w.shape[0], w.shape[1] // 32]
    got = np.frombuffer(bytes(m.buffers[st.buffer].data), np.float16).reshape(st.shape)
    np.testing.assert_array_equal(got.view(np.uint16), O.blockwise_scale_fp16(want["scale"]).view(np.uint16))


def test_weight_only_inserts_dequantize_and_shares_buffers(cuda):
  from aeq_b200 import qtyping, quantizer
  from aeq_b200.utils import tfl_flatbuffer_utils as fu
  from aeq_b200.utils import tfl_model as T
  w = O.synthetic_weight(40, 40, 7)
  model = tfl_fixtures.fc_stack([w, O.synthetic_weight(40, 40, 8), w], share_first_weight_with_last=True,
                                embedding=O.synthetic_weight(100, 40, 9))
  qz = quantizer.Quantizer(T.write_model_to_bytes(model))
  qz.add_weight_only_config(".*", qtyping.TFLOperationName.FULLY_CONNECTED, 8)
  qz.update_quantization_recipe("layer1/", qtyping.TFLOperationName.FULLY_CONNECTED,
                                algorithm_key="no_quantize")
  m = T.read_model_from_bytes(qz.quantize().quantized_model)
  g = m.subgraphs[0]
  codes = [T.builtin_code(m.operatorCodes[o.opcodeIndex]) for o in g.operators]
  B = T.BuiltinOperator
  assert codes == [B.EMBEDDING_LOOKUP, B.DEQUANTIZE, B.FULLY_CONNECTED, B.FULLY_CONNECTED,
                   B.DEQUANTIZE, B.FULLY_CONNECTED]
  want = O.minmax_requant(w, 8, True)
  for name, deq_idx, fc_idx in ((b"layer0/w", 1, 2), (b"layer2/w_shared", 4, 5)):
    t = _tensor(g, name)
    assert t.type == T.TensorType.INT8
    np.testing.assert_array_equal(fu.get_tensor_data(t, m.buffers), want["q"])
    np.testing.assert_array_equal(t.quantization.scale, want["scale"].ravel())
    deq, fc = g.operators[deq_idx], g.operators[fc_idx]
    assert list(deq.inputs) == [next(i for i, x in enumerate(g.tensors) if x is t)]
    out = g.tensors[int(deq.outputs[0])]
    assert out.name == name + b"_dequant" and out.type == T.TensorType.FLOAT32 and out.buffer == 0
    assert int(fc.inputs[1]) == int(deq.outputs[0])
  assert _tensor(g, b"layer0/w").buffer == _tensor(g, b"layer2/w_shared").buffer  # packed once
  assert _tensor(g, b"layer1/w").type == T.TensorType.FLOAT32                      # carved out
  assert _tensor(g, b"embedding/table").type == T.TensorType.FLOAT32               # FC-only recipe


def test_dynamic_recipe_covers_embedding_and_regex_scopes(cuda):
  """`*` expands to every op of the algorithm: the EMBEDDING_LOOKUP table is quantised per row
  like an FC weight; a later scope entry overrides the blanket one for the layers it matches."""
  from aeq_b200 import qtyping, quantizer, recipe
  from aeq_b200.utils import tfl_flatbuffer_utils as fu
  from aeq_b200.utils import tfl_model as T
  table = O.synthetic_weight(100, 64, 5)
  ws = [O.synthetic_weight(32, 64, 6), O.synthetic_weight(64, 32, 7)]
  qz = quantizer.Quantizer(T.write_model_to_bytes(tfl_fixtures.fc_stack(ws, embedding=table)),
                           recipe.dynamic_wi8_afp32())
  qz.add_dynamic_config("layer1/", qtyping.TFLOperationName.FULLY_CONNECTED, 4)
  m = T.read_model_from_bytes(qz.quantize().quantized_model)
  g = m.subgraphs[0]
  t = _tensor(g, b"embedding/table")
  want = O.minmax_requant(table, 8, True)
  assert t.type == T.TensorType.INT8 and t.quantization.quantizedDimension == 0
  np.testing.assert_array_equal(fu.get_tensor_data(t, m.buffers), want["q"])
  np.testing.assert_array_equal(t.quantization.scale, want["scale"].ravel())
  assert _tensor(g, b"layer0/w").type == T.TensorType.INT8
  w1 = _tensor(g, b"layer1/w")
  assert w1.type == T.TensorType.INT4
  np.testing.assert_array_equal(np.asarray(m.buffers[w1.buffer].data),
                                O.pack_bits(4, O.minmax_requant(ws[1], 4, True)["q"]))
  assert len(qz.get_quantization_recipe()) == 2


def test_octav_int4_recipe_and_errors(cuda):
  from aeq_b200 import qtyping, quantizer
  from aeq_b200.utils import tfl_flatbuffer_utils as fu
  from aeq_b200.utils import tfl_model as T
  w = O.synthetic_weight(32, 256, 4)
  data = T.write_model_to_bytes(tfl_fixtures.fc_stack([w]))
  qz = quantizer.Quantizer(data)
  with pytest.raises(RuntimeError, match="without a quantization recipe"):
    qz.quantize()
  qz.add_dynamic_config(".*", qtyping.TFLOperationName.FULLY_CONNECTED, 4, algorithm_key="OCTAV")
  m = T.read_model_from_bytes(qz.quantize().quantized_model)
  t = _tensor(m.subgraphs[0], b"layer0/w")
  want = O.octav_requant(w, 4)
  np.testing.assert_allclose(t.quantization.scale, want["scale"].ravel(), rtol=1e-6)
  packed = np.asarray(m.buffers[t.buffer].data)
  assert t.type == T.TensorType.INT4 and packed.size == w.size // 2
  assert (packed != O.pack_bits(4, want["q"])).mean() <= 2e-3
  with pytest.raises(NotImplementedError, match="LiteRT interpreter"):
    qz.calibrate({})


def test_star_blockwise_recipe_leaves_conv_float(cuda):
  """A `*` blockwise recipe on a model that also holds a CONV_2D: the FC weights become packed
  INT4 with fp16 block scales, the convolution stays float32 (its config check rejects blockwise,
  so the recipe entry is skipped for that op: recipe_manager.py:185-198) — no KeyError in the
  blockwise dimension table."""
  from aeq_b200 import quantizer, recipe
  from aeq_b200.utils import tfl_flatbuffer_utils as fu
  from aeq_b200.utils import tfl_model as T
  ws = [O.synthetic_weight(64, 64, 1), O.synthetic_weight(32, 64, 2)]
  conv = O.synthetic_weight(8 * 9, 4, 3).reshape(8, 3, 3, 4)
  data = T.write_model_to_bytes(tfl_fixtures.fc_stack(ws, conv_front=conv))
  m = T.read_model_from_bytes(quantizer.Quantizer(data, recipe.dynamic_wi4b32_afp32()).quantize().quantized_model)
  g = m.subgraphs[0]
  c = _tensor(g, b"conv/w")
  assert c.type == T.TensorType.FLOAT32 and c.quantization is None
  np.testing.assert_array_equal(fu.get_tensor_data(c, m.buffers), conv)
  for i, w in enumerate(ws):
    t = _tensor(g, b"layer%d/w" % i)
    want = O.minmax_requant(w, 4, True, block=32)
    assert t.type == T.TensorType.INT4
    np.testing.assert_array_equal(np.asarray(m.buffers[t.buffer].data), O.pack_bits(4, want["q"]))
  # the same model under a per-channel `*` recipe quantises the convolution too (dim 0)
  m8 = T.read_model_from_bytes(quantizer.Quantizer(data, recipe.dynamic_wi8_afp32()).quantize().quantized_model)
  c8 = _tensor(m8.subgraphs[0], b"conv/w")
  want = O.minmax_requant(conv.reshape(8, -1), 8, True)
  assert c8.type == T.TensorType.INT8
  np.testing.assert_array_equal(fu.get_tensor_data(c8, m8.buffers).reshape(8, -1), want["q"])


def test_model_over_2gb_writes_external_buffers(cuda):
  """Serialisation switches to external buffers (offset / size after the flatbuffer, 16-byte
  aligned: model_modifier.py:290-377) when the payloads pass 2 GB; forced here on a small model
  so the test stays small, then read back and compared tensor by tensor."""
  from aeq_b200 import quantizer, recipe
  from aeq_b200.utils import tfl_flatbuffer_utils as fu
  from aeq_b200.utils import tfl_model as T
  ws = [O.synthetic_weight(64, 128, 4), O.synthetic_weight(32, 64, 5)]
  qz = quantizer.Quantizer(T.write_model_to_bytes(tfl_fixtures.fc_stack(ws), external_buffers=True),
                           recipe.dynamic_wi8_afp32())
  out = qz.quantize(external_buffers=True).quantized_model
  m = T.read_model_from_bytes(out)
  for i, w in enumerate(ws):
    t = _tensor(m.subgraphs[0], b"layer%d/w" % i)
    np.testing.assert_array_equal(fu.get_tensor_data(t, m.buffers), O.minmax_requant(w, 8, True)["q"])


def test_mse_weights_go_through_the_batched_driver(cuda):
  """SURVEY.md §8f row 1 beyond min-max: MSE per-channel weights with 128-multiple rows travel through
  the host pipeline in ONE batched call (aeqb_host_requant_mse_rows_batch_f32) and land in the cache;
  the result equals the per-op path (mse.get_tensor_quant_params) bit for bit; a weight the batched path
  cannot take (96 columns) is left to the per-op path."""
  import types
  from aeq_b200 import host, qtyping, quantizer
  from aeq_b200.algorithm_manager import AlgorithmName
  from aeq_b200.algorithms.uniform_quantize import mse
  from aeq_b200.utils import tfl_flatbuffer_utils as fu
  from aeq_b200.utils import tfl_model as T
  # 256-, 48- and 256-column rows: the middle one is not a multiple of 128
  ws = [O.synthetic_weight(48, 256, 1), O.synthetic_weight(256, 48, 2), O.synthetic_weight(128, 256, 3)]
  data = T.write_model_to_bytes(tfl_fixtures.fc_stack(ws))
  qz = quantizer.Quantizer(data)
  qz.add_weight_only_config(".*", qtyping.TFLOperationName.FULLY_CONNECTED, 8, algorithm_key=AlgorithmName.MSE)
  m = T.read_model_from_bytes(qz.quantize().quantized_model)
  assert qz.prefetch_stats["quantized"] == 2 and qz.prefetch_stats["left_to_per_op_path"] == 1
  g = m.subgraphs[0]
  cfg = qtyping.TensorQuantizationConfig(num_bits=8, symmetric=True, granularity=qtyping.QuantGranularity.CHANNELWISE)
  info = qtyping.OpInfo(types.SimpleNamespace(inputs=[0, 1, -1], outputs=[2]), qtyping.TFLOperationName.FULLY_CONNECTED,
                        0, qtyping.OpQuantizationConfig(weight_tensor_config=cfg))
  for i, w in enumerate(ws):
    want = mse.get_tensor_quant_params(info, cfg, w, None)
    t = _tensor(g, b"layer%d/w" % i)
    np.testing.assert_array_equal(t.quantization.scale, want.scale.ravel())
    np.testing.assert_array_equal(fu.get_tensor_data(t, m.buffers), want.quantized_data)
    np.testing.assert_allclose(want.scale, O.mse_requant(w, 8)["scale"], rtol=1e-6)
  with pytest.raises(RuntimeError, match="128"):
    host.requant_mse_rows([ws[1]], 8, 0.05408)
